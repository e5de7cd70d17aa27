#!/bin/bash
# r2s: waves — NLHE parity (all), trace, headline
O=gpurun_out
TAG=${1:-r2s}
timeout 900 python -m pytest tests/test_nlhe_gpu.py tests/test_pins.py -x -q -m gpu --timeout 600 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
RBP_NLHE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --epochs-per-step 8 --skip-cpu-baseline > /dev/null 2> $O/trace_${TAG}.txt; tail -n 3 $O/trace_${TAG}.txt | cut -c1-300
for W in 1 2 4; do
RBP_NLHE_WAVES=$W timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_w$W.json 2> $O/bench_${TAG}_w$W.err; tail -2 $O/bench_${TAG}_w$W.err
python - $O/bench_${TAG}_w$W.json $W <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("waves", sys.argv[2], "%.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "launches", d["gpu_launches"])
PY
done
