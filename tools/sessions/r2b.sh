#!/bin/bash
# r2b: merge + chain fold — NLHE parity tests, headline, trace
O=gpurun_out
TAG=${1:-r2b}
mkdir -p $O
timeout 900 python -m pytest tests/test_nlhe_gpu.py tests/test_pins.py -x -q -m gpu --timeout 300 > $O/pytest_${TAG}.log 2>&1; tail -3 $O/pytest_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}.err
timeout 400 python bench.py --steps 5 --warmup 3 --batch 65536 --epochs-per-step 8 --skip-cpu-baseline > $O/bench_${TAG}_nlhe64k_n1.json 2>> $O/bench_${TAG}.err
for f in nlhe_n1 nlhe64k_n1; do python - $O/bench_${TAG}_$f.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "%.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "launches", d["gpu_launches"])
PY
done
tail -n 5 $O/bench_${TAG}.err
