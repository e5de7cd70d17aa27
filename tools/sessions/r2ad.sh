#!/bin/bash
# r2ad: validation of the round's final code: full GPU suite, smoke, the driver's bench command (both arms), configs[4] at full N on the shipped Elkan step
O=gpurun_out
TAG=${1:-r2ad}
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_${TAG}.log 2>&1; tail -1 $O/smoke_${TAG}.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}_nlhe_n1.err; tail -1 $O/bench_${TAG}_nlhe_n1.err
python - $O/bench_${TAG}_nlhe_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("nlhe %.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"])
PY
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_${TAG}_reference_arm.json 2> $O/bench_${TAG}_reference_arm.err; cut -c1-300 $O/bench_${TAG}_reference_arm.json
for K in 100 500; do
timeout 400 python bench.py --workload lloyd_turn --k $K --steps 8 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_lloyd_turn_k$K.json 2> $O/bench_${TAG}_lloyd_turn_k$K.err; tail -1 $O/bench_${TAG}_lloyd_turn_k$K.err
python - $O/bench_${TAG}_lloyd_turn_k$K.json $K <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("lloyd_turn k", sys.argv[2], "%.3f ms/iter" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "reassigned", d["reassigned_last"])
PY
done
