#!/bin/bash
# NLHE value phase with large roots split into per-child tasks: parity, then bench lines.
O=gpurun_out
TAG=${1:-r1t}
mkdir -p $O
timeout 900 python -m pytest tests/test_nlhe_gpu.py -x -q --timeout 600 > $O/pytest_${TAG}.log 2>&1; tail -3 $O/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
timeout 400 python bench.py --workload nlhe --steps 30 > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}.err
timeout 400 python bench.py --workload nlhe --batch 65536 --steps 10 > $O/bench_${TAG}_nlhe64k_n1.json 2>> $O/bench_${TAG}.err
for f in nlhe_n1 nlhe64k_n1; do python - $O/bench_${TAG}_$f.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "%.4g updates/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
PY
done
tail -n 3 $O/bench_${TAG}.err
