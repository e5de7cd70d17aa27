#!/bin/bash
# last session of the round: full GPU suite + smoke on the final code, NLHE lines at 16k / 65k
O=gpurun_out
TAG=${1:-r1x}
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 > $O/pytest_${TAG}.log 2>&1; tail -2 $O/pytest_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
timeout 200 python bench.py --workload nlhe --batch 65536 --steps 10 --skip-cpu-baseline > $O/bench_${TAG}_nlhe64k_n1.json 2> $O/bench_${TAG}.err
timeout 200 python bench.py --workload nlhe --steps 30 --skip-cpu-baseline > $O/bench_${TAG}_nlhe_n1.json 2>> $O/bench_${TAG}.err
for f in nlhe_n1 nlhe64k_n1; do python - $O/bench_${TAG}_$f.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "%.4g updates/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "launches", d["gpu_launches"])
PY
done
tail -n 3 $O/bench_${TAG}.err
