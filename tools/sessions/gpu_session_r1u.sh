#!/bin/bash
# NLHE value phase, bucket-sorted child tasks: parity, then split-threshold sweep at 16k and 64k trees.
O=gpurun_out
TAG=${1:-r1u}
mkdir -p $O
timeout 900 python -m pytest tests/test_nlhe_gpu.py -x -q --timeout 600 > $O/pytest_${TAG}.log 2>&1; tail -3 $O/pytest_${TAG}.log
for sp in 192 96 384 1000000; do
  RBP_NLHE_SPLIT=$sp timeout 300 python bench.py --workload nlhe --steps 20 > $O/bench_${TAG}_nlhe_s$sp.json 2>> $O/bench_${TAG}.err
done
for sp in 192 512 1536; do
  RBP_NLHE_SPLIT=$sp timeout 300 python bench.py --workload nlhe --batch 65536 --steps 8 > $O/bench_${TAG}_nlhe64k_s$sp.json 2>> $O/bench_${TAG}.err
done
for f in $O/bench_${TAG}_nlhe*.json; do python - $f <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "%.4g updates/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()})
PY
done
tail -n 3 $O/bench_${TAG}.err
