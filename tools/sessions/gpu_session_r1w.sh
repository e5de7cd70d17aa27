#!/bin/bash
# windowed scatter: NLHE parity gate, traced probe, bench lines
O=gpurun_out
TAG=${1:-r1w}
mkdir -p $O
timeout 600 python -m pytest tests/test_nlhe_gpu.py -x -q --timeout 300 > $O/pytest_nlhe_${TAG}.log 2>&1 || { tail -20 $O/pytest_nlhe_${TAG}.log; echo "NLHE PARITY FAILED"; exit 1; }
tail -1 $O/pytest_nlhe_${TAG}.log
RBP_NLHE_TRACE=1 timeout 200 python tools/nlhe_probe.py 16384 65536 > $O/${TAG}_nlhe_trace.txt 2>&1; cat $O/${TAG}_nlhe_trace.txt | cut -c1-300
timeout 300 python bench.py --workload nlhe --steps 30 > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}.err
timeout 300 python bench.py --workload nlhe --batch 65536 --steps 10 --skip-cpu-baseline > $O/bench_${TAG}_nlhe64k_n1.json 2>> $O/bench_${TAG}.err
for f in nlhe_n1 nlhe64k_n1; do python - $O/bench_${TAG}_$f.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "%.4g updates/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "launches", d["gpu_launches"])
PY
done
tail -n 3 $O/bench_${TAG}.err
