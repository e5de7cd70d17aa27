#!/bin/bash
# Round-end style session: NLHE parity gate, full GPU suite, smoke, bench lines, NLHE launch list.
O=gpurun_out
TAG=${1:-r1v}
mkdir -p $O
timeout 600 python -m pytest tests/test_nlhe_gpu.py -x -q --timeout 300 > $O/pytest_nlhe_${TAG}.log 2>&1 || { tail -20 $O/pytest_nlhe_${TAG}.log; echo "NLHE PARITY FAILED"; exit 1; }
tail -1 $O/pytest_nlhe_${TAG}.log
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 > $O/pytest_${TAG}.log 2>&1; tail -3 $O/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
timeout 300 python bench.py > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}.err
timeout 300 python bench.py --workload nlhe --steps 30 > $O/bench_${TAG}_nlhe_n1.json 2>> $O/bench_${TAG}.err
timeout 300 python bench.py --workload nlhe --batch 65536 --steps 10 --skip-cpu-baseline > $O/bench_${TAG}_nlhe64k_n1.json 2>> $O/bench_${TAG}.err
for f in n1 nlhe_n1 nlhe64k_n1; do python - $O/bench_${TAG}_$f.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "%.4g updates/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "launches", d["gpu_launches"])
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_nlhe_launches.csv python tools/nlhe_probe.py 16384 > $O/ncu_launch_${TAG}.log 2>&1
tail -n 3 $O/bench_${TAG}.err
