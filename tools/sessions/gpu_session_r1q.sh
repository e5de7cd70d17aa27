#!/bin/bash
# Full GPU suite + smoke + bench lines (Leduc headline, NLHE 16k / 64k) after the Sinkhorn rewrite and the NLHE level prediction.
O=gpurun_out
TAG=${1:-r1q}
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 --durations=8 > $O/pytest_${TAG}.log 2>&1; tail -14 $O/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
timeout 400 python bench.py > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}.err
timeout 400 python bench.py --workload nlhe --steps 30 > $O/bench_${TAG}_nlhe_n1.json 2>> $O/bench_${TAG}.err
timeout 400 python bench.py --workload nlhe --batch 65536 --steps 10 > $O/bench_${TAG}_nlhe64k_n1.json 2>> $O/bench_${TAG}.err
cut -c1-420 $O/bench_${TAG}_n1.json $O/bench_${TAG}_nlhe_n1.json $O/bench_${TAG}_nlhe64k_n1.json
grep -o '"kernel_ms": {[^}]*}' $O/bench_${TAG}_nlhe_n1.json $O/bench_${TAG}_nlhe64k_n1.json
tail -n 3 $O/bench_${TAG}.err
