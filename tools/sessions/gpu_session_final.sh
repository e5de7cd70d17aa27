#!/bin/bash
# Round-end style session on one B200: full GPU test suite, smoke, both bench lines, NLHE launch list and ncu captures.
O=gpurun_out
TAG=${1:-r1i}
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 300 > $O/pytest_${TAG}.log 2>&1; tail -3 $O/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
timeout 400 python bench.py > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}.err
timeout 400 python bench.py --workload nlhe --steps 30 > $O/bench_${TAG}_nlhe_n1.json 2>> $O/bench_${TAG}.err
timeout 400 python bench.py --workload nlhe --batch 65536 --steps 10 > $O/bench_${TAG}_nlhe64k_n1.json 2>> $O/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_nlhe_launches.csv python tools/nlhe_probe.py 16384 > $O/ncu_launch_${TAG}.log 2>&1
for k in value classify expand fold; do
  skip=3; [ $k = expand ] && skip=154; [ $k = classify ] && skip=154
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:nlhe_${k}_kernel --launch-skip $skip -c 1 -o $O/${TAG}_nlhe_$k -f python tools/nlhe_probe.py 16384 > $O/ncu_${k}_${TAG}.log 2>&1
done
cut -c1-300 $O/bench_${TAG}_n1.json $O/bench_${TAG}_nlhe_n1.json $O/bench_${TAG}_nlhe64k_n1.json
tail -3 $O/bench_${TAG}.err
