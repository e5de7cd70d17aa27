#!/bin/bash
# r2q: the k-means layers at stated size — configs[4] turn sweep (k = 100, 500 at N = 13,960,050; k = 2000 at 4 M points) and configs[2] flop layer
O=gpurun_out
TAG=${1:-r2q}
for K in 100 500; do
timeout 600 python bench.py --workload lloyd_turn --k $K --steps 8 --warmup 3 > $O/bench_${TAG}_turn_k$K.json 2> $O/bench_${TAG}_turn_k$K.err; tail -2 $O/bench_${TAG}_turn_k$K.err; cut -c1-400 $O/bench_${TAG}_turn_k$K.json
done
timeout 600 python bench.py --workload lloyd_turn --k 2000 --points 4000000 --steps 6 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_turn_k2000_4m.json 2> $O/bench_${TAG}_turn_k2000.err; tail -2 $O/bench_${TAG}_turn_k2000.err; cut -c1-400 $O/bench_${TAG}_turn_k2000_4m.json
timeout 900 python bench.py --workload lloyd_flop --steps 3 > $O/bench_${TAG}_flop.json 2> $O/bench_${TAG}_flop.err; tail -2 $O/bench_${TAG}_flop.err; cut -c1-1500 $O/bench_${TAG}_flop.json
