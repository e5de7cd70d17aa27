#!/bin/bash
# r2e: fold trace (sub-phase times + longest segment) in a natural run
O=gpurun_out
TAG=${1:-r2e}
RBP_NLHE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --epochs-per-step 8 --skip-cpu-baseline > $O/bench_${TAG}.json 2> $O/trace_${TAG}.txt
tail -n 6 $O/trace_${TAG}.txt; cut -c1-200 $O/bench_${TAG}.json
