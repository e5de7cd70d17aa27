#!/bin/bash
# Sinkhorn (flop layer) session: parity of the rewritten half-sweep, launch-shape tuning, before/after, ncu capture.
O=gpurun_out
TAG=${1:-r1m}
mkdir -p $O
timeout 600 python -m pytest tests/test_sinkhorn_gpu.py -x -q --timeout 300 > $O/pytest_sk_${TAG}.log 2>&1; tail -3 $O/pytest_sk_${TAG}.log
B="python tests/measure/bench_sinkhorn.py --n 8000 --k 200 --cpu-pairs 0"
timeout 300 $B --tag w8b2 --cpu-pairs 1024 > $O/sk_${TAG}_w8b2.json 2> $O/sk_${TAG}.err
RBP_SK_WARPS=8 RBP_SK_BLOCKS_PER_SM=1 timeout 200 $B --tag w8b1 > $O/sk_${TAG}_w8b1.json 2>> $O/sk_${TAG}.err
RBP_SK_WARPS=6 RBP_SK_BLOCKS_PER_SM=3 timeout 200 $B --tag w6b3 > $O/sk_${TAG}_w6b3.json 2>> $O/sk_${TAG}.err
RBP_SK_WARPS=4 RBP_SK_BLOCKS_PER_SM=5 timeout 200 $B --tag w4b5 > $O/sk_${TAG}_w4b5.json 2>> $O/sk_${TAG}.err
RBP_SK_WARPS=4 RBP_SK_BLOCKS_PER_SM=3 timeout 200 $B --tag w4b3 > $O/sk_${TAG}_w4b3.json 2>> $O/sk_${TAG}.err
RBP_LIB_PATH=$PWD/tools/_prev/librbp_b200_prev.so timeout 300 $B --tag prev --sweeps 1 > $O/sk_${TAG}_prev.json 2>> $O/sk_${TAG}.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sk_assign_kernel --launch-skip 1 -c 1 -o $O/${TAG}_sk_assign -f \
  python tests/measure/bench_sinkhorn.py --n 2000 --k 64 --cpu-pairs 0 --sweeps 1 --steps 1 > $O/ncu_sk_${TAG}.log 2>&1
cat $O/sk_${TAG}_*.json | cut -c1-900
tail -5 $O/sk_${TAG}.err
