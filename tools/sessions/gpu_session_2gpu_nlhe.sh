#!/bin/bash
# 2-GPU session: NLHE MCCFR N=1 and N=2 (NCCL ragged all-gather of update records), and the N=2 vs N=1 table equality check
O=gpurun_out
TAG=${1:-r1g}
mkdir -p $O
timeout 300 python bench.py --workload nlhe --steps 20 > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}_nlhe.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --workload nlhe --gpus 2 --steps 20 > $O/bench_${TAG}_nlhe_n2.json 2>> $O/bench_${TAG}_nlhe.err
# parity across the exchange: 2 ranks x 4096 trees must equal 1 rank x 8192 trees, bit for bit
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/nlhe_world_check.py > $O/nlhe_world_check_${TAG}.log 2>&1
tail -3 $O/nlhe_world_check_${TAG}.log
cut -c1-400 $O/bench_${TAG}_nlhe_n1.json $O/bench_${TAG}_nlhe_n2.json
tail -5 $O/bench_${TAG}_nlhe.err
