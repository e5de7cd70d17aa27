#!/bin/bash
O=gpurun_out
TAG=${1:-r1e}
timeout 300 python -m pytest tests/test_mccfr_gpu.py -x -q > $O/pytest_${TAG}.log 2>&1
timeout 200 python bench.py --fold ordered --batch 16384 > $O/bench_${TAG}_ordered16k.json 2> $O/bench_${TAG}.err
timeout 200 python bench.py > $O/bench_${TAG}_n1.json 2>> $O/bench_${TAG}.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > $O/bench_${TAG}_n2.json 2>> $O/bench_${TAG}.err
timeout 300 python tools/bench_lloyd_dist.py --n-per-gpu 4000000 --k 256 > $O/bench_${TAG}_lloyd_dist_n1.json 2>> $O/bench_${TAG}.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/bench_lloyd_dist.py --n-per-gpu 4000000 --k 256 > $O/bench_${TAG}_lloyd_dist_n2.json 2>> $O/bench_${TAG}.err
