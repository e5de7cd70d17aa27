#!/bin/bash
O=gpurun_out
TAG=${1:-r1c}
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1
timeout 300 python bench.py > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}.err
timeout 300 python bench.py --fold ordered --batch 16384 > $O/bench_${TAG}_ordered16k.json 2>> $O/bench_${TAG}.err
RBP_STEP_VARIANT=0 timeout 300 python tests/measure/bench_lloyd.py --n 1000000 --k 256 --iters 4 --cpu-n 0 > $O/bench_${TAG}_lloyd_v0.json 2>> $O/bench_${TAG}.err
RBP_STEP_VARIANT=1 timeout 300 python tests/measure/bench_lloyd.py --n 1000000 --k 256 --iters 4 --cpu-n 0 > $O/bench_${TAG}_lloyd_v1.json 2>> $O/bench_${TAG}.err
RBP_STEP_VARIANT=0 timeout 300 python tests/measure/bench_lloyd.py --n 13960050 --k 256 --iters 3 --cpu-n 0 > $O/bench_${TAG}_lloyd14m_v0.json 2>> $O/bench_${TAG}.err
RBP_STEP_VARIANT=1 timeout 300 python tests/measure/bench_lloyd.py --n 13960050 --k 256 --iters 3 --cpu-n 0 > $O/bench_${TAG}_lloyd14m_v1.json 2>> $O/bench_${TAG}.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:elkan_step_kernel -s 2 -c 1 -f -o $O/${TAG}_elkan_step python tests/measure/bench_lloyd.py --n 1000000 --k 256 --iters 2 --cpu-n 0 > $O/ncu_${TAG}_elkan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assign_kernel -s 1 -c 1 -f -o $O/${TAG}_assign python tests/measure/bench_lloyd.py --n 1000000 --k 256 --iters 2 --cpu-n 0 > $O/ncu_${TAG}_assign.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mccfr_fold -s 4 -c 1 -f -o $O/${TAG}_fold python bench.py --fold ordered --batch 16384 --steps 6 --warmup 3 > $O/ncu_${TAG}_fold.log 2>&1
