#!/bin/bash
O=gpurun_out
TAG=${1:-r1d}
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1
for V in 0 2; do
RBP_STEP_VARIANT=$V timeout 300 python tests/measure/bench_lloyd.py --n 1000000 --k 256 --iters 4 --cpu-n 0 > $O/bench_${TAG}_lloyd_v$V.json 2>> $O/bench_${TAG}.err
RBP_STEP_VARIANT=$V timeout 300 python tests/measure/bench_lloyd.py --n 13960050 --k 256 --iters 3 --cpu-n 0 > $O/bench_${TAG}_lloyd14m_v$V.json 2>> $O/bench_${TAG}.err
done
timeout 300 python tools/bench_lloyd_dist.py --n-per-gpu 4000000 --k 256 > $O/bench_${TAG}_lloyd_dist_n1.json 2>> $O/bench_${TAG}.err
