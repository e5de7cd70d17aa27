#!/bin/bash
# One GPU-box session: parity tests, headline + secondary benches, ncu launch list and per-kernel captures.
# Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ afterwards on the build host.
O=gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1
timeout 300 python bench.py > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}.err
timeout 300 python bench.py --fold ordered --batch 16384 > $O/bench_${TAG}_ordered16k.json 2>> $O/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $O/bench_${TAG}_ref.json 2>> $O/bench_${TAG}.err
timeout 300 python tests/measure/bench_lloyd.py --n 1000000 --k 256 --iters 4 > $O/bench_${TAG}_lloyd.json 2>> $O/bench_${TAG}.err
# launch list of the headline command (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_${TAG}.csv python bench.py --steps 10 --warmup 3 > $O/ncu_${TAG}_0.log 2>&1
for K in mccfr_sample_kernel mccfr_rank_partial mccfr_apply_batched; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o $O/${TAG}_$K python bench.py --steps 6 --warmup 3 > $O/ncu_${TAG}_$K.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:elkan_step_kernel -s 2 -c 1 -f -o $O/${TAG}_elkan_step python tests/measure/bench_lloyd.py --n 1000000 --k 256 --iters 2 --cpu-n 0 > $O/ncu_${TAG}_elkan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assign_kernel -s 1 -c 1 -f -o $O/${TAG}_assign python tests/measure/bench_lloyd.py --n 1000000 --k 256 --iters 2 --cpu-n 0 > $O/ncu_${TAG}_assign.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:river_equity -c 1 -f -o $O/${TAG}_river_equity python tests/measure/bench_deuce.py --n 500000 --cpu-n 100 > $O/ncu_${TAG}_equity.log 2>&1
timeout 900 python tools/abstraction_pipeline.py --iters 8 > $O/pipeline_${TAG}.json 2> $O/pipeline_${TAG}.err
ls -la $O | tail -30
