#!/bin/bash
# ncu --set full captures of the two dominant NLHE kernels (batch 16384), plus the bench line
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nlhe_value_kernel --launch-skip 3 -c 1 -o gpurun_out/r1f_nlhe_value -f python tools/nlhe_probe.py 16384 > gpurun_out/ncu_value.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nlhe_expand_kernel --launch-skip 154 -c 1 -o gpurun_out/r1f_nlhe_expand -f python tools/nlhe_probe.py 16384 > gpurun_out/ncu_expand.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nlhe_fold_kernel --launch-skip 3 -c 1 -o gpurun_out/r1f_nlhe_fold -f python tools/nlhe_probe.py 16384 > gpurun_out/ncu_fold.log 2>&1
timeout 600 python bench.py --workload nlhe --steps 20 --warmup 3 > gpurun_out/bench_r1f_nlhe_n1.json 2> gpurun_out/bench_r1f_nlhe.err
cat gpurun_out/bench_r1f_nlhe_n1.json | cut -c1-1500
tail -3 gpurun_out/ncu_value.log gpurun_out/ncu_expand.log
