#!/bin/bash
# r2ab: ncu --set full with source of the shipped Elkan step (variant 6) at 3 M points, k = 500
O=gpurun_out
TAG=${1:-r2ab}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:elkan_step_kernel -s 4 -c 1 -o $O/${TAG}_elkan_step python bench.py --workload lloyd_turn --k 500 --points 3000000 --steps 3 --warmup 3 --skip-cpu-baseline > $O/${TAG}_ncu.log 2>&1; tail -2 $O/${TAG}_ncu.log | cut -c1-300
ls -la $O/${TAG}_elkan_step.ncu-rep
