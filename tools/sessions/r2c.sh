#!/bin/bash
# r2c: launch list of the NLHE epoch with the merge + chain fold (ncu per-launch durations: cold cache, serialised — shares only)
O=gpurun_out
TAG=${1:-r2c}
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv --log-file $O/${TAG}_nlhe_launches.csv \
  python bench.py --steps 1 --warmup 3 --epochs-per-step 4 --skip-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
tail -2 $O/${TAG}_ncu_bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nlhe_chain_kernel -s 12 -c 1 -o $O/${TAG}_chain \
  python bench.py --steps 1 --warmup 3 --epochs-per-step 4 --skip-cpu-baseline > $O/${TAG}_ncu_chain.log 2>&1
tail -2 $O/${TAG}_ncu_chain.log | cut -c1-300
ls -la $O | tail -5
