#!/bin/bash
# r2v: DRAM traffic + duration of every kernel of NLHE epochs (metrics pass, one wave so that launches are in phase order)
O=gpurun_out
TAG=${1:-r2v}
RBP_NLHE_WAVES=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1400 -c 450 --csv --log-file $O/${TAG}_nlhe_traffic.csv \
  python bench.py --steps 1 --warmup 3 --epochs-per-step 4 --skip-cpu-baseline > $O/${TAG}_ncu.log 2>&1
tail -1 $O/${TAG}_ncu.log | cut -c1-200; wc -l $O/${TAG}_nlhe_traffic.csv
