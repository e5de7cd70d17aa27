#!/bin/bash
# r2p: screen tests + ncu capture of the tcgen05 screen kernel (tensor pipe %, DRAM, registers)
O=gpurun_out
TAG=${1:-r2p}
timeout 300 python -m pytest tests/test_sinkhorn_gpu.py -x -q -m gpu --timeout 120 2>&1 | tail -5
timeout 200 python tests/measure/screen_probe.py --n 40000 --k 200 --alpha 0.02 2>/dev/null | tail -1 > $O/${TAG}_screen_n40000_a002.json
timeout 200 python tests/measure/screen_probe.py --n 10000 --k 200 --alpha 0.3 2>/dev/null | tail -1 > $O/${TAG}_screen_n10000_a03.json
cut -c1-700 $O/${TAG}_screen_n10000_a03.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sk_screen_kernel -c 1 -o $O/${TAG}_screen python tests/measure/screen_probe.py --n 1500 --k 200 --alpha 0.02 > $O/${TAG}_ncu.log 2>&1
tail -2 $O/${TAG}_ncu.log | cut -c1-200
