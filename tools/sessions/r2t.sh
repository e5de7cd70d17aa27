#!/bin/bash
# r2t: source-level ncu captures of the two level kernels at a wide level (where the tree build's time goes), one wave
O=gpurun_out
TAG=${1:-r2t}
export RBP_NLHE_WAVES=1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nlhe_classify_kernel -s 78 -c 1 -o $O/${TAG}_classify python bench.py --steps 1 --warmup 3 --epochs-per-step 2 --skip-cpu-baseline > $O/${TAG}_ncu1.log 2>&1; tail -1 $O/${TAG}_ncu1.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nlhe_expand_kernel -s 78 -c 1 -o $O/${TAG}_expand python bench.py --steps 1 --warmup 3 --epochs-per-step 2 --skip-cpu-baseline > $O/${TAG}_ncu2.log 2>&1; tail -1 $O/${TAG}_ncu2.log | cut -c1-200
