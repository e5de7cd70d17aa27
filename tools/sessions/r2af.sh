#!/bin/bash
# r2af: warp-aggregated accumulate kernel (the recompute of the Elkan iteration): parity tests, then the step at 6 M points with and without the point reorder
O=gpurun_out
TAG=${1:-r2af}
timeout 600 python -m pytest tests/test_lloyd_gpu.py tests/test_sinkhorn_gpu.py -x -q -m gpu --timeout 300 2>&1 | tail -2
RBP_W1_REORDER=0 timeout 600 python -m pytest tests/test_lloyd_gpu.py -x -q -m gpu --timeout 300 2>&1 | tail -2
for R in 1 0; do for K in 100 500; do
RBP_W1_REORDER=$R timeout 400 python bench.py --workload lloyd_turn --k $K --points 6000000 --steps 8 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_r${R}_k$K.json 2> $O/bench_${TAG}_r${R}_k$K.err; tail -1 $O/bench_${TAG}_r${R}_k$K.err
python - $O/bench_${TAG}_r${R}_k$K.json $R $K <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("reorder", sys.argv[2], "k", sys.argv[3], "%.3f ms/iter" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "reassigned", d["reassigned_last"])
PY
done; done
