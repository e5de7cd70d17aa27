#!/bin/bash
# r2y: Elkan step with far prefetches of the bounds stream into L2 (variants 4: 12 centroids ahead, 5: 24) against variant 2
O=gpurun_out
TAG=${1:-r2y}
RBP_STEP_VARIANT=4 timeout 600 python -m pytest tests/test_lloyd_gpu.py -x -q -m gpu --timeout 300 2>&1 | tail -2
for V in 2 4 5; do for K in 100 500; do
RBP_STEP_VARIANT=$V timeout 400 python bench.py --workload lloyd_turn --k $K --points 6000000 --steps 8 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_v${V}_k$K.json 2> $O/bench_${TAG}_v${V}_k$K.err; tail -1 $O/bench_${TAG}_v${V}_k$K.err
python - $O/bench_${TAG}_v${V}_k$K.json $V $K <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("variant", sys.argv[2], "k", sys.argv[3], "%.3f ms/iter" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"])
PY
done; done
