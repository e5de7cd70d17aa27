#!/bin/bash
# r2aa: Elkan step with the bounds stream fetched by cp.async into a per-thread shared-memory ring (variants 6: 8 groups ahead, 96-centroid
# tiles; 7: 16 groups ahead, 64-centroid tiles) against variant 4
O=gpurun_out
TAG=${1:-r2aa}
for V in 6 7; do RBP_STEP_VARIANT=$V timeout 600 python -m pytest tests/test_lloyd_gpu.py -x -q -m gpu --timeout 300 2>&1 | tail -2; done
for V in 4 6 7; do for K in 100 500; do
RBP_STEP_VARIANT=$V timeout 400 python bench.py --workload lloyd_turn --k $K --points 6000000 --steps 8 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_v${V}_k$K.json 2> $O/bench_${TAG}_v${V}_k$K.err; tail -1 $O/bench_${TAG}_v${V}_k$K.err
python - $O/bench_${TAG}_v${V}_k$K.json $V $K <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("variant", sys.argv[2], "k", sys.argv[3], "%.3f ms/iter" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "reassigned", d["reassigned_last"])
PY
done; done
