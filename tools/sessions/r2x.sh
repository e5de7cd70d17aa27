#!/bin/bash
# r2x: Elkan step with TMA bulk copies into a two-stage tile ring (variant 3) against the staged-load kernel (variant 2): parity, then timing
O=gpurun_out
TAG=${1:-r2x}
RBP_STEP_VARIANT=3 timeout 600 python -m pytest tests/test_lloyd_gpu.py tests/test_pins.py tests/test_comm_gpu.py -x -q -m gpu --timeout 300 2>&1 | tail -3
for V in 2 3; do for K in 100 500; do
RBP_STEP_VARIANT=$V timeout 400 python bench.py --workload lloyd_turn --k $K --points 6000000 --steps 8 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_v${V}_k$K.json 2> $O/bench_${TAG}_v${V}_k$K.err; tail -1 $O/bench_${TAG}_v${V}_k$K.err
python - $O/bench_${TAG}_v${V}_k$K.json $V $K <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("variant", sys.argv[2], "k", sys.argv[3], "%.3f ms/iter" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"])
PY
done; done
