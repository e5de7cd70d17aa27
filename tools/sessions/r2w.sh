#!/bin/bash
# r2w: how much of the epoch is the profile table's footprint?  148 k infosets live in the table; slots 2^19 .. 2^25
O=gpurun_out
TAG=${1:-r2w}
for S in 19 20 22 25; do
timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --table-slots $((1<<S)) > $O/bench_${TAG}_slots$S.json 2> $O/bench_${TAG}_slots$S.err; tail -1 $O/bench_${TAG}_slots$S.err
python - $O/bench_${TAG}_slots$S.json $S <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("slots 2^"+sys.argv[2], "%.4g updates/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "rows", d["table_rows"])
PY
done
