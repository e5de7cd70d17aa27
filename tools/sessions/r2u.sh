#!/bin/bash
# r2u: full GPU suite on the current code (world-of-one exchange tests, waves, apply_child_edge), headline with cpu_baseline, reference arm
O=gpurun_out
TAG=${1:-r2u}
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 > $O/pytest_${TAG}.log 2>&1; tail -4 $O/pytest_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}.err; tail -2 $O/bench_${TAG}.err
timeout 400 python bench.py --impl reference --steps 6 --warmup 2 > $O/bench_${TAG}_ref.json 2>> $O/bench_${TAG}.err
python - $O/bench_${TAG}_nlhe_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("%.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "roofline", d["roofline"]["phase"], round(d["roofline"]["frac"],4), "cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"))
PY
cut -c1-250 $O/bench_${TAG}_ref.json
