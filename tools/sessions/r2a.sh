#!/bin/bash
# r2a: first GPU session of round 2 — full GPU suite + smoke on the new code, the NLHE headline through the new bench, trace
O=gpurun_out
TAG=${1:-r2a}
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > $O/pytest_${TAG}.log 2>&1; tail -3 $O/pytest_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}.err
RBP_NLHE_TRACE=1 timeout 200 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > /dev/null 2> $O/trace_${TAG}.txt
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_${TAG}_ref.json 2>> $O/bench_${TAG}.err
python - $O/bench_${TAG}_nlhe_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "%.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "launches", d["gpu_launches"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
tail -n 5 $O/bench_${TAG}.err; tail -n 3 $O/trace_${TAG}.txt; cat $O/bench_${TAG}_ref.json | cut -c1-300
