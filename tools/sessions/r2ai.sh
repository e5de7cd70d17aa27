#!/bin/bash
# r2ai: validation of the round's final code (full GPU suite, smoke, the driver's bench command), configs[4] at full N, and an
# ncu --set full capture of the shipped Elkan step (variant 8) and the recompute kernel
O=gpurun_out
TAG=${1:-r2ai}
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_${TAG}.log 2>&1; tail -1 $O/smoke_${TAG}.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}_nlhe_n1.err; tail -1 $O/bench_${TAG}_nlhe_n1.err
python - $O/bench_${TAG}_nlhe_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("nlhe %.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"])
PY
for K in 100 500; do
timeout 400 python bench.py --workload lloyd_turn --k $K --steps 8 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_lloyd_turn_k$K.json 2> $O/bench_${TAG}_lloyd_turn_k$K.err; tail -1 $O/bench_${TAG}_lloyd_turn_k$K.err
python - $O/bench_${TAG}_lloyd_turn_k$K.json $K <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("lloyd_turn k", sys.argv[2], "%.3f ms/iter" % d["ms_per_step"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "reassigned", d["reassigned_last"])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"elkan_step_kernel|accumulate_kernel" -s 8 -c 2 -o $O/${TAG}_elkan_v8 python bench.py --workload lloyd_turn --k 500 --points 3000000 --steps 3 --warmup 3 --skip-cpu-baseline > $O/${TAG}_ncu.log 2>&1; tail -1 $O/${TAG}_ncu.log | cut -c1-120
ls -la $O/${TAG}_elkan_v8.ncu-rep
