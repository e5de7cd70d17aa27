#!/bin/bash
# Sinkhorn (flop layer) session 3: parity, launch-shape sweep, ncu.
O=gpurun_out
TAG=${1:-r1o}
mkdir -p $O
timeout 600 python -m pytest tests/test_sinkhorn_gpu.py -x -q --timeout 300 > $O/pytest_sk_${TAG}.log 2>&1; tail -3 $O/pytest_sk_${TAG}.log
B="python tests/measure/bench_sinkhorn.py --n 16000 --k 200 --cpu-pairs 0"
for shape in "8 3" "8 4" "8 2" "4 6" "4 8"; do set -- $shape
  RBP_SK_WARPS=$1 RBP_SK_BLOCKS_PER_SM=$2 timeout 200 $B --tag a02_w$1b$2 > $O/sk_${TAG}_a02_w$1b$2.json 2>> $O/sk_${TAG}.err
done
B3="python tests/measure/bench_sinkhorn.py --n 4000 --k 200 --alpha 0.3 --cpu-pairs 0 --sweeps 1"
for shape in "8 3" "8 4"; do set -- $shape
  RBP_SK_WARPS=$1 RBP_SK_BLOCKS_PER_SM=$2 timeout 300 $B3 --tag a30_w$1b$2 > $O/sk_${TAG}_a30_w$1b$2.json 2>> $O/sk_${TAG}.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sk_assign_kernel --launch-skip 1 -c 1 -o $O/${TAG}_sk_assign -f \
  python tests/measure/bench_sinkhorn.py --n 3000 --k 64 --cpu-pairs 0 --sweeps 1 --steps 1 > $O/ncu_sk_${TAG}.log 2>&1
for f in $O/sk_${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["tag"], "pt/cen support %.1f/%.1f" % (d["mean_point_support"], d["mean_centroid_support"]), "assign %.3g solves/s" % d["assign_solves_per_s"],
      "%.3g terms/s" % d["assign_exp_terms_per_s"], "frac %.3f" % d["roofline"]["frac"], "sweeps/solve %.1f" % d["assign_sweeps_per_solve"],
      "step %.1f ms" % d["elkan_step_ms"], "pp %.2fs bounds %.2fs" % (d["init_pp_s"], d["init_bounds_s"]), d.get("cpu_baseline", ""))
PY
done
tail -5 $O/sk_${TAG}.err
