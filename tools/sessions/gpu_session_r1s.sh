#!/bin/bash
# NLHE fold (exact reciprocal division + prefetch) and Sinkhorn term without the no-op max: parity, then bench lines.
O=gpurun_out
TAG=${1:-r1s}
mkdir -p $O
timeout 900 python -m pytest tests/test_nlhe_gpu.py tests/test_sinkhorn_gpu.py tests/test_mccfr_gpu.py -x -q --timeout 600 > $O/pytest_${TAG}.log 2>&1; tail -3 $O/pytest_${TAG}.log
timeout 400 python bench.py --workload nlhe --steps 30 > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}.err
timeout 400 python bench.py --workload nlhe --batch 65536 --steps 10 > $O/bench_${TAG}_nlhe64k_n1.json 2>> $O/bench_${TAG}.err
for f in nlhe_n1 nlhe64k_n1; do python - $O/bench_${TAG}_$f.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "%.4g updates/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
PY
done
timeout 200 python tests/measure/bench_sinkhorn.py --n 16000 --k 200 --cpu-pairs 0 --tag a02_w8b3 > $O/sk_${TAG}_a02_w8b3.json 2>> $O/bench_${TAG}.err
timeout 300 python tests/measure/bench_sinkhorn.py --n 4000 --k 200 --alpha 0.3 --cpu-pairs 0 --sweeps 1 --tag a30_w8b3 > $O/sk_${TAG}_a30_w8b3.json 2>> $O/bench_${TAG}.err
for f in $O/sk_${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["tag"], "assign %.3g solves/s" % d["assign_solves_per_s"], "%.3g terms/s" % d["assign_exp_terms_per_s"], "frac %.3f" % d["roofline"]["frac"], "step %.1f ms" % d["elkan_step_ms"])
PY
done
tail -n 3 $O/bench_${TAG}.err
