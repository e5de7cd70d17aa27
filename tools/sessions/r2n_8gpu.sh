#!/bin/bash
# r2n (8 GPUs): in-library exchange at world 8 — parity against one process, then the bench at N = 8, 4, 1
O=gpurun_out
TAG=${1:-r2n}
mkdir -p $O
RBP_CHECK_BATCH=2048 RBP_CHECK_EPOCHS=3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/nlhe_world_check.py 2>&1 | grep -v Warning | grep "identical" > $O/${TAG}_world_check.txt
cat $O/${TAG}_world_check.txt
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_${TAG}_nlhe_n$N.json 2> $O/bench_${TAG}_n$N.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_nlhe_n1.json 2> $O/bench_${TAG}_n1.err
for N in 1 4 8; do python - $O/bench_${TAG}_nlhe_n$N.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print(d["n_gpus"], "%.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()})
PY
done
tail -2 $O/bench_${TAG}_n8.err
