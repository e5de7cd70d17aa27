#!/bin/bash
# r2m (2 GPUs): in-library exchange — parity of `world` ranks against one process, then the bench at N=2
O=gpurun_out
TAG=${1:-r2m}
mkdir -p $O
RBP_CHECK_BATCH=4096 RBP_CHECK_EPOCHS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nlhe_world_check.py > $O/${TAG}_world_check.txt 2>&1
tail -5 $O/${TAG}_world_check.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_${TAG}_nlhe_n2.json 2> $O/bench_${TAG}.err
tail -3 $O/bench_${TAG}.err
python - $O/bench_${TAG}_nlhe_n2.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print("%.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()})
PY
