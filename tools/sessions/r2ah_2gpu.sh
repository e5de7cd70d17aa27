#!/bin/bash
# r2ah (2 GPUs): the round's final code across ranks — NLHE library exchange (with sampling waves) against one process, the bench at N=2,
# and the sharded turn-layer iteration (fused integer all-reduce inside the library) on the reordered, ring-fed Elkan step
O=gpurun_out
TAG=${1:-r2ah}
mkdir -p $O
RBP_CHECK_BATCH=4096 RBP_CHECK_EPOCHS=4 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nlhe_world_check.py > $O/${TAG}_world_check.txt 2>&1
tail -4 $O/${TAG}_world_check.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_${TAG}_nlhe_n2.json 2> $O/bench_${TAG}.err
tail -2 $O/bench_${TAG}.err
python - $O/bench_${TAG}_nlhe_n2.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print("%.4g updates/s" % d["value"], "e2e %.4g" % d["e2e"]["value"], "%.3f ms/step" % d["ms_per_step"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()})
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload lloyd_turn --k 100 --steps 8 --warmup 3 --skip-cpu-baseline > $O/bench_${TAG}_lloyd_turn_k100_n2.json 2> $O/bench_${TAG}_lloyd.err
tail -2 $O/bench_${TAG}_lloyd.err
python - $O/bench_${TAG}_lloyd_turn_k100_n2.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print("lloyd_turn N=13.96M k=100 on 2 GPUs: %.3f ms/iter" % d["ms_per_step"], "reassigned", d["reassigned_last"])
PY
