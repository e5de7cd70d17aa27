#!/usr/bin/env python
"""Weak-scaling bench of the point-sharded turn-layer k-means: one process per GPU, points sharded by index range,
centroids replicated, ONE integer all-reduce (K x 102 u64 + tallies) per iteration over NCCL.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_lloyd_dist.py --n-per-gpu 4000000 --k 256
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n-per-gpu", type=int, default=4_000_000)
    p.add_argument("--k", type=int, default=256)
    p.add_argument("--iters", type=int, default=8)
    args = p.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from lloyd_data import turn_histograms

    import robopoker_b200 as rbp
    from robopoker_b200.distributed import allreduce_kmeans_step

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    pts = turn_histograms(args.n_per_gpu, seed=100 + rank)
    seeds = torch.from_numpy(pts[: args.k].astype(np.int64)).cuda()
    if world > 1:
        dist.broadcast(seeds, 0)                      # common initial centroids: rank 0's first K points
    layer = rbp.lloyd.Layer(pts, args.k, device=local)
    layer.set_centroids(seeds.cpu().numpy().astype(np.uint64))
    layer.init_bounds()
    d = dist if world > 1 else None
    for _ in range(2):
        allreduce_kmeans_step(layer, d, device=local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    last = None
    for _ in range(args.iters):
        last = allreduce_kmeans_step(layer, d, device=local)   # drains the library stream and the NCCL stream every step
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.iters
    if rank == 0:
        n_total = args.n_per_gpu * world
        print(json.dumps({"bench": "lloyd_turn_w1_sharded", "n_gpus": world, "n_total": n_total, "k": args.k, "ms_per_iteration": ms,
                          "points_per_s": n_total / (ms * 1e-3), "scaling": "weak", "sizes_sum": int(last.sizes.sum()),
                          "reassigned": int(last.reassignment)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
