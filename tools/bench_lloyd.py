"""bench.py --workload lloyd_turn | lloyd_flop — the k-means abstraction layers at the sizes BASELINE.json states.

lloyd_turn = configs[4]: N = 13,960,050 turn histograms over the 101 river-equity buckets (`Equity::variation`), k in {100, 500, 2000}
             (`--k`), one step = one Elkan iteration (`Elkan::step_elkan`, crates/elkan/src/elkan.rs:153-168).  With --gpus N the points
             are sharded by index (strong scaling: the street's point count is fixed) and the one integer all-reduce runs inside the
             library (rbp_kmeans_attach_comm).
lloyd_flop = configs[2]: N = 1,286,792 flop histograms over 256 turn clusters, K = 200, `Sinkhorn::divergence`; measured: the naive
             N x K sweep behind `init_bounds` / `lookup` through the tensor-core screen, then `--steps` (capped at 3) Elkan iterations.
Synthetic points (tests/lloyd_data.py, seeded), centroids = the first K points of rank 0 (k-means++ itself is K more N-distance sweeps).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TURN_N, FLOP_N = 13_960_050, 1_286_792  # crates/deuce/src/street.rs:129-135


def _dist(world, local):
    import torch

    torch.cuda.set_device(local)
    if world == 1:
        return None, None
    import torch.distributed as dist
    from robopoker_b200.comm import Comm

    dist.init_process_group("nccl")
    return dist, Comm.from_torch(dist, device=local)


def _common_centroids(dist, pts, k):
    import numpy as np
    import torch

    seeds = torch.from_numpy(pts[:k].astype(np.int64)).cuda()
    if dist is not None:
        dist.broadcast(seeds, 0)
    return seeds.cpu().numpy().astype(np.uint64)


def _max(dist, x):
    if dist is None:
        return x
    import torch

    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main_turn(args, ranks, Clocks, peaks, oracle_native):
    import numpy as np
    import torch
    from lloyd_data import turn_histograms

    import robopoker_b200 as rbp

    rank, world, local = ranks
    dist, comm = _dist(world, local)
    k = args.k or 256
    n_total = args.points or TURN_N
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    pts = np.concatenate([turn_histograms(min(2_000_000, hi - s), seed=1000 + s) for s in range(lo, hi, 2_000_000)])
    layer = rbp.lloyd.Layer(pts, k, device=local)
    if comm:
        layer.attach_comm(comm)
    layer.set_centroids(_common_centroids(dist, pts, k))
    layer.init_bounds()
    l = rbp.load_library()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        layer.step()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = Clocks(local)
    l0 = l.rbp_kernel_launches()
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):   # each step returns drift / sizes / reassignment to host buffers (K x 8 + 4 bytes): this IS the host-facing call
        last = layer.step()
    torch.cuda.synchronize()
    ms_total = _max(dist, (time.perf_counter() - t0) * 1e3)
    launches = l.rbp_kernel_launches() - l0
    clk = clocks.stop()
    ms = ms_total / args.steps
    dev_ms = layer.timed(0, max(2, min(args.steps, 8))) / max(2, min(args.steps, 8)) if world == 1 else None   # CUDA events on the library stream
    peak, peak_src = peaks()
    # algorithmic bytes per point per iteration: the 112-byte point + the K lower bounds read and written + upper / assign / stale
    alg = (hi - lo) * (112.0 + 8.0 * k + 9.0)
    kernel_s = (dev_ms if dev_ms else ms) * 1e-3
    if rank == 0:
        line = {"metric": "points clustered/sec (one Elkan iteration)", "value": n_total / (ms * 1e-3), "unit": "points/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"configs[4] turn-layer k-means, N = {n_total} histograms x 101 river-equity buckets, k = {k}, Equity::variation, one Elkan iteration per step",
                           "k": k, "points": n_total, "parallelism": f"points sharded x{world}, one integer all-reduce per iteration inside librbp_b200" if world > 1 else "points x1",
                           "l2": f"bounds stream of {(hi - lo) * 4 * k / 1e9:.1f} GB per iteration exceeds the 126 MB L2"},
                "clocks": clk, "e2e": {"value": n_total / (ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * k + 4,
                                       "note": "points are resident (uploaded once at Layer::build); every step returns drift, sizes and the reassignment count to host buffers"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "elkan_step_kernel", "achieved": alg / kernel_s / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg / kernel_s / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                             "kernel_ms": {"step (events on the library stream)": dev_ms}, "note": "whole step against the bounds-stream model 112 + 8k + 9 bytes per point; ncu of the shipped step at 3 M x 500 (profiles/r2ai_elkan_v8_ncu.txt): "
                                     "DRAM 12.45 GB per launch = 8.3 B per (point, centroid) pair, i.e. the model's traffic and nothing re-read; the early, loosely "
                                     "pruned iterations are bound by the distance arithmetic, not by this stream (DESIGN 3)"},
                "reassigned_last": int(last.reassignment), "distance_evals_upper_bound_per_s": n_total * k / (ms * 1e-3)}
        if world == 1 and not args.skip_cpu_baseline:
            oracle = oracle_native()
            threads = os.cpu_count() or 1
            m = min(len(pts), 400_000)
            o = oracle.OracleKmeans(pts[:m], k, threads=threads)
            o.set_centroids_from_points(np.arange(k))
            o.init_bounds()
            o.step()
            c0 = time.perf_counter()
            it = 0
            while it < 8 and (it == 0 or time.perf_counter() - c0 < 10.0):
                o.step(); it += 1
            dt = time.perf_counter() - c0
            line["cpu_baseline"] = {"value": m * it / dt, "unit": "points/s", "cores": threads, "kind": "port",
                                    "sample": f"{it} Elkan iterations over the first {m} points in {dt:.1f}s, C++ restatement of the reference rayon path, -O3 -march=native"}
        print(json.dumps(line))
    layer.close()
    if comm:
        comm.close()
    if dist:
        dist.destroy_process_group()


def main_flop(args, ranks, Clocks, peaks, oracle_native):
    import numpy as np
    import torch
    from lloyd_data import flop_mixture_histograms, synthetic_metric

    import robopoker_b200 as rbp

    rank, world, local = ranks
    dist, comm = _dist(world, local)
    k = args.k or 200
    n_total = args.points or FLOP_N
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    pts = flop_mixture_histograms(n_total, 256, comps=k, alpha=0.02, seed=0)[lo:hi]   # mean support ~12, as measured on real flop projections
    tri = synthetic_metric(256, 0)
    t0 = time.perf_counter()
    layer = rbp.lloyd.Layer(pts, k, metric=tri, device=local)   # uploads the points, OT(x, x) of every point
    t_build = time.perf_counter() - t0
    if comm:
        layer.attach_comm(comm)
    layer.set_centroids(_common_centroids(dist, flop_mixture_histograms(k, 256, comps=k, alpha=0.02, seed=1), k))
    l = rbp.load_library()
    margin = 2e-4
    layer.screen(margin)
    layer.sinkhorn_stats(reset=True)
    clocks = Clocks(local)
    l0 = l.rbp_kernel_launches()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    layer.init_bounds()                                         # `Elkan::init_bounds`: the N x K sweep, screened on the tensor cores
    t_bounds = _max(dist, time.perf_counter() - t0)
    solves_bounds = layer.sinkhorn_stats(reset=True)[0]
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    outs = [layer.step() for _ in range(steps)]
    t_steps = _max(dist, time.perf_counter() - t0)
    solves_steps = layer.sinkhorn_stats(reset=True)[0]
    t0 = time.perf_counter()
    assign = layer.lookup()                                     # `Layer::lookup`: the final N x K sweep, screened
    t_lookup = _max(dist, time.perf_counter() - t0)
    solves_lookup = layer.sinkhorn_stats(reset=True)[0]
    launches = l.rbp_kernel_launches() - l0
    clk = clocks.stop()
    exact = None
    if args.points and args.points <= 100_000 and world == 1:   # small runs also time the unscreened sweep and compare
        layer.screen(-1.0)
        t0 = time.perf_counter(); a2 = layer.lookup(); exact = {"exact_lookup_s": time.perf_counter() - t0, "identical": bool(np.array_equal(assign, a2))}
    if rank == 0:
        pairs = n_total * k
        line = {"metric": "point-centroid pairs resolved/sec (Layer::lookup: argmin + winning Sinkhorn divergence)", "value": pairs / t_lookup, "unit": "pairs/s",
                "n_gpus": world, "steps": steps, "warmup": 0, "ms_per_step": t_steps / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (screen: bf16 hi+lo x bf16, f32 accumulate)", "data": "synthetic",
                "config": {"workload": f"configs[2] flop-layer k-means, N = {n_total} histograms x 256 turn clusters, K = {k}, Sinkhorn::divergence (T 0.025, 128 iterations, tol 5e-4), "
                                       f"screen margin {margin}", "k": k, "points": n_total, "parallelism": f"points sharded x{world}" if world > 1 else "points x1",
                           "l2": "per-point working set streams from HBM (N x K approximate-divergence matrix, 1 GB)"},
                "clocks": clk, "e2e": {"value": pairs / t_lookup, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * (hi - lo),
                                       "note": "wall clock of rbp_kmeans_assign including the D2H copy of the assignments"},
                "gpu_launches": int(launches),
                "phases_s": {"layer_build": t_build, "init_bounds": t_bounds, "elkan_steps": t_steps, "lookup": t_lookup},
                "exact_ot_solves": {"init_bounds": solves_bounds, "elkan_steps": solves_steps, "lookup": solves_lookup, "full_sweep_would_be": (hi - lo) * k},
                "reassigned": [int(o.reassignment) for o in outs], "unscreened": exact,
                "roofline": {"bound": "tensor", "kernel": "sk_screen_kernel", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                             "note": "see profiles/r2p_screen_ncu.txt: the screen is bound by its element-wise epilogue (TMEM <-> registers), the MMAs are small by construction"}}
        print(json.dumps(line))
    layer.close()
    if comm:
        comm.close()
    if dist:
        dist.destroy_process_group()


def main(args, ranks, Clocks, peaks, oracle_native):
    if args.impl == "reference":
        if ranks[0] == 0:
            print(json.dumps({"impl": "reference", "unavailable": "the lloyd workloads carry their CPU baseline in the b200 arm's cpu_baseline (bounded oracle sample)"}))
        return
    return main_turn(args, ranks, Clocks, peaks, oracle_native) if args.workload == "lloyd_turn" else main_flop(args, ranks, Clocks, peaks, oracle_native)
