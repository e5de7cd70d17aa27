#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[2], flop abstraction layer): the Sinkhorn N x K assignment sweep and one
Elkan step on synthetic flop histograms (47 draws from a Dirichlet mixture over 256 turn clusters) against K centroids,
reported as OT solves/s and exp terms/s against the FP32-issue ceiling, next to the oracle on the host cores.

    python tests/measure/bench_sinkhorn.py --n 20000 --k 200
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# FMA-pipe instructions per exp term of a softmin in the shipped SASS (sub, shifter fma + sub, 2 range-reduction fma,
# 5 Horner fma, r*r, fma, +1, running add); the two clamps, MIN_POSITIVE max and exponent add issue on the ALU pipe
FMA_PIPE_PER_TERM = 14


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=20000)
    p.add_argument("--k", type=int, default=200)
    p.add_argument("--bins", type=int, default=256)
    p.add_argument("--sweeps", type=int, default=2, help="timed N x K assignment sweeps")
    p.add_argument("--steps", type=int, default=1, help="timed Elkan steps")
    p.add_argument("--cpu-pairs", type=int, default=2048, help="point-centroid solves for the oracle's bounded sample (0 = skip)")
    p.add_argument("--alpha", type=float, default=0.02,
                   help="Dirichlet concentration of the synthetic mixture: 0.02 = mean point support ~11 (as measured on real flop "
                        "projections), 0.3 = SURVEY 8d config 3 (support ~34)")
    p.add_argument("--tag", default="")
    args = p.parse_args()
    import numpy as np
    from lloyd_data import flop_mixture_histograms, synthetic_metric

    import robopoker_b200 as rbp

    fadd = ctypes.c_float()
    rbp.load_library().rbp_measure_fadd_peak(ctypes.byref(fadd))
    pts = flop_mixture_histograms(args.n, args.bins, comps=args.k, alpha=args.alpha, seed=0)
    tri = synthetic_metric(args.bins, 0)
    t0 = time.perf_counter()
    g = rbp.lloyd.Layer(pts, args.k, metric=tri)
    t_create = time.perf_counter() - t0
    t0 = time.perf_counter(); g.init_centroids(0); t_pp = time.perf_counter() - t0
    s_pp = g.sinkhorn_stats(reset=True)
    t0 = time.perf_counter(); g.init_bounds(); t_bounds = time.perf_counter() - t0
    s_bounds = g.sinkhorn_stats(reset=True)
    g.step()  # centroids become merged member sums (wide supports), as in every later iteration
    g.sinkhorn_stats(reset=True)
    ms_assign = g.timed(1, args.sweeps) / args.sweeps
    solves, sweeps, terms = (x / args.sweeps for x in g.sinkhorn_stats(reset=True))
    solves = args.n * args.k  # one OT solve per (point, centroid); the counters are absent in older builds
    ms_step = g.timed(0, args.steps) / args.steps
    st_solves, st_sweeps, st_terms = (x / args.steps for x in g.sinkhorn_stats(reset=True))
    counts, _ = g.future()
    peak_terms = fadd.value * 1e12 / FMA_PIPE_PER_TERM
    line = {"bench": "lloyd_flop_sinkhorn", "tag": args.tag, "n": args.n, "k": args.k, "bins": args.bins, "alpha": args.alpha,
            "mean_point_support": float((pts > 0).sum(axis=1).mean()), "mean_centroid_support": float((counts > 0).sum(axis=1).mean()),
            "create_self_terms_s": t_create, "init_pp_s": t_pp, "init_pp_solves": s_pp[0], "init_bounds_s": t_bounds,
            "init_bounds_solves_per_s": args.n * args.k / t_bounds,
            "assign_sweep_ms": ms_assign, "assign_solves_per_s": solves / (ms_assign * 1e-3), "assign_sweeps_per_solve": sweeps / max(solves, 1),
            "assign_exp_terms_per_s": terms / (ms_assign * 1e-3),
            "elkan_step_ms": ms_step, "elkan_step_solves": st_solves, "elkan_step_exp_terms_per_s": st_terms / (ms_step * 1e-3),
            "roofline": {"bound": "fp32-issue", "kernel": "sk_assign_kernel", "achieved": terms / (ms_assign * 1e-3) / 1e9, "peak": peak_terms / 1e9,
                         "unit": "G exp terms/s", "frac": terms / (ms_assign * 1e-3) / peak_terms,
                         "note": f"peak = measured FADD issue rate {fadd.value:.1f} T lane-ops/s / {FMA_PIPE_PER_TERM} FMA-pipe instructions per exp term"},
            "launch": {"RBP_SK_WARPS": os.environ.get("RBP_SK_WARPS", "8"), "RBP_SK_BLOCKS_PER_SM": os.environ.get("RBP_SK_BLOCKS_PER_SM", "3")}}
    if args.cpu_pairs:
        from oracle import binding as oracle

        rng = np.random.default_rng(1)
        ia, ib = rng.integers(0, args.n, args.cpu_pairs), rng.integers(0, args.k, args.cpu_pairs)
        a = pts[ia].astype(np.uint32)
        b = np.minimum(counts[ib], 0xFFFFFFFF).astype(np.uint32)
        t0 = time.perf_counter()
        want = oracle.sinkhorn_divergence_batch(a, b, tri, math=0, threads=os.cpu_count() or 1)
        dt = time.perf_counter() - t0
        got = rbp.lloyd.sinkhorn_divergence(a, b, np.arange(args.cpu_pairs), np.arange(args.cpu_pairs), tri)
        line["cpu_baseline"] = {"value": args.cpu_pairs / dt, "unit": "divergences/s (3 OT solves each: cross + both self terms)", "cores": os.cpu_count(),
                                "kind": "port", "sample": f"{args.cpu_pairs} point-centroid pairs",
                                "bit_identical_to_gpu": bool(np.array_equal(np.asarray(got).view(np.uint32), np.asarray(want).view(np.uint32)))}
    print(json.dumps(line), flush=True)
    g.close()


if __name__ == "__main__":
    main()
