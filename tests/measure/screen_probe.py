#!/usr/bin/env python
"""Error and speed of the tensor-core screen (csrc/sk_screen.cuh) against the exact log-domain solver on a flop-style layer.

    python tests/measure/screen_probe.py --n 4000 --k 200 --alpha 0.02
Prints one JSON line: max / mean |approx - exact| over m x k pairs, how many centroids per point survive a margin, the time of the
screened and of the full exact assignment sweep, and whether the assignments / winning distances are bit-identical.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=4000)
    p.add_argument("--k", type=int, default=200)
    p.add_argument("--bins", type=int, default=256)
    p.add_argument("--alpha", type=float, default=0.02)
    p.add_argument("--pairs", type=int, default=256, help="points whose k exact divergences are compared with the screen's")
    p.add_argument("--margin", type=float, default=None, help="default: 4 x the measured max error")
    p.add_argument("--steps", type=int, default=1, help="Elkan steps before the comparison (centroids become merged member sums)")
    args = p.parse_args()
    import numpy as np
    from lloyd_data import flop_mixture_histograms, synthetic_metric

    import robopoker_b200 as rbp
    from robopoker_b200.lloyd import sinkhorn_divergence

    pts = flop_mixture_histograms(args.n, args.bins, comps=args.k, alpha=args.alpha, seed=0)
    tri = synthetic_metric(args.bins, 0)
    g = rbp.lloyd.Layer(pts, args.k, metric=tri)
    g.init_centroids(0)
    g.init_bounds()
    for _ in range(args.steps):
        g.step()
    counts, _ = g.future()
    m = min(args.pairs, args.n)
    print('layer ready', file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    approx, (problems, iters) = g.screen_probe()
    t_screen = time.perf_counter() - t0
    print('screen done', t_screen, problems, iters, file=sys.stderr, flush=True)
    ia = np.repeat(np.arange(args.k), m).astype(np.int32)            # distance(c_j, x): mu = centroid
    ib = np.tile(np.arange(m), args.k).astype(np.int32)
    exact = sinkhorn_divergence(counts.astype(np.uint32), pts[:m].astype(np.uint32), ia, ib, tri).reshape(args.k, m).T
    err = np.abs(approx[:m] - exact)
    margin = args.margin if args.margin is not None else 4.0 * float(err.max())
    amin = approx.min(axis=1, keepdims=True)
    survivors = (approx <= amin + margin).sum(axis=1)
    g.screen(-1.0)
    g.sinkhorn_stats(reset=True)
    t0 = time.perf_counter(); a0, d0 = g.lookup(with_distance=True); t_exact = time.perf_counter() - t0
    s_exact = g.sinkhorn_stats(reset=True)[0]
    g.screen(margin)
    t0 = time.perf_counter(); a1, d1 = g.lookup(with_distance=True); t_scr = time.perf_counter() - t0
    s_scr = g.sinkhorn_stats(reset=True)[0]
    print(json.dumps({"n": args.n, "k": args.k, "alpha": args.alpha, "mean_support": float((pts > 0).sum(axis=1).mean()),
                      "max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()), "exact_mean": float(exact.mean()),
                      "worst_pairs": [[int(i), int(j), float(approx[i, j]), float(exact[i, j])] for i, j in zip(*np.unravel_index(np.argsort(err, axis=None)[-3:], err.shape))],
                      "margin": margin, "survivors_mean": float(survivors.mean()), "survivors_max": int(survivors.max()),
                      "screen_s": t_screen, "screen_pairs_per_s": args.n * args.k / t_screen, "iterations_per_problem": iters / max(problems, 1) / 128.0,
                      "exact_sweep_s": t_exact, "screened_sweep_s": t_scr, "speedup": t_exact / t_scr, "exact_solves": s_exact, "screened_solves": s_scr,
                      "assignments_identical": bool(np.array_equal(a0, a1)), "distances_identical": bool(np.array_equal(d0.view(np.uint32), d1.view(np.uint32)))}))


if __name__ == "__main__":
    main()
