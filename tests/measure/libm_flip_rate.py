#!/usr/bin/env python
"""How often does the exp/ln CONTRACT (include/rbp.h: exp_c / ln_c, fixed IEEE operation sequences) pick a different bucket than the
reference's platform libm (`f32::exp` / `f32::ln`, crates/lloyd/src/sinkhorn.rs:115-136) would?  CPU only (oracle, both maths).

A config-3 sample: `--n` flop-style points x `--k` centroids (member sums of a random partition of the points), `Layer::lookup`
semantics (argmin over distance(c_j, x), first minimum).  Prints one JSON line: the number of points whose argmin differs, the
largest divergence difference, and the gap between best and second best at the flipped points.

    python tests/measure/libm_flip_rate.py --n 20000 --k 200 > profiles/r2_contract_vs_libm_flips.json
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
    p.add_argument("--n", type=int, default=20000)
    p.add_argument("--k", type=int, default=200)
    p.add_argument("--alpha", type=float, default=0.02)
    p.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    a = p.parse_args()
    import numpy as np
    from lloyd_data import flop_mixture_histograms, synthetic_metric

    from oracle import binding as oracle

    pts = flop_mixture_histograms(a.n, 256, comps=a.k, alpha=a.alpha, seed=0).astype(np.uint32)
    tri = synthetic_metric(256, 0)
    rng = np.random.default_rng(1)
    member = rng.integers(0, a.k, a.n)
    cen = np.zeros((a.k, 256), np.uint32)
    np.add.at(cen, member, pts)                      # centroids = member sums (`Absorb`, elkan/src/absorb.rs:16-21)
    cen[cen.sum(axis=1) == 0] = pts[0]
    t0 = time.perf_counter()
    d = {}
    for name, math in (("contract", 0), ("libm", 1)):
        out = np.zeros((a.n, a.k), np.float32)
        for j in range(a.k):                         # distance(c_j, x): mu = centroid
            out[:, j] = oracle.sinkhorn_divergence_batch(np.repeat(cen[j:j + 1], a.n, axis=0), pts, tri, math=math, threads=a.threads)
        d[name] = out
    arg = {k: v.argmin(axis=1) for k, v in d.items()}
    flips = np.flatnonzero(arg["contract"] != arg["libm"])
    srt = np.sort(d["libm"], axis=1)
    gap = srt[:, 1] - srt[:, 0]
    print(json.dumps({"n": a.n, "k": a.k, "alpha": a.alpha, "pairs": a.n * a.k, "seconds": time.perf_counter() - t0,
                      "assignment_flips": int(len(flips)), "flip_rate": len(flips) / a.n,
                      "max_abs_divergence_difference": float(np.max(np.abs(d["contract"] - d["libm"]))),
                      "mean_abs_divergence_difference": float(np.mean(np.abs(d["contract"] - d["libm"]))),
                      "best_to_second_gap_at_flips": [float(x) for x in gap[flips][:20]],
                      "median_best_to_second_gap": float(np.median(gap)),
                      "points_with_gap_below_2e-5": int((gap < 2e-5).sum())}))


if __name__ == "__main__":
    main()
