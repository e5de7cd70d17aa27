#!/usr/bin/env python
"""Contract-vs-libm ASSIGNMENT flip rate on a config-3 sample at 20 000 x 200.

The device computes the full N x K divergence matrix under the exp/ln contract (include/rbp.h).  An assignment can only differ under the
reference's libm where the best and second-best divergences are closer than twice the contract-vs-libm divergence difference (measured
<= 5.1e-7 on 100 000 pairs, profiles/r2_contract_vs_libm_flips_n2000_k50.json; bounded at 2e-5 in tests/test_sinkhorn_gpu.py).  So the
libm restatement (CPU oracle, math=1) is evaluated for every point whose gap is below `--window` (default 1e-4 = 200 x the measured
difference), against all of that point's centroids inside the window, and the argmins are compared.

    python tests/measure/libm_flip_rate_gpu.py --n 20000 --k 200 > profiles/r2_contract_vs_libm_flips_n20000_k200.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=20000)
    p.add_argument("--k", type=int, default=200)
    p.add_argument("--alpha", type=float, default=0.02)
    p.add_argument("--window", type=float, default=1e-4)
    a = p.parse_args()
    import numpy as np
    from lloyd_data import flop_mixture_histograms, synthetic_metric

    import robopoker_b200 as rbp
    from oracle import binding as oracle

    pts = flop_mixture_histograms(a.n, 256, comps=a.k, alpha=a.alpha, seed=0).astype(np.uint32)
    tri = synthetic_metric(256, 0)
    member = np.random.default_rng(1).integers(0, a.k, a.n)
    cen = np.zeros((a.k, 256), np.uint32)
    np.add.at(cen, member, pts)                                  # centroids = member sums (`Absorb`, elkan/src/absorb.rs:16-21)
    cen[cen.sum(axis=1) == 0] = pts[0]
    ia, ib = np.repeat(np.arange(a.k), a.n).astype(np.int32), np.tile(np.arange(a.n), a.k).astype(np.int32)
    d = rbp.lloyd.sinkhorn_divergence(cen, pts, ia, ib, tri).reshape(a.k, a.n).T   # distance(c_j, x) under the contract, on the device
    best = d.argmin(axis=1)
    srt = np.sort(d, axis=1)
    gap = srt[:, 1] - srt[:, 0]
    near = np.flatnonzero(gap < a.window)
    flips, pairs, maxdiff = 0, 0, 0.0
    for i in near:
        cand = np.flatnonzero(d[i] < srt[i, 0] + a.window)
        lib = oracle.sinkhorn_divergence_batch(cen[cand], np.repeat(pts[i:i + 1], len(cand), axis=0), tri, math=1, threads=os.cpu_count() or 8)
        pairs += len(cand)
        maxdiff = max(maxdiff, float(np.max(np.abs(lib - d[i, cand]))))
        flips += int(cand[int(np.argmin(lib))] != best[i])
    print(json.dumps({"n": a.n, "k": a.k, "alpha": a.alpha, "pairs": a.n * a.k, "window": a.window, "points_inside_window": int(len(near)),
                      "libm_pairs_evaluated": pairs, "assignment_flips": flips, "flip_rate": flips / a.n, "max_abs_divergence_difference_inside_window": maxdiff,
                      "median_best_to_second_gap": float(np.median(gap)), "smallest_gaps": [float(x) for x in np.sort(gap)[:8]]}))


if __name__ == "__main__":
    main()
