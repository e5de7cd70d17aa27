"""Pins the lloyd oracle to the properties the reference's own tests assert (no numeric golden vectors exist):
`Equity::variation` symmetric / zero on self (crates/lloyd/src/emd.rs:73-97), Elkan ≡ naive assignment
(crates/lloyd/src/tests.rs:149-160, K=8 N=2048 fixture shape), `Pair` triangular bijection (pair.rs:172-189)."""
import numpy as np

from lloyd_data import turn_histograms


def test_variation_properties(oracle):
    pts = turn_histograms(64, seed=1).astype(np.uint32)
    for i in range(0, 64, 7):
        assert oracle.variation(pts[i], pts[i]) == 0.0
        for j in range(1, 64, 11):
            a, b = oracle.variation(pts[i], pts[j]), oracle.variation(pts[j], pts[i])
            assert a == b and a >= 0.0
    # point masses at buckets 0 and 100: cdfs differ by 1 on 100 of the 101 buckets
    x, y = np.zeros(101, np.uint32), np.zeros(101, np.uint32)
    x[0], y[100] = 5, 9
    assert abs(oracle.variation(x, y) - 100.0 / 101.0) < 1e-6
    # triangle inequality (W1 is a metric)
    for t in range(20):
        a, b, c = pts[3 * t], pts[3 * t + 1], pts[3 * t + 2]
        assert oracle.variation(a, c) <= oracle.variation(a, b) + oracle.variation(b, c) + 1e-6


def test_pair_triangular_bijection():
    # Pair::merge / Pair::split (crates/lloyd/src/pair.rs:30-39)
    import math

    for k in (2, 8, 256):
        seen = set()
        for i in range(k):
            for j in range(i):
                t = i * (i - 1) // 2 + j
                jj = (math.isqrt(1 + 8 * t) + 1) // 2
                ii = t - jj * (jj - 1) // 2
                assert (ii, jj) == (j, i)
                seen.add(t)
        assert seen == set(range(k * (k - 1) // 2))


def test_elkan_equals_naive(oracle):
    # the reference's (ignored) elkan_naive_equivalence test, on the TestLayer shape K=8, N=2048:
    # after every Elkan step the tracked assignment equals a fresh naive argmin against the PRE-step centroids
    pts = turn_histograms(2048, seed=2)
    e = oracle.OracleKmeans(pts, 8, threads=4)
    e.init_centroids(seed=5)
    e.init_bounds()
    for it in range(6):
        naive = e.lookup()          # naive argmin against the centroids the step will use
        e.step()
        assert np.array_equal(e.bounds()[0], naive), it


def test_kmeans_converges(oracle):
    pts = turn_histograms(4096, seed=3)
    e = oracle.OracleKmeans(pts, 16, threads=4)
    chosen = e.init_centroids(seed=0)
    assert len(set(chosen.tolist())) == 16
    e.init_bounds()
    re = [e.step()[2] for _ in range(12)]
    assert re[0] == 0 and re[-1] < re[1]          # first step: centroids are points, nothing moves
    counts, weights = e.future()
    assert int(weights.sum()) == 46 * 4096 and (counts.sum(axis=1) == weights).all()
    tri = e.metric()
    assert abs(tri.max() - 1.0) < 1e-6 and tri.min() > 0.0
