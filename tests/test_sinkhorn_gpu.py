"""GPU parity for the Sinkhorn path (flop layer): batched divergences and the full Elkan k-means, bit-exact against
the oracle under the shared exp/ln contract; plus the reference's own property tests on the device."""
import numpy as np
import pytest

from lloyd_data import flop_histograms, synthetic_metric
from test_oracle_sinkhorn import flop_metric, hist

pytestmark = pytest.mark.gpu


def f32eq(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def test_reference_property_tests_on_device(rbp):
    tri = flop_metric()
    h = hist([(0, 3), (5, 1), (12, 4), (24, 2)])[None]
    assert abs(rbp.lloyd.sinkhorn_divergence(h, h, [0], [0], tri)[0]) < 1e-4            # sinkhorn.rs divergence_is_zero_on_self
    mu, nu = hist([(0, 3), (5, 1), (12, 4)])[None], hist([(2, 2), (8, 5), (20, 1), (24, 3)])[None]
    d12 = rbp.lloyd.sinkhorn_divergence(mu, nu, [0], [0], tri)[0]
    d21 = rbp.lloyd.sinkhorn_divergence(nu, mu, [0], [0], tri)[0]
    assert abs(d12 - d21) < 1e-3 and d12 > 0                                            # divergence_is_symmetric


@pytest.mark.parametrize("bins,seed", [(32, 0), (256, 1)])
def test_batched_divergence_bit_exact(rbp, oracle, bins, seed):
    tri = synthetic_metric(bins, seed)
    a = flop_histograms(96, bins, seed=seed).astype(np.uint32)
    b = flop_histograms(64, bins, seed=seed + 10).astype(np.uint32)
    b[:8] = b[:8] * 1000 + flop_histograms(8, bins, seed=99, draws=400, spread=40.0)    # wide, centroid-like supports
    rng = np.random.default_rng(seed)
    ia, ib = rng.integers(0, 96, 400), rng.integers(0, 64, 400)
    got = rbp.lloyd.sinkhorn_divergence(a, b, ia, ib, tri)
    want = oracle.sinkhorn_divergence_batch(a[ia], b[ib], tri, math=0)
    assert f32eq(got, want)
    # and the contract stays within 2e-5 of the literal libm restatement
    libm = oracle.sinkhorn_divergence_batch(a[ia[:64]], b[ib[:64]], tri, math=1)
    assert np.max(np.abs(got[:64] - libm)) < 2e-5


@pytest.mark.parametrize("n,k,bins,seed", [(400, 5, 32, 0), (1500, 12, 64, 1), (600, 8, 256, 2)])
def test_flop_kmeans_bit_exact(rbp, oracle, n, k, bins, seed):
    pts = flop_histograms(n, bins, seed=seed)
    tri = synthetic_metric(bins, seed)
    g = rbp.lloyd.Layer(pts, k, metric=tri)
    o = oracle.OracleKmeans(pts, k, threads=8)
    o.set_metric(tri)
    assert np.array_equal(g.init_centroids(seed), o.init_centroids(seed))
    g.init_bounds()
    o.init_bounds()
    ga, gu, _, _ = g.bounds()
    oa, ou, _, _ = o.bounds()
    assert np.array_equal(ga, oa) and f32eq(gu, ou)
    for it in range(4):
        s = g.step()
        drift, sizes, re = o.step()
        assert f32eq(s.drift, drift), it
        assert np.array_equal(s.sizes, sizes) and s.reassignment == re, it
        assert np.array_equal(g.future()[0], o.future()[0]), it
    ga, gu, gl, gs = g.bounds(True)
    oa, ou, ol, os_ = o.bounds(True)
    assert np.array_equal(ga, oa) and f32eq(gu, ou) and f32eq(gl, ol) and np.array_equal(gs, os_)
    a, d = g.lookup(with_distance=True)
    oa2, od = o.lookup(with_distance=True)
    assert np.array_equal(a, oa2) and f32eq(d, od)       # bucket assignments, bit-exact
    assert f32eq(g.metric(), o.metric())


@pytest.mark.parametrize("alpha,members", [(0.02, 40), (0.3, 12), (0.05, 3)])
def test_mixture_points_against_merged_centroids_bit_exact(rbp, oracle, alpha, members):
    # the shapes the flop layer meets: narrow points (both half-sweep schedules, 4/8-wide groups and
    # masked tails) against centroids that are integer merges of member points
    from lloyd_data import flop_mixture_histograms

    bins = 256
    tri = synthetic_metric(bins, 3)
    pts = flop_mixture_histograms(48 * members, bins, comps=48, alpha=alpha, seed=7)
    cen = pts.reshape(48, members, bins).astype(np.uint32).sum(axis=1)
    rng = np.random.default_rng(11)
    ia, ib = rng.integers(0, len(pts), 160), rng.integers(0, 48, 160)
    a = pts.astype(np.uint32)
    got_pc = rbp.lloyd.sinkhorn_divergence(a, cen, ia, ib, tri)          # distance(x, c)
    got_cp = rbp.lloyd.sinkhorn_divergence(cen, a, ib, ia, tri)          # distance(c, x): not symmetric in f32
    assert f32eq(got_pc, oracle.sinkhorn_divergence_batch(a[ia], cen[ib], tri, math=0, threads=8))
    assert f32eq(got_cp, oracle.sinkhorn_divergence_batch(cen[ib], a[ia], tri, math=0, threads=8))


def test_sinkhorn_counters_and_metric_validation(rbp):
    from lloyd_data import flop_mixture_histograms

    pts = flop_mixture_histograms(600, 64, comps=8, alpha=0.1, seed=2)
    tri = synthetic_metric(64, 4)
    g = rbp.lloyd.Layer(pts, 8, metric=tri)
    g.init_centroids(0)
    g.sinkhorn_stats(reset=True)
    g.lookup()
    solves, sweeps, terms = g.sinkhorn_stats(reset=True)
    assert solves == 600 * 8 and solves <= sweeps <= 128 * solves and terms > sweeps    # one OT solve per (point, centroid)
    assert g.sinkhorn_stats() == (0, 0, 0)
    bad = tri.copy()
    bad[5] = -0.25
    with pytest.raises(rbp.RbpError):
        rbp.lloyd.Layer(pts, 8, metric=bad)
    with pytest.raises(rbp.RbpError):
        rbp.lloyd.sinkhorn_divergence(pts[:2], pts[:2], [0], [1], bad)


# ── tensor-core screen of the naive sweeps (csrc/sk_screen.cuh: tcgen05 + TMEM + TMA) ─────────────────────────────────────────
SCREEN_MARGIN = 2e-4  # >= 2 x the largest |approx - exact| the tests below allow (1e-4): then the screened sweep is exact


@pytest.mark.parametrize("n,k,alpha,bins", [(600, 200, 0.02, 256),   # two centroid tiles (128 + 72), point supports ~12 (one K step)
                                            (400, 64, 0.3, 256),     # one partial tile, supports ~34 (three K steps of the small operand)
                                            (300, 130, 0.1, 200)])   # fewer bins than the tile's 256
def test_screen_error_bound_and_exact_assignment(rbp, n, k, alpha, bins):
    from lloyd_data import flop_mixture_histograms
    pts = flop_mixture_histograms(n, bins, comps=k, alpha=alpha, seed=3)
    tri = synthetic_metric(bins, 3)
    g = rbp.lloyd.Layer(pts, k, metric=tri)
    g.init_centroids(3)
    g.init_bounds()
    g.step()                                                   # centroids become merged member sums (wide supports)
    counts, _ = g.future()
    approx, (problems, iters) = g.screen_probe()
    assert problems == n * ((k + 127) // 128) and iters > 0
    m = min(n, 96)
    ia, ib = np.repeat(np.arange(k), m).astype(np.int32), np.tile(np.arange(m), k).astype(np.int32)
    exact = rbp.lloyd.sinkhorn_divergence(counts.astype(np.uint32), pts[:m].astype(np.uint32), ia, ib, tri).reshape(k, m).T   # distance(c_j, x)
    assert np.all(np.isfinite(approx)) and np.max(np.abs(approx[:m] - exact)) < 1e-4
    # the screened sweep = the exact sweep, bit for bit, with two orders of magnitude fewer OT solves
    g.sinkhorn_stats(reset=True)
    a0, d0 = g.lookup(with_distance=True)
    full = g.sinkhorn_stats(reset=True)[0]
    g.screen(SCREEN_MARGIN)
    a1, d1 = g.lookup(with_distance=True)
    few = g.sinkhorn_stats(reset=True)[0]
    assert np.array_equal(a0, a1) and f32eq(d0, d1)
    assert full == n * k and n <= few < full // 20


def test_screened_init_bounds_equals_exact(rbp):
    from lloyd_data import flop_mixture_histograms
    pts = flop_mixture_histograms(500, 256, comps=150, alpha=0.05, seed=5)
    tri = synthetic_metric(256, 5)
    outs = []
    for margin in (-1.0, SCREEN_MARGIN):
        g = rbp.lloyd.Layer(pts, 150, metric=tri)
        g.init_centroids(5)
        g.screen(margin)
        g.init_bounds()                                        # elkan.rs:39-47: argmin with first-minimum ties, upper = that distance
        a, u, _, _ = g.bounds()
        steps = [g.step() for _ in range(2)]                   # and the Elkan steps that start from those bounds
        outs.append((a, u, steps[-1].drift, g.future()[0]))
    assert np.array_equal(outs[0][0], outs[1][0]) and f32eq(outs[0][1], outs[1][1])
    assert f32eq(outs[0][2], outs[1][2]) and np.array_equal(outs[0][3], outs[1][3])


def test_screen_refuses_the_w1_layer(rbp):
    from lloyd_data import turn_histograms
    g = rbp.lloyd.Layer(turn_histograms(200, seed=0), 8)
    with pytest.raises(rbp.RbpError):
        g.screen(1e-4)


def test_preflop_layer_is_identity_with_a_sinkhorn_metric(rbp, oracle):
    """`Layer::init_centroids` on the preflop street (crates/lloyd/src/layer.rs:151-154): N = K = 169, every point is its own centroid; the layer's
    outputs are the identity lookup and the 169 x 169 symmetrised Sinkhorn metric between the points (layer.rs:85-101)."""
    from lloyd_data import flop_mixture_histograms
    # preflop-shaped layer: 169 points over the 256 flop clusters.  (A point side holds at most 64 buckets — sized for the flop's 47 children;
    # real preflop histograms are wider and go through rbp_sinkhorn_batch, whose both sides take 256: test_batched_divergence_bit_exact.)
    pts = flop_mixture_histograms(169, 256, comps=40, alpha=0.3, seed=9, draws=60)
    assert (pts > 0).sum(axis=1).max() <= 64
    tri = synthetic_metric(256, 9)
    g = rbp.lloyd.Layer(pts, 169, metric=tri)
    g.set_centroids(pts.astype(np.uint64))
    o = oracle.OracleKmeans(pts, 169, threads=8)
    o.set_metric(tri)
    o.set_centroids_from_points(np.arange(169))
    a, d = g.lookup(with_distance=True)
    assert np.array_equal(a, np.arange(169)) and np.all(d == 0.0)      # divergence(x, x) = max(0, OT - OT/2 - OT/2) = 0, first minimum
    assert f32eq(g.metric(), o.metric())
    counts, weights = g.future()
    assert np.array_equal(counts, pts.astype(np.uint64)) and np.array_equal(weights, pts.sum(axis=1).astype(np.uint64))
