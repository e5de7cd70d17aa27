"""GPU parity for the turn-layer k-means (W1 / Elkan): bit-exact against the oracle at every step."""
import numpy as np
import pytest

from lloyd_data import turn_histograms

pytestmark = pytest.mark.gpu


def f32eq(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("n,k,seed", [(2048, 8, 0), (5000, 37, 1), (20000, 256, 2), (3000, 600, 3)])
def test_full_pipeline_bit_exact(rbp, oracle, n, k, seed):
    pts = turn_histograms(n, seed=seed)
    g = rbp.lloyd.Layer(pts, k)
    o = oracle.OracleKmeans(pts, k, threads=8)
    assert np.array_equal(g.init_centroids(seed), o.init_centroids(seed))          # k-means++ picks
    assert np.array_equal(g.future()[0], o.future()[0])
    g.init_bounds()
    o.init_bounds()
    ga, gu, _, gs = g.bounds()
    oa, ou, _, os_ = o.bounds()
    assert np.array_equal(ga, oa) and f32eq(gu, ou) and np.array_equal(gs, os_)
    for it in range(6):
        s = g.step()
        drift, sizes, re = o.step()
        assert f32eq(s.drift, drift), it
        assert np.array_equal(s.sizes, sizes) and s.reassignment == re, it
        gc, gw = g.future()
        oc, ow = o.future()
        assert np.array_equal(gc, oc) and np.array_equal(gw, ow), it
    with_lower = n * k <= 2_000_000
    ga, gu, gl, gs = g.bounds(with_lower)
    oa, ou, ol, os_ = o.bounds(with_lower)
    assert np.array_equal(ga, oa) and f32eq(gu, ou) and np.array_equal(gs, os_)
    if with_lower:
        assert f32eq(gl, ol)
    a, d = g.lookup(with_distance=True)
    oa2, od = o.lookup(with_distance=True)
    assert np.array_equal(a, oa2) and f32eq(d, od)                                   # bucket assignments
    assert f32eq(g.metric(), o.metric())


def test_explicit_centroids_and_tiny_k(rbp, oracle):
    pts = turn_histograms(300, seed=9)
    g = rbp.lloyd.Layer(pts, 1)
    g.set_centroids(pts[:1].astype(np.uint64))
    g.init_bounds()
    s = g.step()
    assert s.sizes.tolist() == [300] and s.reassignment == 0
    assert (g.lookup() == 0).all() and len(g.metric()) == 0


def test_properties_at_scale(rbp):
    # 1M-point subsample of config 5, K=256: size-independent invariants
    n, k = 1_000_000, 256
    pts = turn_histograms(n, seed=4)
    g = rbp.lloyd.Layer(pts, k)
    g.init_centroids(0)
    g.init_bounds()
    steps = [g.step() for _ in range(4)]
    for s in steps:
        assert int(s.sizes.sum()) == n
    assert steps[0].reassignment == 0
    counts, weights = g.future()
    assert int(weights.sum()) == int(pts.sum())                 # every draw lands in exactly one centroid
    assert np.array_equal(counts.sum(axis=0), pts.sum(axis=0, dtype=np.uint64))
    a1 = g.lookup()
    a2 = g.lookup()
    assert np.array_equal(a1, a2) and a1.max() < k              # idempotent, in range
