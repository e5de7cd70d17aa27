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


def test_reseeding_after_bounds_sees_the_points_in_input_order(rbp, oracle):
    # init_bounds stores the points sorted by their first assignment (coherent warps in the Elkan step); k-means++ draws by a prefix
    # sum over the points in INPUT order, so a second initialisation on the same layer must undo that storage order first
    pts = turn_histograms(6000, seed=11)
    g = rbp.lloyd.Layer(pts, 24)
    o = oracle.OracleKmeans(pts, 24, threads=8)
    g.init_centroids(5); o.init_centroids(5)
    g.init_bounds(); o.init_bounds()
    for _ in range(2):
        g.step(); o.step()
    assert np.array_equal(g.lookup(), o.lookup())                                    # per-point output, input order
    assert np.array_equal(g.init_centroids(6), o.init_centroids(6))                  # re-seeded on the reordered layer
    g.init_bounds(); o.init_bounds()                                                 # second reorder composes with nothing stale
    for it in range(3):
        s = g.step()
        drift, sizes, re = o.step()
        assert f32eq(s.drift, drift) and np.array_equal(s.sizes, sizes) and s.reassignment == re, it
    ga, gu, gl, gs = g.bounds(True)
    oa, ou, ol, os_ = o.bounds(True)
    assert np.array_equal(ga, oa) and f32eq(gu, ou) and f32eq(gl, ol) and np.array_equal(gs, os_)


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


def test_two_point_shards_on_one_gpu_equal_single_layer(rbp):
    # point-sharded multi-GPU step emulated in one process: two handles own halves of the points; the integer
    # accumulators and tallies are summed across the handles (what NCCL all_reduce does), then both finish the step
    import torch

    from robopoker_b200.distributed import _DeviceWords

    n, k = 6000, 24
    pts = turn_histograms(n, seed=6)
    one = rbp.lloyd.Layer(pts, k)
    chosen = one.init_centroids(2)
    one.init_bounds()
    shards = [rbp.lloyd.Layer(pts[: n // 2], k), rbp.lloyd.Layer(pts[n // 2:], k)]
    for s in shards:
        s.set_centroids(pts[chosen].astype(np.uint64))
        s.init_bounds()
    for it in range(4):
        ref = one.step()
        views = []
        for s in shards:
            s.step_local()
            acc_ptr, acc_bytes, sizes_ptr, re_ptr, _ = s.exchange_buffers()
            views.append((torch.as_tensor(_DeviceWords(acc_ptr, acc_bytes, "<i8", 8), device="cuda"),
                          torch.as_tensor(_DeviceWords(sizes_ptr, 4 * k, "<i4", 4), device="cuda"),
                          torch.as_tensor(_DeviceWords(re_ptr, 4, "<i4", 4), device="cuda")))
        torch.cuda.synchronize()
        for f in range(3):
            total = views[0][f] + views[1][f]
            views[0][f].copy_(total)
            views[1][f].copy_(total)
        torch.cuda.synchronize()
        outs = [s.step_finish() for s in shards]
        for o in outs:
            assert f32eq(o.drift, ref.drift) and np.array_equal(o.sizes, ref.sizes) and o.reassignment == ref.reassignment, it
    assert np.array_equal(np.concatenate([s.bounds()[0] for s in shards]), one.bounds()[0])
    assert np.array_equal(shards[0].future()[0], one.future()[0])
