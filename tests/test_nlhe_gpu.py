"""GPU parity: heads-up NLHE MCCFR (robopoker_b200/csrc/nlhe.cu through the C ABI) against the oracle (oracle/nlhe.hpp) on identical
Philox streams.  Bit-exact bar: sampled trees (policy / sampling probabilities / payoffs per node), the blueprint rows
(past, present, choices, edge, weight, regret, payoff, visits) and the telemetry counters must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make(rbp, oracle, batch, seed, regret="LinearRegret", weight="LinearWeight", sampling="PluribusSampling", warmup=None,
         threshold=None, slots=1 << 18):
    from robopoker_b200.nlhe import Nlhe
    hyper = rbp.Hyper(**({"prune_warmup": warmup} if warmup is not None else {}), **({"prune_threshold": threshold} if threshold is not None else {}))
    g = Nlhe(regret=regret, weight=weight, sampling=sampling, batch=batch, seed=seed, hyper=hyper, table_slots=slots)
    hv = [hyper.c.temperature, hyper.c.smoothing, hyper.c.curiosity, hyper.c.prune_threshold, hyper.c.prune_explore]
    o = oracle.OracleNlhe(seed=seed, batch=batch, threads=8, regret=regret, weight=weight, sampling=sampling, hyper=hv,
                          warmup=hyper.c.prune_warmup)
    return g, o


def rows_equal(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for f in ("past", "choices", "edge", "present"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["visits"].astype(np.int64), b["visits"].astype(np.int64))
    for f in ("regret", "weight", "payoff"):
        x, y = a[f].view(np.uint32), b[f].view(np.uint32)
        bad = np.nonzero(x != y)[0]
        assert bad.size == 0, (f, bad.size, a[bad[:3]], b[bad[:3]])


def trees_equal(g, o, trees):
    for t in trees:
        a, b = g.debug_tree(t), o.tree_preorder(t)
        assert len(a) == len(b), (t, len(a), len(b))
        for f in ("depth", "kind", "act"):
            assert np.array_equal(a[f], b[f]), (t, f)
        for f in ("p", "q", "payoff"):
            assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), (t, f, a[f][:8], b[f][:8])


def test_first_epoch_trees(rbp, oracle):
    g, o = make(rbp, oracle, 64, 5)
    trees_equal(g, o, range(64))


CASES = [
    # regret, weight, sampling, batch, epochs, warmup, threshold
    ("LinearRegret", "LinearWeight", "PluribusSampling", 128, 6, None, None),       # Flagship (nlhe/src/lib.rs:86-90), warm-up phase
    ("LinearRegret", "LinearWeight", "PluribusSampling", 96, 8, 2, 20.0),           # pruning active: raise/shove edges start below 20
    ("LinearRegret", "LinearWeight", "PrunableSampling", 64, 6, None, 20.0),
    ("FlooredRegret", "LinearWeight", "ExternalSampling", 200, 5, None, None),
    ("DiscountedRegret", "QuadraticWeight", "ExternalSampling", 64, 6, None, None),
    ("AsymmetricRegret", "ExponentialWeight", "ExternalSampling", 33, 6, None, None),
    ("SummedRegret", "ConstantWeight", "ExternalSampling", 1, 40, None, None),
]


@pytest.mark.parametrize("regret,weight,sampling,batch,epochs,warmup,threshold", CASES)
def test_profile_parity(rbp, oracle, regret, weight, sampling, batch, epochs, warmup, threshold):
    g, o = make(rbp, oracle, batch, 77, regret, weight, sampling, warmup, threshold)
    g.step(epochs), o.step(epochs)
    rows_equal(g.profile(), o.export())
    cg, co = g.counters(), o.counters()
    for k in ("epochs", "nodes", "infos", "updates", "rows"):
        assert cg[k] == co[k], (k, cg, co)
    trees_equal(g, o, range(0, batch, max(1, batch // 7)))  # trees of the next epoch read the trained table


def test_export_import_roundtrip(rbp, oracle):
    g, o = make(rbp, oracle, 64, 3)
    g.step(3), o.step(5)
    rows = g.profile()
    from robopoker_b200.nlhe import Nlhe
    h = Nlhe(batch=64, seed=3, table_slots=1 << 17)
    h.load(rows, 3)
    assert h.profile().tobytes() == rows.tobytes()
    h.step(2)
    rows_equal(h.profile(), o.export())


def test_capacity_errors_are_loud(rbp):
    from robopoker_b200.nlhe import Nlhe
    g = Nlhe(batch=256, seed=1, table_slots=1 << 10)
    with pytest.raises(rbp.RbpError) as e:
        g.step(3)
    assert e.value.status == -4
    g = Nlhe(batch=64, seed=1, max_nodes_per_tree=64)
    with pytest.raises(rbp.RbpError):
        g.step(1)


def test_large_batch_properties(rbp):
    from robopoker_b200.nlhe import Nlhe
    g = Nlhe(batch=8192, seed=9, table_slots=1 << 21)
    g.step(2)
    c = g.counters()
    rows = g.profile()
    assert c["epochs"] == 2 and c["updates"] > 8192 * 2 * 50
    assert int(rows["visits"].astype(np.int64).sum()) >= c["infos"]  # every Decisions visits each edge of its row once
    assert np.all(np.isfinite(rows["regret"])) and np.all(rows["weight"] >= 0)
    keys = np.stack([rows["past"], rows["present"].astype(np.int64), rows["choices"]], axis=1)
    assert len(np.unique(keys, axis=0)) == c["rows"]

