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


def test_profile_parity_with_the_large_epoch_walker_sort(rbp, oracle, monkeypatch):
    # epochs whose preorder arrays exceed L2 break size ties of the walker sort by tree index (32-bit keys); that path is
    # forced here at a size the oracle can follow, together with a split threshold that sends most roots through the
    # per-child tasks
    monkeypatch.setenv("RBP_NLHE_TIEBREAK", "1")
    monkeypatch.setenv("RBP_NLHE_SPLIT", "24")
    g, o = make(rbp, oracle, 160, 31)
    g.step(5), o.step(5)
    rows_equal(g.profile(), o.export())
    cg, co = g.counters(), o.counters()
    for k in ("epochs", "nodes", "infos", "updates", "rows"):
        assert cg[k] == co[k], (k, cg, co)


@pytest.mark.parametrize("batch,epochs", [(16384, 2), (65536, 1)])
def test_profile_parity_at_bench_sizes(rbp, oracle, batch, epochs):
    """The bench configuration itself (Flagship, seed 0, 16384 trees / epoch) and the 65536-tree epoch, where the size-dependent paths switch on by
    themselves — the tree tie-break of the walker sort, the per-child split of large roots, hot fold segments of thousands of Decisions: blueprint
    rows and telemetry bit-identical to the oracle."""
    import os
    from robopoker_b200.nlhe import Nlhe
    g = Nlhe(batch=batch, seed=0, table_slots=1 << 22)
    o = oracle.OracleNlhe(seed=0, batch=batch, threads=os.cpu_count() or 8)
    g.step(epochs)
    o.step(epochs)
    rows_equal(g.profile(), o.export())
    c, d = g.counters(), o.counters()
    assert (c["epochs"], c["nodes"], c["infos"], c["updates"], c["rows"]) == (d["epochs"], d["nodes"], d["infos"], d["updates"], d["rows"])


def test_waves_do_not_change_the_result(rbp, oracle, monkeypatch):
    """An epoch's trees are sampled in concurrent waves (two from 4096 trees on, four from 32768); forced to three at a small size here: same
    trees, same blueprint rows."""
    monkeypatch.setenv("RBP_NLHE_WAVES", "3")
    g, o = make(rbp, oracle, 100, 7)
    trees_equal(g, o, [0, 32, 33, 34, 66, 67, 99])
    g.step(4)
    o.step(4)
    rows_equal(g.profile(), o.export())
    assert g.counters()["nodes"] == o.counters()["nodes"] and g.counters()["updates"] == o.counters()["updates"]


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


def test_blueprint_edge_column_is_the_reference_u64(rbp, oracle):
    """`rbp_nlhe_export` writes `u64::from(Edge)` (kicker/src/edge.rs:185-197) in the edge column and `rbp_nlhe_import` reads it,
    including the legacy BBs form and rows in the reference's (unsorted, HashMap) order; foreign edges are rejected."""
    from robopoker_b200.nlhe import Nlhe
    g = Nlhe(batch=64, seed=3, table_slots=1 << 17)
    g.step(2)
    rows = g.profile()
    legal = {0, 1, 2, 3, 5, 22, 30, 38, 46, 8204, 6156, 4108, 6164, 8220, 2060, 8236, 4124, 2068, 2076}  # written out from edge.rs / pokerkit RAISES, OPENS
    assert set(int(v) for v in np.unique(rows["edge"])) <= legal and 1 in rows["edge"] and 5 in rows["edge"]
    for row in rows[:300]:
        assert oracle.nlhe_edge_from_u64(int(row["edge"])) in oracle.nlhe_unpath(int(row["choices"]))
    legacy = rows.copy()
    opens = (legacy["edge"] & 7) == 6
    assert opens.any()
    legacy["edge"][opens] = 4 | (legacy["edge"][opens] >> 3 & 0xFF) << 3 | 1 << 19   # old Size::BBs encoding (edge.rs:168-172)
    legacy = legacy[np.random.default_rng(0).permutation(len(legacy))]
    h = Nlhe(batch=64, seed=3, table_slots=1 << 17)
    h.load(legacy, 2)
    assert h.profile().tobytes() == rows.tobytes()
    bad = rows[:4].copy()
    bad["edge"][0] = 4 | 7 << 3 | 9 << 11  # Raise(7/9) is not on the Pluribus grid
    with pytest.raises(rbp.RbpError):
        h.load(bad, 2)


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


def _mix64(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _synthetic_bucket(pocket, pub, k):
    with np.errstate(over="ignore"):
        return (_mix64(pocket * np.uint64(0x9E3779B97F4A7C15) ^ _mix64(pub)) % np.uint64(k)).astype(np.uint8)


def test_installed_lookup_tables_equal_the_synthetic_function(rbp):
    """`rbp_nlhe_set_lookup` at the reference's full sizes (169 / 1,286,792 / 13,960,050 / 123,156,254 isomorphisms): tables
    filled with the synthetic bucket function must reproduce the table-free run bit for bit (every observation the sampler
    meets is found, canonicalisation agrees with the enumeration)."""
    from robopoker_b200.nlhe import Nlhe

    plain = Nlhe(batch=512, seed=31, table_slots=1 << 19)
    table = Nlhe(batch=512, seed=31, table_slots=1 << 19)
    for street, k in (("pref", 169), ("flop", 256), ("turn", 256), ("rive", 101)):
        isos = rbp.deuce.IsoSet(street)
        n = len(isos)
        abs_ = np.zeros(n, np.uint8)
        for off in range(0, n, 1 << 23):
            p, b = isos.export(off, min(1 << 23, n - off))
            abs_[off:off + len(p)] = _synthetic_bucket(p, b, k)
        isos.set_abstractions(abs_)
        if street == "flop":  # through the reference's `isomorphism` table rows (obs i64, abs i16) instead of the device handle
            obs, a16 = isos.export_rows()
            assert np.all(a16 >> 8 == 1) and np.array_equal(a16 & 0xFF, abs_)
            table.set_lookup_rows(obs, a16)
        else:
            table.set_lookup(isos)
        isos.close()
    plain.step(3), table.step(3)
    assert plain.profile().tobytes() == table.profile().tobytes()
    assert plain.counters()["updates"] == table.counters()["updates"] > 0


def test_lookup_tables_install_per_street(rbp):
    """Streets with a table use it, the others keep the synthetic lookup; an all-zero column collapses a street to one bucket."""
    from robopoker_b200.nlhe import Nlhe

    g = Nlhe(batch=64, seed=2, table_slots=1 << 16)
    for street in ("pref", "turn"):
        isos = rbp.deuce.IsoSet(street)
        isos.set_abstractions(np.zeros(len(isos), np.uint8))
        g.set_lookup(isos)
        isos.close()
    g.step(2)
    rows = g.profile()
    present = rows["present"].astype(np.int64)
    assert set(np.unique(present[present >> 8 == 0])) == {0} and set(np.unique(present[present >> 8 == 2])) == {2 << 8}
    assert len(np.unique(present[present >> 8 == 1])) > 50  # the flop still spreads over the synthetic buckets
