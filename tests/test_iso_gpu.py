"""GPU parity for isomorphism enumeration, canonicalisation and projections: bit-exact against the oracle, and the
reference's exhaustive counts (crates/deuce/src/street.rs:129-135) on the device — including the 123,156,254 river
isomorphisms the reference only checks in an ignored test."""
import numpy as np
import pytest

from test_oracle_deuce import random_hands

pytestmark = pytest.mark.gpu


def test_counts_and_order_pref_flop(rbp, oracle):
    assert len(rbp.deuce.IsoSet("pref")) == 169
    flop = rbp.deuce.IsoSet("flop")
    n, op, ob = oracle.isomorphisms("flop")
    assert len(flop) == n == 1_286_792
    gp, gb = flop.export()
    assert np.array_equal(gp, op) and np.array_equal(gb, ob)       # same set, same enumeration order


def test_turn_and_river_counts(rbp, oracle):
    turn = rbp.deuce.IsoSet("turn")
    assert len(turn) == 13_960_050
    n, op, ob = oracle.isomorphisms("turn", cap=200_000)
    gp, gb = turn.export(0, 200_000)
    assert np.array_equal(gp, op) and np.array_equal(gb, ob)
    river = rbp.deuce.IsoSet("rive")
    assert len(river) == 123_156_254
    p, b = river.export(123_000_000, 1000)
    key = list(zip(p.tolist(), b.tolist()))
    assert key == sorted(key)                                      # sorted by (pocket, public)
    cp, cb, fl = rbp.deuce.canonical(p, b)
    assert fl.all() and np.array_equal(cp, p) and np.array_equal(cb, b)


def test_canonical_batch(rbp, oracle):
    rng = np.random.default_rng(3)
    for k in (5, 6, 7):
        hands = random_hands(rng, 50_000, k)
        pocket = np.array([int(h) & -int(h) | (int(h) & (int(h) - 1)) & -(int(h) & (int(h) - 1)) for h in hands], dtype=np.uint64)
        public = hands & ~pocket
        g = rbp.deuce.canonical(pocket, public)
        o = oracle.canonical_batch(pocket, public)
        assert all(np.array_equal(x, y) for x, y in zip(g, o))


def test_turn_projection_through_river_lookup(rbp, oracle):
    # Lookup::grow(Rive) then Lookup::projections for the turn layer: histograms over river-equity buckets
    river = rbp.deuce.IsoSet("rive")
    river.river_buckets()
    turn = rbp.deuce.IsoSet("turn")
    for offset in (0, 7_000_000, 13_959_000):
        hist, misses = turn.project(river, 101, offset, 600)
        assert misses == 0
        p, b = turn.export(offset, 600)
        want = oracle.turn_histograms(p, b, threads=8)             # children → equity → bucket, no lookup table
        assert np.array_equal(hist, want)
        assert (hist.sum(axis=1) == 46).all()
    # spot-check the river column itself
    p, b, a = river.export(55_555_000, 2000, with_abs=True)
    assert np.array_equal(a, oracle.river_equity_batch(p, b)[1])


def test_flop_projection_through_a_turn_lookup(rbp, oracle):
    # a synthetic turn lookup (bucket = hash of the iso) projected onto flop observations, vs the oracle's BTreeMap-style path
    turn = rbp.deuce.IsoSet("turn")
    tp, tb = turn.export()
    abs_ = ((tp * np.uint64(0x9E3779B97F4A7C15) ^ tb * np.uint64(0xC2B2AE3D27D4EB4F)) >> np.uint64(56)).astype(np.uint8)
    turn.set_abstractions(abs_)
    flop = rbp.deuce.IsoSet("flop")
    hist, misses = flop.project(turn, 256, 640_000, 300)
    assert misses == 0 and (hist.sum(axis=1) == 47).all()
    p, b = flop.export(640_000, 300)
    want = oracle.project(p, b, tp, tb, abs_, 256, threads=8)
    assert np.array_equal(hist, want)
