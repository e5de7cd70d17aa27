"""Pins the isomorphism oracle to the reference's exhaustive counts (crates/deuce/src/street.rs:129-135; the ignored
tests of isomorphism_iter.rs:42-70 assert exactly these) and its invariance tests (isomorphism.rs tests)."""
import itertools

import numpy as np

from test_oracle_deuce import random_hands


def test_isomorphism_counts(oracle):
    assert oracle.isomorphisms("pref", cap=0)[0] == 169
    n, pocket, public = oracle.isomorphisms("flop")
    assert n == 1_286_792
    # enumeration order is ascending in (pocket, public): the lookup tables rely on it
    key = list(zip(pocket[:50000].tolist(), public[:50000].tolist()))
    assert key == sorted(key)
    assert all(bin(p).count("1") == 2 for p in pocket[:1000].tolist()) and all(bin(b).count("1") == 3 for b in public[:1000].tolist())


def test_turn_count(oracle):
    assert oracle.isomorphisms("turn", cap=0, threads=8)[0] == 13_960_050


def test_canonical_is_invariant_under_all_24_suit_permutations(oracle):
    # isomorphism.rs tests: super_symmetry / false_positives
    rng = np.random.default_rng(0)
    seven = random_hands(rng, 200, 7)
    pocket = np.array([int(h) & -int(h) | (int(h) & (int(h) - 1)) & -(int(h) & (int(h) - 1)) for h in seven], dtype=np.uint64)
    public = seven & ~pocket
    cp, cb, _ = oracle.canonical_batch(pocket, public)

    def relabel(x, perm):
        out = np.zeros_like(x)
        for s in range(4):
            cards = x & np.uint64(0x0001111111111111 << s)
            d = perm[s] - s
            out |= (cards << np.uint64(d)) if d >= 0 else (cards >> np.uint64(-d))
        return out

    for perm in itertools.permutations(range(4)):
        p2, b2, _ = oracle.canonical_batch(relabel(pocket, perm), relabel(public, perm))
        assert np.array_equal(p2, cp) and np.array_equal(b2, cb)
    # canonical forms are fixed points and flagged canonical
    p3, b3, ic = oracle.canonical_batch(cp, cb)
    assert np.array_equal(p3, cp) and np.array_equal(b3, cb) and ic.all()


def test_turn_histogram_shape(oracle):
    n, pocket, public = oracle.isomorphisms("turn", cap=64)
    h = oracle.turn_histograms(pocket, public, threads=8)
    assert (h.sum(axis=1) == 46).all()       # 46 river children per turn observation (street.rs:120-126)


def test_observation_i64_encoding_roundtrip():
    """`i64::from(Observation)` / `Observation::from(i64)` (crates/deuce/src/observation.rs:130-163): board cards, then pocket
    cards, ascending, one byte (1 + card) each, first card highest.  Host-side format conversion of librbp_b200."""
    from robopoker_b200 import deuce

    def card(s):
        return "23456789TJQKA".index(s[0]) * 4 + "cdhs".index(s[1])

    pocket = np.array([1 << card("2c") | 1 << card("Ts")], np.uint64)
    public = np.array([1 << card("Jc") | 1 << card("Js") | 1 << card("3d")], np.uint64)
    want = 0
    for c in ("3d", "Jc", "Js", "2c", "Ts"):
        want = want << 8 | (card(c) + 1)
    assert int(deuce.obs_encode(pocket, public)[0]) == want
    rng = np.random.default_rng(5)
    ps, bs = [], []
    for _ in range(2000):
        cards = rng.choice(52, size=2 + int(rng.choice([0, 3, 4, 5])), replace=False)
        ps.append(sum(1 << int(c) for c in cards[:2]))
        bs.append(sum(1 << int(c) for c in cards[2:]))
    ps, bs = np.array(ps, np.uint64), np.array(bs, np.uint64)
    obs = deuce.obs_encode(ps, bs)
    p2, b2 = deuce.obs_decode(obs)
    assert np.array_equal(p2, ps) and np.array_equal(b2, bs) and len(np.unique(obs)) == len(set(zip(ps.tolist(), bs.tolist())))
