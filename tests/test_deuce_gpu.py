"""GPU parity for hand strength and river equity: bit-exact against the oracle and the reference's known answers."""
import json
import os

import numpy as np
import pytest

from test_oracle_deuce import GOLD, random_hands

pytestmark = pytest.mark.gpu


def test_known_answers_gpu(rbp):
    hands = np.array([c["bits"] for c in GOLD["cases"]], dtype=np.uint64)
    got = rbp.deuce.strength(hands)
    assert [int(g) for g in got] == [c["packed"] for c in GOLD["cases"]]


def test_all_five_card_hands(rbp, oracle):
    from itertools import combinations

    bits = np.fromiter((sum(1 << c for c in comb) for comb in combinations(range(52), 5)), dtype=np.uint64, count=2598960)
    assert np.array_equal(rbp.deuce.strength(bits), oracle.eval_batch(bits))


@pytest.mark.parametrize("k", [5, 6, 7])
def test_random_hands(rbp, oracle, k):
    hands = random_hands(np.random.default_rng(k), 200000, k)
    assert np.array_equal(rbp.deuce.strength(hands), oracle.eval_batch(hands))


def test_empty_batch(rbp):
    assert len(rbp.deuce.strength(np.zeros(0, np.uint64))) == 0


def test_river_equity(rbp, oracle):
    rng = np.random.default_rng(7)
    seven = random_hands(rng, 4096, 7)
    pockets = np.zeros_like(seven)
    for i, h in enumerate(seven):
        h = int(h)
        lo = h & -h
        h2 = h & (h - 1)
        pockets[i] = lo | (h2 & -h2)
    public = seven & ~pockets
    g = rbp.deuce.river_equity(pockets, public)
    o = oracle.river_equity_batch(pockets, public)
    assert np.array_equal(g[2], o[2]) and np.array_equal(g[3], o[3])       # wins, decisive
    assert np.array_equal(g[0].view(np.uint32), o[0].view(np.uint32))      # equity f32, bit-exact
    assert np.array_equal(g[1], o[1])                                      # bucket = round(100 p)
    # structural property at any size: swapping suits (an isomorphism) leaves equity unchanged
    def swap_cd(x):
        c = x & np.uint64(0x0001111111111111)
        d = x & np.uint64(0x0002222222222222)
        return (x & ~np.uint64(0x0003333333333333)) | (c << np.uint64(1)) | (d >> np.uint64(1))
    g2 = rbp.deuce.river_equity(swap_cd(pockets), swap_cd(public))
    assert np.array_equal(g2[2], g[2]) and np.array_equal(g2[3], g[3])
