"""Pins the oracle's evaluator to the reference's own known answers (tests/golden/eval_known_answers.json, generated
from crates/deuce/src/evaluator.rs:186-357 by tests/golden/make_eval_golden.py) and to counting facts."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "eval_known_answers.json")))


def random_hands(rng, n, k):
    out = np.zeros(n, dtype=np.uint64)
    for i in range(n):
        cards = rng.choice(52, size=k, replace=False)
        out[i] = sum(1 << int(c) for c in cards)
    return out


def test_known_answers(oracle):
    assert len(GOLD["cases"]) == 19
    hands = np.array([c["bits"] for c in GOLD["cases"]], dtype=np.uint64)
    got = oracle.eval_batch(hands)
    for c, g in zip(GOLD["cases"], got):
        assert int(g) == c["packed"], (c["name"], hex(int(g)), hex(c["packed"]))


def test_ord_quirk_fullhouse_below_flush(oracle):
    from robopoker_b200.deuce import hand

    fh, fl = oracle.eval_batch([hand("2s 2h 2d 3c 3s"), hand("As Ks Qs Js 9s")])
    assert fh >> 24 == 5 and fl >> 24 == 6 and fh < fl  # ranking.rs:33-44 as written
    # Flush carries only its top rank: equal-high flushes tie (ranking.rs:39)
    a, b = oracle.eval_batch([hand("As Ks Qs Js 9s"), hand("As 8s 7s 3s 2s")])
    assert a == b


def test_category_counts_five_cards(oracle):
    # all C(52,5) hands: category frequencies of standard poker
    from itertools import combinations

    bits = np.fromiter((sum(1 << c for c in comb) for comb in combinations(range(52), 5)), dtype=np.uint64, count=2598960)
    tags = oracle.eval_batch(bits) >> 24
    counts = np.bincount(tags, minlength=9)
    # HighCard, OnePair, TwoPair, Trips, Straight, FullHouse, Flush, Quads, StraightFlush
    assert counts.tolist() == [1302540, 1098240, 123552, 54912, 10200, 3744, 5108, 624, 40]


def test_equity_counts(oracle):
    from robopoker_b200.deuce import hand

    # the nuts: royal flush on board -> every showdown ties -> 0.5 by the reference's convention
    eq, bk, w, t = oracle.river_equity_batch([hand("2c 3d")], [hand("Ts Js Qs Ks As")])
    assert (eq[0], bk[0], w[0], t[0]) == (0.5, 50, 0, 0)
    eq, bk, w, t = oracle.river_equity_batch([hand("As Ks")], [hand("Qs Js Ts 2d 3c")])
    assert (eq[0], bk[0], w[0], t[0]) == (1.0, 100, 990, 990)
    rng = np.random.default_rng(0)
    seven = random_hands(rng, 64, 7)
    pockets = np.array([int(h) & -int(h) | (int(h) & (int(h) - 1)) & -(int(h) & (int(h) - 1)) for h in seven], dtype=np.uint64)
    eq, bk, w, t = oracle.river_equity_batch(pockets, seven & ~pockets)
    assert (t <= 990).all() and (w <= t).all() and ((eq >= 0) & (eq <= 1)).all()
