"""Pins the oracle's safe subgame solver (oracle/subgame.hpp) to what the reference holds for it, and the library's host half
(`rbp_subgame_partition`, `rbp_subgame_entries` — no device needed) to the oracle.

The reference's subgame tests are statistical (kuhn/src/solver.rs `subgame_*`: after 2^16 subgame steps on a 2^18-epoch blueprint, K calls a
bet with probability > 0.90 averaged over the worlds; `restrict_produces_valid_deals`); `Partition::partition` is checked against
hand-evaluated cases of world/partition.rs:27-53.
"""
import hashlib
import itertools
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = os.path.join(HERE, "golden", "subgame_pins.json")
N18, N16 = 1 << 18, 1 << 16


def baseline(cards, external):
    """`KuhnEncoder::baseline` / `LeducEncoder::baseline` (kuhn/src/encoder.rs:75-86): uniform reach over the cards the external player could hold."""
    reach = [0.0, 0.0, 0.0]
    for c in range(6):
        if c != cards[1 - external]:
            reach[c >> 1] += 1.0
    return reach


@pytest.fixture(scope="module")
def kuhn_blueprint(oracle):
    return oracle.OracleSolver("kuhn", "FlooredRegret", "LinearWeight", "ExternalSampling", batch=1, seed=0).solve(N18)


def test_partition_cases(oracle, rbp):
    cases = [([1, 2, 2], 2), ([0, 0, 0], 2), ([5, 1, 1], 2), ([1, 1, 1], 3), ([0.2, 0.5, 0.3], 2), ([3, 3, 3], 4), ([1, 0, 0], 2)]
    # hand-evaluated from world/partition.rs:27-53 (descending stable sort, a world closes when the running mass reaches its quantile)
    want = {0: ([1, 0, 0], [0.8, 0.2]), 1: ([0, 0, 0], [0.5, 0.5]), 2: ([0, 1, 1], [5 / 7, 2 / 7]), 3: ([0, 1, 2], [1 / 3, 1 / 3, 1 / 3]),
            6: ([0, 1, 1], [1.0, 0.0])}
    for i, (reach, w) in enumerate(cases):
        o_world, o_weights = oracle.partition(reach, w)
        g_world, g_weights = rbp.subgame.partition(reach, w)                      # the library's host arithmetic
        assert np.array_equal(o_world, g_world) and np.array_equal(o_weights.view(np.uint32), g_weights.view(np.uint32)), (reach, w)
        assert abs(float(o_weights.sum()) - 1.0) < 1e-6
        if i in want:
            assert o_world.tolist() == want[i][0] and np.allclose(o_weights, want[i][1], atol=1e-7), (reach, o_world, o_weights)


def test_restrict_produces_valid_deals(oracle, kuhn_blueprint):
    # kuhn/src/solver.rs `restrict_produces_valid_deals`: every world's entry has two different cards, and its external card's rank is in the world
    for cards in itertools.permutations(range(6), 2):
        for external in (0, 1):
            world_of, weights = oracle.partition(baseline(cards, external), 2)
            sg = oracle.OracleSubgame(kuhn_blueprint, external, world_of, weights, cards)
            for w in range(2):
                c = sg.entry(w)
                assert c[0] != c[1] and c[1 - external] == cards[1 - external]
                if (world_of == w).any() and any(world_of[k >> 1] == w for k in range(6) if k != cards[1 - external]):
                    assert world_of[c[external] >> 1] == w


def test_host_entries_equal_the_oracle(oracle, rbp):
    # the library's `WorldRestrict` + path walk over the flat enumeration against the oracle's game states: cards and infoset keys
    n = 0
    for game in ("kuhn", "leduc"):
        bp = oracle.OracleSolver(game).step(4)
        paths = [(), (0,), (1,), (0, 1)] if game == "kuhn" else [(0, 0, b) for b in range(4)] + [(0, 0, b, 0) for b in range(4)] + [(1, 1, b) for b in range(4)]
        for c0, c1 in itertools.permutations(range(6), 2):
            for path in paths:
                for external in (0, 1):
                    for world_of in (None, [0, 1, 0], [1, 0, 0], [1, 1, 0]):
                        try:
                            cards, _, keys = rbp.subgame.entries(game, external, world_of, 2, (c0, c1), path)
                        except rbp.RbpError:
                            continue                                                   # not a decision node / a chance node below it
                        sg = oracle.OracleSubgame(bp, external, world_of, [0.5, 0.5], (c0, c1), path)
                        for w in range(2):
                            assert cards[w] == sg.entry(w) and int(keys[w]) == sg.entry_key(w), (game, c0, c1, path, external, world_of, w)
                            n += 1
    assert n > 5000


def test_leduc_first_round_stops_at_the_board_deal(oracle, rbp):
    # `WorldEncoder::branches` (world/encoder.rs:97-106) does not expand chance nodes: from a first-round entry the tree ends at the board deal,
    # and `terminal_value` (mccfr/src/strategy/nash.rs:66-79) gives such a leaf the stored V(I) of the decision node above it
    bp = oracle.OracleSolver("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", batch=1, seed=3).step(4096)
    cards, nodes, keys = rbp.subgame.entries("leduc", 1, None, 2, (0, 3), ())        # accepted by the library's host half
    sg = oracle.OracleSubgame(bp, 1, None, [0.5, 0.5], (0, 3), (), seed=2).step(2000)
    assert cards[0] == sg.entry(0) and int(keys[0]) == sg.entry_key(0)
    rows = np.concatenate([sg.profile_rows(w) for w in range(2)])
    assert len(rows) > 0 and ((rows["info_key"] >> 1) & 3 == 0).all()                 # only first-round infosets (no board) were ever written
    assert np.isfinite(rows["regret"]).all() and np.abs(rows["payoff"]).max() > 0.0   # the chance leaves carried the blueprint's V(I)


def subpolicy(sg, worlds, rank, hist, action):
    from robopoker_b200 import kuhn_info

    return float(np.mean([sg.averaged_distribution(w, kuhn_info(rank, hist))[action] for w in range(worlds)]))


@pytest.mark.parametrize("external,cards,path", [(1, (2, 5), ()), (0, (0, 3), (0,)), (0, (4, 1), (1,))])
def test_kuhn_subgame_nash(oracle, kuhn_blueprint, external, cards, path):
    # kuhn/src/solver.rs `subgame!`, `subgame_after_check`, `subgame_after_bet` + `subgame_nash`: K|B and K|XB call > 0.90
    world_of, weights = oracle.partition(baseline(cards, external), 2)
    sg = oracle.OracleSubgame(kuhn_blueprint, external, world_of, weights, cards, path, seed=1).step(N16)
    assert sg.t == N16 and int(sg.drawn().sum()) == N16
    assert abs(sg.drawn()[0] / N16 - weights[0]) < 0.01                                # worlds are drawn in proportion to the belief
    assert subpolicy(sg, 2, "K", "Bet", 1) > 0.90 and subpolicy(sg, 2, "K", "CheckBet", 1) > 0.90
    assert sg.sum_regret() < 0.05


def test_reach_conditioned_posterior(oracle, rbp, kuhn_blueprint):
    # kuhn/src/solver.rs `subgame_with_reach_conditioned_posterior`: after P0 checks and P1 bets the posterior over P1's hand is weighted by
    # the blueprint's probability of that bet (K bets ~always after a check, J / Q less), the subgame converges and K|XB still calls
    cards, path, external = (2, 5), (0, 1), 1                                        # P0 holds Q; Check, Bet
    prior = oracle.subgame_posterior(kuhn_blueprint, external, cards, path)
    lib_prior = rbp.subgame.posterior("kuhn", kuhn_blueprint.profile_rows(), external, cards, path)   # the library's host arithmetic
    assert np.array_equal(prior.view(np.uint32), lib_prior.view(np.uint32))
    assert prior[2] > prior[0] and prior[2] > prior[1] and prior[2] > 1.8            # two K cards, each bet with probability ~1
    world_of, weights = oracle.partition(prior, 2)
    assert world_of[2] == 0                                                           # K carries the most reach: world 0
    sg = oracle.OracleSubgame(kuhn_blueprint, external, world_of, weights, cards, path, seed=5).step(N16)
    assert sg.sum_regret() < 0.01
    assert subpolicy(sg, 2, "K", "CheckBet", 1) > 0.90


def test_posterior_library_equals_oracle(oracle, rbp):
    n = 0
    for game, epochs in (("kuhn", 3000), ("leduc", 20000)):
        bp = oracle.OracleSolver(game, "FlooredRegret", "LinearWeight", "ExternalSampling", batch=1, seed=2).step(epochs)
        rows = bp.profile_rows()
        paths = [(), (0,), (1,), (0, 1)] if game == "kuhn" else [(0,), (1, 1), (0, 0, 2), (0, 0, 1, 0), (1, 1, 3, 1), (0, 1, 1, 0, 0)]
        for c0, c1 in itertools.permutations(range(6), 2):
            for path in paths:
                for external in (0, 1):
                    try:
                        got = rbp.subgame.posterior(game, rows, external, (c0, c1), path)
                    except rbp.RbpError:
                        continue
                    want = oracle.subgame_posterior(bp, external, (c0, c1), path)
                    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (game, c0, c1, path, external, got, want)
                    n += 1
    assert n > 500


def test_first_write_reads_through_to_the_blueprint(oracle, kuhn_blueprint):
    # world/profile.rs + strategy/profile.rs:94-104: the first update of an edge starts from max(blueprint regret, EPS) and from the
    # warmstart weight = averaged policy * k * (k + 1) / 2 (k = 2^14); at t = 0 LinearWeight adds policy * 0
    from robopoker_b200 import kuhn_info

    cards = (2, 5)
    sg = oracle.OracleSubgame(kuhn_blueprint, 1, None, [1.0], cards, (), seed=3).step(1)
    rows = sg.profile_rows(0)
    assert len(rows) > 0 and (rows["visits"] == 1).all()
    bp = {(int(r["info_key"]), int(r["action"])): r for r in kuhn_blueprint.profile_rows()}
    k = np.float32(16384.0)
    for r in rows:
        key = int(r["info_key"])
        w = np.array([max(np.float32(bp[(key, a)]["weight"]), np.finfo(np.float32).tiny) for a in range(2)], np.float32)
        policy = w[int(r["action"])] / (w[0] + w[1])
        assert np.float32(r["weight"]) == np.float32(np.float32(np.float32(policy * k) * np.float32(k + np.float32(1.0))) / np.float32(2.0))
    assert kuhn_info("Q", "Open") in {int(r["info_key"]) for r in rows}               # t = 0: the walker is seat 0, holding Q


@pytest.mark.parametrize("external,cards,path,entry_node,epochs", [(1, (2, 5), (), 0, 4096), (0, (0, 3), (0,), 1, 4096), (0, (4, 1), (1,), 2, 64), (1, (1, 0), (0, 1), 3, 300)])
def test_dense_table_model_equals_the_world_profile(oracle, external, cards, path, entry_node, epochs):
    # tests/subgame_dense_model.py restates the step the way the device does it (seeded dense tables per world, `visits == 0` -> blueprint
    # weight, LIFO tree from the entry node, ordered fold); the oracle keeps the reference's two-level HashMap profile.  Bit-identical rows.
    from subgame_dense_model import DenseSubgame

    bp = oracle.OracleSolver("kuhn", "FlooredRegret", "LinearWeight", "ExternalSampling", batch=1, seed=7).step(epochs)
    reach = [0.0, 0.0, 0.0]
    for c in range(6):
        if c != cards[1 - external]:
            reach[c >> 1] += 1.0 + 0.25 * (c >> 1)
    world_of, weights = oracle.partition(reach, 2)
    o = oracle.OracleSubgame(bp, external, world_of, weights, cards, path, seed=11)
    d = DenseSubgame(oracle.philox, bp.profile_rows(), 2, weights, [o.entry(w) for w in range(2)], entry_node, seed=11)
    for n in (1, 1, 2, 60, 400):
        o.step(n)
        for _ in range(n):
            d.step()
        assert d.drawn == [int(x) for x in o.drawn()]
        for w in range(2):
            want = o.profile_rows(w)
            got = d.rows(w)
            assert len(got) == len(want), (n, w)
            for g, r in zip(got, want):
                assert g[0] == int(r["info_key"]) and g[1] == int(r["action"]) and g[5] == int(r["visits"])
                for x, name in ((g[2], "weight"), (g[3], "regret"), (g[4], "payoff")):
                    assert np.float32(x).view(np.uint32) == np.float32(r[name]).view(np.uint32), (n, w, g, r)


def digest(rows):
    return hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest()


def pin_cases(oracle):
    out = {}
    for game, epochs, external, cards, path, worlds in (("kuhn", 4096, 1, (2, 5), (), 2), ("kuhn", 4096, 0, (0, 3), (0,), 3),
                                                        ("leduc", 8192, 1, (1, 4), (0, 0, 1), 2), ("leduc", 8192, 0, (5, 2), (1, 1, 0, 0), 2),
                                                        ("leduc", 8192, 1, (0, 3), (), 2), ("leduc", 8192, 0, (4, 1), (1,), 3)):
        bp = oracle.OracleSolver(game, "FlooredRegret", "LinearWeight", "ExternalSampling", batch=1, seed=7).step(epochs)
        reach = [0.0, 0.0, 0.0]
        for c in range(6):
            if c != cards[1 - external]:
                reach[c >> 1] += 1.0 + 0.25 * (c >> 1)
        world_of, weights = oracle.partition(reach, worlds)
        sg = oracle.OracleSubgame(bp, external, world_of, weights, cards, path, seed=11).step(2000)
        out[f"{game}|{external}|{cards}|{path}|{worlds}"] = {"drawn": [int(x) for x in sg.drawn()], "rows": [digest(sg.profile_rows(w)) for w in range(worlds)]}
    return out


def test_committed_pins(oracle):
    # freezes the oracle's subgame outputs (generator: `python tests/test_oracle_subgame.py`): a change of the restatement shows up here
    want = json.load(open(PINS))
    assert pin_cases(oracle) == want


if __name__ == "__main__":
    import sys

    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import binding

    json.dump(pin_cases(binding), open(PINS, "w"), indent=1)
    print("wrote", PINS)
