"""Pins the oracle (CPU restatement) to every known answer the reference holds for the MCCFR path.

The reference has no golden numeric vectors for this path; what it does pin — and what is checked here
— are the Philox generator (Random123 KATs), the Leduc/Kuhn tree shapes, Kuhn's analytic Nash equilibrium and the
exploitability thresholds of the reference's own tests.
"""
import pytest

N18 = 1 << 18


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert oracle.philox([f, f, f, f], [f, f]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_tree_shapes(oracle):
    # Leduc exploitability tree 3,517 nodes / 1,860 terminals / 120 decision infosets
    assert oracle.OracleSolver("leduc").tree_stats() == {"nodes": 3517, "terminals": 1860, "infosets": 120}
    # Kuhn: 12 infosets (crates/kuhn/src/solver.rs:88); 1 + 6 + 30 + 30*8 nodes
    assert oracle.OracleSolver("kuhn").tree_stats() == {"nodes": 277, "terminals": 150, "infosets": 12}


def test_untrained_exploitability(oracle):
    # uniform strategy: exact rational values, Kuhn 17/40, Leduc 277/240
    assert abs(oracle.OracleSolver("kuhn").exploitability() - 0.425) < 1e-6
    assert abs(oracle.OracleSolver("leduc").exploitability() - 277 / 240) < 1e-5


def test_kuhn_nash_equilibrium(oracle):
    # crates/kuhn/src/solver.rs:178-203 (same assertions, same tolerances)
    s = oracle.OracleSolver("kuhn", "FlooredRegret", "LinearWeight", "ExternalSampling").solve(N18)
    from robopoker_b200 import kuhn_info

    def policy(rank, hist, a):
        return s.averaged_distribution(kuhn_info(rank, hist))[a]

    # choices: Open/Check -> [Check, Bet]; Bet/CheckBet -> [Fold, Call]
    assert policy("J", "Bet", 0) > 0.95
    assert policy("J", "CheckBet", 0) > 0.95
    assert policy("K", "Bet", 1) > 0.95
    assert policy("K", "CheckBet", 1) > 0.95
    assert policy("K", "Check", 1) > 0.95
    assert policy("Q", "Open", 0) > 0.85
    assert abs(policy("J", "Open", 1) - 9 / 31) < 0.05
    assert abs(policy("K", "Open", 1) - 27 / 31) < 0.05
    assert abs(policy("Q", "Bet", 1) - 17 / 31) < 0.08
    assert abs(policy("Q", "CheckBet", 1) - 23 / 31) < 0.05
    assert abs(policy("J", "Check", 1) - 9 / 31) < 0.05
    assert abs(policy("Q", "Check", 1) - 8 / 31) < 0.18
    assert abs(policy("K", "Open", 1) / policy("J", "Open", 1) - 3.0) < 0.4
    assert s.exploitability() < 0.020


# crates/kuhn/src/solver.rs:234-277 and crates/roshambo/src/solver.rs:205-250: the reference's full 44-combination test
# matrices (sampling x regret x weight, same iteration counts, same tolerances), transcribed into
# tests/golden/solver_combos.json by tests/golden/make_solver_combos.py
def _combos(key):
    import json
    import os

    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solver_combos.json")))[key]


def _run_all(jobs):
    from concurrent.futures import ThreadPoolExecutor  # the oracle runs outside the GIL

    with ThreadPoolExecutor(max_workers=8) as pool:
        return [f for f in pool.map(lambda job: job(), jobs) if f]


def test_kuhn_exploitability_thresholds_all_44(oracle):
    combos = _combos("kuhn_exploitability_after_2^18")
    assert len(combos) == 44

    def job(c):
        def run():
            e = oracle.OracleSolver("kuhn", c["regret"], c["weight"], c["sampling"]).solve(N18).exploitability()
            return None if e < c["tolerance"] else (c, e)
        return run

    assert _run_all([job(c) for c in combos]) == []


def test_rps_equilibrium_all_44(oracle):
    combos = _combos("rps_equilibrium_after_2^16")
    assert len(combos) == 44

    def job(c):
        def run():
            s = oracle.OracleSolver("rps", c["regret"], c["weight"], c["sampling"]).solve(1 << 16)
            for key in (1, 3):  # P1, P2
                r, p, sc = s.averaged_distribution(key)
                if not (abs(r - 0.40) < c["tolerance"] and abs(p - 0.40) < c["tolerance"] and abs(sc - 0.20) < c["tolerance"]):
                    return (c, key, r, p, sc)
            return None
        return run

    assert _run_all([job(c) for c in combos]) == []


# crates/leduc/src/solver.rs:121-123
@pytest.mark.parametrize("sampling,regret,weight", [
    ("ExternalSampling", "FlooredRegret", "LinearWeight"),
    ("ExternalSampling", "DiscountedRegret", "LinearWeight"),
    ("PrunableSampling", "FlooredRegret", "LinearWeight"),
])
def test_leduc_exploitability_threshold(oracle, sampling, regret, weight):
    s = oracle.OracleSolver("leduc", regret, weight, sampling).solve(N18)
    e = s.exploitability()
    assert e < 0.080, e


def test_threads_do_not_change_results(oracle):
    a = oracle.OracleSolver("leduc", batch=64, seed=3, threads=1).step(20)
    b = oracle.OracleSolver("leduc", batch=64, seed=3, threads=4).step(20)
    assert a.profile_rows().tobytes() == b.profile_rows().tobytes()


def test_rps_equilibrium_and_exploitability(oracle):
    # crates/roshambo/src/solver.rs:157-165,253-257: asymmetric payoffs (scissors double) -> (0.4, 0.4, 0.2), exploitability < 0.03
    s = oracle.OracleSolver("rps", "FlooredRegret", "LinearWeight", "ExternalSampling").solve(1 << 16)
    assert s.tree_stats() == {"nodes": 13, "terminals": 9, "infosets": 2}
    for key in (1, 3):  # P1, P2
        r, p, sc = s.averaged_distribution(key)
        assert abs(r - 0.40) < 0.05 and abs(p - 0.40) < 0.05 and abs(sc - 0.20) < 0.05
    assert s.exploitability() < 0.03
