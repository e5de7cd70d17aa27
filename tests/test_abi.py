"""The C-ABI library loads on a CPU-only host, exports every symbol include/rbp.h declares, and refuses to compute
without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(rbp_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_declares_symbols():
    assert len(declared_symbols()) >= 15


def test_every_declared_symbol_is_exported(rbp):
    l = rbp.load_library()
    missing = [n for n in declared_symbols() if not hasattr(l, n)]
    assert not missing, missing


def test_philox_contract_matches_random123(rbp):
    l = rbp.load_library()
    out = (ctypes.c_uint32 * 4)()
    l.rbp_philox4x32_10((ctypes.c_uint32 * 4)(0, 0, 0, 0), (ctypes.c_uint32 * 2)(0, 0), out)
    assert list(out) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_hyper_defaults(rbp):
    h = rbp.Hyper().c
    # crates/mccfr/src/hyperparams/{sampling.rs:39-50,pruning.rs:40-55,training.rs:52-60}
    assert (h.temperature, h.smoothing, round(h.curiosity, 6)) == (1.0, 2.0, 0.05)
    assert (h.prune_threshold, round(h.prune_explore, 6), h.prune_warmup, h.regret_min) == (-3e5, 0.05, 16384, -4e6)


def test_info_key_packing(rbp):
    assert rbp.kuhn_info("J", "Open") == 1
    assert rbp.kuhn_info("K", "CheckBet") == 1 | (3 << 1) | (2 << 3)
    assert rbp.leduc_info("Q", None, "Open", None) == 1 | (1 << 8)
    assert rbp.leduc_info("K", "J", "Raised", "Checked") == 1 | (1 << 1) | (2 << 3) | (2 << 5) | (2 << 8)


def test_no_cpu_fallback(rbp):
    if rbp.load_library().rbp_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(rbp.RbpError) as e:
        rbp.Solver("kuhn")
    assert e.value.status == -2  # RBP_ERR_NO_DEVICE


def test_every_compute_family_refuses_without_a_device(rbp):
    """NLHE solver, both k-means layers, batched Sinkhorn, hand evaluation, river equity, isomorphism sets: each entry
    point reports RBP_ERR_NO_DEVICE on a CPU-only host instead of computing anything."""
    import numpy as np

    if rbp.load_library().rbp_device_count() > 0:
        pytest.skip("GPU present")
    from robopoker_b200.nlhe import Nlhe

    turn = np.zeros((4, 101), np.uint8); turn[:, 50] = 46
    flop = np.zeros((4, 32), np.uint8); flop[:, 3] = 47
    tri = np.full(32 * 31 // 2, 0.5, np.float32)
    hands = np.array([0x7F], np.uint64)
    calls = [lambda: Nlhe(batch=8, seed=0, table_slots=1 << 10),
             lambda: rbp.lloyd.Layer(turn, 2),
             lambda: rbp.lloyd.Layer(flop, 2, metric=tri),
             lambda: rbp.lloyd.sinkhorn_divergence(flop, flop, [0], [1], tri),
             lambda: rbp.deuce.strength(hands),
             lambda: rbp.deuce.river_equity(np.array([0x3], np.uint64), np.array([0x7C], np.uint64)),
             lambda: rbp.deuce.IsoSet("flop")]
    for call in calls:
        with pytest.raises(rbp.RbpError) as e:
            call()
        assert e.value.status == -2, call
