"""The MCCFR kernels of csrc/mccfr.cu — their actual source — run on the CPU under a SIMT shim (tests/simt/) and compared with the oracle.

Why this exists: the subgame hooks (entry node, per-world table offset, the `visits == 0` read of the blueprint's weight, `subgame_seed_kernel`)
were written when no GPU minutes were left.  The shim runs every CUDA thread of a block as a fiber, so the kernel text itself — not a model of
it — is what produces the tables checked here; the orchestration around the kernels (buffer sizes, launch shapes, EpochArgs, the world draw) is
restated in tests/simt/mccfr_simt.cpp.  It proves nothing about the CUDA runtime side (memory spaces, launch limits, divergence hazards): that is
what tests/test_subgame_gpu.py is for.  As a by-product the flat-game training kernels get a CPU parity check too.
"""
import ctypes
import os
import shutil
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"),
                                reason="needs g++ and the CUDA headers (host compilation of the kernel source)")
SAMPLERS = {"ExternalSampling": 0, "PrunableSampling": 2, "PluribusSampling": 3, "TargetedSampling": 4}


@pytest.fixture(scope="module")
def simt():
    from simt import build

    l = ctypes.CDLL(build.build())
    vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
    l.simt_create.restype = vp
    l.simt_create.argtypes = [i32] * 5 + [u64]
    l.simt_destroy.argtypes = [vp]
    l.simt_step.argtypes = [vp, u64]
    l.simt_export.argtypes = [vp, i32, vp, i32]
    l.simt_import.argtypes = [vp, vp, i32, u64]
    l.simt_subgame_create.restype = vp
    l.simt_subgame_create.argtypes = [vp, i32, u64]
    l.simt_subgame_step.argtypes = [vp, u64, vp, vp, vp]
    return l


def rows_of(simt, oracle, h, world=0):
    buf = np.zeros(4096, dtype=oracle.ROW_DTYPE)
    n = simt.simt_export(h, world, buf.ctypes.data, len(buf))
    return buf[:n]


@pytest.mark.parametrize("game,regret,weight,sampling,batch,steps", [
    ("kuhn", "FlooredRegret", "LinearWeight", "ExternalSampling", 1, 200),
    ("kuhn", "SummedRegret", "LinearWeight", "ExternalSampling", 1, 200),      # the schedules of the subgame solver, template instance <SUMMED, LINEAR, unmasked>
    ("kuhn", "DiscountedRegret", "ExponentialWeight", "PrunableSampling", 130, 6),   # two blocks, masked fold
    ("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", 1, 16),
    ("leduc", "LinearRegret", "QuadraticWeight", "PluribusSampling", 200, 4),
])
def test_training_kernels_under_the_shim_equal_the_oracle(simt, oracle, game, regret, weight, sampling, batch, steps):
    h = simt.simt_create(oracle.GAMES[game], oracle.REGRETS[regret], oracle.WEIGHTS[weight], SAMPLERS[sampling], batch, 7)
    o = oracle.OracleSolver(game, regret, weight, sampling, batch=batch, seed=7)
    simt.simt_step(h, steps)
    o.step(steps)
    a, b = rows_of(simt, oracle, h), o.profile_rows()
    simt.simt_destroy(h)
    assert len(a) == len(b) and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("game,epochs,external,cards,path,worlds,steps", [
    ("kuhn", 4096, 1, (2, 5), (), 2, 300),
    ("kuhn", 4096, 0, (0, 3), (0,), 3, 300),
    ("kuhn", 64, 1, (1, 0), (), 1, 150),            # barely trained blueprint: most rows read through
    ("kuhn", 300, 0, (4, 1), (0, 1), 2, 200),
    ("leduc", 8192, 1, (1, 4), (0, 0, 1), 2, 40),    # second betting round after check-check and the board
    ("leduc", 8192, 0, (5, 2), (1, 1, 0, 0), 2, 40),
    ("leduc", 8192, 1, (0, 3), (), 2, 60),           # first round: the tree stops at the board deal, valued by V(I) of the node above it
    ("leduc", 200, 0, (4, 1), (1,), 2, 60),
])
def test_subgame_kernels_under_the_shim_equal_the_oracle(simt, oracle, rbp, game, epochs, external, cards, path, worlds, steps):
    o_bp = oracle.OracleSolver(game, "FlooredRegret", "LinearWeight", "ExternalSampling", batch=1, seed=7).step(epochs)
    bp = simt.simt_create(oracle.GAMES[game], 1, 1, 0, 1, 7)
    bp_rows = np.ascontiguousarray(o_bp.profile_rows())
    assert simt.simt_import(bp, bp_rows.ctypes.data, len(bp_rows), epochs) == 0
    reach = [0.0, 0.0, 0.0]
    for c in range(6):
        if c != cards[1 - external]:
            reach[c >> 1] += 1.0 + 0.25 * (c >> 1)
    world_of, weights = oracle.partition(reach, worlds)
    o = oracle.OracleSubgame(o_bp, external, world_of, weights, cards, path, seed=11)
    _, nodes, _ = rbp.subgame.entries(game, external, world_of, worlds, cards, path)          # the library's host half supplies the entry nodes
    nodes = np.ascontiguousarray(nodes, np.int32)
    sg = simt.simt_subgame_create(bp, worlds, 11)                                             # subgame_seed_kernel
    drawn = np.zeros(8, np.uint64)
    for chunk in (1, 1, steps - 2):
        simt.simt_subgame_step(sg, chunk, weights.ctypes.data, nodes.ctypes.data, drawn.ctypes.data)   # mccfr_sample_kernel + mccfr_fold_kernel
        o.step(chunk)
        assert np.array_equal(drawn[:worlds], o.drawn())
        for w in range(worlds):
            a, b = rows_of(simt, oracle, sg, w), o.profile_rows(w)
            assert len(a) == len(b) and a.tobytes() == b.tobytes(), (chunk, w)
    simt.simt_destroy(sg)
    simt.simt_destroy(bp)
