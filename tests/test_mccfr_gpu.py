"""GPU parity: the CUDA MCCFR path (through the C ABI) against the oracle on identical Philox streams.

The bar for this path is bit-exact: every f32 operation of sampling, value computation and the ordered fold is
performed in the oracle's order without FMA contraction, so tables, exploitability and counters must be identical
(the north-star's 1e-5 relative / 1e-4 exploitability tolerances are therefore met with margin zero).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rows_equal(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for f in ("info_key", "action", "visits"):
        assert np.array_equal(a[f], b[f]), f
    for f in ("regret", "weight", "payoff"):
        x, y = a[f].view(np.uint32), b[f].view(np.uint32)
        bad = np.nonzero(x != y)[0]
        assert bad.size == 0, (f, a[bad[:4]], b[bad[:4]])


@pytest.mark.parametrize("game", ["kuhn", "leduc"])
def test_game_shape_and_untrained_exploitability(rbp, oracle, game):
    g, o = rbp.Solver(game), oracle.OracleSolver(game)
    st, shape = o.tree_stats(), g.game_shape()
    assert (shape["nodes"], shape["terminals"], shape["infosets"]) == (st["nodes"], st["terminals"], st["infosets"])
    assert np.float32(g.exploitability()).view(np.uint32) == np.float32(o.exploitability()).view(np.uint32)


CASES = [
    # game, regret, weight, sampling, batch, epochs
    ("kuhn", "FlooredRegret", "LinearWeight", "ExternalSampling", 1, 2000),
    ("kuhn", "SummedRegret", "ConstantWeight", "ExternalSampling", 7, 300),
    ("kuhn", "LinearRegret", "QuadraticWeight", "ExternalSampling", 128, 100),
    ("kuhn", "DiscountedRegret", "ExponentialWeight", "ExternalSampling", 300, 60),
    ("kuhn", "AsymmetricRegret", "LinearWeight", "PrunableSampling", 64, 100),
    ("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", 1, 3000),
    ("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", 1024, 64),
    ("leduc", "LinearRegret", "LinearWeight", "ExternalSampling", 129, 100),
    ("leduc", "DiscountedRegret", "LinearWeight", "ExternalSampling", 256, 80),
    ("leduc", "SummedRegret", "ConstantWeight", "PrunableSampling", 500, 40),
    ("leduc", "AsymmetricRegret", "QuadraticWeight", "PluribusSampling", 200, 60),
    ("leduc", "FlooredRegret", "LinearWeight", "TargetedSampling", 333, 60),
    ("kuhn", "LinearRegret", "ConstantWeight", "TargetedSampling", 50, 200),
]


@pytest.mark.parametrize("game,regret,weight,sampling,batch,epochs", CASES)
def test_bit_exact_against_oracle(rbp, oracle, game, regret, weight, sampling, batch, epochs):
    g = rbp.Solver(game, regret, weight, sampling, batch=batch, seed=11)
    o = oracle.OracleSolver(game, regret, weight, sampling, batch=batch, seed=11, threads=4)
    for chunk in (1, 1, epochs - 2):
        g.step(chunk)
        o.step(chunk)
        rows_equal(g.profile_rows(), o.profile_rows())
    assert g.epochs == o.epochs == epochs
    assert g.counters() == o.counters()
    assert np.float32(g.exploitability()).view(np.uint32) == np.float32(o.exploitability()).view(np.uint32)


def test_pluribus_pruning_engages(rbp, oracle):
    # tiny warm-up and a threshold regrets actually cross, so the pruned branch of sample/pluribus.rs:85-99 runs
    hyper = rbp.Hyper(prune_warmup=4, prune_threshold=-0.5, prune_explore=0.25)
    g = rbp.Solver("leduc", "SummedRegret", "LinearWeight", "PluribusSampling", batch=256, seed=5, hyper=hyper)
    o = oracle.OracleSolver("leduc", "SummedRegret", "LinearWeight", "PluribusSampling", batch=256, seed=5, threads=4)
    o.set_hyper(prune_warmup=4, prune_threshold=-0.5, prune_explore=0.25)
    g.step(60)
    o.step(60)
    rows_equal(g.profile_rows(), o.profile_rows())
    assert g.counters() == o.counters()


def test_large_batch_full_size(rbp, oracle):
    # the bench configuration (16384 trees / epoch), a few epochs, still bit-exact
    g = rbp.Solver("leduc", batch=16384, seed=0)
    o = oracle.OracleSolver("leduc", batch=16384, seed=0, threads=8)
    g.step(6)
    o.step(6)
    rows_equal(g.profile_rows(), o.profile_rows())
    assert g.counters() == o.counters()


def test_convergence_properties_at_scale(rbp):
    # size-independent properties at 1M trees: exploitability under the reference's own threshold
    # (crates/leduc/src/solver.rs:121-123), average strategies are distributions, visits add up
    g = rbp.Solver("leduc", batch=1024, seed=1).step(1024)
    assert g.exploitability() < 0.080
    rows = g.profile_rows()
    assert len(rows) == 240
    c = g.counters()
    assert int(rows["visits"].sum()) == 2 * c["infos"]  # every Decisions bumps both rows of its infoset
    assert c["updates"] == 2 * c["infos"]
    for key in np.unique(rows["info_key"]):
        p = g.averaged_distribution(int(key))
        assert abs(sum(p) - 1.0) < 1e-6 and min(p) >= 0.0


def test_import_export_roundtrip_and_resume(rbp, oracle):
    a = rbp.Solver("leduc", batch=64, seed=9).step(30)
    rows, epochs = a.profile_rows().copy(), a.epochs
    b = rbp.Solver("leduc", batch=64, seed=9)
    b.import_rows(rows, epochs)
    rows_equal(b.profile_rows(), rows)
    a.step(10)
    b.step(10)
    rows_equal(a.profile_rows(), b.profile_rows())
    o = oracle.OracleSolver("leduc", batch=64, seed=9, threads=2)
    o.import_rows(rows, epochs)
    o.step(10)
    rows_equal(a.profile_rows(), o.profile_rows())


def test_timed_step_matches_untimed(rbp):
    a = rbp.Solver("leduc", batch=512, seed=2).step(12)
    b = rbp.Solver("leduc", batch=512, seed=2)
    total, s, f = b.step_timed(12, flush_l2=True)
    assert total > 0 and s > 0 and f > 0 and abs(total - (s + f)) < 0.25 * total
    rows_equal(a.profile_rows(), b.profile_rows())


def test_welford_division_is_ieee_exact(rbp):
    # the fold kernel divides by (visits+1) through a precomputed reciprocal + two FMAs; it must equal IEEE `/`
    import ctypes

    bad = ctypes.c_uint64(1)
    st = rbp.load_library().rbp_selftest_div_by_count(1 << 22, 64, ctypes.byref(bad))
    assert st == 0 and bad.value == 0, bad.value


BATCHED = [
    ("kuhn", "FlooredRegret", "LinearWeight", "ExternalSampling", 1, 500),
    ("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", 1000, 40),
    ("leduc", "LinearRegret", "QuadraticWeight", "ExternalSampling", 129, 60),
    ("leduc", "DiscountedRegret", "ExponentialWeight", "PrunableSampling", 300, 50),
    ("leduc", "SummedRegret", "ConstantWeight", "ExternalSampling", 16384, 5),
]


@pytest.mark.parametrize("game,regret,weight,sampling,batch,epochs", BATCHED)
def test_batched_fold_bit_exact(rbp, oracle, game, regret, weight, sampling, batch, epochs):
    g = rbp.Solver(game, regret, weight, sampling, batch=batch, seed=4, fold=rbp.FOLD_BATCHED)
    o = oracle.OracleSolver(game, regret, weight, sampling, batch=batch, seed=4, threads=4)
    o.set_fold(1)
    g.step(epochs)
    o.step(epochs)
    rows_equal(g.profile_rows(), o.profile_rows())
    assert g.counters() == o.counters()


def test_batched_two_ranks_on_one_gpu(rbp, oracle):
    # world_size 2 emulated in one process: two handles own tree ids [0,B) and [B,2B); the exchange is a concat
    import torch

    from robopoker_b200.distributed import _DeviceWords

    B, E = 640, 20
    hs = [rbp.Solver("leduc", "LinearRegret", "LinearWeight", batch=B, seed=8, fold=rbp.FOLD_BATCHED) for _ in range(2)]
    os_ = [oracle.OracleSolver("leduc", "LinearRegret", "LinearWeight", batch=B, seed=8, threads=4) for _ in range(2)]
    for r in range(2):
        hs[r].set_world(r, 2)
        os_[r].set_fold(1, r, 2)
    views = [torch.as_tensor(_DeviceWords(*h.delta_buffer()), device="cuda") for h in hs]
    for _ in range(E):
        for h in hs:
            h.sample()
        gathered = torch.cat(views)
        torch.cuda.synchronize()
        for h in hs:
            h.fold_gathered(gathered.data_ptr(), 2)
        words = np.concatenate([o.sample() for o in os_])
        for o in os_:
            o.fold_gathered(words, 2)
    rows_equal(hs[0].profile_rows(), hs[1].profile_rows())
    rows_equal(hs[0].profile_rows(), os_[0].profile_rows())


@pytest.mark.parametrize("fold", [0, 1])
def test_rps_bit_exact(rbp, oracle, fold):
    # third validation game (crates/roshambo): 3 actions, no deal, an infoset spanning three nodes of one tree
    g = rbp.Solver("rps", "DiscountedRegret", "LinearWeight", "PluribusSampling", batch=97, seed=6, fold=fold)
    o = oracle.OracleSolver("rps", "DiscountedRegret", "LinearWeight", "PluribusSampling", batch=97, seed=6, threads=2)
    o.set_fold(fold)
    assert np.float32(g.exploitability()).view(np.uint32) == np.float32(o.exploitability()).view(np.uint32)
    g.step(300)
    o.step(300)
    rows_equal(g.profile_rows(), o.profile_rows())
    assert g.counters() == o.counters()
    assert np.float32(g.exploitability()).view(np.uint32) == np.float32(o.exploitability()).view(np.uint32)


def test_batched_fold_at_its_old_bench_size(rbp, oracle):
    # the (non-reference) batched fold at 262144 trees / epoch with the flagship-of-round-1 tuple: still the oracle's sums, bit for bit
    g = rbp.Solver("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", batch=262144, seed=0, fold=rbp.FOLD_BATCHED)
    o = oracle.OracleSolver("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", batch=262144, seed=0, threads=8)
    o.set_fold(1)
    g.step(2)
    o.step(2)
    rows_equal(g.profile_rows(), o.profile_rows())
    assert g.counters() == o.counters()


def test_batch_one_exploitability_lock_step_to_2_16(rbp, oracle):
    # BASELINE.json configs[1] in the reference's own regime (batch_size() = 1): the exploitability curve sampled at 2^k epochs is the oracle's, bit
    # for bit, and ends under the reference's threshold for this many iterations (crates/leduc/src/solver.rs:121-123 asserts < 0.080 at 2^18)
    g = rbp.Solver("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", batch=1, seed=0)
    o = oracle.OracleSolver("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", batch=1, seed=0, threads=1)
    done = 0
    for k in range(6, 17):
        g.step((1 << k) - done)
        o.step((1 << k) - done)
        done = 1 << k
        assert np.float32(g.exploitability()).view(np.uint32) == np.float32(o.exploitability()).view(np.uint32), k
    rows_equal(g.profile_rows(), o.profile_rows())
    assert g.exploitability() < 0.15


def test_spend_runs_until_the_deadline(rbp):
    # `Solver::spend` (solver.rs:130-137): the real-time players' wall-clock budget
    g = rbp.Solver("leduc", batch=64, seed=3)
    n, dt = g.spend(0.2)
    assert n > 0 and n == g.epochs and 0.2 <= dt < 1.0
    h = rbp.Solver("leduc", batch=64, seed=3).step(n)
    rows_equal(g.profile_rows(), h.profile_rows())   # the budgeted run is the same run
