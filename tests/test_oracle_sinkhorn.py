"""Pins the Sinkhorn oracle to the reference's own property tests (crates/lloyd/src/sinkhorn.rs:240-293: synthetic
32-abstraction metric d(i,j) = ((7i+13j) mod 97 + 1)/100, self-divergence < 1e-4, symmetry < 1e-3) under BOTH math
policies, and bounds the exp/ln contract's distance from libm."""
import math

import numpy as np


def flop_metric(bins=32):
    # sinkhorn.rs:252-262 flop_metric(): pair (i<j) -> ((i*7 + j*13) % 97 + 1) / 100, stored at Pair::merge(i, j)
    tri = np.zeros(bins * (bins - 1) // 2, np.float32)
    for i in range(bins):
        for j in range(i + 1, bins):
            tri[j * (j - 1) // 2 + i] = np.float32(((i * 7 + j * 13) % 97 + 1)) / np.float32(100.0)
    return tri


def hist(entries, bins=32):
    h = np.zeros(bins, np.uint32)
    for i, c in entries:
        h[i] = c
    return h


def test_exp_ln_contract_close_to_libm(oracle):
    xs = np.concatenate([np.linspace(-87, 88, 4001), np.linspace(-1, 1, 2001), [-87.33654, 88.72283, 0.0]]).astype(np.float32)
    for x in xs:
        got, ref = oracle.exp_c(float(x)), math.exp(float(x))
        assert abs(got - ref) <= 3e-7 * ref + 3e-45, (x, got, ref)
    # saturating outside [ln MIN_POSITIVE, ln MAX]: the softmin clamps at MIN_POSITIVE anyway (sinkhorn.rs:120)
    lo, hi = oracle.exp_c(-87.33654), oracle.exp_c(88.72283)
    assert 0.0 < lo <= 1.1755e-38 and math.isfinite(hi) and hi > 3.4e38
    for x in (-100.0, -103.5, -1e30, float("-inf"), float("nan")):
        assert oracle.exp_c(x) == lo
    for x in (89.0, 1e30, float("inf")):
        assert oracle.exp_c(x) == hi
    ys = np.concatenate([np.logspace(-37, 37, 3001), np.linspace(0.5, 2.0, 2001), [1.17549435e-38, 1e-40]]).astype(np.float32)
    for y in ys:
        got, ref = oracle.ln_c(float(y)), math.log(float(y))
        assert abs(got - ref) <= 3e-7 * abs(ref) + 2e-7, (y, got, ref)
    assert oracle.ln_c(1.0) == 0.0 and oracle.exp_c(0.0) == 1.0


def test_saturating_exp_never_returns_less_than_min_positive(oracle):
    # sinkhorn.rs:120 clamps every softmin term at MIN_POSITIVE; under the saturating contract the clamp is the identity
    # (exhaustive over every float from the lower saturation point up to -80; above that exp(x) > 1e-35), which is why
    # the device softmin (csrc/sinkhorn.cuh sk_term) has no max instruction
    FLT_MIN = 1.17549435e-38
    assert oracle.exp_c_min_over(-87.33654, -80.0) >= FLT_MIN
    assert oracle.exp_c(-87.33654) >= FLT_MIN and oracle.exp_c(float("-inf")) >= FLT_MIN and oracle.exp_c(float("nan")) >= FLT_MIN


def test_reference_property_tests_both_math_policies(oracle):
    tri = flop_metric()
    for m in (0, 1):
        h = hist([(0, 3), (5, 1), (12, 4), (24, 2)])
        d = oracle.sinkhorn_divergence_batch(h[None], h[None], tri, math=m)[0]
        assert abs(d) < 1e-4                                    # divergence_is_zero_on_self
        mu, nu = hist([(0, 3), (5, 1), (12, 4)]), hist([(2, 2), (8, 5), (20, 1), (24, 3)])
        d12 = oracle.sinkhorn_divergence_batch(mu[None], nu[None], tri, math=m)[0]
        d21 = oracle.sinkhorn_divergence_batch(nu[None], mu[None], tri, math=m)[0]
        assert abs(d12 - d21) < 1e-3 and d12 > 0                # divergence_is_symmetric


def test_contract_tracks_libm(oracle):
    rng = np.random.default_rng(0)
    tri = flop_metric()
    a = np.zeros((64, 32), np.uint32)
    b = np.zeros((64, 32), np.uint32)
    for i in range(64):
        a[i, rng.choice(32, size=rng.integers(1, 9), replace=False)] = rng.integers(1, 9, size=1)
        b[i, rng.choice(32, size=rng.integers(1, 9), replace=False)] = rng.integers(1, 9, size=1)
    dc = oracle.sinkhorn_divergence_batch(a, b, tri, math=0)
    dl = oracle.sinkhorn_divergence_batch(a, b, tri, math=1)
    assert np.max(np.abs(dc - dl)) < 2e-5, np.max(np.abs(dc - dl))
    c, iters = oracle.ot_cost(a[0], b[0], tri)
    assert 1 <= iters <= 128 and c >= 0
