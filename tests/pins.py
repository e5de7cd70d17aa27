"""The computations behind tests/golden/oracle_pins.json, written once against a small backend interface so that the
generator (oracle), the CPU test (oracle) and the GPU test (librbp_b200 through the C ABI) run the same recipe."""
import hashlib

import numpy as np

from lloyd_data import flop_histograms, flop_mixture_histograms, synthetic_metric, turn_histograms

NLHE_FIELDS = ("past", "present", "choices", "edge", "weight", "regret", "payoff", "visits")


def bits(x):
    return [int(v) for v in np.asarray(x, np.float32).reshape(-1).view(np.uint32)]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def reference_flop_metric(n=32):
    # sinkhorn.rs:252-262 flop_metric(): pair (i<j) -> ((i*7 + j*13) % 97 + 1) / 100, stored at Pair::merge(i, j)
    tri = np.zeros(n * (n - 1) // 2, np.float32)
    for i in range(n):
        for j in range(i + 1, n):
            tri[j * (j - 1) // 2 + i] = np.float32(((i * 7 + j * 13) % 97 + 1)) / np.float32(100.0)
    return tri


class OracleBackend:
    def __init__(self, oracle):
        self.o = oracle

    def solver(self, game, regret, weight, sampling, batch, seed, batched=False):
        s = self.o.OracleSolver(game, regret, weight, sampling, batch=batch, seed=seed)
        if batched:
            s.set_fold(1)
        return s

    def nlhe(self, seed, batch):
        n = self.o.OracleNlhe(seed=seed, batch=batch)
        n.rows = n.export
        return n

    def divergences(self, a, b, tri):
        return self.o.sinkhorn_divergence_batch(a, b, tri, math=0, threads=4)

    def kmeans(self, pts, k, tri):
        o = self.o.OracleKmeans(pts, k, threads=4)
        if tri is not None:
            o.set_metric(tri)
        o.drift = lambda: o.step()[0]
        return o

    def strengths(self, hands):
        return self.o.eval_batch(hands)


class DeviceBackend:
    def __init__(self, rbp):
        self.r = rbp

    def solver(self, game, regret, weight, sampling, batch, seed, batched=False):
        return self.r.Solver(game, regret, weight, sampling, batch=batch, seed=seed, fold=self.r.FOLD_BATCHED if batched else self.r.FOLD_ORDERED)

    def nlhe(self, seed, batch):
        from robopoker_b200.nlhe import Nlhe

        n = Nlhe(batch=batch, seed=seed, table_slots=1 << 16)
        n.rows = n.profile
        return n

    def divergences(self, a, b, tri):
        idx = np.arange(len(a))
        return self.r.lloyd.sinkhorn_divergence(a, b, idx, idx, tri)

    def kmeans(self, pts, k, tri):
        g = self.r.lloyd.Layer(pts, k, metric=tri)
        g.drift = lambda: g.step().drift
        return g

    def strengths(self, hands):
        return self.r.deuce.strength(hands)


def compute(be):
    pins = {}
    # MCCFR: Leduc / Kuhn exploitability at power-of-two epochs, batch 1 (the reference's batch_size()), seed 0
    s = be.solver("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", 1, 0)
    curve, done = [], 0
    for k in range(0, 13):
        s.step((1 << k) - done); done = 1 << k
        curve.append(bits(s.exploitability())[0])
    pins["leduc_floored_linear_external_batch1_seed0_exploitability_bits_at_2^k"] = curve
    pins["leduc_table_sha256_after_2^12"] = digest(s.profile_rows())
    k = be.solver("kuhn", "LinearRegret", "LinearWeight", "ExternalSampling", 1, 0).step(2048)
    pins["kuhn_linear_linear_external_batch1_seed0_exploitability_bits_at_2048"] = bits(k.exploitability())[0]
    b = be.solver("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling", 512, 3, batched=True).step(16)
    pins["leduc_batched_fold_batch512_seed3_table_sha256_after_16"] = digest(b.profile_rows())
    # NLHE blueprint: 2 Flagship epochs of 16 trees
    n = be.nlhe(5, 16)
    n.step(2)
    rows = n.rows()
    pins["nlhe_flagship_batch16_seed5_rows_sha256_after_2"] = digest(*[rows[f] for f in NLHE_FIELDS])
    pins["nlhe_flagship_batch16_seed5_updates_after_2"] = int(n.counters()["updates"])
    # Sinkhorn divergences under the exp/ln contract: the reference's synthetic metric, seeded histograms
    a = flop_histograms(24, 32, seed=0).astype(np.uint32)
    c = flop_histograms(24, 32, seed=1).astype(np.uint32)
    pins["sinkhorn_divergence_bits_bins32_reference_metric"] = bits(be.divergences(a, c, reference_flop_metric()))
    m = flop_mixture_histograms(96, 256, comps=8, alpha=0.05, seed=4)
    cen = m.reshape(8, 12, 256).astype(np.uint32).sum(axis=1)
    pins["sinkhorn_divergence_bits_bins256_points_vs_merged_centroids"] = bits(
        be.divergences(m[:16].astype(np.uint32), cen[np.arange(16) % 8], synthetic_metric(256, 2)))
    # k-means layers: W1 (turn) and Sinkhorn (flop), three Elkan steps
    for name, pts, kk, tri in (("turn_w1_n600_k8_seed0", turn_histograms(600, seed=0), 8, None),
                               ("flop_sinkhorn_n200_k5_bins32_seed0", flop_histograms(200, 32, seed=0), 5, synthetic_metric(32, 0))):
        o = be.kmeans(pts, kk, tri)
        chosen = o.init_centroids(0)
        o.init_bounds()
        drift = [o.drift() for _ in range(3)]
        assign, dist = o.lookup(with_distance=True)
        pins["kmeans_" + name] = {"chosen": [int(x) for x in chosen], "drift_bits_step3": bits(drift[-1]),
                                  "lookup_sha256": digest(np.asarray(assign, np.uint32)),
                                  "lookup_distance_sha256": digest(np.asarray(dist, np.float32)), "metric_sha256": digest(np.asarray(o.metric(), np.float32))}
    # deuce: strengths of seeded 7-card hands
    rng = np.random.default_rng(0)
    hands = np.array([sum(1 << int(c) for c in rng.choice(52, 7, replace=False)) for _ in range(64)], np.uint64)
    pins["eval7_strength_sha256_seed0_x64"] = digest(np.asarray(be.strengths(hands), np.uint32))
    return pins


def contract_pins(oracle):
    """exp/ln contract values (host side only: the device has no scalar entry point for them)."""
    xs = np.array([-87.33654, -80.0, -40.0, -10.5, -1.0, -1e-3, 0.0, 1e-3, 0.5, 1.0, 10.0, 40.0, 88.0], np.float32)
    ys = np.array([1.17549435e-38, 1e-30, 1e-5, 0.25, 0.70710678, 1.0, 1.5, 47.0, 1e10], np.float32)
    return {"exp_c_bits": {repr(float(x)): bits(oracle.exp_c(float(x)))[0] for x in xs},
            "ln_c_bits": {repr(float(y)): bits(oracle.ln_c(float(y)))[0] for y in ys}}
