"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/build/librbp_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RBP_ORACLE_LIB: bench.py's CPU legs point this at the -O3 -march=native build made on the timing host (build.build_oracle(native=True))
LIB_PATH = os.environ.get("RBP_ORACLE_LIB") or os.path.join(_HERE, "build", "librbp_oracle.so")

ROW_DTYPE = np.dtype([("info_key", "<u4"), ("action", "<u4"), ("weight", "<f4"), ("regret", "<f4"), ("payoff", "<f4"),
                      ("visits", "<u4")])
GAMES = {"kuhn": 0, "leduc": 1, "rps": 2}
REGRETS = {"SummedRegret": 0, "FlooredRegret": 1, "LinearRegret": 2, "DiscountedRegret": 3, "AsymmetricRegret": 4}
WEIGHTS = {"ConstantWeight": 0, "LinearWeight": 1, "QuadraticWeight": 2, "ExponentialWeight": 3}
SAMPLERS = {"ExternalSampling": 0, "VanillaSampling": 1, "PrunableSampling": 2, "PluribusSampling": 3, "TargetedSampling": 4}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run __graft_entry__.build()")
        l = ctypes.CDLL(LIB_PATH)
        vp, u64, i32, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32
        l.orc_solver_create.restype = vp
        l.orc_solver_create.argtypes = [i32, i32, i32, i32, i32, u64, i32]
        l.orc_solver_destroy.argtypes = [vp]
        l.orc_solver_step.argtypes = [vp, u64]
        for f in ("epochs", "updates", "nodes", "infos"):
            getattr(l, "orc_solver_" + f).restype = u64
            getattr(l, "orc_solver_" + f).argtypes = [vp]
        l.orc_solver_exploitability.restype = ctypes.c_float
        l.orc_solver_exploitability.argtypes = [vp]
        l.orc_solver_tree_stats.argtypes = [vp, ctypes.POINTER(i32)]
        l.orc_solver_export.argtypes = [vp, vp, i32]
        l.orc_solver_import.argtypes = [vp, vp, i32, u64]
        l.orc_solver_averaged.argtypes = [vp, u32, ctypes.POINTER(ctypes.c_float)]
        l.orc_solver_set_hyper.argtypes = [vp] + [ctypes.c_float] * 5 + [u32, ctypes.c_float]
        l.orc_solver_set_fold.argtypes = [vp, i32, i32, i32]
        l.orc_solver_partial_words.argtypes = [vp]
        l.orc_solver_sample.argtypes = [vp, vp]
        l.orc_solver_fold_gathered.argtypes = [vp, vp, i32]
        l.orc_philox.argtypes = [u32] * 6 + [ctypes.POINTER(u32)]
        _lib = l
    return _lib


def philox(counter, key):
    out = (ctypes.c_uint32 * 4)()
    lib().orc_philox(*counter, *key, out)
    return list(out)


class OracleSolver:
    def __init__(self, game, regret="FlooredRegret", weight="LinearWeight", sampling="ExternalSampling", batch=1, seed=0, threads=1):
        self.batch = batch
        self._h = lib().orc_solver_create(GAMES[game], REGRETS[regret], WEIGHTS[weight], SAMPLERS[sampling], batch, seed, threads)
        if not self._h:
            raise ValueError("orc_solver_create")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_solver_destroy(self._h)
            self._h = None

    def set_hyper(self, **kw):
        h = dict(temperature=1.0, smoothing=2.0, curiosity=0.05, prune_threshold=-3e5, prune_explore=0.05, prune_warmup=16384,
                 regret_min=-4e6)
        h.update(kw)
        lib().orc_solver_set_hyper(self._h, h["temperature"], h["smoothing"], h["curiosity"], h["prune_threshold"],
                                   h["prune_explore"], h["prune_warmup"], h["regret_min"])

    def set_fold(self, fold_mode, world_rank=0, world_size=1):
        lib().orc_solver_set_fold(self._h, fold_mode, world_rank, world_size)

    def partial_words(self):
        return lib().orc_solver_partial_words(self._h)

    def sample(self):
        """BATCHED fold, first half: this rank's blocked partial sums as a u32 word buffer."""
        out = np.zeros(self.partial_words(), dtype=np.uint32)
        lib().orc_solver_sample(self._h, out.ctypes.data)
        return out

    def fold_gathered(self, words, world):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        lib().orc_solver_fold_gathered(self._h, words.ctypes.data, world)

    def step(self, n=1):
        lib().orc_solver_step(self._h, n)
        return self

    def solve(self, trees):
        return self.step(trees // self.batch)

    @property
    def epochs(self):
        return lib().orc_solver_epochs(self._h)

    def counters(self):
        l = lib()
        return {"nodes": l.orc_solver_nodes(self._h), "infos": l.orc_solver_infos(self._h), "updates": l.orc_solver_updates(self._h)}

    def exploitability(self):
        return lib().orc_solver_exploitability(self._h)

    def tree_stats(self):
        out = (ctypes.c_int * 3)()
        lib().orc_solver_tree_stats(self._h, out)
        return {"nodes": out[0], "terminals": out[1], "infosets": out[2]}

    def profile_rows(self):
        buf = np.zeros(4096, dtype=ROW_DTYPE)
        n = lib().orc_solver_export(self._h, buf.ctypes.data, len(buf))
        return buf[:n]

    def import_rows(self, rows, epochs):
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        lib().orc_solver_import(self._h, rows.ctypes.data, len(rows), epochs)

    def averaged_distribution(self, info_key):
        out = (ctypes.c_float * 8)()
        n = lib().orc_solver_averaged(self._h, info_key, out)
        return [out[i] for i in range(n)]


def _sub():
    l = lib()
    if not getattr(l, "_sub_ready", False):
        vp, u64, i32, u32, f32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_float
        l.orc_partition.argtypes = [vp, i32, i32, vp, vp]
        l.orc_subgame_posterior.argtypes = [vp, i32, i32, i32, vp, i32, vp]
        l.orc_subgame_create.restype = vp
        l.orc_subgame_create.argtypes = [vp, i32, i32, vp, vp, i32, i32, vp, i32, u64]
        l.orc_subgame_destroy.argtypes = [vp]
        l.orc_subgame_step.argtypes = [vp, u64]
        l.orc_subgame_t.restype = u64
        l.orc_subgame_t.argtypes = [vp]
        l.orc_subgame_drawn.argtypes = [vp, vp]
        l.orc_subgame_sum_regret.restype = f32
        l.orc_subgame_sum_regret.argtypes = [vp]
        l.orc_subgame_entry.argtypes = [vp, i32, vp]
        l.orc_subgame_entry_key.restype = u32
        l.orc_subgame_entry_key.argtypes = [vp, i32]
        l.orc_subgame_export.argtypes = [vp, i32, vp, i32]
        l.orc_subgame_averaged.argtypes = [vp, i32, u32, vp]
        l.orc_subgame_harvest.argtypes = [vp, u32, vp, vp, vp]
        l._sub_ready = True
    return l


def partition(reach, worlds):
    """`Partition::partition::<W>` (subgame/src/world/partition.rs:27-53): (world of every secret, per-world weights)."""
    reach = np.ascontiguousarray(reach, dtype=np.float32)
    world_of, weights = np.zeros(len(reach), np.int32), np.zeros(worlds, np.float32)
    _sub().orc_partition(reach.ctypes.data, len(reach), worlds, world_of.ctypes.data, weights.ctypes.data)
    return world_of, weights


def subgame_posterior(blueprint, external, cards, path=()):
    """Reach per rank of the external player's hand given the path (external_reach per card, Posterior::add per rank)."""
    p = np.ascontiguousarray(list(path), dtype=np.uint8)
    out = np.zeros(3, np.float32)
    _sub().orc_subgame_posterior(blueprint._h, external, cards[0], cards[1], p.ctypes.data if len(p) else None, len(p), out.ctypes.data)
    return out


class OracleSubgame:
    """`WorldSolver` / `SubGameSolver` without an origin over an OracleSolver blueprint (oracle/subgame.hpp)."""

    def __init__(self, blueprint, external, world_of_rank, weights, cards, path=(), seed=0):
        self.blueprint = blueprint  # keeps the blueprint alive: the subgame reads through to it
        self.worlds = len(weights)
        w = np.ascontiguousarray(weights, dtype=np.float32)
        m = None if world_of_rank is None else np.ascontiguousarray(world_of_rank, dtype=np.int32)
        p = np.ascontiguousarray(list(path), dtype=np.uint8)
        self._h = _sub().orc_subgame_create(blueprint._h, external, self.worlds, None if m is None else m.ctypes.data, w.ctypes.data,
                                            cards[0], cards[1], p.ctypes.data if len(p) else None, len(p), seed)
        if not self._h:
            raise ValueError("orc_subgame_create")

    def __del__(self):
        if getattr(self, "_h", None):
            _sub().orc_subgame_destroy(self._h)
            self._h = None

    def step(self, n=1):
        _sub().orc_subgame_step(self._h, n)
        return self

    @property
    def t(self):
        return _sub().orc_subgame_t(self._h)

    def drawn(self):
        out = np.zeros(self.worlds, np.uint64)
        _sub().orc_subgame_drawn(self._h, out.ctypes.data)
        return out

    def sum_regret(self):
        return _sub().orc_subgame_sum_regret(self._h)

    def entry(self, world):
        out = np.zeros(2, np.int32)
        _sub().orc_subgame_entry(self._h, world, out.ctypes.data)
        return int(out[0]), int(out[1])

    def entry_key(self, world):
        return int(_sub().orc_subgame_entry_key(self._h, world))

    def profile_rows(self, world):
        buf = np.zeros(4096, dtype=ROW_DTYPE)
        n = _sub().orc_subgame_export(self._h, world, buf.ctypes.data, len(buf))
        return buf[:n]

    def averaged_distribution(self, world, info_key):
        out = np.zeros(8, np.float32)
        n = _sub().orc_subgame_averaged(self._h, world, info_key, out.ctypes.data)
        return out[:n]

    def harvest(self, info_key):
        refined, visits, regret = np.zeros(4, np.float32), np.zeros(4, np.uint32), np.zeros(1, np.float32)
        n = _sub().orc_subgame_harvest(self._h, info_key, refined.ctypes.data, visits.ctypes.data, regret.ctypes.data)
        return refined[:n], visits[:n], float(regret[0])


def _deuce():
    l = lib()
    if not getattr(l, "_deuce_ready", False):
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        l.orc_eval_batch.argtypes = [vp, i64, vp]
        l.orc_river_equity_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp, ctypes.c_int]
        l._deuce_ready = True
    return l


def eval_batch(hands):
    hands = np.ascontiguousarray(hands, dtype=np.uint64)
    out = np.zeros(len(hands), dtype=np.uint32)
    _deuce().orc_eval_batch(hands.ctypes.data, len(hands), out.ctypes.data)
    return out


def river_equity_batch(pocket, public, threads=8):
    pocket = np.ascontiguousarray(pocket, dtype=np.uint64)
    public = np.ascontiguousarray(public, dtype=np.uint64)
    n = len(pocket)
    eq, bk = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    w, t = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    _deuce().orc_river_equity_batch(pocket.ctypes.data, public.ctypes.data, n, eq.ctypes.data, bk.ctypes.data, w.ctypes.data, t.ctypes.data, threads)
    return eq, bk, w, t


class OracleKmeans:
    """oracle/lloyd.hpp: Elkan k-means over W1 histograms (turn layer)."""

    def __init__(self, counts, k, threads=8):
        counts = np.ascontiguousarray(counts, dtype=np.uint8)
        self.n, self.bins = counts.shape
        self.k = k
        l = lib()
        if not getattr(l, "_lloyd_ready", False):
            vp, i32 = ctypes.c_void_p, ctypes.c_int
            l.orc_kmeans_create.restype = vp
            l.orc_kmeans_create.argtypes = [vp, i32, i32, i32, i32]
            l.orc_kmeans_destroy.argtypes = [vp]
            l.orc_kmeans_set_metric.argtypes = [vp, vp]
            l.orc_kmeans_init_pp.argtypes = [vp, ctypes.c_uint64, vp]
            l.orc_kmeans_set_centroids_from_points.argtypes = [vp, vp]
            l.orc_kmeans_init_bounds.argtypes = [vp]
            l.orc_kmeans_set_centroids.argtypes = [vp, vp]
            l.orc_kmeans_step.argtypes = [vp, vp, vp, ctypes.POINTER(ctypes.c_uint32)]
            l.orc_kmeans_state.argtypes = [vp, vp, vp, vp, vp]
            l.orc_kmeans_step_local.argtypes = [vp]
            l.orc_kmeans_step_finish.argtypes = [vp, vp, vp, ctypes.POINTER(ctypes.c_uint32)]
            l.orc_kmeans_acc.restype = ctypes.POINTER(ctypes.c_uint64)
            l.orc_kmeans_acc.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
            l.orc_kmeans_tally.restype = ctypes.POINTER(ctypes.c_uint32)
            l.orc_kmeans_tally.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
            l.orc_kmeans_centroids.argtypes = [vp, vp, vp]
            l.orc_kmeans_assign.argtypes = [vp, vp, vp]
            l.orc_kmeans_metric.argtypes = [vp, vp]
            l.orc_variation.restype = ctypes.c_float
            l.orc_variation.argtypes = [vp, vp, i32]
            l.orc_kmeans_dist_evals.restype = ctypes.c_uint64
            l.orc_kmeans_dist_evals.argtypes = [vp]
            l._lloyd_ready = True
        self._l = l
        self._h = l.orc_kmeans_create(counts.ctypes.data, self.n, self.bins, k, threads)

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.orc_kmeans_destroy(self._h)
            self._h = None

    def set_metric(self, tri):
        tri = np.ascontiguousarray(tri, dtype=np.float32)
        assert len(tri) == self.bins * (self.bins - 1) // 2
        self._l.orc_kmeans_set_metric(self._h, tri.ctypes.data)

    def init_centroids(self, seed=0):
        chosen = np.zeros(self.k, np.int32)
        self._l.orc_kmeans_init_pp(self._h, seed, chosen.ctypes.data)
        return chosen

    def set_centroids_from_points(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._l.orc_kmeans_set_centroids_from_points(self._h, idx.ctypes.data)

    def set_centroids_from_counts(self, counts):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        assert counts.shape == (self.k, self.bins)
        self._l.orc_kmeans_set_centroids(self._h, counts.ctypes.data)

    def init_bounds(self):
        self._l.orc_kmeans_init_bounds(self._h)

    def step(self):
        drift, sizes, re = np.zeros(self.k, np.float32), np.zeros(self.k, np.uint32), ctypes.c_uint32()
        self._l.orc_kmeans_step(self._h, drift.ctypes.data, sizes.ctypes.data, ctypes.byref(re))
        return drift, sizes, re.value

    def step_local(self):
        self._l.orc_kmeans_step_local(self._h)

    def exchange_arrays(self):
        """numpy views of this shard's integer accumulators (u64) and tallies (u32): all-reduce(sum) them in place."""
        n = ctypes.c_int64()
        pa = self._l.orc_kmeans_acc(self._h, ctypes.byref(n))
        acc = np.ctypeslib.as_array(pa, shape=(n.value,))
        pt = self._l.orc_kmeans_tally(self._h, ctypes.byref(n))
        tally = np.ctypeslib.as_array(pt, shape=(n.value,))
        return acc, tally

    def step_finish(self):
        drift, sizes, re = np.zeros(self.k, np.float32), np.zeros(self.k, np.uint32), ctypes.c_uint32()
        self._l.orc_kmeans_step_finish(self._h, drift.ctypes.data, sizes.ctypes.data, ctypes.byref(re))
        return drift, sizes, re.value

    def bounds(self, with_lower=False):
        a, u, st = np.zeros(self.n, np.uint32), np.zeros(self.n, np.float32), np.zeros(self.n, np.uint8)
        lo = np.zeros((self.n, self.k), np.float32) if with_lower else None
        self._l.orc_kmeans_state(self._h, a.ctypes.data, u.ctypes.data, lo.ctypes.data if with_lower else None, st.ctypes.data)
        return a, u, lo, st

    def future(self):
        c, w = np.zeros((self.k, self.bins), np.uint64), np.zeros(self.k, np.uint64)
        self._l.orc_kmeans_centroids(self._h, c.ctypes.data, w.ctypes.data)
        return c, w

    def lookup(self, with_distance=False):
        out, d = np.zeros(self.n, np.uint32), np.zeros(self.n, np.float32)
        self._l.orc_kmeans_assign(self._h, out.ctypes.data, d.ctypes.data)
        return (out, d) if with_distance else out

    def metric(self):
        tri = np.zeros(self.k * (self.k - 1) // 2, np.float32)
        self._l.orc_kmeans_metric(self._h, tri.ctypes.data)
        return tri

    def dist_evals(self):
        return self._l.orc_kmeans_dist_evals(self._h)


def variation(x, y):
    x = np.ascontiguousarray(x, dtype=np.uint32)
    y = np.ascontiguousarray(y, dtype=np.uint32)
    OracleKmeans  # ensure argtypes are set lazily through a constructed instance elsewhere
    l = lib()
    l.orc_variation.restype = ctypes.c_float
    l.orc_variation.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return l.orc_variation(x.ctypes.data, y.ctypes.data, len(x))


def _sink():
    l = lib()
    if not getattr(l, "_sink_ready", False):
        vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        l.orc_exp_c.restype = f32
        l.orc_exp_c.argtypes = [f32]
        l.orc_ln_c.restype = f32
        l.orc_ln_c.argtypes = [f32]
        l.orc_ot_cost.restype = f32
        l.orc_ot_cost.argtypes = [vp, vp, i32, vp, i32, f32, i32, f32, ctypes.POINTER(i32)]
        l.orc_sinkhorn_divergence_batch.argtypes = [vp, vp, ctypes.c_int64, i32, vp, i32, vp, i32]
        l._sink_ready = True
    return l


def exp_c(x):
    return _sink().orc_exp_c(x)


def ln_c(x):
    return _sink().orc_ln_c(x)


def ot_cost(mu, nu, tri, math=0, temperature=0.025, iterations=128, tolerance=0.0005):
    mu = np.ascontiguousarray(mu, dtype=np.uint32)
    nu = np.ascontiguousarray(nu, dtype=np.uint32)
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    it = ctypes.c_int()
    c = _sink().orc_ot_cost(mu.ctypes.data, nu.ctypes.data, len(mu), tri.ctypes.data, math, temperature, iterations, tolerance, ctypes.byref(it))
    return c, it.value


def exp_c_min_over(lo, hi):
    l = _sink()
    l.orc_exp_c_min_over.restype = ctypes.c_float
    l.orc_exp_c_min_over.argtypes = [ctypes.c_float, ctypes.c_float]
    return l.orc_exp_c_min_over(lo, hi)


def sinkhorn_divergence_batch(a, b, tri, math=0, threads=8):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    n, bins = a.shape
    out = np.zeros(n, np.float32)
    _sink().orc_sinkhorn_divergence_batch(a.ctypes.data, b.ctypes.data, n, bins, tri.ctypes.data, math, out.ctypes.data, threads)
    return out


def _iso():
    l = lib()
    if not getattr(l, "_iso_ready", False):
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        l.orc_isomorphisms.restype = i64
        l.orc_isomorphisms.argtypes = [i32, vp, vp, i64, i32]
        l.orc_canonical_batch.argtypes = [vp, vp, i64, vp, vp, vp]
        l.orc_turn_histograms.argtypes = [vp, vp, i64, vp, i32]
        l.orc_project.argtypes = [vp, vp, i64, vp, vp, vp, i64, i32, vp, i32]
        l._iso_ready = True
    return l


STREETS = {"pref": 0, "flop": 1, "turn": 2, "rive": 3}


def isomorphisms(street, cap=None, threads=8):
    """Canonical observations of a street in `IsomorphismIterator` order → (count, pocket u64[], public u64[])."""
    l = _iso()
    st = STREETS[street]
    n = l.orc_isomorphisms(st, None, None, 0, threads)
    m = n if cap is None else min(cap, n)
    pocket, public = np.zeros(m, np.uint64), np.zeros(m, np.uint64)
    if m:
        l.orc_isomorphisms(st, pocket.ctypes.data, public.ctypes.data, m, threads)
    return n, pocket, public


def canonical_batch(pocket, public):
    pocket = np.ascontiguousarray(pocket, dtype=np.uint64)
    public = np.ascontiguousarray(public, dtype=np.uint64)
    po, pu, ic = np.zeros_like(pocket), np.zeros_like(public), np.zeros(len(pocket), np.uint8)
    _iso().orc_canonical_batch(pocket.ctypes.data, public.ctypes.data, len(pocket), po.ctypes.data, pu.ctypes.data, ic.ctypes.data)
    return po, pu, ic


def turn_histograms(pocket, public, threads=8):
    pocket = np.ascontiguousarray(pocket, dtype=np.uint64)
    public = np.ascontiguousarray(public, dtype=np.uint64)
    hist = np.zeros((len(pocket), 101), np.uint8)
    _iso().orc_turn_histograms(pocket.ctypes.data, public.ctypes.data, len(pocket), hist.ctypes.data, threads)
    return hist


def project(pocket, public, next_pocket, next_public, next_abs, bins, threads=8):
    a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (pocket, public, next_pocket, next_public)]
    next_abs = np.ascontiguousarray(next_abs, dtype=np.uint8)
    hist = np.zeros((len(a[0]), bins), np.uint8)
    _iso().orc_project(a[0].ctypes.data, a[1].ctypes.data, len(a[0]), a[2].ctypes.data, a[3].ctypes.data, next_abs.ctypes.data, len(a[2]), bins,
                       hist.ctypes.data, threads)
    return hist


# ── NLHE (oracle/nlhe.hpp) ─────────────────────────────────────────────────────────────────────────
NLHE_PROBE = np.dtype([("pot", "<i2"), ("to_call", "<i2"), ("to_raise", "<i2"), ("to_shove", "<i2"),
                       ("stack", "<i2", 2), ("stake", "<i2", 2), ("spent", "<i2", 2), ("street", "i1"), ("turn", "i1"),
                       ("applied_kind", "u1"), ("pad", "u1"), ("applied_chips", "<i2"), ("abs", "<u2"), ("flags", "<u4"),
                       ("subgame", "<u8"), ("choices", "<u8"), ("board", "<u8"), ("hole", "<u8", 2)])
NLHE_ROW = np.dtype([("past", "<i8"), ("choices", "<i8"), ("edge", "<i8"), ("present", "<i2"), ("pad", "<i2", 3),
                     ("weight", "<f4"), ("regret", "<f4"), ("payoff", "<f4"), ("visits", "<i4")])
NLHE_NODE = np.dtype([("parent", "<i4"), ("edge", "u1"), ("turn", "u1"), ("pot", "<i2"), ("abs", "<u2"), ("pad", "<u2"),
                      ("pad1", "<u4"), ("subgame", "<u8"), ("choices", "<u8"), ("payoff0", "<f4"), ("pad2", "<i4")])
NLHE_FLAGS = {"must_post": 1, "must_stop": 2, "must_deal": 4, "is_everyone_alright": 8, "is_everyone_calling": 16,
              "is_everyone_touched": 32, "is_everyone_matched": 64, "is_everyone_folding": 128, "is_everyone_shoving": 256,
              "may_fold": 512, "may_call": 1024, "may_check": 2048, "may_raise": 4096, "may_shove": 8192}
NLHE_KINDS = {"Draw": 0, "Fold": 1, "Call": 2, "Check": 3, "Raise": 4, "Shove": 5}
NLHE_EDGES = {"Draw": 1, "Fold": 2, "Check": 3, "Call": 4, "Shove": 5}
NLHE_OPENS = [2, 3, 4, 5]
NLHE_RAISES = [(1, 4), (1, 3), (1, 2), (2, 3), (3, 4), (1, 1), (5, 4), (3, 2), (2, 1), (3, 1)]
_nl = None


def nlhe_edge(name, *args):
    """Edge → u8 (kicker/src/edge.rs:117-135): Open(n) → 6+idx, Raise(n, d) → 10+idx."""
    if name == "Open":
        return 6 + NLHE_OPENS.index(args[0])
    if name == "Raise":
        return 10 + NLHE_RAISES.index(tuple(args))
    return NLHE_EDGES[name]


def nlhe_edge_u64(code):
    """5-bit path code → `u64::from(Edge)` (kicker/src/edge.rs:185-197), the blueprint table's `edge` column."""
    return int(_nlhe().orc_nlhe_edge_u64(int(code)))


def nlhe_edge_from_u64(value):
    """`Edge::from(u64)` (kicker/src/edge.rs:160-183, legacy BBs form included) → 5-bit path code (0 = not on the grid)."""
    return int(_nlhe().orc_nlhe_edge_from_u64(int(value)))


def nlhe_path(edges):
    p = 0
    for i, e in enumerate(edges[:12]):
        p |= e << (5 * i)
    return p


def nlhe_unpath(p):
    out = []
    while p & 0x1F:
        out.append(p & 0x1F)
        p >>= 5
    return out


def _nlhe():
    global _nl
    if _nl is None:
        l = lib()
        vp, u64, i32, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32
        l.orc_nlhe_script.argtypes = [u64, u32, u32, i32, vp, vp, i32, vp, vp, vp]
        l.orc_nlhe_aggression.argtypes = [u64]
        l.orc_nlhe_path_push.restype = u64
        l.orc_nlhe_path_push.argtypes = [u64, i32]
        l.orc_nlhe_raises.argtypes = [i32, i32, vp]
        l.orc_nlhe_into_chips.argtypes = [i32, i32]
        l.orc_nlhe_edge_u64.restype = u64
        l.orc_nlhe_edge_u64.argtypes = [i32]
        l.orc_nlhe_edge_from_u64.restype = i32
        l.orc_nlhe_edge_from_u64.argtypes = [u64]
        l.orc_nlhe_default_regret.restype = ctypes.c_float
        l.orc_nlhe_default_regret.argtypes = [i32]
        l.orc_nlhe_deck_draw.argtypes = [ctypes.POINTER(u64), u32]
        l.orc_nlhe_abstraction.restype = ctypes.c_uint16
        l.orc_nlhe_abstraction.argtypes = [u64, u64]
        l.orc_nlhe_showdown.argtypes = [i32, vp, vp, vp, vp]
        l.orc_nlhe_create.restype = vp
        l.orc_nlhe_create.argtypes = [u64, i32, i32, i32, i32, i32]
        l.orc_nlhe_destroy.argtypes = [vp]
        l.orc_nlhe_set_hyper.argtypes = [vp, vp, u32]
        l.orc_nlhe_step.argtypes = [vp, u64]
        l.orc_nlhe_set_world.argtypes = [vp, i32, i32]
        l.orc_nlhe_sample.restype = u64
        l.orc_nlhe_sample.argtypes = [vp, vp, u64]
        l.orc_nlhe_fold.argtypes = [vp, vp, u64]
        l.orc_nlhe_owner.argtypes = [vp, i32]
        l.orc_nlhe_touched.restype = u64
        l.orc_nlhe_touched.argtypes = [vp, vp, u64]
        l.orc_nlhe_apply_rows.argtypes = [vp, vp, u64]
        l.orc_nlhe_counters.argtypes = [vp, vp]
        l.orc_nlhe_export.restype = u64
        l.orc_nlhe_export.argtypes = [vp, vp, u64]
        l.orc_nlhe_tree.argtypes = [vp, i32, vp, i32]
        l.orc_nlhe_tree_preorder.argtypes = [vp, i32, vp, i32]
        _nl = l
    return _nl


def nlhe_script(steps, seed=0, epoch=0, tree=0, snap=False):
    """steps: list of ("Call", 1) / ("Check",) / ("Draw",) / ("Raise", None) (None = the legal amount) / ("Edge", code).
    Returns (probes[n+1], won[2] or None).  Raises ValueError when a step is not `is_allowed`."""
    kinds = np.array([16 + s[1] if s[0] == "Edge" else NLHE_KINDS[s[0]] for s in steps], dtype=np.int32)
    chips = np.array([-1 if (s[0] == "Edge" or len(s) < 2 or s[1] is None) else s[1] for s in steps], dtype=np.int32)
    out = np.zeros(len(steps) + 1, dtype=NLHE_PROBE)
    won = np.full(2, -32768, dtype=np.int16)
    reward = np.zeros(2, dtype=np.int16)
    rc = _nlhe().orc_nlhe_script(seed, epoch, tree, len(steps), kinds.ctypes.data, chips.ctypes.data, 1 if snap else 0,
                                 out.ctypes.data, won.ctypes.data, reward.ctypes.data)
    if rc < 0:
        raise ValueError(f"step {-rc - 1} {steps[-rc - 1]} is not allowed")
    return out, (None if won[0] == -32768 else (won.copy(), reward.copy()))


class OracleNlhe:
    def __init__(self, seed=0, batch=128, threads=1, regret="LinearRegret", weight="LinearWeight", sampling="PluribusSampling",
                 hyper=None, warmup=None):
        self._l = _nlhe()
        self._h = self._l.orc_nlhe_create(seed, batch, threads, REGRETS[regret], WEIGHTS[weight], SAMPLERS[sampling])
        if hyper is not None or warmup is not None:
            f = np.array(hyper if hyper is not None else [1.0, 2.0, 0.05, -3e5, 0.05], dtype=np.float32)
            self._l.orc_nlhe_set_hyper(self._h, f.ctypes.data, 16384 if warmup is None else warmup)

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.orc_nlhe_destroy(self._h)
            self._h = None

    def step(self, n=1):
        self._l.orc_nlhe_step(self._h, n)

    def counters(self):
        c = np.zeros(5, dtype=np.uint64)
        self._l.orc_nlhe_counters(self._h, c.ctypes.data)
        return dict(zip(("epochs", "nodes", "infos", "updates", "rows"), (int(x) for x in c)))

    def export(self):
        n = self._l.orc_nlhe_export(self._h, None, 0)
        out = np.zeros(n, dtype=NLHE_ROW)
        self._l.orc_nlhe_export(self._h, out.ctypes.data, n)
        return out

    # multi-rank exchange (robopoker_b200.distributed.ShardedNlhe on the CPU/gloo path)
    def set_world(self, rank, world):
        self._l.orc_nlhe_set_world(self._h, rank, world)

    def sample_records(self):
        """This rank's Decisions of the epoch as an int32 matrix [count, words]."""
        words = self._l.orc_nlhe_dec_bytes() // 4
        n = self._l.orc_nlhe_sample(self._h, None, 0)
        out = np.zeros((n, words), dtype=np.int32)
        self._l.orc_nlhe_sample(self._h, out.ctypes.data, n)
        return out

    def fold_records(self, records):
        records = np.ascontiguousarray(records, dtype=np.int32)
        self._l.orc_nlhe_fold(self._h, records.ctypes.data, len(records))

    def partition_records(self, world):
        """This rank's Decisions grouped by owner rank (hash(infoset) mod world): (int32 matrix, [count per destination])."""
        recs = self.sample_records()
        owners = np.array([self._l.orc_nlhe_owner(recs[i].ctypes.data, world) for i in range(len(recs))], dtype=np.int64)
        order = np.argsort(owners, kind="stable")
        return np.ascontiguousarray(recs[order]), [int((owners == r).sum()) for r in range(world)]

    def touched_rows(self):
        words = self._l.orc_nlhe_packed_bytes() // 4
        n = self._l.orc_nlhe_touched(self._h, None, 0)
        out = np.zeros((n, words), dtype=np.int32)
        self._l.orc_nlhe_touched(self._h, out.ctypes.data, n)
        return out

    def apply_rows(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        self._l.orc_nlhe_apply_rows(self._h, rows.ctypes.data, len(rows))

    def tree_preorder(self, tree, cap=1 << 16):
        dt = np.dtype([("depth", "u1"), ("kind", "u1"), ("act", "u1"), ("pad", "u1"), ("p", "<f4"), ("q", "<f4"), ("payoff", "<f4")])
        out = np.zeros(cap, dtype=dt)
        n = self._l.orc_nlhe_tree_preorder(self._h, tree, out.ctypes.data, cap)
        return out[:n]

    def tree(self, tree, cap=1 << 16):
        out = np.zeros(cap, dtype=NLHE_NODE)
        n = self._l.orc_nlhe_tree(self._h, tree, out.ctypes.data, cap)
        return out[:n]


def nlhe_showdown(ledger):
    """ledger: list of (risked, status 'P'|'S'|'F', packed strength) → rewards (kicker/src/showdown.rs)."""
    n = len(ledger)
    risked = np.array([x[0] for x in ledger], dtype=np.int16)
    status = np.array(["PSF".index(x[1]) for x in ledger], dtype=np.uint8)
    strength = np.array([x[2] for x in ledger], dtype=np.uint32)
    reward = np.zeros(n, dtype=np.int16)
    _nlhe().orc_nlhe_showdown(n, risked.ctypes.data, status.ctypes.data, strength.ctypes.data, reward.ctypes.data)
    return reward.tolist()
