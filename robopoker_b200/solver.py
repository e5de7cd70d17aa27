"""Host-side mirror of `trait Solver` (crates/mccfr/src/solver/solver.rs:38-350) over the C ABI.

Names follow the reference: `step`, `solve(trees)`, `exploitability`, `profile` rows of `Encounter`s keyed by
(infoset, edge).  Generic parameters R, W, S of `Kuhn<R,W,S>` / `Leduc<R,W,S>` are passed by name.
"""
import ctypes

import numpy as np

from . import _ffi

GAMES = {"kuhn": 0, "leduc": 1, "rps": 2}
REGRETS = {"SummedRegret": 0, "FlooredRegret": 1, "LinearRegret": 2, "DiscountedRegret": 3, "AsymmetricRegret": 4}
WEIGHTS = {"ConstantWeight": 0, "LinearWeight": 1, "QuadraticWeight": 2, "ExponentialWeight": 3}
SAMPLERS = {"ExternalSampling": 0, "VanillaSampling": 1, "PrunableSampling": 2, "PluribusSampling": 3, "TargetedSampling": 4}
FOLD_ORDERED, FOLD_BATCHED = 0, 1

ROW_DTYPE = np.dtype([("info_key", "<u4"), ("action", "<u4"), ("weight", "<f4"), ("regret", "<f4"), ("payoff", "<f4"),
                      ("visits", "<u4")])

RANKS = {"J": 0, "Q": 1, "K": 2}
KUHN_HISTORY = {"Open": 0, "Check": 1, "Bet": 2, "CheckBet": 3}
LEDUC_SPOT = {"Open": 0, "Checked": 1, "Raised": 2, "CheckRaised": 3}


def kuhn_info(rank, history, acting=True):
    """`kuhn_info(acting, rank, node)` (crates/kuhn/src/info.rs:74-76) as the packed key of include/rbp.h."""
    return int(acting) | (KUHN_HISTORY[history] << 1) | (RANKS[rank] << 3)


def leduc_info(rank, board, r1, r2, acting=True):
    """`leduc_info(acting, rank, board, r1, r2)` (crates/leduc/src/info.rs:92-94) as the packed key."""
    b = 0 if board is None else 1 + RANKS[board]
    s2 = 0 if r2 is None else 1 + LEDUC_SPOT[r2]
    return int(acting) | (b << 1) | (LEDUC_SPOT[r1] << 3) | (s2 << 5) | (RANKS[rank] << 8)


class Hyper:
    """Sampling/Pruning/Training hyper-parameter singletons (crates/mccfr/src/hyperparams/*.rs) as one value."""

    def __init__(self, **kw):
        self.c = _ffi.HyperC()
        _ffi.lib().rbp_hyper_default(ctypes.byref(self.c))
        for k, v in kw.items():
            if not hasattr(self.c, k):
                raise AttributeError(k)
            setattr(self.c, k, v)


class Solver:
    def __init__(self, game, regret="FlooredRegret", weight="LinearWeight", sampling="ExternalSampling", batch=1, seed=0,
                 fold=FOLD_ORDERED, hyper=None, device=0):
        self._lib = _ffi.lib()
        self._h = ctypes.c_void_p()
        self.game = game
        self.batch = int(batch)
        hp = ctypes.byref(hyper.c) if hyper is not None else None
        _ffi.check(self._lib.rbp_solver_create(GAMES[game], REGRETS[regret], WEIGHTS[weight], SAMPLERS[sampling], fold,
                                               self.batch, seed, hp, device, ctypes.byref(self._h)), "rbp_solver_create")

    @classmethod
    def kuhn(cls, **kw):
        return cls("kuhn", **kw)

    @classmethod
    def leduc(cls, **kw):
        return cls("leduc", **kw)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.rbp_solver_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_world(self, rank, size):
        _ffi.check(self._lib.rbp_solver_set_world(self._h, rank, size), "rbp_solver_set_world")

    def set_stream(self, cuda_stream):
        _ffi.check(self._lib.rbp_solver_set_stream(self._h, ctypes.c_void_p(cuda_stream)), "rbp_solver_set_stream")

    def step(self, n=1):
        """`Solver::step` × n."""
        _ffi.check(self._lib.rbp_solver_step(self._h, n), "rbp_solver_step")
        return self

    def attach_comm(self, comm):
        """Tree-sharded epochs inside the library (BATCHED fold): `step()` then all-gathers the partial sums itself."""
        _ffi.check(self._lib.rbp_solver_attach_comm(self._h, comm._h), "rbp_solver_attach_comm")
        self._comm = comm
        return self

    def spend(self, seconds):
        """`Solver::spend` (solver.rs:130-137): steps until the wall-clock budget is used; returns (epochs, elapsed seconds)."""
        n, dt = ctypes.c_uint64(), ctypes.c_double()
        _ffi.check(self._lib.rbp_solver_spend(self._h, float(seconds), ctypes.byref(n), ctypes.byref(dt)), "rbp_solver_spend")
        return int(n.value), float(dt.value)

    def step_timed(self, n=1, flush_l2=True):
        """`step` × n with CUDA-event timing on the library stream; returns (total_ms, sample_ms, fold_ms)."""
        t, a, b = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        _ffi.check(self._lib.rbp_solver_step_timed(self._h, n, int(flush_l2), ctypes.byref(t), ctypes.byref(a), ctypes.byref(b)),
                   "rbp_solver_step_timed")
        return t.value, a.value, b.value

    def sample(self):
        """BATCHED fold, first half: sample this rank's trees and reduce them to blocked partial sums (device)."""
        _ffi.check(self._lib.rbp_solver_sample(self._h), "rbp_solver_sample")

    def delta_buffer(self):
        """(device pointer, bytes) of this rank's partial sums — what ranks all-gather."""
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _ffi.check(self._lib.rbp_solver_delta_buffer(self._h, ctypes.byref(p), ctypes.byref(n)), "rbp_solver_delta_buffer")
        return p.value, n.value

    def fold_gathered(self, dev_ptr, world_size):
        """BATCHED fold, second half: rank-ordered sum of the gathered partials + one schedule application per row."""
        _ffi.check(self._lib.rbp_solver_fold_gathered(self._h, ctypes.c_void_p(dev_ptr), world_size), "rbp_solver_fold_gathered")

    def solve(self, trees):
        """`Solver::solve(trees)`: trees / batch_size steps (solver.rs:111-122)."""
        return self.step(trees // self.batch)

    @property
    def epochs(self):
        out = ctypes.c_uint64()
        _ffi.check(self._lib.rbp_solver_epochs(self._h, ctypes.byref(out)), "rbp_solver_epochs")
        return out.value

    def exploitability(self):
        out = ctypes.c_float()
        _ffi.check(self._lib.rbp_solver_exploitability(self._h, ctypes.byref(out)), "rbp_solver_exploitability")
        return out.value

    def counters(self):
        out = (ctypes.c_uint64 * 3)()
        _ffi.check(self._lib.rbp_solver_counters(self._h, out), "rbp_solver_counters")
        return {"nodes": out[0], "infos": out[1], "updates": out[2]}

    def game_shape(self):
        out = (ctypes.c_int * 6)()
        _ffi.check(self._lib.rbp_solver_game_shape(self._h, out), "rbp_solver_game_shape")
        return dict(zip(["nodes", "terminals", "infosets", "rows", "max_tree_nodes", "max_tree_infos"], list(out)))

    def profile_rows(self, out=None):
        """Bulk `RefProf::cum_*`: structured array sorted by (info_key, action)."""
        cap = self.game_shape()["rows"]
        buf = out if out is not None else np.zeros(cap, dtype=ROW_DTYPE)
        n = ctypes.c_int()
        ptr = buf.ctypes.data_as(ctypes.POINTER(_ffi.ProfileRow))
        _ffi.check(self._lib.rbp_profile_export(self._h, ptr, len(buf), ctypes.byref(n)), "rbp_profile_export")
        return buf[: n.value]

    def import_rows(self, rows, epochs):
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        ptr = rows.ctypes.data_as(ctypes.POINTER(_ffi.ProfileRow))
        _ffi.check(self._lib.rbp_profile_import(self._h, ptr, len(rows), epochs), "rbp_profile_import")

    def averaged_distribution(self, info_key):
        probs = (ctypes.c_float * 8)()
        n = ctypes.c_int()
        _ffi.check(self._lib.rbp_profile_averaged(self._h, info_key, probs, 8, ctypes.byref(n)), "rbp_profile_averaged")
        return [probs[i] for i in range(n.value)]
