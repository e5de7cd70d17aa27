"""Host-side mirror of `Nlhe<R, W, S>` (crates/nlhe/src/solver.rs:11, `Flagship` = crates/nlhe/src/lib.rs:86-90) over the C ABI.

`Nlhe(...)` defaults to the flagship configuration (LinearRegret, LinearWeight, PluribusSampling, batch 128).
`profile()` returns the rows of the reference's blueprint table (crates/nlhe/src/profile.rs:143-160:
past, present, choices, edge, weight, regret, payoff, visits)."""
import ctypes

import numpy as np

from . import _ffi
from .solver import REGRETS, SAMPLERS, WEIGHTS

ROW_DTYPE = np.dtype([("past", "<i8"), ("choices", "<i8"), ("edge", "<i8"), ("present", "<i2"), ("pad", "<i2", 3),
                      ("weight", "<f4"), ("regret", "<f4"), ("payoff", "<f4"), ("visits", "<u4")])
NODE_DTYPE = np.dtype([("depth", "u1"), ("kind", "u1"), ("act", "u1"), ("pad", "u1"), ("p", "<f4"), ("q", "<f4"), ("payoff", "<f4")])
# 5-bit codes inside `past` / `choices` Paths (kicker/src/edge.rs:117-135; Open(n) 6.., Raise(odds) 10..) ...
EDGES = {"Draw": 1, "Fold": 2, "Check": 3, "Call": 4, "Shove": 5}
# ... and the `edge` COLUMN of a blueprint row, which is `u64::from(Edge)` (kicker/src/edge.rs:185-197):
EDGE_COLUMN = {"Draw": 0, "Fold": 1, "Check": 2, "Call": 3, "Shove": 5}  # Raise(n/d) = 4 | n << 3 | d << 11, Open(n) = 6 | n << 3


class Nlhe:
    def __init__(self, regret="LinearRegret", weight="LinearWeight", sampling="PluribusSampling", batch=128, seed=0, hyper=None,
                 table_slots=1 << 22, max_nodes_per_tree=4096, device=0):
        self._lib = _ffi.lib()
        self._h = ctypes.c_void_p()
        self.batch = int(batch)
        hp = ctypes.byref(hyper.c) if hyper is not None else None
        _ffi.check(self._lib.rbp_nlhe_create(REGRETS[regret], WEIGHTS[weight], SAMPLERS[sampling], self.batch, seed, hp, table_slots,
                                             max_nodes_per_tree, device, ctypes.byref(self._h)), "rbp_nlhe_create")

    @classmethod
    def flagship(cls, **kw):
        return cls(**kw)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.rbp_nlhe_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_world(self, rank, size):
        _ffi.check(self._lib.rbp_nlhe_set_world(self._h, rank, size), "rbp_nlhe_set_world")

    def attach_comm(self, comm):
        """Sharded `Solver::step` inside the library (`robopoker_b200.comm.Comm`); collective."""
        _ffi.check(self._lib.rbp_nlhe_attach_comm(self._h, comm._h), "rbp_nlhe_attach_comm")
        self._comm = comm  # keep it alive
        return self

    def set_stream(self, cuda_stream):
        _ffi.check(self._lib.rbp_nlhe_set_stream(self._h, ctypes.c_void_p(cuda_stream)), "rbp_nlhe_set_stream")

    def step(self, n=1):
        """`Solver::step` × n."""
        _ffi.check(self._lib.rbp_nlhe_step(self._h, n), "rbp_nlhe_step")
        return self

    def spend(self, seconds):
        """`Solver::spend` (crates/mccfr/src/solver/solver.rs:130-137); returns (epochs, elapsed seconds)."""
        n, dt = ctypes.c_uint64(), ctypes.c_double()
        _ffi.check(self._lib.rbp_nlhe_spend(self._h, float(seconds), ctypes.byref(n), ctypes.byref(dt)), "rbp_nlhe_spend")
        return int(n.value), float(dt.value)

    def solve(self, trees):
        return self.step(trees // self.batch)

    def step_timed(self, n=1, flush_l2=True):
        """Returns device ms (total, tree build, value kernels, resolve+sort, fold, records exchange, rows exchange, -) summed
        over n epochs; the last three are zero without a communicator."""
        ms = (ctypes.c_float * 8)()
        _ffi.check(self._lib.rbp_nlhe_step_timed(self._h, n, int(flush_l2), ms), "rbp_nlhe_step_timed")
        return tuple(ms)

    def counters(self):
        out = (ctypes.c_uint64 * 8)()
        _ffi.check(self._lib.rbp_nlhe_counters(self._h, out), "rbp_nlhe_counters")
        return dict(zip(("epochs", "nodes", "infos", "updates", "rows", "records", "max_tree"), (int(x) for x in out)))

    def traffic_counters(self):
        out = (ctypes.c_uint64 * 4)()
        _ffi.check(self._lib.rbp_nlhe_traffic_counters(self._h, out), "rbp_nlhe_traffic_counters")
        return dict(zip(("walker_nodes", "walker_choices", "opponent_nodes", "opponent_choices"), (int(x) for x in out)))

    def profile(self):
        n = ctypes.c_uint64()
        _ffi.check(self._lib.rbp_nlhe_export(self._h, None, 0, ctypes.byref(n)), "rbp_nlhe_export")
        rows = np.zeros(n.value, dtype=ROW_DTYPE)
        _ffi.check(self._lib.rbp_nlhe_export(self._h, rows.ctypes.data, n.value, ctypes.byref(n)), "rbp_nlhe_export")
        return rows

    def load(self, rows, epochs):
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        _ffi.check(self._lib.rbp_nlhe_import(self._h, rows.ctypes.data, len(rows), epochs), "rbp_nlhe_import")

    def set_lookup(self, isoset):
        """Install one street's abstraction table (`NlheEncoder`'s BTreeMap) from a `robopoker_b200.deuce.IsoSet` whose
        abstraction column is filled (river equity buckets, or the k-means assignments of that street)."""
        _ffi.check(self._lib.rbp_nlhe_set_lookup(self._h, isoset._h), "rbp_nlhe_set_lookup")

    def set_lookup_rows(self, obs, abs_):
        """The same from the reference's `isomorphism` table rows (obs i64, abs i16) of ONE street."""
        obs = np.ascontiguousarray(obs, dtype=np.int64)
        abs_ = np.ascontiguousarray(abs_, dtype=np.int16)
        assert len(obs) == len(abs_)
        _ffi.check(self._lib.rbp_nlhe_set_lookup_rows(self._h, obs.ctypes.data, abs_.ctypes.data, len(obs)), "rbp_nlhe_set_lookup_rows")

    # multi-GPU exchange (robopoker_b200.distributed.ShardedNlhe)
    def sample(self):
        _ffi.check(self._lib.rbp_nlhe_sample(self._h), "rbp_nlhe_sample")

    def records(self):
        p, n, cap, w = ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int()
        _ffi.check(self._lib.rbp_nlhe_records(self._h, ctypes.byref(p), ctypes.byref(n), ctypes.byref(cap), ctypes.byref(w)), "rbp_nlhe_records")
        return p.value, n.value, cap.value, w.value

    def fold_records(self, dev_ptr, count):
        _ffi.check(self._lib.rbp_nlhe_fold_records(self._h, ctypes.c_void_p(dev_ptr), count), "rbp_nlhe_fold_records")

    def partition_records(self, world):
        """This rank's records grouped by owner rank: (device pointer, [count per destination rank])."""
        p, counts = ctypes.c_void_p(), (ctypes.c_uint64 * world)()
        _ffi.check(self._lib.rbp_nlhe_partition_records(self._h, ctypes.byref(p), counts), "rbp_nlhe_partition_records")
        return p.value, [int(c) for c in counts]

    def touched_rows(self):
        p, n, w = ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_int()
        _ffi.check(self._lib.rbp_nlhe_touched_rows(self._h, ctypes.byref(p), ctypes.byref(n), ctypes.byref(w)), "rbp_nlhe_touched_rows")
        return p.value, n.value, w.value

    def apply_rows(self, dev_ptr, count):
        _ffi.check(self._lib.rbp_nlhe_apply_rows(self._h, ctypes.c_void_p(dev_ptr), count), "rbp_nlhe_apply_rows")

    def debug_tree(self, tree, cap=16384):
        out = np.zeros(cap, dtype=NODE_DTYPE)
        n = ctypes.c_int()
        _ffi.check(self._lib.rbp_nlhe_debug_tree(self._h, tree, out.ctypes.data, cap, ctypes.byref(n)), "rbp_nlhe_debug_tree")
        return out[:n.value]
