"""Host-side mirror of the reference's safe subgame solver on the small games.

`WorldSolver` (crates/subgame/src/world/solver.rs:33-146) — which is also what `SubGameSolver::new` (crates/subgame/src/solver.rs:46-70,
origin = None) runs: the depth-limited frontier machinery is not built.  Names and argument meaning follow the reference:

    blueprint = rbp.Solver("kuhn", "FlooredRegret", "LinearWeight", "ExternalSampling").solve(1 << 18)
    belief    = rbp.subgame.partition(prior, worlds=2)                       # Posterior::partition::<2>()
    solver    = rbp.subgame.WorldSolver(blueprint, external=1, belief=belief, cards=(c0, c1), path=[])   # CfrRecall::new(descents, game)
    solver.solve(1 << 16); solver.harvest(info_key)
"""
import ctypes

import numpy as np

from . import _ffi
from .solver import ROW_DTYPE


def partition(reach, worlds):
    """`Partition::partition::<W>` (world/partition.rs:27-53): reach per secret (ascending secret order) -> (world of every secret, weights)."""
    reach = np.ascontiguousarray(reach, dtype=np.float32)
    world_of, weights = np.zeros(len(reach), np.int32), np.zeros(worlds, np.float32)
    _ffi.check(_ffi.lib().rbp_subgame_partition(reach.ctypes.data, len(reach), worlds, world_of.ctypes.data, weights.ctypes.data),
               "rbp_subgame_partition")
    return world_of, weights


def posterior(game, blueprint_rows, external, cards, path=()):
    """Reach of every rank the external player could hold, conditioned on the path (`Solver::external_reach` per card, `Posterior::add` per
    rank); `blueprint_rows` = `Solver.profile_rows()` of the blueprint.  Host arithmetic."""
    from .solver import GAMES
    rows = np.ascontiguousarray(blueprint_rows, dtype=ROW_DTYPE)
    p = np.ascontiguousarray(list(path), dtype=np.uint8)
    out = np.zeros(3, np.float32)
    ptr = rows.ctypes.data_as(ctypes.POINTER(_ffi.ProfileRow))
    _ffi.check(_ffi.lib().rbp_subgame_posterior(GAMES[game], ptr, len(rows), int(external), int(cards[0]), int(cards[1]),
                                                p.ctypes.data if len(p) else None, len(p), out.ctypes.data), "rbp_subgame_posterior")
    return out


def entries(game, external, world_of_rank, worlds, cards, path=()):
    """Host half of `WorldSolver::new` (no device): per world the restricted deal, the flat node and the infoset key of the entry state."""
    from .solver import GAMES
    m = None if world_of_rank is None else np.ascontiguousarray(world_of_rank, dtype=np.int32)
    p = np.ascontiguousarray(list(path), dtype=np.uint8)
    c, n, k = np.zeros(16, np.int32), np.zeros(8, np.int32), np.zeros(8, np.uint32)
    _ffi.check(_ffi.lib().rbp_subgame_entries(GAMES[game], int(external), int(worlds), None if m is None else m.ctypes.data, int(cards[0]), int(cards[1]),
                                              p.ctypes.data if len(p) else None, len(p), c.ctypes.data, n.ctypes.data, k.ctypes.data), "rbp_subgame_entries")
    return [(int(c[2 * w]), int(c[2 * w + 1])) for w in range(worlds)], n[:worlds].copy(), k[:worlds].copy()


class WorldSolver:
    def __init__(self, blueprint, external, belief, cards, path=(), seed=0):
        """belief = (world_of_rank [3] or None, weights [W]) as `partition` returns it; cards = the observed deal (card = 2 * rank + suit);
        path = branch indices from the dealt root to the entry state."""
        self._lib = _ffi.lib()
        self._h = ctypes.c_void_p()
        world_of, weights = belief
        self.worlds = len(weights)
        w = np.ascontiguousarray(weights, dtype=np.float32)
        m = None if world_of is None else np.ascontiguousarray(world_of, dtype=np.int32)
        p = np.ascontiguousarray(list(path), dtype=np.uint8)
        _ffi.check(self._lib.rbp_subgame_create(blueprint._h, int(external), self.worlds, None if m is None else m.ctypes.data, w.ctypes.data,
                                                int(cards[0]), int(cards[1]), p.ctypes.data if len(p) else None, len(p), int(seed),
                                                ctypes.byref(self._h)), "rbp_subgame_create")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.rbp_subgame_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def step(self, n=1):
        _ffi.check(self._lib.rbp_subgame_step(self._h, int(n)), "rbp_subgame_step")
        return self

    solve = step  # batch_size() = 1: `solve(trees)` is `trees` steps

    def spend(self, seconds):
        """`Solver::spend`: steps until the wall-clock budget is used; returns (steps, elapsed seconds)."""
        n, dt = ctypes.c_uint64(), ctypes.c_double()
        _ffi.check(self._lib.rbp_subgame_spend(self._h, float(seconds), ctypes.byref(n), ctypes.byref(dt)), "rbp_subgame_spend")
        return int(n.value), float(dt.value)

    def info(self):
        steps = ctypes.c_uint64()
        drawn, cards = np.zeros(8, np.uint64), np.zeros(16, np.int32)
        _ffi.check(self._lib.rbp_subgame_info(self._h, ctypes.byref(steps), drawn.ctypes.data, cards.ctypes.data), "rbp_subgame_info")
        return {"t": int(steps.value), "drawn": drawn[: self.worlds], "entries": [tuple(int(c) for c in cards[2 * w: 2 * w + 2]) for w in range(self.worlds)]}

    def profile_rows(self, world):
        buf = np.zeros(4096, dtype=ROW_DTYPE)
        n = ctypes.c_int()
        ptr = buf.ctypes.data_as(ctypes.POINTER(_ffi.ProfileRow))
        _ffi.check(self._lib.rbp_subgame_export(self._h, int(world), ptr, len(buf), ctypes.byref(n)), "rbp_subgame_export")
        return buf[: n.value]

    def averaged_distribution(self, world, info_key):
        probs = (ctypes.c_float * 8)()
        n = ctypes.c_int()
        _ffi.check(self._lib.rbp_subgame_averaged(self._h, int(world), int(info_key), probs, 8, ctypes.byref(n)), "rbp_subgame_averaged")
        return np.array([probs[i] for i in range(n.value)], np.float32)

    def harvest(self, info_key):
        """`Harvest::harvest(base)`: (refined policy, visits per edge, positive regret)."""
        refined, visits, regret = np.zeros(4, np.float32), np.zeros(4, np.uint32), np.zeros(1, np.float32)
        n = ctypes.c_int()
        _ffi.check(self._lib.rbp_subgame_harvest(self._h, int(info_key), refined.ctypes.data, visits.ctypes.data, regret.ctypes.data, 4,
                                                 ctypes.byref(n)), "rbp_subgame_harvest")
        return refined[: n.value], visits[: n.value], float(regret[0])
