"""robopoker_b200 — B200-native (sm_100a) drop-in for robopoker's data-parallel training path.

The product is `librbp_b200.so` (hand-written CUDA behind the C ABI of `include/rbp.h`).  This package is the
thin host-side mirror of the reference's trait surface for that path, used by tests and `bench.py`:

* `Solver`  — `trait Solver` (crates/mccfr/src/solver/solver.rs:38-350): `step`, `solve`, `exploitability`, profile rows
* `subgame.WorldSolver` — `WorldSolver` / `SubGameSolver` without an origin (crates/subgame/src/world/solver.rs:33-146)

There is no CPU fallback: every compute call raises `RbpError` without a CUDA device.
"""
from ._ffi import RbpError, lib, load_library  # noqa: F401
from . import deuce  # noqa: F401
from . import lloyd  # noqa: F401
from . import subgame  # noqa: F401
from .solver import (  # noqa: F401
    FOLD_BATCHED, FOLD_ORDERED, GAMES, REGRETS, SAMPLERS, WEIGHTS, Hyper, Solver, kuhn_info, leduc_info,
)

__all__ = ["RbpError", "lib", "load_library", "Solver", "Hyper", "kuhn_info", "leduc_info",
           "GAMES", "REGRETS", "WEIGHTS", "SAMPLERS", "FOLD_ORDERED", "FOLD_BATCHED"]
