#!/usr/bin/env python
"""Regenerates tests/golden/oracle_pins.json: bit patterns / digests of the oracle's outputs on small seeded inputs
(the recipe is tests/pins.py).

The reference holds no numeric golden vectors for this path and cannot be built here (Rust), so these pins
do not come from the reference: they freeze the ORACLE (the CPU restatement every GPU parity test compares against) so
that a compiler flag, a refactor or a contract change cannot move it unnoticed, and tests/test_pins.py checks the
device library against the same committed vectors WITHOUT the oracle in the loop.  Regenerate only on a deliberate
contract change (the exp/ln contract of include/rbp.h, the Philox RNG contract) and say so in the commit.

    python tests/golden/make_oracle_pins.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if __name__ == "__main__":
    import pins
    from oracle import binding as oracle

    out = pins.compute(pins.OracleBackend(oracle))
    out.update(pins.contract_pins(oracle))
    path = os.path.join(HERE, "oracle_pins.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
