#!/usr/bin/env python
"""Transcribes the reference's solver test matrices into tests/golden/solver_combos.json:
`kuhn!(S, R, W, tol)` (crates/kuhn/src/solver.rs:234-277, exploitability after 2^18 iterations) and
`rps!(S, R, W, tol)` (crates/roshambo/src/solver.rs:205-250, |averaged − (0.4, 0.4, 0.2)| after 2^16 iterations).
Run where /root/reference exists; the tests read only the committed JSON."""
import json
import os
import re

REF = "/root/reference/crates"
HERE = os.path.dirname(os.path.abspath(__file__))


def combos(path, macro):
    out = []
    for line in open(path):
        m = re.search(macro + r"!\((\w+),\s*(\w+),\s*(\w+),\s*([0-9.]+)\)", line)
        if m:
            out.append({"sampling": m.group(1), "regret": m.group(2), "weight": m.group(3), "tolerance": float(m.group(4))})
    return out


if __name__ == "__main__":
    data = {"kuhn_exploitability_after_2^18": combos(os.path.join(REF, "kuhn/src/solver.rs"), "kuhn"),
            "rps_equilibrium_after_2^16": combos(os.path.join(REF, "roshambo/src/solver.rs"), "rps")}
    json.dump(data, open(os.path.join(HERE, "solver_combos.json"), "w"), indent=1)
    print({k: len(v) for k, v in data.items()})
