#!/usr/bin/env python
"""Leduc exploitability vs iterations (BASELINE.json `metric`, configs[1]).

Runs `Leduc<R, LinearWeight, ExternalSampling>::solve(trees)` on the GPU and, in lockstep, the oracle (the C++
restatement of the reference's CPU path) with the same Philox seed, sampling `Solver::exploitability()` at every
power-of-two epoch.  One JSON line per (batch, checkpoint): both exploitabilities, whether they are bit-identical, and
the wall-clock of each side up to that checkpoint.  batch=1 x 2^20 epochs is the reference-faithful run
(`batch_size() == 1`, crates/leduc/src/solver.rs); larger batches are the throughput configurations.

    python tests/measure/leduc_curve.py --trees 1048576 --batch 1 1024 16384
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--trees", type=int, default=1 << 20, help="total sampled trees (the reference's 'iterations' at batch 1)")
    p.add_argument("--batch", type=int, nargs="+", default=[1, 1024, 16384])
    p.add_argument("--regret", default="FlooredRegret")
    p.add_argument("--fold", default="ordered", choices=["ordered", "batched"])
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--no-oracle", action="store_true")
    p.add_argument("--oracle-max-trees", type=int, default=1 << 24, help="skip the oracle for runs above this many trees")
    args = p.parse_args()
    import numpy as np

    import robopoker_b200 as rbp
    from robopoker_b200 import solver as S

    oracle = None
    if not args.no_oracle:
        from oracle import binding as oracle
    fold = S.FOLD_BATCHED if args.fold == "batched" else S.FOLD_ORDERED
    for batch in args.batch:
        epochs_total = max(1, args.trees // batch)
        g = rbp.Solver("leduc", args.regret, "LinearWeight", "ExternalSampling", batch=batch, seed=args.seed, fold=fold)
        o = None
        if oracle is not None and args.trees <= args.oracle_max_trees:
            o = oracle.OracleSolver("leduc", args.regret, "LinearWeight", "ExternalSampling", batch=batch, seed=args.seed,
                                    threads=min(os.cpu_count() or 1, max(1, batch // 64)))
            o.set_fold(fold)
        done, t_gpu, t_cpu = 0, 0.0, 0.0
        marks = sorted({min(1 << k, epochs_total) for k in range(0, epochs_total.bit_length() + 1)})
        for mark in marks:
            t0 = time.perf_counter(); g.step(mark - done); eg = g.exploitability(); t_gpu += time.perf_counter() - t0
            line = {"curve": "leduc_exploitability", "regret": args.regret, "fold": args.fold, "batch": batch, "epochs": mark,
                    "trees": mark * batch, "exploitability_gpu": eg, "gpu_seconds": t_gpu}
            if o is not None:
                t0 = time.perf_counter(); o.step(mark - done); eo = o.exploitability(); t_cpu += time.perf_counter() - t0
                line.update({"exploitability_oracle": eo, "oracle_seconds": t_cpu,
                             "bit_identical": bool(np.float32(eg).view(np.uint32) == np.float32(eo).view(np.uint32))})
            done = mark
            print(json.dumps(line), flush=True)
        if o is not None:
            same = g.profile_rows().tobytes() == o.profile_rows().tobytes()
            print(json.dumps({"curve": "leduc_exploitability", "batch": batch, "final_tables_bit_identical": bool(same),
                              "reference_threshold": "exploitability < 0.080 at 2^18 iterations (crates/leduc/src/solver.rs:121-123)"}), flush=True)
        g.close()


if __name__ == "__main__":
    main()
