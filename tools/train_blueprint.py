#!/usr/bin/env python
"""`trainer --fast` on the GPU (crates/forge/src/{trainer,fast}.rs minus Postgres): the blueprint training loop around
`Nlhe::step` with periodic checkpoints in the reference's blueprint row format and resume.

    python tools/train_blueprint.py --epochs 200 --batch 16384 --checkpoint blueprint.npz --every 50
    python tools/train_blueprint.py --epochs 400 --resume blueprint.npz            # continues at the stored epoch

A checkpoint holds the rows `NlheProfile::rows()` would COPY into Postgres (past i64, present i16, choices i64, edge i64,
weight f32, regret f32, payoff f32, visits i32; crates/nlhe/src/profile.rs:143-160) plus the epoch counter
(`daybook::epoch()`'s `current` key, profile.rs:100-108).  Resuming is `Hydrate`: rbp_nlhe_import(rows, epochs).
tests/test_nlhe_gpu.py::test_export_import_roundtrip checks that train → checkpoint → resume → train equals training straight through.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--epochs", type=int, default=100, help="train until this epoch count")
    p.add_argument("--batch", type=int, default=16384, help="trees per epoch (`batch_size()`; the reference's macro says 128)")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--table-slots", type=int, default=1 << 24)
    p.add_argument("--checkpoint", default=None, help="write the blueprint rows here (.npz)")
    p.add_argument("--every", type=int, default=0, help="checkpoint every N epochs (0 = only at the end)")
    p.add_argument("--resume", default=None, help="hydrate from this checkpoint first")
    args = p.parse_args()
    import numpy as np

    from robopoker_b200.nlhe import Nlhe

    solver = Nlhe.flagship(batch=args.batch, seed=args.seed, table_slots=args.table_slots)
    if args.resume:
        ck = np.load(args.resume)
        solver.load(ck["rows"], int(ck["epochs"]))
        print(json.dumps({"event": "hydrate", "rows": int(len(ck["rows"])), "epochs": int(ck["epochs"])}), flush=True)

    def checkpoint():
        if not args.checkpoint:
            return
        t0 = time.perf_counter()
        rows = solver.profile()
        np.savez(args.checkpoint, rows=rows, epochs=np.int64(solver.counters()["epochs"]))
        print(json.dumps({"event": "flush", "rows": int(len(rows)), "seconds": round(time.perf_counter() - t0, 3)}), flush=True)

    done = solver.counters()["epochs"]
    while done < args.epochs:
        n = min(args.every or args.epochs, args.epochs - done)
        c0, t0 = solver.counters(), time.perf_counter()
        solver.step(n)
        c1, dt = solver.counters(), time.perf_counter() - t0
        done = c1["epochs"]
        print(json.dumps({"event": "train", "epochs": done, "trees": done * args.batch, "infosets": c1["rows"],
                          "updates_per_s": (c1["updates"] - c0["updates"]) / dt, "nodes_per_s": (c1["nodes"] - c0["nodes"]) / dt}), flush=True)
        if args.every:
            checkpoint()
    if not args.every:
        checkpoint()


if __name__ == "__main__":
    main()
