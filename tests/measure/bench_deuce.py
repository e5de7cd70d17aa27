#!/usr/bin/env python
"""Secondary benchmark: hand evaluations/s and river-equity observations/s (SURVEY §8d "river equity": ALU-bound,
990 evaluations per observation), next to the oracle on the host cores."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=2_000_000)
    p.add_argument("--cpu-n", type=int, default=20000)
    args = p.parse_args()
    import ctypes

    import numpy as np
    import torch

    import robopoker_b200 as rbp

    rng = np.random.default_rng(0)
    # random 7-card observations: argpartition of random keys picks 7 distinct cards per row
    keys = rng.random((args.n, 52), dtype=np.float32)
    cards = np.argpartition(keys, 7, axis=1)[:, :7].astype(np.uint64)
    bits = (np.uint64(1) << cards)
    pocket = bits[:, 0] | bits[:, 1]
    public = bits[:, 2] | bits[:, 3] | bits[:, 4] | bits[:, 5] | bits[:, 6]
    l = rbp.load_library()
    dp, db = torch.from_numpy(pocket.view(np.int64)).cuda(), torch.from_numpy(public.view(np.int64)).cuda()
    de = torch.empty(args.n, dtype=torch.float32, device="cuda")
    dk = torch.empty(args.n, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run():
        st = l.rbp_river_equity_device(dp.data_ptr(), db.data_ptr(), args.n, de.data_ptr(), dk.data_ptr(), None, None, stream.cuda_stream)
        assert st == 0

    run(); torch.cuda.synchronize()
    e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    line = {"bench": "river_equity", "n": args.n, "ms": ms, "observations_per_s": args.n / (ms * 1e-3),
            "hand_evals_per_s": args.n * 991 / (ms * 1e-3), "full_river_layer_s": 123_156_254 / (args.n / (ms * 1e-3))}
    from oracle import binding as oracle

    t0 = time.perf_counter()
    oracle.river_equity_batch(pocket[: args.cpu_n], public[: args.cpu_n], threads=os.cpu_count() or 1)
    dt = time.perf_counter() - t0
    line["cpu_baseline"] = {"observations_per_s": args.cpu_n / dt, "cores": os.cpu_count(), "kind": "port", "sample": f"{args.cpu_n} observations"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
