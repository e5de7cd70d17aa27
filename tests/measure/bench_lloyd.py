#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[4], turn+river EMD k-means sweep): seconds/iteration, distance
evaluations/s and achieved GB/s of the Elkan step vs the HBM roofline, next to the oracle on the host cores.

    python tests/measure/bench_lloyd.py --n 1000000 --k 100 256 500 2000 --iters 4
"""
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
    p.add_argument("--n", type=int, default=1_000_000)
    p.add_argument("--k", type=int, nargs="+", default=[100, 256, 500])
    p.add_argument("--iters", type=int, default=4)
    p.add_argument("--cpu-n", type=int, default=20000, help="points for the oracle's bounded sample")
    args = p.parse_args()
    import numpy as np
    from lloyd_data import turn_histograms

    import robopoker_b200 as rbp

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    import ctypes
    fadd = ctypes.c_float()
    rbp.load_library().rbp_measure_fadd_peak(ctypes.byref(fadd))
    pts = turn_histograms(args.n, seed=0)
    for k in args.k:
        g = rbp.lloyd.Layer(pts, k)
        t0 = time.perf_counter(); g.init_centroids(0); t_pp = time.perf_counter() - t0
        t0 = time.perf_counter(); g.init_bounds(); t_bounds = time.perf_counter() - t0
        g.step(); g.step()                      # warm-up (first step has nothing to move)
        ms_step = g.timed(0, args.iters) / args.iters
        ms_assign = g.timed(1, 3) / 3
        # algorithmic bytes of one Elkan step: point row + bounds read+write
        bytes_step = args.n * (112 + 2 * 4 * k + 9)
        line = {"bench": "lloyd_turn_w1", "n": args.n, "k": k, "iters": args.iters,
                "s_per_iteration": ms_step * 1e-3, "init_pp_s": t_pp, "init_bounds_s": t_bounds,
                "assign_sweep_ms": ms_assign, "assign_distance_evals_per_s": args.n * k / (ms_assign * 1e-3),
                "assign_fp32_tera_adds": args.n * k * 202 / (ms_assign * 1e-3) / 1e12, "measured_fadd_peak_tera_adds": fadd.value,
                "assign_frac_of_fadd_peak": args.n * k * 202 / (ms_assign * 1e-3) / 1e12 / fadd.value,
                "roofline": {"bound": "hbm", "kernel": "elkan_step_kernel", "achieved": bytes_step / (ms_step * 1e-3) / 1e9, "peak": hbm,
                             "unit": "GB/s", "frac": bytes_step / (ms_step * 1e-3) / 1e9 / hbm,
                             "note": "whole step (pairwise+step+accumulate+cdf+drift kernels) over the step kernel's algorithmic bytes"}}
        if k <= 256 and args.cpu_n:
            from oracle import binding as oracle

            sub = pts[: args.cpu_n]
            o = oracle.OracleKmeans(sub, k, threads=os.cpu_count() or 1)
            o.init_centroids(0); o.init_bounds(); o.step()
            t0 = time.perf_counter()
            for _ in range(2):
                o.step()
            dt = (time.perf_counter() - t0) / 2
            line["cpu_baseline"] = {"s_per_iteration_scaled": dt * args.n / args.cpu_n, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{args.cpu_n} points, scaled linearly to {args.n}"}
        print(json.dumps(line), flush=True)
        g.close()


if __name__ == "__main__":
    main()
