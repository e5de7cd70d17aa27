#!/usr/bin/env python
"""bench.py — headline benchmark of the MCCFR hot path (BASELINE.json configs[1]: Leduc MCCFR).

One step = one `Solver::step` epoch over a batch of externally-sampled Leduc trees (deals are synthetic: Philox
draws).  `value` = infoset-action regret updates per second, whole job, device-timed with the table resident on the
GPU; `e2e` = the same metric through the host-buffer API (profile import → step → profile export every step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "infoset-action updates/sec"
UNIT = "updates/s"
GAME, REGRET, WEIGHT, SAMPLING = "leduc", "FlooredRegret", "LinearWeight", "ExternalSampling"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="leduc", choices=["leduc", "nlhe"],
                   help="leduc = BASELINE.json configs[1] (default); nlhe = configs[3]: heads-up NLHE blueprint MCCFR, synthetic abstraction")
    p.add_argument("--batch", type=int, default=None, help="trees per epoch per GPU (default: 262144 leduc, 16384 nlhe)")
    p.add_argument("--table-slots", type=int, default=1 << 24, help="nlhe: infoset table capacity (power of two)")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--skip-cpu-baseline", action="store_true", help="tuning runs only: leave out the oracle's bounded CPU sample")
    p.add_argument("--fold", default="batched", choices=["ordered", "batched"],
                   help="ordered = reference Solver::step semantics (serial per row); batched = blocked delta sums (scales across GPUs)")
    a = p.parse_args()
    if a.batch is None:
        a.batch = 262144 if a.workload == "leduc" else 16384
    return a


def workload(args, n):
    return {"workload": f"configs[1] Leduc MCCFR ({REGRET},{WEIGHT},{SAMPLING}), {args.batch} trees/epoch/GPU, {args.fold} fold",
            "fold": args.fold,
            "game": GAME, "trees_per_epoch_per_gpu": args.batch, "global_batch": args.batch * n, "parallelism": f"trees x{n}",
            "table_rows": 240, "l2": "flushed between steps (192 MiB write), untimed"}


class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline_run(args, epochs, threads):
    """The oracle (C++ restatement of the reference's rayon path) timed on the host cores: bounded sample."""
    from oracle import binding as oracle

    o = oracle.OracleSolver(GAME, REGRET, WEIGHT, SAMPLING, batch=args.batch, seed=args.seed, threads=threads)
    o.set_fold(1 if args.fold == "batched" else 0)
    o.step(1)
    u0 = o.counters()["updates"]
    t0 = time.perf_counter()
    o.step(epochs)
    dt = time.perf_counter() - t0
    return (o.counters()["updates"] - u0) / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import binding as oracle

    o = oracle.OracleSolver(GAME, REGRET, WEIGHT, SAMPLING, batch=args.batch, seed=args.seed, threads=threads)
    o.set_fold(1 if args.fold == "batched" else 0)
    o.step(args.warmup)
    u0 = o.counters()["updates"]
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    val = (o.counters()["updates"] - u0) / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload(args, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} epochs x {args.batch} trees, C++ restatement of the reference rayon path (Rust toolchain absent)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


NLHE = ("LinearRegret", "LinearWeight", "PluribusSampling")  # `Flagship` (crates/nlhe/src/lib.rs:86-90)


def nlhe_workload(args, n):
    return {"workload": f"configs[3] heads-up NLHE blueprint MCCFR (Flagship: {','.join(NLHE)}), {args.batch} trees/epoch/GPU, "
                        "ordered fold, synthetic hash abstraction 169/256/256/101 (SURVEY 8d config 4)",
            "game": "nlhe-hu", "trees_per_epoch_per_gpu": args.batch, "global_batch": args.batch * n, "parallelism": f"trees x{n}, infosets owned by hash mod {n}" if n > 1 else "trees x1",
            "table_slots": args.table_slots, "l2": "flushed between steps (192 MiB write), untimed"}


def nlhe_reference(args):
    """The oracle's NLHE solver (C++ restatement of the reference's rayon path) on all host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import binding as oracle

    threads = os.cpu_count() or 1
    o = oracle.OracleNlhe(seed=args.seed, batch=args.batch, threads=threads, regret=NLHE[0], weight=NLHE[1], sampling=NLHE[2])
    o.step(min(args.warmup, 1))
    u0 = o.counters()["updates"]
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    val = (o.counters()["updates"] - u0) / dt
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": nlhe_workload(args, 1),
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                       "sample": f"{args.steps} epochs x {args.batch} trees, C++ restatement of the reference rayon path (Rust toolchain absent)"},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def main_nlhe(args):
    import robopoker_b200 as rbp
    from robopoker_b200.nlhe import Nlhe

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl")
        dist = dist_mod
    l = rbp.load_library()
    if l.rbp_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — librbp_b200 has no CPU fallback")
    s = Nlhe(regret=NLHE[0], weight=NLHE[1], sampling=NLHE[2], batch=args.batch, seed=args.seed, table_slots=args.table_slots, device=local)
    warm = max(args.warmup, 3)
    phases = None
    if world == 1:
        s.step_timed(warm, flush_l2=True)
        c0 = s.counters()
        clocks = Clocks(local)
        l0 = l.rbp_kernel_launches()
        phases = s.step_timed(args.steps, flush_l2=True)
        ms_total = phases[0]
        gpu_launches = l.rbp_kernel_launches() - l0
        clk = clocks.stop()
        c1 = s.counters()
        updates, nodes, records = c1["updates"] - c0["updates"], c1["nodes"] - c0["nodes"], c1["records"]
        stepper = lambda: s.step(1)  # noqa: E731
    else:
        import torch
        from robopoker_b200.distributed import ShardedNlhe

        stream = torch.cuda.current_stream()
        s.set_stream(stream.cuda_stream)
        sh = ShardedNlhe(s, dist, device=local)
        flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
        sh.step(warm)
        c0 = s.counters()
        dist.barrier(); torch.cuda.synchronize()
        clocks = Clocks(local)
        l0 = l.rbp_kernel_launches()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            flush.zero_()
            e0.record(stream)
            sh.step(1)  # sample -> records to their owner ranks (NCCL all-to-all) -> fold -> all-gather of the touched rows
            e1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        ms_total = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        gpu_launches = l.rbp_kernel_launches() - l0
        clk = clocks.stop()
        c1 = s.counters()
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        # owner-sharded fold: every rank folds the infosets it owns, so the job's updates are the sum over ranks
        t = torch.tensor([c1["updates"] - c0["updates"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        updates, nodes, records = float(t.item()), None, c1["records"]
        stepper = lambda: sh.step(1)  # noqa: E731
    value = updates / (ms_total * 1e-3)

    # end to end through the public call: step + telemetry read-back per step, wall clock.  The path has no per-step host
    # input (deals come from the device-side Philox contract); the host reads the 64-byte counter block every step.
    e_steps = max(5, min(args.steps, 50))
    u0 = s.counters()["updates"]
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        stepper()
        s.counters()
    e_dt = time.perf_counter() - t0
    e_updates = s.counters()["updates"] - u0
    if dist:
        import torch
        t = torch.tensor([e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_dt = float(t.item())
        t = torch.tensor([e_updates], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        e_updates = float(t.item())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = None
    if phases is not None:
        # dominant phase: the value kernels (child tasks of the large roots + small roots + combine).  Algorithmic bytes per
        # epoch: every preorder node read once (16 B) and one 72-byte update record written per walker node — the scans
        # re-read nodes once per walker ancestor from L1/L2.
        k_s = phases[2] * 1e-3 / args.steps
        traffic = None  # the committed ncu capture of this kernel (profiles/r1k_nlhe_value_ncu.txt) predates the split value phase
        alg = (16.0 * nodes + 72.0 * records * args.steps) / args.steps
        roofline = {"bound": "hbm", "kernel": "nlhe_value_kernel", "achieved": alg / k_s / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": alg / k_s / 1e9 / peak, "peak_source": "measured" if peaks else "fallback", "traffic": traffic,
                    "kernel_ms": {"tree_build": phases[1] / args.steps, "value": phases[2] / args.steps,
                                  "resolve_sort": phases[3] / args.steps, "fold": phases[4] / args.steps},
                    "note": "divergent tree walks and serial per-infoset chains: latency-bound, reported (not claimed) against the HBM roofline (SURVEY 8d)"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": nlhe_workload(args, world), "clocks": clk,
                "e2e": {"value": e_updates / e_dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64, "steps": e_steps,
                        "note": "no per-step host input exists on this path (device-side deals); the host reads the counter block every step"},
                "gpu_launches": int(gpu_launches), "roofline": roofline, "epochs": c1["epochs"] + e_steps, "table_rows": s.counters()["rows"]}
        if world == 1 and not args.skip_cpu_baseline:
            from oracle import binding as oracle

            threads = os.cpu_count() or 1
            cpu_epochs = max(2, int(4.0e5 // args.batch))
            o = oracle.OracleNlhe(seed=args.seed, batch=args.batch, threads=threads, regret=NLHE[0], weight=NLHE[1], sampling=NLHE[2])
            o.step(1)
            u0 = o.counters()["updates"]
            t0 = time.perf_counter()
            o.step(cpu_epochs)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": (o.counters()["updates"] - u0) / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{cpu_epochs} epochs x {args.batch} trees in {dt:.1f}s, C++ restatement of the reference rayon path"}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "nlhe":
        return nlhe_reference(args) if args.impl == "reference" else main_nlhe(args)
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np

    import robopoker_b200 as rbp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl")
        dist = dist_mod
    l = rbp.load_library()
    if l.rbp_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — librbp_b200 has no CPU fallback")

    fold = rbp.FOLD_BATCHED if args.fold == "batched" else rbp.FOLD_ORDERED
    if world > 1 and args.fold != "batched":
        raise SystemExit("bench.py: --gpus > 1 needs --fold batched (the ordered fold is serial per row and does not shard)")
    s = rbp.Solver(GAME, REGRET, WEIGHT, SAMPLING, batch=args.batch, seed=args.seed, device=local, fold=fold)
    warm = max(args.warmup, 3)
    if world == 1:
        s.step_timed(warm, flush_l2=True)
        u0 = s.counters()["updates"]
        clocks = Clocks(local)
        l0 = l.rbp_kernel_launches()
        ms_total, ms_sample, ms_fold = s.step_timed(args.steps, flush_l2=True)
        gpu_launches = l.rbp_kernel_launches() - l0
        clk = clocks.stop()
        updates = s.counters()["updates"] - u0
    else:
        import torch
        from robopoker_b200.distributed import ShardedSolver

        stream = torch.cuda.current_stream()
        s.set_stream(stream.cuda_stream)
        sh = ShardedSolver(s, dist, device=local)
        flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
        sh.step(warm)
        u0 = s.counters()["updates"]
        dist.barrier(); torch.cuda.synchronize()
        clocks = Clocks(local)
        l0 = l.rbp_kernel_launches()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            flush.zero_()          # L2 flush, untimed
            e0.record(stream)
            sh.step(1)             # sample -> all-gather (NCCL) -> fold, all on this stream
            e1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        ms_total = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        ms_sample = ms_fold = float("nan")
        gpu_launches = l.rbp_kernel_launches() - l0
        clk = clocks.stop()
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        t = torch.tensor([s.counters()["updates"] - u0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        updates = float(t.item())
    value = updates / (ms_total * 1e-3)

    # end-to-end through the host-buffer API: import profile (H2D) → step → export profile (D2H), wall clock
    rows = s.profile_rows().copy()
    buf = np.zeros(len(rows) + 16, dtype=rows.dtype)
    e_steps = max(10, min(args.steps, 200))
    epochs = s.epochs
    stepper = (lambda: s.step(1)) if world == 1 else (lambda: sh.step(1))
    for _ in range(3):
        s.import_rows(rows, epochs); stepper(); rows = s.profile_rows(buf).copy(); epochs += 1
    ue0 = s.counters()["updates"]
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        s.import_rows(rows, epochs)   # H2D: the host mirror of the profile (what `MutProf::mut_*` edits)
        stepper()
        rows = s.profile_rows(buf)    # D2H: the refreshed host mirror
        epochs += 1
    e_dt = time.perf_counter() - t0
    e_updates = s.counters()["updates"] - ue0
    if dist:
        import torch
        t = torch.tensor([e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_dt = float(t.item())
        t = torch.tensor([e_updates], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        e_updates = float(t.item())
    e2e_val = e_updates / e_dt
    row_bytes = 24 * len(rows)

    # roofline of the dominant kernel (the ordered fold): algorithmic bytes = 32 B per infoset-action update
    # (16 B Encounter read + 16 B write) over the fold kernel's own CUDA-event time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_launch_updates = updates / world / args.steps
    if args.fold == "ordered":
        dom, dom_s = "mccfr_fold_kernel", ms_fold * 1e-3 / args.steps
        note = "240-row table is L2-resident; the ordered fold is bound by the reference's serial-per-row schedule, not by HBM"
    else:
        dom, dom_s = "mccfr_sample_kernel", (ms_sample if world == 1 else ms_total) * 1e-3 / args.steps
        note = ("sampling kernel: latency/divergence-bound tree walks over L1/L2-resident tables (SURVEY 8d: reported, not claimed "
                "against the HBM roofline); algorithmic bytes = 32 B per infoset-action update")
    achieved = 32.0 * per_launch_updates / dom_s / 1e9
    traffic = None
    try:  # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/)
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[dom]
        if args.batch == 262144 or dom == "mccfr_fold_kernel":
            traffic = t["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "measured" if peaks else "fallback", "traffic": traffic,
                "kernel_ms": {"mccfr_sample_kernel": ms_sample / args.steps, "fold_kernels": ms_fold / args.steps}, "note": note}

    if rank == 0:
        threads = os.cpu_count() or 1
        cpu_epochs = max(4, int(2.0e6 // args.batch))  # ~2M trees of CPU work: 10-30 s of core time
        cpu_val, cpu_dt = cpu_baseline_run(args, cpu_epochs, threads) if world == 1 else (None, None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload(args, world), "clocks": clk,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": row_bytes, "d2h_bytes_per_step": row_bytes, "steps": e_steps},
                "gpu_launches": int(gpu_launches), "roofline": roofline,
                "exploitability": s.exploitability(), "epochs": s.epochs}
        if cpu_val is not None:
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{cpu_epochs} epochs x {args.batch} trees in {cpu_dt:.1f}s, C++ restatement of the reference rayon path"}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
