#!/usr/bin/env python
"""bench.py — headline benchmark of the MCCFR hot path on the north-star workload.

Default workload = BASELINE.json configs[3]: heads-up NLHE blueprint MCCFR, `Flagship` = (LinearRegret, LinearWeight,
PluribusSampling) (crates/nlhe/src/lib.rs:86-90), the reference's ORDERED fold (`Solver::step`,
crates/mccfr/src/solver/solver.rs:96-105: one schedule application per Decisions, in tree order), synthetic hash abstraction
(SURVEY §8d config 4), deals from the device-side Philox contract.

  step   = `--epochs-per-step` epochs (one `rbp_nlhe_step_timed` call) of `--batch` trees per GPU each
  value  = infoset-action regret updates per second, whole job, CUDA events on the library stream, table resident in HBM
  e2e    = the same metric through the host-facing call (`rbp_nlhe_step` + the counter read-back per step), wall clock
  N > 1  = trees sharded over ranks, infosets owned by hash mod N, exchange inside librbp_b200 (rbp_nlhe_attach_comm)

Other workloads: `--workload leduc` (configs[1]), `--workload lloyd_turn` / `lloyd_flop` (configs[4] / configs[2]).
`--impl reference` times the CPU restatement of the reference's rayon path (oracle/, built -O3 -march=native on this host).
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
LEDUC = ("leduc", "FlooredRegret", "LinearWeight", "ExternalSampling")
NLHE = ("LinearRegret", "LinearWeight", "PluribusSampling")  # `Flagship` (crates/nlhe/src/lib.rs:86-90)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="nlhe", choices=["nlhe", "leduc", "lloyd_turn", "lloyd_flop"],
                   help="nlhe = BASELINE.json configs[3] (default, the north-star workload); leduc = configs[1]; lloyd_turn = configs[4]; lloyd_flop = configs[2]")
    p.add_argument("--batch", type=int, default=None, help="trees per epoch per GPU (default: 16384 nlhe and leduc)")
    p.add_argument("--epochs-per-step", type=int, default=None, help="epochs in one bench step (default: 32 nlhe, 64 leduc)")
    p.add_argument("--table-slots", type=int, default=1 << 22, help="nlhe: infoset table capacity (power of two); the synthetic abstraction of config 4 yields ~1.5e5 infosets")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--k", type=int, default=None, help="lloyd_*: clusters (default 256 turn, 200 flop)")
    p.add_argument("--points", type=int, default=None, help="lloyd_*: points (default: the street's isomorphism count)")
    p.add_argument("--skip-cpu-baseline", action="store_true", help="tuning runs only: leave out the oracle's bounded CPU sample")
    p.add_argument("--fold", default="ordered", choices=["ordered", "batched"],
                   help="leduc only.  ordered = reference Solver::step semantics; batched = blocked delta sums, a DIFFERENT algorithm the "
                        "reference does not implement (no reference arm, no vs_reference)")
    a = p.parse_args()
    if a.batch is None:
        a.batch = 16384
    if a.epochs_per_step is None:
        a.epochs_per_step = 32 if a.workload == "nlhe" else 64
    return a


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


def env_ranks():
    return tuple(int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this round (profiles/), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")))[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def oracle_native():
    """The CPU arm's build of the oracle: -O3 -march=native on THIS host, loaded instead of the portable one."""
    from robopoker_b200 import build as b

    os.environ["RBP_ORACLE_LIB"] = b.build_oracle(native=True)
    from oracle import binding as oracle

    return oracle


def all_max(dist, x):
    import torch

    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(dist, x):
    import torch

    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    return float(t.item())


# ───────────────────────────────────────── NLHE blueprint (configs[3]) ─────────────────────────────────────────

def nlhe_config(args, n):
    return {"workload": f"configs[3] heads-up NLHE blueprint MCCFR (Flagship: {','.join(NLHE)}), ordered fold (reference Solver::step), "
                        f"{args.batch} trees/epoch/GPU, step = {args.epochs_per_step} epochs, synthetic hash abstraction 169/256/256/101 (SURVEY 8d config 4)",
            "game": "nlhe-hu", "fold": "ordered", "trees_per_epoch_per_gpu": args.batch, "epochs_per_step": args.epochs_per_step,
            "global_batch": args.batch * n,
            "parallelism": f"trees x{n}, infosets owned by hash mod {n}, exchange inside librbp_b200 (peer-memory stores + NCCL barriers)" if n > 1 else "trees x1",
            "table_slots": args.table_slots,
            "l2": "per-epoch working set (~0.4 GB of node arrays at 16384 trees) exceeds the 126 MB L2; plus one 192 MiB flush write before every step"}


def nlhe_cpu_sample(oracle, args, threads, budget_s, max_epochs):
    """Bounded sample of the same workload on the host cores: whole epochs of `batch` trees until the budget is spent."""
    o = oracle.OracleNlhe(seed=args.seed, batch=args.batch, threads=threads, regret=NLHE[0], weight=NLHE[1], sampling=NLHE[2])
    o.step(1)
    u0 = o.counters()["updates"]
    t0 = time.perf_counter()
    epochs = 0
    while epochs < max_epochs and (epochs == 0 or time.perf_counter() - t0 < budget_s):
        o.step(1)
        epochs += 1
    dt = time.perf_counter() - t0
    return (o.counters()["updates"] - u0) / dt, dt, epochs


def nlhe_reference(args):
    """`--impl reference`: the CPU restatement of the reference's rayon path on all host cores; rank 0 only.  Each step is a
    bounded sample of the GPU arm's step: ONE epoch of `batch` trees (the GPU arm's step is `epochs_per_step` of them)."""
    rank, _, _ = env_ranks()
    if rank != 0:
        return
    oracle = oracle_native()
    threads = os.cpu_count() or 1
    o = oracle.OracleNlhe(seed=args.seed, batch=args.batch, threads=threads, regret=NLHE[0], weight=NLHE[1], sampling=NLHE[2])
    o.step(max(1, min(args.warmup, 2)))
    u0 = o.counters()["updates"]
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    val = (o.counters()["updates"] - u0) / dt
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": nlhe_config(args, 1),
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                       "sample": f"{args.steps} steps x 1 epoch x {args.batch} trees (a 1/{args.epochs_per_step} sample of the GPU arm's step), C++ restatement of "
                                                 "the reference rayon path (Rust toolchain absent), -O3 -march=native, persistent pool"},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def main_nlhe(args):
    import torch

    import robopoker_b200 as rbp
    from robopoker_b200.nlhe import Nlhe

    rank, world, local = env_ranks()
    torch.cuda.set_device(local)
    dist = comm = None
    if world > 1:
        import torch.distributed as dist_mod
        from robopoker_b200.comm import Comm

        dist_mod.init_process_group("nccl")
        dist = dist_mod
        comm = Comm.from_torch(dist, device=local)
    l = rbp.load_library()
    if l.rbp_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — librbp_b200 has no CPU fallback")
    s = Nlhe(regret=NLHE[0], weight=NLHE[1], sampling=NLHE[2], batch=args.batch, seed=args.seed, table_slots=args.table_slots, device=local)
    stream = torch.cuda.Stream(device=local)
    s.set_stream(stream.cuda_stream)  # so that the bracketing events sit on the stream the kernels are launched on
    if comm:
        s.attach_comm(comm)
    E, K, warm = args.epochs_per_step, args.steps, max(args.warmup, 3)
    for _ in range(warm):
        s.step_timed(E, flush_l2=True)
    c0, t0 = s.counters(), s.traffic_counters()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = Clocks(local)
    l0 = l.rbp_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = [0.0] * 8
    e0.record(stream)
    for _ in range(K):
        ph = s.step_timed(E, flush_l2=True)  # E epochs; per-phase CUDA events inside
        phases = [a + b for a, b in zip(phases, ph)]
    e1.record(stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    launches = l.rbp_kernel_launches() - l0
    clk = clocks.stop()
    c1, t1 = s.counters(), s.traffic_counters()
    updates, nodes = c1["updates"] - c0["updates"], c1["nodes"] - c0["nodes"]
    if dist:
        ms_total = all_max(dist, ms_total)
        updates, nodes, launches = all_sum(dist, updates), all_sum(dist, nodes), all_sum(dist, launches)  # every rank folds what it owns
    value = updates / (ms_total * 1e-3)

    # end to end through the host-facing call: rbp_nlhe_step(E) + the 64-byte counter block per step, wall clock.  The path
    # has no per-step host input (deals come from the device-side Philox contract).
    e_steps = max(3, min(K, 10))
    u0 = s.counters()["updates"]
    if dist:
        dist.barrier()
    w0 = time.perf_counter()
    for _ in range(e_steps):
        s.step(E)
        s.counters()
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - w0
    e_updates = s.counters()["updates"] - u0
    if dist:
        e_dt, e_updates = all_max(dist, e_dt), all_sum(dist, e_updates)

    # roofline of the dominant phase (largest device time on rank 0), algorithmic bytes per SURVEY 8d:
    #   tree build: every decision node probes its infoset key (16 B) and reads A regrets (+ A weights at opponent nodes), and every
    #               node is written once as a 16-byte preorder node;  value: 16 B per node read + one 72-byte record per walker node;
    #   resolve+sort: 72-byte record + 16-byte key probe per record;  fold: 32 B per infoset-action update + the records it consumes
    n_epochs = K * E
    tw = {k: t1[k] - t0[k] for k in t0}
    records = tw["walker_nodes"]
    rk_nodes = c1["nodes"] - c0["nodes"]
    rk_updates = c1["updates"] - c0["updates"]
    alg = {"tree_build": 16.0 * rk_nodes + 16.0 * (tw["walker_nodes"] + tw["opponent_nodes"]) + 4.0 * tw["walker_choices"] + 8.0 * tw["opponent_choices"],
           "value": 16.0 * rk_nodes + 72.0 * records,
           "resolve_sort": 88.0 * records,
           "fold": 32.0 * rk_updates + 72.0 * records}
    kms = {"tree_build": phases[1] / n_epochs, "value": phases[2] / n_epochs, "resolve_sort": phases[3] / n_epochs, "fold": phases[4] / n_epochs}
    if world > 1:
        kms["records_to_owners"] = phases[5] / n_epochs
        kms["rows_to_peers"] = phases[6] / n_epochs
    dom = max(("tree_build", "value", "resolve_sort", "fold"), key=lambda k: kms[k])
    kernel_of = {"tree_build": "nlhe_level_kernels", "value": "nlhe_value_kernel", "resolve_sort": "nlhe_resolve_kernel", "fold": "nlhe_fold_kernel"}
    peak, peak_src = peaks()
    ach = alg[dom] / n_epochs / (kms[dom] * 1e-3) / 1e9
    whole = sum(alg.values()) / n_epochs / (phases[0] / n_epochs * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kernel_of[dom], "phase": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peak_src, "traffic": ncu_traffic(kernel_of[dom]),
                "algorithmic_bytes_per_epoch": {k: v / n_epochs for k, v in alg.items()}, "kernel_ms": kms,
                "epoch_ms_sum_of_phases": phases[0] / n_epochs, "whole_epoch_gbs": whole, "whole_epoch_frac": whole / peak,
                "note": "tree walks and per-infoset ordered chains are latency-bound: reported, not claimed, against the HBM roofline (SURVEY 8d); rank 0's phases"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": warm,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": nlhe_config(args, world), "clocks": clk,
                "e2e": {"value": e_updates / e_dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64, "steps": e_steps,
                        "note": "no per-step host input exists on this path (device-side deals); the host reads the counter block every step"},
                "gpu_launches": int(launches), "roofline": roofline, "epochs": c1["epochs"] + e_steps * E, "table_rows": s.counters()["rows"],
                "trees_per_s": n_epochs * args.batch * world / (ms_total * 1e-3)}
        if world == 1 and not args.skip_cpu_baseline:
            oracle = oracle_native()
            threads = os.cpu_count() or 1
            v, dt, ep = nlhe_cpu_sample(oracle, args, threads, budget_s=12.0, max_epochs=12)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{ep} epochs x {args.batch} trees in {dt:.1f}s, C++ restatement of the reference rayon path, -O3 -march=native, persistent pool"}
        print(json.dumps(line))
    s.close()
    if comm:
        comm.close()
    if dist:
        dist.destroy_process_group()


# ───────────────────────────────────────── Leduc (configs[1]) ─────────────────────────────────────────

def leduc_config(args, n):
    game, regret, weight, sampling = LEDUC
    return {"workload": f"configs[1] Leduc MCCFR ({regret},{weight},{sampling}), {args.batch} trees/epoch/GPU, step = {args.epochs_per_step} epochs, {args.fold} fold"
                        + (" — NOT a reference mode" if args.fold == "batched" else " (reference Solver::step)"),
            "fold": args.fold, "game": game, "trees_per_epoch_per_gpu": args.batch, "epochs_per_step": args.epochs_per_step,
            "global_batch": args.batch * n, "parallelism": f"trees x{n}", "table_rows": 240,
            "l2": "one 192 MiB flush write before every epoch, untimed (the 240-row table is cache-resident by nature)"}


def leduc_reference(args):
    rank, _, _ = env_ranks()
    if rank != 0:
        return
    if args.fold == "batched":
        print(json.dumps({"impl": "reference", "unavailable": "the batched fold is not an algorithm the reference implements (solver.rs:96-105 folds per Decisions)"}))
        return
    oracle = oracle_native()
    threads = os.cpu_count() or 1
    game, regret, weight, sampling = LEDUC
    o = oracle.OracleSolver(game, regret, weight, sampling, batch=args.batch, seed=args.seed, threads=threads)
    o.step(max(args.warmup, 1))
    u0 = o.counters()["updates"]
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    val = (o.counters()["updates"] - u0) / dt
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": leduc_config(args, 1),
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                       "sample": f"{args.steps} steps x 1 epoch x {args.batch} trees, C++ restatement of the reference rayon path (Rust toolchain absent), -O3 -march=native"},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def main_leduc(args):
    import numpy as np

    import robopoker_b200 as rbp

    rank, world, local = env_ranks()
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
    game, regret, weight, sampling = LEDUC
    fold = rbp.FOLD_BATCHED if args.fold == "batched" else rbp.FOLD_ORDERED
    if world > 1 and args.fold != "batched":
        raise SystemExit("bench.py --workload leduc: --gpus > 1 needs --fold batched (the reference's ordered fold over a 240-row table is serial per row "
                         "and has nothing to shard); the multi-GPU headline is --workload nlhe")
    s = rbp.Solver(game, regret, weight, sampling, batch=args.batch, seed=args.seed, device=local, fold=fold)
    E, K, warm = args.epochs_per_step, args.steps, max(args.warmup, 3)
    if world == 1:
        s.step_timed(warm * E, flush_l2=True)
        u0 = s.counters()["updates"]
        clocks = Clocks(local)
        l0 = l.rbp_kernel_launches()
        ms_total, ms_sample, ms_fold = s.step_timed(K * E, flush_l2=True)
        gpu_launches = l.rbp_kernel_launches() - l0
        clk = clocks.stop()
        updates = s.counters()["updates"] - u0
        stepper = lambda: s.step(E)  # noqa: E731
    else:
        import torch
        from robopoker_b200.comm import Comm

        stream = torch.cuda.current_stream()
        s.set_stream(stream.cuda_stream)
        comm = Comm.from_torch(dist, device=local)
        s.attach_comm(comm)        # the exchange runs inside the library
        sh = s
        sh.step(warm * E)
        u0 = s.counters()["updates"]
        dist.barrier(); torch.cuda.synchronize()
        clocks = Clocks(local)
        l0 = l.rbp_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sh.step(K * E)             # per epoch: sample -> all-gather of the partial sums (NCCL, inside the library) -> fold, all on this stream
        e1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        ms_total = all_max(dist, e0.elapsed_time(e1))
        ms_sample = ms_fold = None
        gpu_launches = l.rbp_kernel_launches() - l0
        clk = clocks.stop()
        updates = all_sum(dist, s.counters()["updates"] - u0)
        stepper = lambda: sh.step(E)  # noqa: E731
    value = updates / (ms_total * 1e-3)

    # end-to-end through the host-buffer API: import profile (H2D) → step → export profile (D2H), wall clock
    rows = s.profile_rows().copy()
    buf = np.zeros(len(rows) + 16, dtype=rows.dtype)
    e_steps = max(3, min(K, 10))
    epochs = s.epochs
    for _ in range(2):
        s.import_rows(rows, epochs); stepper(); rows = s.profile_rows(buf).copy(); epochs += E
    ue0 = s.counters()["updates"]
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        s.import_rows(rows, epochs)   # H2D: the host mirror of the profile (what `MutProf::mut_*` edits)
        stepper()
        rows = s.profile_rows(buf)    # D2H: the refreshed host mirror
        epochs += E
    e_dt = time.perf_counter() - t0
    e_updates = s.counters()["updates"] - ue0
    if dist:
        e_dt, e_updates = all_max(dist, e_dt), all_sum(dist, e_updates)
    row_bytes = 24 * len(rows)

    peak, peak_src = peaks()
    n_epochs = K * E
    per_launch_updates = updates / world / n_epochs
    if world == 1:
        dom, dom_ms = max((("mccfr_sample_kernel", ms_sample), ("mccfr_fold_kernel", ms_fold)), key=lambda kv: kv[1])
        dom_s = dom_ms * 1e-3 / n_epochs
        kms = {"mccfr_sample_kernel": ms_sample / n_epochs, "fold_kernels": ms_fold / n_epochs}
    else:
        dom, dom_s, kms = "epoch (sample + all-gather + fold)", ms_total * 1e-3 / n_epochs, None
    achieved = 32.0 * per_launch_updates / dom_s / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": None, "kernel_ms": kms,
                "note": "algorithmic bytes = 32 B per infoset-action update (SURVEY 8d); the 240-row table never leaves L1/L2, so this path is latency-bound "
                        "by construction: reported, not claimed, against the HBM roofline"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": warm,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": leduc_config(args, world), "clocks": clk,
                "e2e": {"value": e_updates / e_dt, "unit": UNIT, "h2d_bytes_per_step": row_bytes, "d2h_bytes_per_step": row_bytes, "steps": e_steps},
                "gpu_launches": int(gpu_launches), "roofline": roofline, "exploitability": s.exploitability(), "epochs": s.epochs}
        if world == 1 and not args.skip_cpu_baseline:
            oracle = oracle_native()
            threads = os.cpu_count() or 1
            o = oracle.OracleSolver(game, regret, weight, sampling, batch=args.batch, seed=args.seed, threads=threads)
            o.set_fold(1 if args.fold == "batched" else 0)
            o.step(1)
            cu0 = o.counters()["updates"]
            ct0 = time.perf_counter()
            ep = 0
            while ep < 4096 and (ep == 0 or time.perf_counter() - ct0 < 10.0):
                o.step(8)
                ep += 8
            cdt = time.perf_counter() - ct0
            line["cpu_baseline"] = {"value": (o.counters()["updates"] - cu0) / cdt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{ep} epochs x {args.batch} trees in {cdt:.1f}s, C++ restatement of the reference rayon path, -O3 -march=native, persistent pool"}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "nlhe":
        return nlhe_reference(args) if args.impl == "reference" else main_nlhe(args)
    if args.workload == "leduc":
        return leduc_reference(args) if args.impl == "reference" else main_leduc(args)
    from tools import bench_lloyd  # the k-means workloads live beside their data generators

    return bench_lloyd.main(args, env_ranks(), Clocks, peaks, oracle_native)


if __name__ == "__main__":
    main()
