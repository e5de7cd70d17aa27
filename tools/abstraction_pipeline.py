#!/usr/bin/env python
"""`trainer --cluster` on the GPU (crates/forge/src/pretraining.rs:26-63, minus Postgres): river → turn → flop.

    Rive: Lookup::grow            → rbp_isoset_river_buckets   (123,156,254 isomorphisms x 990 showdowns)
    Turn: Lookup::projections     → rbp_isoset_project         (13,960,050 histograms over 101 equity buckets)
          Layer::cluster          → rbp_kmeans_* (W1)          k-means++ / Elkan / lookup / metric
    Flop: Lookup::projections     → rbp_isoset_project         (1,286,792 histograms over the turn clusters)
          Layer::cluster          → rbp_kmeans_* (Sinkhorn)    on a subsample by default (log-domain parity kernel)

Prints one JSON line per stage with wall-clock seconds.  Sizes are the reference's; --turn-k / --iters / --flop-n scale the work.
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
    p.add_argument("--turn-k", type=int, default=256)
    p.add_argument("--flop-k", type=int, default=32)
    p.add_argument("--iters", type=int, default=32)
    p.add_argument("--flop-n", type=int, default=20000, help="flop points clustered with the Sinkhorn layer (0 = skip)")
    p.add_argument("--flop-iters", type=int, default=4)
    p.add_argument("--blueprint-epochs", type=int, default=0,
                   help="then train the NLHE blueprint on the clustered abstraction for this many epochs (`trainer --blueprint`)")
    p.add_argument("--blueprint-batch", type=int, default=16384)
    args = p.parse_args()
    import numpy as np

    import robopoker_b200 as rbp

    def stage(name, t0, **kw):
        print(json.dumps({"stage": name, "seconds": round(time.perf_counter() - t0, 3), **kw}), flush=True)

    t0 = time.perf_counter()
    river = rbp.deuce.IsoSet("rive")
    stage("river_isomorphisms", t0, n=len(river))
    t0 = time.perf_counter()
    river.river_buckets()
    stage("river_equity_lookup", t0, showdowns=len(river) * 990)
    t0 = time.perf_counter()
    turn = rbp.deuce.IsoSet("turn")
    n_turn = len(turn)
    hist = np.zeros((n_turn, 101), np.uint8)
    misses = 0
    step = 2_000_000
    for off in range(0, n_turn, step):
        h, m = turn.project(river, 101, off, min(step, n_turn - off))
        hist[off:off + len(h)] = h
        misses += m
    stage("turn_projections", t0, n=n_turn, misses=misses)
    if not args.blueprint_epochs:
        river.close()
    t0 = time.perf_counter()
    layer = rbp.lloyd.Layer(hist, args.turn_k)
    layer.init_centroids(0)
    stage("turn_kmeans_pp", t0, k=args.turn_k)
    t0 = time.perf_counter()
    layer.init_bounds()
    stage("turn_init_bounds", t0)
    t0 = time.perf_counter()
    steps = [layer.step() for _ in range(args.iters)]
    stage("turn_elkan_iterations", t0, iters=args.iters, last_reassigned=steps[-1].reassignment, last_drift_max=float(steps[-1].drift.max()),
          min_cluster=int(steps[-1].sizes.min()))
    t0 = time.perf_counter()
    lookup = layer.lookup()
    turn_metric = layer.metric()
    stage("turn_lookup_metric", t0)
    layer.close()
    del hist
    turn.set_abstractions(lookup.astype(np.uint8))
    t0 = time.perf_counter()
    flop = rbp.deuce.IsoSet("flop")
    fh, m = flop.project(turn, args.turn_k)
    stage("flop_projections", t0, n=len(flop), misses=m, mean_support=float((fh > 0).sum(axis=1).mean()))
    if args.flop_n:
        t0 = time.perf_counter()
        sub = fh[:: max(1, len(fh) // args.flop_n)][: args.flop_n]
        fl = rbp.lloyd.Layer(sub, args.flop_k, metric=turn_metric)
        fl.init_centroids(0)
        fl.init_bounds()
        st = [fl.step() for _ in range(args.flop_iters)]
        stage("flop_sinkhorn_kmeans_subsample", t0, n=len(sub), k=args.flop_k, iters=args.flop_iters, last_reassigned=st[-1].reassignment)
        if args.blueprint_epochs:
            t0 = time.perf_counter()
            if len(sub) < len(fh):  # `Layer::lookup` for every flop isomorphism against the learned centroids
                centroids, _ = fl.future()
                fl.close()
                fl = rbp.lloyd.Layer(fh, args.flop_k, metric=turn_metric)
                fl.set_centroids(centroids)
            flop.set_abstractions(fl.lookup().astype(np.uint8))
            stage("flop_lookup_all", t0, n=len(fh), solves=len(fh) * args.flop_k)
        fl.close()
    if args.blueprint_epochs:
        # `trainer --blueprint`: NlheEncoder's isomorphism → abstraction table is the clustering output, street by street
        from robopoker_b200.nlhe import Nlhe

        t0 = time.perf_counter()
        solver = Nlhe(batch=args.blueprint_batch, seed=0, table_slots=1 << 24)
        pref = rbp.deuce.IsoSet("pref")
        pref.set_abstractions(np.arange(len(pref), dtype=np.uint8))  # preflop: one bucket per isomorphism (169)
        for isos in (pref, flop, turn, river):
            solver.set_lookup(isos)
            isos.close()
        stage("blueprint_install_lookups", t0, streets=4)
        t0 = time.perf_counter()
        solver.step(args.blueprint_epochs)
        dt = time.perf_counter() - t0
        c = solver.counters()
        stage("blueprint_mccfr", t0, epochs=args.blueprint_epochs, trees=args.blueprint_epochs * args.blueprint_batch,
              updates=c["updates"], updates_per_s=c["updates"] / dt, infosets=c["rows"])


if __name__ == "__main__":
    main()
