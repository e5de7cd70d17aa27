#!/usr/bin/env python
"""Bring-up of the tcgen05 / TMA screen kernel, one pipeline stage at a time (RBP_SCREEN_DEBUG=1..4): each stage runs in its own
process under a time-out so that a mistake shows up as a stage that does not return, not as a hung session."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, os, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
from lloyd_data import flop_mixture_histograms, synthetic_metric
import robopoker_b200 as rbp
pts = flop_mixture_histograms(300, 256, comps=200, alpha=0.02, seed=0)
tri = synthetic_metric(256, 0)
g = rbp.lloyd.Layer(pts, 200, metric=tri)
g.init_centroids(0)
try:
    approx, stats = g.screen_probe(64)
    print("stage", os.environ.get("RBP_SCREEN_DEBUG"), "returned; stats", stats, "first values", approx.ravel()[:8].tolist(), flush=True)
    lvl = int(os.environ.get("RBP_SCREEN_DEBUG", "0"))
    if lvl in (2, 3):
        print("  column sums of nu (1 expected):", float(approx.ravel()[:200].min()), float(approx.ravel()[:200].max()))
    if lvl == 4:
        # Q[j, y] = (1/|supp x0|) * sum_{x in supp x0} G[x, y] with G = bf16(exp(-C/T))
        idx = np.flatnonzero(pts[0]); C = np.zeros((256, 256), np.float32)
        hi, lo = np.tril_indices(256, -1); C[hi, lo] = tri; C[lo, hi] = tri
        G = np.exp(-(C / np.float32(0.025)).astype(np.float64))
        want = G[idx][:, :32].sum(axis=0) / len(idx)
        got = approx.ravel()[:200 * 32].reshape(200, 32)
        print("  expected Q[0, :4]", want[:4].tolist(), "got rows 0 / 1 / 150", got[0, :4].tolist(), got[1, :4].tolist(), got[150, :4].tolist())
        print("  max rel err over 200 x 32:", float(np.max(np.abs(got - want) / want)))
except Exception as e:
    print("stage", os.environ.get("RBP_SCREEN_DEBUG"), "error:", e, flush=True)
''' % (ROOT, ROOT)
for level in sys.argv[1:] or ["1", "2", "3", "4", "0"]:
    env = dict(os.environ, RBP_SCREEN_DEBUG=level)
    try:
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, timeout=40, capture_output=True, text=True)
        print(r.stdout.strip() or ("stage %s: no output, rc %d, stderr tail: %s" % (level, r.returncode, r.stderr[-300:])), flush=True)
    except subprocess.TimeoutExpired:
        print("stage", level, "did not return within 40 s", flush=True)
        break
