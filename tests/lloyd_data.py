"""Synthetic turn-layer points: histograms of 46 river-equity draws over 101 buckets,
Beta(2,2)-centred, seeded."""
import numpy as np


def turn_histograms(n, seed=0, draws=46, spread=12.0):
    rng = np.random.default_rng(seed)
    centers = rng.beta(2.0, 2.0, size=n) * 100.0
    vals = np.clip(np.rint(rng.normal(centers[:, None], spread, size=(n, draws))), 0, 100).astype(np.int64)
    pts = np.zeros((n, 101), dtype=np.uint8)
    rows = np.repeat(np.arange(n), draws)
    np.add.at(pts, (rows, vals.ravel()), 1)
    return pts
