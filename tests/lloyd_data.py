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


def synthetic_metric(bins, seed=0):
    """Random symmetric ground metric in (0, 1], normalised by its max, triangular Pair::merge order."""
    rng = np.random.default_rng(seed)
    tri = rng.uniform(0.05, 1.0, size=bins * (bins - 1) // 2).astype(np.float32)
    return tri / tri.max()


def flop_histograms(n, bins, seed=0, draws=47, spread=3.0):
    """Synthetic flop-layer points: 47 draws over `bins` next-street clusters, concentrated around a random centre."""
    rng = np.random.default_rng(seed)
    centers = rng.uniform(0, bins - 1, size=n)
    vals = np.clip(np.rint(rng.normal(centers[:, None], spread, size=(n, draws))), 0, bins - 1).astype(np.int64)
    pts = np.zeros((n, bins), dtype=np.uint8)
    np.add.at(pts, (np.repeat(np.arange(n), draws), vals.ravel()), 1)
    return pts


def flop_mixture_histograms(n, bins=256, comps=200, alpha=0.3, seed=0, draws=47):
    """`draws` samples per point from a `comps`-component Dirichlet(alpha) mixture over `bins`
    next-street clusters.  alpha=0.3 is the survey's value (mean support ~34); alpha=0.02 matches the mean support
    (~11) measured on real flop projections (profiles/r1g_pipeline_blueprint.json)."""
    rng = np.random.default_rng(seed)
    comp = rng.dirichlet(np.full(bins, alpha), size=comps)
    z = rng.integers(0, comps, n)
    pts = np.zeros((n, bins), dtype=np.uint8)
    for c in range(comps):
        rows = np.flatnonzero(z == c)
        if len(rows):
            pts[rows] = rng.multinomial(draws, comp[c], size=len(rows)).astype(np.uint8)
    return pts
