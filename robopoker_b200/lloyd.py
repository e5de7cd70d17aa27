"""Host-side mirror of the k-means abstraction layer: `Layer<K,N>` (crates/lloyd/src/layer.rs) driving
`Elkan` (crates/elkan/src/elkan.rs) through the C ABI.  Names follow the reference: `init_centroids`, `init_bounds`,
`step` (= `step_elkan` + `Prior::tally`), `lookup`, `metric`, `future`, `cluster`."""
import ctypes
from collections import namedtuple

import numpy as np

from . import _ffi

KMEANS_W1, KMEANS_SINKHORN = 0, 1
Step = namedtuple("Step", "index drift sizes reassignment")  # crates/elkan/src/step.rs


class Layer:
    """Turn-layer clustering: points are histograms over the 101 river-equity buckets, distance `Equity::variation`."""

    def __init__(self, counts, k, device=0, metric=None):
        """`metric=None`: turn layer (W1 over 101 river-equity buckets).  `metric=tri`: flop-style layer — points are
        histograms over next-street clusters compared by `Sinkhorn::divergence` under the triangular ground metric."""
        counts = np.ascontiguousarray(counts, dtype=np.uint8)
        assert counts.ndim == 2
        self.n, self.bins = counts.shape
        self.k = int(k)
        self._lib = _ffi.lib()
        self._h = ctypes.c_void_p()
        self.index = 0
        kind = KMEANS_W1 if metric is None else KMEANS_SINKHORN
        _ffi.check(self._lib.rbp_kmeans_create(kind, self.n, self.k, self.bins, counts.ctypes.data, device, ctypes.byref(self._h)),
                   "rbp_kmeans_create")
        if metric is not None:
            tri = np.ascontiguousarray(metric, dtype=np.float32)
            assert len(tri) == self.bins * (self.bins - 1) // 2
            _ffi.check(self._lib.rbp_kmeans_set_metric(self._h, tri.ctypes.data, self.bins), "rbp_kmeans_set_metric")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.rbp_kmeans_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_centroids(self, seed=0):
        """`Layer::init_centroids` (k-means++); returns the chosen point indices."""
        chosen = np.zeros(self.k, dtype=np.int32)
        _ffi.check(self._lib.rbp_kmeans_init_pp(self._h, seed, chosen.ctypes.data), "rbp_kmeans_init_pp")
        return chosen

    def set_centroids(self, counts):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        assert counts.shape == (self.k, self.bins)
        _ffi.check(self._lib.rbp_kmeans_set_centroids(self._h, counts.ctypes.data), "rbp_kmeans_set_centroids")

    def init_bounds(self):
        _ffi.check(self._lib.rbp_kmeans_init_bounds(self._h), "rbp_kmeans_init_bounds")

    def step(self):
        drift = np.zeros(self.k, np.float32)
        sizes = np.zeros(self.k, np.uint32)
        re = ctypes.c_uint32()
        _ffi.check(self._lib.rbp_kmeans_step(self._h, drift.ctypes.data, sizes.ctypes.data, ctypes.byref(re)), "rbp_kmeans_step")
        s = Step(self.index, drift, sizes, re.value)
        self.index += 1
        return s

    def step_local(self):
        _ffi.check(self._lib.rbp_kmeans_step_local(self._h), "rbp_kmeans_step_local")

    def step_finish(self):
        drift = np.zeros(self.k, np.float32)
        sizes = np.zeros(self.k, np.uint32)
        re = ctypes.c_uint32()
        _ffi.check(self._lib.rbp_kmeans_step_finish(self._h, drift.ctypes.data, sizes.ctypes.data, ctypes.byref(re)), "rbp_kmeans_step_finish")
        s = Step(self.index, drift, sizes, re.value)
        self.index += 1
        return s

    def exchange_buffers(self):
        """Device pointers a multi-GPU host all-reduces between step_local and step_finish: (acc_ptr, acc_bytes, sizes_ptr, reassigned_ptr, stream)."""
        p, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        _ffi.check(self._lib.rbp_kmeans_accumulator(self._h, ctypes.byref(p), ctypes.byref(nbytes)), "rbp_kmeans_accumulator")
        s, r = ctypes.c_void_p(), ctypes.c_void_p()
        _ffi.check(self._lib.rbp_kmeans_counters(self._h, ctypes.byref(s), ctypes.byref(r)), "rbp_kmeans_counters")
        return p.value, nbytes.value, s.value, r.value, self._lib.rbp_kmeans_stream(self._h)

    def lookup(self, with_distance=False):
        """`Layer::lookup`: fresh naive argmin; returns assignments (u32[n]) [and distances]."""
        out = np.zeros(self.n, np.uint32)
        dist = np.zeros(self.n, np.float32) if with_distance else None
        _ffi.check(self._lib.rbp_kmeans_assign(self._h, out.ctypes.data, dist.ctypes.data if with_distance else None), "rbp_kmeans_assign")
        return (out, dist) if with_distance else out

    def future(self):
        """`Layer::future`: centroid histograms (counts[k][bins], weights[k])."""
        counts = np.zeros((self.k, self.bins), np.uint64)
        weights = np.zeros(self.k, np.uint64)
        _ffi.check(self._lib.rbp_kmeans_centroids(self._h, counts.ctypes.data, weights.ctypes.data), "rbp_kmeans_centroids")
        return counts, weights

    def metric(self):
        tri = np.zeros(self.k * (self.k - 1) // 2, np.float32)
        _ffi.check(self._lib.rbp_kmeans_metric(self._h, tri.ctypes.data), "rbp_kmeans_metric")
        return tri

    def bounds(self, with_lower=False):
        a, u, st = np.zeros(self.n, np.uint32), np.zeros(self.n, np.float32), np.zeros(self.n, np.uint8)
        lo = np.zeros((self.n, self.k), np.float32) if with_lower else None
        _ffi.check(self._lib.rbp_kmeans_bounds(self._h, a.ctypes.data, u.ctypes.data, lo.ctypes.data if with_lower else None, st.ctypes.data),
                   "rbp_kmeans_bounds")
        return a, u, lo, st

    def timed(self, what=0, iters=1):
        ms = ctypes.c_float()
        _ffi.check(self._lib.rbp_kmeans_timed(self._h, what, iters, ctypes.byref(ms)), "rbp_kmeans_timed")
        return ms.value

    def sinkhorn_stats(self, reset=False):
        """Work counters of a Sinkhorn layer: (OT solves, Gauss-Seidel sweeps, exp terms) since creation / the last reset."""
        out = np.zeros(3, np.uint64)
        if not hasattr(self._lib, "rbp_kmeans_sinkhorn_stats"):  # an older build loaded through RBP_LIB_PATH
            return 0, 0, 0
        _ffi.check(self._lib.rbp_kmeans_sinkhorn_stats(self._h, out.ctypes.data, int(reset)), "rbp_kmeans_sinkhorn_stats")
        return int(out[0]), int(out[1]), int(out[2])

    def attach_comm(self, comm):
        """Point-sharded clustering inside the library: `step()` then includes the one integer all-reduce (`robopoker_b200.comm.Comm`)."""
        _ffi.check(self._lib.rbp_kmeans_attach_comm(self._h, comm._h), "rbp_kmeans_attach_comm")
        self._comm = comm
        return self

    def screen(self, margin):
        """Tensor-core screen of the naive sweeps (`init_bounds`, `lookup`) of a Sinkhorn layer: exact solves only for the centroids
        within `margin` of a point's smallest approximate divergence (csrc/sk_screen.cuh).  margin < 0 switches it off."""
        _ffi.check(self._lib.rbp_kmeans_screen(self._h, float(margin)), "rbp_kmeans_screen")
        return self

    def screen_probe(self, m=None):
        """Approximate divergences [m][k] of the first m points against every centroid, and (problems, iterations) counters."""
        m = self.n if m is None else int(m)
        out = np.zeros((m, self.k), np.float32)
        stats = np.zeros(2, np.uint64)
        _ffi.check(self._lib.rbp_kmeans_screen_probe(self._h, m, out.ctypes.data, stats.ctypes.data), "rbp_kmeans_screen_probe")
        return out, (int(stats[0]), int(stats[1]))

    def cluster(self, iterations=32, seed=0):
        """`Layer::cluster` minus persistence: init → bounds → `iterations` Elkan steps → (lookup, metric, future)."""
        self.init_centroids(seed)
        self.init_bounds()
        steps = [self.step() for _ in range(iterations)]
        return self.lookup(), self.metric(), self.future(), steps


def sinkhorn_divergence(a_counts, b_counts, ia, ib, tri, temperature=0.025, iterations=128, tolerance=0.0005):
    """Batched `Metric::emd` → `Sinkhorn::divergence` (crates/lloyd/src/metric.rs:109-115): out[t] = divergence(A[ia[t]], B[ib[t]])."""
    a = np.ascontiguousarray(a_counts, dtype=np.uint32)
    b = np.ascontiguousarray(b_counts, dtype=np.uint32)
    ia = np.ascontiguousarray(ia, dtype=np.int32)
    ib = np.ascontiguousarray(ib, dtype=np.int32)
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    out = np.zeros(len(ia), np.float32)
    _ffi.check(_ffi.lib().rbp_sinkhorn_batch(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], a.shape[1], ia.ctypes.data, ib.ctypes.data,
                                             len(ia), tri.ctypes.data, temperature, iterations, tolerance, out.ctypes.data), "rbp_sinkhorn_batch")
    return out
