// lloyd_w1.cu — the TURN abstraction layer on sm_100a: Elkan/Lloyd k-means over histograms of river-equity
// buckets under the 1-D Wasserstein distance `Equity::variation`, and its C ABI (include/rbp.h).
//
// Reference seams (paths relative to krukah/robopoker):
//   crates/lloyd/src/layer.rs:136-181,195-248   Layer::{distance, init_centroids (k-means++), cluster}
//   crates/elkan/src/elkan.rs:39-168            init_bounds, neighbor, pairwises, midpoints, refresh, rebound,
//                                               recompute, drift, step_elkan;  bounds.rs:57-91 Bounds
//   crates/lloyd/src/equity.rs:41-53            variation(x, y) = Σ_b |cdf_x(b) − cdf_y(b)| / 101
//   crates/lloyd/src/bins.rs:58-60,75-82        density = count/weight (f32), merge = integer add
//   crates/lloyd/src/layer.rs:44-101            lookup (fresh naive argmin), metric (symmetrised, normalised)
//
// Design (B200-first, not a port).  `variation` separates: each histogram's f32 CDF is built once (sequential
// adds, exactly as the reference accumulates it) and a distance is then 101 |a−b| terms summed in bin order.
// A thread owns one point and keeps its 101-entry CDF in REGISTERS; centroid CDFs sit in shared memory (K ≤ 512
// per tile, 104-float rows read as broadcast LDS.128) — so a distance is 26 LDS + 202 FADD and the kernels are
// FP32-pipe bound, while the Elkan bounds (N×K f32, K-major so a warp's accesses coalesce) stream through HBM
// exactly once per iteration: the previous iteration's drift update (`Bounds::update`) is applied lazily when the
// bound is next read, fusing the reference's second N×K pass into the first.
// Elkan's pruning tests are evaluated per point exactly as the reference does (same f32 comparisons, same order,
// including the mid-loop switch of the pairwise row when a point is reassigned), so assignments, centroids and
// drifts are bit-identical to the oracle; the distance itself is only computed when some lane of the warp needs it.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

#include "kmeans_common.cuh"

namespace rbp {

constexpr int kBins = 101;       // KMEANS_EQTY_CLUSTER_COUNT, crates/pokerkit/src/lib.rs:191
constexpr int kRow = 112;        // bytes per stored point row (101 counts + pad, 7 x 16 B)
constexpr int kCdfRow = 104;     // floats per centroid CDF row in shared memory (26 x LDS.128)
constexpr int kTileK = 512;      // centroids per shared-memory tile (512 x 104 x 4 B = 208 KB)
constexpr int kThreads = 128;        // step / k-means++ kernels
constexpr int kStepThreads = 256;
constexpr int kAssignThreads = 256;  // assign kernel: 2 blocks x 8 warps per SM at 128 registers

struct KmDev {
    const uint8_t* pts;   // [N][kRow]
    int64_t n;
    int k;
    float* cdf;           // [K][kCdfRow] centroid CDFs (current)
    float* pair;          // [K][kp], kp = K rounded up to 4
    int kp;
    float* mid;           // [K]
    float* drift;         // [K] drift of the PREVIOUS step, applied lazily
    float* lower;         // [K][N]
    float* upper;         // [N]
    uint32_t* assign;     // [N]
    uint8_t* stale;       // [N]
    unsigned long long* acc;   // [K][kBins + 1] new centroid counts + weight (integer merge)
    unsigned long long* ccount;  // [K][kBins + 1] current centroid counts + weight
    uint32_t* sizes;      // [K]
    uint32_t* reassigned; // [1]
    int pending;          // 1 if `drift` has not yet been folded into lower/upper
};

// ── per-thread point CDF in registers (bins.rs:58-60 density, equity.rs:45-47 running cdf) ──
__device__ __forceinline__ void load_point_cdf(const uint8_t* __restrict__ row, float (&X)[kBins]) {
    uint32_t w[kRow / 4];
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int q = 0; q < kRow / 16; ++q) {
        uint4 v = r4[q];
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
    uint32_t weight = 0;
#pragma unroll
    for (int b = 0; b < kBins; ++b) weight += (w[b >> 2] >> (8 * (b & 3))) & 0xFFu;
    const float wf = (float)weight;
    float c = 0.0f;
#pragma unroll
    for (int b = 0; b < kBins; ++b) {
        const uint32_t cnt = (w[b >> 2] >> (8 * (b & 3))) & 0xFFu;
        c += (float)cnt / wf;   // adding 0/w = +0 leaves c unchanged, as in the reference
        X[b] = c;
    }
}

// distance of the register CDF to up to 4 centroid CDF rows in shared memory (equity.rs:48-52: sequential sum, / 101)
template <int G>
__device__ __forceinline__ void dist_group(const float (&X)[kBins], const float* __restrict__ rows, float (&out)[G]) {
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.0f;
#pragma unroll
    for (int q = 0; q < kCdfRow / 4; ++q) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float4 c = *reinterpret_cast<const float4*>(rows + g * kCdfRow + 4 * q);
            if (4 * q + 0 < kBins) acc[g] += fabsf(X[4 * q + 0 < kBins ? 4 * q + 0 : 0] - c.x);
            if (4 * q + 1 < kBins) acc[g] += fabsf(X[4 * q + 1 < kBins ? 4 * q + 1 : 0] - c.y);
            if (4 * q + 2 < kBins) acc[g] += fabsf(X[4 * q + 2 < kBins ? 4 * q + 2 : 0] - c.z);
            if (4 * q + 3 < kBins) acc[g] += fabsf(X[4 * q + 3 < kBins ? 4 * q + 3 : 0] - c.w);
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) out[g] = acc[g] / (float)kBins;
}

__device__ __forceinline__ void stage_cdf_tile(float* s_cdf, const float* __restrict__ cdf, int j0, int kt) {
    const float4* src = reinterpret_cast<const float4*>(cdf + (size_t)j0 * kCdfRow);
    float4* dst = reinterpret_cast<float4*>(s_cdf);
    for (int t = threadIdx.x; t < kt * (kCdfRow / 4); t += blockDim.x) dst[t] = src[t];
}

// ── centroid CDFs from integer counts (K threads) ──
__global__ void centroid_cdf_kernel(const unsigned long long* __restrict__ counts, int k, float* __restrict__ cdf) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const unsigned long long* c = counts + (size_t)j * (kBins + 1);
    const float wf = (float)c[kBins];
    float run = 0.0f;
    for (int b = 0; b < kBins; ++b) {
        run += (float)c[b] / wf;
        cdf[(size_t)j * kCdfRow + b] = run;
    }
    for (int b = kBins; b < kCdfRow; ++b) cdf[(size_t)j * kCdfRow + b] = 0.0f;
}
__device__ __forceinline__ float cdf_distance(const float* __restrict__ a, const float* __restrict__ b) {
    float acc = 0.0f;
    for (int q = 0; q < kBins; ++q) acc += fabsf(a[q] - b[q]);
    return acc / (float)kBins;
}
// elkan.rs:80-105 pairwises + midpoints (one block per row i)
__global__ void pairwise_kernel(const float* __restrict__ cdf, int k, int kp, float* __restrict__ pair, float* __restrict__ mid) {
    const int i = blockIdx.x;
    __shared__ float s_min[256];
    float m = 3.402823466e+38f;  // f32::MAX
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const float d = i == j ? 0.0f : cdf_distance(cdf + (size_t)i * kCdfRow, cdf + (size_t)j * kCdfRow);
        pair[(size_t)i * kp + j] = d;
        if (j != i) { const float h = d * 0.5f; m = h < m ? h : m; }  // f32::min ignores a NaN operand
    }
    s_min[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { const float o = s_min[threadIdx.x + s]; if (o < s_min[threadIdx.x]) s_min[threadIdx.x] = o; }
        __syncthreads();
    }
    if (threadIdx.x == 0) mid[i] = s_min[0];
}
// elkan.rs:107-109 drift_j = distance(new_j, old_j)
__global__ void drift_kernel(const float* __restrict__ new_cdf, const float* __restrict__ old_cdf, int k, float* __restrict__ drift) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) drift[j] = cdf_distance(new_cdf + (size_t)j * kCdfRow, old_cdf + (size_t)j * kCdfRow);
}
// layer.rs:85-101 metric: (d(i,j) + d(j,i)) / 2 for i > j, triangular index pair.rs:36-39
__global__ void metric_kernel(const float* __restrict__ cdf, int k, float* __restrict__ tri) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = k * (k - 1) / 2;
    if (t >= total) return;
    int hi = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
    while (hi * (hi - 1) / 2 > t) --hi;
    while ((hi + 1) * hi / 2 <= t) ++hi;
    const int lo = t - hi * (hi - 1) / 2;
    const float a = cdf_distance(cdf + (size_t)hi * kCdfRow, cdf + (size_t)lo * kCdfRow);
    const float b = cdf_distance(cdf + (size_t)lo * kCdfRow, cdf + (size_t)hi * kCdfRow);
    tri[t] = (a + b) / 2.0f;
}
__global__ void metric_normalize_kernel(float* tri, int total) {  // metric.rs:127-141
    __shared__ float s_max[256];
    float m = 1.17549435e-38f;
    for (int t = threadIdx.x; t < total; t += blockDim.x) m = tri[t] > m ? tri[t] : m;
    s_max[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s && s_max[threadIdx.x + s] > s_max[threadIdx.x]) s_max[threadIdx.x] = s_max[threadIdx.x + s];
        __syncthreads();
    }
    const float mx = s_max[0];
    for (int t = threadIdx.x; t < total; t += blockDim.x) tri[t] = tri[t] / mx;
}

// ── naive argmin over all centroids: init_bounds (elkan.rs:39-47), lookup (layer.rs:44-60) ──
template <bool INIT_BOUNDS>
__global__ void __launch_bounds__(kAssignThreads, 2)
assign_kernel(KmDev km, uint32_t* __restrict__ out_assign, float* __restrict__ out_dist) {
    extern __shared__ __align__(16) float s_cdf[];
    const int64_t i = blockIdx.x * (int64_t)kAssignThreads + threadIdx.x;
    const bool live = i < km.n;
    float X[kBins];
    if (live) load_point_cdf(km.pts + (size_t)i * kRow, X);
    else {
#pragma unroll
        for (int b = 0; b < kBins; ++b) X[b] = 0.0f;
    }
    float best = 0.0f;
    int bestj = -1;
    for (int j0 = 0; j0 < km.k; j0 += kTileK) {
        const int kt = min(kTileK, km.k - j0);
        __syncthreads();
        stage_cdf_tile(s_cdf, km.cdf, j0, kt);
        __syncthreads();
        int j = 0;
        for (; j + 4 <= kt; j += 4) {
            float d[4];
            dist_group<4>(X, s_cdf + (size_t)j * kCdfRow, d);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                if (bestj < 0 || d[g] < best) { best = d[g]; bestj = j0 + j + g; }  // first minimum (Iterator::min_by)
        }
        for (; j < kt; ++j) {
            float d[1];
            dist_group<1>(X, s_cdf + (size_t)j * kCdfRow, d);
            if (bestj < 0 || d[0] < best) { best = d[0]; bestj = j0 + j; }
        }
    }
    if (!live) return;
    if (INIT_BOUNDS) {  // Bounds::from((j, upper)): lower = 0, stale = false (bounds.rs:108-117)
        km.assign[i] = (uint32_t)bestj;
        km.upper[i] = best;
        km.stale[i] = 0;
        for (int j = 0; j < km.k; ++j) km.lower[(size_t)j * km.n + i] = 0.0f;
    } else {
        out_assign[i] = (uint32_t)bestj;
        if (out_dist) out_dist[i] = best;
    }
}

// ── one Elkan step, point side (elkan.rs:153-164 up to recompute) ──
// TMA bulk copy (1-D: the rows of a centroid tile are contiguous) of `bytes` into shared memory, completion on an mbarrier
__device__ __forceinline__ void tile_bulk_load(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst), b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was last read through the generic proxy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void tile_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
}
// PIPE: the centroid CDF tiles arrive by TMA bulk copies into a two-stage shared-memory ring (tile t + 1 is in flight while tile t is
// worked on); otherwise every tile is staged with plain loads between two block barriers.
// kRing > 0: the bounds stream lower[j][i] is fetched by per-thread cp.async copies kRing groups (of 4 centroids) ahead into a private
// shared-memory ring.  With the one-group register lookahead a warp had 4 x 128 B in flight and an SM (12 warps at 180 registers) about
// 6 KB — by Little's law ~1.1 TB/s over the whole device at HBM latency, which is what the step measured (profiles/r2y, r2z).
// kFast (needs kRing): groups of 4 centroids that are complete, in blocks whose threads all own a point, take a path without the
// per-member tail / liveness tests, with running pointers instead of 64-bit index products and with the member-by-member walk only in
// warps where some lane examines a centroid — bookkeeping was 57 % of the step's instructions (profiles/r2ab_elkan_step_hot_lines.txt).
template <int THREADS, int MINB, int TILE, bool PIPE = false, int kFar = 0, int kRing = 0, bool kFast = false>
__global__ void __launch_bounds__(THREADS, MINB)
elkan_step_kernel(KmDev km) {
    static_assert(!kFast || kRing > 0, "the fast path reads the bounds from the ring");
    extern __shared__ __align__(128) float s_cdf[];
    __shared__ __align__(8) uint64_t s_bar[2];
    static_assert(TILE % 4 == 0, "a group of 4 centroids never straddles two tiles");
    float* s_drift = s_cdf + (PIPE ? (size_t)2 * TILE : (size_t)min(km.k, TILE)) * kCdfRow;  // [K] drift of the previous step
    constexpr int kSlots = kRing == 6 ? 8 : kRing + 2;     // the slot being refilled was read (at least) two groups ago
    float* s_ring = s_drift + ((km.k + 3) & ~3);            // [kSlots][4][THREADS], cell (slot, g, thread) is private to the thread
    for (int j = threadIdx.x; j < km.k; j += blockDim.x) s_drift[j] = km.pending ? km.drift[j] : 0.0f;
    if (PIPE && threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (PIPE && threadIdx.x == 0) tile_bulk_load(s_cdf, km.cdf, (uint32_t)min(TILE, km.k) * kCdfRow * 4u, &s_bar[0]);
    uint32_t bar_phase = 0;  // bit b = parity to wait for on barrier b
    const int64_t i = blockIdx.x * (int64_t)THREADS + threadIdx.x;
    const bool live = i < km.n;
    float X[kBins];
    if (live) load_point_cdf(km.pts + (size_t)i * kRow, X);
    else {
#pragma unroll
        for (int b = 0; b < kBins; ++b) X[b] = 0.0f;
    }
    const int n_groups = (km.k + 3) >> 2;
    auto ring_issue = [&](int G) {  // one commit group per group of centroids, also when there is nothing to fetch: the wait below counts groups
        if (kRing > 0) {
            if (G < n_groups && live) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int j = G * 4 + g;
                    if (j < km.k) {
                        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_ring + ((size_t)(G % kSlots) * 4 + g) * THREADS + threadIdx.x);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(km.lower + (size_t)j * km.n + i) : "memory");
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    if (kRing > 0)
        for (int G = 0; G < kRing; ++G) ring_issue(G);
    uint32_t c = 0, c_prior = 0;
    float u = 0.0f;
    bool stale = false;
    if (live) {
        c = c_prior = km.assign[i];
        u = km.upper[i];
        stale = km.stale[i] != 0;
        if (km.pending) { u += s_drift[c]; stale = true; }  // Bounds::update of the previous step (bounds.rs:65-74)
    }
    const bool act = live && u > km.mid[c];  // step_elkan filter: b.u() > midpoints[b.j()]
    // refresh (elkan.rs:113-117, bounds.rs:76-80): recompute the stale upper bound exactly; lower[c] is written below
    bool refreshed = false;
    float refreshed_d = 0.0f;
    {
        const bool need = act && stale;
        if (__any_sync(0xFFFFFFFFu, need)) {
            if (need) {
                float d[1];
                dist_group<1>(X, km.cdf + (size_t)c * kCdfRow, d);  // global/L2 read of one row
                u = d[0]; stale = false; refreshed = true; refreshed_d = d[0];
            }
        }
    }
    const uint32_t c_refresh = c;
    const bool block_full = (int64_t)(blockIdx.x + 1) * THREADS <= km.n;
    const int n_full_groups = km.k >> 2;
    const size_t n4 = 4 * (size_t)km.n;
    float* lower_ptr = km.lower + i;                      // &lower[j][i] of the group being worked on
    const float* prow = km.pair + (size_t)c * km.kp;      // the pairwise row of the current centroid
    for (int j0 = 0; j0 < km.k; j0 += TILE) {
        const int kt = min(TILE, km.k - j0);
        float* tile = s_cdf;
        if (PIPE) {
            const int t = j0 / TILE, nb = (t + 1) & 1;
            if (threadIdx.x == 0 && j0 + TILE < km.k)   // buffer nb was last read in the previous iteration, which ended with a block barrier
                tile_bulk_load(s_cdf + (size_t)nb * TILE * kCdfRow, km.cdf + (size_t)(j0 + TILE) * kCdfRow, (uint32_t)min(TILE, km.k - j0 - TILE) * kCdfRow * 4u, &s_bar[nb]);
            tile = s_cdf + (size_t)(t & 1) * TILE * kCdfRow;
            tile_wait(&s_bar[t & 1], bar_phase >> (t & 1) & 1u);
            bar_phase ^= 1u << (t & 1);
        } else {
            __syncthreads();
            stage_cdf_tile(s_cdf, km.cdf, j0, kt);
            __syncthreads();
        }
        // software pipeline: the bounds and the pairwise row of group g+1 are requested before group g is worked on,
        // so their HBM/L2 latency hides behind the 808 FADDs of the current group
        float ln[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        float4 prn = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        uint32_t c_pref = c;
        auto prefetch = [&](int jj) {
            const int g_n = min(4, kt - jj);
            if (kRing == 0) {
#pragma unroll
                for (int g = 0; g < 4; ++g) ln[g] = (g < g_n && live) ? km.lower[(size_t)(j0 + jj + g) * km.n + i] : 0.0f;
            }
            prn = *reinterpret_cast<const float4*>(km.pair + (size_t)c * km.kp + j0 + jj);
            c_pref = c;
        };
        prefetch(0);
        for (int jj = 0; jj < kt; jj += 4, lower_ptr += n4) {
            if (kFast && block_full && jj + 4 <= kt) {
                const int jg = j0 + jj, G = jg >> 2;
                if (G + kRing < n_full_groups) {  // fetch group G + kRing: 4 complete rows of the bounds stream
                    const float* src = lower_ptr + (size_t)kRing * n4;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_ring + (size_t)((G + kRing) % kSlots) * 4 * THREADS + threadIdx.x);
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(g * THREADS * 4)), "l"(src + (size_t)g * km.n) : "memory");
                    asm volatile("cp.async.commit_group;" ::: "memory");
                } else ring_issue(G + kRing);
                asm volatile("cp.async.wait_group %0;" ::"n"(kRing) : "memory");
                const float* cell = s_ring + (size_t)(G % kSlots) * 4 * THREADS + threadIdx.x;
                float l[4] = {cell[0], cell[THREADS], cell[2 * THREADS], cell[3 * THREADS]};
                float4 pr = prn;
                if (c_pref != c) pr = *reinterpret_cast<const float4*>(prow + jg);  // reassigned since the prefetch
                if (jj + 4 < kt) { prn = *reinterpret_cast<const float4*>(prow + jg + 4); c_pref = c; }
                const float4 dr = *reinterpret_cast<const float4*>(s_drift + jg);    // zeros on the first step: (l - 0).max(0) = l for l >= +0
                l[0] = l[0] - dr.x; l[0] = l[0] > 0.0f ? l[0] : 0.0f;
                l[1] = l[1] - dr.y; l[1] = l[1] > 0.0f ? l[1] : 0.0f;
                l[2] = l[2] - dr.z; l[2] = l[2] > 0.0f ? l[2] : 0.0f;
                l[3] = l[3] - dr.w; l[3] = l[3] > 0.0f ? l[3] : 0.0f;
                if (refreshed && (c_refresh >> 2) == (uint32_t)G) {  // Bounds::refresh sets lower[j]
#pragma unroll
                    for (int g = 0; g < 4; ++g) if ((c_refresh & 3u) == (uint32_t)g) l[g] = refreshed_d;
                }
                float half[4] = {0.5f * pr.x, 0.5f * pr.y, 0.5f * pr.z, 0.5f * pr.w};
                const uint32_t cg = c - (uint32_t)jg;  // the current centroid's place in the group (>= 4: not in it)
                const bool want = act && ((cg != 0u && u > l[0] && u > half[0]) || (cg != 1u && u > l[1] && u > half[1]) ||
                                          (cg != 2u && u > l[2] && u > half[2]) || (cg != 3u && u > l[3] && u > half[3]));
                if (__any_sync(0xFFFFFFFFu, want)) {
                    float d[4];
                    dist_group<4>(X, tile + (size_t)jj * kCdfRow, d);
                    if (want) {
                        uint32_t c_row = c;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint32_t j = (uint32_t)(jg + g);
                            if (c != c_row) {  // the pairwise row switches when the point is reassigned mid-group
                                const float4 p2 = *reinterpret_cast<const float4*>(prow + jg);
                                half[0] = 0.5f * p2.x; half[1] = 0.5f * p2.y; half[2] = 0.5f * p2.z; half[3] = 0.5f * p2.w;
                                c_row = c;
                            }
                            if (j != c && u > l[g] && u > half[g]) {  // bounds.rs:57-61 has_shifted
                                l[g] = d[g];                           // witness (bounds.rs:81-87)
                                if (d[g] < u) { c = j; u = d[g]; prow = km.pair + (size_t)c * km.kp; }
                            }
                        }
                    }
                }
                lower_ptr[0] = l[0]; lower_ptr[(size_t)km.n] = l[1]; lower_ptr[2 * (size_t)km.n] = l[2]; lower_ptr[3 * (size_t)km.n] = l[3];
                continue;
            }
            const int g_n = min(4, kt - jj);
            float l[4], half[4];
            bool want = false;
            uint32_t c_row = c;
            float4 pr = prn;
            if (c_pref != c) pr = *reinterpret_cast<const float4*>(km.pair + (size_t)c * km.kp + j0 + jj);  // reassigned since the prefetch
            half[0] = 0.5f * pr.x; half[1] = 0.5f * pr.y; half[2] = 0.5f * pr.z; half[3] = 0.5f * pr.w;
            if (kRing > 0) {
                const int G = (j0 + jj) >> 2;
                ring_issue(G + kRing);
                asm volatile("cp.async.wait_group %0;" ::"n"(kRing) : "memory");  // all but the kRing newest groups have landed: group G is in its slot
#pragma unroll
                for (int g = 0; g < 4; ++g) l[g] = (g < g_n && live) ? s_ring[((size_t)(G % kSlots) * 4 + g) * THREADS + threadIdx.x] : 0.0f;
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) l[g] = ln[g];
            }
            if (jj + 4 < kt) prefetch(jj + 4);
            if (kFar > 0 && live && j0 + jj + kFar + 4 <= km.k) {  // pull the bounds of a later group from HBM into L2 (they cross tiles: the stream is [K][N])
#pragma unroll
                for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(km.lower + (size_t)(j0 + jj + kFar + g) * km.n + i));
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g < g_n && live) {
                    const int j = j0 + jj + g;
                    float v = l[g];
                    if (km.pending) { v = v - s_drift[j]; v = v > 0.0f ? v : 0.0f; }  // (lower - movement).max(0.0)
                    if (refreshed && (uint32_t)j == c_refresh) v = refreshed_d;        // Bounds::refresh sets lower[j]
                    l[g] = v;
                    // could this centroid be examined under the current (c, u)?  (re-tested exactly below)
                    want |= act && (uint32_t)j != c && u > v && u > half[g];
                }
            }
            // A reassignment inside the group can enable a later member, but only after an earlier member was
            // examined — so if no member passes under the current state, none is examined at all; otherwise all
            // four distances are produced together.
            float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (__any_sync(0xFFFFFFFFu, want)) {
                if (!kFast && g_n == 4) dist_group<4>(X, tile + (size_t)jj * kCdfRow, d);
                else
                    for (int g = 0; g < g_n; ++g) { float t[1]; dist_group<1>(X, tile + (size_t)(jj + g) * kCdfRow, t); d[g] = t[0]; }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g < g_n && live) {
                    const int j = j0 + jj + g;
                    if (c != c_row) {  // the pairwise row switches when the point is reassigned mid-loop
                        const float4 p2 = *reinterpret_cast<const float4*>(km.pair + (size_t)c * km.kp + j0 + jj);
                        half[0] = 0.5f * p2.x; half[1] = 0.5f * p2.y; half[2] = 0.5f * p2.z; half[3] = 0.5f * p2.w;
                        c_row = c;
                    }
                    if (want && act && (uint32_t)j != c && u > l[g] && u > half[g]) {  // bounds.rs:57-61 has_shifted
                        l[g] = d[g];                                   // witness (bounds.rs:81-87)
                        if (d[g] < u) { c = (uint32_t)j; u = d[g]; prow = km.pair + (size_t)c * km.kp; }
                    }
                    km.lower[(size_t)j * km.n + i] = l[g];
                }
            }
        }
        if (PIPE) __syncthreads();  // every thread is done with this tile's buffer: the next bulk copy may overwrite it
    }
    if (live) {
        km.assign[i] = c;
        km.upper[i] = u;
        km.stale[i] = stale ? 1 : 0;
        if (c != c_prior) atomicAdd(km.reassigned, 1u);
    }
}

// ── recompute (elkan.rs:125-142): integer merge of member points into the new centroids ──
// Each block privatises the K x 102 (bins + weight) u32 accumulators in shared memory over a contiguous slice of
// points, then flushes non-zero cells with 64-bit global atomics: integer addition, so any order is exact.
// A warp walks 32 consecutive points.  Stored sorted by their first assignment (reorder_by_assignment), most of a warp's points belong to
// one cluster, and 32 lanes adding to the same (cluster, bin) cell is a 32-way shared-memory atomic conflict: 1.3 ms of a 4.3 ms iteration
// at 3 M points x k = 100 (profiles/r2ae_lloyd_k100_launches.csv).  So the warp first sums the counts of its LARGEST group of lanes that
// share a cluster (two REDUX per word of 4 bins: bytes widened to 16-bit pairs, 32 x 255 < 2^16) and four lanes add the sums; the other
// lanes (points that have moved since the reorder) add their own counts.  Integer sums: exact in any grouping.
template <bool SMEM>
__global__ void __launch_bounds__(256)
accumulate_kernel(KmDev km, int64_t per_block) {
    extern __shared__ __align__(16) unsigned int s_acc[];
    const int cells = km.k * (kBins + 1);
    if (SMEM) {
        for (int t = threadIdx.x; t < cells; t += blockDim.x) s_acc[t] = 0u;
        __syncthreads();
    }
    auto add_cell = [&](uint32_t c, int b, uint32_t v) {
        if (SMEM) atomicAdd(&s_acc[c * (kBins + 1) + b], v);
        else atomicAdd(km.acc + (size_t)c * (kBins + 1) + b, (unsigned long long)v);
    };
    const int lane = threadIdx.x & 31;
    const int64_t lo = blockIdx.x * per_block, hi = min(km.n, lo + per_block);
    for (int64_t base = lo + (threadIdx.x - lane); base < hi; base += blockDim.x) {
        const int64_t i = base + lane;
        const bool live = i < hi;
        const uint32_t c = live ? km.assign[i] : 0xFFFFFFFFu;
        uint32_t word[kRow / 4];                                       // the point's 112 bytes, loaded once
        {
            const uint4* row4 = reinterpret_cast<const uint4*>(km.pts + (size_t)(live ? i : base) * kRow);
#pragma unroll
            for (int q4 = 0; q4 < kRow / 16; ++q4) {
                const uint4 v4 = live ? row4[q4] : make_uint4(0u, 0u, 0u, 0u);
                word[4 * q4] = v4.x; word[4 * q4 + 1] = v4.y; word[4 * q4 + 2] = v4.z; word[4 * q4 + 3] = v4.w;
            }
        }
        // the largest group of lanes with one cluster; if the dead lanes of a tail warp (they share 0xFFFFFFFF) outnumber every live
        // group, there is no leader and every live lane takes the per-lane path
        const uint32_t same = __match_any_sync(0xFFFFFFFFu, c);
        const uint32_t most = __reduce_max_sync(0xFFFFFFFFu, (uint32_t)__popc(same));
        const uint32_t leaders = __ballot_sync(0xFFFFFFFFu, live && (uint32_t)__popc(same) == most);
        uint32_t grp = 0u, c0 = 0u;
        if (leaders) { const int lead = __ffs(leaders) - 1; grp = __shfl_sync(0xFFFFFFFFu, same, lead); c0 = __shfl_sync(0xFFFFFFFFu, c, lead); }
        const bool in = grp >> lane & 1u;
        if (__popc(grp) >= 4) {
            uint32_t w = 0;
#pragma unroll
            for (int q = 0; q < kRow / 4; ++q) {                       // bins 4q .. 4q+3
                const int valid = kBins - 4 * q;                       // how many of them exist (the row is padded to 112 bytes)
                if (valid <= 0) continue;
                uint32_t v = in ? word[q] : 0u;
                if (valid < 4) v &= (1u << (8 * valid)) - 1u;
                if (!__any_sync(0xFFFFFFFFu, v != 0u)) continue;
                const uint32_t even = __reduce_add_sync(0xFFFFFFFFu, v & 0x00FF00FFu);          // bins 4q (low half) and 4q+2 (high half)
                const uint32_t odd = __reduce_add_sync(0xFFFFFFFFu, (v >> 8) & 0x00FF00FFu);    // bins 4q+1 and 4q+3
                const uint32_t sum = lane == 0 ? (even & 0xFFFFu) : lane == 1 ? (odd & 0xFFFFu) : lane == 2 ? (even >> 16) : (odd >> 16);
                if (lane < 4 && lane < valid && sum) add_cell(c0, 4 * q + lane, sum);
                w += (even & 0xFFFFu) + (odd & 0xFFFFu) + (even >> 16) + (odd >> 16);
            }
            if (lane == 0) {
                add_cell(c0, kBins, w);
                atomicAdd(km.sizes + c0, (unsigned int)__popc(grp));
            }
        } else grp = 0u;
        if (live && !(grp >> lane & 1u)) {  // everyone else: per-lane atomics
            uint32_t w = 0;
#pragma unroll
            for (int q = 0; q < kRow / 4; ++q) {
                const uint32_t v = word[q];
                if (!v) continue;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t cnt = (v >> (8 * t)) & 0xFFu;
                    if (cnt && 4 * q + t < kBins) { add_cell(c, 4 * q + t, cnt); w += cnt; }
                }
            }
            add_cell(c, kBins, w);
            atomicAdd(km.sizes + c, 1u);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int t = threadIdx.x; t < cells; t += blockDim.x)
            if (s_acc[t]) atomicAdd(km.acc + t, (unsigned long long)s_acc[t]);
    }
}

// ── k-means++ (layer.rs:160-180) under the integer-weight contract of include/rbp.h ──
// potentials update against the newly chosen point + per-block integer sums for the next draw
__global__ void __launch_bounds__(kThreads)
pp_update_kernel(KmDev km, float* __restrict__ pot, const int64_t* __restrict__ pick, int first, unsigned long long* __restrict__ bsum) {
    __shared__ float s_row[kCdfRow];
    __shared__ unsigned long long s_sum[kThreads / 32];
    const int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x;
    const bool live = i < km.n;
    float p = 1.0f;
    if (!first) {
        const int64_t pk = *pick;
        if (threadIdx.x < 32) {  // CDF of the chosen point, once per block
            if (threadIdx.x == 0) {
                float Xp[kBins];
                load_point_cdf(km.pts + (size_t)pk * kRow, Xp);
#pragma unroll
                for (int b = 0; b < kBins; ++b) s_row[b] = Xp[b];
                for (int b = kBins; b < kCdfRow; ++b) s_row[b] = 0.0f;
            }
        }
        __syncthreads();
        if (live) {
            float X[kBins];
            load_point_cdf(km.pts + (size_t)i * kRow, X);
            float d[1];
            dist_group<1>(X, s_row, d);  // distance(&x, h): |a-b| is symmetric, so argument order is immaterial
            const float d2 = d[0] * d[0];
            p = pot[i];
            p = d2 < p ? d2 : p;         // Energy::min(d0, d1)
            if (i == pk) p = 0.0f;       // potentials[i] = 0 precedes the min and 0 survives it
            pot[i] = p;
        }
    } else if (live) {
        pot[i] = 1.0f;
    }
    unsigned long long q = live ? quantize_potential(p) : 0ull;
    for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, d);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kThreads / 32; ++w) t += s_sum[w];
        bsum[blockIdx.x] = t;
    }
}
// centroid j := point `pick` (counts + weight)
__global__ void set_centroid_kernel(KmDev km, const int64_t* __restrict__ pick, int j) {
    const int b = threadIdx.x;
    const uint8_t* row = km.pts + (size_t)(*pick) * kRow;
    __shared__ unsigned int s_w;
    if (b == 0) s_w = 0;
    __syncthreads();
    if (b < kBins) {
        km.ccount[(size_t)j * (kBins + 1) + b] = row[b];
        atomicAdd(&s_w, (unsigned int)row[b]);
    }
    __syncthreads();
    if (b == 0) km.ccount[(size_t)j * (kBins + 1) + kBins] = s_w;
}
// fold a still-pending drift into the stored bounds (only for state export / tests)
__global__ void materialize_kernel(KmDev km) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= km.n) return;
    for (int j = 0; j < km.k; ++j) {
        float v = km.lower[(size_t)j * km.n + i] - km.drift[j];
        km.lower[(size_t)j * km.n + i] = v > 0.0f ? v : 0.0f;
    }
    km.upper[i] += km.drift[km.assign[i]];
    km.stale[i] = 1;
}

}  // namespace rbp

using namespace rbp;

struct KmW1;
namespace rbp {
void w1_destroy(KmW1* h);
}

struct KmW1 : rbp_kmeans {
    KmDev d{};
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> owned;
    float* new_cdf = nullptr;
    float* pot = nullptr;
    unsigned long long* bsum = nullptr;
    int64_t* pick = nullptr;
    int32_t* chosen = nullptr;
    float* tri = nullptr;
    uint32_t* tmp_assign = nullptr;
    float* tmp_dist = nullptr;
    int nb = 0;
    size_t smem = 0;
    bool have_centroids = false, have_bounds = false;
    uint64_t dist_evals = 0;
    // After init_bounds the points are stored sorted by their first assignment (position p holds original point perm[p]), so that the 32 points
    // of a warp start in the same cluster and prune the same centroids: the step computes a group of distances whenever ANY lane of the warp
    // needs one, and with points in input order a warp's lanes wanted ~half of all groups between them (profiles/r2z).  Results are per point
    // and the merge is an integer sum: nothing depends on the storage order; every per-point output is returned in input order.
    uint8_t* pts_alt = nullptr;
    uint32_t *perm = nullptr, *perm_alt = nullptr, *order = nullptr, *iota = nullptr, *assign_alt = nullptr;
    float* upper_alt = nullptr;
    void* sort_tmp = nullptr;
    size_t sort_bytes = 0;
    bool permuted = false, reorder = true;
};

namespace {
template <class T>
int kalloc(KmW1* h, size_t n, T** out) {
    void* p = nullptr;
    RBP_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    h->owned.push_back(p);
    RBP_CUDA(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    *out = static_cast<T*>(p);
    return RBP_OK;
}
int refresh_centroid_tables(KmW1* h) {  // CDFs of the current centroids
    centroid_cdf_kernel<<<(h->d.k + 127) / 128, 128, 0, h->stream>>>(h->d.ccount, h->d.k, h->d.cdf);
    RBP_LAUNCHED();
    return RBP_OK;
}
}  // namespace

namespace rbp {

int w1_create(int kind, int64_t n, int k, int bins, const uint8_t* counts, int device, rbp_kmeans_t** out) {
    if (!out) return RBP_ERR_INVALID;
    *out = nullptr;
    if (kind != RBP_KMEANS_W1 || bins != kBins || n < 1 || k < 1 || k > n || !counts) {
        set_last_error("rbp_kmeans_create: only the W1 (Equity::variation, 101 bins) layer is built");
        return RBP_ERR_INVALID;
    }
    if (rbp_device_count() <= device) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    KmW1* h = new KmW1();
    h->kind = RBP_KMEANS_W1;
    h->k = k;
    if (const char* e = getenv("RBP_W1_REORDER")) h->reorder = atoi(e) != 0;
    h->device = device;
    auto fail = [&](int code) { w1_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fail(RBP_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(RBP_ERR_CUDA);
    KmDev& d = h->d;
    d.n = n; d.k = k;
    int st;
    uint8_t* pts = nullptr;
    if ((st = kalloc(h, (size_t)n * kRow, &pts))) return fail(st);
    if (cudaMemcpy2DAsync(pts, kRow, counts, bins, bins, n, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return fail(RBP_ERR_CUDA);
    d.pts = pts;
    if ((st = kalloc(h, (size_t)k * kCdfRow, &d.cdf))) return fail(st);
    if ((st = kalloc(h, (size_t)k * kCdfRow, &h->new_cdf))) return fail(st);
    d.kp = (k + 3) & ~3;
    if ((st = kalloc(h, (size_t)k * d.kp + 4, &d.pair))) return fail(st);
    if ((st = kalloc(h, (size_t)k, &d.mid))) return fail(st);
    if ((st = kalloc(h, (size_t)k, &d.drift))) return fail(st);
    if ((st = kalloc(h, (size_t)k * n, &d.lower))) return fail(st);
    if ((st = kalloc(h, (size_t)n, &d.upper))) return fail(st);
    if ((st = kalloc(h, (size_t)n, &d.assign))) return fail(st);
    if ((st = kalloc(h, (size_t)n, &d.stale))) return fail(st);
    // (+ k + 2 words of tail room: with a communicator the tallies ride in the same all-reduce, kmeans_api.cu)
    if ((st = kalloc(h, (size_t)k * (kBins + 1) + k + 2, &d.acc))) return fail(st);
    if ((st = kalloc(h, (size_t)k * (kBins + 1) + k + 2, &d.ccount))) return fail(st);
    if ((st = kalloc(h, (size_t)k, &d.sizes))) return fail(st);
    if ((st = kalloc(h, 1, &d.reassigned))) return fail(st);
    if ((st = kalloc(h, (size_t)n, &h->pot))) return fail(st);
    h->nb = (int)((n + kThreads - 1) / kThreads);
    if ((st = kalloc(h, (size_t)h->nb, &h->bsum))) return fail(st);
    if ((st = kalloc(h, 1, &h->pick))) return fail(st);
    if ((st = kalloc(h, (size_t)k, &h->chosen))) return fail(st);
    if ((st = kalloc(h, (size_t)k * (k - 1) / 2 + 1, &h->tri))) return fail(st);
    if ((st = kalloc(h, (size_t)n, &h->tmp_assign))) return fail(st);
    if ((st = kalloc(h, (size_t)n, &h->tmp_dist))) return fail(st);
    h->smem = (size_t)std::min(k, kTileK) * kCdfRow * sizeof(float);
    if (cudaFuncSetAttribute(assign_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem) != cudaSuccess ||
        cudaFuncSetAttribute(assign_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 1, kTileK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->smem + (size_t)k * sizeof(float))) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<256, 2, kTileK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->smem + (size_t)k * sizeof(float))) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 3, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->smem + (size_t)k * sizeof(float))) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 3, 128, false, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->smem + (size_t)k * sizeof(float))) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 3, 128, false, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->smem + (size_t)k * sizeof(float))) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 3, 96, false, 0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)96 * kCdfRow * 4 + ((size_t)k + 4) * sizeof(float) + (size_t)10 * 4 * 128 * 4)) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 3, 64, false, 0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)64 * kCdfRow * 4 + ((size_t)k + 4) * sizeof(float) + (size_t)18 * 4 * 128 * 4)) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 3, 56, true, 0, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * 56 * kCdfRow * 4 + ((size_t)k + 4) * sizeof(float) + (size_t)8 * 4 * 128 * 4)) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 3, 96, false, 0, 6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)96 * kCdfRow * 4 + ((size_t)k + 4) * sizeof(float) + (size_t)8 * 4 * 128 * 4)) != cudaSuccess ||
        cudaFuncSetAttribute(elkan_step_kernel<128, 3, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * 64 * kCdfRow * 4 + (size_t)k * sizeof(float))) != cudaSuccess ||
        cudaFuncSetAttribute(accumulate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return fail(RBP_ERR_CUDA);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(RBP_ERR_CUDA);
    *out = h;
    return RBP_OK;
}

void w1_destroy(KmW1* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (void* p : h->owned) cudaFree(p);
    delete h;
}

namespace { int restore_input_order(KmW1* h); }

int w1_init_pp(KmW1* h, uint64_t seed, int32_t* chosen_out) {
    if (!h) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    { const int st = restore_input_order(h); if (st) return st; }
    for (int r = 0; r < h->d.k; ++r) {
        pp_update_kernel<<<h->nb, kThreads, 0, h->stream>>>(h->d, h->pot, h->pick, r == 0, h->bsum);
        RBP_LAUNCHED();
        Philox4 w = philox4x32_10((uint32_t)r, 0u, 0u, TAG_KMEANSPP, (uint32_t)seed, (uint32_t)(seed >> 32));
        pp_pick_kernel<<<1, 1024, 0, h->stream>>>(h->d.n, kThreads, h->pot, h->bsum, h->nb, w.r[0], w.r[1], h->pick, r, h->chosen);
        RBP_LAUNCHED();
        set_centroid_kernel<<<1, 128, 0, h->stream>>>(h->d, h->pick, r);
        RBP_LAUNCHED();
    }
    h->dist_evals += (uint64_t)h->d.n * (h->d.k - 1);
    int st = refresh_centroid_tables(h);
    if (st) return st;
    if (chosen_out) RBP_CUDA(cudaMemcpyAsync(chosen_out, h->chosen, h->d.k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    h->have_centroids = true;
    return RBP_OK;
}

int w1_set_centroids(KmW1* h, const uint64_t* counts) {
    if (!h || !counts) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    std::vector<unsigned long long> host((size_t)h->d.k * (kBins + 1));
    for (int j = 0; j < h->d.k; ++j) {
        unsigned long long w = 0;
        for (int b = 0; b < kBins; ++b) { host[(size_t)j * (kBins + 1) + b] = counts[(size_t)j * kBins + b]; w += counts[(size_t)j * kBins + b]; }
        host[(size_t)j * (kBins + 1) + kBins] = w;
    }
    RBP_CUDA(cudaMemcpyAsync(h->d.ccount, host.data(), host.size() * 8, cudaMemcpyHostToDevice, h->stream));
    int st = refresh_centroid_tables(h);
    if (st) return st;
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    h->have_centroids = true;
    return RBP_OK;
}

__global__ void w1_iota_kernel(uint32_t* __restrict__ v, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}
// new position p <- old position order[p]: the 112-byte point row, its upper bound, and the composed permutation
__global__ void __launch_bounds__(256)
w1_gather_kernel(const uint8_t* __restrict__ pts, const float* __restrict__ upper, const uint32_t* __restrict__ perm_old, const uint32_t* __restrict__ order, int64_t n,
                 uint8_t* __restrict__ pts_out, float* __restrict__ upper_out, uint32_t* __restrict__ perm_out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t p = t / (kRow / 16), q = t % (kRow / 16);
    if (p >= n) return;
    const uint32_t src = order[p];
    reinterpret_cast<uint4*>(pts_out + (size_t)p * kRow)[q] = reinterpret_cast<const uint4*>(pts + (size_t)src * kRow)[q];
    if (q == 0) { upper_out[p] = upper[src]; perm_out[p] = perm_old ? perm_old[src] : src; }
}
// out[perm[p]] = in[p]: per-point results back in input order
template <class T>
__global__ void w1_unpermute_kernel(const T* __restrict__ in, const uint32_t* __restrict__ perm, int64_t n, T* __restrict__ out) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < n) out[perm[p]] = in[p];
}
namespace {
int reorder_by_assignment(KmW1* h) {  // right after init_bounds: every lower bound is 0 and nothing is stale, so only points / assign / upper move
    KmDev& d = h->d;
    if (!h->reorder || d.n < 1024) return RBP_OK;
    int st;
    if (!h->pts_alt) {
        if ((st = kalloc(h, (size_t)d.n * kRow, &h->pts_alt))) return st;
        if ((st = kalloc(h, (size_t)d.n, &h->perm))) return st;
        if ((st = kalloc(h, (size_t)d.n, &h->perm_alt))) return st;
        if ((st = kalloc(h, (size_t)d.n, &h->order))) return st;
        if ((st = kalloc(h, (size_t)d.n, &h->iota))) return st;
        if ((st = kalloc(h, (size_t)d.n, &h->assign_alt))) return st;
        if ((st = kalloc(h, (size_t)d.n, &h->upper_alt))) return st;
        RBP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, h->sort_bytes, d.assign, h->assign_alt, h->iota, h->order, (int)d.n, 0, 32, h->stream));
        if ((st = kalloc(h, h->sort_bytes, reinterpret_cast<unsigned char**>(&h->sort_tmp)))) return st;
    }
    int bits = 1;
    while ((1 << bits) < d.k) ++bits;
    w1_iota_kernel<<<(unsigned)((d.n + 255) / 256), 256, 0, h->stream>>>(h->iota, d.n);
    RBP_LAUNCHED();
    RBP_CUDA(cub::DeviceRadixSort::SortPairs(h->sort_tmp, h->sort_bytes, d.assign, h->assign_alt, h->iota, h->order, (int)d.n, 0, bits, h->stream));  // stable
    const int64_t threads = d.n * (kRow / 16);
    w1_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, h->stream>>>(d.pts, d.upper, h->permuted ? h->perm : nullptr, h->order, d.n, h->pts_alt, h->upper_alt, h->perm_alt);
    RBP_LAUNCHED();
    uint8_t* old_pts = const_cast<uint8_t*>(d.pts);
    d.pts = h->pts_alt; h->pts_alt = old_pts;
    std::swap(d.assign, h->assign_alt);
    std::swap(d.upper, h->upper_alt);
    std::swap(h->perm, h->perm_alt);
    h->permuted = true;
    return RBP_OK;
}
int restore_input_order(KmW1* h) {  // k-means++ draws by a prefix sum over the points IN INPUT ORDER (include/rbp.h contract)
    KmDev& d = h->d;
    if (!h->permuted) return RBP_OK;
    // order := inverse of perm, then the same gather
    w1_iota_kernel<<<(unsigned)((d.n + 255) / 256), 256, 0, h->stream>>>(h->iota, d.n);
    RBP_LAUNCHED();
    w1_unpermute_kernel<uint32_t><<<(unsigned)((d.n + 255) / 256), 256, 0, h->stream>>>(h->iota, h->perm, d.n, h->order);
    RBP_LAUNCHED();
    const int64_t threads = d.n * (kRow / 16);
    w1_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, h->stream>>>(d.pts, d.upper, nullptr, h->order, d.n, h->pts_alt, h->upper_alt, h->perm_alt);
    RBP_LAUNCHED();
    uint8_t* old_pts = const_cast<uint8_t*>(d.pts);
    d.pts = h->pts_alt; h->pts_alt = old_pts;
    h->permuted = false;
    h->have_bounds = false;  // assign / upper were not carried over
    return RBP_OK;
}
}  // namespace

int w1_init_bounds(KmW1* h) {
    if (!h) return RBP_ERR_INVALID;
    if (!h->have_centroids) { set_last_error("init_bounds before centroids"); return RBP_ERR_STATE; }
    RBP_CUDA(cudaSetDevice(h->device));
    assign_kernel<true><<<(unsigned)((h->d.n + kAssignThreads - 1) / kAssignThreads), kAssignThreads, h->smem, h->stream>>>(h->d, nullptr, nullptr);
    RBP_LAUNCHED();
    h->d.pending = 0;
    h->dist_evals += (uint64_t)h->d.n * h->d.k;
    { const int st = reorder_by_assignment(h); if (st) return st; }
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    h->have_bounds = true;
    return RBP_OK;
}

// point side of a step + local integer accumulation; multi-GPU hosts all-reduce the accumulator in between
int w1_step_local(KmW1* h) {
    if (!h) return RBP_ERR_INVALID;
    if (!h->have_bounds) { set_last_error("step before init_bounds"); return RBP_ERR_STATE; }
    RBP_CUDA(cudaSetDevice(h->device));
    KmDev& d = h->d;
    pairwise_kernel<<<d.k, 256, 0, h->stream>>>(d.cdf, d.k, d.kp, d.pair, d.mid);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemsetAsync(d.reassigned, 0, sizeof(uint32_t), h->stream));
    RBP_CUDA(cudaMemsetAsync(d.sizes, 0, d.k * sizeof(uint32_t), h->stream));
    RBP_CUDA(cudaMemsetAsync(d.acc, 0, (size_t)d.k * (kBins + 1) * 8, h->stream));
    {
        static const int variant = getenv("RBP_STEP_VARIANT") ? atoi(getenv("RBP_STEP_VARIANT")) : 8;  // 8 = 3 blocks/SM, TMA double-buffered 56-centroid tiles, bounds 6 groups ahead in a cp.async ring, fast path for complete groups (fastest measured: profiles/r2ac)
        const size_t drift_b = (size_t)d.k * sizeof(float);
        if (variant == 1) elkan_step_kernel<256, 2, kTileK><<<(unsigned)((d.n + 255) / 256), 256, h->smem + drift_b, h->stream>>>(d);
        else if (variant == 2) elkan_step_kernel<128, 3, 128><<<(unsigned)((d.n + 127) / 128), 128, (size_t)std::min(d.k, 128) * kCdfRow * 4 + drift_b, h->stream>>>(d);
        else if (variant == 4) elkan_step_kernel<128, 3, 128, false, 12><<<(unsigned)((d.n + 127) / 128), 128, (size_t)std::min(d.k, 128) * kCdfRow * 4 + drift_b, h->stream>>>(d);
        else if (variant == 5) elkan_step_kernel<128, 3, 128, false, 24><<<(unsigned)((d.n + 127) / 128), 128, (size_t)std::min(d.k, 128) * kCdfRow * 4 + drift_b, h->stream>>>(d);
        else if (variant == 6) elkan_step_kernel<128, 3, 96, false, 0, 8><<<(unsigned)((d.n + 127) / 128), 128, (size_t)std::min(d.k, 96) * kCdfRow * 4 + ((d.k + 3) & ~3) * sizeof(float) + (size_t)10 * 4 * 128 * 4, h->stream>>>(d);
        else if (variant == 7) elkan_step_kernel<128, 3, 64, false, 0, 16><<<(unsigned)((d.n + 127) / 128), 128, (size_t)std::min(d.k, 64) * kCdfRow * 4 + ((d.k + 3) & ~3) * sizeof(float) + (size_t)18 * 4 * 128 * 4, h->stream>>>(d);
        else if (variant == 8) elkan_step_kernel<128, 3, 56, true, 0, 6, true><<<(unsigned)((d.n + 127) / 128), 128, (size_t)2 * 56 * kCdfRow * 4 + ((d.k + 3) & ~3) * sizeof(float) + (size_t)8 * 4 * 128 * 4, h->stream>>>(d);
        else if (variant == 9) elkan_step_kernel<128, 3, 96, false, 0, 6, true><<<(unsigned)((d.n + 127) / 128), 128, (size_t)std::min(d.k, 96) * kCdfRow * 4 + ((d.k + 3) & ~3) * sizeof(float) + (size_t)8 * 4 * 128 * 4, h->stream>>>(d);
        else if (variant == 3) elkan_step_kernel<128, 3, 64, true><<<(unsigned)((d.n + 127) / 128), 128, (size_t)2 * 64 * kCdfRow * 4 + drift_b, h->stream>>>(d);
        else elkan_step_kernel<128, 1, kTileK><<<(unsigned)((d.n + 127) / 128), 128, h->smem + drift_b, h->stream>>>(d);
    }
    RBP_LAUNCHED();
    {
        const size_t acc_smem = (size_t)d.k * (kBins + 1) * sizeof(unsigned int);
        const int blocks = 148 * 2;
        const int64_t per_block = (d.n + blocks - 1) / blocks;
        if (acc_smem <= 200 * 1024) accumulate_kernel<true><<<blocks, 256, acc_smem, h->stream>>>(d, per_block);
        else accumulate_kernel<false><<<blocks, 256, 0, h->stream>>>(d, per_block);
        RBP_LAUNCHED();
    }
    return RBP_OK;
}
int w1_accumulator(KmW1* h, void** dev_ptr, size_t* bytes) {
    if (!h || !dev_ptr || !bytes) return RBP_ERR_INVALID;
    *dev_ptr = h->d.acc;
    *bytes = (size_t)h->d.k * (kBins + 1) * sizeof(unsigned long long);
    return RBP_OK;
}
int w1_counters(KmW1* h, void** dev_sizes, void** dev_reassigned) {
    if (!h) return RBP_ERR_INVALID;
    if (dev_sizes) *dev_sizes = h->d.sizes;
    if (dev_reassigned) *dev_reassigned = h->d.reassigned;
    return RBP_OK;
}
void* w1_stream(KmW1* h) { return h ? (void*)h->stream : nullptr; }

int w1_step_finish(KmW1* h, float* drift_out, uint32_t* sizes_out, uint32_t* reassigned_out) {
    if (!h) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    KmDev& d = h->d;
    centroid_cdf_kernel<<<(d.k + 127) / 128, 128, 0, h->stream>>>(d.acc, d.k, h->new_cdf);
    RBP_LAUNCHED();
    drift_kernel<<<(d.k + 127) / 128, 128, 0, h->stream>>>(h->new_cdf, d.cdf, d.k, d.drift);
    RBP_LAUNCHED();
    std::swap(d.cdf, h->new_cdf);
    std::swap(d.acc, d.ccount);
    d.pending = 1;
    if (drift_out) RBP_CUDA(cudaMemcpyAsync(drift_out, d.drift, d.k * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (sizes_out) RBP_CUDA(cudaMemcpyAsync(sizes_out, d.sizes, d.k * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (reassigned_out) RBP_CUDA(cudaMemcpyAsync(reassigned_out, d.reassigned, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return RBP_OK;
}

int w1_step(KmW1* h, float* drift_out, uint32_t* sizes_out, uint32_t* reassigned_out) {
    int st = w1_step_local(h);
    if (st) return st;
    return w1_step_finish(h, drift_out, sizes_out, reassigned_out);
}

int w1_assign(KmW1* h, uint32_t* assign_out, float* dist_out) {
    if (!h || !assign_out) return RBP_ERR_INVALID;
    if (!h->have_centroids) { set_last_error("assign before centroids"); return RBP_ERR_STATE; }
    RBP_CUDA(cudaSetDevice(h->device));
    assign_kernel<false><<<(unsigned)((h->d.n + kAssignThreads - 1) / kAssignThreads), kAssignThreads, h->smem, h->stream>>>(h->d, h->tmp_assign, h->tmp_dist);
    RBP_LAUNCHED();
    h->dist_evals += (uint64_t)h->d.n * h->d.k;
    const uint32_t* a_src = h->tmp_assign;
    const float* d_src = h->tmp_dist;
    if (h->permuted) {  // back to input order (the scratch of the reorder is free between init_bounds calls)
        const unsigned blocks = (unsigned)((h->d.n + 255) / 256);
        w1_unpermute_kernel<uint32_t><<<blocks, 256, 0, h->stream>>>(h->tmp_assign, h->perm, h->d.n, h->assign_alt);
        RBP_LAUNCHED();
        w1_unpermute_kernel<float><<<blocks, 256, 0, h->stream>>>(h->tmp_dist, h->perm, h->d.n, h->upper_alt);
        RBP_LAUNCHED();
        a_src = h->assign_alt; d_src = h->upper_alt;
    }
    RBP_CUDA(cudaMemcpyAsync(assign_out, a_src, h->d.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (dist_out) RBP_CUDA(cudaMemcpyAsync(dist_out, d_src, h->d.n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return RBP_OK;
}

int w1_centroids(KmW1* h, uint64_t* counts_out, uint64_t* weights_out) {
    if (!h) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    std::vector<unsigned long long> host((size_t)h->d.k * (kBins + 1));
    RBP_CUDA(cudaMemcpyAsync(host.data(), h->d.ccount, host.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    for (int j = 0; j < h->d.k; ++j) {
        if (counts_out) for (int b = 0; b < kBins; ++b) counts_out[(size_t)j * kBins + b] = host[(size_t)j * (kBins + 1) + b];
        if (weights_out) weights_out[j] = host[(size_t)j * (kBins + 1) + kBins];
    }
    return RBP_OK;
}

int w1_metric(KmW1* h, float* tri_out) {
    if (!h || !tri_out) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    const int total = h->d.k * (h->d.k - 1) / 2;
    if (total == 0) return RBP_OK;
    metric_kernel<<<(total + 127) / 128, 128, 0, h->stream>>>(h->d.cdf, h->d.k, h->tri);
    RBP_LAUNCHED();
    metric_normalize_kernel<<<1, 256, 0, h->stream>>>(h->tri, total);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemcpyAsync(tri_out, h->tri, total * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return RBP_OK;
}

int w1_bounds(KmW1* h, uint32_t* assign_out, float* upper_out, float* lower_out, uint8_t* stale_out) {
    if (!h) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    KmDev& d = h->d;
    if (d.pending) {
        materialize_kernel<<<(unsigned)((d.n + 255) / 256), 256, 0, h->stream>>>(d);
        RBP_LAUNCHED();
        d.pending = 0;
    }
    std::vector<uint32_t> perm;
    if (h->permuted) {
        perm.resize(d.n);
        RBP_CUDA(cudaMemcpyAsync(perm.data(), h->perm, d.n * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    std::vector<uint32_t> a(assign_out ? d.n : 0);
    std::vector<float> u(upper_out ? d.n : 0);
    std::vector<uint8_t> sflag(stale_out ? d.n : 0);
    if (assign_out) RBP_CUDA(cudaMemcpyAsync(a.data(), d.assign, d.n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (upper_out) RBP_CUDA(cudaMemcpyAsync(u.data(), d.upper, d.n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (stale_out) RBP_CUDA(cudaMemcpyAsync(sflag.data(), d.stale, d.n, cudaMemcpyDeviceToHost, h->stream));
    std::vector<float> kn(lower_out ? (size_t)d.k * d.n : 0);
    if (lower_out) RBP_CUDA(cudaMemcpyAsync(kn.data(), d.lower, kn.size() * 4, cudaMemcpyDeviceToHost, h->stream));  // device layout is [K][N]; the reference's Bounds is per point
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    for (int64_t p = 0; p < d.n; ++p) {  // position p holds input point perm[p]
        const int64_t i = h->permuted ? (int64_t)perm[p] : p;
        if (assign_out) assign_out[i] = a[p];
        if (upper_out) upper_out[i] = u[p];
        if (stale_out) stale_out[i] = sflag[p];
        if (lower_out) for (int j = 0; j < d.k; ++j) lower_out[(size_t)i * d.k + j] = kn[(size_t)j * d.n + p];
    }
    return RBP_OK;
}

int w1_timed(KmW1* h, int what, int iters, float* ms_out) {
    // what: 0 = full step, 1 = assign (N x K distances), device time by CUDA events on the library stream
    if (!h || !ms_out || iters < 1) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    cudaEvent_t e0, e1;
    RBP_CUDA(cudaEventCreate(&e0));
    RBP_CUDA(cudaEventCreate(&e1));
    RBP_CUDA(cudaEventRecord(e0, h->stream));
    for (int it = 0; it < iters; ++it) {
        int st;
        if (what == 0) st = w1_step(h, nullptr, nullptr, nullptr);
        else {
            assign_kernel<false><<<(unsigned)((h->d.n + kAssignThreads - 1) / kAssignThreads), kAssignThreads, h->smem, h->stream>>>(h->d, h->tmp_assign, h->tmp_dist);
            g_launches.fetch_add(1);
            st = cudaGetLastError() == cudaSuccess ? RBP_OK : RBP_ERR_CUDA;
        }
        if (st) return st;
    }
    RBP_CUDA(cudaEventRecord(e1, h->stream));
    RBP_CUDA(cudaEventSynchronize(e1));
    RBP_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return RBP_OK;
}

}  // namespace rbp
