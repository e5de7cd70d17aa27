// lloyd_sk.cu — the FLOP abstraction layer on sm_100a: Elkan k-means over histograms of next-street clusters under
// the Sinkhorn divergence with a learned ground metric (`Metric::emd` → `Sinkhorn::divergence`,
// crates/lloyd/src/metric.rs:109-115, sinkhorn.rs:166-171), plus the generic batched divergence entry point.
//
// Every distance is a full entropic-OT solve (≈10⁵–10⁷ flops), so the Elkan bookkeeping is negligible and the layer
// is organised around one unit of work: ONE WARP SOLVES ONE OT PROBLEM (sinkhorn.cuh).  A warp owns a point and
// walks the centroids sequentially exactly as `step_elkan` does (elkan.rs:113-123) — the pruning test is warp-uniform,
// so there is no speculation and no divergence; only when the test passes does the warp solve OT(x, c_j).  The
// reference memoises the self terms OT(h,h) per histogram (sinkhorn.rs:172-191); here they are computed once per
// point at creation and once per centroid per iteration.  Argument order is preserved where the divergence is not
// symmetric in floating point: `neighbor` calls distance(c, x), `refresh/rebound` distance(x, c) (elkan.rs:68-77,113-123).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include <cudaTypedefs.h>

#include "kmeans_common.cuh"
#include "sinkhorn.cuh"
#include "sk_screen.cuh"

namespace rbp {

constexpr int kSkWarps = 8;          // max warps (OT problems) per block; the launch picks blockDim = 32 x warps
constexpr int kSkMinBlocks = 4;       // register budget 64: ptxas serialises the 8 exp chains without a budget (see profiles/r1n)
constexpr int kPtMax = 64;           // max support of a point (flop children: 47, crates/deuce/src/street.rs:120-126)

struct SkDev {
    int64_t n;
    int k, bins;
    // points, sparse: ascending bucket ids + counts
    const uint8_t* p_idx;   // [N][kPtMax]
    const uint8_t* p_cnt;   // [N][kPtMax]
    const uint8_t* p_n;     // [N]
    const uint16_t* p_w;    // [N] weight = Σ counts
    float* p_self;          // [N] OT(x, x)
    // centroids: dense integer counts + weight
    unsigned long long* ccount;  // [K][bins + 1]
    unsigned long long* acc;     // [K][bins + 1]
    float* c_self;          // [K]
    float* new_self;        // [K]
    const float* tri;       // ground metric, dense symmetric [256][256] (sk_dense_tables)
    const float* reg;       // metric / temperature, same layout
    float* pair;            // [K][K]
    float* mid;             // [K]
    float* drift;           // [K]
    float* lower;           // [N][K]  (point-major: a warp owns a point)
    float* upper;           // [N]
    uint32_t* assign;       // [N]
    uint8_t* stale;         // [N]
    uint32_t* sizes;        // [K]
    uint32_t* reassigned;   // [1]
    int pending;
    SkParams hp;
    unsigned long long* stats;  // {solves, sweeps, exp terms}
    unsigned long long* queue;  // next unclaimed point of the running sweep (zeroed before each launch)
};

// points are claimed one at a time from a device counter: per-point work varies by orders of magnitude once Elkan
// prunes, and a point costs K full OT solves in the naive sweeps
__device__ __forceinline__ int64_t next_point(const SkDev& d, int lane) {
    unsigned long long i = 0;
    if (lane == 0) i = atomicAdd(d.queue, 1ull);
    return (int64_t)__shfl_sync(0xFFFFFFFFu, i, 0);
}
__device__ __forceinline__ int load_point(const SkDev& d, int64_t i, uint8_t* idx, float* lnd, int lane) {
    const int n = d.p_n[i];
    const float w = (float)d.p_w[i];
    for (int t = lane; t < n; t += 32) {
        idx[t] = d.p_idx[(size_t)i * kPtMax + t];
        lnd[t] = ln_c((float)d.p_cnt[(size_t)i * kPtMax + t] / w);
    }
    __syncwarp();
    return n;
}
__device__ __forceinline__ int load_centroid(const unsigned long long* __restrict__ counts, int bins, uint8_t* idx, float* lnd, int lane) {
    const float w = (float)counts[bins];
    return sk_load_side([&](int b) { return (float)counts[b]; }, w, bins, idx, lnd, lane);
}
template <class W>
__device__ __forceinline__ W& warp_scratch() {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    return reinterpret_cast<W*>(smem_raw)[threadIdx.x >> 5];
}

// OT(h,h) for points / for a centroid table
__global__ void __launch_bounds__(kSkWarps * 32, kSkMinBlocks) sk_self_points_kernel(SkDev d) {
    SkPoint& w = warp_scratch<SkPoint>();
    const int lane = threadIdx.x & 31;
    for (int64_t i = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); i < d.n; i += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        w.n_a = load_point(d, i, w.idx_a, w.lnd_a, lane);
        w.n_b = load_point(d, i, w.idx_b, w.lnd_b, lane);
        const float c = sk_solve(w.a(), w.b(), w.tile, d.tri, d.reg, d.hp, lane, d.stats);
        if (lane == 0) d.p_self[i] = c;
        __syncwarp();
    }
}
__global__ void __launch_bounds__(kSkWarps * 32, kSkMinBlocks) sk_self_centroids_kernel(SkDev d, const unsigned long long* __restrict__ counts, float* __restrict__ out) {
    SkWide& w = warp_scratch<SkWide>();
    const int lane = threadIdx.x & 31;
    for (int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < d.k; j += gridDim.x * (blockDim.x >> 5)) {
        const unsigned long long* c = counts + (size_t)j * (d.bins + 1);
        w.n_a = load_centroid(c, d.bins, w.idx_a, w.lnd_a, lane);
        w.n_b = load_centroid(c, d.bins, w.idx_b, w.lnd_b, lane);
        const float v = sk_solve(w.a(), w.b(), w.tile, d.tri, d.reg, d.hp, lane, d.stats);
        if (lane == 0) out[j] = v;
        __syncwarp();
    }
}

// centroid-vs-centroid divergences: out[t] = divergence(A[ia[t]], B[ib[t]])  (pairwises, drift, metric)
__global__ void __launch_bounds__(kSkWarps * 32, kSkMinBlocks)
sk_centroid_pairs_kernel(SkDev d, const unsigned long long* __restrict__ A, const float* __restrict__ selfA, const unsigned long long* __restrict__ B,
                         const float* __restrict__ selfB, int mode, int total, float* __restrict__ out) {
    // mode 0: all (i, j) of the K x K pairwise table (diagonal = 0);  mode 1: drift, t -> (t, t)
    SkWide& w = warp_scratch<SkWide>();
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < total; t += gridDim.x * (blockDim.x >> 5)) {
        const int i = mode == 0 ? t / d.k : t, j = mode == 0 ? t % d.k : t;
        float v = 0.0f;
        if (!(mode == 0 && i == j)) {  // elkan.rs:91-93 pairwise(i, i) = 0
            w.n_a = load_centroid(A + (size_t)i * (d.bins + 1), d.bins, w.idx_a, w.lnd_a, lane);
            w.n_b = load_centroid(B + (size_t)j * (d.bins + 1), d.bins, w.idx_b, w.lnd_b, lane);
            v = sk_divergence(sk_solve(w.a(), w.b(), w.tile, d.tri, d.reg, d.hp, lane, d.stats), selfA[i], selfB[j]);
        }
        if (lane == 0) out[t] = v;
        __syncwarp();
    }
}
// elkan.rs:95-105 midpoints
__global__ void sk_mid_kernel(const float* __restrict__ pair, int k, float* __restrict__ mid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    float m = 3.402823466e+38f;
    for (int j = 0; j < k; ++j)
        if (j != i) { const float h = pair[(size_t)i * k + j] * 0.5f; m = h < m ? h : m; }
    mid[i] = m;
}
// layer.rs:85-101 + metric.rs:127-141: tri[t] = (pair[i][j] + pair[j][i]) / 2, then / max
__global__ void sk_metric_kernel(const float* __restrict__ pair, int k, float* __restrict__ tri) {
    __shared__ float s_max[256];
    const int total = k * (k - 1) / 2;
    float m = 1.17549435e-38f;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        int hi = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
        while (hi * (hi - 1) / 2 > t) --hi;
        while ((hi + 1) * hi / 2 <= t) ++hi;
        const int lo = t - hi * (hi - 1) / 2;
        const float v = (pair[(size_t)hi * k + lo] + pair[(size_t)lo * k + hi]) / 2.0f;
        tri[t] = v;
        m = v > m ? v : m;
    }
    s_max[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s && s_max[threadIdx.x + s] > s_max[threadIdx.x]) s_max[threadIdx.x] = s_max[threadIdx.x + s];
        __syncthreads();
    }
    const float mx = s_max[0];
    for (int t = threadIdx.x; t < total; t += blockDim.x) tri[t] = tri[t] / mx;
}

// k-means++ potentials: pot_i = min(pot_i, divergence(x_pick, h_i)^2)   (layer.rs:166-178)
__global__ void __launch_bounds__(kSkWarps * 32, kSkMinBlocks) sk_pp_update_kernel(SkDev d, float* __restrict__ pot, const int64_t* __restrict__ pick, int first) {
    SkPoint& w = warp_scratch<SkPoint>();
    const int lane = threadIdx.x & 31;
    for (int64_t i = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); i < d.n; i += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        if (first) { if (lane == 0) pot[i] = 1.0f; continue; }
        const int64_t pk = *pick;
        w.n_a = load_point(d, pk, w.idx_a, w.lnd_a, lane);
        w.n_b = load_point(d, i, w.idx_b, w.lnd_b, lane);
        const float dist = sk_divergence(sk_solve(w.a(), w.b(), w.tile, d.tri, d.reg, d.hp, lane, d.stats), d.p_self[pk], d.p_self[i]);
        if (lane == 0) {
            const float d2 = dist * dist;
            float p = pot[i];
            p = d2 < p ? d2 : p;
            if (i == pk) p = 0.0f;
            pot[i] = p;
        }
        __syncwarp();
    }
}
__global__ void sk_set_centroid_kernel(SkDev d, const int64_t* __restrict__ pick, int j) {
    const int64_t p = *pick;
    unsigned long long* dst = d.ccount + (size_t)j * (d.bins + 1);
    for (int b = threadIdx.x; b <= d.bins; b += blockDim.x) dst[b] = 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = d.p_n[p];
        for (int t = 0; t < n; ++t) dst[d.p_idx[(size_t)p * kPtMax + t]] = d.p_cnt[(size_t)p * kPtMax + t];
        dst[d.bins] = d.p_w[p];
    }
}

// naive argmin over all centroids with distance(c_j, x): init_bounds (elkan.rs:39-47) and lookup (layer.rs:44-60)
// With `approx` (the tensor-core screen, sk_screen.cuh) only the centroids whose approximate divergence is within `margin` of
// the point's smallest one are evaluated — exactly, in centroid order, with the same first-minimum rule.
template <bool INIT_BOUNDS>
__global__ void __launch_bounds__(kSkWarps * 32, kSkMinBlocks) sk_assign_kernel(SkDev d, uint32_t* __restrict__ out_assign, float* __restrict__ out_dist,
                                                                                const float* __restrict__ approx, float margin) {
    SkPoint& w = warp_scratch<SkPoint>();
    const int lane = threadIdx.x & 31;
    for (int64_t i = next_point(d, lane); i < d.n; i = next_point(d, lane)) {
        float best = 0.0f;
        int bestj = -1;
        w.n_a = load_point(d, i, w.idx_a, w.lnd_a, lane);                                             // nu = point, once
        const float self_i = d.p_self[i];
        float bar = INFINITY;
        if (approx) {
            float amin = INFINITY;
            for (int j = lane; j < d.k; j += 32) amin = fminf(amin, approx[(size_t)i * d.k + j]);
            for (int o = 16; o > 0; o >>= 1) amin = fminf(amin, __shfl_xor_sync(0xFFFFFFFFu, amin, o));
            bar = amin + margin;
        }
        for (int j = 0; j < d.k; ++j) {
            if (approx && approx[(size_t)i * d.k + j] > bar) continue;  // screened out (warp-uniform)
            w.n_b = load_centroid(d.ccount + (size_t)j * (d.bins + 1), d.bins, w.idx_b, w.lnd_b, lane);  // mu = centroid
            const float dist = sk_divergence(sk_solve(w.b(), w.a(), w.tile, d.tri, d.reg, d.hp, lane, d.stats), d.c_self[j], self_i);
            if (bestj < 0 || dist < best) { best = dist; bestj = j; }
            __syncwarp();
        }
        if (INIT_BOUNDS) {
            for (int j = lane; j < d.k; j += 32) d.lower[(size_t)i * d.k + j] = 0.0f;
            if (lane == 0) { d.assign[i] = (uint32_t)bestj; d.upper[i] = best; d.stale[i] = 0; }
        } else if (lane == 0) {
            out_assign[i] = (uint32_t)bestj;
            if (out_dist) out_dist[i] = best;
        }
    }
}

// one Elkan step, point side (elkan.rs:153-164): a warp owns a point; all tests are warp-uniform
__global__ void __launch_bounds__(kSkWarps * 32, kSkMinBlocks) sk_step_kernel(SkDev d) {
    SkPoint& w = warp_scratch<SkPoint>();
    const int lane = threadIdx.x & 31;
    for (int64_t i = next_point(d, lane); i < d.n; i = next_point(d, lane)) {
        uint32_t c = d.assign[i];
        const uint32_t c_prior = c;
        float u = d.upper[i];
        bool stale = d.stale[i] != 0;
        float* low = d.lower + (size_t)i * d.k;
        if (d.pending) {  // Bounds::update of the previous step (bounds.rs:65-74), applied on first touch
            for (int j = lane; j < d.k; j += 32) { const float v = low[j] - d.drift[j]; low[j] = v > 0.0f ? v : 0.0f; }
            u += d.drift[c];
            stale = true;
            __syncwarp();
        }
        if (u > d.mid[c]) {  // step_elkan filter
            w.n_a = load_point(d, i, w.idx_a, w.lnd_a, lane);  // mu = point, once
            const float self_i = d.p_self[i];
            auto dist_to = [&](int j) {  // distance(x, c_j): mu = point, nu = centroid
                w.n_b = load_centroid(d.ccount + (size_t)j * (d.bins + 1), d.bins, w.idx_b, w.lnd_b, lane);
                const float v = sk_divergence(sk_solve(w.a(), w.b(), w.tile, d.tri, d.reg, d.hp, lane, d.stats), self_i, d.c_self[j]);
                __syncwarp();
                return v;
            };
            if (stale) {  // refresh (elkan.rs:113-117, bounds.rs:76-80)
                const float v = dist_to((int)c);
                if (lane == 0) low[c] = v;
                u = v; stale = false;
                __syncwarp();
            }
            for (int j = 0; j < d.k; ++j) {  // rebound (elkan.rs:118-123, bounds.rs:57-61,81-87)
                if ((uint32_t)j != c && u > low[j] && u > 0.5f * d.pair[(size_t)c * d.k + j]) {
                    const float v = dist_to(j);
                    if (lane == 0) low[j] = v;
                    if (v < u) { c = (uint32_t)j; u = v; }
                    __syncwarp();
                }
            }
        }
        if (lane == 0) {
            d.assign[i] = c; d.upper[i] = u; d.stale[i] = stale ? 1 : 0;
            if (c != c_prior) atomicAdd(d.reassigned, 1u);
        }
        __syncwarp();
    }
}

// recompute (elkan.rs:125-142): integer merge of sparse member points
__global__ void sk_accumulate_kernel(SkDev d) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t c = d.assign[i];
        unsigned long long* dst = d.acc + (size_t)c * (d.bins + 1);
        const int n = d.p_n[i];
        for (int t = 0; t < n; ++t) atomicAdd(dst + d.p_idx[(size_t)i * kPtMax + t], (unsigned long long)d.p_cnt[(size_t)i * kPtMax + t]);
        atomicAdd(dst + d.bins, (unsigned long long)d.p_w[i]);
        atomicAdd(d.sizes + c, 1u);
    }
}
__global__ void sk_materialize_kernel(SkDev d) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    for (int j = 0; j < d.k; ++j) { const float v = d.lower[(size_t)i * d.k + j] - d.drift[j]; d.lower[(size_t)i * d.k + j] = v > 0.0f ? v : 0.0f; }
    d.upper[i] += d.drift[d.assign[i]];
    d.stale[i] = 1;
}

// generic batch: out[t] = divergence(A[ia[t]], B[ib[t]]) over dense u32 histograms
__global__ void __launch_bounds__(kSkWarps * 32, kSkMinBlocks)
sk_batch_kernel(const uint32_t* __restrict__ A, const uint32_t* __restrict__ B, int bins, const int32_t* __restrict__ ia, const int32_t* __restrict__ ib,
                int64_t n, int mode, const float* __restrict__ selfA, const float* __restrict__ selfB, const float* __restrict__ tri,
                const float* __restrict__ reg, SkParams hp, float* __restrict__ out) {
    // mode 0: self costs of A (n = |A|); mode 1: divergences of the listed pairs
    SkWide& w = warp_scratch<SkWide>();
    const int lane = threadIdx.x & 31;
    for (int64_t t = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); t < n; t += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int a = mode == 0 ? (int)t : ia[t], b = mode == 0 ? (int)t : ib[t];
        const uint32_t* ha = A + (size_t)a * bins;
        const uint32_t* hb = (mode == 0 ? A : B) + (size_t)b * bins;
        float wa = 0.0f, wb = 0.0f;
        {   // weights: exact integer sums, then one conversion (bins.rs weight as f32)
            unsigned long long sa = 0, sb = 0;
            for (int q = lane; q < bins; q += 32) { sa += ha[q]; sb += hb[q]; }
            for (int s = 16; s > 0; s >>= 1) { sa += __shfl_xor_sync(0xFFFFFFFFu, sa, s); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, s); }
            wa = (float)sa; wb = (float)sb;
        }
        w.n_a = sk_load_side([&](int q) { return (float)ha[q]; }, wa, bins, w.idx_a, w.lnd_a, lane);
        w.n_b = sk_load_side([&](int q) { return (float)hb[q]; }, wb, bins, w.idx_b, w.lnd_b, lane);
        const float c = sk_solve(w.a(), w.b(), w.tile, tri, reg, hp, lane);
        if (lane == 0) out[t] = mode == 0 ? c : sk_divergence(c, selfA[a], selfB[b]);
        __syncwarp();
    }
}

}  // namespace rbp

using namespace rbp;

struct KmSk : rbp_kmeans {
    SkDev d{};
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> owned;
    float* pot = nullptr;
    unsigned long long* bsum = nullptr;
    int64_t* pick = nullptr;
    int32_t* chosen = nullptr;
    float* tri_out = nullptr;
    float* tri_dev = nullptr;
    float* reg_dev = nullptr;
    uint32_t* tmp_assign = nullptr;
    float* tmp_dist = nullptr;
    int nb = 0, grid = 0, warps = kSkWarps, threads = kSkWarps * 32;
    size_t smem = 0, smem_wide = 0;  // dynamic shared memory of the point-class / centroid-class kernels
    bool have_metric = false, have_centroids = false, have_bounds = false;
    // tensor-core screen of the naive sweeps (sk_screen.cuh); margin < 0 = off
    float screen_margin = -1.0f;
    int tiles = 0, sms = 148;
    __nv_bfloat16 *gb = nullptr, *gch = nullptr, *gcl = nullptr;
    float *nu_t = nullptr, *c_inv_n = nullptr, *approx = nullptr;
    unsigned long long *squeue = nullptr, *sstats = nullptr;
    CUtensorMap nu_map{};
    std::vector<float> host_metric, host_reg;  // the dense tables (kept for the screen's bf16 Gibbs kernel)
};

namespace rbp {

void sk_destroy(KmSk* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    for (void* p : h->owned) cudaFree(p);
    delete h;
}
namespace {
template <class T>
int salloc(KmSk* h, size_t n, T** out) {
    void* p = nullptr;
    RBP_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    h->owned.push_back(p);
    RBP_CUDA(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    *out = static_cast<T*>(p);
    return RBP_OK;
}
template <class T>
int supload(KmSk* h, const std::vector<T>& v, const T** out) {
    T* p = nullptr;
    int st = salloc(h, v.size(), &p);
    if (st) return st;
    RBP_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    *out = p;
    return RBP_OK;
}
int centroid_selfs(KmSk* h, const unsigned long long* counts, float* out) {
    sk_self_centroids_kernel<<<std::min(h->grid, (h->d.k + h->warps - 1) / h->warps), h->threads, h->smem_wide, h->stream>>>(h->d, counts, out);
    RBP_LAUNCHED();
    return RBP_OK;
}
}  // namespace

int sk_create(int64_t n, int k, int bins, const uint8_t* counts, int device, rbp_kmeans_t** out) {
    if (bins < 2 || bins > kSkMaxSupport || n < 1 || k < 1 || k > n || !counts) return RBP_ERR_INVALID;
    KmSk* h = new KmSk();
    h->kind = RBP_KMEANS_SINKHORN;
    h->k = k;
    h->device = device;
    auto fail = [&](int code) { sk_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fail(RBP_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(RBP_ERR_CUDA);
    std::vector<uint8_t> idx((size_t)n * kPtMax, 0), cnt((size_t)n * kPtMax, 0), pn(n, 0);
    std::vector<uint16_t> pw(n, 0);
    for (int64_t i = 0; i < n; ++i) {
        int m = 0, w = 0;
        for (int b = 0; b < bins; ++b) {
            const uint8_t c = counts[(size_t)i * bins + b];
            if (!c) continue;
            if (m >= kPtMax) { set_last_error("point support exceeds 64 buckets"); return fail(RBP_ERR_CAPACITY); }
            idx[(size_t)i * kPtMax + m] = (uint8_t)b; cnt[(size_t)i * kPtMax + m] = c; ++m; w += c;
        }
        if (m == 0) { set_last_error("empty histogram"); return fail(RBP_ERR_INVALID); }
        pn[i] = (uint8_t)m; pw[i] = (uint16_t)w;
    }
    SkDev& d = h->d;
    d.n = n; d.k = k; d.bins = bins;
    d.hp = SkParams{0.025f, 128, 0.0005f};  // lloyd/src/hyperparams/sinkhorn.rs:17-23
    int st;
    if ((st = supload(h, idx, &d.p_idx))) return fail(st);
    if ((st = supload(h, cnt, &d.p_cnt))) return fail(st);
    if ((st = supload(h, pn, &d.p_n))) return fail(st);
    if ((st = supload(h, pw, &d.p_w))) return fail(st);
    if ((st = salloc(h, (size_t)n, &d.p_self))) return fail(st);
    // (+ k + 2 words of tail room: with a communicator the tallies ride in the same all-reduce, kmeans_api.cu)
    if ((st = salloc(h, (size_t)k * (bins + 1) + k + 2, &d.ccount))) return fail(st);
    if ((st = salloc(h, (size_t)k * (bins + 1) + k + 2, &d.acc))) return fail(st);
    if ((st = salloc(h, (size_t)k, &d.c_self))) return fail(st);
    if ((st = salloc(h, (size_t)k, &d.new_self))) return fail(st);
    if ((st = salloc(h, (size_t)k * k, &d.pair))) return fail(st);
    if ((st = salloc(h, (size_t)k, &d.mid))) return fail(st);
    if ((st = salloc(h, (size_t)k, &d.drift))) return fail(st);
    if ((st = salloc(h, (size_t)n * k, &d.lower))) return fail(st);
    if ((st = salloc(h, (size_t)n, &d.upper))) return fail(st);
    if ((st = salloc(h, (size_t)n, &d.assign))) return fail(st);
    if ((st = salloc(h, (size_t)n, &d.stale))) return fail(st);
    if ((st = salloc(h, (size_t)k, &d.sizes))) return fail(st);
    if ((st = salloc(h, 1, &d.reassigned))) return fail(st);
    if ((st = salloc(h, (size_t)n, &h->pot))) return fail(st);
    h->nb = (int)((n + 127) / 128);
    if ((st = salloc(h, (size_t)h->nb, &h->bsum))) return fail(st);
    if ((st = salloc(h, 1, &h->pick))) return fail(st);
    if ((st = salloc(h, (size_t)k, &h->chosen))) return fail(st);
    if ((st = salloc(h, (size_t)k * (k - 1) / 2 + 1, &h->tri_out))) return fail(st);
    if ((st = salloc(h, (size_t)kSkLd * kSkLd, &h->tri_dev))) return fail(st);
    if ((st = salloc(h, (size_t)kSkLd * kSkLd, &h->reg_dev))) return fail(st);
    if ((st = salloc(h, (size_t)n, &h->tmp_assign))) return fail(st);
    if ((st = salloc(h, (size_t)n, &h->tmp_dist))) return fail(st);
    d.tri = h->tri_dev; d.reg = h->reg_dev;
    if ((st = salloc(h, 3, &d.stats))) return fail(st);
    if ((st = salloc(h, 1, &d.queue))) return fail(st);
    // launch shape: 8 warps x 3 blocks per SM by default (5.2 KB of scratch per warp for point problems); RBP_SK_WARPS / RBP_SK_BLOCKS_PER_SM
    // override it for tuning runs
    int sms = 148, blocks_per_sm = 3;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    h->sms = sms;
    if (const char* e = getenv("RBP_SK_WARPS")) h->warps = std::max(1, std::min(kSkWarps, atoi(e)));
    if (const char* e = getenv("RBP_SK_BLOCKS_PER_SM")) blocks_per_sm = std::max(1, std::min(16, atoi(e)));
    h->threads = h->warps * 32;
    h->smem = h->warps * sizeof(SkPoint);
    h->smem_wide = h->warps * sizeof(SkWide);
    h->grid = sms * blocks_per_sm;
    const void* kernels[] = {(const void*)sk_self_points_kernel, (const void*)sk_self_centroids_kernel, (const void*)sk_centroid_pairs_kernel,
                             (const void*)sk_pp_update_kernel, (const void*)sk_assign_kernel<true>, (const void*)sk_assign_kernel<false>,
                             (const void*)sk_step_kernel, (const void*)sk_batch_kernel};
    for (const void* f : kernels)
        if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_wide) != cudaSuccess) return fail(RBP_ERR_CUDA);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(RBP_ERR_CUDA);
    *out = h;
    return RBP_OK;
}

int sk_set_metric(KmSk* h, const float* tri, int bins) {
    if (!h || !tri || bins != h->d.bins) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    if (!sk_metric_valid(tri, bins)) { set_last_error("ground metric entries must be finite and non-negative"); return RBP_ERR_INVALID; }
    std::vector<float>& metric = h->host_metric;
    std::vector<float>& reg = h->host_reg;
    metric.assign((size_t)kSkLd * kSkLd, 0.0f); reg.assign((size_t)kSkLd * kSkLd, 0.0f);
    sk_dense_tables(tri, bins, h->d.hp.temperature, metric.data(), reg.data());
    RBP_CUDA(cudaMemcpyAsync(h->tri_dev, metric.data(), metric.size() * 4, cudaMemcpyHostToDevice, h->stream));
    RBP_CUDA(cudaMemcpyAsync(h->reg_dev, reg.data(), reg.size() * 4, cudaMemcpyHostToDevice, h->stream));
    sk_self_points_kernel<<<h->grid, h->threads, h->smem, h->stream>>>(h->d);
    RBP_LAUNCHED();
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    h->have_metric = true;
    return RBP_OK;
}

int sk_init_pp(KmSk* h, uint64_t seed, int32_t* chosen_out) {
    if (!h->have_metric) { set_last_error("set the ground metric first"); return RBP_ERR_STATE; }
    RBP_CUDA(cudaSetDevice(h->device));
    for (int r = 0; r < h->d.k; ++r) {
        sk_pp_update_kernel<<<h->grid, h->threads, h->smem, h->stream>>>(h->d, h->pot, h->pick, r == 0);
        RBP_LAUNCHED();
        pp_blocksum_kernel<<<h->nb, 128, 0, h->stream>>>(h->pot, h->d.n, h->bsum);
        RBP_LAUNCHED();
        Philox4 w = philox4x32_10((uint32_t)r, 0u, 0u, TAG_KMEANSPP, (uint32_t)seed, (uint32_t)(seed >> 32));
        pp_pick_kernel<<<1, 1024, 0, h->stream>>>(h->d.n, 128, h->pot, h->bsum, h->nb, w.r[0], w.r[1], h->pick, r, h->chosen);
        RBP_LAUNCHED();
        sk_set_centroid_kernel<<<1, 128, 0, h->stream>>>(h->d, h->pick, r);
        RBP_LAUNCHED();
    }
    int st = centroid_selfs(h, h->d.ccount, h->d.c_self);
    if (st) return st;
    if (chosen_out) RBP_CUDA(cudaMemcpyAsync(chosen_out, h->chosen, h->d.k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    h->have_centroids = true;
    return RBP_OK;
}

int sk_set_centroids(KmSk* h, const uint64_t* counts) {
    if (!h->have_metric) return RBP_ERR_STATE;
    RBP_CUDA(cudaSetDevice(h->device));
    const int B = h->d.bins;
    std::vector<unsigned long long> host((size_t)h->d.k * (B + 1));
    for (int j = 0; j < h->d.k; ++j) {
        unsigned long long w = 0;
        for (int b = 0; b < B; ++b) { host[(size_t)j * (B + 1) + b] = counts[(size_t)j * B + b]; w += counts[(size_t)j * B + b]; }
        host[(size_t)j * (B + 1) + B] = w;
    }
    RBP_CUDA(cudaMemcpyAsync(h->d.ccount, host.data(), host.size() * 8, cudaMemcpyHostToDevice, h->stream));
    int st = centroid_selfs(h, h->d.ccount, h->d.c_self);
    if (st) return st;
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    h->have_centroids = true;
    return RBP_OK;
}

// the screen's view of the centroids: nu^T[tile][y][j] = density of bucket y in centroid tile*128 + j (bins.rs:58-60), and 1 / |support|
__global__ void __launch_bounds__(256)
sk_screen_prepare_kernel(const unsigned long long* __restrict__ ccount, int k, int bins, int tiles, float* __restrict__ nu_t, float* __restrict__ c_inv_n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = tiles * skt::kBinsT * skt::kLanes;
    if (t < total) {
        const int j = t % skt::kLanes, y = (t / skt::kLanes) % skt::kBinsT, tile = t / (skt::kLanes * skt::kBinsT), c = tile * skt::kLanes + j;
        float v = 0.0f;
        if (c < k && y < bins) {
            const unsigned long long* row = ccount + (size_t)c * (bins + 1);
            if (row[y]) v = (float)row[y] / (float)row[bins];
        }
        nu_t[t] = v;
    }
    if (t < tiles * skt::kLanes) {
        int n = 0;
        if (t < k) for (int y = 0; y < bins; ++y) n += ccount[(size_t)t * (bins + 1) + y] != 0ull;
        c_inv_n[t] = n ? 1.0f / (float)n : 0.0f;
    }
}
namespace {
int screen_setup(KmSk* h) {  // buffers, the bf16 Gibbs kernel and the TMA descriptor of the centroid tile
    if (h->approx) return RBP_OK;
    const SkDev& d = h->d;
    h->tiles = (d.k + skt::kLanes - 1) / skt::kLanes;
    int st;
    if ((st = salloc(h, (size_t)skt::kBinsT * skt::kBinsT, &h->gb))) return st;
    if ((st = salloc(h, (size_t)skt::kBinsT * skt::kBinsT, &h->gch))) return st;
    if ((st = salloc(h, (size_t)skt::kBinsT * skt::kBinsT, &h->gcl))) return st;
    if ((st = salloc(h, (size_t)h->tiles * skt::kBinsT * skt::kLanes, &h->nu_t))) return st;
    if ((st = salloc(h, (size_t)h->tiles * skt::kLanes, &h->c_inv_n))) return st;
    if ((st = salloc(h, (size_t)h->tiles, &h->squeue))) return st;
    if ((st = salloc(h, 4, &h->sstats))) return st;
    if ((st = salloc(h, (size_t)d.n * d.k, &h->approx))) return st;
    // G = exp(-C/T) rounded once to bf16 (a fixed perturbation of the cost by <= T * 2^-9), and G∘C as bf16 hi + lo planes
    std::vector<__nv_bfloat16> gb((size_t)skt::kBinsT * skt::kBinsT), gh(gb.size()), gl(gb.size());
    for (size_t t = 0; t < gb.size(); ++t) {
        const __nv_bfloat16 g = __float2bfloat16_rn((float)std::exp(-(double)h->host_reg[t]));
        const float gc = __bfloat162float(g) * h->host_metric[t];
        gb[t] = g;
        gh[t] = __float2bfloat16_rn(gc);
        gl[t] = __float2bfloat16_rn(gc - __bfloat162float(gh[t]));
    }
    RBP_CUDA(cudaMemcpyAsync(h->gb, gb.data(), gb.size() * 2, cudaMemcpyHostToDevice, h->stream));
    RBP_CUDA(cudaMemcpyAsync(h->gch, gh.data(), gh.size() * 2, cudaMemcpyHostToDevice, h->stream));
    RBP_CUDA(cudaMemcpyAsync(h->gcl, gl.data(), gl.size() * 2, cudaMemcpyHostToDevice, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    // TMA descriptor: nu^T as a 2-D fp32 tensor [tiles * 256 rows][128], boxes of 128 x 128 (the driver entry point is fetched
    // through the runtime: the library does not link libcuda)
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    RBP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { set_last_error("cuTensorMapEncodeTiled is not available in this driver"); return RBP_ERR_CUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)skt::kLanes, (cuuint64_t)h->tiles * skt::kBinsT};
    const cuuint64_t strides[1] = {(cuuint64_t)skt::kLanes * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)skt::kLanes, 128u}, estr[2] = {1u, 1u};
    const CUresult r = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn)(&h->nu_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->nu_t, dims, strides, box, estr,
                                                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"); return RBP_ERR_CUDA; }
    RBP_CUDA(cudaFuncSetAttribute((const void*)skt::sk_screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)skt::kSmemTotal));
    return RBP_OK;
}
// approximate divergences of points [0, m) against every current centroid → h->approx[m][k]
int screen_run(KmSk* h, int64_t m) {
    int st = screen_setup(h);
    if (st) return st;
    const SkDev& d = h->d;
    RBP_CUDA(cudaMemsetAsync(h->squeue, 0, (size_t)h->tiles * 8, h->stream));
    const int total = h->tiles * skt::kBinsT * skt::kLanes;
    sk_screen_prepare_kernel<<<(total + 255) / 256, 256, 0, h->stream>>>(d.ccount, d.k, d.bins, h->tiles, h->nu_t, h->c_inv_n);
    RBP_LAUNCHED();
    skt::ScreenArgs a{};
    a.n = m; a.k = d.k; a.tiles = h->tiles; a.iterations = d.hp.iterations; a.tolerance = d.hp.tolerance;
    a.p_idx = d.p_idx; a.p_cnt = d.p_cnt; a.p_n = d.p_n; a.p_w = d.p_w; a.pt_stride = kPtMax;
    a.p_self = d.p_self; a.c_self = d.c_self; a.gb = h->gb; a.gch = h->gch; a.gcl = h->gcl; a.c_inv_n = h->c_inv_n;
    a.approx = h->approx; a.queue = h->squeue; a.stats = h->sstats;
    if (const char* e = getenv("RBP_SCREEN_DEBUG")) a.debug = atoi(e);
    const int grid = std::max(h->tiles, (h->sms / h->tiles) * h->tiles);  // one CTA per SM (all of TMEM, 182 KB of shared memory), a multiple of the tile count
    skt::sk_screen_kernel<<<grid, skt::kThreads, skt::kSmemTotal, h->stream>>>(h->nu_map, a);
    RBP_LAUNCHED();
    return RBP_OK;
}
// a pipeline time-out inside the screen kernel (a barrier that never completed) is an error, never a silent fallback
int screen_check(KmSk* h) {
    unsigned long long code = 0;
    RBP_CUDA(cudaMemcpyAsync(&code, h->sstats + 2, 8, cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    if (code) {
        static const char* stage[] = {"", "TMA load of the centroid tile", "centroid-side MMA", "point-side MMA", "cost MMA"};
        set_last_error(std::string("sk_screen_kernel: time-out waiting for the ") + stage[(code >> 32) & 7] + " at point " + std::to_string((unsigned)(code & 0xFFFFFFFFu)));
        return RBP_ERR_CUDA;
    }
    return RBP_OK;
}
}  // namespace

int sk_screen(KmSk* h, float margin) {
    if (!h->have_metric) { set_last_error("set the ground metric first"); return RBP_ERR_STATE; }
    if (h->d.bins > skt::kBinsT) return RBP_ERR_CAPACITY;
    RBP_CUDA(cudaSetDevice(h->device));
    if (margin >= 0.0f) { const int st = screen_setup(h); if (st) return st; }
    h->screen_margin = margin;
    return RBP_OK;
}
int sk_screen_probe(KmSk* h, int64_t m, float* out, uint64_t* stats2) {
    if (!h->have_centroids || !out || m < 1 || m > h->d.n) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    int st = screen_setup(h);
    if (st) return st;
    RBP_CUDA(cudaMemsetAsync(h->sstats, 0, 32, h->stream));
    if ((st = screen_run(h, m))) return st;
    RBP_CUDA(cudaMemcpyAsync(out, h->approx, (size_t)m * h->d.k * 4, cudaMemcpyDeviceToHost, h->stream));
    if (stats2) RBP_CUDA(cudaMemcpyAsync(stats2, h->sstats, 16, cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return screen_check(h);
}

int sk_init_bounds(KmSk* h) {
    if (!h->have_centroids) return RBP_ERR_STATE;
    RBP_CUDA(cudaSetDevice(h->device));
    const bool screened = h->screen_margin >= 0.0f;
    if (screened) { int st = screen_run(h, h->d.n); if (st) return st; if ((st = screen_check(h))) return st; }
    RBP_CUDA(cudaMemsetAsync(h->d.queue, 0, 8, h->stream));
    sk_assign_kernel<true><<<h->grid, h->threads, h->smem, h->stream>>>(h->d, nullptr, nullptr, screened ? h->approx : nullptr, h->screen_margin);
    RBP_LAUNCHED();
    h->d.pending = 0;
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    h->have_bounds = true;
    return RBP_OK;
}

int sk_step_local(KmSk* h) {
    if (!h->have_bounds) return RBP_ERR_STATE;
    RBP_CUDA(cudaSetDevice(h->device));
    SkDev& d = h->d;
    sk_centroid_pairs_kernel<<<h->grid, h->threads, h->smem_wide, h->stream>>>(d, d.ccount, d.c_self, d.ccount, d.c_self, 0, d.k * d.k, d.pair);
    RBP_LAUNCHED();
    sk_mid_kernel<<<(d.k + 127) / 128, 128, 0, h->stream>>>(d.pair, d.k, d.mid);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemsetAsync(d.reassigned, 0, 4, h->stream));
    RBP_CUDA(cudaMemsetAsync(d.sizes, 0, d.k * 4, h->stream));
    RBP_CUDA(cudaMemsetAsync(d.acc, 0, (size_t)d.k * (d.bins + 1) * 8, h->stream));
    RBP_CUDA(cudaMemsetAsync(d.queue, 0, 8, h->stream));
    sk_step_kernel<<<h->grid, h->threads, h->smem, h->stream>>>(d);
    RBP_LAUNCHED();
    sk_accumulate_kernel<<<148 * 4, 256, 0, h->stream>>>(d);
    RBP_LAUNCHED();
    return RBP_OK;
}

int sk_step_finish(KmSk* h, float* drift_out, uint32_t* sizes_out, uint32_t* reassigned_out) {
    RBP_CUDA(cudaSetDevice(h->device));
    SkDev& d = h->d;
    int st = centroid_selfs(h, d.acc, d.new_self);
    if (st) return st;
    // drift_j = distance(new_j, old_j)  (elkan.rs:107-109)
    sk_centroid_pairs_kernel<<<h->grid, h->threads, h->smem_wide, h->stream>>>(d, d.acc, d.new_self, d.ccount, d.c_self, 1, d.k, d.drift);
    RBP_LAUNCHED();
    std::swap(d.acc, d.ccount);
    std::swap(d.new_self, d.c_self);
    d.pending = 1;
    if (drift_out) RBP_CUDA(cudaMemcpyAsync(drift_out, d.drift, d.k * 4, cudaMemcpyDeviceToHost, h->stream));
    if (sizes_out) RBP_CUDA(cudaMemcpyAsync(sizes_out, d.sizes, d.k * 4, cudaMemcpyDeviceToHost, h->stream));
    if (reassigned_out) RBP_CUDA(cudaMemcpyAsync(reassigned_out, d.reassigned, 4, cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return RBP_OK;
}

int sk_assign(KmSk* h, uint32_t* assign_out, float* dist_out) {
    if (!h->have_centroids) return RBP_ERR_STATE;
    RBP_CUDA(cudaSetDevice(h->device));
    const bool screened = h->screen_margin >= 0.0f;
    if (screened) { int st = screen_run(h, h->d.n); if (st) return st; if ((st = screen_check(h))) return st; }
    RBP_CUDA(cudaMemsetAsync(h->d.queue, 0, 8, h->stream));
    sk_assign_kernel<false><<<h->grid, h->threads, h->smem, h->stream>>>(h->d, h->tmp_assign, h->tmp_dist, screened ? h->approx : nullptr, h->screen_margin);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemcpyAsync(assign_out, h->tmp_assign, h->d.n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (dist_out) RBP_CUDA(cudaMemcpyAsync(dist_out, h->tmp_dist, h->d.n * 4, cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return RBP_OK;
}

int sk_centroids(KmSk* h, uint64_t* counts_out, uint64_t* weights_out) {
    RBP_CUDA(cudaSetDevice(h->device));
    const int B = h->d.bins;
    std::vector<unsigned long long> host((size_t)h->d.k * (B + 1));
    RBP_CUDA(cudaMemcpyAsync(host.data(), h->d.ccount, host.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    for (int j = 0; j < h->d.k; ++j) {
        if (counts_out) for (int b = 0; b < B; ++b) counts_out[(size_t)j * B + b] = host[(size_t)j * (B + 1) + b];
        if (weights_out) weights_out[j] = host[(size_t)j * (B + 1) + B];
    }
    return RBP_OK;
}

int sk_metric(KmSk* h, float* tri_out) {
    RBP_CUDA(cudaSetDevice(h->device));
    SkDev& d = h->d;
    const int total = d.k * (d.k - 1) / 2;
    if (total == 0) return RBP_OK;
    sk_centroid_pairs_kernel<<<h->grid, h->threads, h->smem_wide, h->stream>>>(d, d.ccount, d.c_self, d.ccount, d.c_self, 0, d.k * d.k, d.pair);
    RBP_LAUNCHED();
    sk_metric_kernel<<<1, 256, 0, h->stream>>>(d.pair, d.k, h->tri_out);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemcpyAsync(tri_out, h->tri_out, total * 4, cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return RBP_OK;
}

int sk_bounds(KmSk* h, uint32_t* assign_out, float* upper_out, float* lower_out, uint8_t* stale_out) {
    RBP_CUDA(cudaSetDevice(h->device));
    SkDev& d = h->d;
    if (d.pending) {
        sk_materialize_kernel<<<(unsigned)((d.n + 255) / 256), 256, 0, h->stream>>>(d);
        RBP_LAUNCHED();
        d.pending = 0;
    }
    if (assign_out) RBP_CUDA(cudaMemcpyAsync(assign_out, d.assign, d.n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (upper_out) RBP_CUDA(cudaMemcpyAsync(upper_out, d.upper, d.n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (stale_out) RBP_CUDA(cudaMemcpyAsync(stale_out, d.stale, d.n, cudaMemcpyDeviceToHost, h->stream));
    if (lower_out) RBP_CUDA(cudaMemcpyAsync(lower_out, d.lower, (size_t)d.n * d.k * 4, cudaMemcpyDeviceToHost, h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return RBP_OK;
}

int sk_accumulator(KmSk* h, void** dev_ptr, size_t* bytes) {
    *dev_ptr = h->d.acc;
    *bytes = (size_t)h->d.k * (h->d.bins + 1) * 8;
    return RBP_OK;
}
int sk_counters(KmSk* h, void** dev_sizes, void** dev_reassigned) {
    if (dev_sizes) *dev_sizes = h->d.sizes;
    if (dev_reassigned) *dev_reassigned = h->d.reassigned;
    return RBP_OK;
}
void* sk_stream(KmSk* h) { return (void*)h->stream; }
int sk_stats(KmSk* h, uint64_t* out, int reset) {
    RBP_CUDA(cudaSetDevice(h->device));
    if (out) RBP_CUDA(cudaMemcpyAsync(out, h->d.stats, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    if (reset) RBP_CUDA(cudaMemsetAsync(h->d.stats, 0, 3 * sizeof(uint64_t), h->stream));
    RBP_CUDA(cudaStreamSynchronize(h->stream));
    return RBP_OK;
}

int sk_timed(KmSk* h, int what, int iters, float* ms_out) {
    RBP_CUDA(cudaSetDevice(h->device));
    cudaEvent_t e0, e1;
    RBP_CUDA(cudaEventCreate(&e0));
    RBP_CUDA(cudaEventCreate(&e1));
    RBP_CUDA(cudaEventRecord(e0, h->stream));
    for (int it = 0; it < iters; ++it) {
        int st;
        if (what == 0) { st = sk_step_local(h); if (!st) st = sk_step_finish(h, nullptr, nullptr, nullptr); }
        else {
            RBP_CUDA(cudaMemsetAsync(h->d.queue, 0, 8, h->stream));
            sk_assign_kernel<false><<<h->grid, h->threads, h->smem, h->stream>>>(h->d, h->tmp_assign, h->tmp_dist, nullptr, 0.0f);
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

// `Metric::emd` for explicit pairs (SURVEY §8b "EMD" row): dense u32 histograms, pair list, triangular ground metric
int sk_batch(const uint32_t* a_counts, int na, const uint32_t* b_counts, int nb, int bins, const int32_t* ia, const int32_t* ib, int64_t n,
             const float* tri, float temperature, int iterations, float tolerance, float* out) {
    if (bins < 2 || bins > kSkMaxSupport || na < 1 || nb < 1 || n < 0 || !a_counts || !b_counts || !tri || (n > 0 && (!ia || !ib || !out)))
        return RBP_ERR_INVALID;
    if (rbp_device_count() < 1) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    if (n == 0) return RBP_OK;
    if (!(temperature > 0.0f) || !sk_metric_valid(tri, bins)) { set_last_error("temperature must be positive, ground metric entries finite and non-negative"); return RBP_ERR_INVALID; }
    const size_t T = (size_t)kSkLd * kSkLd;
    std::vector<float> metric(T), reg(T);
    sk_dense_tables(tri, bins, temperature, metric.data(), reg.data());
    uint32_t *dA = nullptr, *dB = nullptr;
    int32_t *dia = nullptr, *dib = nullptr;
    float *dtri = nullptr, *dreg = nullptr, *dsa = nullptr, *dsb = nullptr, *dout = nullptr;
    RBP_CUDA(cudaMalloc(&dA, (size_t)na * bins * 4)); RBP_CUDA(cudaMalloc(&dB, (size_t)nb * bins * 4));
    RBP_CUDA(cudaMalloc(&dia, n * 4)); RBP_CUDA(cudaMalloc(&dib, n * 4));
    RBP_CUDA(cudaMalloc(&dtri, T * 4)); RBP_CUDA(cudaMalloc(&dreg, T * 4));
    RBP_CUDA(cudaMalloc(&dsa, na * 4)); RBP_CUDA(cudaMalloc(&dsb, nb * 4)); RBP_CUDA(cudaMalloc(&dout, n * 4));
    RBP_CUDA(cudaMemcpy(dA, a_counts, (size_t)na * bins * 4, cudaMemcpyHostToDevice));
    RBP_CUDA(cudaMemcpy(dB, b_counts, (size_t)nb * bins * 4, cudaMemcpyHostToDevice));
    RBP_CUDA(cudaMemcpy(dia, ia, n * 4, cudaMemcpyHostToDevice));
    RBP_CUDA(cudaMemcpy(dib, ib, n * 4, cudaMemcpyHostToDevice));
    RBP_CUDA(cudaMemcpy(dtri, metric.data(), T * 4, cudaMemcpyHostToDevice));
    RBP_CUDA(cudaMemcpy(dreg, reg.data(), T * 4, cudaMemcpyHostToDevice));
    const size_t smem = kSkWarps * sizeof(SkWide);
    RBP_CUDA(cudaFuncSetAttribute(sk_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SkParams hp{temperature, iterations, tolerance};
    const int grid = 148 * 2;
    sk_batch_kernel<<<grid, kSkWarps * 32, smem>>>(dA, dA, bins, nullptr, nullptr, na, 0, nullptr, nullptr, dtri, dreg, hp, dsa);
    RBP_LAUNCHED();
    sk_batch_kernel<<<grid, kSkWarps * 32, smem>>>(dB, dB, bins, nullptr, nullptr, nb, 0, nullptr, nullptr, dtri, dreg, hp, dsb);
    RBP_LAUNCHED();
    sk_batch_kernel<<<grid, kSkWarps * 32, smem>>>(dA, dB, bins, dia, dib, n, 1, dsa, dsb, dtri, dreg, hp, dout);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dia); cudaFree(dib); cudaFree(dtri); cudaFree(dreg); cudaFree(dsa); cudaFree(dsb); cudaFree(dout);
    return RBP_OK;
}

}  // namespace rbp
