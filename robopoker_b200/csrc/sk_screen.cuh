// sk_screen.cuh — tensor-core SCREEN for the flop layer's naive N x K sweeps (`Elkan::init_bounds`, crates/elkan/src/elkan.rs:39-47,
// 68-77; `Layer::lookup`, crates/lloyd/src/layer.rs:62-84).  Those sweeps need, per point, only argmin_j and that one distance.
// This kernel computes an APPROXIMATE Sinkhorn divergence of a point against 128 centroids at a time in the scaling domain
// (u = e^phi, v = e^psi, G = exp(-C/T)): the two softmin half-steps of sinkhorn.rs:96-129 become the contractions
//     Q[j, y] = sum_x U[j, x] G[x, y]      V = nu (/) Q        (centroid side first: distance(c, x) has mu = centroid)
//     R[j, x] = sum_y V[j, y] G[y, x]      U = mu (/) R
// i.e. "dense cost-matrix x scaling-vector" products — GEMMs with M = 128 centroids (one per TMEM lane), on tcgen05.mma:
//   * A operands (the scaling vectors, rewritten every half-step) live in TENSOR MEMORY as bf16 hi + lo planes (written by the
//     epilogue threads with tcgen05.st, 2^-17 relative), accumulators in TMEM fp32, read back with tcgen05.ld;
//   * B operands are the rows of G selected by the point's support (K-major, no-swizzle core-matrix layout) in shared memory;
//   * the centroid density tile nu^T [256][128] fp32 (128 KB) is staged ONCE per CTA by TMA (cp.async.bulk.tensor + mbarrier).
// One thread owns one centroid column: every per-pair reduction (the L1 stopping rule of sinkhorn.rs:85-94, the cost read-out
// u^T (G∘C) v) is thread-local.  The exact log-domain warp kernel (sinkhorn.cuh) then re-evaluates only the centroids whose
// approximate divergence is within `margin` of the point's minimum, in centroid order with the reference's first-minimum
// rule — so assignments and winning distances stay bit-identical to the full exact sweep as long as margin >= 2 x the
// screen's error (measured by rbp_kmeans_screen_probe; tests/test_sinkhorn_gpu.py pins both).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace rbp {
namespace skt {

constexpr int kLanes = 128;     // centroids per tile = TMEM lanes
constexpr int kThreads = 256;   // two threads per centroid column: warps w and w + 4 share TMEM lane quadrant w and split the bins
constexpr int kBinsT = 256;     // histogram bins (KMEANS_MAX_CLUSTER_COUNT)
constexpr int kMaxSx = 48;      // padded support of a point (flop children: 47)
// TMEM columns (512 allocated)
constexpr uint32_t kColVH = 0, kColVL = 128, kColQ = 256, kColR = 384, kColUH = 432, kColUL = 456;
// shared memory
constexpr uint32_t kSmemNu = kBinsT * kLanes * 4;             // 131072
constexpr uint32_t kSmemB = kMaxSx * 512;                     // 24576 each (B1 / B2; reused for the hi / lo cost tiles)
constexpr uint32_t kSmemMisc = 2048;
constexpr uint32_t kSmemTotal = kSmemNu + 2 * kSmemB + kSmemMisc + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded: a descriptor or pipeline mistake must surface as an error code, not as a hung GPU (returns false after ~0.2 s)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
        if (clock64() - t0 > 400000000ll) return false;
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] · B[smem]^T, bf16 inputs, fp32 accumulation; one thread issues
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): fp32 accumulate, bf16 x bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) { return 1u << 4 | 1u << 7 | 1u << 10 | (n >> 3) << 17 | (128u >> 4) << 24; }
// shared-memory descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes, contiguous (128 B); `lbo` = byte stride
// between the two 16-byte K chunks of one MMA, `sbo` = byte stride between 8-row groups  (UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (uint64_t)(lbo >> 4) << 16 | (uint64_t)(sbo >> 4) << 32 | 1ull << 46;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
                 "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// v = hi + lo with hi, lo bf16: pairs (even element in the low half) as the MMA reads them from a 32-bit TMEM column
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float2 join_pair(uint32_t hi, uint32_t lo) {
    const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi)), l = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lo));
    return make_float2(h.x + l.x, h.y + l.y);
}

struct ScreenArgs {
    int64_t n;                 // points
    int k;                     // centroids
    int tiles;                 // ceil(k / 128)
    int iterations;
    float tolerance;
    const uint8_t *p_idx, *p_cnt, *p_n;   // sparse points (lloyd_sk.cu SkDev), row stride `pt_stride`
    const uint16_t* p_w;
    int pt_stride;
    const float *p_self, *c_self;
    const __nv_bfloat16* gb;   // [256][256] bf16(exp(-C/T)), symmetric
    const __nv_bfloat16 *gch, *gcl;  // hi / lo planes of G∘C
    const float* c_inv_n;      // [tiles * 128] 1 / |support| of the centroid; 0 = padding lane or empty centroid
    float* approx;             // [n][k] approximate divergences (NaN-free: a failed column reads 0 = "always a candidate")
    unsigned long long* queue; // [tiles] next unclaimed point of each tile's stream
    int debug;                 // bring-up: stop after 1 = barriers + TMEM allocation, 2 = + TMA load, 3 = + TMEM store/load round trip, 4 = + the first MMA
    unsigned long long* stats; // [0] (point, tile) problems, [1] iterations summed over them, [2] first pipeline time-out (0 = none): stage << 32 | point
};

// One CTA = one centroid tile (its nu^T staged once by TMA) x a stream of points.
__global__ void __launch_bounds__(kThreads, 1)
sk_screen_kernel(const __grid_constant__ CUtensorMap nu_map, ScreenArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_nu = reinterpret_cast<float*>(smem);                                  // [256 y][128 j]
    unsigned char* s_b1 = smem + kSmemNu;                                           // [32 chunks][sxp rows][16 B]   rows of G picked by the support
    unsigned char* s_b2 = s_b1 + kSmemB;                                            // [sxp/8 chunks][256 rows][16 B] the same, transposed
    unsigned char* misc = s_b2 + kSmemB;
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(misc);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(misc + 16);
    unsigned long long* s_next = reinterpret_cast<unsigned long long*>(misc + 24);
    uint8_t* s_idx = misc + 64;                                                     // [64]
    float* s_mu = reinterpret_cast<float*>(misc + 128);                             // [64]
    float* s_err = reinterpret_cast<float*>(misc + 512);                           // [128] the upper half's share of a column's L1 change
    uint8_t* s_frozen = misc + 1024;                                                // [128] column converged (set by the lower half)
    const int tid = threadIdx.x, warp = tid >> 5, tile = blockIdx.x % a.tiles;
    const int col = tid & (kLanes - 1), part = tid >> 7;                            // TMEM lane = centroid column; which half of the bins
    const int j = tile * kLanes + col;                                              // this thread's centroid
    if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);          // this warp's quadrant of TMEM lanes
    if (a.debug == 1) {
        if (tid == 0 && blockIdx.x == 0) a.approx[0] = __uint_as_float(tmem);
        __syncthreads();
        if (warp == 0) tmem_dealloc(tmem, 512);
        return;
    }
    if (tid == 0) {  // the tile's densities, once: 2 boxes of [128 y][128 j] fp32
        mbar_expect_tx(bar_tma, kSmemNu);
        tma_load_2d(s_nu, &nu_map, 0, tile * kBinsT, bar_tma);
        tma_load_2d(s_nu + 128 * kLanes, &nu_map, 0, tile * kBinsT + 128, bar_tma);
    }
    bool alive = mbar_wait(bar_tma, 0);
    if (!alive && tid == 0) atomicCAS(&a.stats[2], 0ull, 1ull << 32);
    if (a.debug == 2 || a.debug == 3) {
        float sum = 0.0f;
        if (alive) for (int y = 0; y < kBinsT; ++y) sum += s_nu[y * kLanes + col];   // = 1 for a real centroid
        if (a.debug == 3 && alive) {  // TMEM round trip: store 16 words, load them back
            uint32_t w[16], r[16];
            for (int q = 0; q < 16; ++q) w[q] = (uint32_t)(tid * 100 + q);
            tmem_st16(lane_addr + kColQ + 16 * part, w);
            tmem_st_wait();
            tmem_ld16(lane_addr + kColQ + 16 * part, r);
            tmem_ld_wait();
            for (int q = 0; q < 16; ++q) if (r[q] != w[q]) sum = -1000.0f;
        }
        if (blockIdx.x < a.tiles && j < a.k && (part == 0 || sum < 0.0f)) a.approx[j] = sum;
        tc_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc(tmem, 512);
        return;
    }
    const float inv_n = a.c_inv_n[j];
    const bool real = j < a.k && inv_n > 0.0f;
    const float self_c = j < a.k ? a.c_self[j] : 0.0f;
    uint32_t phase = 0;
    unsigned long long n_prob = 0, n_iter = 0;
    while (alive) {
        if (tid == 0) *s_next = atomicAdd(&a.queue[tile], 1ull);
        __syncthreads();
        const int64_t i = (int64_t)*s_next;
        if (i >= a.n) break;
        const int sx = a.p_n[i], sxp = (sx + 15) & ~15;
        if (tid < 64) {
            const bool live = tid < sx;
            s_idx[tid] = live ? a.p_idx[(size_t)i * a.pt_stride + tid] : 0;
            s_mu[tid] = live ? (float)a.p_cnt[(size_t)i * a.pt_stride + tid] / (float)a.p_w[i] : 0.0f;   // bins.rs:58-60 density
        }
        __syncthreads();
        // B1[c][n] = G[idx_n][8c .. 8c+7]; rows n >= sx are zero
        for (int t = tid; t < sxp * 32; t += kThreads) {
            const int n = t >> 5, c = t & 31;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (n < sx) v = __ldg(reinterpret_cast<const uint4*>(a.gb + (size_t)s_idx[n] * kBinsT + c * 8));
            *reinterpret_cast<uint4*>(s_b1 + ((size_t)c * sxp + n) * 16) = v;
        }
        // B2[c][y] = G[y][idx_{8c} .. idx_{8c+7}] = G[idx][y] (symmetric)
        for (int t = tid; t < (sxp >> 3) * kBinsT; t += kThreads) {
            const int y = t & (kBinsT - 1), c = t >> 8;
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int x0 = c * 8 + 2 * q, x1 = x0 + 1;
                const uint16_t lo = x0 < sx ? __ldg(reinterpret_cast<const uint16_t*>(a.gb) + (size_t)s_idx[x0] * kBinsT + y) : (uint16_t)0;
                const uint16_t hi = x1 < sx ? __ldg(reinterpret_cast<const uint16_t*>(a.gb) + (size_t)s_idx[x1] * kBinsT + y) : (uint16_t)0;
                w[q] = (uint32_t)lo | (uint32_t)hi << 16;
            }
            *reinterpret_cast<uint4*>(s_b2 + ((size_t)c * kBinsT + y) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        // initial potentials (phi.rs:25-30 uniform): u = 1/|supp x|, v = 1/|supp c| on the supports
        float u[kMaxSx];
        const float u0 = 1.0f / (float)sx;
#pragma unroll
        for (int x = 0; x < kMaxSx; ++x) u[x] = x < sx ? u0 : 0.0f;
#pragma unroll
        for (int x0 = 0; x0 < kMaxSx; x0 += 16) {
            if (x0 < sxp && part == 0) {   // (part is warp-uniform: the TMEM accesses stay warp-collective)
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) split_pair(u[x0 + 2 * q], u[x0 + 2 * q + 1], hi[q], lo[q]);
                tmem_st8(lane_addr + kColUH + (x0 >> 1), hi);
                tmem_st8(lane_addr + kColUL + (x0 >> 1), lo);
            }
        }
        for (int y0 = 128 * part; y0 < 128 * part + 128; y0 += 32) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float v0 = s_nu[(y0 + 2 * q) * kLanes + col] > 0.0f ? inv_n : 0.0f, v1 = s_nu[(y0 + 2 * q + 1) * kLanes + col] > 0.0f ? inv_n : 0.0f;
                split_pair(v0, v1, hi[q], lo[q]);
            }
            tmem_st16(lane_addr + kColVH + (y0 >> 1), hi);
            tmem_st16(lane_addr + kColVL + (y0 >> 1), lo);
        }
        tmem_st_wait();
        fence_proxy_async();   // B1 / B2 were written through the generic proxy; the tensor core reads them through the async proxy
        tc_fence_before();
        __syncthreads();
        bool frozen = !real;
        if (part == 0) s_frozen[col] = frozen;
        int it = 0;
        for (; it < a.iterations; ++it) {
            float err = 0.0f;
            // centroid side: Q = U·G, V = nu / Q — two halves of 128 bins through one 128-column accumulator
            for (int h = 0; h < 2; ++h) {
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t idesc = make_idesc(128);
                    for (int s = 0; s < (sxp >> 4); ++s) {
                        const uint64_t bd = make_sdesc(smem_u32(s_b2) + (uint32_t)h * 128u * 16u + (uint32_t)s * 2u * (kBinsT * 16u), kBinsT * 16u, 128u);
                        umma_ts(tmem + kColQ, tmem + kColUH + s * 8, bd, idesc, s > 0);
                        umma_ts(tmem + kColQ, tmem + kColUL + s * 8, bd, idesc, 1u);
                    }
                    umma_commit(bar_mma);
                }
                if (!mbar_wait(bar_mma, phase)) { alive = false; if (tid == 0) atomicCAS(&a.stats[2], 0ull, 2ull << 32 | (unsigned long long)i); }
                if (!__syncthreads_and(alive)) { alive = false; break; }
                phase ^= 1u;
                tc_fence_after();
                if (a.debug == 4) {  // Q[j, y] of the first half-step: (1/|supp x|) * sum_x G[x, y]
                    uint32_t q[32];
                    tmem_ld32(lane_addr + kColQ, q);
                    tmem_ld_wait();
                    if (j < a.k && i == 0 && h == 0 && part == 0) for (int y = 0; y < 32; ++y) a.approx[(size_t)j * 32 + y] = __uint_as_float(q[y]);
                    alive = false;
                    break;
                }
                for (int y0 = 64 * part; y0 < 64 * part + 64; y0 += 32) {
                    uint32_t q[32], oh[16], ol[16], nh[16], nl[16];
                    tmem_ld32(lane_addr + kColQ + y0, q);
                    tmem_ld16(lane_addr + kColVH + ((h * 128 + y0) >> 1), oh);
                    tmem_ld16(lane_addr + kColVL + ((h * 128 + y0) >> 1), ol);
                    tmem_ld_wait();
#pragma unroll
                    for (int p = 0; p < 16; ++p) {
                        const int y = h * 128 + y0 + 2 * p;
                        const float n0 = s_nu[y * kLanes + col], n1 = s_nu[(y + 1) * kLanes + col];
                        const float v0 = n0 > 0.0f ? __fdividef(n0, __uint_as_float(q[2 * p])) : 0.0f;
                        const float v1 = n1 > 0.0f ? __fdividef(n1, __uint_as_float(q[2 * p + 1])) : 0.0f;
                        const float2 old = join_pair(oh[p], ol[p]);
                        err += fabsf(v0 - old.x) + fabsf(v1 - old.y);
                        split_pair(v0, v1, nh[p], nl[p]);
                        if (frozen) { nh[p] = oh[p]; nl[p] = ol[p]; }   // a converged column keeps its iterate
                    }
                    // tcgen05.st is warp-collective (.sync.aligned): every lane stores, converged columns store what they had
                    tmem_st16(lane_addr + kColVH + ((h * 128 + y0) >> 1), nh);
                    tmem_st16(lane_addr + kColVL + ((h * 128 + y0) >> 1), nl);
                }
                if (part == 1 && h == 1) s_err[col] = err;   // the upper half's share of the column's L1 change
                tmem_st_wait();
                tc_fence_before();
                __syncthreads();
            }
            if (!alive) break;
            // point side: R = V·G, U = mu / R
            if (tid == 0) {
                tc_fence_after();
                const uint32_t idesc = make_idesc((uint32_t)sxp);
                for (int s = 0; s < kBinsT / 16; ++s) {
                    const uint64_t bd = make_sdesc(smem_u32(s_b1) + (uint32_t)s * 2u * ((uint32_t)sxp * 16u), (uint32_t)sxp * 16u, 128u);
                    umma_ts(tmem + kColR, tmem + kColVH + s * 8, bd, idesc, s > 0);
                    umma_ts(tmem + kColR, tmem + kColVL + s * 8, bd, idesc, 1u);
                }
                umma_commit(bar_mma);
            }
            if (!mbar_wait(bar_mma, phase)) { alive = false; if (tid == 0) atomicCAS(&a.stats[2], 0ull, 3ull << 32 | (unsigned long long)i); }
            if (!__syncthreads_and(alive)) { alive = false; break; }
            phase ^= 1u;
            tc_fence_after();
#pragma unroll
            for (int x0 = 0; x0 < kMaxSx; x0 += 16) {
                if (x0 < sxp && part == 0) {
                    uint32_t r[16], hi[8], lo[8];
                    tmem_ld16(lane_addr + kColR + x0, r);
                    tmem_ld_wait();
                    float nu_[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float m = s_mu[x0 + q];
                        nu_[q] = m > 0.0f ? __fdividef(m, __uint_as_float(r[q])) : 0.0f;
                        err += fabsf(nu_[q] - u[x0 + q]);
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) u[x0 + q] = frozen ? u[x0 + q] : nu_[q];
#pragma unroll
                    for (int q = 0; q < 8; ++q) split_pair(u[x0 + 2 * q], u[x0 + 2 * q + 1], hi[q], lo[q]);
                    tmem_st8(lane_addr + kColUH + (x0 >> 1), hi);   // warp-collective: every lane stores
                    tmem_st8(lane_addr + kColUL + (x0 >> 1), lo);
                }
            }
            tmem_st_wait();
            if (part == 0) {
                if (!frozen) ++n_iter;
                err += s_err[col];
                frozen = frozen || !(err >= a.tolerance);   // sinkhorn.rs:85-94: stop once the L1 change of both sides is below the tolerance (NaN stops too)
                s_frozen[col] = frozen;
            }
            tc_fence_before();
            __syncthreads();
            frozen = s_frozen[col] != 0;
            if (__syncthreads_and(frozen)) { ++it; break; }
        }
        if (!alive) break;
        // cost read-out (sinkhorn.rs:131-139): W[j, x] = sum_y V[j, y] (G∘C)[y, x];  cost = sum_x u_x W_x
        for (int t = tid; t < sxp * 32; t += kThreads) {
            const int n = t >> 5, c = t & 31;
            uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
            if (n < sx) {
                vh = __ldg(reinterpret_cast<const uint4*>(a.gch + (size_t)s_idx[n] * kBinsT + c * 8));
                vl = __ldg(reinterpret_cast<const uint4*>(a.gcl + (size_t)s_idx[n] * kBinsT + c * 8));
            }
            *reinterpret_cast<uint4*>(s_b1 + ((size_t)c * sxp + n) * 16) = vh;
            *reinterpret_cast<uint4*>(s_b2 + ((size_t)c * sxp + n) * 16) = vl;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = make_idesc((uint32_t)sxp);
            for (int s = 0; s < kBinsT / 16; ++s) {
                const uint32_t off = (uint32_t)s * 2u * ((uint32_t)sxp * 16u);
                const uint64_t bh = make_sdesc(smem_u32(s_b1) + off, (uint32_t)sxp * 16u, 128u), bl = make_sdesc(smem_u32(s_b2) + off, (uint32_t)sxp * 16u, 128u);
                umma_ts(tmem + kColR, tmem + kColVH + s * 8, bh, idesc, s > 0);
                umma_ts(tmem + kColR, tmem + kColVL + s * 8, bh, idesc, 1u);
                umma_ts(tmem + kColR, tmem + kColVH + s * 8, bl, idesc, 1u);
            }
            umma_commit(bar_mma);
        }
        if (!mbar_wait(bar_mma, phase)) { alive = false; if (tid == 0) atomicCAS(&a.stats[2], 0ull, 4ull << 32 | (unsigned long long)i); }
        if (!__syncthreads_and(alive)) break;
        phase ^= 1u;
        tc_fence_after();
        float cost = 0.0f;
#pragma unroll
        for (int x0 = 0; x0 < kMaxSx; x0 += 16) {
            if (x0 < sxp && part == 0) {
                uint32_t r[16];
                tmem_ld16(lane_addr + kColR + x0, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; ++q) cost += u[x0 + q] * __uint_as_float(r[q]);
            }
        }
        if (j < a.k && part == 0) {
            float d = cost - 0.5f * self_c - 0.5f * a.p_self[i];
            d = d > 0.0f ? d : 0.0f;                      // NaN or an empty centroid → 0: always re-evaluated exactly
            a.approx[(size_t)i * a.k + j] = real ? d : 0.0f;
        }
        ++n_prob;
        tc_fence_before();
        __syncthreads();   // the next point rewrites B1 / B2, the index scratch and the accumulators
    }
    if (tid == 0 && a.stats) { atomicAdd(&a.stats[0], n_prob); }
    if (a.stats && n_iter) atomicAdd(&a.stats[1], n_iter);
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace skt
}  // namespace rbp
