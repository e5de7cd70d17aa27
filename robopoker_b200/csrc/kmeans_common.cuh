// kmeans_common.cuh — pieces shared by the k-means layers: the opaque handle base and the k-means++ integer draw.
#pragma once
#include "common.cuh"

struct rbp_comm;
struct rbp_kmeans {  // opaque `rbp_kmeans_t`; concrete layers derive from it
    int kind = 0;
    int k = 0;                  // clusters (for the fused exchange)
    rbp_comm* comm = nullptr;   // rbp_kmeans_attach_comm: points are sharded over its ranks
};

namespace rbp {

// k-means++ weight contract (include/rbp.h): q = (u64)(min(potential, 2^20) * 2^32)
__device__ __forceinline__ unsigned long long quantize_potential(float p) {
    const float q = p < 1048576.0f ? p : 1048576.0f;
    return (unsigned long long)((double)q * 4294967296.0);
}

// per-group integer sums of the quantised potentials (group = `group` consecutive points)
static __global__ void __launch_bounds__(128)
pp_blocksum_kernel(const float* __restrict__ pot, int64_t n, unsigned long long* __restrict__ bsum) {
    __shared__ unsigned long long s_sum[4];
    const int64_t i = blockIdx.x * (int64_t)128 + threadIdx.x;
    unsigned long long q = i < n ? quantize_potential(pot[i]) : 0ull;
    for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, d);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) bsum[blockIdx.x] = s_sum[0] + s_sum[1] + s_sum[2] + s_sum[3];
}

// draw: x = mulhi64(word, T); pick = first i with x < Σ_{k<=i} q_k   (one block; integer sums ⇒ order-free exact)
static __global__ void __launch_bounds__(1024)
pp_pick_kernel(int64_t n, int group, const float* __restrict__ pot, const unsigned long long* __restrict__ bsum, int nb, uint32_t w0,
               uint32_t w1, int64_t* __restrict__ pick, int round, int32_t* __restrict__ chosen) {
    __shared__ unsigned long long s_part[1024];
    __shared__ unsigned long long s_x, s_base;
    __shared__ int s_blk;
    __shared__ bool s_zero;
    const int tid = threadIdx.x;
    const int per = (nb + 1023) / 1024;
    const int lo = min(nb, tid * per), hi = min(nb, lo + per);
    unsigned long long mine = 0;
    for (int b = lo; b < hi; ++b) mine += bsum[b];
    s_part[tid] = mine;
    __syncthreads();
    if (tid == 0) {
        unsigned long long T = 0;
        for (int t = 0; t < 1024; ++t) T += s_part[t];
        const unsigned long long word = (unsigned long long)w0 << 32 | w1;
        const unsigned long long x = __umul64hi(word, T);
        s_x = x;
        s_zero = T == 0ull;
        unsigned long long cum = 0;
        int owner = 1023;
        for (int t = 0; t < 1024; ++t) { if (x < cum + s_part[t]) { owner = t; break; } cum += s_part[t]; }
        const int olo = min(nb, owner * per), ohi = min(nb, olo + per);
        int blk = ohi - 1;
        for (int b = olo; b < ohi; ++b) { if (x < cum + bsum[b]) { blk = b; break; } cum += bsum[b]; }
        s_blk = blk < 0 ? 0 : blk;
        s_base = cum;
        const int64_t start = (int64_t)s_blk * group;
        const int64_t end = min(n, start + group);
        int64_t p = end - 1;
        for (int64_t i = start; i < end; ++i) {
            cum += quantize_potential(pot[i]);
            if (x < cum) { p = i; break; }
        }
        if (s_zero) p = n - 1;
        *pick = p;
        chosen[round] = (int32_t)p;
    }
}

}  // namespace rbp
