// common.cuh — status plumbing, launch accounting and the Philox RNG contract (include/rbp.h) for librbp_b200.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/rbp.h"

namespace rbp {

extern std::atomic<uint64_t> g_launches;
void set_last_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define RBP_CUDA(expr)                                                           \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) return ::rbp::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define RBP_LAUNCHED()                                   \
    do {                                                 \
        ::rbp::g_launches.fetch_add(1, std::memory_order_relaxed); \
        RBP_CUDA(cudaGetLastError());                    \
    } while (0)

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw — SC'11), the counter-based generator of the RNG contract.
struct Philox4 {
    uint32_t r[4];
};
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{{c0, c1, c2, c3}};
}
enum : uint32_t { TAG_NODE = 0, TAG_ROOT = 1, TAG_COIN = 2, TAG_KMEANSPP = 3 };
__host__ __device__ __forceinline__ uint32_t draw_range(uint32_t r, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __umulhi(r, n);
#else
    return (uint32_t)(((uint64_t)r * n) >> 32);
#endif
}
__host__ __device__ __forceinline__ float draw_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// (shared by the ordered folds of mccfr.cu and nlhe.cu; `rbp_selftest_div_by_count` checks it on the device)
// a / b for b = (float)(visits + 1), with the reciprocal prepared off the dependent chain.
// rb = RN(1/b); q = RN(a*rb); r = a - b*q (exact, FMA); q' = RN(q + r*rb) is the correctly rounded quotient
// (Markstein's theorem) whenever no intermediate under/overflows and b's significand is not all ones — both
// guarded, falling back to IEEE division.  tests/test_mccfr_gpu.py checks it against `/` exhaustively in b.
__device__ __forceinline__ float div_by_count(float a, float b, float rb) {
    // exponent of a within [2^-64, 2^63] (zero takes the slow path too), evaluated beside the FMA chain
    const bool safe = ((((__float_as_uint(a) >> 23 & 0xFFu) - 63u) < 128u) | (__float_as_uint(a) == 0u)) & ((__float_as_uint(b) & 0x7FFFFFu) != 0x7FFFFFu);
    const float q = a * rb;
    const float r = __fmaf_rn(-b, q, a);
    const float fast = __fmaf_rn(r, rb, q);
    if (__builtin_expect(!safe, 0)) return a / b;
    return fast;
}

constexpr float kEps = 1.17549435e-38f;  // pokerkit/src/lib.rs:204 EPSILON = f32::MIN_POSITIVE

}  // namespace rbp
