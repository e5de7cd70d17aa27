// sinkhorn.cuh — warp-cooperative entropic OT (`Sinkhorn`, crates/lloyd/src/sinkhorn.rs:22-139,201-217) for sm_100a.
//
// One warp solves one (mu, nu) problem.  The reference's arithmetic is reproduced operation for operation: log-domain
// Gauss-Seidel sweeps, every Σ sequential in ascending-bucket order, MIN_POSITIVE clamps, the L1 stopping rule on
// exp(potential), the single-accumulator row-major cost.  Parallelism comes only from independent rows: a lane owns
// one x (resp. one y) and runs its softmin sum sequentially; the order-sensitive scalar sums (`delta`, `cost`) are
// fed lane by lane through shuffles.  exp/ln follow the contract of include/rbp.h (`exp_c` / `ln_c`: fixed IEEE
// operation sequences, no contraction — this file is compiled -fmad=false), so results are bit-identical to the oracle.
#pragma once
#include "common.cuh"

namespace rbp {

constexpr int kSkMaxSupport = 256;  // KMEANS_MAX_CLUSTER_COUNT (crates/pokerkit/src/lib.rs:185)

__device__ __forceinline__ float exp_c(float x) {
    if (!(x < 88.72283f)) return x != x ? x : INFINITY;
    if (x < -103.0f) return 0.0f;
    const float kf = rintf(x * 1.44269504088896341f);
    int k = (int)kf;
    float r = __fmaf_rn(kf, -0.693359375f, x);
    r = __fmaf_rn(kf, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = p * r + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    float y = p * (r * r) + r + 1.0f;
    if (k < -125) { y = y * __uint_as_float((uint32_t)(127 - 100) << 23); k += 100; }
    if (k > 127) { y = y * __uint_as_float((uint32_t)(127 + 100) << 23); k -= 100; }
    return y * __uint_as_float((uint32_t)(k + 127) << 23);
}
__device__ __forceinline__ float ln_c(float x) {
    if (!(x > 0.0f)) return x == 0.0f ? -INFINITY : NAN;
    if (x == INFINITY) return x;
    uint32_t u = __float_as_uint(x);
    int e = (int)(u >> 23) - 126;
    if ((u >> 23) == 0) { x = x * 16777216.0f; u = __float_as_uint(x); e = (int)(u >> 23) - 126 - 24; }
    float m = __uint_as_float((u & 0x007FFFFFu) | 0x3F000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    const float z = m * m;
    float y = 7.0376836292e-2f;
    y = y * m + -1.1514610310e-1f;
    y = y * m + 1.1676998740e-1f;
    y = y * m + -1.2420140846e-1f;
    y = y * m + 1.4249322787e-1f;
    y = y * m + -1.6668057665e-1f;
    y = y * m + 2.0000714765e-1f;
    y = y * m + -2.4999993993e-1f;
    y = y * m + 3.3333331174e-1f;
    y = y * m * z;
    const float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y + -0.5f * z;
    float r = m + y;
    r = r + 0.693359375f * fe;
    return r;
}

struct SkParams {
    float temperature;  // 0.025
    int iterations;     // 128
    float tolerance;    // 5e-4
};

// per-warp scratch in shared memory
struct SkWarp {
    float lhs[kSkMaxSupport], rhs[kSkMaxSupport], nxt[kSkMaxSupport];
    float lnmu[kSkMaxSupport], lnnu[kSkMaxSupport];
    uint16_t ix[kSkMaxSupport], iy[kSkMaxSupport];
    int nx, ny;
};

// ground metric: `tri` in Pair::merge order; `reg` = tri / temperature precomputed element-wise (same f32 quotient
// the reference forms on every access, sinkhorn.rs:127-129)
__device__ __forceinline__ float tri_at(const float* __restrict__ t, int x, int y) {
    if (x == y) return 0.0f;
    const int lo = x < y ? x : y, hi = x < y ? y : x;
    return __ldg(t + (size_t)hi * (hi - 1) / 2 + lo);
}

// Load a dense histogram (count accessor) into a sparse side of the warp scratch: ascending support, ln(density).
template <class CountAt>
__device__ __forceinline__ int sk_load_side(CountAt cnt, float weight, int bins, uint16_t* idx, float* lnd, int lane) {
    int n = 0;
    for (int b0 = 0; b0 < bins; b0 += 32) {
        const int b = b0 + lane;
        const float c = b < bins ? cnt(b) : 0.0f;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, c > 0.0f);
        if (c > 0.0f) {
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            idx[pos] = (uint16_t)b;
            lnd[pos] = ln_c(c / weight);  // bins.rs:58-60 density, then `.ln()` (sinkhorn.rs:113)
        }
        n += __popc(m);
    }
    __syncwarp();
    return n;
}

// ordered sum of one value per lane (lanes 0..m-1), continuing `acc`
__device__ __forceinline__ float ordered_add(float acc, float v, int m) {
    for (int k = 0; k < m; ++k) acc = acc + __shfl_sync(0xFFFFFFFFu, v, k);
    return acc;
}

// OT cost of the problem currently loaded in `w` (both sides filled by sk_load_side)
__device__ __forceinline__ float sk_solve(SkWarp& w, const float* __restrict__ tri, const float* __restrict__ reg, const SkParams hp, int lane) {
    const int nx = w.nx, ny = w.ny;
    const float lx0 = ln_c(1.0f / (float)nx), ly0 = ln_c(1.0f / (float)ny);  // Phi::uniform (phi.rs:25-30)
    for (int i = lane; i < nx; i += 32) w.lhs[i] = lx0;
    for (int j = lane; j < ny; j += 32) w.rhs[j] = ly0;
    __syncwarp();
    for (int t = 0; t < hp.iterations; ++t) {
        // lhs(): softmin over y for every x
        for (int i0 = 0; i0 < nx; i0 += 32) {
            const int i = i0 + lane;
            if (i < nx) {
                const int xi = w.ix[i];
                float s = 0.0f;
                for (int j = 0; j < ny; ++j) {
                    const float e = exp_c(w.rhs[j] - tri_at(reg, xi, w.iy[j]));
                    s = s + (e > kEps ? e : kEps);
                }
                w.nxt[i] = w.lnmu[i] - ln_c(s);
            }
        }
        __syncwarp();
        float lerr = 0.0f;  // delta(prev, next) — sequential over the support
        for (int i0 = 0; i0 < nx; i0 += 32) {
            const int i = i0 + lane;
            const float v = i < nx ? fabsf(exp_c(w.nxt[i]) - exp_c(w.lhs[i])) : 0.0f;
            lerr = ordered_add(lerr, v, min(32, nx - i0));
        }
        for (int i = lane; i < nx; i += 32) w.lhs[i] = w.nxt[i];
        __syncwarp();
        // rhs(): softmin over x for every y, against the NEW lhs
        for (int j0 = 0; j0 < ny; j0 += 32) {
            const int j = j0 + lane;
            if (j < ny) {
                const int yj = w.iy[j];
                float s = 0.0f;
                for (int i = 0; i < nx; ++i) {
                    const float e = exp_c(w.lhs[i] - tri_at(reg, w.ix[i], yj));
                    s = s + (e > kEps ? e : kEps);
                }
                w.nxt[j] = w.lnnu[j] - ln_c(s);
            }
        }
        __syncwarp();
        float rerr = 0.0f;
        for (int j0 = 0; j0 < ny; j0 += 32) {
            const int j = j0 + lane;
            const float v = j < ny ? fabsf(exp_c(w.nxt[j]) - exp_c(w.rhs[j])) : 0.0f;
            rerr = ordered_add(rerr, v, min(32, ny - j0));
        }
        for (int j = lane; j < ny; j += 32) w.rhs[j] = w.nxt[j];
        __syncwarp();
        if (lerr + rerr < hp.tolerance) break;
    }
    // Coupling::cost: Σ_x Σ_y exp(lhs_x + rhs_y − reg)·C_xy, one accumulator, row-major
    float cost = 0.0f;
    for (int i = 0; i < nx; ++i) {
        const int xi = w.ix[i];
        const float li = w.lhs[i];
        for (int j0 = 0; j0 < ny; j0 += 32) {
            const int j = j0 + lane;
            float v = 0.0f;
            if (j < ny) {
                const int yj = w.iy[j];
                v = exp_c(li + w.rhs[j] - tri_at(reg, xi, yj)) * tri_at(tri, xi, yj);
            }
            cost = ordered_add(cost, v, min(32, ny - j0));
        }
    }
    return cost;
}

__device__ __forceinline__ float sk_divergence(float xy, float xx, float yy) {  // sinkhorn.rs:166-171
    const float d = xy - 0.5f * xx - 0.5f * yy;
    return d > 0.0f ? d : 0.0f;
}

}  // namespace rbp
