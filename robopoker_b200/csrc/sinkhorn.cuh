// sinkhorn.cuh — warp-cooperative entropic OT (`Sinkhorn`, crates/lloyd/src/sinkhorn.rs:22-139,201-217) for sm_100a.
//
// One warp solves one (mu, nu) problem.  The reference's arithmetic is reproduced operation for operation: log-domain
// Gauss-Seidel sweeps, every Σ sequential in ascending-bucket order, MIN_POSITIVE clamps, the L1 stopping rule on
// exp(potential), the single-accumulator row-major cost.  Only the ADDITIONS of a softmin are order-sensitive; its
// exp terms are independent, so a half-sweep picks, per problem shape, the cheaper of two bit-identical schedules:
//   rows:       a lane owns a target and runs its sum over the sources in a register (full lanes when the target
//               side is wide — centroids);
//   transposed: lanes span the SOURCES, the exp terms of one 32-source slab go through a [target][source] shared
//               tile, and lane t then adds row t left to right (narrow target side — a point's ~11 buckets — where
//               `rows` would leave two thirds of the warp idle in front of the 256-term loop).
// Either way a lane has up to 8 independent exp chains in flight.  Scratch is sized per problem class (a point side
// holds <= 64 buckets): 5.2 KB per warp for point problems, so 24-32 warps per SM leave the L1 large enough to keep the
// table rows in use resident (the sweeps re-read the same nx x ny entries ~230 times per solve).
// The ground metric is a dense symmetric [256][256] table (L1/L2 resident): the broadcast index picks the row, the
// lane's index the column, so every load is one coalesced line.  exp/ln follow the contract of include/rbp.h
// (`exp_c` / `ln_c`: fixed IEEE operation sequences with explicit fma — this file is compiled -fmad=false so nothing
// else contracts), so results are bit-identical to the oracle.
#pragma once
#include "common.cuh"

namespace rbp {

constexpr int kSkMaxSupport = 256;  // KMEANS_MAX_CLUSTER_COUNT (crates/pokerkit/src/lib.rs:185)
constexpr int kSkLd = 256;          // row stride of the dense ground-metric tables
constexpr int kSkTrMax = 16;        // widest target side the transposed schedule takes
constexpr int kSkTile = 36;         // row stride of its [target][32 sources] tile (16-byte rows, conflict-free LDS.128)

// exp contract (include/rbp.h): saturating, x clamped to [ln MIN_POSITIVE, ln MAX]; k = rint(x·log2e) through the
// 1.5·2^23 shifter; r = x − k·ln2 (two-term split); degree-5 Cephes polynomial in Horner form, all fma; 2^k applied
// by adding k to the exponent field.
__device__ __forceinline__ float exp_c(float x) {
    x = fminf(fmaxf(x, -87.33654f), 88.72283f);
    const float t = __fmaf_rn(x, 1.44269504f, 12582912.0f);
    const float kf = t - 12582912.0f;
    float r = __fmaf_rn(kf, -0.693359375f, x);
    r = __fmaf_rn(kf, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float y = __fmaf_rn(p, r * r, r) + 1.0f;
    return __uint_as_float(__float_as_uint(y) + (__float_as_uint(t) << 23));
}
// exp_c for callers that guarantee x <= 88.72283 (every softmin / delta argument: potentials are <= -ln MIN_POSITIVE
// and the ground metric is non-negative) — the upper clamp is then the identity
__device__ __forceinline__ float exp_c_le(float x) {
    x = fmaxf(x, -87.33654f);
    const float t = __fmaf_rn(x, 1.44269504f, 12582912.0f);
    const float kf = t - 12582912.0f;
    float r = __fmaf_rn(kf, -0.693359375f, x);
    r = __fmaf_rn(kf, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float y = __fmaf_rn(p, r * r, r) + 1.0f;
    return __uint_as_float(__float_as_uint(y) + (__float_as_uint(t) << 23));
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
    y = __fmaf_rn(y, m, -1.1514610310e-1f);
    y = __fmaf_rn(y, m, 1.1676998740e-1f);
    y = __fmaf_rn(y, m, -1.2420140846e-1f);
    y = __fmaf_rn(y, m, 1.4249322787e-1f);
    y = __fmaf_rn(y, m, -1.6668057665e-1f);
    y = __fmaf_rn(y, m, 2.0000714765e-1f);
    y = __fmaf_rn(y, m, -2.4999993993e-1f);
    y = __fmaf_rn(y, m, 3.3333331174e-1f);
    y = y * m * z;
    const float fe = (float)e;
    y = __fmaf_rn(-2.12194440e-4f, fe, y);
    y = __fmaf_rn(-0.5f, z, y);
    float r = m + y;
    r = __fmaf_rn(0.693359375f, fe, r);
    return r;
}

struct SkParams {
    float temperature;  // 0.025
    int iterations;     // 128
    float tolerance;    // 5e-4
};

// One side of an OT problem as it sits in the warp scratch: potential, ln(density), ascending support
struct SkSide {
    float* pot;
    const float* lnd;
    const uint8_t* idx;
    int n;
};
// per-warp scratch in shared memory: slot A holds <= NA buckets (points: 64), slot B <= NB (centroids: 256).
// <64,256> = 5200 B, <256,256> = 6928 B.
template <int NA, int NB>
struct __align__(16) SkScratch {
    float pot_a[NA], lnd_a[NA], pot_b[NB], lnd_b[NB];
    float tile[kSkTrMax * kSkTile];
    uint8_t idx_a[NA], idx_b[NB];  // bucket ids < 256
    int n_a, n_b;
    __device__ __forceinline__ SkSide a() { return SkSide{pot_a, lnd_a, idx_a, n_a}; }
    __device__ __forceinline__ SkSide b() { return SkSide{pot_b, lnd_b, idx_b, n_b}; }
};
using SkPoint = SkScratch<64, kSkMaxSupport>;             // point x centroid, point x point
using SkWide = SkScratch<kSkMaxSupport, kSkMaxSupport>;   // centroid x centroid, generic batches

// Dense symmetric ground-metric tables from `tri` in Pair::merge order (pair.rs:36-39): metric[x][y] = raw_distance
// (0 on the diagonal, metric.rs:42-54); reg = metric / temperature element-wise — the same f32 quotient the reference
// forms on every access (sinkhorn.rs:127-129).  Host side.
inline bool sk_metric_valid(const float* tri, int bins) {  // distances: finite, >= 0 (exp_c_le's domain argument relies on it)
    for (size_t t = 0; t < (size_t)bins * (bins - 1) / 2; ++t)
        if (!(tri[t] >= 0.0f) || !(tri[t] <= 3.0e38f)) return false;
    return true;
}
inline void sk_dense_tables(const float* tri, int bins, float temperature, float* metric, float* reg) {
    for (size_t i = 0; i < (size_t)kSkLd * kSkLd; ++i) { metric[i] = 0.0f; reg[i] = 0.0f; }
    for (int hi = 1; hi < bins; ++hi)
        for (int lo = 0; lo < hi; ++lo) {
            const float c = tri[(size_t)hi * (hi - 1) / 2 + lo], q = c / temperature;
            metric[(size_t)hi * kSkLd + lo] = c; metric[(size_t)lo * kSkLd + hi] = c;
            reg[(size_t)hi * kSkLd + lo] = q; reg[(size_t)lo * kSkLd + hi] = q;
        }
}

// Load a dense histogram (count accessor) into a sparse side of the warp scratch: ascending support, ln(density).
template <class CountAt>
__device__ __forceinline__ int sk_load_side(CountAt cnt, float weight, int bins, uint8_t* idx, float* lnd, int lane) {
    int n = 0;
    for (int b0 = 0; b0 < bins; b0 += 32) {
        const int b = b0 + lane;
        const float c = b < bins ? cnt(b) : 0.0f;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, c > 0.0f);
        if (c > 0.0f) {
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            idx[pos] = (uint8_t)b;
            lnd[pos] = ln_c(c / weight);  // bins.rs:58-60 density, then `.ln()` (sinkhorn.rs:113)
        }
        n += __popc(m);
    }
    __syncwarp();
    return n;
}

// acc + v(lane 0) + v(lane 1) + … + v(lane 31), left to right, on every lane.  Lanes outside the support pass +0.0,
// which is an exact no-op on a sum that starts at +0.0.  `buf`: 32 floats of warp scratch, 16-byte aligned.
__device__ __forceinline__ float ordered_sum32(float acc, float v, float* buf, int lane) {
    buf[lane] = v;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 w = reinterpret_cast<const float4*>(buf)[q];
        acc = acc + w.x; acc = acc + w.y; acc = acc + w.z; acc = acc + w.w;
    }
    __syncwarp();
    return acc;
}

// softmin term max(exp(p − c), MIN_POSITIVE)  (sinkhorn.rs:118-120).  The saturating exp_c never returns less than
// MIN_POSITIVE (its lower clamp is ln MIN_POSITIVE rounded up; checked exhaustively by
// tests/test_oracle_sinkhorn.py::test_saturating_exp_never_returns_less_than_min_positive), so the max is the identity
// and is not issued.
__device__ __forceinline__ float sk_term(float p, float c) { return exp_c_le(p - c); }

// s + Σ_{i<8} term(pot[i], col[idx[i]·ld]) left to right; MASKED: only the first m (< 8) sources exist — the other
// slots read stale scratch and contribute an exact +0.0
template <bool MASKED>
__device__ __forceinline__ float sk_rows8(float s, const float* __restrict__ reg, unsigned col, const uint8_t* __restrict__ idx, const float* __restrict__ pot, int m) {
    const uint2 r8 = *reinterpret_cast<const uint2*>(idx);
    const float4 pa = *reinterpret_cast<const float4*>(pot), pb = *reinterpret_cast<const float4*>(pot + 4);
    const float c0 = __ldg(reg + (col + ((unsigned)(r8.x & 0xFFu) << 8))), c1 = __ldg(reg + (col + ((unsigned)((r8.x >> 8) & 0xFFu) << 8)));
    const float c2 = __ldg(reg + (col + ((unsigned)((r8.x >> 16) & 0xFFu) << 8))), c3 = __ldg(reg + (col + ((unsigned)(r8.x >> 24) << 8)));
    const float c4 = __ldg(reg + (col + ((unsigned)(r8.y & 0xFFu) << 8))), c5 = __ldg(reg + (col + ((unsigned)((r8.y >> 8) & 0xFFu) << 8)));
    const float c6 = __ldg(reg + (col + ((unsigned)((r8.y >> 16) & 0xFFu) << 8))), c7 = __ldg(reg + (col + ((unsigned)(r8.y >> 24) << 8)));
    float e0 = sk_term(pa.x, c0), e1 = sk_term(pa.y, c1), e2 = sk_term(pa.z, c2), e3 = sk_term(pa.w, c3);
    float e4 = sk_term(pb.x, c4), e5 = sk_term(pb.y, c5), e6 = sk_term(pb.z, c6), e7 = sk_term(pb.w, c7);
    if (MASKED) {  // 5 <= m <= 7
        e5 = m > 5 ? e5 : 0.0f; e6 = m > 6 ? e6 : 0.0f; e7 = 0.0f;
    }
    s = s + e0; s = s + e1; s = s + e2; s = s + e3; s = s + e4; s = s + e5; s = s + e6; s = s + e7;
    return s;
}
// the same for 1 <= m <= 4 sources
__device__ __forceinline__ float sk_rows4(float s, const float* __restrict__ reg, unsigned col, const uint8_t* __restrict__ idx, const float* __restrict__ pot, int m) {
    const uchar4 r4 = *reinterpret_cast<const uchar4*>(idx);
    const float4 pa = *reinterpret_cast<const float4*>(pot);
    const float c0 = __ldg(reg + (col + ((unsigned)r4.x << 8))), c1 = __ldg(reg + (col + ((unsigned)r4.y << 8)));
    const float c2 = __ldg(reg + (col + ((unsigned)r4.z << 8))), c3 = __ldg(reg + (col + ((unsigned)r4.w << 8)));
    const float e0 = sk_term(pa.x, c0);
    float e1 = sk_term(pa.y, c1), e2 = sk_term(pa.z, c2), e3 = sk_term(pa.w, c3);
    e1 = m > 1 ? e1 : 0.0f; e2 = m > 2 ? e2 : 0.0f; e3 = m > 3 ? e3 : 0.0f;
    s = s + e0; s = s + e1; s = s + e2; s = s + e3;
    return s;
}

// One Gauss-Seidel half-sweep (sinkhorn.rs:96-129 `lhs()` / `rhs()` + `delta`): for every target t
//   next_t = ln dens_t − ln Σ_s max(exp(pot_s − reg[t][s]), MIN_POSITIVE)      (Σ sequential over the source support)
// replaces pot_t and returns Σ_t |exp(next_t) − exp(prev_t)| (sequential over the target support).
__device__ __forceinline__ float sk_half_sweep(const SkSide tg, const SkSide sr, float* __restrict__ tile, const float* __restrict__ reg, int lane) {
    float* __restrict__ pot_t = tg.pot;
    const float* __restrict__ lnd_t = tg.lnd;
    const uint8_t* __restrict__ idx_t = tg.idx;
    const float* __restrict__ pot_s = sr.pot;
    const uint8_t* __restrict__ idx_s = sr.idx;
    const int n_t = tg.n, n_s = sr.n;
    float err = 0.0f;
    // issue-slot model: rows = ceil(n_t/32)·n_s exp steps; transposed = ceil4(n_t)·ceil(n_s/32) exp steps + the row adds
    const int n_t4 = (n_t + 3) & ~3;
    const int cost_rows = ((n_t + 31) >> 5) * n_s;
    const int cost_tr = n_t4 * ((n_s + 31) >> 5) + (n_s >> 4) + 2;
    if (n_t <= kSkTrMax && cost_tr < cost_rows) {
        float s = 0.0f;  // lane t: running sum of target t
        for (int s0 = 0; s0 < n_s; s0 += 32) {
            const int si = s0 + lane;
            const bool live = si < n_s;
            const float ps = live ? pot_s[si] : 0.0f;
            const unsigned col = live ? (unsigned)idx_s[si] : 0u;
            // rows n_t..n_t4-1 of the tile are written from stale bucket ids (valid table rows) and never read
            int t = 0;
            for (; t + 8 <= n_t4; t += 8) {
                const uint2 r8 = *reinterpret_cast<const uint2*>(idx_t + t);
                const float c0 = __ldg(reg + (col + ((unsigned)(r8.x & 0xFFu) << 8))), c1 = __ldg(reg + (col + ((unsigned)((r8.x >> 8) & 0xFFu) << 8)));
                const float c2 = __ldg(reg + (col + ((unsigned)((r8.x >> 16) & 0xFFu) << 8))), c3 = __ldg(reg + (col + ((unsigned)(r8.x >> 24) << 8)));
                const float c4 = __ldg(reg + (col + ((unsigned)(r8.y & 0xFFu) << 8))), c5 = __ldg(reg + (col + ((unsigned)((r8.y >> 8) & 0xFFu) << 8)));
                const float c6 = __ldg(reg + (col + ((unsigned)((r8.y >> 16) & 0xFFu) << 8))), c7 = __ldg(reg + (col + ((unsigned)(r8.y >> 24) << 8)));
                const float e0 = sk_term(ps, c0), e1 = sk_term(ps, c1), e2 = sk_term(ps, c2), e3 = sk_term(ps, c3);
                const float e4 = sk_term(ps, c4), e5 = sk_term(ps, c5), e6 = sk_term(ps, c6), e7 = sk_term(ps, c7);
                float* __restrict__ w = tile + t * kSkTile + lane;
                w[0] = live ? e0 : 0.0f; w[kSkTile] = live ? e1 : 0.0f; w[2 * kSkTile] = live ? e2 : 0.0f; w[3 * kSkTile] = live ? e3 : 0.0f;
                w[4 * kSkTile] = live ? e4 : 0.0f; w[5 * kSkTile] = live ? e5 : 0.0f; w[6 * kSkTile] = live ? e6 : 0.0f; w[7 * kSkTile] = live ? e7 : 0.0f;
            }
            if (t < n_t4) {
                const uchar4 r4 = *reinterpret_cast<const uchar4*>(idx_t + t);
                const float c0 = __ldg(reg + (col + ((unsigned)r4.x << 8))), c1 = __ldg(reg + (col + ((unsigned)r4.y << 8)));
                const float c2 = __ldg(reg + (col + ((unsigned)r4.z << 8))), c3 = __ldg(reg + (col + ((unsigned)r4.w << 8)));
                const float e0 = sk_term(ps, c0), e1 = sk_term(ps, c1), e2 = sk_term(ps, c2), e3 = sk_term(ps, c3);
                float* __restrict__ w = tile + t * kSkTile + lane;
                w[0] = live ? e0 : 0.0f; w[kSkTile] = live ? e1 : 0.0f; w[2 * kSkTile] = live ? e2 : 0.0f; w[3 * kSkTile] = live ? e3 : 0.0f;
            }
            __syncwarp();
            if (lane < n_t) {  // dead sources hold +0.0: an exact no-op on a non-negative sum
                const float4* __restrict__ row = reinterpret_cast<const float4*>(tile + lane * kSkTile);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 w = row[q];
                    s = s + w.x; s = s + w.y; s = s + w.z; s = s + w.w;
                }
            }
            __syncwarp();
        }
        float v = 0.0f;
        if (lane < n_t) {
            const float prev = pot_t[lane];
            const float nv = lnd_t[lane] - ln_c(s);
            v = fabsf(exp_c_le(nv) - exp_c_le(prev));
            pot_t[lane] = nv;
        }
        err = ordered_sum32(err, v, tile, lane);
    } else {
        const int n_s8 = n_s & ~7, tail = n_s - n_s8;
        for (int t0 = 0; t0 < n_t; t0 += 32) {
            const int t = t0 + lane;
            const bool live = t < n_t;
            const unsigned col = live ? (unsigned)idx_t[t] : 0u;
            float s = 0.0f;
            for (int q = 0; q < n_s8; q += 8) s = sk_rows8<false>(s, reg, col, idx_s + q, pot_s + q, 8);
            if (tail > 4) s = sk_rows8<true>(s, reg, col, idx_s + n_s8, pot_s + n_s8, tail);
            else if (tail > 0) s = sk_rows4(s, reg, col, idx_s + n_s8, pot_s + n_s8, tail);
            float v = 0.0f;
            if (live) {
                const float prev = pot_t[t];
                const float nv = lnd_t[t] - ln_c(s);
                v = fabsf(exp_c_le(nv) - exp_c_le(prev));
                pot_t[t] = nv;
            }
            err = ordered_sum32(err, v, tile, lane);
        }
    }
    return err;
}

// OT cost of (mu, nu) — two sides of the warp scratch already filled with support and ln(density); `metric` / `reg`
// are the dense tables of sk_dense_tables, `tile` the scratch's tile.  `stats` (nullable): {solves, sweeps, exp terms}.
__device__ __forceinline__ float sk_solve(const SkSide mu, const SkSide nu, float* __restrict__ tile, const float* __restrict__ metric,
                                          const float* __restrict__ reg, const SkParams hp, int lane,
                                          unsigned long long* __restrict__ stats = nullptr) {
    const int nx = mu.n, ny = nu.n;
    const float lx0 = ln_c(1.0f / (float)nx), ly0 = ln_c(1.0f / (float)ny);  // Phi::uniform (phi.rs:25-30)
    for (int i = lane; i < nx; i += 32) mu.pot[i] = lx0;
    for (int j = lane; j < ny; j += 32) nu.pot[j] = ly0;
    __syncwarp();
    int sweeps = 0;
    for (int t = 0; t < hp.iterations; ++t) {
        const float lerr = sk_half_sweep(mu, nu, tile, reg, lane);
        const float rerr = sk_half_sweep(nu, mu, tile, reg, lane);  // against the NEW lhs
        ++sweeps;
        if (lerr + rerr < hp.tolerance) break;
    }
    if (stats && lane == 0) {
        atomicAdd(stats + 0, 1ull);
        atomicAdd(stats + 1, (unsigned long long)sweeps);
        atomicAdd(stats + 2, (unsigned long long)nx * ny * (2ull * sweeps + 1ull));
    }
    // Coupling::cost: Σ_x Σ_y exp(lhs_x + rhs_y − reg)·C_xy, one accumulator, row-major
    float cost = 0.0f;
    for (int i = 0; i < nx; ++i) {
        const int row = (int)mu.idx[i] * kSkLd;
        const float li = mu.pot[i];
        for (int j0 = 0; j0 < ny; j0 += 32) {
            const int j = j0 + lane;
            float v = 0.0f;
            if (j < ny) {
                const int at = row + (int)nu.idx[j];
                v = exp_c(li + nu.pot[j] - __ldg(reg + at)) * __ldg(metric + at);
            }
            cost = ordered_sum32(cost, v, tile, lane);
        }
    }
    return cost;
}

__device__ __forceinline__ float sk_divergence(float xy, float xx, float yy) {  // sinkhorn.rs:166-171
    const float d = xy - 0.5f * xx - 0.5f * yy;
    return d > 0.0f ? d : 0.0f;
}

}  // namespace rbp
