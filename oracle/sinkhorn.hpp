// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).  CPU restatement of the reference's entropic optimal
// transport (`crates/lloyd/src/sinkhorn.rs:22-139,166-217`, `phi.rs:20-35`, `metric.rs:42-54`), strict left-to-right
// f32, sequential sums in support (ascending bucket) order.
//
// exp / ln.  The reference calls `f32::exp` / `f32::ln` (platform libm) — "parity unpinned" (SURVEY §8c): no test in
// the reference pins a value beyond 1e-4 properties, and CUDA's expf/logf differ from glibc's in the last ulp, which
// is enough to flip a near-tie bucket assignment.  As with the RNG, the path is therefore defined on a CONTRACT both
// sides implement with identical IEEE operations (explicit fma, nothing else contracted): `exp_c` / `ln_c` below
// (Cephes-style single precision kernels, ≤ 2 ulp from libm; exp_c saturates instead of under/overflowing).
// `Math::Libm` keeps the literal libm restatement so tests can bound the contract's distance from it; the
// reference's own property tests (self-divergence < 1e-4, symmetry < 1e-3 on the
// synthetic metric of sinkhorn.rs:252-262) are checked under both.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

inline float f_from_bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t bits_of(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// exp contract: saturating — x clamped to [ln MIN_POSITIVE, ln MAX] (NaN → lower bound); k = rint(x·log2e) through
// the 1.5·2^23 shifter; r = x − k·ln2 (two-term split); degree-5 Cephes polynomial in Horner form, every step one
// fma; 2^k applied by adding k to the exponent field.  Identical IEEE operations on both sides (fmaf is exact).
inline float exp_c(float x) {
    x = fminf(fmaxf(x, -87.33654f), 88.72283f);
    const float t = fmaf(x, 1.44269504f, 12582912.0f);
    const float kf = t - 12582912.0f;
    float r = fmaf(kf, -0.693359375f, x);
    r = fmaf(kf, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    const float y = fmaf(p, r * r, r) + 1.0f;
    return f_from_bits(bits_of(y) + (bits_of(t) << 23));
}
// ln contract (x > 0): x = m·2^e with m in [sqrt(1/2), sqrt 2); Cephes logf polynomial in (m − 1), Horner steps fma
inline float ln_c(float x) {
    if (!(x > 0.0f)) return x == 0.0f ? -INFINITY : NAN;
    if (x == INFINITY) return x;
    uint32_t u = bits_of(x);
    int e = (int)(u >> 23) - 126;
    if ((u >> 23) == 0) {  // denormal: rescale exactly
        x = x * 16777216.0f; u = bits_of(x); e = (int)(u >> 23) - 126 - 24;
    }
    float m = f_from_bits((u & 0x007FFFFFu) | 0x3F000000u);  // [0.5, 1)
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    const float z = m * m;
    float y = 7.0376836292e-2f;
    y = fmaf(y, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = y * m * z;
    const float fe = (float)e;
    y = fmaf(-2.12194440e-4f, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(0.693359375f, fe, r);
    return r;
}

enum class Math { Contract, Libm };
template <Math M> inline float exp_m(float x) { return M == Math::Contract ? exp_c(x) : expf(x); }
template <Math M> inline float ln_m(float x) { return M == Math::Contract ? ln_c(x) : logf(x); }

struct SinkhornParams {  // lloyd/src/hyperparams/sinkhorn.rs:17-23
    float temperature = 0.025f;
    int iterations = 128;
    float tolerance = 0.0005f;
};

// Ground metric over `bins` abstractions: triangular f32 table in `Pair::merge` order (pair.rs:36-39), 0 on the diagonal
struct GroundMetric {
    int bins = 0;
    std::vector<float> tri;
    float raw(int x, int y) const {  // metric.rs:42-54 raw_distance
        if (x == y) return 0.0f;
        const int lo = x < y ? x : y, hi = x < y ? y : x;
        return tri[(size_t)hi * (hi - 1) / 2 + lo];
    }
};

// A histogram as (support indices ascending, densities = count as f32 / weight as f32)  (bins.rs:58-60,83-87)
struct Measure {
    std::vector<int> idx;
    std::vector<float> dens;
    template <class C>
    static Measure from_counts(const C* counts, int bins) {
        Measure m;
        uint64_t w = 0;
        for (int b = 0; b < bins; ++b) w += (uint64_t)counts[b];
        for (int b = 0; b < bins; ++b)
            if (counts[b] > 0) { m.idx.push_back(b); m.dens.push_back((float)(uint64_t)counts[b] / (float)w); }
        return m;
    }
};

// sinkhorn.rs:22-139 + Coupling::cost (201-217): returns OT cost and the iteration count used
template <Math M>
inline float ot_cost(const Measure& mu, const Measure& nu, const GroundMetric& g, const SinkhornParams& hp, int* iters_out = nullptr) {
    const int nx = (int)mu.idx.size(), ny = (int)nu.idx.size();
    std::vector<float> lhs(nx, ln_m<M>(1.0f / (float)nx)), rhs(ny, ln_m<M>(1.0f / (float)ny));  // phi.rs:25-30 uniform
    std::vector<float> reg((size_t)nx * ny);
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) reg[(size_t)i * ny + j] = g.raw(mu.idx[i], nu.idx[j]) / hp.temperature;  // regularization
    std::vector<float> next_l(nx), next_r(ny);
    int t = 0;
    for (; t < hp.iterations; ++t) {
        for (int i = 0; i < nx; ++i) {  // lhs(): softmin(x, mu, rhs)
            float s = 0.0f;
            for (int j = 0; j < ny; ++j) {
                float e = exp_m<M>(rhs[j] - reg[(size_t)i * ny + j]);
                s = s + (e > FLT_MIN ? e : FLT_MIN);
            }
            next_l[i] = ln_m<M>(mu.dens[i]) - ln_m<M>(s);
        }
        float lerr = 0.0f;
        for (int i = 0; i < nx; ++i) lerr = lerr + fabsf(exp_m<M>(next_l[i]) - exp_m<M>(lhs[i]));  // delta(prev, next)
        lhs.swap(next_l);
        for (int j = 0; j < ny; ++j) {  // rhs(): softmin(y, nu, lhs) — uses the NEW lhs; regularization(y, x) is symmetric
            float s = 0.0f;
            for (int i = 0; i < nx; ++i) {
                float e = exp_m<M>(lhs[i] - reg[(size_t)i * ny + j]);
                s = s + (e > FLT_MIN ? e : FLT_MIN);
            }
            next_r[j] = ln_m<M>(nu.dens[j]) - ln_m<M>(s);
        }
        float rerr = 0.0f;
        for (int j = 0; j < ny; ++j) rerr = rerr + fabsf(exp_m<M>(next_r[j]) - exp_m<M>(rhs[j]));
        rhs.swap(next_r);
        if (lerr + rerr < hp.tolerance) { ++t; break; }
    }
    if (iters_out) *iters_out = t;
    float cost = 0.0f;  // single accumulator, row-major over (x, y)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            cost = cost + exp_m<M>(lhs[i] + rhs[j] - reg[(size_t)i * ny + j]) * g.raw(mu.idx[i], nu.idx[j]);
    return cost;
}

// sinkhorn.rs:166-171 divergence, with the self terms supplied (the reference memoises them per histogram)
inline float divergence_from(float xy, float xx, float yy) {
    float d = xy - 0.5f * xx - 0.5f * yy;
    return d > 0.0f ? d : 0.0f;
}
template <Math M>
inline float divergence(const Measure& mu, const Measure& nu, const GroundMetric& g, const SinkhornParams& hp) {
    return divergence_from(ot_cost<M>(mu, nu, g, hp), ot_cost<M>(mu, mu, g, hp), ot_cost<M>(nu, nu, g, hp));
}

}  // namespace orc
