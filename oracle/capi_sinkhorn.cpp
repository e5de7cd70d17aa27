// ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes-facing C API over oracle/sinkhorn.hpp.
#include <thread>

#include "sinkhorn.hpp"

using namespace orc;

extern "C" {
float orc_exp_c(float x) { return exp_c(x); }
float orc_ln_c(float x) { return ln_c(x); }
// min of exp_c over every float in [lo, hi] (both negative, lo <= hi): the GPU softmin drops the reference's
// `.max(MIN_POSITIVE)` because the saturating contract never returns less — this is the exhaustive check of that claim
float orc_exp_c_min_over(float lo, float hi) {
    float m = INFINITY;
    for (uint32_t u = bits_of(lo); u >= bits_of(hi); --u) { const float e = exp_c(f_from_bits(u)); if (!(e >= m)) m = e; }
    return m;
}
// counts are dense u32[bins]; tri is the triangular ground metric; math: 0 contract, 1 libm
float orc_ot_cost(const uint32_t* mu, const uint32_t* nu, int bins, const float* tri, int math, float temperature, int iterations,
                  float tolerance, int* iters_out) {
    GroundMetric g; g.bins = bins; g.tri.assign(tri, tri + (size_t)bins * (bins - 1) / 2);
    SinkhornParams hp; hp.temperature = temperature; hp.iterations = iterations; hp.tolerance = tolerance;
    Measure a = Measure::from_counts(mu, bins), b = Measure::from_counts(nu, bins);
    return math == 0 ? ot_cost<Math::Contract>(a, b, g, hp, iters_out) : ot_cost<Math::Libm>(a, b, g, hp, iters_out);
}
// batch of divergences: pairs (a[i], b[i]) of dense u32 histograms
void orc_sinkhorn_divergence_batch(const uint32_t* a, const uint32_t* b, int64_t n, int bins, const float* tri, int math, float* out,
                                   int threads) {
    GroundMetric g; g.bins = bins; g.tri.assign(tri, tri + (size_t)bins * (bins - 1) / 2);
    SinkhornParams hp;
    if (threads < 1) threads = 1;
    auto work = [&](int t) {
        for (int64_t i = t; i < n; i += threads) {
            Measure x = Measure::from_counts(a + (size_t)i * bins, bins), y = Measure::from_counts(b + (size_t)i * bins, bins);
            out[i] = math == 0 ? divergence<Math::Contract>(x, y, g, hp) : divergence<Math::Libm>(x, y, g, hp);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}
}
