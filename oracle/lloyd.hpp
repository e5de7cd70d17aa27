// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).  CPU restatement of the reference's hierarchical
// k-means abstraction stage for the TURN layer: points are histograms over the 101 river-equity buckets, the
// distance is the 1-D Wasserstein `Equity::variation`, clustering is Elkan-accelerated Lloyd with integer
// centroid merges.  Strict left-to-right f32 (build -ffp-contract=off).
//
// Follows:
//   crates/elkan/src/elkan.rs:39-47,68-77   init_bounds / neighbor (first minimum)
//   crates/elkan/src/elkan.rs:80-168        pairwises, midpoints, refresh, rebound, recompute, drift, step_elkan
//   crates/elkan/src/bounds.rs:57-91        has_shifted, update, refresh, witness
//   crates/lloyd/src/layer.rs:62-113,140-181  lookup (naive argmin), metric (symmetrised, normalised), k-means++
//   crates/lloyd/src/equity.rs:41-53        variation
//   crates/lloyd/src/bins.rs:58-60,75-82    density = count as f32 / weight as f32; merge = integer add
//   crates/lloyd/src/metric.rs:127-141, pair.rs:36-39   normalise by max, triangular index
//
// Parity status: no golden vectors exist in the reference for centroids/assignments; pinned here
// are the reference's own properties (variation symmetric / zero on self — lloyd/src/emd.rs:73-97; Elkan ≡ naive
// — lloyd/src/tests.rs:149-160; Pair bijection — pair.rs:172-189).  k-means++ seeding is "parity unpinned":
// the reference draws from SmallRng(SipHash(street)) + WeightedIndex<f32>; the contract here is
//   round r: word = Philox(counter=(r,0,0,TAG_KMEANSPP), key=seed); x = mulhi64(r0:r1, T),
//   weights q_i = (u64)(min(potential_i, 2^20) * 2^32), T = Σ q_i, pick = first i with x < Σ_{k<=i} q_k
// (integer prefix sums: associative, so any parallel reduction order reproduces it exactly).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "rng.hpp"
#include "sinkhorn.hpp"

namespace orc {

struct Kmeans {
    int N = 0, K = 0, B = 101;
    std::vector<uint8_t> points;      // [N][B] counts (<= 255 per bin)
    std::vector<uint32_t> pweight;    // [N]
    std::vector<uint64_t> ccounts;    // [K][B] centroid counts (sum of member counts)
    std::vector<uint64_t> cweight;    // [K]
    std::vector<uint32_t> assign;     // bounds.j
    std::vector<float> upper;         // bounds.error
    std::vector<float> lower;         // [N][K]
    std::vector<uint8_t> stale;
    std::vector<uint32_t> prior;      // assignments before the step (Prior::tally)
    int threads = 1;
    uint64_t dist_evals = 0;
    // kind 0: W1 `Equity::variation` (turn layer);  kind 1: `Sinkhorn::divergence` under a ground metric (flop layer,
    // metric.rs:109-115).  For kind 1 the self terms OT(h,h) are memoised per histogram as the reference does
    // (sinkhorn.rs:172-191) and argument order is preserved (the divergence is not symmetric in f32).
    int kind = 0;
    GroundMetric ground;
    SinkhornParams hp;
    std::vector<Measure> pmeas, cmeas;
    std::vector<float> pself, cself;

    void build_point_measures() {
        pmeas.resize(N); pself.resize(N);
        parallel(N, [&](int i) {
            pmeas[i] = Measure::from_counts(&points[(size_t)i * B], B);
            pself[i] = ot_cost<Math::Contract>(pmeas[i], pmeas[i], ground, hp);
        });
    }
    static void centroid_measures(const Kmeans& km, const std::vector<uint64_t>& counts, std::vector<Measure>& meas, std::vector<float>& self) {
        meas.resize(km.K); self.resize(km.K);
        km.parallel(km.K, [&](int j) {
            meas[j] = Measure::from_counts(&counts[(size_t)j * km.B], km.B);
            self[j] = meas[j].idx.empty() ? 0.0f : ot_cost<Math::Contract>(meas[j], meas[j], km.ground, km.hp);
        });
    }
    float sk(const Measure& a, float sa, const Measure& b, float sb) const {
        return divergence_from(ot_cost<Math::Contract>(a, b, ground, hp), sa, sb);
    }

    // bins.rs:58-60
    static inline float dens(uint64_t count, uint64_t weight) { return (float)count / (float)weight; }
    // equity.rs:41-53 over generic (count, weight) accessors
    template <class FX, class FY>
    inline float variation(FX fx, uint64_t wx, FY fy, uint64_t wy) const {
        float cx = 0.0f, cy = 0.0f, acc = 0.0f;
        for (int b = 0; b < B; ++b) {
            cx += dens(fx(b), wx);
            cy += dens(fy(b), wy);
            acc += fabsf(cx - cy);
        }
        return acc / (float)B;
    }
    float d_pc(int i, int j) const {  // distance(x_i, c_j): refresh / rebound (elkan.rs:113-123)
        if (kind == 1) return sk(pmeas[i], pself[i], cmeas[j], cself[j]);
        return d_point_centroid(i, j);
    }
    float d_cp(int j, int i) const {  // distance(c_j, x_i): neighbor (elkan.rs:68-77)
        if (kind == 1) return sk(cmeas[j], cself[j], pmeas[i], pself[i]);
        return d_point_centroid(i, j);
    }
    float d_point_centroid(int i, int j) const {
        const uint8_t* p = &points[(size_t)i * B];
        const uint64_t* c = &ccounts[(size_t)j * B];
        return variation([&](int b) { return (uint64_t)p[b]; }, pweight[i], [&](int b) { return c[b]; }, cweight[j]);
    }
    float d_pp(int i, int k) const {  // distance(x_i, x_k): k-means++ (layer.rs:171)
        if (kind == 1) return sk(pmeas[i], pself[i], pmeas[k], pself[k]);
        return d_point_point(i, k);
    }
    float d_point_point(int i, int k) const {
        const uint8_t* p = &points[(size_t)i * B];
        const uint8_t* q = &points[(size_t)k * B];
        return variation([&](int b) { return (uint64_t)p[b]; }, pweight[i], [&](int b) { return (uint64_t)q[b]; }, pweight[k]);
    }
    float d_centroids(const uint64_t* a, uint64_t wa, const uint64_t* c, uint64_t wc) const {
        return variation([&](int b) { return a[b]; }, wa, [&](int b) { return c[b]; }, wc);
    }

    template <class F>
    void parallel(int n, F f) const {
        int T = threads < 1 ? 1 : threads;
        if (T == 1 || n < 2 * T) { for (int i = 0; i < n; ++i) f(i); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([=]() { for (int i = (int)((int64_t)n * t / T); i < (int)((int64_t)n * (t + 1) / T); ++i) f(i); });
        for (auto& x : th) x.join();
    }

    void set_centroid_from_point(int j, int i) {
        for (int b = 0; b < B; ++b) ccounts[(size_t)j * B + b] = points[(size_t)i * B + b];
        cweight[j] = pweight[i];
    }

    // layer.rs:140-181 k-means++ (potentials start at 1, pot = min(pot, d(x_new, h)^2), pot[chosen] = 0 first)
    std::vector<int> init_plusplus(uint64_t seed) {
        std::vector<float> pot(N, 1.0f);
        std::vector<int> chosen;
        Draw rng{seed};
        for (int r = 0; r < K; ++r) {
            Philox4 w = rng.at((uint32_t)r, 0, 0, TAG_KMEANSPP);
            unsigned __int128 total = 0;
            std::vector<uint64_t> q(N);
            for (int i = 0; i < N; ++i) {
                float p = pot[i] < 1048576.0f ? pot[i] : 1048576.0f;
                q[i] = (uint64_t)((double)p * 4294967296.0);
                total += q[i];
            }
            uint64_t T = (uint64_t)total;
            uint64_t word = (uint64_t)w.r[0] << 32 | w.r[1];
            uint64_t x = (uint64_t)(((unsigned __int128)word * T) >> 64);
            uint64_t cum = 0;
            int pick = N - 1;
            for (int i = 0; i < N; ++i) { cum += q[i]; if (x < cum) { pick = i; break; } }
            chosen.push_back(pick);
            set_centroid_from_point(r, pick);
            pot[pick] = 0.0f;
            parallel(N, [&](int i) {
                float d = d_pp(pick, i);  // distance(&x, h)
                float d2 = d * d;
                pot[i] = d2 < pot[i] ? d2 : pot[i];  // Energy::min(d0, d1)
            });
            dist_evals += N;
        }
        if (kind == 1) centroid_measures(*this, ccounts, cmeas, cself);
        return chosen;
    }

    // elkan.rs:68-77 neighbor: argmin_j distance(c_j, x), first minimum (Iterator::min_by)
    void neighbor(int i, uint32_t* j_out, float* d_out) const {
        int best = 0;
        float bd = d_cp(0, i);
        for (int j = 1; j < K; ++j) {
            float d = d_cp(j, i);
            if (d < bd) { bd = d; best = j; }
        }
        *j_out = (uint32_t)best;
        *d_out = bd;
    }
    // elkan.rs:39-47 init_bounds
    void init_bounds() {
        assign.assign(N, 0); upper.assign(N, 0.0f); stale.assign(N, 0);
        lower.assign((size_t)N * K, 0.0f);
        parallel(N, [&](int i) { neighbor(i, &assign[i], &upper[i]); });
        dist_evals += (uint64_t)N * K;
    }

    struct StepOut { std::vector<float> drift; std::vector<uint32_t> sizes; uint32_t reassigned; };
    // elkan.rs:153-168 step_elkan, split at the one exchange point of a point-sharded run: `step_local` is the point
    // pass plus this shard's integer accumulators [K][B+1] (counts, weight) and tallies; `step_finish` consumes the
    // (all-reduced) accumulators.  step() = both.
    std::vector<uint64_t> acc;        // [K][B + 1]
    std::vector<uint32_t> tally;      // sizes[K] + reassigned
    StepOut step() {
        step_local();
        return step_finish();
    }
    void step_local() {
        prior = assign;
        std::vector<float> pair((size_t)K * K, 0.0f), mid(K, FLT_MAX);
        parallel(K, [&](int i) {  // elkan.rs:80-93 pairwises (both triangles computed)
            for (int j = 0; j < K; ++j)
                pair[(size_t)i * K + j] = i == j ? 0.0f
                    : (kind == 1 ? sk(cmeas[i], cself[i], cmeas[j], cself[j])
                                 : d_centroids(&ccounts[(size_t)i * B], cweight[i], &ccounts[(size_t)j * B], cweight[j]));
        });
        for (int i = 0; i < K; ++i)  // elkan.rs:95-105 midpoints
            for (int j = 0; j < K; ++j)
                if (j != i) { float h = pair[(size_t)i * K + j] * 0.5f; mid[i] = mid[i] < h ? mid[i] : h; }
        std::vector<uint64_t> evals(threads < 1 ? 1 : threads, 0);
        parallel(N, [&](int i) {
            if (!(upper[i] > mid[assign[i]])) return;  // filter(|b| b.u() > midpoints[b.j()])
            float* l = &lower[(size_t)i * K];
            if (stale[i]) {  // elkan.rs:113-117 + bounds.rs:76-80 refresh
                float d = d_pc(i, (int)assign[i]);
                l[assign[i]] = d; upper[i] = d; stale[i] = 0;
            }
            for (int j = 0; j < K; ++j) {  // elkan.rs:118-123 rebound
                uint32_t c = assign[i];
                if ((int)c != j && upper[i] > l[j] && upper[i] > 0.5f * pair[(size_t)c * K + j]) {  // bounds.rs:57-61
                    float d = d_pc(i, j);
                    l[j] = d;  // bounds.rs:81-87 witness
                    if (d < upper[i]) { assign[i] = (uint32_t)j; upper[i] = d; }
                }
            }
        });
        // elkan.rs:125-142 recompute: integer merge of member points (bins.rs:75-82)
        acc.assign((size_t)K * (B + 1), 0);
        tally.assign(K + 1, 0);
        for (int i = 0; i < N; ++i) {
            uint32_t j = assign[i];
            acc[(size_t)j * (B + 1) + B] += pweight[i];
            for (int b = 0; b < B; ++b) acc[(size_t)j * (B + 1) + b] += points[(size_t)i * B + b];
            tally[j]++;                      // elkan/src/prior.rs:35-47
            tally[K] += assign[i] != prior[i];
        }
    }
    StepOut step_finish() {
        std::vector<uint64_t> nc((size_t)K * B, 0), nw(K, 0);
        for (int j = 0; j < K; ++j) {
            nw[j] = acc[(size_t)j * (B + 1) + B];
            for (int b = 0; b < B; ++b) nc[(size_t)j * B + b] = acc[(size_t)j * (B + 1) + b];
        }
        StepOut out;
        out.drift.resize(K);
        std::vector<Measure> nmeas;
        std::vector<float> nself;
        if (kind == 1) centroid_measures(*this, nc, nmeas, nself);
        for (int j = 0; j < K; ++j)  // elkan.rs:107-109 drift = distance(new, old)
            out.drift[j] = kind == 1 ? sk(nmeas[j], nself[j], cmeas[j], cself[j])
                                     : d_centroids(&nc[(size_t)j * B], nw[j], &ccounts[(size_t)j * B], cweight[j]);
        parallel(N, [&](int i) {  // bounds.rs:65-74 update
            float* l = &lower[(size_t)i * K];
            for (int j = 0; j < K; ++j) { float v = l[j] - out.drift[j]; l[j] = v > 0.0f ? v : 0.0f; }
            upper[i] += out.drift[assign[i]];
            stale[i] = 1;
        });
        ccounts.swap(nc); cweight.swap(nw);
        if (kind == 1) { cmeas.swap(nmeas); cself.swap(nself); }
        out.sizes.assign(tally.begin(), tally.begin() + K);
        out.reassigned = tally[K];
        return out;
    }

    // layer.rs:85-101 metric + metric.rs:127-141 normalisation; triangular layout pair.rs:36-39
    std::vector<float> metric() const {
        std::vector<float> tri((size_t)K * (K - 1) / 2, 0.0f);
        float mx = FLT_MIN;
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < i; ++j) {
                float a = kind == 1 ? sk(cmeas[i], cself[i], cmeas[j], cself[j])
                                    : d_centroids(&ccounts[(size_t)i * B], cweight[i], &ccounts[(size_t)j * B], cweight[j]);
                float b = kind == 1 ? sk(cmeas[j], cself[j], cmeas[i], cself[i])
                                    : d_centroids(&ccounts[(size_t)j * B], cweight[j], &ccounts[(size_t)i * B], cweight[i]);
                float d = (a + b) / 2.0f;
                tri[(size_t)i * (i - 1) / 2 + j] = d;
                mx = d > mx ? d : mx;  // fold(f32::MIN_POSITIVE, f32::max)
            }
        for (float& v : tri) v = v / mx;
        return tri;
    }
};

}  // namespace orc
