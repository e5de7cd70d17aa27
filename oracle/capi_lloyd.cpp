// ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes-facing C API over oracle/lloyd.hpp.
#include <cstring>

#include "lloyd.hpp"

using orc::Kmeans;

extern "C" {
Kmeans* orc_kmeans_create(const uint8_t* counts, int n, int bins, int k, int threads) {
    Kmeans* m = new Kmeans();
    m->N = n; m->K = k; m->B = bins; m->threads = threads;
    m->points.assign(counts, counts + (size_t)n * bins);
    m->pweight.resize(n);
    for (int i = 0; i < n; ++i) {
        uint32_t w = 0;
        for (int b = 0; b < bins; ++b) w += counts[(size_t)i * bins + b];
        m->pweight[i] = w;
    }
    m->ccounts.assign((size_t)k * bins, 0);
    m->cweight.assign(k, 0);
    return m;
}
void orc_kmeans_destroy(Kmeans* m) { delete m; }
// flop layer: switch to Sinkhorn divergence under the ground metric `tri` (Pair::merge order)
void orc_kmeans_set_metric(Kmeans* m, const float* tri) {
    m->kind = 1;
    m->ground.bins = m->B;
    m->ground.tri.assign(tri, tri + (size_t)m->B * (m->B - 1) / 2);
    m->build_point_measures();
}
void orc_kmeans_init_pp(Kmeans* m, uint64_t seed, int* chosen) {
    std::vector<int> c = m->init_plusplus(seed);
    if (chosen) std::memcpy(chosen, c.data(), c.size() * sizeof(int));
}
void orc_kmeans_set_centroids_from_points(Kmeans* m, const int* idx) {
    for (int j = 0; j < m->K; ++j) m->set_centroid_from_point(j, idx[j]);
    if (m->kind == 1) Kmeans::centroid_measures(*m, m->ccounts, m->cmeas, m->cself);
}
void orc_kmeans_set_centroids(Kmeans* m, const uint64_t* counts) {
    for (int j = 0; j < m->K; ++j) {
        uint64_t w = 0;
        for (int b = 0; b < m->B; ++b) { m->ccounts[(size_t)j * m->B + b] = counts[(size_t)j * m->B + b]; w += counts[(size_t)j * m->B + b]; }
        m->cweight[j] = w;
    }
    if (m->kind == 1) Kmeans::centroid_measures(*m, m->ccounts, m->cmeas, m->cself);
}
void orc_kmeans_init_bounds(Kmeans* m) { m->init_bounds(); }
void orc_kmeans_step(Kmeans* m, float* drift, uint32_t* sizes, uint32_t* reassigned) {
    Kmeans::StepOut o = m->step();
    if (drift) std::memcpy(drift, o.drift.data(), o.drift.size() * sizeof(float));
    if (sizes) std::memcpy(sizes, o.sizes.data(), o.sizes.size() * sizeof(uint32_t));
    if (reassigned) *reassigned = o.reassigned;
}
void orc_kmeans_step_local(Kmeans* m) { m->step_local(); }
uint64_t* orc_kmeans_acc(Kmeans* m, int64_t* n) { *n = (int64_t)m->acc.size(); return m->acc.data(); }
uint32_t* orc_kmeans_tally(Kmeans* m, int64_t* n) { *n = (int64_t)m->tally.size(); return m->tally.data(); }
void orc_kmeans_step_finish(Kmeans* m, float* drift, uint32_t* sizes, uint32_t* reassigned) {
    Kmeans::StepOut o = m->step_finish();
    if (drift) std::memcpy(drift, o.drift.data(), o.drift.size() * sizeof(float));
    if (sizes) std::memcpy(sizes, o.sizes.data(), o.sizes.size() * sizeof(uint32_t));
    if (reassigned) *reassigned = o.reassigned;
}
void orc_kmeans_state(Kmeans* m, uint32_t* assign, float* upper, float* lower, uint8_t* stale) {
    if (assign) std::memcpy(assign, m->assign.data(), m->assign.size() * 4);
    if (upper) std::memcpy(upper, m->upper.data(), m->upper.size() * 4);
    if (lower) std::memcpy(lower, m->lower.data(), m->lower.size() * 4);
    if (stale) std::memcpy(stale, m->stale.data(), m->stale.size());
}
void orc_kmeans_centroids(Kmeans* m, uint64_t* counts, uint64_t* weights) {
    if (counts) std::memcpy(counts, m->ccounts.data(), m->ccounts.size() * 8);
    if (weights) std::memcpy(weights, m->cweight.data(), m->cweight.size() * 8);
}
// layer.rs:44-60 lookup: fresh naive argmin against the current centroids
void orc_kmeans_assign(Kmeans* m, uint32_t* out, float* dist) {
    m->parallel(m->N, [&](int i) {
        uint32_t j; float d;
        m->neighbor(i, &j, &d);
        out[i] = j;
        if (dist) dist[i] = d;
    });
}
void orc_kmeans_metric(Kmeans* m, float* tri) {
    std::vector<float> t = m->metric();
    std::memcpy(tri, t.data(), t.size() * 4);
}
float orc_variation(const uint32_t* x, const uint32_t* y, int bins) {
    Kmeans m; m.B = bins;
    uint64_t wx = 0, wy = 0;
    for (int b = 0; b < bins; ++b) { wx += x[b]; wy += y[b]; }
    return m.variation([&](int b) { return (uint64_t)x[b]; }, wx, [&](int b) { return (uint64_t)y[b]; }, wy);
}
uint64_t orc_kmeans_dist_evals(Kmeans* m) { return m->dist_evals; }
}
