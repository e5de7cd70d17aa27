// rbp.hpp — C++ host-side mirror of the reference's trait surface over the C ABI of include/rbp.h.
//
// The reference is Rust; this image has no Rust toolchain, so the host side above the C ABI is provided in C++ (and in
// Python/ctypes, robopoker_b200/*.py, for the tests).  Names, argument meaning and error behaviour follow the reference:
//   Solver            trait Solver                         crates/mccfr/src/solver/solver.rs:38-350
//   Nlhe              Nlhe<R, W, S> (mccfr! expansion)     crates/nlhe/src/solver.rs:11, profile.rs:97-162
//   Layer             Layer<K, N> + trait Elkan            crates/lloyd/src/layer.rs:44-272, crates/elkan/src/elkan.rs:27-207
//   IsoSet            IsomorphismIterator + Lookup         crates/deuce/src/isomorphism_iter.rs, crates/lloyd/src/lookup.rs:46-192
//   strength / river_equity / sinkhorn_divergence          Strength::from(Hand), Observation::equity, Metric::emd
// The reference's convention on this path is to panic (`expect`), not to return `Result`: every wrapper throws rbp::Error
// on a non-zero status (message = rbp_status_string + rbp_last_error).  Handles are move-only RAII owners; nothing here
// computes on the host — without a CUDA device every constructor throws RBP_ERR_NO_DEVICE.  Header-only; link librbp_b200.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rbp.h"

namespace rbp {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& where)
        : std::runtime_error(where + ": " + rbp_status_string(s) + " (" + std::to_string(s) + ") " + rbp_last_error()), status(s) {}
};
inline void ok(int status, const char* where) {
    if (status != RBP_OK) throw Error(status, where);
}

enum class Game { Kuhn = RBP_GAME_KUHN, Leduc = RBP_GAME_LEDUC, Rps = RBP_GAME_RPS };
enum class Regret { Summed = RBP_REGRET_SUMMED, Floored = RBP_REGRET_FLOORED, Linear = RBP_REGRET_LINEAR, Discounted = RBP_REGRET_DISCOUNTED,
                    Asymmetric = RBP_REGRET_ASYMMETRIC };                                                        // crates/mccfr/src/regret/
enum class Weight { Constant = RBP_WEIGHT_CONSTANT, Linear = RBP_WEIGHT_LINEAR, Quadratic = RBP_WEIGHT_QUADRATIC,
                    Exponential = RBP_WEIGHT_EXPONENTIAL };                                                      // crates/mccfr/src/policy/
enum class Sampling { External = RBP_SAMPLING_EXTERNAL, Vanilla = RBP_SAMPLING_VANILLA, Prunable = RBP_SAMPLING_PRUNABLE,
                      Pluribus = RBP_SAMPLING_PLURIBUS, Targeted = RBP_SAMPLING_TARGETED };                      // crates/mccfr/src/sample/
enum class Fold { Ordered = RBP_FOLD_ORDERED, Batched = RBP_FOLD_BATCHED };
enum class Street { Pref = 0, Flop = 1, Turn = 2, Rive = 3 };                                                    // crates/deuce/src/street.rs

inline rbp_hyper_t hyper_default() {
    rbp_hyper_t h;
    rbp_hyper_default(&h);
    return h;
}

namespace detail {
template <class T, void (*Destroy)(T*)>
class Handle {
public:
    Handle() = default;
    explicit Handle(T* p) : p_(p) {}
    Handle(Handle&& o) noexcept : p_(std::exchange(o.p_, nullptr)) {}
    Handle& operator=(Handle&& o) noexcept {
        if (this != &o) { reset(); p_ = std::exchange(o.p_, nullptr); }
        return *this;
    }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    ~Handle() { reset(); }
    T* get() const { return p_; }

private:
    void reset() { if (p_) Destroy(p_); p_ = nullptr; }
    T* p_ = nullptr;
};
}  // namespace detail

// `Kuhn::<R, W, S>::default()` / `Leduc::<R, W, S>::default()` / `Rps::<R, W, S>::default()` with `batch_size()` = batch.
class Solver {
public:
    Solver(Game game, Regret regret = Regret::Floored, Weight weight = Weight::Linear, Sampling sampling = Sampling::External, int batch = 1,
           uint64_t seed = 0, Fold fold = Fold::Ordered, const rbp_hyper_t* hyper = nullptr, int device = 0)
        : batch_(batch) {
        rbp_solver_t* h = nullptr;
        ok(rbp_solver_create((int)game, (int)regret, (int)weight, (int)sampling, (int)fold, batch, seed, hyper, device, &h), "rbp_solver_create");
        h_ = decltype(h_)(h);
    }
    int batch_size() const { return batch_; }
    Solver& step(uint64_t n = 1) { ok(rbp_solver_step(h_.get(), n), "rbp_solver_step"); return *this; }                 // Solver::step x n
    Solver& solve(uint64_t trees) { return step(trees / (uint64_t)batch_); }                                            // Solver::solve (solver.rs:111-122)
    uint64_t epochs() const { uint64_t t = 0; ok(rbp_solver_epochs(h_.get(), &t), "rbp_solver_epochs"); return t; }     // RefProf::t
    float exploitability() const { float e = 0; ok(rbp_solver_exploitability(h_.get(), &e), "rbp_solver_exploitability"); return e; }
    std::vector<rbp_profile_row_t> profile() const {                                                                    // RefProf::cum_* in bulk
        int shape[6];
        ok(rbp_solver_game_shape(h_.get(), shape), "rbp_solver_game_shape");
        std::vector<rbp_profile_row_t> rows((size_t)shape[3]);
        int n = 0;
        ok(rbp_profile_export(h_.get(), rows.data(), (int)rows.size(), &n), "rbp_profile_export");
        rows.resize((size_t)n);
        return rows;
    }
    void storage(const std::vector<rbp_profile_row_t>& rows, uint64_t epochs) {                                         // MutProf::mut_* in bulk
        ok(rbp_profile_import(h_.get(), rows.data(), (int)rows.size(), epochs), "rbp_profile_import");
    }
    std::vector<float> averaged_distribution(uint32_t info_key) const {                                                 // RefProf::averaged_distribution
        float p[16];
        int n = 0;
        ok(rbp_profile_averaged(h_.get(), info_key, p, 16, &n), "rbp_profile_averaged");
        return std::vector<float>(p, p + n);
    }
    rbp_solver_t* raw() const { return h_.get(); }

private:
    detail::Handle<rbp_solver_t, rbp_solver_destroy> h_;
    int batch_;
};

// `WorldSolver` (crates/subgame/src/world/solver.rs:33-146) = `SubGameSolver::new` (crates/subgame/src/solver.rs:46-70) on Kuhn / Leduc
struct Belief {                                                                                                          // world/belief.rs:20-27
    std::vector<int32_t> world_of_rank;  // empty = no members: every secret is remembered
    std::vector<float> weights;
};
inline Belief partition(const std::vector<float>& reach, int worlds) {                                                   // world/partition.rs:27-53
    Belief b{std::vector<int32_t>(reach.size()), std::vector<float>((size_t)worlds)};
    ok(rbp_subgame_partition(reach.data(), (int)reach.size(), worlds, b.world_of_rank.data(), b.weights.data()), "rbp_subgame_partition");
    return b;
}
class WorldSolver {
public:
    struct Harvested { std::vector<float> refined; std::vector<uint32_t> visits; float regret; };                        // mccfr Harvested<E>
    WorldSolver(const Solver& blueprint, int external, const Belief& belief, int c0, int c1, const std::vector<uint8_t>& path = {}, uint64_t seed = 0)
        : worlds_((int)belief.weights.size()) {
        rbp_subgame_t* h = nullptr;
        ok(rbp_subgame_create(blueprint.raw(), external, worlds_, belief.world_of_rank.empty() ? nullptr : belief.world_of_rank.data(), belief.weights.data(), c0, c1,
                              path.empty() ? nullptr : path.data(), (int)path.size(), seed, &h), "rbp_subgame_create");
        h_ = decltype(h_)(h);
    }
    WorldSolver& step(uint64_t n = 1) { ok(rbp_subgame_step(h_.get(), n), "rbp_subgame_step"); return *this; }           // world/solver.rs:118-146
    WorldSolver& solve(uint64_t trees) { return step(trees); }                                                          // batch_size() = 1
    std::pair<uint64_t, double> spend(double seconds) {                                                                 // Solver::spend
        uint64_t n = 0; double dt = 0;
        ok(rbp_subgame_spend(h_.get(), seconds, &n, &dt), "rbp_subgame_spend");
        return {n, dt};
    }
    std::vector<rbp_profile_row_t> into_profile(int world) const {                                                       // WorldProfile::local of one world
        std::vector<rbp_profile_row_t> rows(4096);
        int n = 0;
        ok(rbp_subgame_export(h_.get(), world, rows.data(), (int)rows.size(), &n), "rbp_subgame_export");
        rows.resize((size_t)n);
        return rows;
    }
    Harvested harvest(uint32_t info_key) const {                                                                         // world/solver.rs:148-191
        float refined[8], regret = 0; uint32_t visits[8]; int n = 0;
        ok(rbp_subgame_harvest(h_.get(), info_key, refined, visits, &regret, 8, &n), "rbp_subgame_harvest");
        return Harvested{std::vector<float>(refined, refined + n), std::vector<uint32_t>(visits, visits + n), regret};
    }
    int worlds() const { return worlds_; }

private:
    detail::Handle<rbp_subgame_t, rbp_subgame_destroy> h_;
    int worlds_;
};

class IsoSet;

// `Nlhe<R, W, S>`; the defaults are `Flagship` (crates/nlhe/src/lib.rs:86-90).
class Nlhe {
public:
    Nlhe(int batch = 128, uint64_t seed = 0, Regret regret = Regret::Linear, Weight weight = Weight::Linear, Sampling sampling = Sampling::Pluribus,
         uint64_t table_slots = 0, const rbp_hyper_t* hyper = nullptr, int max_nodes_per_tree = 0, int device = 0) {
        rbp_nlhe_t* h = nullptr;
        ok(rbp_nlhe_create((int)regret, (int)weight, (int)sampling, batch, seed, hyper, table_slots, max_nodes_per_tree, device, &h), "rbp_nlhe_create");
        h_ = decltype(h_)(h);
    }
    Nlhe& step(uint64_t n = 1) { ok(rbp_nlhe_step(h_.get(), n), "rbp_nlhe_step"); return *this; }
    struct Counters { uint64_t epochs, nodes, infos, updates, rows, records, max_tree; };
    Counters counters() const {
        uint64_t c[8];
        ok(rbp_nlhe_counters(h_.get(), c), "rbp_nlhe_counters");
        return Counters{c[0], c[1], c[2], c[3], c[4], c[5], c[6]};
    }
    std::vector<rbp_nlhe_row_t> rows() const {                                                                          // NlheProfile::rows (profile.rs:143-160)
        uint64_t n = 0;
        ok(rbp_nlhe_export(h_.get(), nullptr, 0, &n), "rbp_nlhe_export");
        std::vector<rbp_nlhe_row_t> out((size_t)n);
        ok(rbp_nlhe_export(h_.get(), out.data(), n, &n), "rbp_nlhe_export");
        out.resize((size_t)n);
        return out;
    }
    void hydrate(const std::vector<rbp_nlhe_row_t>& rows, uint64_t epochs) {                                            // Hydrate (profile.rs:97-141)
        ok(rbp_nlhe_import(h_.get(), rows.data(), rows.size(), epochs), "rbp_nlhe_import");
    }
    inline void set_lookup(IsoSet& isos);                                                                               // NlheEncoder's table of one street
    void set_lookup_rows(const std::vector<int64_t>& obs, const std::vector<int16_t>& abs) {
        if (obs.size() != abs.size()) throw Error(RBP_ERR_INVALID, "Nlhe::set_lookup_rows");
        ok(rbp_nlhe_set_lookup_rows(h_.get(), obs.data(), abs.data(), (int64_t)obs.size()), "rbp_nlhe_set_lookup_rows");
    }
    rbp_nlhe_t* raw() const { return h_.get(); }

private:
    detail::Handle<rbp_nlhe_t, rbp_nlhe_destroy> h_;
};

// `IsomorphismIterator::from(street)` resident on the device, with its `Lookup` column.
class IsoSet {
public:
    explicit IsoSet(Street street, int device = 0) {
        rbp_isoset_t* h = nullptr;
        ok(rbp_isoset_create((int)street, device, &h), "rbp_isoset_create");
        h_ = decltype(h_)(h);
    }
    int64_t size() const { return rbp_isoset_size(h_.get()); }
    void river_buckets() { ok(rbp_isoset_river_buckets(h_.get()), "rbp_isoset_river_buckets"); }                       // Lookup::grow(Street::Rive)
    void set_abstractions(const std::vector<uint8_t>& abs) {
        if ((int64_t)abs.size() != size()) throw Error(RBP_ERR_INVALID, "IsoSet::set_abstractions");
        ok(rbp_isoset_set_abstractions(h_.get(), abs.data()), "rbp_isoset_set_abstractions");
    }
    // Lookup::projections: histograms [count][bins] of the parents [offset, offset + count) over this (child) set's abstractions
    std::vector<uint8_t> project_from(IsoSet& parent, int bins, int64_t offset, int64_t count, uint64_t* misses = nullptr) {
        std::vector<uint8_t> hist((size_t)count * (size_t)bins);
        uint64_t m = 0;
        ok(rbp_isoset_project(parent.h_.get(), h_.get(), bins, offset, count, hist.data(), &m), "rbp_isoset_project");
        if (misses) *misses = m;
        return hist;
    }
    void rows(int64_t offset, int64_t count, std::vector<int64_t>& obs, std::vector<int16_t>& abs) const {              // Streamable::rows
        obs.resize((size_t)count); abs.resize((size_t)count);
        ok(rbp_isoset_export_rows(h_.get(), offset, count, obs.data(), abs.data()), "rbp_isoset_export_rows");
    }
    rbp_isoset_t* raw() const { return h_.get(); }

private:
    detail::Handle<rbp_isoset_t, rbp_isoset_destroy> h_;
};
inline void Nlhe::set_lookup(IsoSet& isos) { ok(rbp_nlhe_set_lookup(h_.get(), isos.raw()), "rbp_nlhe_set_lookup"); }

// `Layer<K, N>`: k-means over `n` dense u8 histograms of `bins` buckets.  Without a metric it is the turn layer
// (`Equity::variation`); with `set_metric` called right after construction it must have been created as a Sinkhorn layer.
class Layer {
public:
    struct Step { std::vector<float> drift; std::vector<uint32_t> sizes; uint32_t reassignment; };                       // crates/elkan/src/step.rs
    Layer(const uint8_t* counts, int64_t n, int k, int bins, bool sinkhorn = false, int device = 0) : n_(n), k_(k), bins_(bins) {
        rbp_kmeans_t* h = nullptr;
        ok(rbp_kmeans_create(sinkhorn ? RBP_KMEANS_SINKHORN : RBP_KMEANS_W1, n, k, bins, counts, device, &h), "rbp_kmeans_create");
        h_ = decltype(h_)(h);
    }
    void set_metric(const std::vector<float>& tri) {                                                                    // Metric of the next street
        if ((int64_t)tri.size() != (int64_t)bins_ * (bins_ - 1) / 2) throw Error(RBP_ERR_INVALID, "Layer::set_metric");
        ok(rbp_kmeans_set_metric(h_.get(), tri.data(), bins_), "rbp_kmeans_set_metric");
    }
    std::vector<int32_t> init_centroids(uint64_t seed = 0) {                                                            // Layer::init_centroids
        std::vector<int32_t> chosen((size_t)k_);
        ok(rbp_kmeans_init_pp(h_.get(), seed, chosen.data()), "rbp_kmeans_init_pp");
        return chosen;
    }
    void init_bounds() { ok(rbp_kmeans_init_bounds(h_.get()), "rbp_kmeans_init_bounds"); }                             // Elkan::init_bounds
    Step step_elkan() {                                                                                                 // Elkan::step_elkan
        Step s{std::vector<float>((size_t)k_), std::vector<uint32_t>((size_t)k_), 0u};
        ok(rbp_kmeans_step(h_.get(), s.drift.data(), s.sizes.data(), &s.reassignment), "rbp_kmeans_step");
        return s;
    }
    std::vector<uint32_t> lookup(std::vector<float>* distance = nullptr) {                                              // Layer::lookup
        std::vector<uint32_t> a((size_t)n_);
        if (distance) distance->resize((size_t)n_);
        ok(rbp_kmeans_assign(h_.get(), a.data(), distance ? distance->data() : nullptr), "rbp_kmeans_assign");
        return a;
    }
    std::vector<float> metric() {                                                                                       // Layer::metric
        std::vector<float> tri((size_t)k_ * (size_t)(k_ - 1) / 2);
        if (!tri.empty()) ok(rbp_kmeans_metric(h_.get(), tri.data()), "rbp_kmeans_metric");
        return tri;
    }
    std::vector<uint64_t> future(std::vector<uint64_t>* weights = nullptr) {                                            // Layer::future
        std::vector<uint64_t> counts((size_t)k_ * (size_t)bins_);
        if (weights) weights->resize((size_t)k_);
        ok(rbp_kmeans_centroids(h_.get(), counts.data(), weights ? weights->data() : nullptr), "rbp_kmeans_centroids");
        return counts;
    }
    rbp_kmeans_t* raw() const { return h_.get(); }

private:
    detail::Handle<rbp_kmeans_t, rbp_kmeans_destroy> h_;
    int64_t n_;
    int k_, bins_;
};

// `Strength::from(Hand)` for a batch (packed so that integer order = the reference's `Ord`)
inline std::vector<uint32_t> strength(const std::vector<uint64_t>& hands) {
    std::vector<uint32_t> out(hands.size());
    ok(rbp_eval_batch(hands.data(), (int64_t)hands.size(), out.data()), "rbp_eval_batch");
    return out;
}
// `Observation::equity` on river observations; returns equities, optionally the `Abstraction::from(equity)` buckets
inline std::vector<float> river_equity(const std::vector<uint64_t>& pocket, const std::vector<uint64_t>& pub, std::vector<uint8_t>* bucket = nullptr) {
    if (pocket.size() != pub.size()) throw Error(RBP_ERR_INVALID, "river_equity");
    std::vector<float> eq(pocket.size());
    if (bucket) bucket->resize(pocket.size());
    ok(rbp_river_equity_batch(pocket.data(), pub.data(), (int64_t)pocket.size(), eq.data(), bucket ? bucket->data() : nullptr, nullptr, nullptr),
       "rbp_river_equity_batch");
    return eq;
}
// `Metric::emd` → `Sinkhorn::divergence` for explicit pairs of dense u32 histograms (hyper-parameters of
// crates/lloyd/src/hyperparams/sinkhorn.rs:17-23)
inline std::vector<float> sinkhorn_divergence(const std::vector<uint32_t>& a, int na, const std::vector<uint32_t>& b, int nb, int bins,
                                              const std::vector<int32_t>& ia, const std::vector<int32_t>& ib, const std::vector<float>& tri,
                                              float temperature = 0.025f, int iterations = 128, float tolerance = 0.0005f) {
    if (ia.size() != ib.size() || a.size() != (size_t)na * (size_t)bins || b.size() != (size_t)nb * (size_t)bins) throw Error(RBP_ERR_INVALID, "sinkhorn_divergence");
    std::vector<float> out(ia.size());
    ok(rbp_sinkhorn_batch(a.data(), na, b.data(), nb, bins, ia.data(), ib.data(), (int64_t)ia.size(), tri.data(), temperature, iterations, tolerance,
                          out.data()),
       "rbp_sinkhorn_batch");
    return out;
}

}  // namespace rbp
