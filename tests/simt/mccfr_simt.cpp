// TEST INFRASTRUCTURE ONLY — drives the MCCFR kernels of robopoker_b200/csrc/mccfr.cu (their SOURCE, pasted in by tests/simt/build.py between
// the "device-side views" marker and the end of `namespace rbp`) under the SIMT shim.  The orchestration below restates what the host half of
// the library does around them (rbp_solver_create / rbp_solver_step / rbp_subgame_create / rbp_subgame_step): buffer sizes, launch shapes,
// EpochArgs, the world draw.  Compared against the oracle in tests/test_simt_mccfr.py.
#include "simt.hpp"

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "flat_game.hpp"

namespace rbp {
std::atomic<uint64_t> g_launches{0};
void set_last_error(const std::string&) {}
int cuda_fail(cudaError_t, const char*, const char*, int) { return RBP_ERR_CUDA; }
#include "mccfr_kernels.inc"

using namespace rbp;

namespace {
struct Sim {
    FlatGame G;
    DevGame d{};
    Scratch sc{};
    std::vector<rbp_encounter_t> table;
    std::vector<float> fb;
    std::vector<float> pay, dr, bp_f;
    std::vector<uint8_t> mask;
    std::vector<int32_t> m_off, m_cnt;
    std::vector<uint32_t> bp_na;
    unsigned long long counters[3] = {};
    int regret = 0, weight = 0, sampling = 0, batch = 1, worlds = 0, world = 0, entry_plus1 = 0;
    uint64_t seed = 0, epochs = 0;
    rbp_hyper_t hyper{};
    size_t sample_smem = 0, fold_smem = 0;
};
EpochArgs epoch_args(const Sim& s) {  // mccfr.cu epoch_args
    EpochArgs ep{};
    ep.seed_lo = (uint32_t)s.seed; ep.seed_hi = (uint32_t)(s.seed >> 32);
    ep.epoch = (uint32_t)s.epochs;
    ep.walker = (int)(s.epochs % 2);
    ep.batch = s.batch;
    ep.tree_base = s.worlds ? s.world : 0;
    ep.entry_plus1 = s.entry_plus1;
    ep.sampling = s.sampling;
    ep.hyper = s.hyper;
    ep.regret_sched = s.regret; ep.weight_sched = s.weight; ep.fold_mode = RBP_FOLD_ORDERED;
    ep.t = (float)s.epochs;
    ep.disc_pos = powf(ep.t / 1.0f, 1.5f);
    ep.disc_neg = powf(ep.t / 1.0f, 0.5f);
    return ep;
}
template <int RS, int WS>
void fold_rw(Sim& s, const EpochArgs& ep, rbp_encounter_t* table) {
    const DevGame d = s.d; const Scratch sc = s.sc;
    if (s.sampling != RBP_SAMPLING_EXTERNAL) simt::launch((unsigned)d.n_infos, 96, [=] { mccfr_fold_kernel<RS, WS, true>(d, table, sc, ep); });
    else simt::launch((unsigned)d.n_infos, 96, [=] { mccfr_fold_kernel<RS, WS, false>(d, table, sc, ep); });
}
template <int RS>
void fold_r(Sim& s, const EpochArgs& ep, rbp_encounter_t* t) {
    switch (s.weight) {
        case RBP_WEIGHT_CONSTANT: return fold_rw<RS, RBP_WEIGHT_CONSTANT>(s, ep, t);
        case RBP_WEIGHT_LINEAR: return fold_rw<RS, RBP_WEIGHT_LINEAR>(s, ep, t);
        case RBP_WEIGHT_QUADRATIC: return fold_rw<RS, RBP_WEIGHT_QUADRATIC>(s, ep, t);
        default: return fold_rw<RS, RBP_WEIGHT_EXPONENTIAL>(s, ep, t);
    }
}
void fold(Sim& s, const EpochArgs& ep, rbp_encounter_t* t) {
    switch (s.regret) {
        case RBP_REGRET_SUMMED: return fold_r<RBP_REGRET_SUMMED>(s, ep, t);
        case RBP_REGRET_FLOORED: return fold_r<RBP_REGRET_FLOORED>(s, ep, t);
        case RBP_REGRET_LINEAR: return fold_r<RBP_REGRET_LINEAR>(s, ep, t);
        case RBP_REGRET_DISCOUNTED: return fold_r<RBP_REGRET_DISCOUNTED>(s, ep, t);
        default: return fold_r<RBP_REGRET_ASYMMETRIC>(s, ep, t);
    }
}
void one_epoch(Sim& s) {
    const EpochArgs ep = epoch_args(s);
    rbp_encounter_t* table = s.table.data() + (size_t)s.world * s.G.n_rows;
    const float* fb = s.worlds ? s.fb.data() : nullptr;
    const DevGame d = s.d; const Scratch sc = s.sc;
    simt::launch((unsigned)sc.nblk, kTreesPerBlock, [=] { mccfr_sample_kernel(d, table, fb, sc, ep); });
    fold(s, ep, table);
    s.epochs += 1;
}
}  // namespace

extern "C" {
// rbp_solver_create's sizing, host memory instead of device memory
void* simt_create(int game, int regret, int weight, int sampling, int batch, uint64_t seed) {
    Sim* s = new Sim();
    if (!build_flat_game(game, &s->G)) { delete s; return nullptr; }
    const FlatGame& G = s->G;
    s->regret = regret; s->weight = weight; s->sampling = sampling; s->batch = batch; s->seed = seed;
    s->hyper = rbp_hyper_t{};  // rbp_hyper_default (mccfr.cu): hyperparams/{sampling,pruning,training}.rs
    s->hyper.temperature = 1.0f; s->hyper.smoothing = 2.0f; s->hyper.curiosity = 0.05f;
    s->hyper.prune_threshold = -3e5f; s->hyper.prune_explore = 0.05f; s->hyper.prune_warmup = 16384; s->hyper.regret_min = -4e6f;
    DevGame& d = s->d;
    d.nodes = G.nodes.data(); d.payoff1 = G.payoff1.data(); d.root_table = G.root_table.data(); d.info_row = G.info_row.data();
    d.info_actions = G.info_actions.data(); d.info_player = G.info_player.data(); d.parent = G.parent.data(); d.level_start = G.level_start.data();
    d.span_start = G.span_start.data(); d.span_nodes = G.span_nodes.data();
    d.n_nodes = (int)G.nodes.size(); d.n_infos = (int)G.info_key.size(); d.n_rows = G.n_rows; d.n_levels = (int)G.level_start.size() - 1; d.deck = G.deck;
    s->table.assign(G.n_rows, rbp_encounter_t{0.0f, 0.0f, 0.0f, 0u});
    Scratch& sc = s->sc;
    sc.nblk = (batch + kTreesPerBlock - 1) / kTreesPerBlock;
    sc.cap = kTreesPerBlock * G.max_tree_infos;
    const size_t total = (size_t)sc.nblk * sc.cap;
    s->pay.assign(total, 0.0f); s->dr.assign(total * kMaxActions, 0.0f); s->mask.assign(total, 0);
    s->m_off.assign((size_t)d.n_infos * sc.nblk, 0); s->m_cnt.assign((size_t)d.n_infos * sc.nblk, 0);
    sc.pay = s->pay.data(); sc.dr = s->dr.data(); sc.mask = s->mask.data(); sc.m_off = s->m_off.data(); sc.m_cnt = s->m_cnt.data();
    sc.counters = s->counters; sc.bp_f = nullptr; sc.bp_na = nullptr;
    return s;
}
void simt_destroy(void* h) { delete static_cast<Sim*>(h); }
void simt_step(void* h, uint64_t n) { Sim* s = static_cast<Sim*>(h); for (uint64_t i = 0; i < n; ++i) one_epoch(*s); }
// rows with visits > 0 of table `world`, (info_key, action) order; returns the count
int simt_export(void* h, int world, rbp_profile_row_t* rows, int cap) {
    Sim* s = static_cast<Sim*>(h);
    const rbp_encounter_t* t = s->table.data() + (size_t)world * s->G.n_rows;
    std::vector<rbp_profile_row_t> all;
    for (size_t x = 0; x < s->G.info_key.size(); ++x)
        for (int a = 0; a < s->G.info_actions[x]; ++a) {
            const int r = s->G.info_row[x] + a;
            if (t[r].visits > 0) all.push_back(rbp_profile_row_t{s->G.info_key[x], (uint32_t)a, t[r]});
        }
    std::sort(all.begin(), all.end(), [](const rbp_profile_row_t& p, const rbp_profile_row_t& q) { return p.info_key != q.info_key ? p.info_key < q.info_key : p.action < q.action; });
    for (int i = 0; i < (int)all.size() && i < cap; ++i) rows[i] = all[i];
    return (int)all.size();
}
// rbp_profile_import: rows into table 0
int simt_import(void* h, const rbp_profile_row_t* rows, int n, uint64_t epochs) {
    Sim* s = static_cast<Sim*>(h);
    std::fill(s->table.begin(), s->table.begin() + s->G.n_rows, rbp_encounter_t{0.0f, 0.0f, 0.0f, 0u});
    for (int i = 0; i < n; ++i) {
        int x = -1;
        for (size_t k = 0; k < s->G.info_key.size(); ++k) if (s->G.info_key[k] == rows[i].info_key) { x = (int)k; break; }
        if (x < 0 || rows[i].action >= s->G.info_actions[x]) return -1;
        s->table[s->G.info_row[x] + (int)rows[i].action] = rows[i].row;
    }
    s->epochs = epochs;
    return 0;
}
// rbp_subgame_create's device half: W tables seeded by subgame_seed_kernel from the blueprint simulator's table
void* simt_subgame_create(void* blueprint, int worlds, uint64_t seed) {
    Sim* bp = static_cast<Sim*>(blueprint);
    int game = -1;
    for (int id = 0; id < 3 && game < 0; ++id) { FlatGame probe; if (build_flat_game(id, &probe) && probe.nodes.size() == bp->G.nodes.size() && probe.n_rows == bp->G.n_rows) game = id; }
    Sim* s = static_cast<Sim*>(simt_create(game, RBP_REGRET_SUMMED, RBP_WEIGHT_LINEAR, RBP_SAMPLING_EXTERNAL, 1, seed));
    s->hyper = bp->hyper;
    s->worlds = worlds;
    s->table.assign((size_t)worlds * s->G.n_rows, rbp_encounter_t{0.0f, 0.0f, 0.0f, 0u});
    s->fb.assign((size_t)2 * s->G.n_rows, 0.0f);  // [weights | payoffs]
    const DevGame d = s->d;
    const rbp_encounter_t* src = bp->table.data();
    rbp_encounter_t* dst = s->table.data();
    float* fb = s->fb.data();
    simt::launch((unsigned)((d.n_infos + 127) / 128), 128, [=] { subgame_seed_kernel(d, src, 16384.0f, worlds, dst, fb); });
    return s;
}
// rbp_subgame_step: the world draw of the RNG contract, then one epoch from that world's entry node into that world's table
void simt_subgame_step(void* h, uint64_t n, const float* weights, const int32_t* entry_nodes, uint64_t* drawn) {
    Sim* s = static_cast<Sim*>(h);
    for (uint64_t i = 0; i < n; ++i) {
        const Philox4 p = philox4x32_10((uint32_t)s->epochs, 0u, 0xFFFFFFFEu, 5u, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
        float total = 0.0f;
        for (int w = 0; w < s->worlds; ++w) total = total + weights[w];
        const float x = draw_unit(p.r[0]) * total;
        float cum = 0.0f;
        int world = s->worlds - 1;
        for (int w = 0; w < s->worlds; ++w) { cum = cum + weights[w]; if (x < cum) { world = w; break; } }
        drawn[world] += 1;
        s->world = world;
        s->entry_plus1 = entry_nodes[world] + 1;
        one_epoch(*s);
    }
}
}
