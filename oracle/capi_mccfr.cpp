// ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes-facing C API over oracle/mccfr.hpp (tests, smoke and the
// cpu_baseline leg of bench.py are the only permitted callers).
#include <algorithm>
#include <cstdio>

#include "mccfr.hpp"
#include "subgame.hpp"

using namespace orc;

struct OrcSolver {
    int game;  // 0 kuhn, 1 leduc, 2 rps
    Solver<KuhnGame> kuhn;
    Solver<LeducGame> leduc;
    Solver<RpsGame> rps;
};

template <class F>
static auto with(OrcSolver* s, F f) {
    return s->game == 0 ? f(s->kuhn) : (s->game == 1 ? f(s->leduc) : f(s->rps));
}

extern "C" {

struct OrcRow {
    uint32_t info_key;
    uint32_t action;
    float weight, regret, payoff;
    uint32_t visits;
};

void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    Philox4 p = philox4x32_10(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out[i] = p.r[i];
}

OrcSolver* orc_solver_create(int game, int regret, int weight, int sampling, int batch, uint64_t seed, int threads) {
    if (game < 0 || game > 2) return nullptr;
    OrcSolver* s = new OrcSolver();
    s->game = game;
    auto init = [&](auto& sv) {
        sv.regret_sched = regret; sv.weight_sched = weight; sv.sampling = sampling;
        sv.batch = batch; sv.threads = threads; sv.rng.seed = seed;
        return 0;
    };
    with(s, init);
    return s;
}
void orc_solver_destroy(OrcSolver* s) { delete s; }
void orc_solver_step(OrcSolver* s, uint64_t n) {
    with(s, [&](auto& sv) { for (uint64_t i = 0; i < n; ++i) sv.step(); return 0; });
}
uint64_t orc_solver_epochs(OrcSolver* s) { return with(s, [](auto& sv) { return sv.profile.epochs; }); }
uint64_t orc_solver_updates(OrcSolver* s) { return with(s, [](auto& sv) { return sv.updates; }); }
uint64_t orc_solver_nodes(OrcSolver* s) { return with(s, [](auto& sv) { return sv.nodes; }); }
uint64_t orc_solver_infos(OrcSolver* s) { return with(s, [](auto& sv) { return sv.infos; }); }
float orc_solver_exploitability(OrcSolver* s) { return with(s, [](auto& sv) { return sv.exploitability(nullptr); }); }
void orc_solver_tree_stats(OrcSolver* s, int* out3) {
    with(s, [&](auto& sv) {
        typename std::decay_t<decltype(sv)>::ExplStats st{};
        sv.exploitability(&st);
        out3[0] = st.nodes; out3[1] = st.terminals; out3[2] = st.infosets;
        return 0;
    });
}
// rows sorted by (info_key, action); only rows that exist in the reference's HashMap sense
int orc_solver_export(OrcSolver* s, OrcRow* out, int cap) {
    return with(s, [&](auto& sv) {
        std::vector<OrcRow> rows;
        for (auto& kv : sv.profile.rows)
            for (int a = 0; a < kv.second.n; ++a)
                if (kv.second.present[a]) {
                    const Encounter& e = kv.second.e[a];
                    rows.push_back(OrcRow{kv.first, (uint32_t)a, e.weight, e.regret, e.payoff, e.visits});
                }
        std::sort(rows.begin(), rows.end(), [](const OrcRow& x, const OrcRow& y) {
            return x.info_key != y.info_key ? x.info_key < y.info_key : x.action < y.action;
        });
        int n = (int)rows.size();
        for (int i = 0; i < n && i < cap; ++i) out[i] = rows[i];
        return n;
    });
}
void orc_solver_import(OrcSolver* s, const OrcRow* in, int n, uint64_t epochs) {
    with(s, [&](auto& sv) {
        sv.profile.rows.clear();
        for (int i = 0; i < n; ++i) {
            Encounter& e = sv.profile.mut_row(in[i].info_key, (int)in[i].action);
            e = Encounter{in[i].weight, in[i].regret, in[i].payoff, in[i].visits};
        }
        sv.profile.epochs = epochs;
        return 0;
    });
}
void orc_solver_set_hyper(OrcSolver* s, float temperature, float smoothing, float curiosity, float prune_threshold,
                          float prune_explore, uint32_t prune_warmup, float regret_min) {
    with(s, [&](auto& sv) {
        sv.profile.hyper = Hyper{temperature, smoothing, curiosity, prune_threshold, prune_explore, prune_warmup, regret_min};
        return 0;
    });
}
void orc_solver_set_fold(OrcSolver* s, int fold_mode, int world_rank, int world_size) {
    with(s, [&](auto& sv) { sv.fold_mode = fold_mode; sv.world_rank = world_rank; sv.world_size = world_size; return 0; });
}
int orc_solver_partial_words(OrcSolver* s) {
    return with(s, [](auto& sv) { sv.ensure_info_order(); return (int)(sv.info_order.size() * sizeof(Partial) / 4); });
}
void orc_solver_sample(OrcSolver* s, uint32_t* words_out) {
    with(s, [&](auto& sv) {
        std::vector<Partial> p = sv.sample_partials();
        std::memcpy(words_out, p.data(), p.size() * sizeof(Partial));
        return 0;
    });
}
void orc_solver_fold_gathered(OrcSolver* s, const uint32_t* words, int world) {
    with(s, [&](auto& sv) { sv.fold_gathered(reinterpret_cast<const Partial*>(words), world); return 0; });
}
// averaged policy (Nash approximation) for one infoset: profile.rs:41-45
int orc_solver_averaged(OrcSolver* s, uint32_t info_key, float* out) {
    return with(s, [&](auto& sv) {
        auto v = sv.averaged(info_key);
        for (int a = 0; a < v.n; ++a) out[a] = v.p[a];
        return v.n;
    });
}

}  // extern "C"

// ── safe subgame solving on the small games (oracle/subgame.hpp) ──
struct OrcSubgame {
    int game;
    SubSolver<KuhnGame>* kuhn = nullptr;
    SubSolver<LeducGame>* leduc = nullptr;
    ~OrcSubgame() { delete kuhn; delete leduc; }
};
template <class G>
static typename G::State entry_state(typename G::State s, const uint8_t* path, int n) {
    for (int i = 0; i < n; ++i) {  // path = branch indices in `branches()` order
        uint8_t ed[MAX_BRANCH];
        const int k = G::branches(s, ed);
        if (path[i] >= k) break;
        s = G::apply(s, ed[path[i]]);
    }
    return s;
}
template <class F>
static auto with_sub(OrcSubgame* g, F f) { return g->game == 0 ? f(*g->kuhn) : f(*g->leduc); }

extern "C" {
// kuhn/src/solver.rs `subgame_with_reach_conditioned_posterior` + mccfr/src/solver/solver.rs:198-211 external_reach: per card the external
// player could hold, the product along the path of the blueprint's averaged policy at the external player's decisions; summed per rank
void orc_subgame_posterior(OrcSolver* bp, int external, int c0, int c1, const uint8_t* path, int path_len, float* reach3) {
    for (int r = 0; r < 3; ++r) reach3[r] = 0.0f;
    if (bp->game == 0) {
        using G = KuhnGame;
        G::State obs = entry_state<G>(G::State{{(uint8_t)c0, (uint8_t)c1}, G::Open}, path, path_len);
        for (uint8_t c = 0; c < 6; ++c) {
            if (c == obs.hole[1 - external] || G::board_is(obs, c)) continue;
            G::State s{{(uint8_t)c0, (uint8_t)c1}, G::Open};
            s.hole[external] = c;
            float reach = 1.0f;
            G::State o = G::State{{(uint8_t)c0, (uint8_t)c1}, G::Open};  // the observed trajectory names the edges
            for (int i = 0; i < path_len; ++i) {
                uint8_t ed[MAX_BRANCH];
                G::branches(o, ed);
                const uint8_t edge = ed[path[i]];
                if ((int)G::turn(s) == external) reach = reach * bp->kuhn.averaged_policy(G::info_key(s), edge);
                s = G::apply(s, edge); o = G::apply(o, edge);
            }
            reach3[G::rank(c)] += reach;
        }
    } else {
        using G = LeducGame;
        const G::State root{{(uint8_t)c0, (uint8_t)c1}, G::R1, 0, G::SOpen, 0, 0};
        G::State obs = entry_state<G>(root, path, path_len);
        for (uint8_t c = 0; c < 6; ++c) {
            if (c == obs.hole[1 - external] || G::board_is(obs, c)) continue;
            G::State s = root;
            s.hole[external] = c;
            float reach = 1.0f;
            G::State o = root;
            for (int i = 0; i < path_len; ++i) {
                uint8_t ed[MAX_BRANCH];
                G::branches(o, ed);
                const uint8_t edge = ed[path[i]];
                if ((int)G::turn(s) == external) reach = reach * bp->leduc.averaged_policy(G::info_key(s), edge);
                s = G::apply(s, edge); o = G::apply(o, edge);
            }
            reach3[G::rank(c)] += reach;
        }
    }
}
void orc_partition(const float* reach, int n, int worlds, int32_t* world_of, float* weights) { partition(reach, n, worlds, world_of, weights); }
OrcSubgame* orc_subgame_create(OrcSolver* bp, int external, int worlds, const int32_t* world_of_rank, const float* weights, int c0, int c1,
                               const uint8_t* path, int path_len, uint64_t seed) {
    if (!bp || bp->game > 1 || worlds < 1 || worlds > MAX_WORLDS) return nullptr;
    OrcSubgame* g = new OrcSubgame();
    g->game = bp->game;
    if (bp->game == 0) {
        KuhnGame::State root{{(uint8_t)c0, (uint8_t)c1}, KuhnGame::Open};
        g->kuhn = new SubSolver<KuhnGame>(&bp->kuhn, external, worlds, world_of_rank, weights, entry_state<KuhnGame>(root, path, path_len), seed);
    } else {
        LeducGame::State root{{(uint8_t)c0, (uint8_t)c1}, LeducGame::R1, 0, LeducGame::SOpen, 0, 0};
        g->leduc = new SubSolver<LeducGame>(&bp->leduc, external, worlds, world_of_rank, weights, entry_state<LeducGame>(root, path, path_len), seed);
    }
    return g;
}
void orc_subgame_destroy(OrcSubgame* g) { delete g; }
void orc_subgame_step(OrcSubgame* g, uint64_t n) { with_sub(g, [&](auto& sv) { for (uint64_t i = 0; i < n; ++i) sv.step(); return 0; }); }
uint64_t orc_subgame_t(OrcSubgame* g) { return with_sub(g, [](auto& sv) { return sv.t; }); }
void orc_subgame_drawn(OrcSubgame* g, uint64_t* out) { with_sub(g, [&](auto& sv) { for (int w = 0; w < sv.worlds; ++w) out[w] = sv.drawn[w]; return 0; }); }
float orc_subgame_sum_regret(OrcSubgame* g) { return with_sub(g, [](auto& sv) { return sv.sum_regret(); }); }
void orc_subgame_entry(OrcSubgame* g, int world, int* cards2) {
    with_sub(g, [&](auto& sv) { auto s = sv.restrict(world); cards2[0] = s.hole[0]; cards2[1] = s.hole[1]; return 0; });
}
uint32_t orc_subgame_entry_key(OrcSubgame* g, int world) {
    return with_sub(g, [&](auto& sv) { using GG = typename std::decay_t<decltype(sv)>::Game; return (uint32_t)GG::info_key(sv.restrict(world)); });
}
// rows of one world that exist locally, sorted by (info_key, action)
int orc_subgame_export(OrcSubgame* g, int world, OrcRow* out, int cap) {
    return with_sub(g, [&](auto& sv) {
        std::vector<OrcRow> rows;
        for (auto& kv : sv.local[world].profile.rows)
            for (int a = 0; a < kv.second.n; ++a)
                if (kv.second.present[a]) {
                    const Encounter& e = kv.second.e[a];
                    rows.push_back(OrcRow{kv.first, (uint32_t)a, e.weight, e.regret, e.payoff, e.visits});
                }
        std::sort(rows.begin(), rows.end(), [](const OrcRow& x, const OrcRow& y) { return x.info_key != y.info_key ? x.info_key < y.info_key : x.action < y.action; });
        const int n = (int)rows.size();
        for (int i = 0; i < n && i < cap; ++i) out[i] = rows[i];
        return n;
    });
}
// CfrNash::averaged_policy over WorldInfo(world, info): local weights, the blueprint's (floored) where the edge was never written
int orc_subgame_averaged(OrcSubgame* g, int world, uint32_t info_key, float* out) {
    return with_sub(g, [&](auto& sv) {
        auto v = sv.local[world].averaged(info_key);
        for (int a = 0; a < v.n; ++a) out[a] = v.p[a];
        return v.n;
    });
}
// Harvest (world/solver.rs:148-191) at a base infoset: refined[a] = sum over worlds of iterated policy / W, visits[a] = sum of cum_visits,
// regret = sum over edges and worlds of max(cum_regret, 0)
int orc_subgame_harvest(OrcSubgame* g, uint32_t info_key, float* refined, uint32_t* visits, float* regret) {
    return with_sub(g, [&](auto& sv) {
        int n = 0;
        float reg = 0.0f;
        for (int a = 0; a < MAXA; ++a) { refined[a] = 0.0f; visits[a] = 0u; }
        for (int w = 0; w < sv.worlds; ++w) {
            InfoView v = view_of(sv.local[w].profile, info_key);
            n = v.n;
            for (int a = 0; a < v.n; ++a) refined[a] += v.r[a] / v.rd / (float)sv.worlds;
        }
        for (int a = 0; a < n; ++a)
            for (int w = 0; w < sv.worlds; ++w) visits[a] += sv.local[w].profile.cum_visits(info_key, a);
        const InfoView v0 = view_of(sv.local[0].profile, info_key);
        for (int a = 0; a < n; ++a)
            for (int w = 0; w < sv.worlds; ++w) {
                const float r = sv.local[w].profile.cum_regret(info_key, a, v0.edges[a]);
                reg += r > 0.0f ? r : 0.0f;
            }
        *regret = reg;
        return n;
    });
}
}
