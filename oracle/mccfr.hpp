// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).  CPU restatement of the reference's
// MCCFR path in plain C++17, strict left-to-right f32 arithmetic (build with -ffp-contract=off, no
// fast-math).  Parity status: the reference holds NO golden numeric vectors for this path and its
// RNG stream is unreproducible — pinned here are the reference's own statistical
// tests (Kuhn analytic Nash, Kuhn/Leduc exploitability thresholds, Leduc tree counts); the sampled
// stream itself is "parity unpinned" and is defined by the Philox contract in rng.hpp.
//
// Follows (all under /root/reference/crates/mccfr/src):
//   solver/solver.rs:96-105,143-192,225-305,327-338   step / update_* / batch / tree / exploitability
//   solver/builder.rs:74-161                          TreeBuilder (LIFO expansion)
//   state/tree.rs:76-97, state/node.rs:82-170         Tree::seed/grow/partition, Node edges (petgraph
//                                                     0.6 adjacency: newest out-edge first)
//   strategy/flow.rs:20-87,166-216                    regret matching, sampling distribution, dfs
//   strategy/profile.rs:31-51, strategy/book.rs:40-145 floored reads, defaults, walker, epochs
//   strategy/nash.rs:31-193                           exploitability
//   sample/{external,pruning,pluribus,vanilla,mod}.rs samplers
//   regret/*.rs, policy/*.rs                          schedules
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

#include "games.hpp"
#include "pool.hpp"
#include "rng.hpp"

namespace orc {

constexpr float EPS = FLT_MIN;  // pokerkit/src/lib.rs:204  EPSILON = f32::MIN_POSITIVE
constexpr int MAXA = 4;

struct Encounter {  // solver/encounter.rs:21-27
    float weight, regret, payoff;
    uint32_t visits;
};

enum RegretSched { R_SUMMED = 0, R_FLOORED = 1, R_LINEAR = 2, R_DISCOUNTED = 3, R_ASYMMETRIC = 4 };
enum WeightSched { W_CONSTANT = 0, W_LINEAR = 1, W_QUADRATIC = 2, W_EXPONENTIAL = 3 };
enum SamplingKind { S_EXTERNAL = 0, S_VANILLA = 1, S_PRUNABLE = 2, S_PLURIBUS = 3, S_TARGETED = 4 };
enum FoldMode { FOLD_ORDERED = 0, FOLD_BATCHED = 1 };

struct Hyper {
    float temperature = 1.0f, smoothing = 2.0f, curiosity = 0.05f;      // hyperparams/sampling.rs:39-50
    float prune_threshold = -3e5f, prune_explore = 0.05f;               // hyperparams/pruning.rs:40-55
    uint32_t prune_warmup = 16384;
    float regret_min = -4e6f;                                           // hyperparams/training.rs:58
};

// regret/*.rs — `gain` = max(accumulate, floor)
inline float regret_gain(int sched, float net, float add, uint64_t epoch, const Hyper& h) {
    float t = (float)epoch;
    float acc, floor = h.regret_min;
    switch (sched) {
        case R_SUMMED: acc = net + add; floor = -INFINITY; break;           // summed.rs:13-19
        case R_FLOORED: acc = net + add; floor = 0.0f; break;               // floored.rs:13-19
        case R_LINEAR: { float d = t / (t + 1.0f); acc = net * d + add; break; }  // linear.rs:13-17
        case R_DISCOUNTED: {                                                // discounted.rs:27-45 (PERIOD=1)
            float x;
            if (net > 0.0f) x = powf(t / 1.0f, 1.5f);
            else if (net < 0.0f) x = powf(t / 1.0f, 0.5f);
            else x = t / 1.0f;
            float d = x / (x + 1.0f);
            acc = net * d + add;
            break;
        }
        default: {                                                          // asymmetric.rs:13-21
            if (net > 0.0f) acc = net + add;
            else { float d = t / (t + 1.0f); acc = net * d + add; }
            break;
        }
    }
    return acc > floor ? acc : floor;  // f32::max; acc is never NaN here
}
// policy/*.rs — `learn` = max(accumulate, EPSILON)
inline float weight_learn(int sched, float net, float add, uint64_t epoch) {
    float t = (float)epoch, acc;
    switch (sched) {
        case W_CONSTANT: acc = net + add; break;
        case W_LINEAR: acc = net + add * t; break;
        case W_QUADRATIC: acc = net + add * t * t; break;
        default: acc = net * 0.9999f + add; break;
    }
    return acc > EPS ? acc : EPS;
}

template <class G>
struct Profile {  // strategy/book.rs (HashMap<I, HashMap<E, Encounter>> + epochs)
    struct Row {
        Encounter e[MAXA];
        bool present[MAXA];
        uint8_t edges[MAXA];
        int n;
    };
    std::unordered_map<uint32_t, Row> rows;
    uint64_t epochs = 0;
    Hyper hyper;
    // subgame/src/world/profile.rs:25-36: a WorldProfile is this table (`local`) over a frozen blueprint (`global`); null for a plain profile
    const Profile* blueprint = nullptr;
    float prior_strength = 16384.0f;  // hyperparams/warmstart.rs:27-33 (1 << 14)

    const Row* find(uint32_t info) const {
        auto it = rows.find(info);
        return it == rows.end() ? nullptr : &it->second;
    }
    Row& entry(uint32_t info) {
        auto it = rows.find(info);
        if (it != rows.end()) return it->second;
        Row r{};
        uint8_t ed[MAX_BRANCH];
        r.n = G::choices(info, ed);
        for (int a = 0; a < r.n; ++a) r.edges[a] = ed[a];
        return rows.emplace(info, r).first->second;
    }
    Encounter& mut_row(uint32_t info, int a) {  // book.rs:40-90: or_insert_with(Encounter::from(edge))
        Row& r = entry(info);
        if (!r.present[a]) {
            r.present[a] = true;
            // world/profile.rs:62-109: or_insert_with(|| blueprint.warmstart(&info.inner(), edge))
            r.e[a] = blueprint ? blueprint->warmstart(info, a, r.edges[a], prior_strength)
                               : Encounter{0.0f /*default_policy*/, G::default_regret(r.edges[a]), 0.0f, 0};
        }
        return r.e[a];
    }
    // book.rs:93-122; world/profile.rs:118-145: an edge never written falls through to the blueprint, floored at EPSILON
    float cum_regret(uint32_t info, int a, uint8_t edge) const {
        const Row* r = find(info);
        if (r && r->present[a]) return r->e[a].regret;
        if (blueprint) { const float g = blueprint->cum_regret(info, a, edge); return g > EPS ? g : EPS; }
        return G::default_regret(edge);
    }
    float cum_weight(uint32_t info, int a) const {
        const Row* r = find(info);
        if (r && r->present[a]) return r->e[a].weight;
        if (blueprint) { const float g = blueprint->cum_weight(info, a); return g > EPS ? g : EPS; }
        return 0.0f;
    }
    uint32_t cum_visits(uint32_t info, int a) const {
        const Row* r = find(info);
        if (r && r->present[a]) return r->e[a].visits;
        return blueprint ? blueprint->cum_visits(info, a) : 0u;
    }
    float cum_payoff(uint32_t info, int a) const {  // book.rs:93-122; world/profile.rs:133-138 (no floor)
        const Row* r = find(info);
        if (r && r->present[a]) return r->e[a].payoff;
        return blueprint ? blueprint->cum_payoff(info, a) : 0.0f;
    }
    // strategy/profile.rs:94-104 warmstart (on the blueprint): weight = averaged policy * k * (k + 1) / 2, regret = cum_regret * k / max(t, 1)
    Encounter warmstart(uint32_t info, int a, uint8_t edge, float k) const {
        uint8_t ed[MAX_BRANCH];
        const int n = G::choices(info, ed);
        float w[MAXA], sum = 0.0f;  // profile.rs:40-44 averaged_distribution
        for (int b = 0; b < n; ++b) { const float cw = cum_weight(info, b); w[b] = cw > EPS ? cw : EPS; sum = sum + w[b]; }
        const float policy = w[a] / sum;
        const float regret_scale = k / (float)(epochs > 1 ? epochs : 1);
        return Encounter{policy * k * (k + 1.0f) / 2.0f, cum_regret(info, a, edge) * regret_scale, 0.0f, 0u};
    }
    int walker() const { return (int)(epochs % 2); }  // book.rs:142-144
};

// Everything flow.rs derives from one infoset's rows, computed once (same op order as the reference
// recomputes it per call).
struct InfoView {
    int n;
    uint8_t edges[MAXA];
    float r[MAXA];      // profile.rs:31-33  regret = max(cum_regret, EPS)
    float rd;           // flow.rs:20-22     Σ r  (choices order, from 0.0)
    float w[MAXA];      // profile.rs:35-37
    float wsum;         // Σ w
    float denom;        // flow.rs:24-26     Σ w + smoothing
    float sw[MAXA];     // flow.rs:30-32     max(((w/τ)+β)/denom, ε)
    float z;            // Σ sw
};

template <class G>
inline InfoView view_of(const Profile<G>& p, uint32_t info) {
    InfoView v;
    uint8_t ed[MAX_BRANCH];
    v.n = G::choices(info, ed);
    float rd = 0.0f, ws = 0.0f;
    for (int a = 0; a < v.n; ++a) {
        v.edges[a] = ed[a];
        float cr = p.cum_regret(info, a, ed[a]);
        v.r[a] = cr > EPS ? cr : EPS;
        rd = rd + v.r[a];
        float cw = p.cum_weight(info, a);
        v.w[a] = cw > EPS ? cw : EPS;
        ws = ws + v.w[a];
    }
    v.rd = rd;
    v.wsum = ws;
    v.denom = ws + p.hyper.smoothing;
    float z = 0.0f;
    for (int a = 0; a < v.n; ++a) {
        float s = (v.w[a] / p.hyper.temperature + p.hyper.smoothing) / v.denom;
        v.sw[a] = s > p.hyper.curiosity ? s : p.hyper.curiosity;
        z = z + v.sw[a];
    }
    v.z = z;
    return v;
}
inline int action_of(const InfoView& v, uint8_t edge) {
    for (int a = 0; a < v.n; ++a) if (v.edges[a] == edge) return a;
    return -1;
}

template <class G>
struct Tree {  // state/tree.rs + petgraph::Graph adjacency (newest out-edge first)
    using State = typename G::State;
    int id = 0;
    std::vector<State> game;
    std::vector<uint32_t> info;
    std::vector<int> parent, head, next;  // head[n] = newest child, next[c] = previous sibling
    std::vector<uint8_t> incoming;
    int add(const State& s, uint32_t i, int par, uint8_t edge) {
        int idx = (int)game.size();
        game.push_back(s); info.push_back(i); parent.push_back(par); incoming.push_back(edge);
        head.push_back(-1); next.push_back(-1);
        if (par >= 0) { next[idx] = head[par]; head[par] = idx; }
        return idx;
    }
    int n() const { return (int)game.size(); }
    int width(int node) const { int k = 0; for (int c = head[node]; c >= 0; c = next[c]) ++k; return k; }
    int step(int node, uint8_t edge) const {  // node.rs:112-117
        for (int c = head[node]; c >= 0; c = next[c]) if (incoming[c] == edge) return c;
        return -1;
    }
};

// BATCHED fold (include/rbp.h RBP_FOLD_BATCHED; not a reference mode — the north-star's "allreduce of deltas" form):
// per epoch and infoset the Decisions of all trees are summed in a fixed blocked order (sequentially inside blocks
// of 128 trees, then over a rank's blocks, then over ranks) and every schedule is applied ONCE per row.
struct Partial {  // 12 words per infoset, the unit ranks exchange
    float dr[MAXA];
    float pay;
    uint32_t n;
    uint32_t na[MAXA];
    uint32_t pad[2];
};

struct Decisions {  // solver/decisions.rs:23-32
    uint32_t info;
    int tree;              // global tree id
    int n;                 // choices
    bool explored[MAXA];
    float regret[MAXA];    // only where explored
    float policy[MAXA];    // all choices
    float payoff;
};

template <class G>
struct Solver {
    using State = typename G::State;
    Profile<G> profile;
    int regret_sched = R_FLOORED, weight_sched = W_LINEAR, sampling = S_EXTERNAL;
    int batch = 1, threads = 1, fold_mode = FOLD_ORDERED, fold_block = 128;
    mutable std::shared_ptr<Pool> pool;  // rayon's persistent pool (created on first use)
    int world_rank = 0, world_size = 1;
    std::vector<uint32_t> info_order;  // all decision infosets of the game, ascending key: index space of Partial buffers
    Draw rng{0};
    // telemetry (metrics/mod.rs: nodes / infos counters)
    uint64_t nodes = 0, infos = 0, updates = 0;

    struct Leaf { uint8_t edge; State game; int head; };

    // sample/*.rs
    void sample(const Tree<G>& tree, int node, std::vector<Leaf>& br, int kind) const {
        if (br.empty() || kind == S_VANILLA) return;
        const State& g = tree.game[node];
        Turn p = G::turn(g);
        uint32_t info = tree.info[node];
        int walker = profile.walker();
        uint32_t epoch = (uint32_t)profile.epochs;
        if ((int)p == walker) {
            if (kind == S_EXTERNAL || kind == S_TARGETED) return;  // external.rs:36, targeted.rs:27
            if (kind == S_PLURIBUS) {        // pluribus.rs:85-93
                if (profile.epochs < profile.hyper.prune_warmup) return;
                Philox4 c = rng.at(epoch, (uint32_t)tree.id, info, TAG_COIN);
                if (draw_unit(c.r[0]) < profile.hyper.prune_explore) return;
            }
            InfoView v = view_of(profile, info);
            std::vector<Leaf> kept;
            for (const Leaf& l : br) {  // pruning.rs:58-64, pluribus.rs:94-99
                int a = action_of(v, l.edge);
                bool keep = profile.cum_regret(info, a, l.edge) > profile.hyper.prune_threshold;
                if (kind == S_PLURIBUS && G::turn(l.game) == TURN_TERMINAL) keep = true;
                if (keep) kept.push_back(l);
            }
            if (!kept.empty()) br.swap(kept);
            return;
        }
        Philox4 c = rng.at(epoch, (uint32_t)tree.id, info, TAG_NODE);
        int pick;
        if (p == TURN_CHANCE) {
            pick = (int)draw_range(c.r[0], (uint32_t)br.size());  // sample/mod.rs:68-82
        } else {  // external.rs:42-64: weights = sampling_distribution density, floored at EPSILON
            InfoView v = view_of(profile, info);
            float w[MAX_BRANCH];
            for (size_t i = 0; i < br.size(); ++i) {
                int a = action_of(v, br[i].edge);
                if (kind == S_TARGETED) {  // targeted.rs:34-62: iterated distribution floored at `curiosity`
                    float sg = v.r[a] / v.rd;
                    w[i] = sg > profile.hyper.curiosity ? sg : profile.hyper.curiosity;
                    continue;
                }
                float q = v.sw[a] / v.z;
                w[i] = q > EPS ? q : EPS;
            }
            pick = draw_weighted(c.r[0], w, (int)br.size());
        }
        Leaf chosen = br[pick];
        br.clear();
        br.push_back(chosen);
    }

    bool block_chance = false;  // subgame/src/world/encoder.rs:97-106: a subgame tree does not expand chance nodes (no frontier here)
    std::vector<Leaf> branches(const Tree<G>& tree, int node) const {
        uint8_t ed[MAX_BRANCH];
        if (block_chance && G::turn(tree.game[node]) == TURN_CHANCE) return {};
        int n = G::branches(tree.game[node], ed);
        std::vector<Leaf> out;
        for (int i = 0; i < n; ++i) out.push_back(Leaf{ed[i], G::apply(tree.game[node], ed[i]), node});
        return out;
    }

    // builder.rs:74-161
    Tree<G> build(const State& root, int id, int kind) const {
        Tree<G> tree;
        tree.id = id;
        tree.add(root, G::info_key(root), -1, 0);
        std::vector<Leaf> todo = branches(tree, 0);
        sample(tree, 0, todo, kind);
        while (!todo.empty()) {
            Leaf leaf = todo.back();
            todo.pop_back();
            int node = tree.add(leaf.game, G::info_key(leaf.game), leaf.head, leaf.edge);
            std::vector<Leaf> kids = branches(tree, node);
            sample(tree, node, kids, kind);
            for (const Leaf& k : kids) todo.push_back(k);
        }
        return tree;
    }

    // flow.rs:166-174 ancestor_reach (node.rs:153-161 decisions(): upward, chance skipped)
    float ancestor_reach(const Tree<G>& tree, int root, int walker) const {
        float cf = 1.0f, sm = 1.0f;
        int node = root;
        while (tree.parent[node] >= 0) {
            int par = tree.parent[node];
            uint8_t edge = tree.incoming[node];
            Turn t = G::turn(tree.game[par]);
            if (t != TURN_CHANCE && (int)t != walker) {
                InfoView v = view_of(profile, tree.info[par]);
                int a = action_of(v, edge);
                cf = cf * (v.r[a] / v.rd);
                sm = sm * (v.sw[a] / v.z);
            }
            node = par;
        }
        return cf / sm;
    }
    // nash.rs:50-79 terminal_value of a childless node: the game's payoff at a terminal; at a chance node that was not expanded (a
    // subgame tree stops at chance nodes, subgame/src/world/encoder.rs:97-106) `frontier_payoff` = cum_payoff(info, first choice) of the
    // nearest non-chance ancestor's infoset — V(I) as stored, taken for `hero` without a change of sign, as the reference does
    float terminal_value(const Tree<G>& tree, int node, int hero) const {
        const Turn t = G::turn(tree.game[node]);
        if (t == TURN_TERMINAL) return G::payoff(tree.game[node], hero);
        int a = node;
        if (t == TURN_CHANCE) { do { a = tree.parent[a]; } while (a >= 0 && G::turn(tree.game[a]) == TURN_CHANCE); }
        if (a < 0) return 0.0f;  // "chance node cannot be root of game"
        return profile.cum_payoff(tree.info[a], 0);
    }
    // flow.rs:182-216 recursed_value
    float recursed_value(const Tree<G>& tree, int hero, int node, float rel, float smp) const {
        if (tree.head[node] < 0) return rel / smp * terminal_value(tree, node, hero);
        Turn t = G::turn(tree.game[node]);
        bool chance = t == TURN_CHANCE, walk = (int)t == profile.walker();
        InfoView v{};
        if (!chance) v = view_of(profile, tree.info[node]);
        float sum = 0.0f;
        for (int c = tree.head[node]; c >= 0; c = tree.next[c]) {
            float r2 = rel, s2 = smp;
            if (!chance) {
                int a = action_of(v, tree.incoming[c]);
                r2 = rel * (v.r[a] / v.rd);
                if (!walk) s2 = smp * (v.sw[a] / v.z);
            } else {
                r2 = rel * 1.0f; s2 = smp * 1.0f;
            }
            sum = sum + recursed_value(tree, hero, c, r2, s2);
        }
        return sum;
    }
    // flow.rs:64-87 dfs + solver.rs:296-305 update_vector
    Decisions update_vector(const Tree<G>& tree, const std::vector<int>& span) const {
        Decisions d{};
        int head = span[0];
        d.info = tree.info[head];
        InfoView v = view_of(profile, d.info);
        d.n = v.n;
        for (int a = 0; a < v.n; ++a) d.policy[a] = v.r[a] / v.rd;  // profile.rs:47-51
        int hero = (int)G::turn(tree.game[head]);
        float payoff = 0.0f;
        for (int root : span) {
            float reach = ancestor_reach(tree, root, profile.walker());
            float val[MAXA]; int act[MAXA]; int k = 0;
            for (int c = tree.head[root]; c >= 0; c = tree.next[c]) {
                act[k] = action_of(v, tree.incoming[c]);
                val[k] = reach * recursed_value(tree, hero, c, 1.0f, 1.0f);
                ++k;
            }
            float ev = 0.0f;
            for (int i = 0; i < k; ++i) ev = ev + v.r[act[i]] / v.rd * val[i];
            payoff += ev;
            for (int i = 0; i < k; ++i) {
                int a = act[i];
                if (!d.explored[a]) { d.explored[a] = true; d.regret[a] = 0.0f; }
                d.regret[a] += val[i] - ev;
            }
        }
        d.payoff = payoff;
        return d;
    }

    // solver.rs:263-275 record_infosets + tree.rs:88-97 partition (first-seen order; rows of one
    // tree are disjoint so intra-tree order is immaterial)
    void tree_decisions(const Tree<G>& tree, std::vector<Decisions>& out) const {
        std::vector<uint32_t> keys;
        std::vector<std::vector<int>> spans;
        for (int n = 0; n < tree.n(); ++n) {
            if (tree.head[n] < 0) continue;
            uint32_t k = tree.info[n];
            size_t i = 0;
            for (; i < keys.size(); ++i) if (keys[i] == k) break;
            if (i == keys.size()) { keys.push_back(k); spans.emplace_back(); }
            spans[i].push_back(n);
        }
        int walker = profile.walker();
        for (size_t i = 0; i < keys.size(); ++i) {
            if ((int)G::turn(tree.game[spans[i][0]]) != walker) continue;
            out.push_back(update_vector(tree, spans[i]));
            out.back().tree = tree.id;
        }
    }

    State root_of(int tree_id) const {
        return G::root(rng.at((uint32_t)profile.epochs, (uint32_t)tree_id, 0xFFFFFFFFu, TAG_ROOT));
    }

    // solver.rs:225-240 batch (rayon over trees; order-preserving collect)
    std::vector<Decisions> run_batch(uint64_t* node_count) const {
        int T = threads < 1 ? 1 : threads;
        if (T > batch) T = batch;
        if (!pool || pool->size() != T) pool = std::make_shared<Pool>(T);
        constexpr int kChunk = 256;  // trees per claimed chunk (small-game trees are ~20 nodes)
        const int chunks = (batch + kChunk - 1) / kChunk;
        std::vector<std::vector<Decisions>> parts(chunks);
        std::vector<uint64_t> ncount(chunks, 0);
        pool->run(chunks, [&](int c, int) {
            for (int i = c * kChunk; i < std::min(batch, (c + 1) * kChunk); ++i) {
                const int id = world_rank * batch + i;
                Tree<G> tree = build(root_of(id), id, sampling);
                ncount[c] += tree.n();
                tree_decisions(tree, parts[c]);
            }
        });
        std::vector<Decisions> all;
        for (int c = 0; c < chunks; ++c) {
            all.insert(all.end(), parts[c].begin(), parts[c].end());
            *node_count += ncount[c];
        }
        return all;
    }

    // solver.rs:143-192
    void apply(const Decisions& d) {
        uint64_t epoch = profile.epochs;
        // every update reads `profile().cum_*` (for a WorldProfile: the blueprint until the edge exists locally) and only then takes
        // `storage().mut_*`, which is what creates the local edge (from the blueprint's warmstart)
        uint8_t ed[MAX_BRANCH];
        G::choices(d.info, ed);
        for (int a = 0; a < d.n; ++a) {
            if (!d.explored[a]) continue;
            const float total = profile.cum_regret(d.info, a, ed[a]);
            profile.mut_row(d.info, a).regret = regret_gain(regret_sched, total, d.regret[a], epoch, profile.hyper);
            ++updates;
        }
        for (int a = 0; a < d.n; ++a) {
            const float total = profile.cum_weight(d.info, a);
            profile.mut_row(d.info, a).weight = weight_learn(weight_sched, total, d.policy[a], epoch);
        }
        for (int a = 0; a < d.n; ++a) {
            const uint32_t n = profile.cum_visits(d.info, a);
            Encounter& e = profile.mut_row(d.info, a);
            e.payoff += (d.payoff - e.payoff) / (float)(n + 1);
        }
        for (int a = 0; a < d.n; ++a) profile.mut_row(d.info, a).visits += 1;
    }

    void ensure_info_order() {
        if (!info_order.empty()) return;
        Tree<G> tree = build(G::exploitability_root(), 0, S_VANILLA);
        for (int n = 0; n < tree.n(); ++n) {
            Turn t = G::turn(tree.game[n]);
            if (tree.head[n] >= 0 && (t == TURN_P0 || t == TURN_P1)) info_order.push_back(tree.info[n]);
        }
        std::sort(info_order.begin(), info_order.end());
        info_order.erase(std::unique(info_order.begin(), info_order.end()), info_order.end());
    }
    int info_index(uint32_t key) const {
        return (int)(std::lower_bound(info_order.begin(), info_order.end(), key) - info_order.begin());
    }
    // this rank's blocked partial sums for the current epoch (K1 + block/rank reduction on the GPU)
    std::vector<Partial> sample_partials() {
        ensure_info_order();
        uint64_t nc = 0;
        std::vector<Decisions> all = run_batch(&nc);
        nodes += nc;
        infos += all.size();
        const size_t I = info_order.size();
        std::vector<Partial> rank(I), block(I);
        std::vector<uint8_t> touched(I, 0);
        auto zero = [](Partial& p) { std::memset(&p, 0, sizeof p); };
        for (auto& p : rank) zero(p);
        for (auto& p : block) zero(p);
        auto flush = [&]() {  // rank += block, every infoset (dense adds, exactly what the device kernel does)
            for (size_t x = 0; x < I; ++x) {
                for (int a = 0; a < MAXA; ++a) { rank[x].dr[a] = rank[x].dr[a] + block[x].dr[a]; rank[x].na[a] += block[x].na[a]; }
                rank[x].pay = rank[x].pay + block[x].pay;
                rank[x].n += block[x].n;
                zero(block[x]);
            }
        };
        int cur = 0;
        for (const Decisions& d : all) {
            int b = (d.tree - world_rank * batch) / fold_block;
            while (cur < b) { flush(); ++cur; }
            Partial& k = block[info_index(d.info)];
            for (int a = 0; a < d.n; ++a)
                if (d.explored[a]) { k.dr[a] = k.dr[a] + d.regret[a]; k.na[a] += 1; }
            k.pay = k.pay + d.payoff;
            k.n += 1;
        }
        const int nblk = (batch + fold_block - 1) / fold_block;
        while (cur < nblk) { flush(); ++cur; }
        return rank;
    }
    // sum the ranks' partials in rank order and apply each schedule once per row; advances the epoch
    void fold_gathered(const Partial* gathered, int world) {
        ensure_info_order();
        const size_t I = info_order.size();
        const uint64_t epoch = profile.epochs;
        for (size_t x = 0; x < I; ++x) {
            Partial t;
            std::memset(&t, 0, sizeof t);
            for (int r = 0; r < world; ++r) {
                const Partial& g = gathered[(size_t)r * I + x];
                for (int a = 0; a < MAXA; ++a) { t.dr[a] = t.dr[a] + g.dr[a]; t.na[a] += g.na[a]; }
                t.pay = t.pay + g.pay;
                t.n += g.n;
            }
            if (t.n == 0) continue;
            const uint32_t info = info_order[x];
            InfoView v = view_of(profile, info);  // regret matching on the pre-fold regrets
            for (int a = 0; a < v.n; ++a) {
                Encounter& e = profile.mut_row(info, a);
                if (t.na[a] > 0) {
                    e.regret = regret_gain(regret_sched, e.regret, t.dr[a], epoch, profile.hyper);
                    updates += gathered[(size_t)(world > 1 ? world_rank : 0) * I + x].na[a];  // telemetry counts this rank's own trees
                }
                e.weight = weight_learn(weight_sched, e.weight, (float)t.n * (v.r[a] / v.rd), epoch);
                const float mean = t.pay / (float)t.n;
                e.payoff += (mean - e.payoff) * (float)t.n / (float)(e.visits + t.n);
                e.visits += t.n;
            }
        }
        profile.epochs += 1;
    }

    // solver.rs:96-105 step
    void step() {
        if (fold_mode == FOLD_BATCHED) {
            std::vector<Partial> mine = sample_partials();
            fold_gathered(mine.data(), 1);
            return;
        }
        uint64_t nc = 0;
        std::vector<Decisions> all = run_batch(&nc);
        nodes += nc;
        infos += all.size();
        for (const Decisions& d : all) apply(d);
        profile.epochs += 1;  // book.rs:138-140
    }

    // ── exploitability (solver.rs:327-338, nash.rs:31-193) ──
    struct AvgView { int n; uint8_t edges[MAXA]; float p[MAXA]; };
    AvgView averaged(uint32_t info) const {  // profile.rs:41-45
        AvgView v;
        uint8_t ed[MAX_BRANCH];
        v.n = G::choices(info, ed);
        float w[MAXA], sum = 0.0f;
        for (int a = 0; a < v.n; ++a) {
            v.edges[a] = ed[a];
            float cw = profile.cum_weight(info, a);
            w[a] = cw > EPS ? cw : EPS;
            sum = sum + w[a];
        }
        for (int a = 0; a < v.n; ++a) v.p[a] = w[a] / sum;
        return v;
    }
    float averaged_policy(uint32_t info, uint8_t edge) const {
        AvgView v = averaged(info);
        for (int a = 0; a < v.n; ++a) if (v.edges[a] == edge) return v.p[a];
        return 0.0f;
    }
    // nash.rs:103-133 subgamed_payoff
    float subgamed(const Tree<G>& tree, int node, int hero, const std::unordered_map<uint32_t, uint8_t>* br) const {
        int n = tree.width(node);
        if (n == 0) return G::payoff(tree.game[node], hero);
        Turn t = G::turn(tree.game[node]);
        if (t == TURN_CHANCE) {
            float sum = 0.0f;
            for (int c = tree.head[node]; c >= 0; c = tree.next[c]) sum = sum + subgamed(tree, c, hero, br);
            return sum / (float)n;
        }
        if ((int)t == hero && br) {
            uint8_t e = br->at(tree.info[node]);
            return subgamed(tree, tree.step(node, e), hero, br);
        }
        float sum = 0.0f;
        for (int c = tree.head[node]; c >= 0; c = tree.next[c])
            sum = sum + averaged_policy(tree.info[node], tree.incoming[c]) * subgamed(tree, c, hero, br);
        return sum;
    }
    // nash.rs:140-145 external_reach (upward product over non-hero decision ancestors)
    float external_reach(const Tree<G>& tree, int node, int hero) const {
        float prod = 1.0f;
        while (tree.parent[node] >= 0) {
            int par = tree.parent[node];
            Turn t = G::turn(tree.game[par]);
            if (t != TURN_CHANCE && (int)t != hero) prod = prod * averaged_policy(tree.info[par], tree.incoming[node]);
            node = par;
        }
        return prod;
    }
    struct ExplStats { int nodes, terminals, infosets; };
    float exploitability(ExplStats* st = nullptr) const {
        Tree<G> tree = build(G::exploitability_root(), 0, S_VANILLA);
        std::vector<uint32_t> keys;
        std::vector<std::vector<int>> spans;
        std::unordered_map<uint32_t, size_t> idx;
        int terminals = 0, decision_infos = 0;
        for (int n = 0; n < tree.n(); ++n) {
            if (tree.head[n] < 0) { ++terminals; continue; }
            uint32_t k = tree.info[n];
            auto it = idx.find(k);
            if (it == idx.end()) { idx[k] = keys.size(); keys.push_back(k); spans.emplace_back(); spans.back().push_back(n); }
            else spans[it->second].push_back(n);
        }
        for (size_t i = 0; i < keys.size(); ++i) {
            Turn t = G::turn(tree.game[spans[i][0]]);
            if (t == TURN_P0 || t == TURN_P1) ++decision_infos;
        }
        if (st) { st->nodes = tree.n(); st->terminals = terminals; st->infosets = decision_infos; }
        float total = 0.0f;
        for (int hero = 0; hero < 2; ++hero) {
            std::unordered_map<uint32_t, uint8_t> br;
            for (size_t i = 0; i < keys.size(); ++i) {
                if ((int)G::turn(tree.game[spans[i][0]]) != hero) continue;
                uint8_t ed[MAX_BRANCH];
                int n = G::choices(keys[i], ed);
                float best = 0.0f; int besta = -1;
                for (int a = 0; a < n; ++a) {  // nash.rs:171-193
                    float sum = 0.0f;
                    for (int node : spans[i]) {
                        int c = tree.step(node, ed[a]);
                        if (c < 0) continue;
                        sum = sum + external_reach(tree, c, hero) * subgamed(tree, c, hero, nullptr);
                    }
                    if (besta < 0 || !(sum < best)) { best = sum; besta = a; }  // max_by keeps the LAST maximum
                }
                br[keys[i]] = ed[besta];
            }
            total = total + subgamed(tree, 0, hero, &br);
        }
        return total / 2.0f;
    }
};

}  // namespace orc
