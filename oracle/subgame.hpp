// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).  CPU restatement of the reference's safe subgame solver on the small
// games: `WorldSolver` (crates/subgame/src/world/solver.rs:33-146) = `SubGameSolver` with `origin = None`
// (crates/subgame/src/solver.rs:46-146: frontier detection off, `DepthView` / `DepthInfo::Game` / `DepthEdge::Game` are pass-through).
// Depth-limited frontiers (`DepthGame` with an origin, `Continuation` picks, `Payoffs`) are NOT restated.
//
// Follows (all under /root/reference/crates):
//   subgame/src/world/partition.rs:27-53     Posterior -> Belief (quantile worlds, highest reach first)
//   subgame/src/world/belief.rs:33-52        world(), remember() (an empty belief remembers every secret)
//   subgame/src/world/solver.rs:118-146      step: sample a world, restrict the entry, one ExternalSampling tree, SummedRegret +
//                                            LinearWeight updates, advance
//   subgame/src/world/encoder.rs:60-106      infosets tagged with the world (a table per world here); chance and terminal nodes
//                                            do not expand
//   subgame/src/world/profile.rs:62-181      local rows over a frozen blueprint (restated inside Profile, mccfr.hpp)
//   kuhn/src/encoder.rs:47-66, leduc/src/encoder.rs:48-70   WorldRestrict: the first card in Card::ALL order that is free and whose
//                                            rank the belief puts in the world; the observed state if there is none
//   subgame/src/world/solver.rs:148-191      Harvest
// Parity status: the reference draws the world from the thread RNG (`rand::rng()`), unseeded — "parity unpinned"; here it is the
// Philox contract of rng.hpp with counter (step, 0, 0xFFFFFFFE, TAG_WORLD) and `weighted()`.  The tree's draws use tree id = world.
// Pinned to the reference's own subgame assertions (kuhn/src/solver.rs `subgame_nash`: K|B and K|XB call > 0.90 averaged over the
// worlds after 2^16 steps on a 2^18-epoch blueprint; `restrict_produces_valid_deals`) in tests/test_oracle_subgame.py.
#pragma once
#include <algorithm>
#include <vector>

#include "mccfr.hpp"

namespace orc {

enum : uint32_t { TAG_WORLD = 5 };
constexpr int MAX_WORLDS = 8;

// world/partition.rs:27-53.  `reach[i]` is the posterior mass of secret i (secrets in ascending order, as the BTreeMap iterates);
// out: the world of every secret and the per-world weights.
inline void partition(const float* reach, int n, int W, int32_t* world_of, float* weights) {
    float total = 0.0f;
    for (int i = 0; i < n; ++i) total += reach[i];  // Posterior::total: .sum::<f32>()
    for (int w = 0; w < W; ++w) weights[w] = 0.0f;
    if (total <= 0.0f) {
        for (int i = 0; i < n; ++i) world_of[i] = 0;
        for (int w = 0; w < W; ++w) weights[w] = 1.0f / (float)W;
        return;
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return reach[a] > reach[b]; });  // sort_by(b.1.partial_cmp(a.1)): stable, descending
    const float segment = total / (float)W;
    int index = 0;
    float bucket = 0.0f, accumulated = 0.0f;
    for (int i : order) {
        bucket += reach[i];
        accumulated += reach[i];
        world_of[i] = index;
        if (accumulated >= segment * (float)(index + 1) && index < W - 1) {
            weights[index] = bucket / total;
            index += 1;
            bucket = 0.0f;
        }
    }
    weights[index] = bucket / total;
}

template <class G>
struct SubSolver {
    using Game = G;
    using State = typename G::State;
    const Solver<G>* blueprint;
    int worlds = 1, external = 1;
    std::vector<Solver<G>> local;   // one table per world: WorldInfo(world, info) keys
    float weights[MAX_WORLDS];
    int32_t world_of_rank[3];       // Belief::members (secret = Rank); all -1 = empty belief
    bool empty_belief = true;
    State observed;                 // CfrRecall::game()
    uint64_t t = 0;                 // WorldProfile::t
    Draw rng{0};
    uint64_t drawn[MAX_WORLDS] = {};

    SubSolver(const Solver<G>* bp, int external_, int W, const int32_t* world_of, const float* w, const State& obs, uint64_t seed)
        : blueprint(bp), worlds(W), external(external_), observed(obs), rng{seed} {
        for (int r = 0; r < 3; ++r) { world_of_rank[r] = world_of ? world_of[r] : -1; if (world_of_rank[r] >= 0) empty_belief = false; }
        for (int i = 0; i < W; ++i) weights[i] = w[i];
        local.resize(W);
        for (Solver<G>& s : local) {
            s.regret_sched = R_SUMMED; s.weight_sched = W_LINEAR; s.sampling = S_EXTERNAL;  // world/solver.rs:100-102
            s.batch = 1; s.threads = 1; s.block_chance = true;
            s.profile.hyper = bp->profile.hyper;          // temperature / smoothing / curiosity read through to the blueprint
            s.profile.blueprint = &bp->profile;
            s.rng = rng;
        }
    }
    bool remember(int rank, int world) const { return empty_belief || world_of_rank[rank] == world; }  // belief.rs:44-46
    // kuhn/src/encoder.rs:47-66, leduc/src/encoder.rs:48-70
    State restrict(int world) const {
        if (external != 0 && external != 1) return observed;
        for (uint8_t c = 0; c < 6; ++c) {
            if (c == observed.hole[1 - external]) continue;
            if (G::board_is(observed, c)) continue;
            State s = observed;
            s.hole[external] = c;
            if (remember(G::rank(c), world)) return s;
        }
        return observed;
    }
    int draw_world() const {  // world/solver.rs:70-76 WeightedIndex over the belief weights
        const Philox4 p = rng.at((uint32_t)t, 0u, 0xFFFFFFFEu, TAG_WORLD);
        return draw_weighted(p.r[0], weights, worlds);
    }
    void step() {  // world/solver.rs:118-146
        const int w = draw_world();
        drawn[w] += 1;
        Solver<G>& s = local[w];
        s.profile.epochs = t;                                  // WorldProfile::t: walker = t % 2, LinearWeight's t
        const State entry = restrict(w);
        Tree<G> tree = s.build(entry, /*tree id*/ w, S_EXTERNAL);
        s.nodes += tree.n();
        std::vector<Decisions> all;
        s.tree_decisions(tree, all);
        s.infos += all.size();
        for (const Decisions& d : all) s.apply(d);
        t += 1;
    }
    // world/profile.rs:147-155 sum_regret
    float sum_regret() const {
        float sum = 0.0f;
        for (const Solver<G>& s : local)
            for (const auto& kv : s.profile.rows)
                for (int a = 0; a < kv.second.n; ++a)
                    if (kv.second.present[a]) sum += kv.second.e[a].regret > 0.0f ? kv.second.e[a].regret : 0.0f;
        return sum / (float)(t > 1 ? t : 1);
    }
};

}  // namespace orc
