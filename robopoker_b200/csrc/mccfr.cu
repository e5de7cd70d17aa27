// mccfr.cu — external-sampling MCCFR for flat games on sm_100a, and its C ABI (include/rbp.h).
//
// One epoch (= one `Solver::step`, crates/mccfr/src/solver/solver.rs:96-105) is two kernels:
//
//   mccfr_sample_kernel   one thread per tree.  Restates `Solver::batch` (solver.rs:225-240): deal the root,
//                         grow the sampled tree in the reference's LIFO order (solver/builder.rs:74-161) with
//                         External / Prunable / Pluribus sampling (sample/*.rs), then for every walker infoset
//                         the fused regret+value pass `CfrFlow::dfs` (strategy/flow.rs:64-87,166-216) with the
//                         reference's exact f32 operation order (compiled -fmad=false), producing `Decisions`
//                         records.  The block then stably multisplits its records by infoset (smem lane bitmaps
//                         → ranks) so that each infoset's records are contiguous and in tree order.
//   mccfr_fold_kernel     one warp per infoset, one lane per (row, field) chain.  Restates the serial fold of
//                         `update_regret/weight/payoff/visits` (solver.rs:143-192): every Decisions is applied
//                         in tree order with the schedule of regret/*.rs and policy/*.rs.  Rows are 16-byte
//                         `Encounter`s resident in HBM (L2-resident for the validation games).
//
// Because every f32 operation and its order match the oracle's, results are bit-identical, so the sampled
// trajectories of GPU and oracle never diverge.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "comm.hpp"
#include "common.cuh"
#include "flat_game.hpp"

namespace rbp {

std::atomic<uint64_t> g_launches{0};
static std::mutex g_err_mu;
static std::string g_err;
void set_last_error(const std::string& msg) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = msg;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s:%d: %s -> %s", file, line, what, cudaGetErrorString(e));
    set_last_error(buf);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? RBP_ERR_NO_DEVICE : RBP_ERR_CUDA;
}

// ───────────────────────────── device-side views ─────────────────────────────
constexpr int kTreesPerBlock = 128;
constexpr int kMaxTreeNodes = 48;   // local sampled-tree capacity (Leduc needs 29)
constexpr int kMaxDepth = 16;
constexpr int kMaxRecords = 8;      // walker infosets per sampled tree (Leduc needs 7)
constexpr int kMaxInfos = 512;      // smem-resident infoset views

struct DevGame {
    const FlatNode* nodes;
    const float* payoff1;
    const int32_t* root_table;
    const int32_t* info_row;
    const uint8_t* info_actions;
    const uint8_t* info_player;
    const int32_t* parent;
    const int32_t* level_start;
    const int32_t* span_start;
    const int32_t* span_nodes;
    int n_nodes, n_infos, n_rows, n_levels, deck;
};

struct Scratch {        // per-epoch Decisions, block-sorted by infoset
    float* pay;         // [nblk * cap]
    float* dr;          // [kMaxActions][nblk * cap]
    uint8_t* mask;      // [nblk * cap] explored bits
    int32_t* m_off;     // [I][nblk]
    int32_t* m_cnt;     // [I][nblk]
    unsigned long long* counters;  // nodes, infos, updates
    int nblk, cap;      // cap = records per block
    // BATCHED fold: per-block partial sums, [field][I][nblk]; fields 0-3 regret deltas, 4 payoff / explored counts
    float* bp_f;        // [5][I][nblk]
    uint32_t* bp_na;    // [4][I][nblk]
};

// BATCHED fold exchange unit, 12 words per infoset (include/rbp.h rbp_solver_delta_buffer)
struct Partial {
    float dr[kMaxActions];
    float pay;
    uint32_t n;
    uint32_t na[kMaxActions];
    uint32_t pad[2];
};

struct EpochArgs {
    uint32_t seed_lo, seed_hi, epoch;
    int walker, batch, tree_base, sampling;
    rbp_hyper_t hyper;
    int regret_sched, weight_sched, fold_mode;
    float t;                // epoch as f32
    float disc_pos, disc_neg;  // DiscountedRegret x = t^1.5, t^0.5 (host libm, regret/discounted.rs:33,37)
    int entry_plus1;        // subgame: flat node the tree starts from, + 1 (0 = the game's own root rule)
};

__device__ __forceinline__ float fmax_ref(float a, float b) { return a > b ? a : b; }

// ───────────────────────────── K1: sample + value + multisplit ─────────────────────────────
__global__ void __launch_bounds__(kTreesPerBlock)
mccfr_sample_kernel(DevGame g, const rbp_encounter_t* __restrict__ table, const float* __restrict__ fb_weight, Scratch sc, EpochArgs ep) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int I = g.n_infos;
    float* s_sigma = reinterpret_cast<float*>(smem_raw);        // [I][4]  regret-matching policy
    float* s_q = s_sigma + I * kMaxActions;                     // [I][4]  sampling distribution
    float* s_cumr = s_q + I * kMaxActions;                      // [I][4]  raw cumulative regret (pruning)
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_cumr + I * kMaxActions);  // [4][I] lane bitmaps
    uint32_t* s_woff = s_bits + 4 * I;                          // [4][I] warp offsets
    uint32_t* s_cnt = s_woff + 4 * I;                           // [I]
    uint32_t* s_off = s_cnt + I;                                // [I]
    float* s_pay = reinterpret_cast<float*>(s_off + I);         // [I]  V(I) = cum_payoff(info, first choice) (subgame chance leaves)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // infoset views (strategy/profile.rs:31-51, flow.rs:20-59), once per block instead of once per node visit
    for (int x = tid; x < I; x += kTreesPerBlock) {
        const int A = g.info_actions[x], row = g.info_row[x];
        float r[kMaxActions], w[kMaxActions], sw[kMaxActions];
        float rd = 0.0f, ws = 0.0f;
        for (int a = 0; a < A; ++a) {
            rbp_encounter_t e = table[row + a];
            s_cumr[x * kMaxActions + a] = e.regret;
            r[a] = fmax_ref(e.regret, kEps);
            rd = rd + r[a];
            // subgame/src/world/profile.rs:118-138: an edge the subgame never wrote (visits 0) reads the blueprint's weight and payoff
            if (fb_weight != nullptr && e.visits == 0u) { e.weight = fb_weight[row + a]; e.payoff = fb_weight[g.n_rows + row + a]; }
            if (a == 0) s_pay[x] = e.payoff;  // nash.rs:50-56 frontier_payoff: the first choice's stored V(I)
            w[a] = fmax_ref(e.weight, kEps);
            ws = ws + w[a];
        }
        const float denom = ws + ep.hyper.smoothing;
        float z = 0.0f;
        for (int a = 0; a < A; ++a) {
            float s = (w[a] / ep.hyper.temperature + ep.hyper.smoothing) / denom;
            sw[a] = fmax_ref(s, ep.hyper.curiosity);
            z = z + sw[a];
        }
        for (int a = 0; a < A; ++a) {
            s_sigma[x * kMaxActions + a] = r[a] / rd;
            s_q[x * kMaxActions + a] = sw[a] / z;
        }
        for (int wv = 0; wv < 4; ++wv) s_bits[wv * I + x] = 0u;
    }
    __syncthreads();

    const int local = blockIdx.x * kTreesPerBlock + tid;
    const bool active = local < ep.batch;
    const uint32_t tree = (uint32_t)(ep.tree_base + local);

    // Decisions of this tree
    int nrec = 0;
    int16_t rec_info[kMaxRecords];
    uint8_t rec_mask[kMaxRecords];
    float rec_pay[kMaxRecords];
    float rec_dr[kMaxRecords][kMaxActions];
    int ln = 0;

    if (active) {
        // sampled tree, reference node order (petgraph-style adjacency: head = newest child)
        int16_t l_flat[kMaxTreeNodes];
        int8_t l_parent[kMaxTreeNodes], l_act[kMaxTreeNodes], l_head[kMaxTreeNodes], l_next[kMaxTreeNodes];
        int16_t t_flat[kMaxTreeNodes];
        int8_t t_parent[kMaxTreeNodes], t_act[kMaxTreeNodes];
        int tn = 0;

        // `CfrGame::root()` (kuhn/leduc game.rs root()): two-swap Fisher-Yates of the identity deck
        {
            if (ep.entry_plus1 > 0) {
                l_flat[0] = (int16_t)(ep.entry_plus1 - 1);  // subgame: `WorldRestrict::restrict`ed entry state (world/solver.rs:121-125)
            } else if (g.deck > 1) {
                Philox4 pr = philox4x32_10(ep.epoch, tree, 0xFFFFFFFFu, TAG_ROOT, ep.seed_lo, ep.seed_hi);
                uint32_t i = draw_range(pr.r[0], (uint32_t)g.deck);
                uint32_t j = 1u + draw_range(pr.r[1], (uint32_t)g.deck - 1u);
                uint32_t c0 = i, c1 = (j == i) ? 0u : j;
                l_flat[0] = (int16_t)g.root_table[c0 * g.deck + c1];
            } else {
                l_flat[0] = (int16_t)g.root_table[0];  // games without a deal (roshambo game.rs:11-13)
            }
            l_parent[0] = -1; l_act[0] = 0; l_head[0] = -1; l_next[0] = -1;
            ln = 1;
        }
        // S::sample + encoder.branches for local node k (sample/{external,pruning,pluribus}.rs)
        auto expand = [&](int k) {
            const FlatNode nd = g.nodes[l_flat[k]];
            const int n = nd.n_child;
            if (n == 0) return;
            if (ep.entry_plus1 > 0 && nd.turn == TURN_CHANCE) return;  // subgame/src/world/encoder.rs:97-106: a subgame tree stops at chance nodes
            if (nd.turn == ep.walker) {
                uint32_t keep = (1u << n) - 1u;
                bool prune = ep.sampling == RBP_SAMPLING_PRUNABLE;
                if (ep.sampling == RBP_SAMPLING_PLURIBUS && ep.epoch >= ep.hyper.prune_warmup) {
                    Philox4 pc = philox4x32_10(ep.epoch, tree, nd.info_key, TAG_COIN, ep.seed_lo, ep.seed_hi);
                    prune = !(draw_unit(pc.r[0]) < ep.hyper.prune_explore);
                }
                if (prune) {
                    uint32_t kept = 0;
                    for (int a = 0; a < n; ++a) {
                        bool ok = s_cumr[nd.info * kMaxActions + a] > ep.hyper.prune_threshold;
                        if (ep.sampling == RBP_SAMPLING_PLURIBUS && g.nodes[nd.first_child + a].turn == TURN_TERMINAL) ok = true;
                        if (ok) kept |= 1u << a;
                    }
                    if (kept) keep = kept;
                }
                for (int a = 0; a < n; ++a)
                    if (keep >> a & 1u) { t_flat[tn] = (int16_t)(nd.first_child + a); t_parent[tn] = (int8_t)k; t_act[tn] = (int8_t)a; ++tn; }
                return;
            }
            Philox4 p = philox4x32_10(ep.epoch, tree, nd.info_key, TAG_NODE, ep.seed_lo, ep.seed_hi);
            int pick;
            if (nd.turn == TURN_CHANCE) {
                pick = (int)draw_range(p.r[0], (uint32_t)n);  // sample/mod.rs:68-82
            } else {  // sample/external.rs:42-64
                // external.rs:42-64: sampling distribution floored at EPSILON; targeted.rs:34-62: iterated distribution
                // floored at `curiosity`
                const bool targeted = ep.sampling == RBP_SAMPLING_TARGETED;
                const float* q = (targeted ? s_sigma : s_q) + nd.info * kMaxActions;
                const float lo = targeted ? ep.hyper.curiosity : kEps;
                float total = 0.0f;
                for (int a = 0; a < n; ++a) total = total + fmax_ref(q[a], lo);
                const float x = draw_unit(p.r[0]) * total;
                float cum = 0.0f;
                pick = n - 1;
                for (int a = 0; a < n; ++a) {
                    cum = cum + fmax_ref(q[a], lo);
                    if (x < cum) { pick = a; break; }
                }
            }
            t_flat[tn] = (int16_t)(nd.first_child + pick); t_parent[tn] = (int8_t)k; t_act[tn] = (int8_t)pick; ++tn;
        };
        expand(0);
        while (tn > 0) {  // builder.rs:141-160: pop the newest branch
            --tn;
            const int k = ln++;
            const int par = t_parent[tn];
            l_flat[k] = t_flat[tn]; l_parent[k] = (int8_t)par; l_act[k] = t_act[tn];
            l_head[k] = -1; l_next[k] = l_head[par]; l_head[par] = (int8_t)k;
            expand(k);
        }

        // flow.rs:64-87 dfs over every walker node, node-index order
        const int hero = ep.walker;
        for (int k = 0; k < ln; ++k) {
            const FlatNode nd = g.nodes[l_flat[k]];
            if (nd.turn != hero || l_head[k] < 0) continue;
            // ancestor_reach (flow.rs:166-174): upward over non-walker decision ancestors
            float cf = 1.0f, sm = 1.0f;
            for (int node = k; l_parent[node] >= 0; node = l_parent[node]) {
                const FlatNode pn = g.nodes[l_flat[l_parent[node]]];
                if (pn.turn <= TURN_P1 && pn.turn != hero) {
                    cf = cf * s_sigma[pn.info * kMaxActions + l_act[node]];
                    sm = sm * s_q[pn.info * kMaxActions + l_act[node]];
                }
            }
            const float reach = cf / sm;
            float val[kMaxActions];
            int act[kMaxActions], nk = 0;
            for (int c0 = l_head[k]; c0 >= 0; c0 = l_next[c0]) {
                // recursed_value(root, child, 1, 1) (flow.rs:182-216), explicit stack
                int s_node[kMaxDepth], s_it[kMaxDepth];
                float s_rel[kMaxDepth], s_smp[kMaxDepth], s_acc[kMaxDepth];
                int sp = 0;
                s_node[0] = c0; s_it[0] = l_head[c0]; s_rel[0] = 1.0f; s_smp[0] = 1.0f; s_acc[0] = 0.0f;
                float ret = 0.0f;
                while (true) {
                    const int n = s_node[sp];
                    bool done = false;
                    if (l_head[n] < 0) {  // nash.rs:66-79 terminal_value
                        const int fn = l_flat[n];
                        float u = hero == 0 ? g.nodes[fn].payoff0 : g.payoff1[fn];
                        if (g.nodes[fn].turn == TURN_CHANCE) {  // an unexpanded chance node (subgame): V(I) of the nearest non-chance ancestor
                            int anc = n;
                            do { anc = l_parent[anc]; } while (anc >= 0 && g.nodes[l_flat[anc]].turn == TURN_CHANCE);
                            u = anc >= 0 ? s_pay[g.nodes[l_flat[anc]].info] : 0.0f;
                        }
                        ret = s_rel[sp] / s_smp[sp] * u;
                        done = true;
                    } else if (s_it[sp] < 0) {
                        ret = s_acc[sp];
                        done = true;
                    }
                    if (done) {
                        if (sp == 0) break;
                        --sp;
                        s_acc[sp] = s_acc[sp] + ret;
                        continue;
                    }
                    const int c = s_it[sp];
                    s_it[sp] = l_next[c];
                    const FlatNode pn = g.nodes[l_flat[n]];
                    float r2 = s_rel[sp], s2 = s_smp[sp];
                    if (pn.turn <= TURN_P1) {
                        r2 = r2 * s_sigma[pn.info * kMaxActions + l_act[c]];
                        if (pn.turn != hero) s2 = s2 * s_q[pn.info * kMaxActions + l_act[c]];
                    }
                    ++sp;
                    s_node[sp] = c; s_it[sp] = l_head[c]; s_rel[sp] = r2; s_smp[sp] = s2; s_acc[sp] = 0.0f;
                }
                act[nk] = l_act[c0];
                val[nk] = reach * ret;
                ++nk;
            }
            const float* sg = s_sigma + nd.info * kMaxActions;
            float ev = 0.0f;
            for (int i = 0; i < nk; ++i) ev = ev + sg[act[i]] * val[i];
            int slot = 0;
            for (; slot < nrec; ++slot) if (rec_info[slot] == nd.info) break;
            if (slot == nrec) {
                rec_info[slot] = nd.info; rec_mask[slot] = 0; rec_pay[slot] = 0.0f;
                for (int a = 0; a < kMaxActions; ++a) rec_dr[slot][a] = 0.0f;
                ++nrec;
            }
            rec_pay[slot] += ev;
            for (int i = 0; i < nk; ++i) {
                rec_dr[slot][act[i]] += val[i] - ev;
                rec_mask[slot] |= (uint8_t)(1u << act[i]);
            }
        }
    }

    // ── stable multisplit of the block's records by infoset (tree order preserved) ──
    for (int s = 0; s < nrec; ++s) atomicOr(&s_bits[warp * I + rec_info[s]], 1u << lane);
    __syncthreads();
    for (int x = tid; x < I; x += kTreesPerBlock) {
        uint32_t run = 0;
        for (int wv = 0; wv < 4; ++wv) {
            s_woff[wv * I + x] = run;
            run += __popc(s_bits[wv * I + x]);
        }
        s_cnt[x] = run;
    }
    __syncthreads();
    // exclusive scan of s_cnt over infosets → s_off (warp 0, 32 infosets per step)
    if (warp == 0) {
        uint32_t carry = 0;
        for (int base = 0; base < I; base += 32) {
            const int x = base + lane;
            const uint32_t v = x < I ? s_cnt[x] : 0u;
            uint32_t inc = v;
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += up;
            }
            if (x < I) s_off[x] = carry + inc - v;
            carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
    }
    __syncthreads();
    const size_t total = (size_t)sc.nblk * sc.cap;
    const size_t base = (size_t)blockIdx.x * sc.cap;
    for (int s = 0; s < nrec; ++s) {
        const int x = rec_info[s];
        const uint32_t bits = s_bits[warp * I + x];
        const uint32_t pos = s_off[x] + s_woff[warp * I + x] + __popc(bits & ((1u << lane) - 1u));
        sc.pay[base + pos] = rec_pay[s];
        sc.mask[base + pos] = rec_mask[s];
        const int A = g.info_actions[x];
        for (int a = 0; a < A; ++a) sc.dr[(size_t)a * total + base + pos] = rec_dr[s][a];
    }
    for (int x = tid; x < I; x += kTreesPerBlock) {
        sc.m_off[(size_t)x * sc.nblk + blockIdx.x] = (int32_t)s_off[x];
        sc.m_cnt[(size_t)x * sc.nblk + blockIdx.x] = (int32_t)s_cnt[x];
    }
    if (ep.fold_mode == RBP_FOLD_BATCHED) {
        // per-block sums of this block's Decisions, sequential in tree order (the innermost level of the blocked order)
        __syncthreads();
        for (int x = tid; x < I; x += kTreesPerBlock) {
            const int n = (int)s_cnt[x], A = g.info_actions[x];
            float dr[kMaxActions] = {0.0f, 0.0f, 0.0f, 0.0f}, pay = 0.0f;
            uint32_t na[kMaxActions] = {0u, 0u, 0u, 0u};
            const size_t first = base + s_off[x];
            for (int e = 0; e < n; ++e) {
                const uint32_t m = sc.mask[first + e];
                for (int a = 0; a < A; ++a)
                    if (m >> a & 1u) { dr[a] = dr[a] + sc.dr[(size_t)a * total + first + e]; na[a] += 1u; }
                pay = pay + sc.pay[first + e];
            }
            for (int a = 0; a < kMaxActions; ++a) {
                sc.bp_f[((size_t)a * I + x) * sc.nblk + blockIdx.x] = dr[a];
                sc.bp_na[((size_t)a * I + x) * sc.nblk + blockIdx.x] = na[a];
            }
            sc.bp_f[((size_t)kMaxActions * I + x) * sc.nblk + blockIdx.x] = pay;
        }
    }
    // telemetry (metrics/mod.rs): nodes, infosets
    unsigned long long nn = (unsigned long long)ln, ni = (unsigned long long)nrec;
    for (int d = 16; d > 0; d >>= 1) {
        nn += __shfl_down_sync(0xFFFFFFFFu, nn, d);
        ni += __shfl_down_sync(0xFFFFFFFFu, ni, d);
    }
    if (lane == 0) {
        atomicAdd(&sc.counters[0], nn);
        atomicAdd(&sc.counters[1], ni);
    }
}

// ───────────────────────────── K3: ordered fold ─────────────────────────────
// One block (3 warps) per walker infoset.  warp 0 folds regrets, warp 1 weights, warp 2 payoff+visits — the three
// chains of solver.rs:143-192 are independent, so they run concurrently; inside a warp lane `a` owns row `a`.
// Each chain is inherently serial (one schedule application per Decisions, in tree order), so the kernel is built
// around chain latency: the infoset's records are staged 32 tree-blocks at a time into shared memory with
// coalesced loads, and the chain loop then runs on shared-memory operands only.
constexpr int kFoldChunkBlocks = 32;
constexpr int kFoldCap = kFoldChunkBlocks * kTreesPerBlock;  // records staged per chunk (one per tree at most)

struct RegretConst {  // loop-invariant parts of regret/*.rs
    float d_lin, d_pos, d_neg, d_zero, floor;
};
template <int RS>
__device__ __forceinline__ float regret_gain(const RegretConst& c, float net, float add) {
    float acc;
    if (RS == RBP_REGRET_SUMMED || RS == RBP_REGRET_FLOORED) acc = net + add;
    else if (RS == RBP_REGRET_LINEAR) acc = net * c.d_lin + add;
    else if (RS == RBP_REGRET_DISCOUNTED) acc = net * (net > 0.0f ? c.d_pos : (net < 0.0f ? c.d_neg : c.d_zero)) + add;
    else acc = net > 0.0f ? net + add : net * c.d_lin + add;
    return fmax_ref(acc, c.floor);
}

template <int RS, int WS, bool MASKED>
__global__ void __launch_bounds__(96)
mccfr_fold_kernel(DevGame g, rbp_encounter_t* __restrict__ table, Scratch sc, EpochArgs ep) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int x = blockIdx.x;
    if (g.info_player[x] != ep.walker) return;
    const int A = g.info_actions[x], row = g.info_row[x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_dr = reinterpret_cast<float*>(smem_raw);          // [A][kFoldCap]  (warp 0)
    float* s_pay = s_dr + (size_t)A * kFoldCap;                // [kFoldCap]     (warp 2)
    uint8_t* s_mask = reinterpret_cast<uint8_t*>(s_pay + kFoldCap);  // [kFoldCap] (warp 0)
    const size_t total = (size_t)sc.nblk * sc.cap;
    const int32_t* off = sc.m_off + (size_t)x * sc.nblk;
    const int32_t* cnt = sc.m_cnt + (size_t)x * sc.nblk;

    // regret matching on the PRE-fold regrets: the policy every Decisions of this epoch carries
    // (strategy/profile.rs:47-51); all warps read before warp 0 stores
    const int a = lane < A ? lane : 0;
    float rd = 0.0f, ra = 0.0f;
    for (int k = 0; k < A; ++k) {
        float r = fmax_ref(table[row + k].regret, kEps);
        rd = rd + r;
        if (k == a) ra = r;
    }
    const float sigma = ra / rd;
    __syncthreads();

    if (warp == 1) {  // solver.rs:158-167 update_weight: n identical applications
        int n = 0;
        for (int b = lane; b < sc.nblk; b += 32) n += cnt[b];
        for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, d);
        if (lane < A) {
            float W = table[row + lane].weight;
            float add = sigma;
            if (WS == RBP_WEIGHT_LINEAR) add = sigma * ep.t;
            if (WS == RBP_WEIGHT_QUADRATIC) add = sigma * ep.t * ep.t;
            if (WS == RBP_WEIGHT_EXPONENTIAL) {
                for (int i = 0; i < n; ++i) W = fmax_ref(W * 0.9999f + add, kEps);
            } else {
#pragma unroll 8
                for (int i = 0; i < n; ++i) W = fmax_ref(W + add, kEps);
            }
            table[row + lane].weight = W;
        }
        return;
    }

    RegretConst rc;
    rc.d_lin = ep.t / (ep.t + 1.0f);
    rc.d_pos = ep.disc_pos / (ep.disc_pos + 1.0f);
    rc.d_neg = ep.disc_neg / (ep.disc_neg + 1.0f);
    rc.d_zero = rc.d_lin;
    rc.floor = RS == RBP_REGRET_SUMMED ? -INFINITY : (RS == RBP_REGRET_FLOORED ? 0.0f : ep.hyper.regret_min);

    float R = 0.0f, ev = 0.0f;
    uint32_t visits = 0;
    unsigned long long ups = 0;
    if (lane < A) {
        R = table[row + lane].regret;
        ev = table[row + lane].payoff;
        visits = table[row + lane].visits;
    }
    // rows of one infoset are always visited together, so their visit counters agree unless a profile with
    // unequal counters was imported; the shared reciprocal table below needs them equal
    const uint32_t visits0 = __shfl_sync(0xFFFFFFFFu, visits, 0);
    const bool uniform_visits = __all_sync(0xFFFFFFFFu, lane >= A || visits == visits0);
    float* s_cnt = reinterpret_cast<float*>(s_mask + kFoldCap);  // [kFoldCap] (float)(visits+1)   (warp 2)
    float* s_rcp = s_cnt + kFoldCap;                             // [kFoldCap] RN(1/(visits+1))    (warp 2)
    __shared__ int s_pfx[3][33];
    __shared__ uint32_t s_src[3][32];
    uint32_t done = 0;
    for (int c = 0; c < sc.nblk; c += kFoldChunkBlocks) {
        const int b = c + lane;
        const int n_mine = b < sc.nblk ? cnt[b] : 0;
        const int o_mine = b < sc.nblk ? off[b] : 0;
        int incl = n_mine;
        for (int d = 1; d < 32; d <<= 1) {
            int up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += up;
        }
        const int chunk_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const int p_mine = incl - n_mine;
        // stage the chunk into shared memory: segment s = this infoset's records of tree-block c+s (contiguous,
        // tree order).  Each lane resolves flat chunk positions to (segment, offset) by a 5-step search over the
        // segment prefix, so a batch of independent loads is in flight before the first store.
        s_pfx[warp][lane] = p_mine;
        s_src[warp][lane] = (uint32_t)((size_t)b * sc.cap + o_mine);
        if (lane == 0) s_pfx[warp][32] = chunk_total;
        __syncwarp();
        constexpr int U = 4;
        for (int e0 = 0; e0 < chunk_total; e0 += 32 * U) {
            uint32_t src[U];
            float v0[U], v1[U], v2[U], v3[U];
            uint8_t mk[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * 32 + lane;
                int sg = 0;
                if (e < chunk_total) {
                    if (s_pfx[warp][16] <= e) sg = 16;
                    if (s_pfx[warp][sg + 8] <= e) sg += 8;
                    if (s_pfx[warp][sg + 4] <= e) sg += 4;
                    if (s_pfx[warp][sg + 2] <= e) sg += 2;
                    if (s_pfx[warp][sg + 1] <= e) sg += 1;
                    src[u] = s_src[warp][sg] + (uint32_t)(e - s_pfx[warp][sg]);
                } else {
                    src[u] = 0xFFFFFFFFu;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v0[u] = v1[u] = v2[u] = v3[u] = 0.0f; mk[u] = 0;
                if (src[u] != 0xFFFFFFFFu) {
                    if (warp == 0) {
                        v0[u] = sc.dr[src[u]];
                        if (A > 1) v1[u] = sc.dr[total + src[u]];
                        if (A > 2) v2[u] = sc.dr[2 * total + src[u]];
                        if (A > 3) v3[u] = sc.dr[3 * total + src[u]];
                        if (MASKED) mk[u] = sc.mask[src[u]];
                    } else {
                        v0[u] = sc.pay[src[u]];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * 32 + lane;
                if (src[u] != 0xFFFFFFFFu) {
                    if (warp == 0) {
                        s_dr[e] = v0[u];
                        if (A > 1) s_dr[kFoldCap + e] = v1[u];
                        if (A > 2) s_dr[2 * kFoldCap + e] = v2[u];
                        if (A > 3) s_dr[3 * kFoldCap + e] = v3[u];
                        if (MASKED) s_mask[e] = mk[u];
                    } else {
                        s_pay[e] = v0[u];
                    }
                }
            }
        }
        if (warp == 2 && uniform_visits) {  // divisors and their reciprocals for this chunk, all lanes, off the chain
            for (int e = lane; e < chunk_total; e += 32) {
                const float bc = (float)(visits0 + done + (uint32_t)e + 1u);
                s_cnt[e] = bc;
                s_rcp[e] = __frcp_rn(bc);
            }
        }
        __syncwarp();
        if (lane < A) {
            if (warp == 0) {  // solver.rs:143-152 update_regret
                const float* d = s_dr + lane * kFoldCap;
#pragma unroll 8
                for (int e = 0; e < chunk_total; ++e) {
                    if (!MASKED || (s_mask[e] >> lane & 1u)) { R = regret_gain<RS>(rc, R, d[e]); ++ups; }
                }
            } else if (uniform_visits) {  // solver.rs:174-192 update_payoff (Welford) then update_visits
                // operands are pulled into registers 8 entries at a time (LDS.128), so only the arithmetic is on the chain
                int e = 0;
                for (; e + 8 <= chunk_total; e += 8) {
                    const float4 p0 = *reinterpret_cast<const float4*>(s_pay + e), p1 = *reinterpret_cast<const float4*>(s_pay + e + 4);
                    const float4 c0 = *reinterpret_cast<const float4*>(s_cnt + e), c1 = *reinterpret_cast<const float4*>(s_cnt + e + 4);
                    const float4 r0 = *reinterpret_cast<const float4*>(s_rcp + e), r1 = *reinterpret_cast<const float4*>(s_rcp + e + 4);
                    // branch-free fast path for the 8 entries; the guard of div_by_count is accumulated beside the chain
                    // and, if any entry fell outside it, the group is redone with IEEE division from the saved state
                    const float ev_in = ev;
                    bool ok = true;
                    const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                    const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                    const float rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float a = pv[k] - ev;
                        ok &= ((((__float_as_uint(a) >> 23 & 0xFFu) - 63u) < 128u) | (__float_as_uint(a) == 0u)) & ((__float_as_uint(cv[k]) & 0x7FFFFFu) != 0x7FFFFFu);
                        const float q = a * rv[k];
                        const float r = __fmaf_rn(-cv[k], q, a);
                        ev += __fmaf_rn(r, rv[k], q);
                    }
                    if (__builtin_expect(!ok, 0)) {
                        ev = ev_in;
#pragma unroll
                        for (int k = 0; k < 8; ++k) ev += (pv[k] - ev) / cv[k];
                    }
                }
                for (; e < chunk_total; ++e) ev += div_by_count(s_pay[e] - ev, s_cnt[e], s_rcp[e]);
                visits += (uint32_t)chunk_total;
            } else {
                for (int e = 0; e < chunk_total; ++e) {
                    ev += (s_pay[e] - ev) / (float)(visits + 1u);
                    visits += 1u;
                }
            }
        }
        done += (uint32_t)chunk_total;
        __syncwarp();
    }
    if (lane < A) {
        if (warp == 0) {
            table[row + lane].regret = R;
            if (ups) atomicAdd(&sc.counters[2], ups);
        } else {
            table[row + lane].payoff = ev;
            table[row + lane].visits = visits;
        }
    }
}

// ───────────────────────────── K3b: BATCHED fold ─────────────────────────────
// rank level of the blocked order: one block per infoset, warp f sums field f over this rank's tree-blocks in
// order.  Lanes fetch 32 consecutive block partials with one coalesced load; the sum itself stays sequential
// (shuffle-fed), because the blocked order is part of the contract.
constexpr int kPartialFields = 2 * kMaxActions + 2;  // dr[4], pay, na[4], n
__global__ void __launch_bounds__(32 * kPartialFields)
mccfr_rank_partial_kernel(DevGame g, Scratch sc, Partial* __restrict__ out) {
    const int x = blockIdx.x, lane = threadIdx.x & 31, f = threadIdx.x >> 5, I = g.n_infos;
    if (f < kMaxActions + 1) {
        const float* src = sc.bp_f + ((size_t)f * I + x) * sc.nblk;
        float acc = 0.0f;
        for (int b0 = 0; b0 < sc.nblk; b0 += 32) {
            const float v = b0 + lane < sc.nblk ? src[b0 + lane] : 0.0f;
            const int m = min(32, sc.nblk - b0);
            for (int k = 0; k < m; ++k) acc = acc + __shfl_sync(0xFFFFFFFFu, v, k);
        }
        if (lane == 0) { if (f < kMaxActions) out[x].dr[f] = acc; else out[x].pay = acc; }
    } else {
        const bool is_n = f == 2 * kMaxActions + 1;
        const uint32_t* src = is_n ? reinterpret_cast<const uint32_t*>(sc.m_cnt + (size_t)x * sc.nblk)
                                   : sc.bp_na + ((size_t)(f - kMaxActions - 1) * I + x) * sc.nblk;
        uint32_t acc = 0;
        for (int b = lane; b < sc.nblk; b += 32) acc += src[b];
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
        if (lane == 0) {
            if (is_n) { out[x].n = acc; out[x].pad[0] = out[x].pad[1] = 0u; }
            else out[x].na[f - kMaxActions - 1] = acc;
        }
    }
}
__device__ __forceinline__ float regret_gain_rt(const EpochArgs& ep, float net, float add) {
    RegretConst rc;
    rc.d_lin = ep.t / (ep.t + 1.0f);
    rc.d_pos = ep.disc_pos / (ep.disc_pos + 1.0f);
    rc.d_neg = ep.disc_neg / (ep.disc_neg + 1.0f);
    rc.d_zero = rc.d_lin;
    switch (ep.regret_sched) {
        case RBP_REGRET_SUMMED: rc.floor = -INFINITY; return regret_gain<RBP_REGRET_SUMMED>(rc, net, add);
        case RBP_REGRET_FLOORED: rc.floor = 0.0f; return regret_gain<RBP_REGRET_FLOORED>(rc, net, add);
        case RBP_REGRET_LINEAR: rc.floor = ep.hyper.regret_min; return regret_gain<RBP_REGRET_LINEAR>(rc, net, add);
        case RBP_REGRET_DISCOUNTED: rc.floor = ep.hyper.regret_min; return regret_gain<RBP_REGRET_DISCOUNTED>(rc, net, add);
        default: rc.floor = ep.hyper.regret_min; return regret_gain<RBP_REGRET_ASYMMETRIC>(rc, net, add);
    }
}
// world level + schedules: sum the ranks' partials in rank order, then ONE schedule application per row
__global__ void __launch_bounds__(128)
mccfr_apply_batched_kernel(DevGame g, rbp_encounter_t* __restrict__ table, const Partial* __restrict__ gathered, int world, int rank,
                           EpochArgs ep, unsigned long long* __restrict__ counters) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, I = g.n_infos;
    if (x >= I) return;
    float dr[kMaxActions] = {0.0f, 0.0f, 0.0f, 0.0f}, pay = 0.0f;
    uint32_t na[kMaxActions] = {0u, 0u, 0u, 0u}, n = 0;
    for (int r = 0; r < world; ++r) {
        const Partial p = gathered[(size_t)r * I + x];
        for (int a = 0; a < kMaxActions; ++a) { dr[a] = dr[a] + p.dr[a]; na[a] += p.na[a]; }
        pay = pay + p.pay;
        n += p.n;
    }
    if (n == 0) return;
    const int A = g.info_actions[x], row = g.info_row[x];
    float r[kMaxActions], rd = 0.0f;
    for (int a = 0; a < A; ++a) { r[a] = fmax_ref(table[row + a].regret, kEps); rd = rd + r[a]; }
    unsigned long long ups = 0;
    for (int a = 0; a < A; ++a) {
        rbp_encounter_t e = table[row + a];
        if (na[a] > 0) { e.regret = regret_gain_rt(ep, e.regret, dr[a]); ups += gathered[(size_t)rank * I + x].na[a]; }  // telemetry counts this rank's own trees
        const float add = (float)n * (r[a] / rd);
        float acc;
        switch (ep.weight_sched) {
            case RBP_WEIGHT_CONSTANT: acc = e.weight + add; break;
            case RBP_WEIGHT_LINEAR: acc = e.weight + add * ep.t; break;
            case RBP_WEIGHT_QUADRATIC: acc = e.weight + add * ep.t * ep.t; break;
            default: acc = e.weight * 0.9999f + add; break;
        }
        e.weight = fmax_ref(acc, kEps);
        const float mean = pay / (float)n;
        e.payoff += (mean - e.payoff) * (float)n / (float)(e.visits + n);
        e.visits += n;
        table[row + a] = e;
    }
    if (ups) atomicAdd(&counters[2], ups);
}

// self-test of div_by_count against IEEE division: every count b in [1, max_count] against `samples` dividends
// (Philox bits reinterpreted so that all exponents in the guarded range, both signs, occur)
__global__ void div_selftest_kernel(uint32_t max_count, uint32_t samples, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x + 1; b <= max_count; b += gridDim.x * blockDim.x) {
        const float fb = (float)b, rb = __frcp_rn(fb);
        for (uint32_t k = 0; k < samples; k += 4) {
            Philox4 p = philox4x32_10(b, k, 0x5e1f7e57u, 0u, 0x1234u, 0x5678u);
            for (int j = 0; j < 4; ++j) {
                uint32_t bits = p.r[j];
                uint32_t expo = 64u + ((bits >> 23) & 0xFFu) % 128u;  // 2^-63 .. 2^64
                float a = __uint_as_float((bits & 0x807FFFFFu) | (expo << 23));
                if ((k + j) % 16 == 0) a = (float)((int)(bits % 2001u) - 1000) * 0.25f;  // small payoffs like Leduc's
                if ((k + j) % 64 == 1) a = 0.0f;
                if ((k + j) % 64 == 2) a = -0.0f;
                if (__float_as_uint(div_by_count(a, fb, rb)) != __float_as_uint(a / fb)) ++bad;
            }
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ───────────────────────────── K4: exploitability ─────────────────────────────
// solver.rs:327-338 + strategy/nash.rs:31-193 on the enumerated tree.  One block; bottom-up level sweeps
// replace the reference's recursion (pure function of the subtree ⇒ same f32 results); the upward
// `external_reach` product keeps the reference's nearest-ancestor-first order.
__global__ void __launch_bounds__(256)
mccfr_exploit_kernel(DevGame g, const rbp_encounter_t* __restrict__ table, float* __restrict__ U,
                     float* __restrict__ cfv, int32_t* __restrict__ br, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_avg = reinterpret_cast<float*>(smem_raw);  // [I][4]
    const int I = g.n_infos, tid = threadIdx.x, T = blockDim.x;
    for (int x = tid; x < I; x += T) {  // profile.rs:41-45 averaged_distribution
        const int A = g.info_actions[x], row = g.info_row[x];
        float w[kMaxActions], sum = 0.0f;
        for (int a = 0; a < A; ++a) { w[a] = fmax_ref(table[row + a].weight, kEps); sum = sum + w[a]; }
        for (int a = 0; a < A; ++a) s_avg[x * kMaxActions + a] = w[a] / sum;
    }
    __syncthreads();
    float total = 0.0f;
    for (int hero = 0; hero < 2; ++hero) {
        for (int pass = 0; pass < 2; ++pass) {  // pass 0: average strategy everywhere; pass 1: hero plays br[]
            for (int d = g.n_levels - 1; d >= 0; --d) {
                for (int n = g.level_start[d] + tid; n < g.level_start[d + 1]; n += T) {
                    const FlatNode nd = g.nodes[n];
                    float u;
                    if (nd.n_child == 0) {
                        u = hero == 0 ? nd.payoff0 : g.payoff1[n];
                    } else if (nd.turn == TURN_CHANCE) {  // nash.rs:116
                        float s = 0.0f;
                        for (int k = 0; k < nd.n_child; ++k) s = s + U[nd.first_child + k];
                        u = s / (float)nd.n_child;
                    } else if (pass == 1 && nd.turn == hero) {  // nash.rs:118-124
                        u = U[nd.first_child + br[nd.info]];
                    } else {  // nash.rs:125-132
                        float s = 0.0f;
                        for (int k = 0; k < nd.n_child; ++k) s = s + s_avg[nd.info * kMaxActions + k] * U[nd.first_child + k];
                        u = s;
                    }
                    U[n] = u;
                }
                __syncthreads();
            }
            if (pass == 1) break;
            // nash.rs:171-193 optimal_cfactual_choice
            for (int xa = tid; xa < I * kMaxActions; xa += T) {
                const int x = xa / kMaxActions, a = xa - x * kMaxActions;
                if (g.info_player[x] != hero || a >= g.info_actions[x]) continue;
                float sum = 0.0f;
                for (int s = g.span_start[x]; s < g.span_start[x + 1]; ++s) {
                    const int c = g.nodes[g.span_nodes[s]].first_child + a;
                    float reach = 1.0f;  // nash.rs:140-145 external_reach
                    for (int node = c; g.parent[node] >= 0; node = g.parent[node]) {
                        const int par = g.parent[node];
                        const FlatNode pn = g.nodes[par];
                        if (pn.turn <= TURN_P1 && pn.turn != hero) reach = reach * s_avg[pn.info * kMaxActions + (node - pn.first_child)];
                    }
                    sum = sum + reach * U[c];
                }
                cfv[xa] = sum;
            }
            __syncthreads();
            for (int x = tid; x < I; x += T) {
                if (g.info_player[x] != hero) continue;
                float best = 0.0f; int besta = -1;
                for (int a = 0; a < g.info_actions[x]; ++a) {
                    const float v = cfv[x * kMaxActions + a];
                    if (besta < 0 || !(v < best)) { best = v; besta = a; }  // Iterator::max_by keeps the last maximum
                }
                br[x] = besta;
            }
            __syncthreads();
        }
        total = total + U[0];
        __syncthreads();
    }
    if (tid == 0) *out = total / 2.0f;
}

// L2 flush between timed steps (bench hygiene): stream a buffer larger than the 126 MB L2
__global__ void l2_flush_kernel(uint4* buf, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        buf[i] = make_uint4((uint32_t)i, 0u, 0u, 0u);
}

// table reset: every row reads as the reference's "missing" row (book.rs:93-122; default_regret = 0 for flat games)
__global__ void table_reset_kernel(rbp_encounter_t* table, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) table[i] = rbp_encounter_t{0.0f, 0.0f, 0.0f, 0u};
}


// ───────────────────────────── subgame: local tables over a frozen blueprint ─────────────────────────────
// subgame/src/world/profile.rs:62-145 + mccfr/src/strategy/profile.rs:94-104.  A WorldProfile edge that was never written reads the
// blueprint (regret and weight floored at EPSILON); `update_regret` reads that value and then `mut_regret` creates the local edge from
// `blueprint.warmstart` and overwrites its regret; `update_weight` then finds the edge locally with the warmstart weight.  So every
// world's table starts at {weight: averaged policy * k * (k + 1) / 2, regret: max(blueprint regret, EPS), payoff 0, visits 0} and an
// edge with visits = 0 reads `fb_weight` = max(blueprint weight, EPS) instead of its own weight (and, for the V(I) of a chance leaf, the
// blueprint's payoff, second half of `fb_weight`).  One thread per infoset.
__global__ void __launch_bounds__(128)
subgame_seed_kernel(DevGame g, const rbp_encounter_t* __restrict__ blueprint, float k, int worlds, rbp_encounter_t* __restrict__ table, float* __restrict__ fb_weight) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.n_infos) return;
    const int A = g.info_actions[x], row = g.info_row[x];
    float w[kMaxActions], sum = 0.0f;  // profile.rs:40-44 averaged_distribution of the blueprint
    for (int a = 0; a < A; ++a) { w[a] = fmax_ref(blueprint[row + a].weight, kEps); sum = sum + w[a]; }
    for (int a = 0; a < A; ++a) {
        const float policy = w[a] / sum;
        rbp_encounter_t e;
        e.weight = policy * k * (k + 1.0f) / 2.0f;
        e.regret = fmax_ref(blueprint[row + a].regret, kEps);
        e.payoff = 0.0f;
        e.visits = 0u;
        for (int wd = 0; wd < worlds; ++wd) table[(size_t)wd * g.n_rows + row + a] = e;
        fb_weight[row + a] = w[a];
        fb_weight[g.n_rows + row + a] = blueprint[row + a].payoff;  // world/profile.rs:133-138: cum_payoff falls through unfloored
    }
}

}  // namespace rbp

// ───────────────────────────── host: handle + C ABI ─────────────────────────────
using namespace rbp;

struct rbp_solver {
    FlatGame game;
    DevGame dev{};
    Scratch sc{};
    rbp_encounter_t* table = nullptr;
    float *U = nullptr, *cfv = nullptr, *expl = nullptr;
    int32_t* br = nullptr;
    std::vector<void*> owned;
    std::vector<uint8_t> touched;  // rows written by import (export lists visits>0 or touched)
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int device = 0, regret = 0, weight = 0, sampling = 0, fold_mode = 0, batch = 1;
    int world_rank = 0, world_size = 1;
    rbp_comm* comm = nullptr;   // rbp_solver_attach_comm
    Partial* gathered = nullptr;  // [world][infosets] partial sums of every rank
    bool sampled = false;  // rbp_solver_sample ran for the current epoch (its partial sums are what fold_gathered consumes)
    uint64_t seed = 0, epochs = 0;
    rbp_hyper_t hyper{};
    size_t sample_smem = 0, fold_smem = 0;
    uint4* flush_buf = nullptr;
    Partial* delta = nullptr;  // this rank's BATCHED partial sums
    std::vector<cudaEvent_t> events;
    // subgame (rbp_subgame_*): `table` holds sub_worlds x n_rows rows, one table per world (WorldInfo(world, info) keys)
    int sub_worlds = 0, sub_world = 0, sub_entry_plus1 = 0;
    float* fb_weight = nullptr;  // [2][n_rows] the blueprint's weights (floored) and payoffs, read for edges the subgame has not written yet
};
static inline rbp_encounter_t* tbl(const rbp_solver* s) { return s->table + (size_t)s->sub_world * (size_t)s->game.n_rows; }

namespace {
template <class T>
int upload(rbp_solver* s, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    RBP_CUDA(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
    s->owned.push_back(p);
    RBP_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T*>(p);
    return RBP_OK;
}
template <class T>
int alloc(rbp_solver* s, size_t n, T** out) {
    void* p = nullptr;
    RBP_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    s->owned.push_back(p);
    RBP_CUDA(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    *out = static_cast<T*>(p);
    return RBP_OK;
}
EpochArgs epoch_args(const rbp_solver* s) {
    EpochArgs ep{};
    ep.seed_lo = (uint32_t)s->seed; ep.seed_hi = (uint32_t)(s->seed >> 32);
    ep.epoch = (uint32_t)s->epochs;
    ep.walker = (int)(s->epochs % 2);  // book.rs:142-144
    ep.batch = s->batch;
    ep.tree_base = s->sub_worlds ? s->sub_world : s->world_rank * s->batch;  // subgame trees draw with tree id = world
    ep.entry_plus1 = s->sub_entry_plus1;
    ep.sampling = s->sampling;
    ep.hyper = s->hyper;
    ep.regret_sched = s->regret; ep.weight_sched = s->weight; ep.fold_mode = s->fold_mode;
    ep.t = (float)s->epochs;
    ep.disc_pos = powf(ep.t / 1.0f, 1.5f);
    ep.disc_neg = powf(ep.t / 1.0f, 0.5f);
    return ep;
}
int launch_sample(rbp_solver* s, const EpochArgs& ep) {
    mccfr_sample_kernel<<<s->sc.nblk, kTreesPerBlock, s->sample_smem, s->stream>>>(s->dev, tbl(s), s->fb_weight, s->sc, ep);
    RBP_LAUNCHED();
    return RBP_OK;
}
int launch_rank_partial(rbp_solver* s) {
    mccfr_rank_partial_kernel<<<s->dev.n_infos, 32 * kPartialFields, 0, s->stream>>>(s->dev, s->sc, s->delta);
    RBP_LAUNCHED();
    return RBP_OK;
}
int launch_apply_batched(rbp_solver* s, const EpochArgs& ep, const Partial* gathered, int world) {
    mccfr_apply_batched_kernel<<<(s->dev.n_infos + 127) / 128, 128, 0, s->stream>>>(s->dev, s->table, gathered, world, world > 1 ? s->world_rank : 0, ep, s->sc.counters);
    RBP_LAUNCHED();
    return RBP_OK;
}
template <int RS, int WS>
int launch_fold_rw(rbp_solver* s, const EpochArgs& ep) {
    const bool masked = s->sampling != RBP_SAMPLING_EXTERNAL;
    if (masked) mccfr_fold_kernel<RS, WS, true><<<s->dev.n_infos, 96, s->fold_smem, s->stream>>>(s->dev, tbl(s), s->sc, ep);
    else mccfr_fold_kernel<RS, WS, false><<<s->dev.n_infos, 96, s->fold_smem, s->stream>>>(s->dev, tbl(s), s->sc, ep);
    RBP_LAUNCHED();
    return RBP_OK;
}
template <int RS>
int launch_fold_r(rbp_solver* s, const EpochArgs& ep) {
    switch (s->weight) {
        case RBP_WEIGHT_CONSTANT: return launch_fold_rw<RS, RBP_WEIGHT_CONSTANT>(s, ep);
        case RBP_WEIGHT_LINEAR: return launch_fold_rw<RS, RBP_WEIGHT_LINEAR>(s, ep);
        case RBP_WEIGHT_QUADRATIC: return launch_fold_rw<RS, RBP_WEIGHT_QUADRATIC>(s, ep);
        default: return launch_fold_rw<RS, RBP_WEIGHT_EXPONENTIAL>(s, ep);
    }
}
int launch_fold(rbp_solver* s, const EpochArgs& ep) {
    switch (s->regret) {
        case RBP_REGRET_SUMMED: return launch_fold_r<RBP_REGRET_SUMMED>(s, ep);
        case RBP_REGRET_FLOORED: return launch_fold_r<RBP_REGRET_FLOORED>(s, ep);
        case RBP_REGRET_LINEAR: return launch_fold_r<RBP_REGRET_LINEAR>(s, ep);
        case RBP_REGRET_DISCOUNTED: return launch_fold_r<RBP_REGRET_DISCOUNTED>(s, ep);
        default: return launch_fold_r<RBP_REGRET_ASYMMETRIC>(s, ep);
    }
}
template <int RS, int WS>
int fold_attr_rw(size_t smem) {
    RBP_CUDA(cudaFuncSetAttribute(mccfr_fold_kernel<RS, WS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RBP_CUDA(cudaFuncSetAttribute(mccfr_fold_kernel<RS, WS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return RBP_OK;
}
template <int RS>
int fold_attr_r(size_t smem) {
    int st;
    if ((st = fold_attr_rw<RS, RBP_WEIGHT_CONSTANT>(smem))) return st;
    if ((st = fold_attr_rw<RS, RBP_WEIGHT_LINEAR>(smem))) return st;
    if ((st = fold_attr_rw<RS, RBP_WEIGHT_QUADRATIC>(smem))) return st;
    return fold_attr_rw<RS, RBP_WEIGHT_EXPONENTIAL>(smem);
}
int fold_attr(size_t smem) {
    int st;
    if ((st = fold_attr_r<RBP_REGRET_SUMMED>(smem))) return st;
    if ((st = fold_attr_r<RBP_REGRET_FLOORED>(smem))) return st;
    if ((st = fold_attr_r<RBP_REGRET_LINEAR>(smem))) return st;
    if ((st = fold_attr_r<RBP_REGRET_DISCOUNTED>(smem))) return st;
    return fold_attr_r<RBP_REGRET_ASYMMETRIC>(smem);
}
}  // namespace

extern "C" {

const char* rbp_status_string(int status) {
    switch (status) {
        case RBP_OK: return "ok";
        case RBP_ERR_INVALID: return "invalid argument";
        case RBP_ERR_NO_DEVICE: return "no CUDA device (librbp_b200 has no CPU fallback)";
        case RBP_ERR_CUDA: return "CUDA error";
        case RBP_ERR_CAPACITY: return "capacity exceeded";
        case RBP_ERR_STATE: return "bad call sequence";
        default: return "unknown status";
    }
}
const char* rbp_last_error(void) {
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_err_mu);
    copy = g_err;
    return copy.c_str();
}
uint64_t rbp_kernel_launches(void) { return g_launches.load(); }
int rbp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
void rbp_hyper_default(rbp_hyper_t* h) {
    h->temperature = 1.0f; h->smoothing = 2.0f; h->curiosity = 0.05f;
    h->prune_threshold = -3e5f; h->prune_explore = 0.05f; h->prune_warmup = 16384; h->regret_min = -4e6f;
}
void rbp_philox4x32_10(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
    Philox4 p = philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
    for (int i = 0; i < 4; ++i) out[i] = p.r[i];
}

int rbp_solver_create(int game, int regret, int weight, int sampling, int fold_mode, int batch, uint64_t seed,
                      const rbp_hyper_t* hyper, int device, rbp_solver_t** out) {
    if (!out) return RBP_ERR_INVALID;
    *out = nullptr;
    if (regret < 0 || regret > 4 || weight < 0 || weight > 3 || batch < 1) return RBP_ERR_INVALID;
    if (sampling != RBP_SAMPLING_EXTERNAL && sampling != RBP_SAMPLING_PRUNABLE && sampling != RBP_SAMPLING_PLURIBUS && sampling != RBP_SAMPLING_TARGETED) {
        set_last_error("training needs a sampling scheme (VanillaSampling is exploitability-only, sample/vanilla.rs)");
        return RBP_ERR_INVALID;
    }
    if (fold_mode != RBP_FOLD_ORDERED && fold_mode != RBP_FOLD_BATCHED) return RBP_ERR_INVALID;
    if (rbp_device_count() <= device) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    rbp_solver* s = new rbp_solver();
    if (!build_flat_game(game, &s->game)) { delete s; return RBP_ERR_INVALID; }
    const FlatGame& G = s->game;
    if (G.max_tree_nodes > kMaxTreeNodes || G.max_tree_infos > kMaxRecords || G.max_depth + 1 > kMaxDepth ||
        (int)G.info_key.size() > kMaxInfos || (int)G.nodes.size() > 32767) {
        set_last_error("flat game exceeds compiled capacities");
        delete s;
        return RBP_ERR_CAPACITY;
    }
    s->device = device; s->regret = regret; s->weight = weight; s->sampling = sampling; s->fold_mode = fold_mode;
    s->batch = batch; s->seed = seed;
    if (hyper) s->hyper = *hyper; else rbp_hyper_default(&s->hyper);
    int st = RBP_OK;
    auto fail = [&](int code) { rbp_solver_destroy(s); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fail(RBP_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(RBP_ERR_CUDA);
    DevGame& d = s->dev;
    if ((st = upload(s, G.nodes, &d.nodes))) return fail(st);
    if ((st = upload(s, G.payoff1, &d.payoff1))) return fail(st);
    if ((st = upload(s, G.root_table, &d.root_table))) return fail(st);
    if ((st = upload(s, G.info_row, &d.info_row))) return fail(st);
    if ((st = upload(s, G.info_actions, &d.info_actions))) return fail(st);
    if ((st = upload(s, G.info_player, &d.info_player))) return fail(st);
    if ((st = upload(s, G.parent, &d.parent))) return fail(st);
    if ((st = upload(s, G.level_start, &d.level_start))) return fail(st);
    if ((st = upload(s, G.span_start, &d.span_start))) return fail(st);
    if ((st = upload(s, G.span_nodes, &d.span_nodes))) return fail(st);
    d.n_nodes = (int)G.nodes.size(); d.n_infos = (int)G.info_key.size(); d.n_rows = G.n_rows;
    d.n_levels = (int)G.level_start.size() - 1; d.deck = G.deck;
    if ((st = alloc(s, (size_t)G.n_rows, &s->table))) return fail(st);
    s->touched.assign(G.n_rows, 0);
    Scratch& sc = s->sc;
    sc.nblk = (batch + kTreesPerBlock - 1) / kTreesPerBlock;
    sc.cap = kTreesPerBlock * G.max_tree_infos;
    const size_t total = (size_t)sc.nblk * sc.cap;
    if ((st = alloc(s, total, &sc.pay))) return fail(st);
    if ((st = alloc(s, total * kMaxActions, &sc.dr))) return fail(st);
    if ((st = alloc(s, total, &sc.mask))) return fail(st);
    if ((st = alloc(s, (size_t)d.n_infos * sc.nblk, &sc.m_off))) return fail(st);
    if ((st = alloc(s, (size_t)d.n_infos * sc.nblk, &sc.m_cnt))) return fail(st);
    if ((st = alloc(s, 3, &sc.counters))) return fail(st);
    if (fold_mode == RBP_FOLD_BATCHED) {
        if ((st = alloc(s, (size_t)5 * d.n_infos * sc.nblk, &sc.bp_f))) return fail(st);
        if ((st = alloc(s, (size_t)4 * d.n_infos * sc.nblk, &sc.bp_na))) return fail(st);
        if ((st = alloc(s, (size_t)d.n_infos, &s->delta))) return fail(st);
    }
    if ((st = alloc(s, (size_t)d.n_nodes, &s->U))) return fail(st);
    if ((st = alloc(s, (size_t)d.n_infos * kMaxActions, &s->cfv))) return fail(st);
    if ((st = alloc(s, (size_t)d.n_infos, &s->br))) return fail(st);
    if ((st = alloc(s, 1, &s->expl))) return fail(st);
    s->sample_smem = (size_t)d.n_infos * (3 * kMaxActions * sizeof(float) + 10 * sizeof(uint32_t) + sizeof(float));
    if (cudaFuncSetAttribute(mccfr_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->sample_smem) != cudaSuccess)
        return fail(RBP_ERR_CUDA);
    int max_a = 1;
    for (uint8_t na : G.info_actions) max_a = std::max<int>(max_a, na);
    s->fold_smem = (size_t)kFoldCap * ((max_a + 3) * sizeof(float) + 1);
    if ((st = fold_attr(s->fold_smem))) return fail(st);
    *out = s;
    return RBP_OK;
}

int rbp_solver_set_stream(rbp_solver_t* s, void* cuda_stream) {
    if (!s) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(s->device));
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    if (s->own_stream) cudaStreamDestroy(s->stream);
    s->stream = (cudaStream_t)cuda_stream;
    s->own_stream = false;
    return RBP_OK;
}

int rbp_solver_set_world(rbp_solver_t* s, int world_rank, int world_size) {
    if (!s || world_size < 1 || world_rank < 0 || world_rank >= world_size) return RBP_ERR_INVALID;
    s->world_rank = world_rank; s->world_size = world_size;
    return RBP_OK;
}

void rbp_solver_destroy(rbp_solver_t* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) { cudaStreamSynchronize(s->stream); if (s->own_stream) cudaStreamDestroy(s->stream); }
    for (void* p : s->owned) cudaFree(p);
    for (cudaEvent_t e : s->events) cudaEventDestroy(e);
    delete s;
}

// Small games across ranks inside the library (BATCHED fold only: the ordered fold over a few hundred rows is serial per row): every
// rank samples its shard of the epoch's trees, the blocked partial sums (48 B per infoset) are all-gathered on the solver's stream
// and every rank folds them in rank order — tables stay bit-identical.  Collective.
int rbp_solver_attach_comm(rbp_solver_t* s, rbp_comm_t* c) {
    if (!s || !c) return RBP_ERR_INVALID;
    if (s->fold_mode != RBP_FOLD_BATCHED) { set_last_error("rbp_solver_attach_comm needs RBP_FOLD_BATCHED"); return RBP_ERR_STATE; }
    if (c->device != s->device) { set_last_error("communicator lives on another device"); return RBP_ERR_INVALID; }
    RBP_CUDA(cudaSetDevice(s->device));
    int st = alloc(s, (size_t)c->world * s->dev.n_infos, &s->gathered);
    if (st) return st;
    s->comm = c; s->world_rank = c->rank; s->world_size = c->world;
    return RBP_OK;
}

int rbp_solver_step(rbp_solver_t* s, uint64_t n_epochs) {
    if (!s) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(s->device));
    if (s->fold_mode == RBP_FOLD_BATCHED && s->world_size > 1 && !s->comm) {
        set_last_error("BATCHED fold with world_size > 1 and no communicator: rbp_solver_attach_comm, or drive rbp_solver_sample / exchange / rbp_solver_fold_gathered");
        return RBP_ERR_STATE;
    }
    for (uint64_t i = 0; i < n_epochs; ++i) {
        EpochArgs ep = epoch_args(s);
        int st;
        if ((st = launch_sample(s, ep))) return st;
        if (s->fold_mode == RBP_FOLD_BATCHED && s->comm && s->world_size > 1) {
            if ((st = launch_rank_partial(s))) return st;
            if ((st = comm::all_gather_bytes(s->comm, s->delta, s->gathered, (size_t)s->dev.n_infos * sizeof(Partial), s->stream))) return st;
            if ((st = launch_apply_batched(s, ep, s->gathered, s->world_size))) return st;
        } else if (s->fold_mode == RBP_FOLD_BATCHED) {
            if ((st = launch_rank_partial(s))) return st;
            if ((st = launch_apply_batched(s, ep, s->delta, 1))) return st;
        } else if ((st = launch_fold(s, ep))) return st;
        s->epochs += 1;  // book.rs:138-140
    }
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    return RBP_OK;
}

// `Solver::spend` (crates/mccfr/src/solver/solver.rs:130-137): epochs in a tight loop until the wall-clock budget is used up —
// what the real-time (subgame) players call with their per-decision budget
int rbp_solver_spend(rbp_solver_t* s, double seconds, uint64_t* epochs_out, double* elapsed_out) {
    if (!s || !(seconds >= 0.0)) return RBP_ERR_INVALID;
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    uint64_t n = 0;
    while (elapsed() < seconds) {
        const int rc = rbp_solver_step(s, 8);  // a Kuhn / Leduc epoch is tens of microseconds: check the clock every 8
        if (rc != RBP_OK) return rc;
        n += 8;
    }
    if (epochs_out) *epochs_out = n;
    if (elapsed_out) *elapsed_out = elapsed();
    return RBP_OK;
}
int rbp_solver_step_timed(rbp_solver_t* s, uint64_t n_epochs, int flush_l2, float* ms_total, float* ms_sample, float* ms_fold) {
    if (!s || !ms_total) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(s->device));
    const size_t flush_n = (size_t)192 << 20 >> 4;  // 192 MiB of uint4
    if (flush_l2 && !s->flush_buf) {
        int st = alloc(s, flush_n, &s->flush_buf);
        if (st) return st;
    }
    while (s->events.size() < 3 * n_epochs) {
        cudaEvent_t e;
        RBP_CUDA(cudaEventCreate(&e));
        s->events.push_back(e);
    }
    for (uint64_t i = 0; i < n_epochs; ++i) {
        EpochArgs ep = epoch_args(s);
        int st;
        if (flush_l2) {
            l2_flush_kernel<<<148 * 8, 256, 0, s->stream>>>(s->flush_buf, flush_n);
            RBP_CUDA(cudaGetLastError());
        }
        RBP_CUDA(cudaEventRecord(s->events[3 * i], s->stream));
        if ((st = launch_sample(s, ep))) return st;
        RBP_CUDA(cudaEventRecord(s->events[3 * i + 1], s->stream));
        if (s->fold_mode == RBP_FOLD_BATCHED) {
            if (s->world_size > 1) return RBP_ERR_STATE;
            if ((st = launch_rank_partial(s))) return st;
            if ((st = launch_apply_batched(s, ep, s->delta, 1))) return st;
        } else if ((st = launch_fold(s, ep))) return st;
        RBP_CUDA(cudaEventRecord(s->events[3 * i + 2], s->stream));
        s->epochs += 1;
    }
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    double tot = 0, a = 0, b = 0;
    for (uint64_t i = 0; i < n_epochs; ++i) {
        float x;
        RBP_CUDA(cudaEventElapsedTime(&x, s->events[3 * i], s->events[3 * i + 2])); tot += x;
        RBP_CUDA(cudaEventElapsedTime(&x, s->events[3 * i], s->events[3 * i + 1])); a += x;
        RBP_CUDA(cudaEventElapsedTime(&x, s->events[3 * i + 1], s->events[3 * i + 2])); b += x;
    }
    *ms_total = (float)tot;
    if (ms_sample) *ms_sample = (float)a;
    if (ms_fold) *ms_fold = (float)b;
    return RBP_OK;
}

int rbp_selftest_div_by_count(uint32_t max_count, uint32_t samples, uint64_t* mismatches) {
    if (!mismatches) return RBP_ERR_INVALID;
    if (rbp_device_count() < 1) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    unsigned long long* d = nullptr;
    RBP_CUDA(cudaMalloc(&d, sizeof *d));
    RBP_CUDA(cudaMemset(d, 0, sizeof *d));
    div_selftest_kernel<<<148 * 8, 256>>>(max_count, samples, d);
    RBP_LAUNCHED();
    unsigned long long h = 0;
    RBP_CUDA(cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost));
    RBP_CUDA(cudaFree(d));
    *mismatches = h;
    return RBP_OK;
}

int rbp_solver_epochs(rbp_solver_t* s, uint64_t* out) {
    if (!s || !out) return RBP_ERR_INVALID;
    *out = s->epochs;
    return RBP_OK;
}

int rbp_solver_exploitability(rbp_solver_t* s, float* out) {
    if (!s || !out) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(s->device));
    size_t smem = (size_t)s->dev.n_infos * kMaxActions * sizeof(float);
    mccfr_exploit_kernel<<<1, 256, smem, s->stream>>>(s->dev, s->table, s->U, s->cfv, s->br, s->expl);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemcpyAsync(out, s->expl, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    return RBP_OK;
}

int rbp_solver_counters(rbp_solver_t* s, uint64_t out[3]) {
    if (!s || !out) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(s->device));
    unsigned long long tmp[3];
    RBP_CUDA(cudaMemcpyAsync(tmp, s->sc.counters, sizeof tmp, cudaMemcpyDeviceToHost, s->stream));
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < 3; ++i) out[i] = tmp[i];
    return RBP_OK;
}

int rbp_profile_export(rbp_solver_t* s, rbp_profile_row_t* rows, int cap, int* n_out) {
    if (!s || !n_out || (cap > 0 && !rows)) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(s->device));
    std::vector<rbp_encounter_t> host(s->game.n_rows);
    RBP_CUDA(cudaMemcpyAsync(host.data(), tbl(s), host.size() * sizeof(rbp_encounter_t), cudaMemcpyDeviceToHost, s->stream));
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<rbp_profile_row_t> all;
    for (size_t x = 0; x < s->game.info_key.size(); ++x)
        for (int a = 0; a < s->game.info_actions[x]; ++a) {
            int r = s->game.info_row[x] + a;
            if (host[r].visits > 0 || s->touched[r]) all.push_back(rbp_profile_row_t{s->game.info_key[x], (uint32_t)a, host[r]});
        }
    std::sort(all.begin(), all.end(), [](const rbp_profile_row_t& p, const rbp_profile_row_t& q) {
        return p.info_key != q.info_key ? p.info_key < q.info_key : p.action < q.action;
    });
    *n_out = (int)all.size();
    for (int i = 0; i < (int)all.size() && i < cap; ++i) rows[i] = all[i];
    return (int)all.size() > cap && cap > 0 ? RBP_ERR_CAPACITY : RBP_OK;
}

int rbp_profile_import(rbp_solver_t* s, const rbp_profile_row_t* rows, int n, uint64_t epochs) {
    if (!s || (n > 0 && !rows) || n < 0) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(s->device));
    std::vector<rbp_encounter_t> host(s->game.n_rows, rbp_encounter_t{0.0f, 0.0f, 0.0f, 0u});
    std::fill(s->touched.begin(), s->touched.end(), 0);
    for (int i = 0; i < n; ++i) {
        int x = -1;
        for (size_t k = 0; k < s->game.info_key.size(); ++k)
            if (s->game.info_key[k] == rows[i].info_key) { x = (int)k; break; }
        if (x < 0 || rows[i].action >= s->game.info_actions[x]) { set_last_error("import: unknown (info_key, action)"); return RBP_ERR_INVALID; }
        int r = s->game.info_row[x] + (int)rows[i].action;
        host[r] = rows[i].row;
        s->touched[r] = 1;
    }
    RBP_CUDA(cudaMemcpyAsync(s->table, host.data(), host.size() * sizeof(rbp_encounter_t), cudaMemcpyHostToDevice, s->stream));
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    s->epochs = epochs;
    return RBP_OK;
}

int rbp_profile_averaged(rbp_solver_t* s, uint32_t info_key, float* probs, int cap, int* n_out) {
    if (!s || !probs || !n_out) return RBP_ERR_INVALID;
    int x = -1;
    for (size_t k = 0; k < s->game.info_key.size(); ++k)
        if (s->game.info_key[k] == info_key) { x = (int)k; break; }
    if (x < 0) return RBP_ERR_INVALID;
    const int A = s->game.info_actions[x];
    if (cap < A) return RBP_ERR_CAPACITY;
    RBP_CUDA(cudaSetDevice(s->device));
    rbp_encounter_t e[kMaxActions];
    RBP_CUDA(cudaMemcpyAsync(e, tbl(s) + s->game.info_row[x], A * sizeof(rbp_encounter_t), cudaMemcpyDeviceToHost, s->stream));
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    float w[kMaxActions], sum = 0.0f;  // profile.rs:41-45 (a read-out of device rows, not a compute path)
    for (int a = 0; a < A; ++a) { w[a] = e[a].weight > kEps ? e[a].weight : kEps; sum = sum + w[a]; }
    for (int a = 0; a < A; ++a) probs[a] = w[a] / sum;
    *n_out = A;
    return RBP_OK;
}

int rbp_solver_game_shape(rbp_solver_t* s, int out[6]) {
    if (!s || !out) return RBP_ERR_INVALID;
    out[0] = (int)s->game.nodes.size(); out[1] = s->game.n_terminals; out[2] = (int)s->game.info_key.size();
    out[3] = s->game.n_rows; out[4] = s->game.max_tree_nodes; out[5] = s->game.max_tree_infos;
    return RBP_OK;
}

int rbp_solver_sample(rbp_solver_t* s) {
    if (!s) return RBP_ERR_INVALID;
    if (s->fold_mode != RBP_FOLD_BATCHED) { set_last_error("rbp_solver_sample needs RBP_FOLD_BATCHED"); return RBP_ERR_STATE; }
    RBP_CUDA(cudaSetDevice(s->device));
    EpochArgs ep = epoch_args(s);
    int st;
    if ((st = launch_sample(s, ep))) return st;
    if ((st = launch_rank_partial(s))) return st;
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    s->sampled = true;
    return RBP_OK;
}
int rbp_solver_delta_buffer(rbp_solver_t* s, void** dev_ptr, size_t* bytes) {
    if (!s || !dev_ptr || !bytes) return RBP_ERR_INVALID;
    if (s->fold_mode != RBP_FOLD_BATCHED) return RBP_ERR_STATE;
    *dev_ptr = s->delta;
    *bytes = (size_t)s->dev.n_infos * sizeof(Partial);
    return RBP_OK;
}
int rbp_solver_fold_gathered(rbp_solver_t* s, const void* dev_gathered, int world_size) {
    if (!s || !dev_gathered || world_size < 1) return RBP_ERR_INVALID;
    if (s->fold_mode != RBP_FOLD_BATCHED) return RBP_ERR_STATE;
    if (world_size != s->world_size) { set_last_error("rbp_solver_fold_gathered: world_size differs from rbp_solver_set_world"); return RBP_ERR_INVALID; }
    if (!s->sampled) { set_last_error("rbp_solver_fold_gathered before rbp_solver_sample"); return RBP_ERR_STATE; }
    RBP_CUDA(cudaSetDevice(s->device));
    EpochArgs ep = epoch_args(s);
    int st;
    if ((st = launch_apply_batched(s, ep, static_cast<const Partial*>(dev_gathered), world_size))) return st;
    s->sampled = false;
    s->epochs += 1;
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    return RBP_OK;
}

}  // extern "C"

// ───────────────────────────── safe subgame solving on the small games ─────────────────────────────
// `WorldSolver` (crates/subgame/src/world/solver.rs:33-146) = `SubGameSolver` without an origin (crates/subgame/src/solver.rs:46-146):
// every step samples a world from the belief, re-deals the external player's card for that world (`WorldRestrict::restrict`,
// kuhn/src/encoder.rs:47-66, leduc/src/encoder.rs:48-70), samples ONE ExternalSampling tree from that entry state and folds it with
// SummedRegret + LinearWeight into the world's table, which reads through to the frozen blueprint.  The epoch kernels are the
// training ones; what is new is the entry node, the per-world table offset and the read-through of unwritten weights.
// RNG contract: world = weighted(philox(step, 0, 0xFFFFFFFE, TAG_WORLD = 5)); the tree draws with tree id = world.
struct rbp_subgame {
    rbp_solver* local = nullptr;
    int worlds = 1, external = 1;
    float weights[8] = {};
    int32_t world_of_rank[3] = {-1, -1, -1};
    bool empty_belief = true;
    int entry[8] = {};       // flat node of every world's restricted entry state
    int cards[8][2] = {};    // its hole cards
    uint64_t seed = 0, drawn[8] = {};
    std::vector<rbp_encounter_t> blueprint_rows;  // host copy (Harvest and averaged read-outs fall through to it)
};

namespace {
constexpr uint32_t TAG_WORLD = 5;
// the flat node reached from the deal (c0, c1) by `path`; at a chance node the step names a CARD (the board survives a re-deal of a hole
// card, its index among the remaining cards does not)
int walk_entry(const FlatGame& G, int c0, int c1, const std::vector<int>& path_cards_or_actions, const std::vector<uint8_t>& is_card) {
    int node = G.root_table[c0 * G.deck + c1];
    if (node < 0) return -1;
    for (size_t i = 0; i < path_cards_or_actions.size(); ++i) {
        const FlatNode& nd = G.nodes[node];
        int idx = path_cards_or_actions[i];
        if (is_card[i]) {  // deals() order: Card::ALL ascending without the two holes
            const int card = idx;
            if (card == c0 || card == c1) return -1;
            idx = card - (card > c0 ? 1 : 0) - (card > c1 ? 1 : 0);
        }
        if (idx < 0 || idx >= nd.n_child) return -1;
        node = nd.first_child + idx;
    }
    return node;
}
}  // namespace

extern "C" {

// `Partition::partition::<W>` (crates/subgame/src/world/partition.rs:27-53): secrets (in ascending order) with their posterior reach ->
// the world of every secret (world 0 = highest reach) and the probability mass of every world.  Host arithmetic on <= a few hundred floats.
int rbp_subgame_partition(const float* reach, int n_secrets, int worlds, int32_t* world_of_secret, float* weights) {
    if (!reach || !world_of_secret || !weights || n_secrets < 1 || worlds < 1 || worlds > 8) return RBP_ERR_INVALID;
    float total = 0.0f;
    for (int i = 0; i < n_secrets; ++i) total += reach[i];
    for (int w = 0; w < worlds; ++w) weights[w] = 0.0f;
    if (total <= 0.0f) {
        for (int i = 0; i < n_secrets; ++i) world_of_secret[i] = 0;
        for (int w = 0; w < worlds; ++w) weights[w] = 1.0f / (float)worlds;
        return RBP_OK;
    }
    std::vector<int> order(n_secrets);
    for (int i = 0; i < n_secrets; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return reach[a] > reach[b]; });
    const float segment = total / (float)worlds;
    int index = 0;
    float bucket = 0.0f, accumulated = 0.0f;
    for (int i : order) {
        bucket += reach[i];
        accumulated += reach[i];
        world_of_secret[i] = index;
        if (accumulated >= segment * (float)(index + 1) && index < worlds - 1) {
            weights[index] = bucket / total;
            index += 1;
            bucket = 0.0f;
        }
    }
    weights[index] = bucket / total;
    return RBP_OK;
}

// Host half of `WorldSolver::new`: the restricted entry state of every world (`WorldRestrict::restrict` on the observed state) as hole
// cards, flat node and infoset key.  No device needed.
int rbp_subgame_entries(int game, int external, int worlds, const int32_t* world_of_rank, int c0, int c1, const uint8_t* path, int path_len,
                        int32_t* cards16, int32_t* nodes8, uint32_t* info_keys8) {
    if (worlds < 1 || worlds > 8 || (external != 0 && external != 1) || path_len < 0 || (path_len > 0 && !path)) return RBP_ERR_INVALID;
    FlatGame G;
    if (!build_flat_game(game, &G)) return RBP_ERR_INVALID;
    if (G.deck != 6 || c0 < 0 || c0 >= 6 || c1 < 0 || c1 >= 6 || c0 == c1) { set_last_error("subgame: Kuhn / Leduc and two distinct cards"); return RBP_ERR_INVALID; }
    // the observed state: follow the path once on the observed deal; a step at a chance node is remembered as the card it deals
    std::vector<int> steps(path_len);
    std::vector<uint8_t> is_card(path_len, 0);
    int board = -1;
    {
        int node = G.root_table[c0 * 6 + c1];
        for (int i = 0; i < path_len; ++i) {
            const FlatNode& nd = G.nodes[node];
            if (path[i] >= nd.n_child) { set_last_error("subgame: path leaves the game tree"); return RBP_ERR_INVALID; }
            steps[i] = path[i];
            if (nd.turn == TURN_CHANCE) {
                int seen = -1, card = -1;
                for (int c = 0; c < 6; ++c) { if (c == c0 || c == c1) continue; if (++seen == (int)path[i]) { card = c; break; } }
                steps[i] = card; is_card[i] = 1; board = card;
            }
            node = nd.first_child + path[i];
        }
        if (G.nodes[node].turn > TURN_P1) { set_last_error("subgame: the entry state is not a decision node"); return RBP_ERR_INVALID; }
    }
    bool empty_belief = true;
    for (int r = 0; r < 3; ++r) if (world_of_rank && world_of_rank[r] >= 0) empty_belief = false;
    const int observed[2] = {c0, c1};
    for (int w = 0; w < worlds; ++w) {  // WorldRestrict::restrict: the first free card whose rank the belief puts in world w
        int hole[2] = {c0, c1};
        for (int c = 0; c < 6; ++c) {
            if (c == observed[1 - external] || c == board) continue;
            if (empty_belief || world_of_rank[c >> 1] == w) { hole[external] = c; break; }
        }
        const int node = walk_entry(G, hole[0], hole[1], steps, is_card);
        if (node < 0) { set_last_error("subgame: the restricted entry state is not in the game tree"); return RBP_ERR_INVALID; }
        if (cards16) { cards16[2 * w] = hole[0]; cards16[2 * w + 1] = hole[1]; }
        if (nodes8) nodes8[w] = node;
        if (info_keys8) info_keys8[w] = G.nodes[node].info_key;
    }
    return RBP_OK;
}

// The action-conditioned posterior over the external player's rank at the observed state (kuhn/src/solver.rs
// `subgame_with_reach_conditioned_posterior`): for every card the external player could hold, `Solver::external_reach`
// (mccfr/src/solver/solver.rs:198-211) = the product, along the recall's path, of the blueprint's averaged policy at the external
// player's own decisions; summed per rank (`Posterior::add`).  Host arithmetic over the blueprint's exported rows; no device needed.
int rbp_subgame_posterior(int game, const rbp_profile_row_t* rows, int n_rows, int external, int c0, int c1, const uint8_t* path, int path_len,
                          float* reach3) {
    if (!reach3 || (n_rows > 0 && !rows) || n_rows < 0 || (external != 0 && external != 1) || path_len < 0 || (path_len > 0 && !path)) return RBP_ERR_INVALID;
    FlatGame G;
    if (!build_flat_game(game, &G)) return RBP_ERR_INVALID;
    if (G.deck != 6 || c0 < 0 || c0 >= 6 || c1 < 0 || c1 >= 6 || c0 == c1) { set_last_error("subgame: Kuhn / Leduc and two distinct cards"); return RBP_ERR_INVALID; }
    std::vector<float> weight(G.n_rows, 0.0f);  // RefProf::cum_weight, 0 where the blueprint holds no row
    for (int i = 0; i < n_rows; ++i) {
        int x = -1;
        for (size_t k = 0; k < G.info_key.size(); ++k) if (G.info_key[k] == rows[i].info_key) { x = (int)k; break; }
        if (x < 0 || rows[i].action >= G.info_actions[x]) { set_last_error("subgame: unknown (info_key, action) in the blueprint rows"); return RBP_ERR_INVALID; }
        weight[G.info_row[x] + (int)rows[i].action] = rows[i].row.weight;
    }
    std::vector<int> steps(path_len);
    std::vector<uint8_t> is_card(path_len, 0);
    int board = -1;
    {
        int node = G.root_table[c0 * 6 + c1];
        for (int i = 0; i < path_len; ++i) {
            const FlatNode& nd = G.nodes[node];
            if (path[i] >= nd.n_child) { set_last_error("subgame: path leaves the game tree"); return RBP_ERR_INVALID; }
            steps[i] = path[i];
            if (nd.turn == TURN_CHANCE) {
                int seen = -1, card = -1;
                for (int c = 0; c < 6; ++c) { if (c == c0 || c == c1) continue; if (++seen == (int)path[i]) { card = c; break; } }
                steps[i] = card; is_card[i] = 1; board = card;
            }
            node = nd.first_child + path[i];
        }
    }
    const int observed[2] = {c0, c1};
    for (int r = 0; r < 3; ++r) reach3[r] = 0.0f;
    for (int c = 0; c < 6; ++c) {
        if (c == observed[1 - external] || c == board) continue;
        int hole[2] = {c0, c1};
        hole[external] = c;
        int node = G.root_table[hole[0] * 6 + hole[1]];
        float reach = 1.0f;  // Iterator::product over f32
        for (int i = 0; i < path_len && node >= 0; ++i) {
            const FlatNode& nd = G.nodes[node];
            int idx = steps[i];
            if (is_card[i]) idx = idx - (idx > hole[0] ? 1 : 0) - (idx > hole[1] ? 1 : 0);
            if (nd.turn == external) {  // averaged_policy(info, edge) = max(w, EPS) / sum (profile.rs:40-44)
                const int A = G.info_actions[nd.info], row = G.info_row[nd.info];
                float sum = 0.0f;
                for (int a = 0; a < A; ++a) sum = sum + (weight[row + a] > kEps ? weight[row + a] : kEps);
                const float w = weight[row + idx] > kEps ? weight[row + idx] : kEps;
                reach = reach * (w / sum);
            }
            node = nd.first_child + idx;
        }
        reach3[c >> 1] += reach;
    }
    return RBP_OK;
}

int rbp_subgame_create(rbp_solver_t* blueprint, int external, int worlds, const int32_t* world_of_rank, const float* weights,
                       int c0, int c1, const uint8_t* path, int path_len, uint64_t seed, rbp_subgame_t** out) {
    if (!out) return RBP_ERR_INVALID;
    *out = nullptr;
    if (!blueprint || !weights) return RBP_ERR_INVALID;
    const FlatGame& G = blueprint->game;
    if (blueprint->sub_worlds) { set_last_error("rbp_subgame_create: the blueprint is itself a subgame"); return RBP_ERR_INVALID; }
    int game_id = -1;
    for (int id = 0; id < 3 && game_id < 0; ++id) { FlatGame probe; if (build_flat_game(id, &probe) && probe.nodes.size() == G.nodes.size() && probe.n_rows == G.n_rows && probe.deck == G.deck) game_id = id; }
    if (game_id < 0) return RBP_ERR_INVALID;
    std::unique_ptr<rbp_subgame> g(new rbp_subgame());
    g->worlds = worlds; g->external = external; g->seed = seed;
    {
        int32_t cards16[16], nodes8[8];
        const int st0 = rbp_subgame_entries(game_id, external, worlds, world_of_rank, c0, c1, path, path_len, cards16, nodes8, nullptr);
        if (st0) return st0;
        for (int w = 0; w < worlds; ++w) { g->weights[w] = weights[w]; g->entry[w] = nodes8[w]; g->cards[w][0] = cards16[2 * w]; g->cards[w][1] = cards16[2 * w + 1]; }
        for (int r = 0; r < 3; ++r) { g->world_of_rank[r] = world_of_rank ? world_of_rank[r] : -1; if (g->world_of_rank[r] >= 0) g->empty_belief = false; }
    }
    rbp_solver_t* local = nullptr;
    int st = rbp_solver_create(game_id, RBP_REGRET_SUMMED, RBP_WEIGHT_LINEAR, RBP_SAMPLING_EXTERNAL, RBP_FOLD_ORDERED, 1, seed, &blueprint->hyper, blueprint->device, &local);
    if (st) return st;
    g->local = local;
    auto fail = [&](int code) { rbp_solver_destroy(local); return code; };
    if (cudaSetDevice(local->device) != cudaSuccess) return fail(RBP_ERR_CUDA);
    rbp_encounter_t* big = nullptr;
    if ((st = alloc(local, (size_t)worlds * G.n_rows, &big))) return fail(st);
    if ((st = alloc(local, (size_t)2 * G.n_rows, &local->fb_weight))) return fail(st);  // [weights | payoffs]
    // the blueprint's rows: a host copy for the read-outs, a device copy (on the local stream) to seed from
    g->blueprint_rows.resize(G.n_rows);
    if (cudaStreamSynchronize(blueprint->stream) != cudaSuccess) return fail(RBP_ERR_CUDA);
    if (cudaMemcpy(g->blueprint_rows.data(), blueprint->table, (size_t)G.n_rows * sizeof(rbp_encounter_t), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(RBP_ERR_CUDA);
    rbp_encounter_t* bp_dev = local->table;  // the n_rows table rbp_solver_create made: scratch for the blueprint copy
    if (cudaMemcpyAsync(bp_dev, g->blueprint_rows.data(), (size_t)G.n_rows * sizeof(rbp_encounter_t), cudaMemcpyHostToDevice, local->stream) != cudaSuccess) return fail(RBP_ERR_CUDA);
    const float k = 16384.0f;  // WarmstartHyperParams::default().prior_strength = 1 << 14 (hyperparams/warmstart.rs:27-33)
    subgame_seed_kernel<<<(local->dev.n_infos + 127) / 128, 128, 0, local->stream>>>(local->dev, bp_dev, k, worlds, big, local->fb_weight);
    g_launches.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(local->stream) != cudaSuccess) return fail(RBP_ERR_CUDA);
    local->table = big;
    local->sub_worlds = worlds;
    local->sub_world = 0;
    local->sub_entry_plus1 = g->entry[0] + 1;
    *out = g.release();
    return RBP_OK;
}
void rbp_subgame_destroy(rbp_subgame_t* g) {
    if (!g) return;
    rbp_solver_destroy(g->local);
    delete g;
}
// `WorldSolver::step` n times (world/solver.rs:118-146); one stream synchronisation at the end
int rbp_subgame_step(rbp_subgame_t* g, uint64_t n) {
    if (!g) return RBP_ERR_INVALID;
    rbp_solver* s = g->local;
    RBP_CUDA(cudaSetDevice(s->device));
    for (uint64_t i = 0; i < n; ++i) {
        const Philox4 p = philox4x32_10((uint32_t)s->epochs, 0u, 0xFFFFFFFEu, TAG_WORLD, (uint32_t)g->seed, (uint32_t)(g->seed >> 32));
        float total = 0.0f;  // weighted() of the RNG contract over the belief weights
        for (int w = 0; w < g->worlds; ++w) total = total + g->weights[w];
        const float x = draw_unit(p.r[0]) * total;
        float cum = 0.0f;
        int world = g->worlds - 1;
        for (int w = 0; w < g->worlds; ++w) { cum = cum + g->weights[w]; if (x < cum) { world = w; break; } }
        g->drawn[world] += 1;
        s->sub_world = world;
        s->sub_entry_plus1 = g->entry[world] + 1;
        const EpochArgs ep = epoch_args(s);
        int st;
        if ((st = launch_sample(s, ep))) return st;
        if ((st = launch_fold(s, ep))) return st;
        s->epochs += 1;  // WorldProfile::increment
    }
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    return RBP_OK;
}
// `Solver::spend` (mccfr/src/solver/solver.rs:130-137) for the real-time caller
int rbp_subgame_spend(rbp_subgame_t* g, double seconds, uint64_t* steps_out, double* elapsed_out) {
    if (!g || !(seconds >= 0.0)) return RBP_ERR_INVALID;
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    uint64_t n = 0;
    while (elapsed() < seconds) {
        const int rc = rbp_subgame_step(g, 8);
        if (rc != RBP_OK) return rc;
        n += 8;
    }
    if (steps_out) *steps_out = n;
    if (elapsed_out) *elapsed_out = elapsed();
    return RBP_OK;
}
int rbp_subgame_info(rbp_subgame_t* g, uint64_t* steps, uint64_t* drawn8, int32_t* entry_cards16) {
    if (!g) return RBP_ERR_INVALID;
    if (steps) *steps = g->local->epochs;
    if (drawn8) for (int w = 0; w < 8; ++w) drawn8[w] = g->drawn[w];
    if (entry_cards16) for (int w = 0; w < 8; ++w) { entry_cards16[2 * w] = g->cards[w][0]; entry_cards16[2 * w + 1] = g->cards[w][1]; }
    return RBP_OK;
}
// the rows of one world that the subgame has written (the reference's local HashMap), sorted by (info_key, action)
int rbp_subgame_export(rbp_subgame_t* g, int world, rbp_profile_row_t* rows, int cap, int* n_out) {
    if (!g || world < 0 || world >= g->worlds) return RBP_ERR_INVALID;
    const int keep = g->local->sub_world;
    g->local->sub_world = world;
    const int st = rbp_profile_export(g->local, rows, cap, n_out);
    g->local->sub_world = keep;
    return st;
}
namespace {
int subgame_rows(rbp_subgame_t* g, uint32_t info_key, int* A_out, rbp_encounter_t (*e)[kMaxActions]) {  // [world][action], fall-through applied
    rbp_solver* s = g->local;
    int x = -1;
    for (size_t k = 0; k < s->game.info_key.size(); ++k) if (s->game.info_key[k] == info_key) { x = (int)k; break; }
    if (x < 0) return RBP_ERR_INVALID;
    const int A = s->game.info_actions[x], row = s->game.info_row[x];
    RBP_CUDA(cudaSetDevice(s->device));
    for (int w = 0; w < g->worlds; ++w)
        RBP_CUDA(cudaMemcpyAsync(e[w], s->table + (size_t)w * s->game.n_rows + row, A * sizeof(rbp_encounter_t), cudaMemcpyDeviceToHost, s->stream));
    RBP_CUDA(cudaStreamSynchronize(s->stream));
    for (int w = 0; w < g->worlds; ++w)
        for (int a = 0; a < A; ++a)
            if (e[w][a].visits == 0u) {  // never written: world/profile.rs:118-145
                const rbp_encounter_t& b = g->blueprint_rows[row + a];
                e[w][a] = rbp_encounter_t{b.weight > kEps ? b.weight : kEps, b.regret > kEps ? b.regret : kEps, b.payoff, b.visits};
            }
    *A_out = A;
    return RBP_OK;
}
}  // namespace
// `CfrNash::averaged_policy` over `WorldInfo(world, info)` (profile.rs:40-44 on the WorldProfile)
int rbp_subgame_averaged(rbp_subgame_t* g, int world, uint32_t info_key, float* probs, int cap, int* n_out) {
    if (!g || !probs || !n_out || world < 0 || world >= g->worlds) return RBP_ERR_INVALID;
    rbp_encounter_t e[8][kMaxActions];
    int A = 0, st;
    if ((st = subgame_rows(g, info_key, &A, e))) return st;
    if (cap < A) return RBP_ERR_CAPACITY;
    float w[kMaxActions], sum = 0.0f;
    for (int a = 0; a < A; ++a) { w[a] = e[world][a].weight > kEps ? e[world][a].weight : kEps; sum = sum + w[a]; }
    for (int a = 0; a < A; ++a) probs[a] = w[a] / sum;
    *n_out = A;
    return RBP_OK;
}
// `Harvest::harvest` (world/solver.rs:148-191) at a base infoset: refined[a] = sum over worlds of the iterated (regret-matching) policy / W,
// visits[a] = sum of cum_visits, regret = sum over edges and worlds of max(cum_regret, 0)
int rbp_subgame_harvest(rbp_subgame_t* g, uint32_t info_key, float* refined, uint32_t* visits, float* regret, int cap, int* n_out) {
    if (!g || !refined || !visits || !regret || !n_out) return RBP_ERR_INVALID;
    rbp_encounter_t e[8][kMaxActions];
    int A = 0, st;
    if ((st = subgame_rows(g, info_key, &A, e))) return st;
    if (cap < A) return RBP_ERR_CAPACITY;
    for (int a = 0; a < A; ++a) { refined[a] = 0.0f; visits[a] = 0u; }
    for (int w = 0; w < g->worlds; ++w) {
        float r[kMaxActions], rd = 0.0f;
        for (int a = 0; a < A; ++a) { r[a] = e[w][a].regret > kEps ? e[w][a].regret : kEps; rd = rd + r[a]; }
        for (int a = 0; a < A; ++a) refined[a] += r[a] / rd / (float)g->worlds;
    }
    float reg = 0.0f;
    for (int a = 0; a < A; ++a)
        for (int w = 0; w < g->worlds; ++w) { visits[a] += e[w][a].visits; reg += e[w][a].regret > 0.0f ? e[w][a].regret : 0.0f; }
    *regret = reg;
    *n_out = A;
    return RBP_OK;
}

}  // extern "C"
