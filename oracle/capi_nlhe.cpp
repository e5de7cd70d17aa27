// ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes-facing C API over oracle/nlhe.hpp.
#include <algorithm>
#include <tuple>

#include "nlhe.hpp"

using namespace orc;
using namespace orc::nlhe;

extern "C" {

struct OrcNlheProbe {  // one row per scripted step (row 0 = root)
    int16_t pot, to_call, to_raise, to_shove;
    int16_t stack[2], stake[2], spent[2];
    int8_t street, turn;  // turn: 0/1 seat, 2 chance, 3 terminal
    uint8_t applied_kind, pad;
    int16_t applied_chips;
    uint16_t abs;
    uint32_t flags;
    uint64_t subgame, choices, board, hole[2];
};
enum : uint32_t {
    F_MUST_POST = 1, F_MUST_STOP = 2, F_MUST_DEAL = 4, F_ALRIGHT = 8, F_CALLING = 16, F_TOUCHED = 32, F_MATCHED = 64,
    F_FOLDING = 128, F_SHOVING = 256, F_MAY_FOLD = 512, F_MAY_CALL = 1024, F_MAY_CHECK = 2048, F_MAY_RAISE = 4096, F_MAY_SHOVE = 8192,
};

static void fill(OrcNlheProbe* p, const State& s) {
    const Game& g = s.game;
    p->pot = g.pot; p->street = (int8_t)g.street(); p->turn = (int8_t)g.turn();
    p->to_call = g.to_call(); p->to_raise = g.to_raise(); p->to_shove = g.to_shove();
    for (int i = 0; i < 2; ++i) { p->stack[i] = g.seats[i].stack; p->stake[i] = g.seats[i].stake; p->spent[i] = g.seats[i].spent; p->hole[i] = g.seats[i].cards; }
    p->board = g.board;
    uint32_t f = 0;
    if (g.must_post()) f |= F_MUST_POST;
    if (g.must_stop()) f |= F_MUST_STOP;
    if (g.must_deal()) f |= F_MUST_DEAL;
    if (g.everyone_alright()) f |= F_ALRIGHT;
    if (g.everyone_calling()) f |= F_CALLING;
    if (g.everyone_touched()) f |= F_TOUCHED;
    if (g.everyone_matched()) f |= F_MATCHED;
    if (g.everyone_folding()) f |= F_FOLDING;
    if (g.everyone_shoving()) f |= F_SHOVING;
    if (g.may_fold()) f |= F_MAY_FOLD;
    if (g.may_call()) f |= F_MAY_CALL;
    if (g.may_check()) f |= F_MAY_CHECK;
    if (g.may_raise()) f |= F_MAY_RAISE;
    if (g.may_shove()) f |= F_MAY_SHOVE;
    p->flags = f;
    const Info i = info_of(s);
    p->subgame = i.subgame; p->choices = i.choices; p->abs = i.abs;
}

// kicker game.rs:282-305 is_allowed
static bool allowed(const Game& g, const Action& a) {
    if (a.kind == A_RAISE) return g.may_raise() && !g.must_stop() && !g.must_deal() && a.chips >= g.to_raise() && a.chips < g.to_shove();
    if (a.kind == A_DRAW) return g.must_deal() && !g.must_stop();
    if (g.must_stop() || g.must_deal()) return false;
    switch (a.kind) {
        case A_SHOVE: return g.may_shove() && a.chips == g.to_shove();
        case A_CALL: return g.may_call() && a.chips == g.to_call();
        case A_FOLD: return g.may_fold();
        case A_CHECK: return g.may_check();
        default: return false;
    }
}

// Steps: kind 0 Draw (dealt by the contract's RNG) · 1 Fold · 2 Call(chips) · 3 Check · 4 Raise(chips) · 5 Shove(chips);
// chips < 0 means "the legal amount" (to_call / to_raise / to_shove); kind 16+e applies abstract edge e through NlheGame::apply.
// mode bit0: snap the action first (Game::snap).  Returns the number of steps applied, or -(i+1) if step i was not allowed.
int orc_nlhe_script(uint64_t seed, uint32_t epoch, uint32_t tree, int n, const int32_t* kinds, const int32_t* chips, int mode,
                    OrcNlheProbe* out, int16_t* won2, int16_t* reward2) {
    const Ctx cx{Draw{seed}, epoch, tree};
    State s = root(cx);
    fill(&out[0], s);
    for (int i = 0; i < n; ++i) {
        if (kinds[i] >= 16) {
            s = apply(s, (uint8_t)(kinds[i] - 16), cx);
            out[i + 1].applied_kind = 0xFF;
        } else {
            Action a{};
            s.hist = mix64(s.hist ^ (uint64_t)(0x100 + kinds[i]));
            switch (kinds[i]) {
                case 0: a = reveal(s.game, cx, s.hist); break;
                case 1: a = Action{A_FOLD, 0, 0}; break;
                case 2: a = Action{A_CALL, chips[i] < 0 ? s.game.to_call() : (Chips)chips[i], 0}; break;
                case 3: a = Action{A_CHECK, 0, 0}; break;
                case 4: a = Action{A_RAISE, chips[i] < 0 ? s.game.to_raise() : (Chips)chips[i], 0}; break;
                default: a = Action{A_SHOVE, chips[i] < 0 ? s.game.to_shove() : (Chips)chips[i], 0}; break;
            }
            if (mode & 1) a = s.game.snap(a);
            if (!allowed(s.game, a)) return -(i + 1);
            s.game.force_act(a);
            out[i + 1].applied_kind = a.kind;
            out[i + 1].applied_chips = a.chips;
        }
        fill(&out[i + 1], s);
    }
    if (s.game.must_stop() && won2) {
        Chips w[2];
        s.game.settle(w);
        for (int i = 0; i < 2; ++i) { won2[i] = w[i]; if (reward2) reward2[i] = (Chips)(w[i] + s.game.seats[i].spent); }
    }
    return n;
}

// Showdown::settle on an explicit ledger (showdown.rs tests): risked, status (0 betting 1 shoving 2 folding), packed strength
void orc_nlhe_showdown(int n, const int16_t* risked, const uint8_t* status, const uint32_t* strength, int16_t* reward) {
    showdown(n, risked, status, strength, reward);
}
int orc_nlhe_aggression(uint64_t path) { return path_aggression(path); }
uint64_t orc_nlhe_path_push(uint64_t path, int edge) { return path_push(path, (uint8_t)edge); }
int orc_nlhe_raises(int street, int depth, uint8_t* out) { return raises(street, depth, out); }
int orc_nlhe_into_chips(int edge, int pot) { return into_chips((uint8_t)edge, (Chips)pot); }
uint64_t orc_nlhe_edge_u64(int edge) { return e_to_u64((uint8_t)edge); }
int orc_nlhe_edge_from_u64(uint64_t v) { return e_from_u64(v); }
float orc_nlhe_default_regret(int edge) { return e_default_regret((uint8_t)edge); }
int orc_nlhe_deck_draw(uint64_t* deck, uint32_t word) { return deck_draw(deck, word); }
uint16_t orc_nlhe_abstraction(uint64_t pocket, uint64_t board) {
    Game g{};
    g.board = board; g.seats[0].cards = pocket; g.dealer = 0; g.ticker = 0;
    return abstraction_of(g);
}

struct OrcNlheRow {  // the reference's blueprint row (nlhe/src/profile.rs:143-160)
    int64_t past;
    int64_t choices;
    int64_t edge;
    int16_t present;
    int16_t pad[3];
    float weight, regret, payoff;
    int32_t visits;
};

nlhe::Solver* orc_nlhe_create(uint64_t seed, int batch, int threads, int regret, int weight, int sampling) {
    nlhe::Solver* s = new nlhe::Solver();
    s->rng.seed = seed; s->batch = batch; s->threads = threads;
    s->regret_sched = regret; s->weight_sched = weight; s->sampling = sampling;
    return s;
}
void orc_nlhe_destroy(nlhe::Solver* s) { delete s; }
void orc_nlhe_set_hyper(nlhe::Solver* s, const float* f5, uint32_t warmup) {
    s->hyper.temperature = f5[0]; s->hyper.smoothing = f5[1]; s->hyper.curiosity = f5[2];
    s->hyper.prune_threshold = f5[3]; s->hyper.prune_explore = f5[4]; s->hyper.prune_warmup = warmup;
}
void orc_nlhe_set_world(nlhe::Solver* s, int rank, int world) { s->world_rank = rank; s->world_size = world; }
int orc_nlhe_dec_bytes() { return (int)sizeof(Dec); }
// multi-rank exchange: this rank's Decisions as raw bytes (call with out = NULL to size), then the fold of every rank's
uint64_t orc_nlhe_sample(nlhe::Solver* s, void* out, uint64_t cap) {
    static thread_local std::vector<Dec> pending;
    if (!out) { pending = s->sample_decs(); return pending.size(); }
    const uint64_t n = std::min<uint64_t>(cap, pending.size());
    std::memcpy(out, pending.data(), n * sizeof(Dec));
    return n;
}
void orc_nlhe_fold(nlhe::Solver* s, const void* decs, uint64_t count) {
    std::vector<Dec> v(count);
    std::memcpy(v.data(), decs, count * sizeof(Dec));
    s->fold_decs(v);
}
// owner-sharded exchange: owner of a Decisions' infoset, the rows the last fold touched, and their installation
int orc_nlhe_owner(const void* dec, int world) { return (int)((InfoHash()(static_cast<const Dec*>(dec)->info) >> 40) % (uint64_t)world); }
struct OrcNlhePacked { Info info; Row row; };
int orc_nlhe_packed_bytes() { return (int)sizeof(OrcNlhePacked); }
uint64_t orc_nlhe_touched(nlhe::Solver* s, void* out, uint64_t cap) {
    if (out)
        for (uint64_t i = 0; i < s->touched.size() && i < cap; ++i) {
            OrcNlhePacked p{};
            p.info = s->touched[i]; p.row = s->rows.at(s->touched[i]);
            std::memcpy(static_cast<char*>(out) + i * sizeof(OrcNlhePacked), &p, sizeof(p));
        }
    return s->touched.size();
}
void orc_nlhe_apply_rows(nlhe::Solver* s, const void* rows, uint64_t count) {
    for (uint64_t i = 0; i < count; ++i) {
        OrcNlhePacked p;
        std::memcpy(&p, static_cast<const char*>(rows) + i * sizeof(OrcNlhePacked), sizeof(p));
        s->rows[p.info] = p.row;
    }
}
void orc_nlhe_step(nlhe::Solver* s, uint64_t n) { for (uint64_t i = 0; i < n; ++i) s->step(); }
void orc_nlhe_counters(nlhe::Solver* s, uint64_t* out5) {
    out5[0] = s->epochs; out5[1] = s->nodes; out5[2] = s->infos; out5[3] = s->updates; out5[4] = s->rows.size();
}
// rows sorted by (past, present, choices, edge position)
uint64_t orc_nlhe_export(nlhe::Solver* s, OrcNlheRow* out, uint64_t cap) {
    std::vector<const std::pair<const Info, Row>*> v;
    for (const auto& kv : s->rows) v.push_back(&kv);
    std::sort(v.begin(), v.end(), [](auto a, auto b) {
        return std::tie(a->first.subgame, a->first.abs, a->first.choices) < std::tie(b->first.subgame, b->first.abs, b->first.choices);
    });
    uint64_t k = 0;
    for (auto p : v)
        for (int a = 0; a < p->second.n; ++a) {
            if (!p->second.present[a]) continue;
            if (out && k < cap) {
                const Encounter& e = p->second.e[a];
                out[k] = OrcNlheRow{(int64_t)p->first.subgame, (int64_t)p->first.choices, (int64_t)e_to_u64(p->second.edges[a]), (int16_t)p->first.abs, {0, 0, 0},
                                    e.weight, e.regret, e.payoff, (int32_t)e.visits};
            }
            ++k;
        }
    return k;
}
// one sampled tree of the CURRENT epoch, for debugging the device builder: per node (parent, edge, turn, pot, subgame, choices, abs)
struct OrcNlheNode { int32_t parent; uint8_t edge, turn; int16_t pot; uint16_t abs; uint16_t pad; uint64_t subgame, choices; float payoff0; int32_t pad2; };
int orc_nlhe_tree(nlhe::Solver* s, int tree, OrcNlheNode* out, int cap) {
    const nlhe::Solver::TreeN t = s->build(tree);
    const int n = (int)t.game.size();
    for (int i = 0; i < n && i < cap; ++i) {
        const Game& g = t.game[i].game;
        out[i] = OrcNlheNode{t.parent[i], t.incoming[i], (uint8_t)g.turn(), g.pot, t.info[i].abs, 0, t.info[i].subgame, t.info[i].choices,
                             g.turn() == Game::T_TERMINAL ? payoff(t.game[i], 0) : 0.0f, 0};
    }
    return n;
}

// the same tree in the device builder's layout: preorder with children in choices() order; per node depth, kind
// (0 walker 1 opponent 2 chance 3 terminal), action index in the parent's choices, policy p / sampling q of the incoming
// edge (flow.rs:20-44), terminal payoff for the walker
struct OrcNlhePre { uint8_t depth, kind, act, pad; float p, q, payoff; };
int orc_nlhe_tree_preorder(nlhe::Solver* s, int tree, OrcNlhePre* out, int cap) {
    const nlhe::Solver::TreeN t = s->build(tree);
    const int walker = s->walker();
    int n = 0;
    struct Item { int node, depth; };
    std::vector<Item> stack{{0, 0}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const Game& g = t.game[it.node].game;
        const int turn = g.turn();
        OrcNlhePre o{(uint8_t)it.depth, (uint8_t)(turn == Game::T_TERMINAL ? 3 : (turn == Game::T_CHANCE ? 2 : (turn == walker ? 0 : 1))), 0, 0, 1.0f, 1.0f, 0.0f};
        if (turn == Game::T_TERMINAL) o.payoff = payoff(t.game[it.node], walker);
        const int par = t.parent[it.node];
        if (par >= 0) {
            const int pt = t.game[par].game.turn();
            if (pt < 2) {
                const View v = s->view(t.info[par]);
                const int a = nlhe::Solver::act_of(v, t.incoming[it.node]);
                o.act = (uint8_t)a;
                o.p = v.r[a] / v.rd;
                if (pt != walker) o.q = v.sw[a] / v.z;
            }
        }
        if (n < cap) out[n] = o;
        ++n;
        std::vector<int> kids;
        for (int c = t.head[it.node]; c >= 0; c = t.next[c]) kids.push_back(c);
        for (int k = (int)kids.size() - 1; k >= 0; --k) stack.push_back(Item{kids[k], it.depth + 1});
    }
    return n;
}

}  // extern "C"
