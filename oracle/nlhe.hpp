// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).  CPU restatement of the heads-up no-limit hold'em plug-in
// of the MCCFR path: the `kicker::Game` state machine, the Pluribus action grid, `Path`/`Edge` packing, showdown
// settlement, and the `nlhe` crate's game/info/encoder glue.
//
// Follows (crates/…):
//   pokerkit/src/lib.rs:60-160        N=2, STACK=200, blinds 1/2, MAX_RAISE_REPEATS=3, OPENS, RAISES, PLURIBUS_INDICES
//   kicker/src/game.rs:54-86,148-320,385-720,721-855   preblind/root, turn, legal, act/bet/fold/show, next_player,
//                                     must_* / may_* / to_*, settlements, choices/unfold/actionize/snap
//   kicker/src/edge.rs:6-135          Edge, default regrets (bias.rs:40-67: fold 100, raise 10, shove 0, other 50),
//                                     raises(), into_chips(), u8 code
//   kicker/src/path.rs:6-20,138-175   Path: 5-bit edges, first edge lowest, ≤12 edges; aggression()
//   kicker/src/size.rs:112-162        grid row = street*3 + min(depth,2); depth > 3 → no raises
//   kicker/src/seat.rs, showdown.rs:23-110, settlement.rs, pnl.rs     chips bookkeeping and pot distribution
//   deuce/src/deck.rs:22-46           Deck::draw (its bias — lowest card twice as likely, highest never — is kept)
//   nlhe/src/game.rs:25-63            NlheGame::{apply (auto-reveal, snap), payoff}
//   nlhe/src/info.rs:62-108, encoder.rs:15-52, public.rs   info = (current-street subgame Path, choices Path, Abstraction)
//
// RNG contract additions (the reference draws hole cards and boards from the thread RNG):
//   hole cards: Philox(epoch, tree, 0xFFFFFFFF, TAG_ROOT) words 0..3 → Deck::draw indices range(52), 51, 50, 49
//   boards:     Philox(epoch, tree, lo32(hist), TAG_DRAW) words 0.. → one Deck::draw per card, hist = running hash of
//               every edge applied since the root (so different branches of one tree see independent boards)
//   infoset word for node draws: lo32(mix64(subgame ^ mix64(choices ^ mix64(abs))))
// Abstraction lookup (`NlheEncoder::abstraction`, a BTreeMap<Isomorphism, Abstraction> loaded from the clustering
// stage): SURVEY §8d config 4 prescribes a synthetic table — bucket = mix64(canonical pocket, canonical public) mod K,
// K = 169 / 256 / 256 / 101 per street; `Abstraction` = street << 8 | bucket (kicker/src/abstraction.rs:50-56).
#pragma once
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

#include "iso.hpp"
#include "mccfr.hpp"
#include "pool.hpp"

namespace orc {
namespace nlhe {

using Chips = int16_t;
constexpr Chips kStack = 200, kBB = 2, kSB = 1;
constexpr int kMaxRaiseRepeats = 3, kMaxPathEdges = 12, kMaxEdges = 10;
enum : uint32_t { TAG_DRAW = 4 };
enum EdgeCode : uint8_t { E_DRAW = 1, E_FOLD = 2, E_CHECK = 3, E_CALL = 4, E_SHOVE = 5, E_OPEN0 = 6, E_RAISE0 = 10 };
constexpr Chips kOpens[4] = {2, 3, 4, 5};
constexpr Chips kRaises[10][2] = {{1, 4}, {1, 3}, {1, 2}, {2, 3}, {3, 4}, {1, 1}, {5, 4}, {3, 2}, {2, 1}, {3, 1}};
constexpr int kGridLen[12] = {0, 2, 1, 5, 2, 1, 4, 2, 1, 4, 2, 1};
constexpr int kGrid[12][5] = {{}, {5, 8}, {5}, {0, 2, 4, 5, 8}, {2, 5}, {5}, {1, 2, 5, 8}, {5, 8}, {5}, {1, 2, 5, 8}, {5, 8}, {5}};

inline bool e_is_choice(uint8_t e) { return e != E_DRAW; }
inline bool e_is_aggro(uint8_t e) { return e >= E_SHOVE; }
inline float e_default_regret(uint8_t e) {  // edge.rs:41-53 + bias.rs defaults
    if (e >= E_OPEN0) return 10.0f;
    if (e == E_CHECK || e == E_CALL) return 50.0f;
    if (e == E_SHOVE) return 0.0f;
    return 100.0f;  // Fold
}
inline int raises(int street, int depth, uint8_t* out) {  // edge.rs:79-88, size.rs:136-153
    if (depth > kMaxRaiseRepeats) return 0;
    if (street == 0 && depth == 0) { for (int i = 0; i < 4; ++i) out[i] = (uint8_t)(E_OPEN0 + i); return 4; }
    const int row = street * 3 + (depth > 2 ? 2 : depth);
    for (int i = 0; i < kGridLen[row]; ++i) out[i] = (uint8_t)(E_RAISE0 + kGrid[row][i]);
    return kGridLen[row];
}
inline Chips f32_to_chips(float v) {  // Rust `as i16`: truncation toward zero, saturating, NaN → 0
    if (!(v == v)) return 0;
    if (v >= 32767.0f) return 32767;
    if (v <= -32768.0f) return -32768;
    return (Chips)v;
}
inline Chips into_chips(uint8_t e, Chips pot) {  // edge.rs:89-95
    if (e >= E_RAISE0) return f32_to_chips((float)pot * ((float)kRaises[e - E_RAISE0][0] / (float)kRaises[e - E_RAISE0][1]));
    if (e >= E_OPEN0) return (Chips)(kOpens[e - E_OPEN0] * kBB);
    return 0;
}
// edge.rs:185-197 `impl From<Edge> for u64` — the blueprint table's `edge` column (nlhe/src/profile.rs:143-160), NOT the 5-bit
// path code: Draw 0, Fold 1, Check 2, Call 3, Raise(n/d) 4 | n << 3 | d << 11, Shove 5, Open(n) 6 | n << 3
inline uint64_t e_to_u64(uint8_t e) {
    if (e >= E_RAISE0) return 4u | (uint64_t)kRaises[e - E_RAISE0][0] << 3 | (uint64_t)kRaises[e - E_RAISE0][1] << 11;
    if (e >= E_OPEN0) return 6u | (uint64_t)kOpens[e - E_OPEN0] << 3;
    return e == E_SHOVE ? 5u : (uint64_t)(e - 1);
}
// edge.rs:160-183 `impl From<u64> for Edge` incl. the legacy form (tag 4 with bit 19 = BBs, read as Open); 0 = not on this grid
inline uint8_t e_from_u64(uint64_t v) {
    const uint64_t lo = v >> 3 & 0xFF, hi = v >> 11 & 0xFF;
    auto open = [&](uint64_t n) -> uint8_t { for (int i = 0; i < 4; ++i) if ((uint64_t)kOpens[i] == n) return (uint8_t)(E_OPEN0 + i); return 0; };
    switch (v & 0b111) {
        case 0: return E_DRAW;
        case 1: return E_FOLD;
        case 2: return E_CHECK;
        case 3: return E_CALL;
        case 5: return E_SHOVE;
        case 6: return open(lo);
        case 4:
            if (v & (1ull << 19)) return open(lo);
            for (int i = 0; i < 10; ++i) if ((uint64_t)kRaises[i][0] == lo && (uint64_t)kRaises[i][1] == hi) return (uint8_t)(E_RAISE0 + i);
            return 0;
        default: return 0;
    }
}

// path.rs: 5 bits per edge, first edge in the low bits, at most 12 edges
inline uint64_t path_push(uint64_t p, uint8_t e) {
    int len = 0;
    for (uint64_t q = p; q & 0x1F; q >>= 5) ++len;
    return len >= kMaxPathEdges ? p : p | (uint64_t)e << (5 * len);
}
inline int path_aggression(uint64_t p) {  // path.rs:14-20: aggro edges among the trailing choice edges
    uint8_t es[kMaxPathEdges];
    int n = 0;
    for (uint64_t q = p; q & 0x1F; q >>= 5) es[n++] = (uint8_t)(q & 0x1F);
    int a = 0;
    for (int i = n - 1; i >= 0 && e_is_choice(es[i]); --i) a += e_is_aggro(es[i]);
    return a;
}

enum SeatState : uint8_t { BETTING = 0, SHOVING = 1, FOLDING = 2 };
struct Seat { uint8_t state; Chips stack, stake, spent; uint64_t cards; };
enum ActKind : uint8_t { A_DRAW, A_FOLD, A_CALL, A_CHECK, A_RAISE, A_SHOVE, A_BLIND };
struct Action { uint8_t kind; Chips chips; uint64_t cards; };

inline uint64_t mix64(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

// deck.rs:22-39 Deck::draw with index i = range(n): i in {0,1} → lowest card, i = k → k-th lowest
inline int deck_draw(uint64_t* deck, uint32_t word) {
    const int n = popc64(*deck);
    const uint32_t i = draw_range(word, (uint32_t)n);
    uint64_t d = *deck;
    int card = __builtin_ctzll(d);
    for (uint32_t ones = 0; ones < i; ++ones) { card = __builtin_ctzll(d); d &= d - 1; }
    *deck &= ~(1ull << card);
    return card;
}

struct Game {  // kicker GameN<2>
    Chips pot;
    uint64_t board;
    Seat seats[2];
    uint8_t dealer, ticker;

    int n_board() const { return popc64(board); }
    int street() const { const int b = n_board(); return b == 0 ? 0 : b - 2; }  // board.rs:26-28
    int actor_idx() const { return (dealer + ticker) % 2; }
    const Seat& actor() const { return seats[actor_idx()]; }
    Seat& actor_mut() { return seats[actor_idx()]; }
    Chips max_stake() const { return seats[0].stake > seats[1].stake ? seats[0].stake : seats[1].stake; }
    bool everyone_folding() const { return (seats[0].state != FOLDING) + (seats[1].state != FOLDING) == 1; }
    bool everyone_shoving() const {
        for (const Seat& s : seats) if (s.state != FOLDING && s.state != SHOVING) return false;
        return true;
    }
    bool everyone_touched() const { return ticker > 2 + (street() == 0 ? 1 : 0); }  // game.rs:489-492 (P == 2: offset 1)
    bool everyone_matched() const {
        const Chips st = max_stake();
        for (const Seat& s : seats) if (s.state == BETTING && s.stake != st) return false;
        return true;
    }
    bool everyone_calling() const { return everyone_touched() && everyone_matched(); }
    bool everyone_alright() const { return everyone_calling() || everyone_folding() || everyone_shoving(); }
    bool must_stop() const { return street() == 3 ? everyone_alright() : everyone_folding(); }
    bool must_deal() const { return street() != 3 && everyone_alright(); }
    bool must_post() const { return street() == 0 && pot < kSB + kBB; }
    enum { T_TERMINAL = 3, T_CHANCE = 2 };
    int turn() const { return must_stop() ? T_TERMINAL : (must_deal() ? T_CHANCE : actor_idx()); }
    bool is_choice() const { return turn() < 2; }
    Chips to_call() const { return (Chips)(max_stake() - actor().stake); }
    Chips to_shove() const { return actor().stack; }
    Chips to_post() const { const Chips b = pot < kSB ? kSB : kBB; return b < actor().stack ? b : actor().stack; }
    Chips to_raise() const {  // game.rs:556-576
        Chips most = 0, next = 0;
        for (const Seat& s : seats) {
            if (s.state == FOLDING) continue;
            if (s.stake > most) { next = most; most = s.stake; } else if (s.stake > next) next = s.stake;
        }
        const Chips relative = (Chips)(most - actor().stake), marginal = (Chips)(most - next);
        return (Chips)(relative + (marginal > kBB ? marginal : kBB));
    }
    bool may_fold() const { return is_choice() && to_call() > 0; }
    bool may_call() const { return is_choice() && may_fold() && to_call() < to_shove(); }
    bool may_check() const { return is_choice() && max_stake() == actor().stake; }
    bool may_raise() const { return is_choice() && to_raise() < to_shove(); }
    bool may_shove() const { return is_choice() && to_shove() > 0; }
    Action passive() const { return may_check() ? Action{A_CHECK, 0, 0} : Action{A_FOLD, 0, 0}; }

    void next_player() {  // game.rs:448-460
        if (!everyone_alright()) {
            for (;;) { ticker += 1; if (actor().state == BETTING) break; }
        }
    }
    void bet(Chips c) {
        pot = (Chips)(pot + c);
        Seat& s = actor_mut();
        s.stack = (Chips)(s.stack - c); s.stake = (Chips)(s.stake + c); s.spent = (Chips)(s.spent + c);
        if (s.stack == 0) s.state = SHOVING;
    }
    void force_act(const Action& a) {  // game.rs:395-415
        switch (a.kind) {
            case A_CHECK: next_player(); break;
            case A_FOLD: actor_mut().state = FOLDING; next_player(); break;
            case A_DRAW:
                ticker = 0; board |= a.cards;
                next_player();
                seats[0].stake = seats[1].stake = 0;
                break;
            default: bet(a.chips); next_player(); break;
        }
    }
    // game.rs:835-855 snap
    Action snap(Action a) const {
        switch (a.kind) {
            case A_RAISE:
                if (a.chips >= to_shove() || !may_raise()) return snap(Action{A_SHOVE, to_shove(), 0});
                if (a.chips < to_raise()) return Action{A_RAISE, to_raise(), 0};
                return a;
            case A_SHOVE:
                if (may_shove()) return Action{A_SHOVE, to_shove(), 0};
                if (may_call()) return Action{A_CALL, to_call(), 0};
                return passive();
            case A_CALL:
                if (may_call()) return Action{A_CALL, to_call(), 0};
                if (may_shove()) return Action{A_SHOVE, to_shove(), 0};
                return passive();
            case A_CHECK:
                if (may_check()) return a;
                if (may_call()) return Action{A_CALL, to_call(), 0};
                return Action{A_FOLD, 0, 0};
            case A_FOLD:
                if (may_fold()) return a;
                return Action{A_CHECK, 0, 0};
            default: return a;
        }
    }
    // game.rs:613-635 settlements: ledger of (spent, state, Strength(hole ∪ board)) → Showdown::settle; returns `won` per seat
    void settle(Chips won[2]) const;
};

// showdown.rs:36-110 Showdown::settle over an n-seat ledger (strength = packed Strength, 0xFFFFFFFF = Ranking::MAX)
inline void showdown(int n, const Chips* risked, const uint8_t* status, const uint32_t* str, Chips* reward) {
    for (int i = 0; i < n; ++i) reward[i] = 0;
    uint32_t best = 0xFFFFFFFFu;
    Chips distributing = 0, distributed = 0;
    for (;;) {  // 'winners: strongest() = max strength below `best` among non-folded
        bool any = false;
        uint32_t top = 0;
        for (int i = 0; i < n; ++i)
            if (str[i] < best && status[i] != FOLDING && (!any || str[i] > top)) { top = str[i]; any = true; }
        if (!any) return;
        best = top;
        for (;;) {  // 'pots: remaining() = smallest stake above what is already distributed, among this tier
            distributed = distributing;
            bool have = false;
            Chips amount = 0;
            for (int i = 0; i < n; ++i)
                if (str[i] == best && risked[i] > distributed && status[i] != FOLDING && (!have || risked[i] < amount)) { amount = risked[i]; have = true; }
            if (!have) break;
            distributing = amount;
            Chips chips = 0;  // winnings()
            for (int i = 0; i < n; ++i) {
                Chips s = risked[i] < distributing ? risked[i] : distributing;
                s = (Chips)(s - distributed);
                chips = (Chips)(chips + (s > 0 ? s : 0));
            }
            int nw = 0;  // distribute(): share to every eligible winner, the remainder one chip each in seat order
            for (int i = 0; i < n; ++i) nw += status[i] != FOLDING && str[i] == best && risked[i] > distributed;
            const Chips share = (Chips)(chips / nw);
            Chips bonus = (Chips)(chips % nw);
            for (int i = 0; i < n; ++i)
                if (status[i] != FOLDING && str[i] == best && risked[i] > distributed) {
                    reward[i] = (Chips)(reward[i] + share);
                    if (bonus > 0) { reward[i] = (Chips)(reward[i] + 1); --bonus; }
                }
            int staked = 0, paid = 0;  // is_complete()
            for (int i = 0; i < n; ++i) { staked += risked[i]; paid += reward[i]; }
            if (staked == paid) return;
        }
    }
}
inline void Game::settle(Chips won[2]) const {
    uint32_t str[2];
    Chips risked[2], reward[2];
    uint8_t status[2];
    for (int i = 0; i < 2; ++i) { str[i] = strength(seats[i].cards | board); risked[i] = seats[i].spent; status[i] = seats[i].state; }
    showdown(2, risked, status, str, reward);
    for (int i = 0; i < 2; ++i) won[i] = (Chips)(reward[i] - risked[i]);
}

// tree-level state: the game plus what `NlheInfo` reads off the tree path
struct State {
    Game game;
    uint64_t subgame;  // current-street choice edges, chronological (info.rs:94-108)
    uint64_t hist;     // running hash of every applied edge (RNG key of board draws)
};
struct Info {
    uint64_t subgame, choices;
    uint16_t abs;
    bool operator==(const Info& o) const { return subgame == o.subgame && choices == o.choices && abs == o.abs; }
};
struct InfoHash { size_t operator()(const Info& i) const { return (size_t)mix64(i.subgame ^ mix64(i.choices ^ mix64(i.abs))); } };
inline uint32_t info_word(const Info& i) { return (uint32_t)mix64(i.subgame ^ mix64(i.choices ^ mix64(i.abs))); }

struct Ctx {  // per-draw RNG context
    Draw rng;
    uint32_t epoch, tree;
};

// synthetic abstraction lookup (see header)
inline uint16_t abstraction_of(const Game& g) {
    const Obs c = canonical(Obs{g.actor().cards, g.board});
    static const int K[4] = {169, 256, 256, 101};
    const int street = g.street();
    return (uint16_t)(street << 8 | (int)(mix64(c.pocket * 0x9E3779B97F4A7C15ull ^ mix64(c.pub)) % (uint64_t)K[street]));
}
// game.rs:724-739 choices(depth): legal actions unfolded onto the grid, in `legal()` order: raises, shove, call, fold, check
inline int choices_of(const Game& g, int depth, uint8_t* out) {
    if (g.must_stop()) return 0;
    if (g.must_deal()) { out[0] = E_DRAW; return 1; }
    int n = 0;
    if (g.may_raise()) n += raises(g.street(), depth, out + n);
    if (g.may_shove()) out[n++] = E_SHOVE;
    if (g.may_call()) out[n++] = E_CALL;
    if (g.may_fold()) out[n++] = E_FOLD;
    if (g.may_check()) out[n++] = E_CHECK;
    return n;
}
inline uint64_t pack_path(const uint8_t* e, int n) { uint64_t p = 0; for (int i = 0; i < n && i < kMaxPathEdges; ++i) p |= (uint64_t)e[i] << (5 * i); return p; }
inline Info info_of(const State& s) {
    uint8_t ch[kMaxEdges];
    const int n = choices_of(s.game, path_aggression(s.subgame), ch);
    // the bucket is only read at decision nodes (chance nodes have the single Draw branch, terminals none): skipped elsewhere
    return Info{s.subgame, pack_path(ch, n), s.game.is_choice() ? abstraction_of(s.game) : (uint16_t)0xFFFF};
}
inline int info_choices(const Info& i, uint8_t* out) { int n = 0; for (uint64_t q = i.choices; q & 0x1F; q >>= 5) out[n++] = (uint8_t)(q & 0x1F); return n; }

// game.rs:605-607 reveal: Action::Draw(deck.deal(street)) with the contract's draws
inline Action reveal(const Game& g, const Ctx& cx, uint64_t hist) {
    uint64_t deck = ~(g.board | g.seats[0].cards | g.seats[1].cards) & 0x000FFFFFFFFFFFFFull;
    const int n = g.street() == 0 ? 3 : 1;  // street.next().n_revealed()
    const Philox4 w = cx.rng.at(cx.epoch, cx.tree, (uint32_t)hist, TAG_DRAW);
    uint64_t cards = 0;
    for (int k = 0; k < n; ++k) cards |= 1ull << deck_draw(&deck, w.r[k]);
    return Action{A_DRAW, 0, cards};
}
inline Action actionize(const Game& g, uint8_t e, const Ctx& cx, uint64_t hist) {  // game.rs:741-752
    switch (e) {
        case E_FOLD: return Action{A_FOLD, 0, 0};
        case E_DRAW: return reveal(g, cx, hist);
        case E_CALL: return Action{A_CALL, g.to_call(), 0};
        case E_CHECK: return Action{A_CHECK, 0, 0};
        case E_SHOVE: return Action{A_SHOVE, g.to_shove(), 0};
        default: return Action{A_RAISE, into_chips(e, g.pot), 0};
    }
}
// nlhe/src/game.rs:35-55 NlheGame::apply
inline State apply(const State& s, uint8_t edge, const Ctx& cx) {
    State out = s;
    Game& game = out.game;
    if (game.turn() == Game::T_TERMINAL) return out;
    if (e_is_choice(edge)) {
        while (game.turn() == Game::T_CHANCE) {
            out.hist = mix64(out.hist ^ E_DRAW);
            game.force_act(reveal(game, cx, out.hist));
        }
        if (game.turn() == Game::T_TERMINAL) return out;
    }
    if (!e_is_choice(edge) && game.turn() != Game::T_CHANCE) return out;
    out.hist = mix64(out.hist ^ edge);
    const Action a = game.snap(actionize(game, edge, cx, out.hist));
    game.force_act(a);
    out.subgame = e_is_choice(edge) ? path_push(s.subgame, edge) : 0;  // info.rs:97-103: trailing choice edges incl. the incoming one
    return out;
}
// kicker game.rs:59-78 root(): fresh deck, two holes, both blinds posted
inline State root(const Ctx& cx) {
    Game g{};
    g.pot = 0; g.board = 0; g.dealer = 0; g.ticker = 0;  // ticker = usize::from(P != 2)
    uint64_t deck = 0x000FFFFFFFFFFFFFull;
    const Philox4 w = cx.rng.at(cx.epoch, cx.tree, 0xFFFFFFFFu, TAG_ROOT);
    for (int i = 0; i < 2; ++i) {
        const int a = deck_draw(&deck, w.r[2 * i]), b = deck_draw(&deck, w.r[2 * i + 1]);
        g.seats[i] = Seat{BETTING, kStack, 0, 0, 1ull << a | 1ull << b};
    }
    for (int k = 0; k < 2; ++k) g.force_act(Action{A_BLIND, g.to_post(), 0});
    return State{g, 0, 0};
}
inline float payoff(const State& s, int player) {  // nlhe/src/game.rs:57-63
    Chips won[2];
    s.game.settle(won);
    return (float)won[player];
}

// ── MCCFR over this game: same arithmetic as oracle/mccfr.hpp (flow.rs / solver.rs), sparse profile ──
struct Row { Encounter e[kMaxEdges]; bool present[kMaxEdges]; uint8_t edges[kMaxEdges]; int n; };
struct View { int n; uint8_t edges[kMaxEdges]; float r[kMaxEdges], rd, sw[kMaxEdges], z; };
struct Dec { Info info; int n; int tree; bool explored[kMaxEdges]; float regret[kMaxEdges], policy[kMaxEdges], payoff; };  // POD: also the unit ranks exchange

struct Solver {
    std::unordered_map<Info, Row, InfoHash> rows;
    uint64_t epochs = 0;
    Hyper hyper;
    int regret_sched = R_LINEAR, weight_sched = W_LINEAR, sampling = S_PLURIBUS, batch = 128, threads = 1;
    std::unique_ptr<Pool> pool;  // rayon's persistent pool (created on first use)
    int world_rank = 0, world_size = 1;  // this handle samples tree ids [rank*batch, (rank+1)*batch) of a world_size*batch epoch
    Draw rng{0};
    uint64_t nodes = 0, infos = 0, updates = 0;
    std::vector<Info> touched;  // infosets the last fold wrote (owner-sharded exchange: the rows a rank broadcasts)

    struct TreeN {
        int id;
        std::vector<State> game; std::vector<Info> info; std::vector<int> parent, head, next; std::vector<uint8_t> incoming;
        int add(const State& s, const Info& i, int par, uint8_t e) {
            int k = (int)game.size();
            game.push_back(s); info.push_back(i); parent.push_back(par); incoming.push_back(e); head.push_back(-1); next.push_back(-1);
            if (par >= 0) { next[k] = head[par]; head[par] = k; }
            return k;
        }
    };
    int walker() const { return (int)(epochs % 2); }
    float cum_regret(const Info& i, int a, uint8_t e) const {
        auto it = rows.find(i);
        return (it != rows.end() && it->second.present[a]) ? it->second.e[a].regret : e_default_regret(e);
    }
    float cum_weight(const Info& i, int a) const {
        auto it = rows.find(i);
        return (it != rows.end() && it->second.present[a]) ? it->second.e[a].weight : 0.0f;
    }
    Encounter& mut_row(const Info& i, int a) {
        auto it = rows.find(i);
        if (it == rows.end()) { Row r{}; r.n = info_choices(i, r.edges); it = rows.emplace(i, r).first; }
        Row& r = it->second;
        if (!r.present[a]) { r.present[a] = true; r.e[a] = Encounter{0.0f, e_default_regret(r.edges[a]), 0.0f, 0}; }
        return r.e[a];
    }
    View view(const Info& i) const {
        View v;
        v.n = info_choices(i, v.edges);
        float rd = 0.0f, ws = 0.0f, w[kMaxEdges];
        for (int a = 0; a < v.n; ++a) {
            const float cr = cum_regret(i, a, v.edges[a]);
            v.r[a] = cr > EPS ? cr : EPS; rd = rd + v.r[a];
            const float cw = cum_weight(i, a);
            w[a] = cw > EPS ? cw : EPS; ws = ws + w[a];
        }
        v.rd = rd;
        const float denom = ws + hyper.smoothing;
        float z = 0.0f;
        for (int a = 0; a < v.n; ++a) { const float s = (w[a] / hyper.temperature + hyper.smoothing) / denom; v.sw[a] = s > hyper.curiosity ? s : hyper.curiosity; z = z + v.sw[a]; }
        v.z = z;
        return v;
    }
    static int act_of(const View& v, uint8_t e) { for (int a = 0; a < v.n; ++a) if (v.edges[a] == e) return a; return -1; }

    struct Leaf { uint8_t edge; State game; int head; };
    void branches(const TreeN& t, int node, const Ctx& cx, std::vector<Leaf>& out) const {
        uint8_t ed[kMaxEdges];
        const int n = info_choices(t.info[node], ed);  // node.branches(): info.choices() → apply
        for (int k = 0; k < n; ++k) out.push_back(Leaf{ed[k], apply(t.game[node], ed[k], cx), node});
    }
    void sample(const TreeN& t, int node, std::vector<Leaf>& br) const {
        if (br.empty()) return;
        const int turn = t.game[node].game.turn();
        const Info& info = t.info[node];
        if (turn == walker()) {
            if (sampling == S_EXTERNAL) return;
            if (sampling == S_PLURIBUS) {
                if (epochs < hyper.prune_warmup) return;
                const Philox4 c = rng.at((uint32_t)epochs, (uint32_t)t.id, info_word(info), TAG_COIN);
                if (draw_unit(c.r[0]) < hyper.prune_explore) return;
            }
            const View v = view(info);
            std::vector<Leaf> kept;
            for (const Leaf& l : br) {
                const int a = act_of(v, l.edge);
                bool keep = cum_regret(info, a, l.edge) > hyper.prune_threshold;
                if (sampling == S_PLURIBUS && l.game.game.turn() == Game::T_TERMINAL) keep = true;
                if (keep) kept.push_back(l);
            }
            if (!kept.empty()) br.swap(kept);
            return;
        }
        const Philox4 c = rng.at((uint32_t)epochs, (uint32_t)t.id, info_word(info), TAG_NODE);
        int pick;
        if (turn == Game::T_CHANCE) pick = (int)draw_range(c.r[0], (uint32_t)br.size());
        else {
            const View v = view(info);
            float w[kMaxEdges];
            for (size_t k = 0; k < br.size(); ++k) { const float q = v.sw[act_of(v, br[k].edge)] / v.z; w[k] = q > EPS ? q : EPS; }
            pick = draw_weighted(c.r[0], w, (int)br.size());
        }
        const Leaf chosen = br[pick];
        br.clear();
        br.push_back(chosen);
    }
    TreeN build(int id) const {
        TreeN t;
        t.id = id;
        const Ctx cx{rng, (uint32_t)epochs, (uint32_t)id};
        const State r = root(cx);
        t.add(r, info_of(r), -1, 0);
        std::vector<Leaf> todo;
        branches(t, 0, cx, todo);
        sample(t, 0, todo);
        while (!todo.empty()) {
            const Leaf leaf = todo.back();
            todo.pop_back();
            const int node = t.add(leaf.game, info_of(leaf.game), leaf.head, leaf.edge);
            std::vector<Leaf> kids;
            branches(t, node, cx, kids);
            sample(t, node, kids);
            for (const Leaf& k : kids) todo.push_back(k);
        }
        return t;
    }
    float ancestor_reach(const TreeN& t, int root_) const {
        float cf = 1.0f, sm = 1.0f;
        for (int node = root_; t.parent[node] >= 0; node = t.parent[node]) {
            const int par = t.parent[node], turn = t.game[par].game.turn();
            if (turn < 2 && turn != walker()) {
                const View v = view(t.info[par]);
                const int a = act_of(v, t.incoming[node]);
                cf = cf * (v.r[a] / v.rd);
                sm = sm * (v.sw[a] / v.z);
            }
        }
        return cf / sm;
    }
    float recursed(const TreeN& t, int hero, int node, float rel, float smp) const {
        if (t.head[node] < 0) return rel / smp * payoff(t.game[node], hero);
        const int turn = t.game[node].game.turn();
        const bool chance = turn == Game::T_CHANCE, walk = turn == walker();
        View v{};
        if (!chance) v = view(t.info[node]);
        float sum = 0.0f;
        for (int c = t.head[node]; c >= 0; c = t.next[c]) {
            float r2 = rel, s2 = smp;
            if (!chance) {
                const int a = act_of(v, t.incoming[c]);
                r2 = rel * (v.r[a] / v.rd);
                if (!walk) s2 = smp * (v.sw[a] / v.z);
            }
            sum = sum + recursed(t, hero, c, r2, s2);
        }
        return sum;
    }
    void tree_decisions(const TreeN& t, std::vector<Dec>& out) const {
        std::unordered_map<Info, std::vector<int>, InfoHash> spans;
        std::vector<Info> order;
        for (int n = 0; n < (int)t.game.size(); ++n) {
            if (t.head[n] < 0) continue;
            auto it = spans.find(t.info[n]);
            if (it == spans.end()) { order.push_back(t.info[n]); spans[t.info[n]].push_back(n); } else it->second.push_back(n);
        }
        for (const Info& key : order) {
            const std::vector<int>& span = spans[key];
            if (t.game[span[0]].game.turn() != walker()) continue;
            Dec d{};
            d.info = key;
            const View v = view(key);
            d.n = v.n;
            for (int a = 0; a < v.n; ++a) d.policy[a] = v.r[a] / v.rd;
            float pay = 0.0f;
            for (int root_ : span) {
                const float reach = ancestor_reach(t, root_);
                float val[kMaxEdges]; int act[kMaxEdges]; int k = 0;
                for (int c = t.head[root_]; c >= 0; c = t.next[c]) { act[k] = act_of(v, t.incoming[c]); val[k] = reach * recursed(t, walker(), c, 1.0f, 1.0f); ++k; }
                float ev = 0.0f;
                for (int i = 0; i < k; ++i) ev = ev + v.r[act[i]] / v.rd * val[i];
                pay += ev;
                for (int i = 0; i < k; ++i) { const int a = act[i]; if (!d.explored[a]) { d.explored[a] = true; d.regret[a] = 0.0f; } d.regret[a] += val[i] - ev; }
            }
            d.payoff = pay;
            d.tree = t.id;
            out.push_back(d);
        }
    }
    void apply_dec(const Dec& d) {
        for (int a = 0; a < d.n; ++a) {
            if (!d.explored[a]) continue;
            Encounter& e = mut_row(d.info, a);
            e.regret = regret_gain(regret_sched, e.regret, d.regret[a], epochs, hyper);
            ++updates;
        }
        for (int a = 0; a < d.n; ++a) { Encounter& e = mut_row(d.info, a); e.weight = weight_learn(weight_sched, e.weight, d.policy[a], epochs); }
        for (int a = 0; a < d.n; ++a) { Encounter& e = mut_row(d.info, a); e.payoff += (d.payoff - e.payoff) / (float)(e.visits + 1); }
        for (int a = 0; a < d.n; ++a) mut_row(d.info, a).visits += 1;
    }
    // solver.rs:225-240 batch: this rank's trees → Decisions (tree order, first-seen infoset order inside a tree)
    std::vector<Dec> sample_decs() {
        const int T = threads < 1 ? 1 : (threads > batch ? batch : threads);
        if (!pool || pool->size() != T) pool = std::make_unique<Pool>(T);
        constexpr int kChunk = 16;  // trees per claimed chunk
        const int chunks = (batch + kChunk - 1) / kChunk;
        std::vector<std::vector<Dec>> parts(chunks);
        std::vector<uint64_t> nc(chunks, 0);
        pool->run(chunks, [&](int c, int) {
            for (int i = c * kChunk; i < std::min(batch, (c + 1) * kChunk); ++i) {
                const TreeN t = build(world_rank * batch + i);
                nc[c] += t.game.size();
                tree_decisions(t, parts[c]);
            }
        });
        std::vector<Dec> all;
        for (int c = 0; c < chunks; ++c) { nodes += nc[c]; all.insert(all.end(), parts[c].begin(), parts[c].end()); }
        return all;
    }
    // solver.rs:96-105: every Decisions of the epoch (all ranks'), applied in global tree order
    void fold_decs(std::vector<Dec>& decs) {
        std::stable_sort(decs.begin(), decs.end(), [](const Dec& a, const Dec& b) { return a.tree < b.tree; });
        infos += decs.size();
        touched.clear();
        std::unordered_map<Info, bool, InfoHash> seen;
        for (const Dec& d : decs) {
            apply_dec(d);
            if (seen.emplace(d.info, true).second) touched.push_back(d.info);
        }
        epochs += 1;
    }
    void step() {
        std::vector<Dec> decs = sample_decs();
        fold_decs(decs);
    }
};

}  // namespace nlhe
}  // namespace orc
