// flat_game.cpp — host-side enumeration of the reference's validation games into FlatGame tables.
//
// Both games are instances of one "single-raise betting round" automaton over a 6-card deck
// (3 ranks × 2 suits, card = 2*rank + suit, order J♠ J♥ Q♠ Q♥ K♠ K♥ as crates/{kuhn,leduc}/src/card.rs):
//   Kuhn  — one round, stakes 1/+1, high card wins           (crates/kuhn/src/game.rs:32-64,134-152)
//   Leduc — two rounds with a board card between them, ante 1, raise +2 / +4, pair-with-board beats
//           high card                                         (crates/leduc/src/game.rs:57-109,195-223)
// The betting automaton (spot × action → spot | round-over | fold) is data, not code.
#include "flat_game.hpp"

#include <algorithm>
#include <deque>

namespace rbp {
namespace {

enum Phase : uint8_t { START, DEALT, BET, BOARD, FOLDED, SHOWN };
enum Spot : uint8_t { OPEN = 0, CHECKED = 1, RAISED = 2, CHECKRAISED = 3 };

struct Pos {
    uint8_t hole[2];
    uint8_t phase;
    uint8_t round;    // 0 or 1
    uint8_t spot[2];  // per round
    int8_t board;     // card or -1
    uint8_t who;      // folder
};

// spot × action(0/1) → next spot, or 0x10 = round ends quietly, 0x11 = round ends raised+called, 0x2p = player p folds
constexpr uint8_t kNext[4][2] = {
    /* OPEN        */ {CHECKED, RAISED},
    /* CHECKED     */ {0x10, CHECKRAISED},
    /* RAISED      */ {0x21, 0x11},
    /* CHECKRAISED */ {0x20, 0x11},
};
inline int actor_of(uint8_t spot) { return (spot == OPEN || spot == CHECKRAISED) ? 0 : 1; }
inline bool was_raised(uint8_t spot) { return spot == RAISED || spot == CHECKRAISED; }

struct Rules {
    int rounds;  // 1 = Kuhn, 2 = Leduc, 0 = Rock-Paper-Scissors (crates/roshambo/src/game.rs: 13 states, id kept in hole[0])
    uint8_t turn(const Pos& p) const {
        if (rounds == 0) return p.hole[0] == 0 ? TURN_P0 : (p.hole[0] <= 3 ? TURN_P1 : TURN_TERMINAL);  // roshambo game.rs:14-21
        switch (p.phase) {
            case START: case DEALT: case BOARD: return TURN_CHANCE;
            case BET: return (uint8_t)actor_of(p.spot[p.round]);
            default: return TURN_TERMINAL;
        }
    }
    // children in `branches()` order: deals ascending by card, actions in `choices()` order
    int expand(const Pos& p, Pos* out) const {
        int n = 0;
        if (rounds == 0) {  // roshambo game.rs:23-39: R, P, S in `choices()` order (turn.rs:43-49)
            if (p.hole[0] > 3) return 0;
            for (int e = 0; e < 3; ++e) { Pos q = p; q.hole[0] = (uint8_t)(p.hole[0] == 0 ? 1 + e : 4 + 3 * (p.hole[0] - 1) + e); out[n++] = q; }
            return n;
        }
        if (p.phase == START || p.phase == DEALT || p.phase == BOARD) {
            for (int c = 0; c < 6; ++c) {
                if (p.phase != START && c == p.hole[0]) continue;
                if (p.phase == BOARD && c == p.hole[1]) continue;
                Pos q = p;
                if (p.phase == START) { q.hole[0] = (uint8_t)c; q.phase = DEALT; }
                else if (p.phase == DEALT) { q.hole[1] = (uint8_t)c; q.phase = BET; q.round = 0; q.spot[0] = OPEN; }
                else { q.board = (int8_t)c; q.phase = BET; q.round = 1; q.spot[1] = OPEN; }
                out[n++] = q;
            }
            return n;
        }
        if (p.phase != BET) return 0;
        for (int a = 0; a < 2; ++a) {
            Pos q = p;
            uint8_t nx = kNext[p.spot[p.round]][a];
            if (nx < 0x10) q.spot[p.round] = nx;
            else if (nx & 0x20) { q.phase = FOLDED; q.who = nx & 1; }
            else {
                if (nx == 0x11) q.spot[p.round] = p.spot[p.round];  // spot already records the raise
                if (p.round + 1 < rounds) q.phase = BOARD; else q.phase = SHOWN;
            }
            out[n++] = q;
        }
        return n;
    }
    float payoff(const Pos& p, int me) const {
        if (rounds == 0) {  // roshambo game.rs:41-63 with ASYMMETRIC_UTILITY = 2 (pokerkit lib.rs:198): scissors counts double
            static const float first_player[13] = {0, 0, 0, 0, 0, -1, 2, 1, 0, -2, -2, 2, 0};
            const float v = 0.0f + first_player[p.hole[0]];
            return (me == 0 ? 1.0f : -1.0f) * v;
        }
        int r0 = p.hole[0] >> 1, r1 = p.hole[1] >> 1;
        if (rounds == 1) {  // Kuhn
            if (p.phase == FOLDED) return p.who == me ? -1.0f : 1.0f;
            float stake = was_raised(p.spot[0]) ? 2.0f : 1.0f;
            if (r0 == r1) return 0.0f;
            int winner = r0 > r1 ? 0 : 1;
            return winner == me ? stake : -stake;
        }
        // Leduc: what the folder / each showdown player has put in
        int base = 1;
        if (p.phase == FOLDED) {
            if (p.round == 1 && was_raised(p.spot[0])) base = 3;
            return p.who == me ? -(float)base : (float)base;
        }
        if (was_raised(p.spot[0])) base = 3;
        if (was_raised(p.spot[1])) base += 4;
        int br = p.board >> 1;
        bool pair0 = r0 == br, pair1 = r1 == br;
        int winner = -1;
        if (pair0 != pair1) winner = pair0 ? 0 : 1;
        else if (r0 != r1) winner = r0 > r1 ? 0 : 1;
        if (winner < 0) return 0.0f;
        return winner == me ? (float)base : -(float)base;
    }
    // include/rbp.h key layouts
    uint32_t info_key(const Pos& p) const {
        if (rounds == 0) { const uint8_t t0 = turn(p); return (t0 == TURN_TERMINAL ? 0u : 1u) | ((uint32_t)(t0 == TURN_TERMINAL ? 2 : t0) << 1); }  // encoder.rs:15-25
        uint8_t t = turn(p);
        uint32_t acting = t <= TURN_P1 ? 1u : 0u;
        int actor = t == TURN_P1 ? 1 : 0;
        uint32_t rank = p.hole[actor] >> 1;
        if (rounds == 1) {
            uint32_t hist = p.phase == BET ? p.spot[0] : 0u;  // kuhn game.rs:78-86: non-decision nodes read Open
            return acting | (hist << 1) | (rank << 3);
        }
        uint32_t r1 = 0, r2c = 0, boardc = 0;  // leduc game.rs:113-123 spots(), :135-147 board()
        switch (p.phase) {
            case START: case DEALT: break;
            case BET:
                r1 = p.spot[0];
                if (p.round == 1) { r2c = 1 + p.spot[1]; boardc = 1 + (p.board >> 1); }
                break;
            case BOARD: r1 = p.spot[0]; r2c = 1 + OPEN; break;
            case FOLDED:
                if (p.round == 1) { r1 = p.spot[0]; boardc = 1 + (p.board >> 1); }
                break;
            case SHOWN: r1 = p.spot[0]; r2c = 1 + p.spot[1]; boardc = 1 + (p.board >> 1); break;
        }
        return acting | (boardc << 1) | (r1 << 3) | (r2c << 5) | (rank << 8);
    }
};

}  // namespace

bool build_flat_game(int game_id, FlatGame* g) {
    if (game_id < 0 || game_id > 2) return false;
    Rules rules{game_id == 0 ? 1 : (game_id == 1 ? 2 : 0)};
    *g = FlatGame();
    std::vector<Pos> pos;
    pos.push_back(Pos{{0, 0}, (uint8_t)(game_id == 2 ? BET : START), 0, {OPEN, OPEN}, -1, 0});
    g->parent.push_back(-1);
    std::vector<int> depth{0};
    for (size_t i = 0; i < pos.size(); ++i) {  // breadth-first: children contiguous, levels contiguous
        Pos kids[8];
        int n = rules.expand(pos[i], kids);
        FlatNode nd{};
        nd.first_child = n ? (int32_t)pos.size() : -1;
        nd.n_child = (uint8_t)n;
        nd.turn = rules.turn(pos[i]);
        nd.info_key = rules.info_key(pos[i]);
        nd.info = -1;
        nd.payoff0 = nd.turn == TURN_TERMINAL ? rules.payoff(pos[i], 0) : 0.0f;
        g->nodes.push_back(nd);
        g->payoff1.push_back(nd.turn == TURN_TERMINAL ? rules.payoff(pos[i], 1) : 0.0f);
        if (nd.turn == TURN_TERMINAL) g->n_terminals++;
        for (int k = 0; k < n; ++k) {
            pos.push_back(kids[k]);
            g->parent.push_back((int32_t)i);
            depth.push_back(depth[i] + 1);
        }
    }
    const int N = (int)g->nodes.size();
    for (int i = 0; i < N; ++i) {
        if (i == 0 || depth[i] != depth[i - 1]) g->level_start.push_back(i);
    }
    g->level_start.push_back(N);
    // dense infoset ids in first-seen (BFS) order; rows contiguous per infoset
    for (int i = 0; i < N; ++i) {
        FlatNode& nd = g->nodes[i];
        if (nd.turn > TURN_P1) continue;
        int id = -1;
        for (size_t k = 0; k < g->info_key.size(); ++k)
            if (g->info_key[k] == nd.info_key) { id = (int)k; break; }
        if (id < 0) {
            id = (int)g->info_key.size();
            g->info_key.push_back(nd.info_key);
            g->info_player.push_back(nd.turn);
            g->info_actions.push_back(nd.n_child);
            g->info_row.push_back(g->n_rows);
            g->n_rows += nd.n_child;
        }
        nd.info = (int16_t)id;
    }
    // reference node order of the Vanilla tree: pop the LAST pushed branch first (builder.rs:141-160)
    g->lifo_index.assign(N, -1);
    {
        std::vector<int> todo{0};
        int idx = 0;
        while (!todo.empty()) {
            int n = todo.back();
            todo.pop_back();
            g->lifo_index[n] = idx++;
            for (int k = 0; k < g->nodes[n].n_child; ++k) todo.push_back(g->nodes[n].first_child + k);
        }
    }
    // spans (tree.rs:88-97 partition pushes nodes in ascending node index)
    const int I = (int)g->info_key.size();
    std::vector<std::vector<int>> spans(I);
    for (int i = 0; i < N; ++i)
        if (g->nodes[i].info >= 0) spans[g->nodes[i].info].push_back(i);
    for (int x = 0; x < I; ++x) {
        std::sort(spans[x].begin(), spans[x].end(), [&](int a, int b) { return g->lifo_index[a] < g->lifo_index[b]; });
        g->span_start.push_back((int32_t)g->span_nodes.size());
        for (int n : spans[x]) g->span_nodes.push_back(n);
    }
    g->span_start.push_back((int32_t)g->span_nodes.size());
    // `root()`: hole pair after the two deal edges
    if (game_id == 2) {  // `RpsGame::root()` is deterministic
        g->deck = 0;
        g->root_table.assign(1, 0);
    } else {
    g->deck = 6;
    g->root_table.assign(36, -1);
    for (int c0 = 0; c0 < 6; ++c0) {
        int dealt = g->nodes[0].first_child + c0;
        for (int c1 = 0; c1 < 6; ++c1) {
            if (c1 == c0) continue;
            g->root_table[c0 * 6 + c1] = g->nodes[dealt].first_child + (c1 < c0 ? c1 : c1 - 1);
        }
    }
    }
    // sizing DP, bottom-up (BFS order ⇒ children have larger ids)
    std::vector<int> mn[2], mi[2], md(N, 0);
    for (int w = 0; w < 2; ++w) { mn[w].assign(N, 1); mi[w].assign(N, 0); }
    for (int i = N - 1; i >= 0; --i) {
        const FlatNode& nd = g->nodes[i];
        for (int w = 0; w < 2; ++w) {
            int sn = 0, si = 0, xn = 0, xi = 0;
            for (int k = 0; k < nd.n_child; ++k) {
                int c = nd.first_child + k;
                sn += mn[w][c]; si += mi[w][c];
                xn = std::max(xn, mn[w][c]); xi = std::max(xi, mi[w][c]);
            }
            if (nd.turn == w) { mn[w][i] = 1 + sn; mi[w][i] = 1 + si; }
            else { mn[w][i] = 1 + xn; mi[w][i] = xi; }
        }
        for (int k = 0; k < nd.n_child; ++k) md[i] = std::max(md[i], 1 + md[nd.first_child + k]);
    }
    for (int r : g->root_table) {
        if (r < 0) continue;
        for (int w = 0; w < 2; ++w) {
            g->max_tree_nodes = std::max(g->max_tree_nodes, mn[w][r]);
            g->max_tree_infos = std::max(g->max_tree_infos, mi[w][r]);
        }
        g->max_depth = std::max(g->max_depth, md[r]);
    }
    return true;
}

}  // namespace rbp
