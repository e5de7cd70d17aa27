// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).
//
// Rule-based restatements of the reference's two validation games.  Each function cites the
// reference lines it follows.  The packed `info_key` layouts are part of the RNG contract
// (the key is a Philox counter word) and are documented in include/rbp.h.
#pragma once
#include <cstdint>

#include "rng.hpp"

namespace orc {

enum Turn : uint8_t { TURN_P0 = 0, TURN_P1 = 1, TURN_CHANCE = 2, TURN_TERMINAL = 3 };
constexpr int MAX_BRANCH = 8;  // max(choices, deals) for the small games

// ───────────────────────────── Kuhn (crates/kuhn/src) ─────────────────────────────
struct KuhnGame {
    // crates/kuhn/src/game.rs:8-30
    enum Node : uint8_t { Start, Dealt, Open, Check, Bet, CheckBet, OverFold0, OverFold1, OverShow1, OverShow2 };
    enum Edge : uint8_t { ECheck = 0, EBet = 1, ECall = 2, EFold = 3, EDeal = 8 };  // crates/kuhn/src/edge.rs:6-12
    struct State {
        uint8_t hole[2];
        uint8_t node;
    };
    static const char* name() { return "kuhn"; }
    static int rank(uint8_t card) { return card >> 1; }  // crates/kuhn/src/card.rs Card::ALL order J♠ J♥ Q♠ Q♥ K♠ K♥
    static bool board_is(const State&, uint8_t) { return false; }  // no board in Kuhn (WorldRestrict, kuhn/src/encoder.rs:47-66)

    // crates/kuhn/src/game.rs:115-123 (Fisher-Yates on ALL with two range draws)
    static State root(const Philox4& p) {
        uint8_t cards[6] = {0, 1, 2, 3, 4, 5};
        uint32_t i = draw_range(p.r[0], 6);
        uint8_t t = cards[0]; cards[0] = cards[i]; cards[i] = t;
        uint32_t j = 1 + draw_range(p.r[1], 5);
        t = cards[1]; cards[1] = cards[j]; cards[j] = t;
        return State{{cards[0], cards[1]}, Open};
    }
    static State exploitability_root() { return State{{0, 0}, Start}; }  // game.rs:161-166
    // game.rs:125-132
    static Turn turn(const State& s) {
        switch (s.node) {
            case Start: case Dealt: return TURN_CHANCE;
            case Open: case CheckBet: return TURN_P0;
            case Check: case Bet: return TURN_P1;
            default: return TURN_TERMINAL;
        }
    }
    // game.rs:134-152
    static State apply(const State& s, uint8_t e) {
        State n = s;
        if (s.node == Start) { n.hole[0] = e - EDeal; n.node = Dealt; return n; }
        if (s.node == Dealt) { n.hole[1] = e - EDeal; n.node = Open; return n; }
        switch (s.node) {
            case Open: n.node = (e == ECheck) ? Check : Bet; break;
            case Check: n.node = (e == ECheck) ? OverShow1 : CheckBet; break;
            case Bet: n.node = (e == ECall) ? OverShow2 : OverFold1; break;
            case CheckBet: n.node = (e == ECall) ? OverShow2 : OverFold0; break;
            default: break;
        }
        return n;
    }
    // game.rs:32-64
    static float payoff(const State& s, int p) {
        if (s.node == OverFold0 || s.node == OverFold1) {
            int who = (s.node == OverFold0) ? 0 : 1;
            return who == p ? -1.0f : 1.0f;
        }
        float stake = (s.node == OverShow2) ? 2.0f : 1.0f;
        int r0 = rank(s.hole[0]), r1 = rank(s.hole[1]);
        if (r0 > r1) return p == 0 ? stake : -stake;
        if (r0 < r1) return p == 1 ? stake : -stake;
        return 0.0f;
    }
    // crates/kuhn/src/encoder.rs:20-33 (seed ≡ resume on every root this build uses), info.rs:9-76
    static uint32_t info_key(const State& s) {
        Turn t = turn(s);
        uint32_t acting = (t == TURN_P0 || t == TURN_P1) ? 1u : 0u;
        int actor = (t == TURN_P1) ? 1 : 0;
        uint32_t hist = 0;  // game.rs:78-86: Open/Check/Bet/CheckBet, everything else Open
        if (s.node == Check) hist = 1; else if (s.node == Bet) hist = 2; else if (s.node == CheckBet) hist = 3;
        return acting | (hist << 1) | ((uint32_t)rank(s.hole[actor]) << 3);
    }
    // info.rs:31-37
    static int choices(uint32_t key, uint8_t* out) {
        uint32_t hist = (key >> 1) & 3;
        if (hist == 0 || hist == 1) { out[0] = ECheck; out[1] = EBet; } else { out[0] = EFold; out[1] = ECall; }
        return 2;
    }
    // game.rs:88-95 (deals) + encoder.rs:35-46 (branches)
    static int branches(const State& s, uint8_t* out) {
        Turn t = turn(s);
        if (t == TURN_TERMINAL) return 0;
        if (t == TURN_CHANCE) {
            bool h0 = s.node != Start, h1 = !(s.node == Start || s.node == Dealt);
            int n = 0;
            for (uint8_t c = 0; c < 6; ++c) {
                if (h0 && c == s.hole[0]) continue;
                if (h1 && c == s.hole[1]) continue;
                out[n++] = EDeal + c;
            }
            return n;
        }
        return choices(info_key(s), out);
    }
    static float default_regret(uint8_t) { return 0.0f; }  // crates/mccfr/src/state/edge.rs:29-34
};

// ───────────────────────────── Leduc (crates/leduc/src) ─────────────────────────────
struct LeducGame {
    enum Spot : uint8_t { SOpen = 0, SChecked = 1, SRaised = 2, SCheckRaised = 3 };  // game.rs:6-11
    enum Kind : uint8_t { Start, Dealt, R1, Deal, R2, FoldR1, FoldR2, Showdown };     // game.rs:14-28
    enum Edge : uint8_t { EFold = 0, ECheck = 1, ECall = 2, ERaise = 3, EDeal = 8 };  // edge.rs:6-12
    struct State {
        uint8_t hole[2];
        uint8_t kind;
        uint8_t board;  // card index, valid for R2/FoldR2/Showdown
        uint8_t r1;     // Spot
        uint8_t r2;     // Spot (R2/Showdown)
        uint8_t who;    // folder (FoldR1/FoldR2)
    };
    static const char* name() { return "leduc"; }
    static int rank(uint8_t card) { return card >> 1; }
    static bool board_is(const State& s, uint8_t c) { return (s.kind == R2 || s.kind == FoldR2 || s.kind == Showdown) && s.board == c; }  // leduc/src/encoder.rs:60 board() != Some(c)
    static bool raised(uint8_t spot) { return spot == SRaised || spot == SCheckRaised; }       // game.rs:37-39
    static int actor(uint8_t spot) { return (spot == SOpen || spot == SCheckRaised) ? 0 : 1; }  // game.rs:41-46

    // game.rs:177-185
    static State root(const Philox4& p) {
        uint8_t cards[6] = {0, 1, 2, 3, 4, 5};
        uint32_t i = draw_range(p.r[0], 6);
        uint8_t t = cards[0]; cards[0] = cards[i]; cards[i] = t;
        uint32_t j = 1 + draw_range(p.r[1], 5);
        t = cards[1]; cards[1] = cards[j]; cards[j] = t;
        return State{{cards[0], cards[1]}, R1, 0, SOpen, 0, 0};
    }
    static State exploitability_root() { return State{{0, 0}, Start, 0, 0, 0, 0}; }  // game.rs:239-244
    // game.rs:187-193
    static Turn turn(const State& s) {
        switch (s.kind) {
            case Start: case Dealt: case Deal: return TURN_CHANCE;
            case R1: return actor(s.r1) == 0 ? TURN_P0 : TURN_P1;
            case R2: return actor(s.r2) == 0 ? TURN_P0 : TURN_P1;
            default: return TURN_TERMINAL;
        }
    }
    // game.rs:195-223
    static State apply(const State& s, uint8_t e) {
        State n = s;
        if (s.kind == Start) { n.hole[0] = e - EDeal; n.kind = Dealt; return n; }
        if (s.kind == Dealt) { n.hole[1] = e - EDeal; n.kind = R1; n.r1 = SOpen; return n; }
        if (s.kind == R1) {
            switch (s.r1) {
                case SOpen: n.r1 = (e == ECheck) ? SChecked : SRaised; break;
                case SChecked: if (e == ECheck) { n.kind = Deal; } else { n.r1 = SCheckRaised; } break;
                case SRaised: if (e == ECall) { n.kind = Deal; } else { n.kind = FoldR1; n.who = 1; } break;
                case SCheckRaised: if (e == ECall) { n.kind = Deal; } else { n.kind = FoldR1; n.who = 0; } break;
            }
            return n;
        }
        if (s.kind == Deal) { n.kind = R2; n.board = e - EDeal; n.r2 = SOpen; return n; }
        if (s.kind == R2) {
            switch (s.r2) {
                case SOpen: n.r2 = (e == ECheck) ? SChecked : SRaised; break;
                case SChecked: if (e == ECheck) { n.kind = Showdown; } else { n.r2 = SCheckRaised; } break;
                case SRaised: if (e == ECall) { n.kind = Showdown; } else { n.kind = FoldR2; n.who = 1; } break;
                case SCheckRaised: if (e == ECall) { n.kind = Showdown; } else { n.kind = FoldR2; n.who = 0; } break;
            }
            return n;
        }
        return n;
    }
    // game.rs:57-109 (pot + payoff)
    static float payoff(const State& s, int p) {
        int pot[2];
        if (s.kind == FoldR1) {
            pot[0] = pot[1] = 1; pot[1 - s.who] += 2;
        } else if (s.kind == FoldR2) {
            int base = raised(s.r1) ? 3 : 1;
            pot[0] = pot[1] = base; pot[1 - s.who] += 4;
        } else {
            int base = raised(s.r1) ? 3 : 1, extra = raised(s.r2) ? 4 : 0;
            pot[0] = pot[1] = base + extra;
        }
        if (s.kind == FoldR1 || s.kind == FoldR2) return s.who == p ? -(float)pot[p] : (float)pot[s.who];
        int br = rank(s.board), r0 = rank(s.hole[0]), r1 = rank(s.hole[1]);
        bool pair0 = r0 == br, pair1 = r1 == br;
        int winner;
        if (pair0 && !pair1) winner = 0;
        else if (!pair0 && pair1) winner = 1;
        else winner = r0 > r1 ? 0 : (r0 < r1 ? 1 : -1);
        if (winner < 0) return 0.0f;
        return winner == p ? (float)pot[1 - p] : -(float)pot[p];
    }
    // game.rs:113-123 spots(), :135-147 board(); encoder.rs:21-35 resume(); info.rs:12-24
    static uint32_t info_key(const State& s) {
        Turn t = turn(s);
        uint32_t acting = (t == TURN_P0 || t == TURN_P1) ? 1u : 0u;
        int act = (t == TURN_P1) ? 1 : 0;
        uint32_t r1 = SOpen, r2c = 0, boardc = 0;
        switch (s.kind) {
            case Start: case Dealt: r1 = SOpen; r2c = 0; break;
            case R1: r1 = s.r1; r2c = 0; break;
            case Deal: r1 = s.r1; r2c = 1 + SOpen; break;
            case R2: r1 = s.r1; r2c = 1 + s.r2; boardc = 1 + rank(s.board); break;
            case FoldR1: r1 = SOpen; r2c = 0; break;
            case FoldR2: r1 = s.r1; r2c = 0; boardc = 1 + rank(s.board); break;
            case Showdown: r1 = s.r1; r2c = 1 + s.r2; boardc = 1 + rank(s.board); break;
        }
        return acting | (boardc << 1) | (r1 << 3) | (r2c << 5) | ((uint32_t)rank(s.hole[act]) << 8);
    }
    // info.rs:30-36
    static int choices(uint32_t key, uint8_t* out) {
        uint32_t r1 = (key >> 3) & 3, r2c = (key >> 5) & 7;
        uint32_t spot = r2c ? r2c - 1 : r1;
        if (spot == SOpen || spot == SChecked) { out[0] = ECheck; out[1] = ERaise; } else { out[0] = EFold; out[1] = ECall; }
        return 2;
    }
    // game.rs:153-162 deals(); encoder.rs:37-48 branches()
    static int branches(const State& s, uint8_t* out) {
        Turn t = turn(s);
        if (t == TURN_TERMINAL) return 0;
        if (t == TURN_CHANCE) {
            bool h0 = s.kind != Start, h1 = !(s.kind == Start || s.kind == Dealt);
            int n = 0;
            for (uint8_t c = 0; c < 6; ++c) {
                if (h0 && c == s.hole[0]) continue;
                if (h1 && c == s.hole[1]) continue;
                out[n++] = EDeal + c;  // board() is None at every chance node, so no third filter applies
            }
            return n;
        }
        return choices(info_key(s), out);
    }
    static float default_regret(uint8_t) { return 0.0f; }
};

// ───────────────────────────── Rock-Paper-Scissors (crates/roshambo/src) ─────────────────────────────
struct RpsGame {
    enum Edge : uint8_t { ER = 0, EP = 1, ES = 2 };  // edge.rs:6-10
    struct State { uint8_t id; };                     // game.rs:6: RpsGame(u8), 0 root, 1-3 after P1, 4-12 terminal
    static const char* name() { return "rps"; }
    static State root(const Philox4&) { return State{0}; }            // game.rs:11-13 (no deal)
    static State exploitability_root() { return State{0}; }
    static Turn turn(const State& s) { return s.id == 0 ? TURN_P0 : (s.id <= 3 ? TURN_P1 : TURN_TERMINAL); }  // game.rs:14-21
    static State apply(const State& s, uint8_t e) { return State{(uint8_t)(s.id == 0 ? 1 + e : 4 + 3 * (s.id - 1) + e)}; }  // game.rs:23-39
    // game.rs:41-63: ASYMMETRIC_UTILITY = 2 (crates/pokerkit/src/lib.rs:198): scissors wins/losses count double
    static float payoff(const State& s, int p) {
        const float direction = p == 0 ? 0.0f + 1.0f : 0.0f - 1.0f;
        float v;
        switch (s.id) {
            case 7: v = 0.0f + 1.0f; break;   // P > R
            case 5: v = 0.0f - 1.0f; break;   // R < P
            case 6: v = 0.0f + 2.0f; break;   // R > S
            case 11: v = 0.0f + 2.0f; break;  // S > P
            case 10: v = 0.0f - 2.0f; break;  // S < R
            case 9: v = 0.0f - 2.0f; break;   // P < S
            default: v = 0.0f; break;
        }
        return direction * v;
    }
    // encoder.rs:15-25: the infoset is the turn itself
    static uint32_t info_key(const State& s) {
        Turn t = turn(s);
        return (t == TURN_TERMINAL ? 0u : 1u) | ((uint32_t)(t == TURN_TERMINAL ? 2 : (int)t) << 1);
    }
    static int choices(uint32_t key, uint8_t* out) {  // turn.rs:43-49
        if (!(key & 1u)) return 0;
        out[0] = ER; out[1] = EP; out[2] = ES;
        return 3;
    }
    static int branches(const State& s, uint8_t* out) { return turn(s) == TURN_TERMINAL ? 0 : choices(info_key(s), out); }
    static float default_regret(uint8_t) { return 0.0f; }
};

}  // namespace orc
