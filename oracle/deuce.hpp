// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).  CPU restatement of the reference's hand evaluator
// and river equity (crates/deuce/src).  Integer work: parity bar is bit-exact.  Pinned by the 19 known-answer
// cases of crates/deuce/src/evaluator.rs:186-357 (tests/golden/eval_known_answers.json) and the opponent count
// C(45,2) = 990 (crates/deuce/src/observation.rs:294-299).
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

// crates/deuce/src/ranking.rs:33-44 (default, non-shortdeck build): derived Ord = declaration order.
// NOTE FullHouse < Flush in this build — reproduced as written.
enum RankingTag : uint32_t { HighCard = 0, OnePair = 1, TwoPair = 2, ThreeOAK = 3, Straight = 4, FullHouse = 5, Flush = 6, FourOAK = 7, StraightFlush = 8 };

// Strength = (Ranking, Kickers) compared lexicographically (strength.rs:6-10): packed so that integer order = Ord
//   bits 24-27 tag | 20-23 first rank | 16-19 second rank | 0-12 kicker rank bits (kicks.rs:4)
inline uint32_t pack_strength(uint32_t tag, uint32_t r1, uint32_t r2, uint32_t kick) { return tag << 24 | r1 << 20 | r2 << 16 | kick; }

// hand.rs:66-86 Hand::ranks — 13-bit rank presence
inline uint16_t hand_ranks(uint64_t h) {
    uint16_t y = 0;
    for (int r = 0; r < 13; ++r)
        if ((h >> (4 * r)) & 0xF) y |= (uint16_t)(1u << r);
    return y;
}
inline int msb16(uint16_t v) { int m = -1; for (int i = 0; i < 16; ++i) if (v >> i & 1) m = i; return m; }  // rank.rs:75-80
inline int popc64(uint64_t v) { int c = 0; while (v) { v &= v - 1; ++c; } return c; }

// evaluator.rs:151-176 find_rank_of_n_oak_skip: highest rank (≠ skip) holding at least n cards
inline int rank_of_n_oak(uint64_t h, int n, int skip) {
    for (int r = 12; r >= 0; --r) {
        if (r == skip) continue;
        if (popc64((h >> (4 * r)) & 0xF) >= n) return r;
    }
    return -1;
}
// evaluator.rs:114-130 find_rank_of_straight (WHEEL = 0b1000000001111 → Five)
inline int rank_of_straight(uint16_t ranks) {
    uint16_t bits = ranks;
    bits &= bits << 1; bits &= bits << 1; bits &= bits << 1; bits &= bits << 1;
    if (bits) return msb16(bits);
    if ((ranks & 0x100F) == 0x100F) return 3;
    return -1;
}
// evaluator.rs:138-148 find_suit_of_flush: first suit (C,D,H,S) with >= 5 cards
inline int suit_of_flush(uint64_t h) {
    for (int s = 0; s < 4; ++s)
        if (popc64(h & (0x0001111111111111ull << s)) >= 5) return s;
    return -1;
}
// evaluator.rs:51-68 find_kickers: top n ranks outside the ranking's own ranks (ranking.rs:47-66)
inline uint16_t kickers_of(uint16_t ranks, uint16_t exclude, int n) {
    uint16_t k = ranks & (uint16_t)~exclude;
    int c = 0;
    for (int i = 0; i < 16; ++i) c += k >> i & 1;
    while (c > n) { k &= (uint16_t)(k - 1); --c; }
    return k;
}

// strength.rs:19-31 + evaluator.rs:39-50 find_ranking order
inline uint32_t strength(uint64_t hand) {
    hand &= 0x000FFFFFFFFFFFFFull;
    const uint16_t ranks = hand_ranks(hand);
    int s = suit_of_flush(hand);
    if (s >= 0) {  // find_straight_flush
        int r = rank_of_straight(hand_ranks(hand & (0x0001111111111111ull << s)));
        if (r >= 0) return pack_strength(StraightFlush, r, 0, 0);
    }
    int q = rank_of_n_oak(hand, 4, -1);
    if (q >= 0) return pack_strength(FourOAK, q, 0, kickers_of(ranks, 1u << q, 1));
    int t = rank_of_n_oak(hand, 3, -1);
    if (t >= 0) {
        int p = rank_of_n_oak(hand, 2, t);
        if (p >= 0) return pack_strength(FullHouse, t, p, 0);
    }
    if (s >= 0) return pack_strength(Flush, msb16(hand_ranks(hand & (0x0001111111111111ull << s))), 0, 0);
    int st = rank_of_straight(ranks);
    if (st >= 0) return pack_strength(Straight, st, 0, 0);
    if (t >= 0) return pack_strength(ThreeOAK, t, 0, kickers_of(ranks, 1u << t, 2));
    int hi = rank_of_n_oak(hand, 2, -1);
    if (hi >= 0) {
        int lo = rank_of_n_oak(hand, 2, hi);
        if (lo >= 0) return pack_strength(TwoPair, hi, lo, kickers_of(ranks, (1u << hi) | (1u << lo), 1));
        return pack_strength(OnePair, hi, 0, kickers_of(ranks, 1u << hi, 3));
    }
    int h1 = rank_of_n_oak(hand, 1, -1);
    return pack_strength(HighCard, h1, 0, kickers_of(ranks, 1u << h1, 4));
}

// observation.rs:45-62 equity: wins / (wins + losses) over every villain hole from the 45 unseen cards;
// ties dropped, 0.5 when nothing is decisive.  Also returns the counts.
inline float river_equity(uint64_t pocket, uint64_t pub, uint32_t* wins_out = nullptr, uint32_t* total_out = nullptr) {
    const uint64_t seen = pocket | pub;
    const uint32_t hero = strength(seen);
    uint32_t won = 0, sum = 0;
    for (int i = 0; i < 52; ++i) {
        if (seen >> i & 1) continue;
        for (int j = i + 1; j < 52; ++j) {
            if (seen >> j & 1) continue;
            uint32_t v = strength(pub | 1ull << i | 1ull << j);
            if (hero > v) { ++won; ++sum; } else if (hero < v) { ++sum; }
        }
    }
    if (wins_out) *wins_out = won;
    if (total_out) *total_out = sum;
    return sum == 0 ? 0.5f : (float)won / (float)sum;
}
// kicker/src/abstraction.rs:43-45 quantize: round(p * 100), half away from zero (f32::round)
inline uint8_t equity_bucket(float p) {
    return (uint8_t)roundf(p * 100.0f);
}

}  // namespace orc
