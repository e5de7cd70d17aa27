// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).  CPU restatement of the reference's suit-isomorphism
// machinery and of the histogram projections that turn one street's lookup into the next layer's points.
//
// Follows (crates/deuce/src): permutation.rs:9-21,40-54 (canonical suit relabelling: suits sorted by pocket count,
// public count, min/max ranks, suit id), isomorphism.rs:9-15,40-44, hand_iter.rs:14-44,62-76 (Gosper enumeration
// with a mask), observation_iter.rs:13-93, isomorphism_iter.rs:7-21, observation.rs:35-40 (children),
// street.rs (cards per street); and crates/lloyd/src/lookup.rs:46-66 (projections), histogram.rs:168-176.
// Pinned by the reference's counts: 169 / 1,286,792 / 13,960,050 / 123,156,254 isomorphisms per street
// (street.rs:129-135, tests isomorphism_iter.rs:42-70) and C(52,2) = 1326 (hand_iter.rs:118-121).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "deuce.hpp"

namespace orc {

struct Obs {
    uint64_t pocket, pub;
};
inline int street_public(int street) { return street == 0 ? 0 : street + 2; }  // Pref 0, Flop 3, Turn 4, Rive 5 cards

struct SuitKey {  // permutation.rs:40-54 order(): lexicographic, Option<Rank> with None < Some
    int pocket_n, public_n, pmin, bmin, pmax, bmax, suit;
    bool operator<(const SuitKey& o) const {
        if (pocket_n != o.pocket_n) return pocket_n < o.pocket_n;
        if (public_n != o.public_n) return public_n < o.public_n;
        if (pmin != o.pmin) return pmin < o.pmin;
        if (bmin != o.bmin) return bmin < o.bmin;
        if (pmax != o.pmax) return pmax < o.pmax;
        if (bmax != o.bmax) return bmax < o.bmax;
        return suit < o.suit;
    }
};
inline SuitKey suit_key(const Obs& o, int s) {
    const uint64_t m = 0x0001111111111111ull << s;
    const uint64_t p = o.pocket & m, b = o.pub & m;
    auto lo = [](uint64_t h) { return h ? (int)(__builtin_ctzll(h) / 4) : -1; };   // hand.rs min_rank: None -> -1
    auto hi = [](uint64_t h) { return h ? (int)((63 - __builtin_clzll(h)) / 4) : -1; };
    return SuitKey{popc64(p), popc64(b), lo(p), lo(b), hi(p), hi(b), s};
}
// permutation.rs:9-21: perm[old suit] = position of that suit in the sorted order
inline void suit_permutation(const Obs& o, int perm[4]) {
    SuitKey k[4] = {suit_key(o, 0), suit_key(o, 1), suit_key(o, 2), suit_key(o, 3)};
    std::stable_sort(k, k + 4);
    for (int i = 0; i < 4; ++i) perm[k[i].suit] = i;
}
inline bool is_canonical(const Obs& o) {  // isomorphism.rs:40-44
    int p[4];
    suit_permutation(o, p);
    return p[0] == 0 && p[1] == 1 && p[2] == 2 && p[3] == 3;
}
inline uint64_t permute_hand(uint64_t h, const int perm[4]) {  // permutation.rs:28-33,55-66 image/shift
    uint64_t out = 0;
    for (int s = 0; s < 4; ++s) {
        const uint64_t cards = h & (0x0001111111111111ull << s);
        const int shift = perm[s] - s;
        out |= shift >= 0 ? cards << shift : cards >> -shift;
    }
    return out;
}
inline Obs canonical(const Obs& o) {  // isomorphism.rs:9-15
    int p[4];
    suit_permutation(o, p);
    return Obs{permute_hand(o.pocket, p), permute_hand(o.pub, p)};
}

// hand_iter.rs:22-33 Gosper's hack
inline uint64_t gosper(uint64_t x) {
    const uint64_t a = x | (x - 1), b = a + 1, c = ~a, d = c & b, e = d - 1;
    const int f = 1 + __builtin_ctzll(x);
    return b | (f >= 64 ? 0 : e >> f);
}
struct HandIter {  // hand_iter.rs
    uint64_t next, mask;
    HandIter(int n, uint64_t m) : next(n == 0 ? 0 : (1ull << n) - 1), mask(m) {
        while ((next & mask) && !exhausted()) next = gosper(next);
    }
    bool exhausted() const { return next == 0 || (64 - 52) > __builtin_clzll(next); }
    bool step(uint64_t* out) {
        if (exhausted()) return false;
        *out = next;
        do { next = gosper(next); } while (next & mask);
        return true;
    }
};

// isomorphism_iter.rs + observation_iter.rs: canonical observations of a street in enumeration order
// (pocket-major, Gosper order inside); `pockets` optionally restricts to a [lo, hi) slice of the 1326 pockets
inline void enumerate_isomorphisms(int street, std::vector<Obs>& out, int pocket_lo = 0, int pocket_hi = 1326) {
    const int nb = street_public(street);
    HandIter outer(2, 0);
    uint64_t pocket;
    int pi = 0;
    while (outer.step(&pocket)) {
        if (pi >= pocket_lo && pi < pocket_hi) {
            if (nb == 0) {
                Obs o{pocket, 0};
                if (is_canonical(o)) out.push_back(o);
            } else {
                HandIter inner(nb, pocket);
                uint64_t pub;
                while (inner.step(&pub)) {
                    Obs o{pocket, pub};
                    if (is_canonical(o)) out.push_back(o);
                }
            }
        }
        ++pi;
    }
}

// observation.rs:35-40 children(): one more public card (flop→turn, turn→river), in HandIterator order
inline int children(const Obs& o, Obs* out) {
    HandIter it(1, o.pocket | o.pub);
    uint64_t c;
    int n = 0;
    while (it.step(&c)) out[n++] = Obs{o.pocket, o.pub | c};
    return n;
}

// lookup.rs:46-66 projections for the TURN layer's points: histogram over the river-equity buckets of the 46
// children (histogram.rs:168-176: children().map(equity).map(Abstraction::from))
inline void turn_histogram(const Obs& turn, uint8_t* hist101) {
    for (int b = 0; b < 101; ++b) hist101[b] = 0;
    Obs kids[52];
    const int n = children(turn, kids);
    for (int k = 0; k < n; ++k) hist101[equity_bucket(river_equity(kids[k].pocket, kids[k].pub))] += 1;
}
// generic projection through a next-street lookup table (sorted by (pocket, public) = enumeration order)
inline int lookup_bucket(const std::vector<Obs>& next_isos, const std::vector<uint8_t>& next_abs, const Obs& child) {
    const Obs c = canonical(child);
    auto it = std::lower_bound(next_isos.begin(), next_isos.end(), c, [](const Obs& a, const Obs& b) {
        return a.pocket != b.pocket ? a.pocket < b.pocket : a.pub < b.pub;
    });
    if (it == next_isos.end() || it->pocket != c.pocket || it->pub != c.pub) return -1;
    return next_abs[it - next_isos.begin()];
}

}  // namespace orc
