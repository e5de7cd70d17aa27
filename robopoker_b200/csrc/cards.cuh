// cards.cuh — device-side card primitives shared by deuce.cu, iso.cu and nlhe.cu: the SWAR hand evaluator
// (`Strength::from(Hand)`, crates/deuce/src/evaluator.rs:39-177) and suit canonicalisation
// (`Isomorphism::from(Observation)`, crates/deuce/src/isomorphism.rs:9-15, permutation.rs:9-66).
#pragma once
#include "common.cuh"

namespace rbp {

__host__ __device__ __forceinline__ uint32_t gather_nibbles(uint64_t x) {  // bit 4i -> bit i, i < 16
    x &= 0x1111111111111111ull;
    x = (x | (x >> 3)) & 0x0303030303030303ull;
    x = (x | (x >> 6)) & 0x000F000F000F000Full;
    x = (x | (x >> 12)) & 0x000000FF000000FFull;
    x = (x | (x >> 24)) & 0xFFFFull;
    return (uint32_t)x;
}
__device__ __forceinline__ int msb(uint32_t v) { return 31 - __clz(v); }
__device__ __forceinline__ uint32_t top_n(uint32_t k, int n) {  // evaluator.rs:51-68: drop low ranks until n remain
    int c = __popc(k);
    while (c > n) { k &= k - 1; --c; }
    return k;
}
__device__ __forceinline__ int straight_top(uint32_t ranks) {  // evaluator.rs:114-130
    uint32_t b = ranks & (ranks << 1);
    b &= b << 2;            // runs of 4
    b &= ranks << 4;        // runs of 5, marked at the top rank
    b &= 0x1FFFu;
    if (b) return msb(b);
    return (ranks & 0x100Fu) == 0x100Fu ? 3 : -1;  // wheel -> Five
}

__device__ __forceinline__ uint32_t strength_of(uint64_t h) {
    h &= 0x000FFFFFFFFFFFFFull;
    // per-rank counts (nibble = 0..4)
    uint64_t c = (h & 0x5555555555555555ull) + ((h >> 1) & 0x5555555555555555ull);
    c = (c & 0x3333333333333333ull) + ((c >> 2) & 0x3333333333333333ull);
    const uint32_t m1 = gather_nibbles(c | (c >> 1) | (c >> 2));
    const uint32_t m2 = gather_nibbles((c >> 1) | (c >> 2));
    const uint32_t m3 = gather_nibbles((c >> 2) | ((c >> 1) & c));
    const uint32_t m4 = gather_nibbles(c >> 2);
    // first suit (C,D,H,S) holding >= 5 cards (evaluator.rs:138-148); at most one exists for <= 9 cards
    int suit = -1;
#pragma unroll
    for (int s = 3; s >= 0; --s)
        if (__popcll(h & (0x0001111111111111ull << s)) >= 5) suit = s;
    uint32_t suited = 0;
    if (suit >= 0) {
        suited = gather_nibbles(h >> suit);
        const int sf = straight_top(suited);
        if (sf >= 0) return 8u << 24 | (uint32_t)sf << 20;                                    // StraightFlush
    }
    if (m4) { const int q = msb(m4); return 7u << 24 | (uint32_t)q << 20 | top_n(m1 & ~(1u << q), 1); }  // FourOAK
    const int t = m3 ? msb(m3) : -1;
    if (t >= 0) {
        const uint32_t rest = m2 & ~(1u << t);
        if (rest) return 5u << 24 | (uint32_t)t << 20 | (uint32_t)msb(rest) << 16;             // FullHouse
    }
    if (suit >= 0) return 6u << 24 | (uint32_t)msb(suited) << 20;                              // Flush(top rank only)
    const int st = straight_top(m1);
    if (st >= 0) return 4u << 24 | (uint32_t)st << 20;                                         // Straight
    if (t >= 0) return 3u << 24 | (uint32_t)t << 20 | top_n(m1 & ~(1u << t), 2);               // ThreeOAK
    if (m2) {
        const int hi = msb(m2);
        const uint32_t rest = m2 & ~(1u << hi);
        if (rest) { const int lo = msb(rest); return 2u << 24 | (uint32_t)hi << 20 | (uint32_t)lo << 16 | top_n(m1 & ~(1u << hi) & ~(1u << lo), 1); }
        return 1u << 24 | (uint32_t)hi << 20 | top_n(m1 & ~(1u << hi), 3);                     // OnePair
    }
    const int h1 = msb(m1);
    return (uint32_t)h1 << 20 | top_n(m1 & ~(1u << h1), 4);                                    // HighCard
}

__device__ __forceinline__ uint32_t suit_key(uint64_t pocket, uint64_t pub, int s) {  // permutation.rs:40-54 (without the suit tiebreak)
    const uint64_t m = 0x0001111111111111ull << s;
    const uint64_t p = pocket & m, b = pub & m;
    const uint32_t pmin = p ? (uint32_t)((__ffsll((long long)p) - 1) >> 2) + 1u : 0u;   // Option<Rank>: None < Some
    const uint32_t bmin = b ? (uint32_t)((__ffsll((long long)b) - 1) >> 2) + 1u : 0u;
    const uint32_t pmax = p ? (uint32_t)((63 - __clzll((long long)p)) >> 2) + 1u : 0u;
    const uint32_t bmax = b ? (uint32_t)((63 - __clzll((long long)b)) >> 2) + 1u : 0u;
    return (uint32_t)__popcll(p) << 20 | (uint32_t)__popcll(b) << 16 | pmin << 12 | bmin << 8 | pmax << 4 | bmax;
}
__device__ __forceinline__ bool is_canonical(uint64_t pocket, uint64_t pub) {  // isomorphism.rs:40-44
    const uint32_t k0 = suit_key(pocket, pub, 0), k1 = suit_key(pocket, pub, 1), k2 = suit_key(pocket, pub, 2), k3 = suit_key(pocket, pub, 3);
    return k0 <= k1 && k1 <= k2 && k2 <= k3;  // stable sort with the suit as tiebreak leaves equal keys in place
}
__device__ __forceinline__ void canonicalize(uint64_t& pocket, uint64_t& pub) {  // permutation.rs:9-33,55-66
    uint32_t k[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) k[s] = suit_key(pocket, pub, s) << 2 | (uint32_t)s;  // suit id = final tiebreak
    int perm[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        int r = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) r += k[t] < k[s];
        perm[s] = r;  // rank of suit s in the sorted order = its new suit
    }
    uint64_t np = 0, nb = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint64_t m = 0x0001111111111111ull << s;
        const int sh = perm[s] - s;
        const uint64_t p = pocket & m, b = pub & m;
        np |= sh >= 0 ? p << sh : p >> -sh;
        nb |= sh >= 0 ? b << sh : b >> -sh;
    }
    pocket = np; pub = nb;
}

}  // namespace rbp
