// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under robopoker_b200/ may include, link or call this.
//
// Counter-based RNG contract shared (by specification, not by code) between the oracle and the
// CUDA path.  The reference seeds a fresh SmallRng per sampled node from
// SipHash13(epoch, info, tree_id) (crates/mccfr/src/strategy/flow.rs:285-295) and draws root deals
// from the thread RNG (crates/leduc/src/game.rs:177-185, crates/kuhn/src/game.rs:115-123); neither
// stream is reproducible outside Rust's std/rand (SURVEY §8c "parity unpinned").  The contract below
// keeps the one property the reference relies on — the same (epoch, tree, infoset) always draws the
// same number — and replaces the generator by Philox4x32-10 (Salmon et al., SC'11):
//
//   key     = (seed_lo, seed_hi)
//   counter = (epoch, tree_id, info_key, tag)      tag: 0 node draw, 1 root deal, 2 pluribus coin
//   range(n)     = (u64(r0) * n) >> 32                      (stands in for rand::random_range)
//   unit()       = (r0 >> 8) * 2^-24                        (stands in for rand::random::<f32>)
//   weighted(w)  = first i with unit()*Σw < Σ_{j<=i} w_j, sequential f32 sums, else last index
//                                                           (stands in for WeightedIndex<f32>)
#pragma once
#include <cstdint>

namespace orc {

struct Philox4 {
    uint32_t r[4];
};

inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

enum : uint32_t { TAG_NODE = 0, TAG_ROOT = 1, TAG_COIN = 2, TAG_KMEANSPP = 3 };

struct Draw {
    uint64_t seed;
    Philox4 at(uint32_t epoch, uint32_t tree, uint32_t info, uint32_t tag) const {
        return philox4x32_10(epoch, tree, info, tag, (uint32_t)seed, (uint32_t)(seed >> 32));
    }
};

inline uint32_t draw_range(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }
inline float draw_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
inline int draw_weighted(uint32_t r, const float* w, int n) {
    float total = 0.0f;
    for (int i = 0; i < n; ++i) total = total + w[i];
    float x = draw_unit(r) * total;
    float cum = 0.0f;
    for (int i = 0; i < n; ++i) {
        cum = cum + w[i];
        if (x < cum) return i;
    }
    return n - 1;
}

}  // namespace orc
