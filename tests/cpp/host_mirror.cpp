// Exercises include/rbp.hpp (the C++ host-side mirror of the reference's trait surface) against librbp_b200.so.
// Without a CUDA device: every owner type must refuse with RBP_ERR_NO_DEVICE (no CPU fallback).
// With a device: a short Kuhn / Leduc run through the mirror (the parity bar itself is tests/test_*_gpu.py).
#include <cmath>
#include <cstdio>
#include <vector>

#include "rbp.hpp"

template <class F>
static int refuses(const char* what, F make) {
    try {
        make();
    } catch (const rbp::Error& e) {
        if (e.status == RBP_ERR_NO_DEVICE) return 0;
        std::printf("%s: wrong status %d (%s)\n", what, e.status, e.what());
        return 1;
    }
    std::printf("%s: computed without a device\n", what);
    return 1;
}

int main() {
    int bad = 0;
    if (rbp_device_count() < 1) {
        std::vector<uint8_t> turn(4 * 101, 0), flop(4 * 32, 0);
        for (int i = 0; i < 4; ++i) { turn[i * 101 + 50] = 46; flop[i * 32 + 3] = 47; }
        bad += refuses("Solver", [] { rbp::Solver s(rbp::Game::Kuhn); });
        bad += refuses("Nlhe", [] { rbp::Nlhe n(8, 0, rbp::Regret::Linear, rbp::Weight::Linear, rbp::Sampling::Pluribus, 1 << 10); });
        bad += refuses("Layer/W1", [&] { rbp::Layer l(turn.data(), 4, 2, 101); });
        bad += refuses("Layer/Sinkhorn", [&] { rbp::Layer l(flop.data(), 4, 2, 32, true); });
        bad += refuses("IsoSet", [] { rbp::IsoSet s(rbp::Street::Flop); });
        bad += refuses("strength", [] { rbp::strength({0x7Full}); });
        bad += refuses("river_equity", [] { rbp::river_equity({0x3ull}, {0x7Cull}); });
        std::printf("no device: %d failures\n", bad);
        return bad;
    }
    rbp::Solver kuhn(rbp::Game::Kuhn, rbp::Regret::Floored, rbp::Weight::Linear, rbp::Sampling::External, 64, 0);
    const float before = kuhn.exploitability();
    kuhn.solve(1 << 16);
    const float after = kuhn.exploitability();
    if (!(std::fabs(before - 0.425f) < 1e-5f) || !(after < 0.05f) || kuhn.epochs() != (1u << 16) / 64) {
        std::printf("kuhn: exploitability %.5f -> %.5f after %llu epochs\n", before, after, (unsigned long long)kuhn.epochs());
        ++bad;
    }
    auto rows = kuhn.profile();
    rbp::Solver copy(rbp::Game::Kuhn, rbp::Regret::Floored, rbp::Weight::Linear, rbp::Sampling::External, 64, 0);
    copy.storage(rows, kuhn.epochs());
    if (copy.exploitability() != after || copy.profile().size() != rows.size()) { std::printf("profile round trip differs\n"); ++bad; }
    auto s = rbp::strength({0x1Full /* 2c2d2h2s3c: four deuces */, 0x7Full});
    if (!((s[0] >> 24) == 7u)) { std::printf("strength tag %u\n", s[0] >> 24); ++bad; }
    rbp::Solver moved = std::move(copy);
    (void)moved.step(1);
    std::printf("device: %d failures, kuhn exploitability %.5f\n", bad, after);
    return bad;
}
