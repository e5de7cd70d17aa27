// ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes-facing C API over oracle/iso.hpp.
#include <thread>

#include "iso.hpp"

using namespace orc;

extern "C" {
// canonical observations of a street (enumeration order); returns the count, fills up to `cap` (pocket, public) pairs
int64_t orc_isomorphisms(int street, uint64_t* pocket_out, uint64_t* public_out, int64_t cap, int threads) {
    if (threads < 1) threads = 1;
    std::vector<std::vector<Obs>> parts(threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([&, t]() { enumerate_isomorphisms(street, parts[t], 1326 * t / threads, 1326 * (t + 1) / threads); });
    for (auto& x : th) x.join();
    int64_t n = 0;
    for (auto& p : parts)
        for (const Obs& o : p) {
            if (n < cap) { if (pocket_out) pocket_out[n] = o.pocket; if (public_out) public_out[n] = o.pub; }
            ++n;
        }
    return n;
}
void orc_canonical_batch(const uint64_t* pocket, const uint64_t* pub, int64_t n, uint64_t* pocket_out, uint64_t* public_out, uint8_t* is_canon) {
    for (int64_t i = 0; i < n; ++i) {
        Obs o{pocket[i], pub[i]}, c = canonical(o);
        pocket_out[i] = c.pocket; public_out[i] = c.pub;
        if (is_canon) is_canon[i] = is_canonical(o) ? 1 : 0;
    }
}
// turn-layer points: histogram over river-equity buckets of the 46 children of each turn observation
void orc_turn_histograms(const uint64_t* pocket, const uint64_t* pub, int64_t n, uint8_t* hist, int threads) {
    if (threads < 1) threads = 1;
    auto work = [&](int t) { for (int64_t i = t; i < n; i += threads) turn_histogram(Obs{pocket[i], pub[i]}, hist + i * 101); };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}
// generic projection: children of each observation looked up (after canonicalisation) in the next street's table
void orc_project(const uint64_t* pocket, const uint64_t* pub, int64_t n, const uint64_t* next_pocket, const uint64_t* next_public,
                 const uint8_t* next_abs, int64_t next_n, int bins, uint8_t* hist, int threads) {
    std::vector<Obs> isos(next_n);
    std::vector<uint8_t> abs(next_abs, next_abs + next_n);
    for (int64_t i = 0; i < next_n; ++i) isos[i] = Obs{next_pocket[i], next_public[i]};
    if (threads < 1) threads = 1;
    auto work = [&](int t) {
        for (int64_t i = t; i < n; i += threads) {
            uint8_t* h = hist + i * bins;
            for (int b = 0; b < bins; ++b) h[b] = 0;
            Obs kids[52];
            const int k = children(Obs{pocket[i], pub[i]}, kids);
            for (int c = 0; c < k; ++c) { int b = lookup_bucket(isos, abs, kids[c]); if (b >= 0) h[b] += 1; }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}
}
