// ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes-facing C API over oracle/deuce.hpp.
#include <thread>
#include <vector>

#include "deuce.hpp"

extern "C" {
void orc_eval_batch(const uint64_t* hands, int64_t n, uint32_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = orc::strength(hands[i]);
}
void orc_river_equity_batch(const uint64_t* pocket, const uint64_t* pub, int64_t n, float* equity, uint8_t* bucket, uint32_t* wins,
                            uint32_t* total, int threads) {
    if (threads < 1) threads = 1;
    auto work = [&](int t) {
        for (int64_t i = t; i < n; i += threads) {
            uint32_t w, s;
            float e = orc::river_equity(pocket[i], pub[i], &w, &s);
            if (equity) equity[i] = e;
            if (bucket) bucket[i] = orc::equity_bucket(e);
            if (wins) wins[i] = w;
            if (total) total[i] = s;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}
}
