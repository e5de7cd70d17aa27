// pool.hpp — TEST INFRASTRUCTURE (oracle): a persistent worker pool standing in for rayon's global pool
// (crates/mccfr/src/solver/solver.rs:225-240 `into_par_iter().map().collect()`): threads are created once, chunks of the
// index range are claimed dynamically (work stealing's load balance: sampled trees vary 50-3500 nodes), and results are
// concatenated in index order, like rayon's order-preserving collect.  Not part of the product path.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace orc {

class Pool {
  public:
    explicit Pool(int threads) : n_(threads < 1 ? 1 : threads) {
        for (int t = 1; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; ++generation_; }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    int size() const { return n_; }
    // fn(chunk index, thread index) for every chunk in [0, chunks); returns when all are done
    void run(int chunks, const std::function<void(int, int)>& fn) {
        if (n_ == 1 || chunks <= 1) { for (int c = 0; c < chunks; ++c) fn(c, 0); return; }
        { std::lock_guard<std::mutex> g(m_); fn_ = &fn; chunks_ = chunks; next_.store(0); pending_ = n_ - 1; ++generation_; }
        cv_.notify_all();
        drain(0);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void drain(int t) { for (int c; (c = next_.fetch_add(1)) < chunks_;) (*fn_)(c, t); }
    void loop(int t) {
        uint64_t seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> g(m_); cv_.wait(g, [&] { return generation_ != seen; }); seen = generation_; if (stop_) return; }
            drain(t);
            { std::lock_guard<std::mutex> g(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)>* fn_ = nullptr;
    std::atomic<int> next_{0};
    int chunks_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

}  // namespace orc
