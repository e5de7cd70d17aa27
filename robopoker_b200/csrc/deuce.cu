// deuce.cu — 5..7-card hand strength and river equity on sm_100a (integer ALU work; bit-exact bar).
//
// Replaces `Strength::from(Hand)` (crates/deuce/src/strength.rs:19-31 → evaluator.rs:39-177) and
// `Observation::equity` (crates/deuce/src/observation.rs:45-62), the inner loop of the river abstraction layer
// (`Lookup::grow(Street::Rive)`, crates/lloyd/src/lookup.rs:177-184: 123,156,254 isomorphisms x 990 villain holes).
//
// The reference walks ranks with loops; here the 52-bit hand (card = 4*rank + suit, hand.rs) is reduced with SWAR:
// per-rank counts in nibbles, "count >= k" flags gathered into 13-bit rank masks by a 4-step bit-gather, and the
// category is a short decision chain over those masks.  Packed result keeps the reference's derived `Ord`
// (ranking.rs:33-44 — FullHouse < Flush in the default build, as written):
//   bits 24-27 tag | 20-23 first rank | 16-19 second rank | 0-12 kicker rank bits.
#include "cards.cuh"

namespace rbp {


__global__ void __launch_bounds__(256) eval_kernel(const uint64_t* __restrict__ hands, int64_t n, uint32_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = strength_of(hands[i]);
}

// One warp per observation: the 990 villain holes C(45,2) are split over the 32 lanes; the 45 unseen cards are
// listed once per warp in shared memory.  16 B in, 4 B (+1 B bucket) out, ~990 evaluations: ALU-bound.
constexpr int kEquityWarps = 8;
__global__ void __launch_bounds__(kEquityWarps * 32)
river_equity_kernel(const uint64_t* __restrict__ pocket, const uint64_t* __restrict__ pub, int64_t n, float* __restrict__ equity,
                    uint8_t* __restrict__ bucket, uint32_t* __restrict__ wins, uint32_t* __restrict__ total) {
    __shared__ uint8_t s_cards[kEquityWarps][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t i = blockIdx.x * (int64_t)kEquityWarps + warp; i < n; i += (int64_t)gridDim.x * kEquityWarps) {
        const uint64_t board = pub[i] & 0x000FFFFFFFFFFFFFull, seen = (pocket[i] | pub[i]) & 0x000FFFFFFFFFFFFFull;
        const uint64_t free_cards = ~seen & 0x000FFFFFFFFFFFFFull;
        const int m = __popcll(free_cards);  // 45 on the river
        // lane l lists cards 2l and 2l+1 of the unseen set
        for (int k = lane; k < m; k += 32) {
            uint64_t f = free_cards;
            for (int d = 0; d < k; ++d) f &= f - 1;  // k-th set bit (m <= 52, done once per observation)
            s_cards[warp][k] = (uint8_t)(__ffsll((long long)f) - 1);
        }
        __syncwarp();
        const uint32_t hero = strength_of(seen);
        const int pairs = m * (m - 1) / 2;
        uint32_t won = 0, sum = 0;
        // pair p = (a, b), a < b, enumerated row-major; lane strides by 32 and advances (a, b) incrementally
        int a = 0, b = 1 + lane;
        while (b >= m && a < m - 1) { ++a; b = b - m + a + 1; }
        for (int p = lane; p < pairs; p += 32) {
            const uint64_t villain = board | 1ull << s_cards[warp][a] | 1ull << s_cards[warp][b];
            const uint32_t v = strength_of(villain);
            won += hero > v;
            sum += hero != v;
            b += 32;
            while (b >= m && a < m - 1) { ++a; b = b - m + a + 1; }
        }
        for (int d = 16; d > 0; d >>= 1) {
            won += __shfl_xor_sync(0xFFFFFFFFu, won, d);
            sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        }
        if (lane == 0) {
            const float e = sum == 0 ? 0.5f : (float)won / (float)sum;  // observation.rs:58-61
            if (equity) equity[i] = e;
            if (bucket) bucket[i] = (uint8_t)roundf(e * 100.0f);        // kicker/src/abstraction.rs:43-45
            if (wins) wins[i] = won;
            if (total) total[i] = sum;
        }
        __syncwarp();
    }
}

}  // namespace rbp

using namespace rbp;

extern "C" {

int rbp_eval_batch(const uint64_t* hands, int64_t n, uint32_t* strength_out) {
    if (n < 0 || (n > 0 && (!hands || !strength_out))) return RBP_ERR_INVALID;
    if (rbp_device_count() < 1) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    if (n == 0) return RBP_OK;
    uint64_t* d_in = nullptr;
    uint32_t* d_out = nullptr;
    RBP_CUDA(cudaMalloc(&d_in, n * sizeof(uint64_t)));
    RBP_CUDA(cudaMalloc(&d_out, n * sizeof(uint32_t)));
    RBP_CUDA(cudaMemcpy(d_in, hands, n * sizeof(uint64_t), cudaMemcpyHostToDevice));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    eval_kernel<<<blocks, 256>>>(d_in, n, d_out);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemcpy(strength_out, d_out, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    cudaFree(d_in);
    cudaFree(d_out);
    return RBP_OK;
}

int rbp_river_equity_device(const uint64_t* d_pocket, const uint64_t* d_public, int64_t n, float* d_equity, uint8_t* d_bucket,
                            uint32_t* d_wins, uint32_t* d_total, void* stream) {
    if (n < 0 || (n > 0 && (!d_pocket || !d_public))) return RBP_ERR_INVALID;
    if (n == 0) return RBP_OK;
    const int blocks = (int)std::min<int64_t>((n + kEquityWarps - 1) / kEquityWarps, 148 * 8);
    river_equity_kernel<<<blocks, kEquityWarps * 32, 0, (cudaStream_t)stream>>>(d_pocket, d_public, n, d_equity, d_bucket, d_wins, d_total);
    RBP_LAUNCHED();
    return RBP_OK;
}

int rbp_river_equity_batch(const uint64_t* pocket, const uint64_t* pub, int64_t n, float* equity_out, uint8_t* bucket_out,
                           uint32_t* wins_out, uint32_t* total_out) {
    if (n < 0 || (n > 0 && (!pocket || !pub))) return RBP_ERR_INVALID;
    if (rbp_device_count() < 1) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    if (n == 0) return RBP_OK;
    uint64_t *d_p = nullptr, *d_b = nullptr;
    float* d_e = nullptr;
    uint8_t* d_k = nullptr;
    uint32_t *d_w = nullptr, *d_t = nullptr;
    RBP_CUDA(cudaMalloc(&d_p, n * 8)); RBP_CUDA(cudaMalloc(&d_b, n * 8));
    RBP_CUDA(cudaMalloc(&d_e, n * 4)); RBP_CUDA(cudaMalloc(&d_k, n));
    RBP_CUDA(cudaMalloc(&d_w, n * 4)); RBP_CUDA(cudaMalloc(&d_t, n * 4));
    RBP_CUDA(cudaMemcpy(d_p, pocket, n * 8, cudaMemcpyHostToDevice));
    RBP_CUDA(cudaMemcpy(d_b, pub, n * 8, cudaMemcpyHostToDevice));
    int st = rbp_river_equity_device(d_p, d_b, n, d_e, d_k, d_w, d_t, nullptr);
    if (st) return st;
    if (equity_out) RBP_CUDA(cudaMemcpy(equity_out, d_e, n * 4, cudaMemcpyDeviceToHost));
    if (bucket_out) RBP_CUDA(cudaMemcpy(bucket_out, d_k, n, cudaMemcpyDeviceToHost));
    if (wins_out) RBP_CUDA(cudaMemcpy(wins_out, d_w, n * 4, cudaMemcpyDeviceToHost));
    if (total_out) RBP_CUDA(cudaMemcpy(total_out, d_t, n * 4, cudaMemcpyDeviceToHost));
    RBP_CUDA(cudaDeviceSynchronize());
    cudaFree(d_p); cudaFree(d_b); cudaFree(d_e); cudaFree(d_k); cudaFree(d_w); cudaFree(d_t);
    return RBP_OK;
}

}  // extern "C"
