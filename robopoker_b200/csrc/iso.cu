// iso.cu — suit-isomorphism enumeration, canonicalisation and histogram projections on sm_100a (integer work; bit-exact).
//
// Replaces, for the abstraction pipeline (crates/lloyd/src/lookup.rs:46-66,177-192):
//   IsomorphismIterator::from(street)   crates/deuce/src/isomorphism_iter.rs:7-21 (+ observation_iter.rs, hand_iter.rs)
//   Isomorphism::from(Observation)      crates/deuce/src/isomorphism.rs:9-15, permutation.rs:9-66
//   Observation::children               crates/deuce/src/observation.rs:35-40
//   Lookup::projections                 children → canonical form → BTreeMap lookup → Histogram::increment
//
// The reference walks 2.8 G river observations with Gosper's hack and filters the canonical ones.  Here an
// observation is addressed by (pocket index, colex rank of the board among the 50 other cards) and UNRANKED in
// parallel; Gosper order = ascending mask value = colex order, so a stable compaction of the canonical ones
// reproduces the reference's enumeration order exactly.  Only 247 of the 1326 pockets can be canonical at all (the
// sort key starts with the per-suit pocket count), which prunes the search 5x before any board is looked at.
// A set is stored sorted by (pocket, public) — its enumeration order — so projections use binary search instead of
// the reference's BTreeMap.
#include <algorithm>
#include <vector>

#include "cards.cuh"
#include "isoset.hpp"

namespace rbp {

__constant__ unsigned long long c_binom[53][6];  // C(n, k), n <= 52, k <= 5


// t-th k-subset (colex order) of the cards not in `skip`, as a card mask
__device__ __forceinline__ uint64_t unrank_board(unsigned long long t, int k, uint64_t skip) {
    const int skipped = __popcll(skip);
    uint64_t board = 0;
    int hi = 52 - skipped;  // remaining-card indices are < hi
    for (int j = k; j >= 1; --j) {
        int c = j - 1;
        // largest c < hi with C(c, j) <= t
        int lo = j - 1, up = hi - 1;
        while (lo < up) { const int mid = (lo + up + 1) >> 1; if (c_binom[mid][j] <= t) lo = mid; else up = mid - 1; }
        c = lo;
        t -= c_binom[c][j];
        hi = c;
        // c-th remaining card → actual card: skip over the excluded cards below it
        int card = c;
        uint64_t sk = skip;
        while (sk) { const int b = __ffsll((long long)sk) - 1; if (b <= card) ++card; else break; sk &= sk - 1; }
        board |= 1ull << card;
    }
    return board;
}

// pass 1: canonical flags counted per block; pass 2: stable write at the scanned offsets
template <bool WRITE>
__global__ void __launch_bounds__(256)
enumerate_kernel(const uint64_t* __restrict__ pockets, int n_pockets, unsigned long long boards_per_pocket, int nb_cards,
                 unsigned long long* __restrict__ block_counts, const unsigned long long* __restrict__ block_offsets,
                 uint64_t* __restrict__ out_pocket, uint64_t* __restrict__ out_public) {
    __shared__ unsigned int s_warp[8];
    const unsigned long long total = (unsigned long long)n_pockets * boards_per_pocket;
    const unsigned long long gid = blockIdx.x * 256ull + threadIdx.x;
    bool keep = false;
    uint64_t pocket = 0, board = 0;
    if (gid < total) {
        const int pi = (int)(gid / boards_per_pocket);
        const unsigned long long t = gid - (unsigned long long)pi * boards_per_pocket;
        pocket = pockets[pi];
        board = nb_cards ? unrank_board(t, nb_cards, pocket) : 0ull;
        keep = is_canonical(pocket, board);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (!WRITE) {
        if (threadIdx.x == 0) { unsigned c = 0; for (int w = 0; w < 8; ++w) c += s_warp[w]; block_counts[blockIdx.x] = c; }
    } else if (keep) {
        unsigned long long pos = block_offsets[blockIdx.x];
        for (int w = 0; w < warp; ++w) pos += s_warp[w];
        pos += __popc(m & ((1u << lane) - 1u));
        out_pocket[pos] = pocket;
        out_public[pos] = board;
    }
}
__global__ void scan_blocks_kernel(const unsigned long long* __restrict__ counts, long long n, unsigned long long* __restrict__ offsets,
                                   unsigned long long* __restrict__ total) {
    // single block, sequential over chunks: n is at most a few million
    __shared__ unsigned long long s_part[1024];
    const int tid = threadIdx.x;
    const long long per = (n + 1023) / 1024;
    const long long lo = min(n, tid * per), hi = min(n, lo + per);
    unsigned long long mine = 0;
    for (long long b = lo; b < hi; ++b) mine += counts[b];
    s_part[tid] = mine;
    __syncthreads();
    if (tid == 0) { unsigned long long run = 0; for (int t = 0; t < 1024; ++t) { const unsigned long long v = s_part[t]; s_part[t] = run; run += v; } *total = run; }
    __syncthreads();
    unsigned long long run = s_part[tid];
    for (long long b = lo; b < hi; ++b) { offsets[b] = run; run += counts[b]; }
}

__global__ void canonical_kernel(const uint64_t* __restrict__ pocket, const uint64_t* __restrict__ pub, int64_t n, uint64_t* __restrict__ po,
                                 uint64_t* __restrict__ bo, uint8_t* __restrict__ flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t p = pocket[i], b = pub[i];
    if (flag) flag[i] = is_canonical(p, b) ? 1 : 0;
    canonicalize(p, b);
    po[i] = p; bo[i] = b;
}

// index of (pocket, public) in a set sorted by (pocket, public); -1 if absent
__device__ __forceinline__ long long find_iso(const uint64_t* __restrict__ sp, const uint64_t* __restrict__ sb, long long n, uint64_t p, uint64_t b) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const uint64_t mp = sp[mid];
        const bool less = mp < p || (mp == p && sb[mid] < b);
        if (less) lo = mid + 1; else hi = mid;
    }
    return (lo < n && sp[lo] == p && sb[lo] == b) ? lo : -1;
}

// Lookup::projections (lookup.rs:46-66): one thread per parent observation; children in HandIterator order
// (ascending card, observation.rs:35-40), canonicalised, looked up in the child street's table, tallied.
__global__ void __launch_bounds__(128)
project_kernel(const uint64_t* __restrict__ pp, const uint64_t* __restrict__ pb, int64_t n, const uint64_t* __restrict__ cp,
               const uint64_t* __restrict__ cb, const uint8_t* __restrict__ cabs, long long cn, int bins, uint8_t* __restrict__ hist,
               unsigned long long* __restrict__ misses) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t pocket = pp[i], pub = pb[i];
    uint8_t* h = hist + (size_t)i * bins;
    for (int b = 0; b < bins; ++b) h[b] = 0;
    uint64_t free_cards = ~(pocket | pub) & 0x000FFFFFFFFFFFFFull;
    unsigned long long miss = 0;
    while (free_cards) {
        const uint64_t card = free_cards & (0 - free_cards);
        free_cards &= free_cards - 1;
        uint64_t p = pocket, b = pub | card;
        canonicalize(p, b);
        const long long at = find_iso(cp, cb, cn, p, b);
        if (at >= 0) h[cabs[at]] += 1; else ++miss;
    }
    if (miss) atomicAdd(misses, miss);
}

}  // namespace rbp

using namespace rbp;

namespace {
int board_cards(int street) { return street == 0 ? 0 : street + 2; }
bool g_binom_ready[16] = {false};
int ensure_binom(int device) {
    if (device < 16 && g_binom_ready[device]) return RBP_OK;
    unsigned long long b[53][6];
    for (int n = 0; n <= 52; ++n)
        for (int k = 0; k <= 5; ++k) b[n][k] = k == 0 ? 1ull : (n == 0 ? 0ull : b[n - 1][k - 1] + b[n - 1][k]);
    RBP_CUDA(cudaMemcpyToSymbol(c_binom, b, sizeof b));
    if (device < 16) g_binom_ready[device] = true;
    return RBP_OK;
}
}  // namespace

extern "C" {

int rbp_isoset_create(int street, int device, rbp_isoset_t** out) {
    if (!out || street < 0 || street > 3) return RBP_ERR_INVALID;
    *out = nullptr;
    if (rbp_device_count() <= device) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    RBP_CUDA(cudaSetDevice(device));
    int st = ensure_binom(device);
    if (st) return st;
    // pockets in Gosper (ascending mask) order that can be canonical: per-suit pocket counts must be non-decreasing
    std::vector<uint64_t> pockets;
    for (int hi = 1; hi < 52; ++hi)
        for (int lo = 0; lo < hi; ++lo) pockets.push_back(1ull << hi | 1ull << lo);
    std::sort(pockets.begin(), pockets.end());
    std::vector<uint64_t> keep;
    for (uint64_t p : pockets) {
        int c[4];
        for (int s = 0; s < 4; ++s) c[s] = __builtin_popcountll(p & (0x0001111111111111ull << s));
        if (c[0] <= c[1] && c[1] <= c[2] && c[2] <= c[3]) keep.push_back(p);
    }
    const int nbc = board_cards(street);
    unsigned long long per = 1;
    for (int j = 1; j <= nbc; ++j) per = per * (unsigned long long)(50 - j + 1) / j;  // C(50, nbc)
    const unsigned long long total = (unsigned long long)keep.size() * per;
    const long long blocks = (long long)((total + 255) / 256);
    uint64_t* d_pockets = nullptr;
    unsigned long long *d_counts = nullptr, *d_offsets = nullptr, *d_total = nullptr;
    RBP_CUDA(cudaMalloc(&d_pockets, keep.size() * 8));
    RBP_CUDA(cudaMemcpy(d_pockets, keep.data(), keep.size() * 8, cudaMemcpyHostToDevice));
    RBP_CUDA(cudaMalloc(&d_counts, blocks * 8));
    RBP_CUDA(cudaMalloc(&d_offsets, blocks * 8));
    RBP_CUDA(cudaMalloc(&d_total, 8));
    enumerate_kernel<false><<<(unsigned)blocks, 256>>>(d_pockets, (int)keep.size(), per, nbc, d_counts, nullptr, nullptr, nullptr);
    RBP_LAUNCHED();
    scan_blocks_kernel<<<1, 1024>>>(d_counts, blocks, d_offsets, d_total);
    RBP_LAUNCHED();
    unsigned long long n = 0;
    RBP_CUDA(cudaMemcpy(&n, d_total, 8, cudaMemcpyDeviceToHost));
    rbp_isoset* h = new rbp_isoset();
    h->street = street; h->device = device; h->n = (int64_t)n;
    RBP_CUDA(cudaMalloc(&h->pocket, std::max<size_t>(n, 1) * 8));
    RBP_CUDA(cudaMalloc(&h->pub, std::max<size_t>(n, 1) * 8));
    RBP_CUDA(cudaMalloc(&h->abs, std::max<size_t>(n, 1)));
    enumerate_kernel<true><<<(unsigned)blocks, 256>>>(d_pockets, (int)keep.size(), per, nbc, nullptr, d_offsets, h->pocket, h->pub);
    RBP_LAUNCHED();
    RBP_CUDA(cudaDeviceSynchronize());
    cudaFree(d_pockets); cudaFree(d_counts); cudaFree(d_offsets); cudaFree(d_total);
    *out = h;
    return RBP_OK;
}
void rbp_isoset_destroy(rbp_isoset_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->pocket); cudaFree(h->pub); cudaFree(h->abs);
    delete h;
}
int64_t rbp_isoset_size(rbp_isoset_t* h) { return h ? h->n : -1; }
int rbp_isoset_export(rbp_isoset_t* h, int64_t offset, int64_t count, uint64_t* pocket_out, uint64_t* public_out, uint8_t* abs_out) {
    if (!h || offset < 0 || count < 0 || offset + count > h->n) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    if (pocket_out) RBP_CUDA(cudaMemcpy(pocket_out, h->pocket + offset, count * 8, cudaMemcpyDeviceToHost));
    if (public_out) RBP_CUDA(cudaMemcpy(public_out, h->pub + offset, count * 8, cudaMemcpyDeviceToHost));
    if (abs_out) {
        if (!h->have_abs) { set_last_error("isoset has no abstraction column yet"); return RBP_ERR_STATE; }
        RBP_CUDA(cudaMemcpy(abs_out, h->abs + offset, count, cudaMemcpyDeviceToHost));
    }
    return RBP_OK;
}
// `i64::from(Observation)` / `Observation::from(i64)` (crates/deuce/src/observation.rs:130-163): the cards of the board, then of the
// pocket, each in ascending card order, one byte per card (1 + card), first card in the highest used byte.
void rbp_obs_encode(const uint64_t* pocket, const uint64_t* pub, int64_t n, int64_t* obs_out) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t acc = 0;
        for (uint64_t h : {pub[i], pocket[i]})
            for (; h; h &= h - 1) acc = acc << 8 | (uint64_t)(1 + __builtin_ctzll(h));
        obs_out[i] = (int64_t)acc;
    }
}
void rbp_obs_decode(const int64_t* obs, int64_t n, uint64_t* pocket_out, uint64_t* pub_out) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t pocket = 0, pub = 0;
        int k = 0;
        for (uint64_t bits = (uint64_t)obs[i]; bits & 0xFF; bits >>= 8, ++k) {
            const uint64_t card = 1ull << ((bits & 0xFF) - 1);
            if (k < 2) pocket |= card; else pub |= card;
        }
        pocket_out[i] = pocket; pub_out[i] = pub;
    }
}
// rows of the reference's `isomorphism` table (crates/lloyd/src/lookup.rs Streamable rows: obs i64, abs i16)
int rbp_isoset_export_rows(rbp_isoset_t* h, int64_t offset, int64_t count, int64_t* obs_out, int16_t* abs_out) {
    if (!h || !obs_out || offset < 0 || count < 0 || offset + count > h->n) return RBP_ERR_INVALID;
    if (abs_out && !h->have_abs) { set_last_error("isoset has no abstraction column yet"); return RBP_ERR_STATE; }
    RBP_CUDA(cudaSetDevice(h->device));
    std::vector<uint64_t> p(count), b(count);
    std::vector<uint8_t> a(abs_out ? count : 0);
    RBP_CUDA(cudaMemcpy(p.data(), h->pocket + offset, count * 8, cudaMemcpyDeviceToHost));
    RBP_CUDA(cudaMemcpy(b.data(), h->pub + offset, count * 8, cudaMemcpyDeviceToHost));
    rbp_obs_encode(p.data(), b.data(), count, obs_out);
    if (abs_out) {
        RBP_CUDA(cudaMemcpy(a.data(), h->abs + offset, count, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < count; ++i) abs_out[i] = (int16_t)(h->street << 8 | a[i]);  // kicker/src/abstraction.rs:15-60
    }
    return RBP_OK;
}
int rbp_isoset_set_abstractions(rbp_isoset_t* h, const uint8_t* abs) {
    if (!h || !abs) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    RBP_CUDA(cudaMemcpy(h->abs, abs, h->n, cudaMemcpyHostToDevice));
    h->have_abs = true;
    return RBP_OK;
}
// `Lookup::grow(Street::Rive)` (lookup.rs:177-184): abstraction = Abstraction::from(equity) for every river isomorphism
int rbp_isoset_river_buckets(rbp_isoset_t* h) {
    if (!h || h->street != 3) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(h->device));
    int st = rbp_river_equity_device(h->pocket, h->pub, h->n, nullptr, h->abs, nullptr, nullptr, nullptr);
    if (st) return st;
    RBP_CUDA(cudaDeviceSynchronize());
    h->have_abs = true;
    return RBP_OK;
}
// `Lookup::projections` (lookup.rs:46-66): histograms of `parent`'s observations over `child`'s abstraction column
int rbp_isoset_project(rbp_isoset_t* parent, rbp_isoset_t* child, int bins, int64_t offset, int64_t count, uint8_t* hist_out, uint64_t* misses_out) {
    if (!parent || !child || !hist_out || bins < 1 || bins > 256 || offset < 0 || count < 0 || offset + count > parent->n) return RBP_ERR_INVALID;
    if (child->street != parent->street + 1) { set_last_error("child set must be the next street"); return RBP_ERR_INVALID; }
    if (!child->have_abs) { set_last_error("child set has no abstraction column"); return RBP_ERR_STATE; }
    if (count == 0) return RBP_OK;
    RBP_CUDA(cudaSetDevice(parent->device));
    uint8_t* d_hist = nullptr;
    unsigned long long* d_miss = nullptr;
    RBP_CUDA(cudaMalloc(&d_hist, (size_t)count * bins));
    RBP_CUDA(cudaMalloc(&d_miss, 8));
    RBP_CUDA(cudaMemset(d_miss, 0, 8));
    project_kernel<<<(unsigned)((count + 127) / 128), 128>>>(parent->pocket + offset, parent->pub + offset, count, child->pocket, child->pub,
                                                            child->abs, child->n, bins, d_hist, d_miss);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemcpy(hist_out, d_hist, (size_t)count * bins, cudaMemcpyDeviceToHost));
    unsigned long long miss = 0;
    RBP_CUDA(cudaMemcpy(&miss, d_miss, 8, cudaMemcpyDeviceToHost));
    if (misses_out) *misses_out = miss;
    cudaFree(d_hist); cudaFree(d_miss);
    return RBP_OK;
}
// `Isomorphism::from(Observation)` for a batch; flag_out (nullable) = `Isomorphism::is_canonical`
int rbp_canonical_batch(const uint64_t* pocket, const uint64_t* pub, int64_t n, uint64_t* pocket_out, uint64_t* public_out, uint8_t* flag_out) {
    if (n < 0 || (n > 0 && (!pocket || !pub || !pocket_out || !public_out))) return RBP_ERR_INVALID;
    if (rbp_device_count() < 1) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    if (n == 0) return RBP_OK;
    uint64_t *dp, *db, *op, *ob;
    uint8_t* df = nullptr;
    RBP_CUDA(cudaMalloc(&dp, n * 8)); RBP_CUDA(cudaMalloc(&db, n * 8)); RBP_CUDA(cudaMalloc(&op, n * 8)); RBP_CUDA(cudaMalloc(&ob, n * 8));
    RBP_CUDA(cudaMalloc(&df, n));
    RBP_CUDA(cudaMemcpy(dp, pocket, n * 8, cudaMemcpyHostToDevice));
    RBP_CUDA(cudaMemcpy(db, pub, n * 8, cudaMemcpyHostToDevice));
    canonical_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dp, db, n, op, ob, df);
    RBP_LAUNCHED();
    RBP_CUDA(cudaMemcpy(pocket_out, op, n * 8, cudaMemcpyDeviceToHost));
    RBP_CUDA(cudaMemcpy(public_out, ob, n * 8, cudaMemcpyDeviceToHost));
    if (flag_out) RBP_CUDA(cudaMemcpy(flag_out, df, n, cudaMemcpyDeviceToHost));
    cudaFree(dp); cudaFree(db); cudaFree(op); cudaFree(ob); cudaFree(df);
    return RBP_OK;
}

}  // extern "C"
