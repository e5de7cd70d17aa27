// kmeans_api.cu — the `rbp_kmeans_*` C ABI: dispatch on the layer kind (W1 turn layer / Sinkhorn flop layer).
#include "comm.hpp"
#include "kmeans_common.cuh"

struct KmW1;
struct KmSk;
namespace rbp {
int w1_create(int kind, int64_t n, int k, int bins, const uint8_t* counts, int device, rbp_kmeans_t** out);
void w1_destroy(KmW1*);
int w1_init_pp(KmW1*, uint64_t, int32_t*);
int w1_set_centroids(KmW1*, const uint64_t*);
int w1_init_bounds(KmW1*);
int w1_step_local(KmW1*);
int w1_accumulator(KmW1*, void**, size_t*);
int w1_counters(KmW1*, void**, void**);
void* w1_stream(KmW1*);
int w1_step_finish(KmW1*, float*, uint32_t*, uint32_t*);
int w1_assign(KmW1*, uint32_t*, float*);
int w1_centroids(KmW1*, uint64_t*, uint64_t*);
int w1_metric(KmW1*, float*);
int w1_bounds(KmW1*, uint32_t*, float*, float*, uint8_t*);
int w1_timed(KmW1*, int, int, float*);
int sk_create(int64_t n, int k, int bins, const uint8_t* counts, int device, rbp_kmeans_t** out);
void sk_destroy(KmSk*);
int sk_set_metric(KmSk*, const float*, int);
int sk_init_pp(KmSk*, uint64_t, int32_t*);
int sk_set_centroids(KmSk*, const uint64_t*);
int sk_init_bounds(KmSk*);
int sk_step_local(KmSk*);
int sk_accumulator(KmSk*, void**, size_t*);
int sk_counters(KmSk*, void**, void**);
void* sk_stream(KmSk*);
int sk_step_finish(KmSk*, float*, uint32_t*, uint32_t*);
int sk_assign(KmSk*, uint32_t*, float*);
int sk_centroids(KmSk*, uint64_t*, uint64_t*);
int sk_metric(KmSk*, float*);
int sk_bounds(KmSk*, uint32_t*, float*, float*, uint8_t*);
int sk_timed(KmSk*, int, int, float*);
int sk_stats(KmSk*, uint64_t*, int);
int sk_screen(KmSk*, float);
int sk_screen_probe(KmSk*, int64_t, float*, uint64_t*);
int sk_batch(const uint32_t*, int, const uint32_t*, int, int, const int32_t*, const int32_t*, int64_t, const float*, float, int, float, float*);
}  // namespace rbp
using namespace rbp;

#define W1(h) reinterpret_cast<KmW1*>(h)
#define SK(h) reinterpret_cast<KmSk*>(h)
#define DISPATCH(h, call_w1, call_sk) (!(h) ? RBP_ERR_INVALID : ((h)->kind == RBP_KMEANS_W1 ? (call_w1) : (call_sk)))

namespace rbp {
// dependent-free FADD streams: the FP32-add issue ceiling the W1 distance kernels are measured against
__global__ void __launch_bounds__(256) fadd_peak_kernel(float* out, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1.0f, a2 = a0 + 2.0f, a3 = a0 + 3.0f, a4 = a0 + 4.0f, a5 = a0 + 5.0f, a6 = a0 + 6.0f, a7 = a0 + 7.0f;
    const float d = seed * 1.0e-3f;
    for (int i = 0; i < iters; ++i) {
        a0 += d; a1 += d; a2 += d; a3 += d; a4 += d; a5 += d; a6 += d; a7 += d;
        a0 += a4; a1 += a5; a2 += a6; a3 += a7; a4 += d; a5 += d; a6 += d; a7 += d;
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) out[0] = a0;
}
}  // namespace rbp

extern "C" {

int rbp_measure_fadd_peak(float* tera_adds_per_s) {
    if (!tera_adds_per_s) return RBP_ERR_INVALID;
    if (rbp_device_count() < 1) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    float* d = nullptr;
    RBP_CUDA(cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    RBP_CUDA(cudaEventCreate(&e0));
    RBP_CUDA(cudaEventCreate(&e1));
    const int iters = 20000, blocks = 148 * 8;
    fadd_peak_kernel<<<blocks, 256>>>(d, 1000, 1.0f);
    RBP_CUDA(cudaEventRecord(e0));
    fadd_peak_kernel<<<blocks, 256>>>(d, iters, 1.0f);
    RBP_LAUNCHED();
    RBP_CUDA(cudaEventRecord(e1));
    RBP_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    RBP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *tera_adds_per_s = (float)((double)blocks * 256 * 16.0 * iters / (ms * 1e-3) / 1e12);
    cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return RBP_OK;
}

int rbp_kmeans_create(int kind, int64_t n, int k, int bins, const uint8_t* counts, int device, rbp_kmeans_t** out) {
    if (!out) return RBP_ERR_INVALID;
    *out = nullptr;
    if (rbp_device_count() <= device) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    if (kind == RBP_KMEANS_W1) return w1_create(kind, n, k, bins, counts, device, out);
    if (kind == RBP_KMEANS_SINKHORN) return sk_create(n, k, bins, counts, device, out);
    return RBP_ERR_INVALID;
}
void rbp_kmeans_destroy(rbp_kmeans_t* h) {
    if (!h) return;
    if (h->kind == RBP_KMEANS_W1) w1_destroy(W1(h)); else sk_destroy(SK(h));
}
int rbp_kmeans_set_metric(rbp_kmeans_t* h, const float* tri, int bins) {
    if (!h) return RBP_ERR_INVALID;
    if (h->kind == RBP_KMEANS_W1) { set_last_error("the W1 layer has no ground metric (river buckets are ordered)"); return RBP_ERR_STATE; }
    return sk_set_metric(SK(h), tri, bins);
}
int rbp_kmeans_init_pp(rbp_kmeans_t* h, uint64_t seed, int32_t* chosen_out) { return DISPATCH(h, w1_init_pp(W1(h), seed, chosen_out), sk_init_pp(SK(h), seed, chosen_out)); }
int rbp_kmeans_set_centroids(rbp_kmeans_t* h, const uint64_t* counts) {
    if (!counts) return RBP_ERR_INVALID;
    return DISPATCH(h, w1_set_centroids(W1(h), counts), sk_set_centroids(SK(h), counts));
}
int rbp_kmeans_init_bounds(rbp_kmeans_t* h) { return DISPATCH(h, w1_init_bounds(W1(h)), sk_init_bounds(SK(h))); }
int rbp_kmeans_step_local(rbp_kmeans_t* h) { return DISPATCH(h, w1_step_local(W1(h)), sk_step_local(SK(h))); }
int rbp_kmeans_accumulator(rbp_kmeans_t* h, void** p, size_t* b) {
    if (!p || !b) return RBP_ERR_INVALID;
    return DISPATCH(h, w1_accumulator(W1(h), p, b), sk_accumulator(SK(h), p, b));
}
int rbp_kmeans_counters(rbp_kmeans_t* h, void** s, void** r) { return DISPATCH(h, w1_counters(W1(h), s, r), sk_counters(SK(h), s, r)); }
void* rbp_kmeans_stream(rbp_kmeans_t* h) { return !h ? nullptr : (h->kind == RBP_KMEANS_W1 ? w1_stream(W1(h)) : sk_stream(SK(h))); }
int rbp_kmeans_step_finish(rbp_kmeans_t* h, float* d, uint32_t* s, uint32_t* r) { return DISPATCH(h, w1_step_finish(W1(h), d, s, r), sk_step_finish(SK(h), d, s, r)); }
// With a communicator the point pass is followed by ONE integer all-reduce on the layer's stream: the K x (bins + 1) u64 member sums with the K
// cluster sizes and the reassignment counter packed behind them (integer sums: exact in any order, so every rank forms bit-identical
// centroids — elkan.rs:125-142 over point shards).
__global__ void kmeans_pack_tallies_kernel(unsigned long long* __restrict__ tail, uint32_t* __restrict__ sizes, uint32_t* __restrict__ reassigned, int k, int unpack) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > k) return;
    uint32_t* w = i < k ? sizes + i : reassigned;
    if (unpack) *w = (uint32_t)tail[i]; else tail[i] = *w;
}
int rbp_kmeans_step(rbp_kmeans_t* h, float* d, uint32_t* s, uint32_t* r) {
    int st = rbp_kmeans_step_local(h);
    if (st) return st;
    if (h->comm && h->comm->world > 1) {
        void *acc = nullptr, *sizes = nullptr, *re = nullptr;
        size_t bytes = 0;
        if ((st = rbp_kmeans_accumulator(h, &acc, &bytes)) || (st = rbp_kmeans_counters(h, &sizes, &re))) return st;
        cudaStream_t stream = static_cast<cudaStream_t>(rbp_kmeans_stream(h));
        unsigned long long* tail = static_cast<unsigned long long*>(acc) + bytes / 8;
        kmeans_pack_tallies_kernel<<<(h->k + 128) / 128, 128, 0, stream>>>(tail, static_cast<uint32_t*>(sizes), static_cast<uint32_t*>(re), h->k, 0);
        RBP_LAUNCHED();
        if ((st = comm::all_reduce_sum_u64(h->comm, acc, bytes / 8 + (size_t)h->k + 1, stream))) return st;
        kmeans_pack_tallies_kernel<<<(h->k + 128) / 128, 128, 0, stream>>>(tail, static_cast<uint32_t*>(sizes), static_cast<uint32_t*>(re), h->k, 1);
        RBP_LAUNCHED();
    }
    return rbp_kmeans_step_finish(h, d, s, r);
}
int rbp_kmeans_attach_comm(rbp_kmeans_t* h, rbp_comm_t* c) {
    if (!h || !c) return RBP_ERR_INVALID;
    h->comm = c;
    return RBP_OK;
}
int rbp_kmeans_assign(rbp_kmeans_t* h, uint32_t* a, float* d) {
    if (!a) return RBP_ERR_INVALID;
    return DISPATCH(h, w1_assign(W1(h), a, d), sk_assign(SK(h), a, d));
}
int rbp_kmeans_centroids(rbp_kmeans_t* h, uint64_t* c, uint64_t* w) { return DISPATCH(h, w1_centroids(W1(h), c, w), sk_centroids(SK(h), c, w)); }
int rbp_kmeans_metric(rbp_kmeans_t* h, float* tri) {
    if (!tri) return RBP_ERR_INVALID;
    return DISPATCH(h, w1_metric(W1(h), tri), sk_metric(SK(h), tri));
}
int rbp_kmeans_bounds(rbp_kmeans_t* h, uint32_t* a, float* u, float* l, uint8_t* s) { return DISPATCH(h, w1_bounds(W1(h), a, u, l, s), sk_bounds(SK(h), a, u, l, s)); }
int rbp_kmeans_timed(rbp_kmeans_t* h, int what, int iters, float* ms) {
    if (!ms || iters < 1) return RBP_ERR_INVALID;
    return DISPATCH(h, w1_timed(W1(h), what, iters, ms), sk_timed(SK(h), what, iters, ms));
}
int rbp_kmeans_sinkhorn_stats(rbp_kmeans_t* h, uint64_t* out3, int reset) {
    if (!h) return RBP_ERR_INVALID;
    if (h->kind != RBP_KMEANS_SINKHORN) { set_last_error("only Sinkhorn layers count OT solves"); return RBP_ERR_STATE; }
    return sk_stats(SK(h), out3, reset);
}
int rbp_kmeans_screen(rbp_kmeans_t* h, float margin) {
    if (!h) return RBP_ERR_INVALID;
    if (h->kind != RBP_KMEANS_SINKHORN) { set_last_error("the screen belongs to Sinkhorn layers (the W1 layer's distance is already exact and cheap)"); return RBP_ERR_STATE; }
    return sk_screen(SK(h), margin);
}
int rbp_kmeans_screen_probe(rbp_kmeans_t* h, int64_t m, float* out, uint64_t* stats2) {
    if (!h) return RBP_ERR_INVALID;
    if (h->kind != RBP_KMEANS_SINKHORN) return RBP_ERR_STATE;
    return sk_screen_probe(SK(h), m, out, stats2);
}
int rbp_sinkhorn_batch(const uint32_t* a_counts, int na, const uint32_t* b_counts, int nb, int bins, const int32_t* ia, const int32_t* ib,
                       int64_t n, const float* tri, float temperature, int iterations, float tolerance, float* out) {
    return sk_batch(a_counts, na, b_counts, nb, bins, ia, ib, n, tri, temperature, iterations, tolerance, out);
}

}  // extern "C"
