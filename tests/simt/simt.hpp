// TEST INFRASTRUCTURE ONLY — a minimal SIMT shim that runs the library's CUDA kernel SOURCE on the CPU, one block at a time.
//
// Every CUDA thread of a block is a user-level fiber (ucontext) on one OS thread; `__syncthreads()` and the warp collectives are
// barriers between fibers, shared memory is static storage, "device" pointers are host pointers.  It exists so that kernel code written
// when no GPU was available (the subgame hooks of csrc/mccfr.cu) can be checked against the oracle before its first hardware run, and it
// doubles as a CPU check of the flat-game MCCFR kernels.  It says nothing about performance, memory models or divergence hazards: a
// kernel that is correct only by accident of warp scheduling passes here and fails on the device, and the other way round.
// Nothing under robopoker_b200/ includes this file.
#pragma once
#include <cuda_runtime.h>   // vector types, host-side no-op definitions of __device__ / __global__ / __shared__ ...
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __launch_bounds__(...)
using std::max;
using std::min;

namespace simt {
constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 256 * 1024;
struct Fiber { ucontext_t ctx; bool done = true; };
struct Block {
    int n = 0, cur = 0, live = 0;
    ucontext_t main;
    Fiber f[kMaxThreads];
    std::vector<char> stacks;
    unsigned long long block_gen = 0; int block_arrived = 0;
    unsigned long long warp_gen[kMaxThreads / 32] = {}; int warp_arrived[kMaxThreads / 32] = {}; int warp_live[kMaxThreads / 32] = {};
    unsigned long long slot[kMaxThreads] = {};
    std::function<void()> body;
};
inline Block& B() { static Block b; return b; }
}  // namespace simt

// the CUDA built-ins, as plain globals switched by the fiber scheduler
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
alignas(16) inline unsigned char smem_raw[1 << 20];

namespace simt {
inline void yield() { Block& b = B(); swapcontext(&b.f[b.cur].ctx, &b.main); }
inline void release_checks(Block& b) {  // a finished fiber counts as arrived
    if (b.live > 0 && b.block_arrived == b.live) { b.block_arrived = 0; ++b.block_gen; }
    for (int w = 0; w < (b.n + 31) / 32; ++w)
        if (b.warp_live[w] > 0 && b.warp_arrived[w] == b.warp_live[w]) { b.warp_arrived[w] = 0; ++b.warp_gen[w]; }
}
inline void trampoline() {
    Block& b = B();
    b.body();
    b.f[b.cur].done = true;
    --b.live; --b.warp_live[b.cur / 32];
    release_checks(b);
    swapcontext(&b.f[b.cur].ctx, &b.main);
}
inline void block_barrier() {
    Block& b = B();
    const unsigned long long g = b.block_gen;
    ++b.block_arrived;
    release_checks(b);
    while (b.block_gen == g) yield();
}
inline void warp_barrier() {
    Block& b = B();
    const int w = b.cur / 32;
    const unsigned long long g = b.warp_gen[w];
    ++b.warp_arrived[w];
    release_checks(b);
    while (b.warp_gen[w] == g) yield();
}
template <class F>
void launch(unsigned grid, unsigned block, F&& kernel_call) {
    Block& b = B();
    if (block > (unsigned)kMaxThreads) { fprintf(stderr, "simt: block too large\n"); abort(); }
    if (b.stacks.size() < kStack * block) b.stacks.resize(kStack * block);
    gridDim = dim3(grid); blockDim = dim3(block);
    b.body = kernel_call;
    for (unsigned blk = 0; blk < grid; ++blk) {
        b.n = (int)block; b.live = (int)block; b.block_arrived = 0;
        for (int w = 0; w < kMaxThreads / 32; ++w) { b.warp_arrived[w] = 0; b.warp_live[w] = 0; }
        for (unsigned t = 0; t < block; ++t) {
            getcontext(&b.f[t].ctx);
            b.f[t].ctx.uc_stack.ss_sp = b.stacks.data() + kStack * t;
            b.f[t].ctx.uc_stack.ss_size = kStack;
            b.f[t].ctx.uc_link = &b.main;
            makecontext(&b.f[t].ctx, trampoline, 0);
            b.f[t].done = false;
            ++b.warp_live[t / 32];
        }
        while (b.live > 0) {
            bool progressed = false;
            for (unsigned t = 0; t < block; ++t) {
                if (b.f[t].done) continue;
                b.cur = (int)t;
                blockIdx = uint3{blk, 0, 0}; threadIdx = uint3{t, 0, 0};
                swapcontext(&b.main, &b.f[t].ctx);
                progressed = true;
            }
            if (!progressed) break;
        }
    }
}
inline unsigned long long exchange(unsigned long long v, int src_lane) {  // every live lane deposits, then reads lane `src_lane` of its warp
    Block& b = B();
    const int base = b.cur / 32 * 32;
    b.slot[b.cur] = v;
    warp_barrier();
    const unsigned long long r = b.slot[base + (src_lane & 31)];
    warp_barrier();
    return r;
}
template <class T> unsigned long long bits_of(T v) { unsigned long long u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> T from_bits(unsigned long long u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
}  // namespace simt

inline void __syncthreads() { simt::block_barrier(); }
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { simt::warp_barrier(); }
template <class T> T __shfl_sync(unsigned, T v, int src, int = 32) { return simt::from_bits<T>(simt::exchange(simt::bits_of(v), src)); }
template <class T> T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    const int lane = (int)(threadIdx.x & 31);
    const T r = simt::from_bits<T>(simt::exchange(simt::bits_of(v), lane >= (int)d ? lane - (int)d : lane));
    return r;
}
template <class T> T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    const int lane = (int)(threadIdx.x & 31);
    return simt::from_bits<T>(simt::exchange(simt::bits_of(v), lane + (int)d < 32 ? lane + (int)d : lane));
}
template <class T> T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return simt::from_bits<T>(simt::exchange(simt::bits_of(v), (int)(threadIdx.x & 31) ^ m)); }
inline unsigned __ballot_sync(unsigned, int pred) {
    simt::Block& b = simt::B();
    const int base = b.cur / 32 * 32;
    b.slot[b.cur] = pred ? 1ull : 0ull;
    simt::warp_barrier();
    unsigned m = 0;
    for (int l = 0; l < 32 && base + l < b.n; ++l) if (!b.f[base + l].done && b.slot[base + l]) m |= 1u << l;
    simt::warp_barrier();
    return m;
}
inline int __all_sync(unsigned mask, int pred) {
    simt::Block& b = simt::B();
    const int base = b.cur / 32 * 32;
    b.slot[b.cur] = pred ? 1ull : 0ull;
    simt::warp_barrier();
    int all = 1;
    for (int l = 0; l < 32 && base + l < b.n; ++l) if (!b.f[base + l].done && !b.slot[base + l]) all = 0;
    simt::warp_barrier();
    (void)mask;
    return all;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline float __frcp_rn(float x) { return 1.0f / x; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
// fibers never run concurrently: plain read-modify-write is atomic here
template <class T, class U> T atomicAdd(T* p, U v) { const T old = *p; *p = (T)(old + (T)v); return old; }
template <class T, class U> T atomicOr(T* p, U v) { const T old = *p; *p = (T)(old | (T)v); return old; }
