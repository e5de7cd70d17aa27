// comm.hpp — the one exchange step of each sharded path, inside the library (SURVEY §8b last row, §8e).
//
// One process per GPU.  NCCL (loaded with dlopen: the library has no link-time dependency on it, and loads without it
// when no communicator is ever created) provides the bootstrap, the barrier and the plain collectives (integer
// all-reduce of the k-means accumulators, all-gather of the small-game partial sums).  The NLHE record / row exchange
// does not go through NCCL: every rank maps its peers' receive buffers (CUDA IPC over NVLink / NVSwitch) and the
// partition kernel stores each record straight into its owner's memory — compute and transfer in one kernel, no
// host-side counts, no staging copy.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <vector>

#include "common.cuh"

struct rbp_comm {
    ncclComm_t nccl = nullptr;
    int rank = 0, world = 1, device = 0;
    int* token = nullptr;            // device: payload of the barrier
    unsigned char* stage = nullptr;  // device: world x 64 B staging for the IPC handle exchange
};

namespace rbp {
namespace comm {

constexpr int kMaxWorld = 16;

// stream-ordered: returns after enqueueing; every rank's stream passes the barrier only when all have reached it
int barrier(rbp_comm* c, cudaStream_t stream);
int all_reduce_sum_u64(rbp_comm* c, void* buf, size_t count, cudaStream_t stream);
int all_gather_bytes(rbp_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream);
// peers[r] = this process's mapping of rank r's `local` allocation (peers[rank] = local).  `local` must be the base of a
// cudaMalloc allocation.  Collective: every rank calls it with its own buffer, in the same order.
int share(rbp_comm* c, void* local, void** peers, cudaStream_t stream);
int unshare(rbp_comm* c, void** peers, cudaStream_t stream);  // collective: closes the peer mappings, then a barrier

}  // namespace comm
}  // namespace rbp
