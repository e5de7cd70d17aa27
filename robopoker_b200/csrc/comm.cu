// comm.cu — rbp_comm_*: communicator of one rank (one process per GPU) and the collectives the sharded paths use.
// Replaces nothing in the reference's fast path (rayon is single-process); it is the "multi-GPU via rbp_comm_init inside
// the library" row of SURVEY §8b, so that a Rust host needs no collective code of its own.
#include "comm.hpp"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

namespace rbp {
namespace comm {
namespace {

struct Api {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};
Api g_api;
std::once_flag g_once;

void load_once() {
    std::vector<std::string> names;
    if (const char* e = getenv("RBP_NCCL_LIB")) names.push_back(e);
    names.push_back("libnccl.so.2");  // already mapped when the host process imported torch; else LD_LIBRARY_PATH
    names.push_back("libnccl.so");
    for (const std::string& n : names) {
        g_api.handle = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (g_api.handle) break;
        g_api.why = dlerror();
    }
    if (!g_api.handle) return;
    auto sym = [&](const char* name) { void* p = dlsym(g_api.handle, name); if (!p) g_api.why = std::string("missing symbol ") + name; return p; };
    g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(sym("ncclCommInitRank"));
    g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(sym("ncclCommDestroy"));
    g_api.AllReduce = reinterpret_cast<decltype(g_api.AllReduce)>(sym("ncclAllReduce"));
    g_api.AllGather = reinterpret_cast<decltype(g_api.AllGather)>(sym("ncclAllGather"));
    g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(sym("ncclGetErrorString"));
}
int load() {
    std::call_once(g_once, load_once);
    if (!g_api.handle || !g_api.GetUniqueId || !g_api.CommInitRank || !g_api.CommDestroy || !g_api.AllReduce || !g_api.AllGather || !g_api.GetErrorString) {
        set_last_error("NCCL is not loadable (set RBP_NCCL_LIB or LD_LIBRARY_PATH to libnccl.so.2): " + g_api.why);
        return RBP_ERR_STATE;
    }
    return RBP_OK;
}
int nccl_fail(ncclResult_t r, const char* what) {
    set_last_error(std::string(what) + ": " + (g_api.GetErrorString ? g_api.GetErrorString(r) : "nccl error"));
    return RBP_ERR_CUDA;
}
#define RBP_NCCL(expr)                                             \
    do {                                                           \
        ncclResult_t _r = (expr);                                  \
        if (_r != ncclSuccess) return nccl_fail(_r, #expr);        \
    } while (0)

}  // namespace

int barrier(rbp_comm* c, cudaStream_t stream) {
    if (!c || c->world == 1) return RBP_OK;
    RBP_NCCL(g_api.AllReduce(c->token, c->token, 1, ncclInt32, ncclMax, c->nccl, stream));
    return RBP_OK;
}
int all_reduce_sum_u64(rbp_comm* c, void* buf, size_t count, cudaStream_t stream) {
    if (!c || c->world == 1 || count == 0) return RBP_OK;
    RBP_NCCL(g_api.AllReduce(buf, buf, count, ncclUint64, ncclSum, c->nccl, stream));
    return RBP_OK;
}
int all_gather_bytes(rbp_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream) {
    if (!c || c->world == 1) {
        if (send != recv) RBP_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, stream));
        return RBP_OK;
    }
    RBP_NCCL(g_api.AllGather(send, recv, bytes_per_rank, ncclUint8, c->nccl, stream));
    return RBP_OK;
}
int share(rbp_comm* c, void* local, void** peers, cudaStream_t stream) {
    if (!c || !local || !peers) return RBP_ERR_INVALID;
    for (int r = 0; r < c->world; ++r) peers[r] = nullptr;
    peers[c->rank] = local;
    if (c->world == 1) return RBP_OK;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t mine;
    RBP_CUDA(cudaIpcGetMemHandle(&mine, local));
    RBP_CUDA(cudaMemcpyAsync(c->stage + 64 * (size_t)c->rank, &mine, 64, cudaMemcpyHostToDevice, stream));
    RBP_NCCL(g_api.AllGather(c->stage + 64 * (size_t)c->rank, c->stage, 64, ncclUint8, c->nccl, stream));
    std::vector<cudaIpcMemHandle_t> all(c->world);
    RBP_CUDA(cudaMemcpyAsync(all.data(), c->stage, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, stream));
    RBP_CUDA(cudaStreamSynchronize(stream));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_last_error(std::string("cudaIpcOpenMemHandle of rank ") + std::to_string(r) + "'s buffer failed (peer access over NVLink is required): " + cudaGetErrorString(e));
            return RBP_ERR_CUDA;
        }
        peers[r] = p;  // the caller owns the mapping: comm::unshare before its own buffer is freed
    }
    // nobody may free / reuse before every rank has opened: a second collective closes the window
    return barrier(c, stream);
}

// close this process's mappings of the peers' buffers; nobody frees its own buffer before every rank has closed
int unshare(rbp_comm* c, void** peers, cudaStream_t stream) {
    if (!c || !peers) return RBP_ERR_INVALID;
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && peers[r]) { cudaIpcCloseMemHandle(peers[r]); peers[r] = nullptr; }
    const int rc = barrier(c, stream);
    if (rc != RBP_OK) return rc;
    RBP_CUDA(cudaStreamSynchronize(stream));
    return RBP_OK;
}

}  // namespace comm
}  // namespace rbp

using namespace rbp;

extern "C" {

int rbp_comm_unique_id(uint8_t out[128]) {
    if (!out) return RBP_ERR_INVALID;
    const int rc = comm::load();
    if (rc != RBP_OK) return rc;
    ncclUniqueId id;
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    const ncclResult_t r = comm::g_api.GetUniqueId(&id);
    if (r != ncclSuccess) return comm::nccl_fail(r, "ncclGetUniqueId");
    std::memcpy(out, &id, 128);
    return RBP_OK;
}
int rbp_comm_init(int world_rank, int world_size, const uint8_t id[128], int device, rbp_comm_t** out) {
    if (!out || !id || world_size < 1 || world_size > comm::kMaxWorld || world_rank < 0 || world_rank >= world_size) return RBP_ERR_INVALID;
    *out = nullptr;
    if (rbp_device_count() <= device) { set_last_error("no CUDA device"); return RBP_ERR_NO_DEVICE; }
    const int rc = comm::load();
    if (rc != RBP_OK) return rc;
    RBP_CUDA(cudaSetDevice(device));
    rbp_comm* c = new rbp_comm();
    c->rank = world_rank; c->world = world_size; c->device = device;
    ncclUniqueId uid;
    std::memcpy(&uid, id, 128);
    const ncclResult_t r = comm::g_api.CommInitRank(&c->nccl, world_size, uid, world_rank);
    if (r != ncclSuccess) { delete c; return comm::nccl_fail(r, "ncclCommInitRank"); }
    if (cudaMalloc(&c->token, sizeof(int)) != cudaSuccess || cudaMemset(c->token, 0, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&c->stage, 64 * (size_t)comm::kMaxWorld) != cudaSuccess) {
        rbp_comm_destroy(c);
        set_last_error("rbp_comm_init: device allocation failed");
        return RBP_ERR_CUDA;
    }
    *out = c;
    return RBP_OK;
}
void rbp_comm_destroy(rbp_comm_t* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->token) cudaFree(c->token);
    if (c->stage) cudaFree(c->stage);
    if (c->nccl) comm::g_api.CommDestroy(c->nccl);
    delete c;
}
int rbp_comm_rank(rbp_comm_t* c) { return c ? c->rank : -1; }
int rbp_comm_size(rbp_comm_t* c) { return c ? c->world : 0; }
int rbp_comm_barrier(rbp_comm_t* c) {
    if (!c) return RBP_ERR_INVALID;
    RBP_CUDA(cudaSetDevice(c->device));
    const int rc = comm::barrier(c, nullptr);
    if (rc != RBP_OK) return rc;
    RBP_CUDA(cudaStreamSynchronize(nullptr));
    return RBP_OK;
}

}  // extern "C"
