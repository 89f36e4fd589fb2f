// s2_nccl.cpp — the one cross-GPU step of the path: a reduce (sum) of the per-GPU buses into the master bus
// (SURVEY.md 8b "s2_bank_reduce_bus(comm, ...) wrapping the NCCL reduce", 8e).  The reference has no collective;
// the requirement is the north star's "one NCCL reduce over NVLink sums per-GPU buses into the master buffer".
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already mapped into the process if there is one,
// e.g. PyTorch's, else the system's), so the library itself links against the CUDA runtime only and a
// single-GPU host needs no NCCL at all.
#include "../../include/s2_cuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <new>

namespace s2 {
int set_error(int code, const char* fmt, ...);
}

namespace {

struct NcclUid { char internal[S2_COMM_UNIQUE_ID_BYTES]; };       // ncclUniqueId (nccl.h: 128 opaque bytes), passed by value
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUid*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUid, int);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*ReduceFn)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);

struct Nccl {
    void* handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    ReduceFn reduce = nullptr;
    GetErrorStringFn error_string = nullptr;
    GetVersionFn get_version = nullptr;
    bool ok = false;
};
Nccl g_nccl;
std::once_flag g_once;

const Nccl& nccl() {
    std::call_once(g_once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.handle) break;
        }
        if (!g_nccl.handle) return;
        g_nccl.get_unique_id = (GetUniqueIdFn)dlsym(g_nccl.handle, "ncclGetUniqueId");
        g_nccl.comm_init_rank = (CommInitRankFn)dlsym(g_nccl.handle, "ncclCommInitRank");
        g_nccl.comm_destroy = (CommDestroyFn)dlsym(g_nccl.handle, "ncclCommDestroy");
        g_nccl.reduce = (ReduceFn)dlsym(g_nccl.handle, "ncclReduce");
        g_nccl.error_string = (GetErrorStringFn)dlsym(g_nccl.handle, "ncclGetErrorString");
        g_nccl.get_version = (GetVersionFn)dlsym(g_nccl.handle, "ncclGetVersion");
        g_nccl.ok = g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.comm_destroy && g_nccl.reduce;
    });
    return g_nccl;
}

int need_nccl() {
    if (!nccl().ok) return s2::set_error(S2_ERR_NO_DEVICE, "NCCL is not available (dlopen libnccl.so.2: %s)", dlerror() ? dlerror() : "symbols missing");
    return S2_OK;
}

int nccl_fail(const char* what, int rc) {
    const Nccl& n = nccl();
    return s2::set_error(S2_ERR_CUDA, "%s failed: %s", what, n.error_string ? n.error_string(rc) : "NCCL error");
}

constexpr int kNcclFloat32 = 7, kNcclSum = 0;      // nccl.h: ncclFloat32, ncclSum

}  // namespace

struct s2_comm {
    NcclComm comm = nullptr;
    int n_ranks = 1, rank = 0, device = 0;
    bool owned = false;
};

extern "C" {

int s2_comm_version(int* version) {
    if (!version) return s2::set_error(S2_ERR_INVALID, "null version");
    int rc = need_nccl();
    if (rc) return rc;
    *version = 0;
    if (nccl().get_version) nccl().get_version(version);
    return S2_OK;
}

int s2_comm_unique_id(uint8_t* id) {
    if (!id) return s2::set_error(S2_ERR_INVALID, "null id");
    int rc = need_nccl();
    if (rc) return rc;
    NcclUid uid;
    memset(&uid, 0, sizeof uid);
    const int e = nccl().get_unique_id(&uid);
    if (e) return nccl_fail("ncclGetUniqueId", e);
    memcpy(id, uid.internal, sizeof uid.internal);
    return S2_OK;
}

int s2_comm_create(const uint8_t* id, int n_ranks, int rank, int device, s2_comm** out) {
    if (!out) return s2::set_error(S2_ERR_INVALID, "null out");
    *out = nullptr;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return s2::set_error(S2_ERR_INVALID, "bad communicator arguments");
    int rc = need_nccl();
    if (rc) return rc;
    if (cudaSetDevice(device) != cudaSuccess) return s2::set_error(S2_ERR_NO_DEVICE, "device %d is not usable", device);
    s2_comm* c = new (std::nothrow) s2_comm;
    if (!c) return s2::set_error(S2_ERR_NOMEM, "out of host memory");
    NcclUid uid;
    memcpy(uid.internal, id, sizeof uid.internal);
    const int e = nccl().comm_init_rank(&c->comm, n_ranks, uid, rank);
    if (e) { delete c; return nccl_fail("ncclCommInitRank", e); }
    c->n_ranks = n_ranks; c->rank = rank; c->device = device; c->owned = true;
    *out = c;
    return S2_OK;
}

int s2_comm_adopt(void* nccl_comm, int n_ranks, int rank, int device, s2_comm** out) {
    if (!out) return s2::set_error(S2_ERR_INVALID, "null out");
    *out = nullptr;
    if (!nccl_comm || n_ranks < 1 || rank < 0 || rank >= n_ranks) return s2::set_error(S2_ERR_INVALID, "bad communicator arguments");
    int rc = need_nccl();
    if (rc) return rc;
    s2_comm* c = new (std::nothrow) s2_comm;
    if (!c) return s2::set_error(S2_ERR_NOMEM, "out of host memory");
    c->comm = nccl_comm; c->n_ranks = n_ranks; c->rank = rank; c->device = device; c->owned = false;
    *out = c;
    return S2_OK;
}

void s2_comm_destroy(s2_comm* c) {
    if (!c) return;
    if (c->owned && c->comm && nccl().ok) {
        cudaSetDevice(c->device);
        nccl().comm_destroy(c->comm);
    }
    delete c;
}

int s2_bank_reduce_bus(s2_bank* bank, s2_comm* comm, int root, const float* d_bus_in, float* d_bus_out, size_t frames,
                       void* stream) {
    if (!bank || !comm || !d_bus_in) return s2::set_error(S2_ERR_INVALID, "null argument");
    if (root < 0 || root >= comm->n_ranks) return s2::set_error(S2_ERR_INVALID, "root %d out of range", root);
    if (comm->rank == root && !d_bus_out) return s2::set_error(S2_ERR_INVALID, "the root rank needs an output buffer");
    if (frames == 0) return S2_OK;
    // the bus is complete once every voice range of the bank and its mix stream have been joined
    int rc = s2_bank_join(bank, stream);
    if (rc) return rc;
    if (cudaSetDevice(comm->device) != cudaSuccess) return s2::set_error(S2_ERR_NO_DEVICE, "device %d is not usable", comm->device);
    const int e = nccl().reduce(d_bus_in, d_bus_out ? d_bus_out : const_cast<float*>(d_bus_in), frames, kNcclFloat32, kNcclSum,
                                root, comm->comm, (cudaStream_t)stream);
    if (e) return nccl_fail("ncclReduce", e);
    return S2_OK;
}

}  // extern "C"
