// Final gather of the time-sharded outputs: one ncclAllGather, called from inside libtemd so that a
// caller of the C ABI has the whole multi-GPU path without torch.distributed.  The reference is single-process
// NumPy and has no counterpart; the independence that makes the record shard by time is sph_zonal_mean.py:244-251.
//
// NCCL is not linked: libnccl.so.2 is resolved at run time (the copy the process already loaded - e.g. the one
// bundled with PyTorch - else TEMD_NCCL_LIB, else the system library), so libtemd.so loads on machines without it.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/temd.h"
#include "temd_internal.h"

namespace temd {

struct NcclUniqueId { char internal[128]; };          // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
enum { NCCL_FLOAT64 = 8 };                            // ncclDouble / ncclFloat64

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*);
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
    int (*CommDestroy)(NcclComm);
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
    const char* (*GetErrorString)(int);
    bool ok;
};

static NcclApi* nccl_api() {
    static NcclApi api = {};
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        const char* env = getenv("TEMD_NCCL_LIB");
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // already in the process (PyTorch's copy)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
    });
    return api.ok ? &api : nullptr;
}

static int nccl_fail(const NcclApi* a, int rc, const char* what) {
    return temd_set_error(-7, "%s failed: NCCL error %d (%s)", what, rc, a->GetErrorString(rc));
}

}  // namespace temd

using namespace temd;

struct temd_comm {
    NcclComm comm;
    int dev, nranks, rank;
};

extern "C" int temd_comm_unique_id(char* id128_host) {
    if (id128_host == nullptr) return temd_set_error(-1, "comm_unique_id: null argument");
    NcclApi* a = nccl_api();
    if (!a) return temd_set_error(-7, "comm_unique_id: libnccl.so.2 not found (set TEMD_NCCL_LIB)");
    NcclUniqueId id;
    const int rc = a->GetUniqueId(&id);
    if (rc) return nccl_fail(a, rc, "ncclGetUniqueId");
    memcpy(id128_host, id.internal, sizeof(id.internal));
    return 0;
}

extern "C" int temd_comm_init(int device, int nranks, int rank, const char* id128_host, temd_comm** out) {
    if (out == nullptr || id128_host == nullptr || nranks < 1 || rank < 0 || rank >= nranks)
        return temd_set_error(-1, "comm_init: bad arguments (nranks %d, rank %d)", nranks, rank);
    *out = nullptr;
    NcclApi* a = nccl_api();
    if (!a) return temd_set_error(-7, "comm_init: libnccl.so.2 not found (set TEMD_NCCL_LIB)");
    int prev = -1;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess) e = cudaSetDevice(device);
    if (e != cudaSuccess) return temd_set_error((int)e, "comm_init: %s", cudaGetErrorString(e));
    NcclUniqueId id;
    memcpy(id.internal, id128_host, sizeof(id.internal));
    NcclComm c = nullptr;
    const int rc = a->CommInitRank(&c, nranks, id, rank);
    cudaSetDevice(prev);
    if (rc) return nccl_fail(a, rc, "ncclCommInitRank");
    temd_comm* t = new temd_comm{c, device, nranks, rank};
    *out = t;
    return 0;
}

extern "C" int temd_comm_destroy(temd_comm* c) {
    if (c == nullptr) return 0;
    NcclApi* a = nccl_api();
    if (a && c->comm) a->CommDestroy(c->comm);
    delete c;
    return 0;
}

// recv[r][0..count) = rank r's send[0..count)  (float64, device pointers, enqueued on `stream`)
extern "C" int temd_allgather_outputs(temd_comm* c, const double* send, double* recv, size_t count, void* stream) {
    if (c == nullptr || send == nullptr || recv == nullptr) return temd_set_error(-1, "allgather_outputs: null argument");
    NcclApi* a = nccl_api();
    if (!a) return temd_set_error(-7, "allgather_outputs: libnccl.so.2 not found");
    if (count == 0) return 0;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != c->dev) cudaSetDevice(c->dev);
    const int rc = a->AllGather(send, recv, count, NCCL_FLOAT64, c->comm, reinterpret_cast<cudaStream_t>(stream));
    if (prev != c->dev) cudaSetDevice(prev);
    if (rc) return nccl_fail(a, rc, "ncclAllGather");
    return 0;
}
