// The data-parallel exchange of the path (SURVEY.md 8(b) `flat_allreduce`, 8(e)): ONE NCCL all-reduce (sum) of the flat fp32
// gradient buffer, issued by this library on the caller's stream.  Replaces the gradient reduction that nn.DataParallel's
// backward performs in the reference (/root/reference/dpc/main.py:65,230) in the one-process-per-GPU layout.
//
// NCCL is resolved at run time from the libnccl.so.2 already loaded into the process (the one PyTorch links), so the library
// has no link-time dependency on it; the communicator is created from a unique id that the host side distributes with
// torch.distributed (plumbing), and the collective itself runs on the compute stream -- no hop through another library's
// internal stream.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};

const NcclApi* nccl() {
    static NcclApi api = [] {
        NcclApi a{};
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return a;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString;
        return a;
    }();
    return &api;
}

}  // namespace

#define DPC_NCCL(call)                                                                          \
    do {                                                                                        \
        ncclResult_t r__ = (call);                                                              \
        if (r__ != ncclSuccess) {                                                               \
            dpc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, nccl()->GetErrorString(r__)); \
            return DPC_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

extern "C" int dpc_comm_unique_id(void* id128) {
    DPC_REQUIRE(id128 != nullptr, "dpc_comm_unique_id: null pointer");
    DPC_REQUIRE(nccl()->ok, "dpc_comm_unique_id: libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == DPC_COMM_ID_BYTES, "ncclUniqueId size");
    DPC_NCCL(nccl()->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
    return DPC_OK;
}

extern "C" int dpc_comm_init(const void* id128, int rank, int world, void** comm) {
    DPC_REQUIRE(id128 && comm, "dpc_comm_init: null pointer");
    DPC_REQUIRE(world >= 1 && rank >= 0 && rank < world, "dpc_comm_init: bad rank %d of %d", rank, world);
    DPC_REQUIRE(nccl()->ok, "dpc_comm_init: libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    DPC_NCCL(nccl()->CommInitRank(&c, world, id, rank));
    *comm = c;
    return DPC_OK;
}

extern "C" int dpc_flat_allreduce(void* comm, float* buf, int64_t n, void* stream) {
    DPC_REQUIRE(comm && buf, "dpc_flat_allreduce: null pointer");
    DPC_REQUIRE(n > 0, "dpc_flat_allreduce: bad length %lld", (long long)n);
    DPC_NCCL(nccl()->AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, reinterpret_cast<ncclComm_t>(comm),
                               reinterpret_cast<cudaStream_t>(stream)));
    return DPC_OK;
}

extern "C" int dpc_comm_destroy(void* comm) {
    if (!comm) return DPC_OK;
    DPC_REQUIRE(nccl()->ok, "dpc_comm_destroy: libnccl.so.2 could not be loaded");
    DPC_NCCL(nccl()->CommDestroy(reinterpret_cast<ncclComm_t>(comm)));
    return DPC_OK;
}
