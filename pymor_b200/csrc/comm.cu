// NCCL behind the C ABI: the k x k / k partial results of the sharded reductions (vectorarrays/mpi.py:394-416:
// Gather + host sum in the reference) are summed over the GPUs of the box by ncclAllReduce on the CONTEXT's stream,
// right behind the reduction kernel that produced them -- no stream hand-over, no Python in the data path.
//
// libnccl is not a link-time dependency: its entry points are looked up in the process at run time (the copy the
// host application already loaded -- PyTorch ships one -- else the system's libnccl.so.2), so that the library and
// its host never drive two different NCCL builds.  The 128-byte ncclUniqueId is created by one rank
// (pmc_comm_unique_id) and handed to the others by the HOST (any channel: torch.distributed store, MPI, a file).
#include <dlfcn.h>

#include "common.cuh"

namespace {

// the few NCCL declarations needed (stable across NCCL 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
    bool ok;
};

NcclApi *nccl_api() {
    static NcclApi api = {};
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy the host process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
    return api.ok ? &api : nullptr;
}

#define PMC_CHECK_NCCL(api, expr)                                                                    \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != 0) {                                                                               \
            pmc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r));  \
            return PMC_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

}  // namespace

extern "C" int pmc_comm_unique_id(void *id_out) {
    PMC_REQUIRE(id_out != nullptr, "id_out is NULL");
    NcclApi *api = nccl_api();
    if (!api) {
        pmc_set_error("libnccl.so.2 not found in the process or on the library path");
        return PMC_ERR_UNSUPPORTED;
    }
    ncclUniqueId id;
    PMC_CHECK_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return PMC_OK;
}

extern "C" int pmc_comm_init(pmc_ctx *ctx, int rank, int size, const void *id_bytes) {
    PMC_REQUIRE(ctx && id_bytes && size >= 1 && rank >= 0 && rank < size, "bad arguments");
    PMC_REQUIRE(ctx->nccl_comm == nullptr, "the context has a communicator already");
    NcclApi *api = nccl_api();
    if (!api) {
        pmc_set_error("libnccl.so.2 not found in the process or on the library path");
        return PMC_ERR_UNSUPPORTED;
    }
    PMC_CHECK_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    PMC_CHECK_NCCL(api, api->CommInitRank(&comm, size, id, rank));
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_size = size;
    return PMC_OK;
}

extern "C" int pmc_comm_destroy(pmc_ctx *ctx) {
    if (!ctx || !ctx->nccl_comm) return PMC_OK;
    NcclApi *api = nccl_api();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (api) api->CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
    ctx->comm_size = 0;
    return PMC_OK;
}

extern "C" int pmc_comm_size(pmc_ctx *ctx) { return (ctx && ctx->nccl_comm) ? ctx->comm_size : 0; }

extern "C" int pmc_allreduce_sum(pmc_ctx *ctx, double *buf, int64_t count, uint32_t flags) {
    PMC_REQUIRE(ctx && count >= 0, "bad arguments");
    PMC_REQUIRE(ctx->nccl_comm != nullptr, "no communicator: call pmc_comm_init first");
    if (count == 0) return pmc_after_launch(ctx, 0, flags);
    PMC_REQUIRE(buf != nullptr, "buf is NULL");
    NcclApi *api = nccl_api();
    PMC_REQUIRE(api != nullptr, "NCCL is gone");
    PMC_CHECK_NCCL(api, api->AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum,
                                       static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
    return pmc_after_launch(ctx, 0, flags);          // (NCCL's kernel is not one of ours: not counted as a launch)
}
