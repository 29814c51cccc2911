// comm.cu — the single collective of the path: sum the per-GPU slice accumulators.
//
// The reference has no distributed code (README.md:22 only advertises "Distributed contraction"), so there
// is nothing to mirror; the shape is dictated by slicing: independent slices are dealt round-robin to one
// process per GPU, each adds its slices into a local accumulator, ONE ncclAllReduce(sum) over NVLink 5 /
// NVSwitch ends the job.  The message is the output tensor (an amplitude: 8 or 16 bytes), so it is latency
// bound and deliberately not fused with the last GEMM.  NCCL is dlopen'ed so that libtnb200.so loads on
// hosts without it (and in CPU-only CI); a missing libnccl is TNB_ENCCL at tnb_comm_* time, never a fallback.
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "tnb_internal.h"

typedef struct { char internal[128]; } ncclUniqueId_t;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
    int (*CommInitRank)(void**, int, ncclUniqueId_t, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi* load_nccl(std::string* why) {
    static NcclApi api;
    static bool tried = false;
    static std::string err;
    if (!tried) {
        tried = true;
        const char* env = getenv("TNB_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
            err = dlerror();
        }
        if (api.handle) {
            api.GetUniqueId = (int (*)(ncclUniqueId_t*))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void**, int, ncclUniqueId_t, int))dlsym(api.handle, "ncclCommInitRank");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
            api.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclBroadcast");
            api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
            api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
            if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
                err = "libnccl lacks a required symbol";
                dlclose(api.handle);
                api.handle = nullptr;
            }
        }
    }
    if (!api.handle) { if (why) *why = err.empty() ? "libnccl.so.2 not found" : err; return nullptr; }
    return &api;
}

// in-place broadcast of raw bytes from `root` (tnb_multi_contract_path ships the leaves to the other devices with it)
int tnb_comm_broadcast_bytes(tnb_ctx* ctx, void* ptr, size_t bytes, int root) {
    if (!ctx || !ctx->comm) return tnb_set_error(ctx, TNB_ENCCL, "communicator not initialised");
    if (!ctx->nccl->Broadcast) return tnb_set_error(ctx, TNB_ENCCL, "libnccl lacks ncclBroadcast");
    if (!bytes) return TNB_OK;
    cudaSetDevice(ctx->device);
    int r = ctx->nccl->Broadcast(ptr, ptr, bytes, 0 /* ncclInt8 */, root, ctx->comm, ctx->stream);
    if (r) return tnb_set_error(ctx, TNB_ENCCL, "ncclBroadcast: %s", ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "error");
    return TNB_OK;
}

extern "C" {

int tnb_comm_unique_id(void* id128) {
    if (!id128) return TNB_EINVAL;
    std::string why;
    NcclApi* api = load_nccl(&why);
    if (!api) return tnb_set_error(nullptr, TNB_ENCCL, "cannot load NCCL: %s", why.c_str());
    ncclUniqueId_t id;
    int r = api->GetUniqueId(&id);
    if (r) return tnb_set_error(nullptr, TNB_ENCCL, "ncclGetUniqueId: %s", api->GetErrorString ? api->GetErrorString(r) : "error");
    memcpy(id128, &id, sizeof id);
    return TNB_OK;
}

int tnb_comm_init(tnb_ctx* ctx, const void* id128, int32_t rank, int32_t nranks) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return tnb_set_error(ctx, TNB_EINVAL, "comm_init: bad arguments");
    std::string why;
    NcclApi* api = load_nccl(&why);
    if (!api) return tnb_set_error(ctx, TNB_ENCCL, "cannot load NCCL: %s", why.c_str());
    if (ctx->comm) return tnb_set_error(ctx, TNB_EINVAL, "communicator already initialised");
    cudaSetDevice(ctx->device);
    ncclUniqueId_t id;
    memcpy(&id, id128, sizeof id);
    void* comm = nullptr;
    int r = api->CommInitRank(&comm, nranks, id, rank);
    if (r) return tnb_set_error(ctx, TNB_ENCCL, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(r) : "error");
    ctx->nccl = api;
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return TNB_OK;
}

int tnb_comm_allreduce_sum(tnb_ctx* ctx, tnb_buf* buf, size_t offset_bytes, int64_t count, int32_t dtype) {
    if (!ctx || !buf || count < 0) return tnb_set_error(ctx, TNB_EINVAL, "allreduce: bad arguments");
    const size_t esz = tnb_dtype_size(dtype);
    if (!esz) return tnb_set_error(ctx, TNB_EUNSUPPORTED, "allreduce: bad dtype");
    if (offset_bytes + (size_t)count * esz > buf->cap) return tnb_set_error(ctx, TNB_EINVAL, "allreduce: range exceeds buffer");
    if (ctx->nranks == 1 && !ctx->comm) return TNB_OK;
    if (!ctx->comm) return tnb_set_error(ctx, TNB_ENCCL, "communicator not initialised");
    cudaSetDevice(ctx->device);
    // complex = pairs of reals; ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0
    const bool dbl = dtype == TNB_C128 || dtype == TNB_F64;
    const size_t n = (size_t)count * (tnb_dtype_complex(dtype) ? 2 : 1);
    void* p = (char*)buf->ptr + offset_bytes;
    int r = ctx->nccl->AllReduce(p, p, n, dbl ? 8 : 7, 0, ctx->comm, ctx->stream);
    if (r) return tnb_set_error(ctx, TNB_ENCCL, "ncclAllReduce: %s", ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "error");
    return TNB_OK;
}

// ranks of the communicator bound to this context (1 when none): lets a caller whose launcher says world > 1 refuse to
// return a PARTIAL slice sum when tnb_comm_init was never called on this context
int32_t tnb_comm_size(const tnb_ctx* ctx) { return (ctx && ctx->comm) ? ctx->nranks : 1; }

int tnb_comm_destroy(tnb_ctx* ctx) {
    if (!ctx) return TNB_EINVAL;
    if (ctx->comm && ctx->nccl) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        ctx->nccl->CommDestroy(ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    return TNB_OK;
}

}  // extern "C"
