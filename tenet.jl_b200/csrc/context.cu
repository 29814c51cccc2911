// context.cu — context, error reporting, the stream-ordered caching allocator and host<->device copies.
//
// Replaces, for the hot path, the Julia `Array` storage behind `parent(tensor)` in the reference
// (/root/reference/src/Components/MPS.jl:83 etc.): tensors live in device buffers handed out by a
// size-class caching allocator.  All work of a context is ordered on ONE stream, so a block released by
// tnb_free may be handed out again immediately: any later kernel that writes it is ordered after every
// earlier kernel that read it.  tnb_free only takes the context mutex and is safe from finalizer threads.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "tnb_internal.h"

static thread_local std::string g_create_error;

int tnb_set_error(tnb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

extern "C" {

int tnb_ctx_create(int device, tnb_ctx** out) {
    if (!out) return tnb_set_error(nullptr, TNB_EINVAL, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return tnb_set_error(nullptr, TNB_ECUDA, "no CUDA device available (%s); libtnb200 has no CPU fallback",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return tnb_set_error(nullptr, TNB_EINVAL, "device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return tnb_set_error(nullptr, TNB_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return tnb_set_error(nullptr, TNB_EUNSUPPORTED, "device %d is sm_%d%d; libtnb200 is built for sm_100a (B200) only",
                             device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return tnb_set_error(nullptr, TNB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    tnb_ctx* ctx = new (std::nothrow) tnb_ctx();
    if (!ctx) return tnb_set_error(nullptr, TNB_ENOMEM, "host allocation failed");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("TNB_GEMM_PAIR")) ctx->gemm_pair = atoi(e) ? 1 : 0;   // default of TNB_OPT_GEMM_PAIR
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return tnb_set_error(nullptr, TNB_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return TNB_OK;
}

int tnb_comm_destroy(tnb_ctx* ctx);

int tnb_ctx_destroy(tnb_ctx* ctx) {
    if (!ctx) return TNB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    tnb_multi_release(ctx);
    tnb_comm_destroy(ctx);
    for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
    ctx->free_blocks.clear();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TNB_OK;
}

int tnb_ctx_set_option(tnb_ctx* ctx, int option, int64_t value) {
    if (!ctx) return TNB_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    switch (option) {
        case TNB_OPT_C64_MODE:
            if (value != TNB_C64_SIMT && value != TNB_C64_TF32X3 && value != TNB_C64_TF32X3_FAST) return tnb_set_error(ctx, TNB_EINVAL, "bad c64 mode");
            ctx->c64_mode = (int)value;
            return TNB_OK;
        case TNB_OPT_FORCE_KERNEL:
            ctx->force_generic = value ? 1 : 0;
            return TNB_OK;
        case TNB_OPT_CUDA_GRAPH:
            ctx->use_graphs = value ? 1 : 0;
            return TNB_OK;
        case TNB_OPT_GEMM_PAIR:
            ctx->gemm_pair = value ? 1 : 0;
            return TNB_OK;
    }
    return tnb_set_error(ctx, TNB_EINVAL, "unknown option %d", option);
}

const char* tnb_last_error(tnb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tnb_sync(tnb_ctx* ctx) {
    if (!ctx) return TNB_EINVAL;
    TNB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return TNB_OK;
}

void* tnb_ctx_stream(tnb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t tnb_ctx_launch_count(tnb_ctx* ctx) { return ctx ? ctx->launches : 0; }
int tnb_ctx_last_kernel(tnb_ctx* ctx) { return ctx ? ctx->last_kernel : -1; }

static size_t size_class(size_t bytes) {
    if (bytes < 512) return 512;
    if (bytes <= ((size_t)1 << 20)) {
        size_t c = 512;
        while (c < bytes) c <<= 1;
        return c;
    }
    const size_t G = (size_t)2 << 20;
    return (bytes + G - 1) / G * G;
}

int tnb_alloc(tnb_ctx* ctx, size_t bytes, tnb_buf** out) {
    if (!ctx || !out) return TNB_EINVAL;
    *out = nullptr;
    std::lock_guard<std::mutex> g(ctx->mu);
    const size_t cap = size_class(bytes);
    void* ptr = nullptr;
    size_t got = cap;
    auto it = ctx->free_blocks.lower_bound(cap);
    if (it != ctx->free_blocks.end() && it->first <= cap + cap / 4) {
        ptr = it->second;
        got = it->first;
        ctx->cached -= got;
        ctx->free_blocks.erase(it);
    } else {
        cudaSetDevice(ctx->device);
        cudaError_t e = cudaMalloc(&ptr, cap);
        if (e != cudaSuccess) {
            cudaGetLastError();
            // give cached blocks back to the driver and retry once
            cudaStreamSynchronize(ctx->stream);
            for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
            ctx->free_blocks.clear();
            ctx->cached = 0;
            e = cudaMalloc(&ptr, cap);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return tnb_set_error(ctx, TNB_ENOMEM, "cudaMalloc(%zu bytes) failed: %s (in use %zu)", cap,
                                     cudaGetErrorString(e), ctx->in_use);
            }
        }
    }
    tnb_buf* b = new (std::nothrow) tnb_buf();
    if (!b) {
        ctx->free_blocks.emplace(got, ptr);
        ctx->cached += got;
        return tnb_set_error(ctx, TNB_ENOMEM, "host allocation failed");
    }
    b->ptr = ptr;
    b->bytes = bytes;
    b->cap = got;
    ctx->in_use += got;
    if (ctx->in_use > ctx->peak) ctx->peak = ctx->in_use;
    *out = b;
    return TNB_OK;
}

int tnb_free(tnb_ctx* ctx, tnb_buf* buf) {
    if (!ctx) return TNB_EINVAL;
    if (!buf) return TNB_OK;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->free_blocks.emplace(buf->cap, buf->ptr);
    ctx->in_use -= buf->cap;
    ctx->cached += buf->cap;
    delete buf;
    return TNB_OK;
}

int tnb_mem_stats(tnb_ctx* ctx, size_t* in_use, size_t* cached, size_t* peak) {
    if (!ctx) return TNB_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (in_use) *in_use = ctx->in_use;
    if (cached) *cached = ctx->cached;
    if (peak) *peak = ctx->peak;
    return TNB_OK;
}

int tnb_mem_trim(tnb_ctx* ctx) {
    if (!ctx) return TNB_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    TNB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
    ctx->free_blocks.clear();
    ctx->cached = 0;
    return TNB_OK;
}

int tnb_upload(tnb_ctx* ctx, tnb_buf* dst, size_t off, const void* host, size_t bytes) {
    if (!ctx || !dst || (!host && bytes)) return tnb_set_error(ctx, TNB_EINVAL, "upload: NULL argument");
    if (off + bytes > dst->cap) return tnb_set_error(ctx, TNB_EINVAL, "upload: %zu+%zu bytes exceed buffer of %zu", off, bytes, dst->cap);
    if (!bytes) return TNB_OK;
    TNB_CUDA_CHECK(ctx, cudaMemcpyAsync((char*)dst->ptr + off, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return TNB_OK;
}

int tnb_download(tnb_ctx* ctx, const tnb_buf* src, size_t off, void* host, size_t bytes) {
    if (!ctx || !src || (!host && bytes)) return tnb_set_error(ctx, TNB_EINVAL, "download: NULL argument");
    if (off + bytes > src->cap) return tnb_set_error(ctx, TNB_EINVAL, "download: %zu+%zu bytes exceed buffer of %zu", off, bytes, src->cap);
    if (bytes) TNB_CUDA_CHECK(ctx, cudaMemcpyAsync(host, (const char*)src->ptr + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TNB_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return TNB_OK;
}

int tnb_memset_zero(tnb_ctx* ctx, tnb_buf* dst, size_t off, size_t bytes) {
    if (!ctx || !dst) return tnb_set_error(ctx, TNB_EINVAL, "memset: NULL argument");
    if (off + bytes > dst->cap) return tnb_set_error(ctx, TNB_EINVAL, "memset: range exceeds buffer");
    if (bytes) TNB_CUDA_CHECK(ctx, cudaMemsetAsync((char*)dst->ptr + off, 0, bytes, ctx->stream));
    return TNB_OK;
}

void* tnb_buf_ptr(const tnb_buf* buf) { return buf ? buf->ptr : nullptr; }
size_t tnb_buf_bytes(const tnb_buf* buf) { return buf ? buf->bytes : 0; }

}  // extern "C"
