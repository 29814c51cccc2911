// linalg.cu — device-resident thin QR / SVD of a tensor viewed as a matrix.
//
// What it replaces: Muscle.tensor_qr_thin / tensor_svd_thin behind canonize!, compress!, evolve! and two-site DMRG
// (/root/reference/src/Operations/canonize.jl:41,58,97; evolve.jl:62,92; src/Algorithms/DMRG.jl:338,437).  In the
// reference these are LAPACK calls on host arrays BETWEEN the einsums of the hot path; with the tensors living in device
// buffers a host factorisation would cost a download + upload per site.  The factorisation itself is library work
// (cuSOLVER geqrf / orgqr / gesvd — SURVEY §8f accepts that; it is dlopen'ed, so libtnb200.so still loads without it and a
// missing library is TNB_EUNSUPPORTED at call time, never a CPU fallback); what this file adds is the tensor <-> matrix
// plumbing on the device: the strided / conjugated operand is gathered into a dense column-major matrix and the factors
// are scattered into the caller's layouts by the engine's own einsum kernel (outer product with the scalar 1).
//
//   A[rows.., cols..]  =  sum_k  Q[rows.., k] R[k, cols..]                     (k = min(m, n), R upper triangular)
//   A[rows.., cols..]  =  sum_k  U[rows.., k] S[k] Vh[k, cols..]               (S real, descending)
#include <cusolverDn.h>
#include <dlfcn.h>
#include <library_types.h>
#include <algorithm>
#include <cstring>
#include <set>
#include <string>
#include <vector>
#include "tnb_internal.h"

namespace {

struct SolverApi {
    void* lib = nullptr;
    cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
    cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
    cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*CreateParams)(cusolverDnParams_t*) = nullptr;
    cusolverStatus_t (*XgeqrfBuf)(cusolverDnHandle_t, cusolverDnParams_t, int64_t, int64_t, cudaDataType, const void*, int64_t, cudaDataType,
                                  const void*, cudaDataType, size_t*, size_t*) = nullptr;
    cusolverStatus_t (*Xgeqrf)(cusolverDnHandle_t, cusolverDnParams_t, int64_t, int64_t, cudaDataType, void*, int64_t, cudaDataType, void*,
                               cudaDataType, void*, size_t, void*, size_t, int*) = nullptr;
    cusolverStatus_t (*XgesvdBuf)(cusolverDnHandle_t, cusolverDnParams_t, signed char, signed char, int64_t, int64_t, cudaDataType,
                                  const void*, int64_t, cudaDataType, const void*, cudaDataType, const void*, int64_t, cudaDataType,
                                  const void*, int64_t, cudaDataType, size_t*, size_t*) = nullptr;
    cusolverStatus_t (*Xgesvd)(cusolverDnHandle_t, cusolverDnParams_t, signed char, signed char, int64_t, int64_t, cudaDataType, void*,
                               int64_t, cudaDataType, void*, cudaDataType, void*, int64_t, cudaDataType, void*, int64_t, cudaDataType,
                               void*, size_t, void*, size_t, int*) = nullptr;
    // Q formation has no generic entry point: typed orgqr / ungqr, called through void* (identical signatures up to the element type)
    cusolverStatus_t (*gqrBuf[4])(cusolverDnHandle_t, int, int, int, const void*, int, const void*, int*) = {};
    cusolverStatus_t (*gqr[4])(cusolverDnHandle_t, int, int, int, void*, int, const void*, void*, int, int*) = {};
    cusolverDnHandle_t handle = nullptr;      // one per process is enough: the stream is set per call under the context mutex
    cusolverDnParams_t params = nullptr;
};

SolverApi* load_solver(std::string* why) {
    static SolverApi api;
    static bool tried = false;
    static std::string err;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (!tried) {
        tried = true;
        const char* env = getenv("TNB_CUSOLVER_LIB");
        const char* names[] = {env, "libcusolver.so.11", "libcusolver.so.12", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
            err = dlerror();
        }
        if (api.lib) {
#define TNB_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name)
            TNB_SYM(Create, "cusolverDnCreate"); TNB_SYM(Destroy, "cusolverDnDestroy"); TNB_SYM(SetStream, "cusolverDnSetStream");
            TNB_SYM(CreateParams, "cusolverDnCreateParams");
            TNB_SYM(XgeqrfBuf, "cusolverDnXgeqrf_bufferSize"); TNB_SYM(Xgeqrf, "cusolverDnXgeqrf");
            TNB_SYM(XgesvdBuf, "cusolverDnXgesvd_bufferSize"); TNB_SYM(Xgesvd, "cusolverDnXgesvd");
            const char* gq[4] = {"cusolverDnZungqr", "cusolverDnCungqr", "cusolverDnDorgqr", "cusolverDnSorgqr"};   // TNB_C128, C64, F64, F32
            for (int i = 0; i < 4; i++) {
                *(void**)(&api.gqr[i]) = dlsym(api.lib, gq[i]);
                *(void**)(&api.gqrBuf[i]) = dlsym(api.lib, (std::string(gq[i]) + "_bufferSize").c_str());
            }
#undef TNB_SYM
            bool ok = api.Create && api.SetStream && api.CreateParams && api.XgeqrfBuf && api.Xgeqrf && api.XgesvdBuf && api.Xgesvd;
            for (int i = 0; i < 4; i++) ok = ok && api.gqr[i] && api.gqrBuf[i];
            if (ok) ok = api.Create(&api.handle) == CUSOLVER_STATUS_SUCCESS && api.CreateParams(&api.params) == CUSOLVER_STATUS_SUCCESS;
            if (!ok) { err = "libcusolver lacks a required symbol or cusolverDnCreate failed"; api.lib = nullptr; }
        }
    }
    if (!api.lib) { if (why) *why = err.empty() ? "libcusolver.so.11 not found" : err; return nullptr; }
    return &api;
}

std::mutex& solver_mutex() { static std::mutex m; return m; }

cudaDataType cuda_type(int dtype) {
    switch (dtype) {
        case TNB_C128: return CUDA_C_64F;
        case TNB_C64: return CUDA_C_32F;
        case TNB_F64: return CUDA_R_64F;
        default: return CUDA_R_32F;
    }
}
cudaDataType cuda_real_type(int dtype) { return (dtype == TNB_C128 || dtype == TNB_F64) ? CUDA_R_64F : CUDA_R_32F; }
int real_dtype(int dtype) { return (dtype == TNB_C128 || dtype == TNB_F64) ? TNB_F64 : TNB_F32; }

// zero the strictly lower triangle of the leading k x n block of a column-major matrix (R of geqrf)
template <typename T>
__global__ void upper_kernel(const T* __restrict__ W, int64_t ldw, T* __restrict__ R, int64_t k, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < k * n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i % k, c = i / k;
        R[i] = r <= c ? W[r + c * ldw] : T{};
    }
}
// real singular values -> the tensor dtype (S is returned in the operand's dtype so that binary_einsum(s, V; dims=[]) works)
template <typename R, typename T>
__global__ void widen_kernel(const R* __restrict__ s, T* __restrict__ out, int64_t k) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
        T v{};
        v.x = s[i];
        out[i] = v;
    }
}

struct Split {
    std::vector<int32_t> row_modes, col_modes;
    std::vector<int64_t> row_ext, col_ext;
    int64_t m = 1, n = 1;
};

int split_modes(tnb_ctx* ctx, const tnb_tensor* A, const int32_t* row_modes, int32_t nrow, Split* sp) {
    std::set<int32_t> rows(row_modes, row_modes + nrow);
    if ((int)rows.size() != nrow) return tnb_set_error(ctx, TNB_EINVAL, "factorisation: repeated row mode");
    for (int i = 0; i < nrow; i++) {
        int r = 0;
        for (; r < A->rank; r++) if (A->mode[r] == row_modes[i]) break;
        if (r == A->rank) return tnb_set_error(ctx, TNB_EINVAL, "factorisation: row mode %d is not a mode of A", row_modes[i]);
        sp->row_modes.push_back(row_modes[i]); sp->row_ext.push_back(A->extent[r]); sp->m *= A->extent[r];
    }
    for (int r = 0; r < A->rank; r++)
        if (!rows.count(A->mode[r])) { sp->col_modes.push_back(A->mode[r]); sp->col_ext.push_back(A->extent[r]); sp->n *= A->extent[r]; }
    return TNB_OK;
}

// dense descriptor [modes...] with column-major strides over `buf`
struct Dense {
    std::vector<int64_t> ext, stride;
    std::vector<int32_t> mode;
    tnb_tensor t;
    Dense(tnb_buf* buf, int dtype, const std::vector<int32_t>& modes, const std::vector<int64_t>& exts, int conj = 0) : ext(exts), mode(modes) {
        int64_t s = 1;
        for (size_t i = 0; i < ext.size(); i++) { stride.push_back(s); s *= ext[i]; }
        t.buf = buf; t.offset_elems = 0; t.dtype = dtype; t.rank = (int32_t)ext.size();
        t.extent = ext.data(); t.stride_elems = stride.data(); t.mode = mode.data(); t.conj = conj;
    }
};

std::vector<int32_t> cat(const std::vector<int32_t>& a, const std::vector<int32_t>& b) { std::vector<int32_t> r(a); r.insert(r.end(), b.begin(), b.end()); return r; }
std::vector<int64_t> cat(const std::vector<int64_t>& a, const std::vector<int64_t>& b) { std::vector<int64_t> r(a); r.insert(r.end(), b.begin(), b.end()); return r; }

struct Scratch {
    tnb_ctx* ctx;
    std::vector<tnb_buf*> bufs;
    explicit Scratch(tnb_ctx* c) : ctx(c) {}
    ~Scratch() { for (tnb_buf* b : bufs) tnb_free(ctx, b); }
    int get(size_t bytes, tnb_buf** out) {
        int rc = tnb_alloc(ctx, bytes ? bytes : 16, out);
        if (!rc) bufs.push_back(*out);
        return rc;
    }
};

// C = A (x) 1 : moves / permutes / conjugates a tensor between layouts with the engine's own einsum kernel
int move_tensor(tnb_ctx* ctx, const tnb_tensor* src, const tnb_tensor* dst, tnb_buf* one) {
    tnb_tensor o;
    memset(&o, 0, sizeof o);
    o.buf = one; o.dtype = src->dtype; o.rank = 0;
    return tnb_binary_einsum(ctx, src, &o, dst, nullptr, 0, nullptr, nullptr);
}

int make_one(tnb_ctx* ctx, Scratch& sc, int dtype, tnb_buf** one) {
    int rc = sc.get(16, one);
    if (rc) return rc;
    const double d2[2] = {1.0, 0.0};
    const float f2[2] = {1.f, 0.f};
    const bool dbl = dtype == TNB_C128 || dtype == TNB_F64;
    return tnb_upload(ctx, *one, 0, dbl ? (const void*)d2 : (const void*)f2, tnb_dtype_size(dtype));
}

int check_info(tnb_ctx* ctx, tnb_buf* dinfo, const char* what) {
    int info = 0;
    int rc = tnb_download(ctx, dinfo, 0, &info, sizeof info);
    if (rc) return rc;
    if (info != 0) return tnb_set_error(ctx, TNB_ECUDA, "%s: cuSOLVER info = %d", what, info);
    return TNB_OK;
}

}  // namespace

extern "C" {

int tnb_qr_thin(tnb_ctx* ctx, const tnb_tensor* A, const int32_t* row_modes, int32_t nrow, int32_t virtual_mode,
                const tnb_tensor* Q, const tnb_tensor* R) {
    if (!ctx || !A || !Q || !R || (nrow > 0 && !row_modes) || nrow < 0) return tnb_set_error(ctx, TNB_EINVAL, "qr_thin: bad arguments");
    if (A->dtype != Q->dtype || A->dtype != R->dtype) return tnb_set_error(ctx, TNB_EINVAL, "qr_thin: dtype mismatch");
    std::string why;
    SolverApi* api = load_solver(&why);
    if (!api) return tnb_set_error(ctx, TNB_EUNSUPPORTED, "qr_thin: cannot load cuSOLVER: %s", why.c_str());
    Split sp;
    int rc = split_modes(ctx, A, row_modes, nrow, &sp);
    if (rc) return rc;
    const int64_t m = sp.m, n = sp.n, k = std::min(m, n);
    if (m >= (1ll << 31) || n >= (1ll << 31)) return tnb_set_error(ctx, TNB_EUNSUPPORTED, "qr_thin: matrix too large");
    const int dtype = A->dtype;
    const size_t esz = tnb_dtype_size(dtype);
    cudaSetDevice(ctx->device);
    Scratch sc(ctx);
    tnb_buf *W, *tau, *Rd, *one, *dinfo, *dwork = nullptr;
    if ((rc = sc.get((size_t)m * n * esz, &W)) || (rc = sc.get((size_t)k * esz, &tau)) || (rc = sc.get((size_t)k * n * esz, &Rd)) ||
        (rc = sc.get(sizeof(int), &dinfo)) || (rc = make_one(ctx, sc, dtype, &one)))
        return rc;
    // 1. gather A -> W[rows.., cols..] dense column-major (conj applied)
    Dense Wd(W, dtype, cat(sp.row_modes, sp.col_modes), cat(sp.row_ext, sp.col_ext));
    if ((rc = move_tensor(ctx, A, &Wd.t, one))) return rc;
    // 2. geqrf (the cuSOLVER handle is process-wide: one factorisation at a time)
    std::lock_guard<std::mutex> g(solver_mutex());
    api->SetStream(api->handle, ctx->stream);
    size_t wdev = 0, whost = 0;
    const cudaDataType ct = cuda_type(dtype);
    if (api->XgeqrfBuf(api->handle, api->params, m, n, ct, W->ptr, m, ct, tau->ptr, ct, &wdev, &whost) != CUSOLVER_STATUS_SUCCESS)
        return tnb_set_error(ctx, TNB_ECUDA, "qr_thin: cusolverDnXgeqrf_bufferSize failed");
    std::vector<char> hwork(whost);
    if ((rc = sc.get(wdev, &dwork))) return rc;
    if (api->Xgeqrf(api->handle, api->params, m, n, ct, W->ptr, m, ct, tau->ptr, ct, dwork->ptr, wdev, hwork.data(), whost,
                    (int*)dinfo->ptr) != CUSOLVER_STATUS_SUCCESS)
        return tnb_set_error(ctx, TNB_ECUDA, "qr_thin: cusolverDnXgeqrf failed");
    // 3. R = upper triangle of the leading k x n block
    {
        const int64_t cnt = k * n;
        const unsigned blocks = (unsigned)std::min<int64_t>((cnt + 255) / 256, 148 * 8);
        switch (dtype) {
            case TNB_C128: upper_kernel<double2><<<blocks, 256, 0, ctx->stream>>>((const double2*)W->ptr, m, (double2*)Rd->ptr, k, n); break;
            case TNB_C64: upper_kernel<float2><<<blocks, 256, 0, ctx->stream>>>((const float2*)W->ptr, m, (float2*)Rd->ptr, k, n); break;
            case TNB_F64: upper_kernel<double><<<blocks, 256, 0, ctx->stream>>>((const double*)W->ptr, m, (double*)Rd->ptr, k, n); break;
            default: upper_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float*)W->ptr, m, (float*)Rd->ptr, k, n); break;
        }
        ctx->launches++;
    }
    // 4. Q = first k columns of the product of the reflectors
    int lwork = 0;
    if (api->gqrBuf[dtype](api->handle, (int)m, (int)k, (int)k, W->ptr, (int)m, tau->ptr, &lwork) != CUSOLVER_STATUS_SUCCESS)
        return tnb_set_error(ctx, TNB_ECUDA, "qr_thin: orgqr_bufferSize failed");
    tnb_buf* qwork = nullptr;
    if ((rc = sc.get((size_t)lwork * esz, &qwork))) return rc;
    if (api->gqr[dtype](api->handle, (int)m, (int)k, (int)k, W->ptr, (int)m, tau->ptr, qwork->ptr, lwork, (int*)dinfo->ptr) != CUSOLVER_STATUS_SUCCESS)
        return tnb_set_error(ctx, TNB_ECUDA, "qr_thin: orgqr failed");
    // 5. scatter into the caller's layouts
    Dense Qd(W, dtype, cat(sp.row_modes, {virtual_mode}), cat(sp.row_ext, {k}));
    Dense Rdd(Rd, dtype, cat({virtual_mode}, sp.col_modes), cat({k}, sp.col_ext));
    rc = move_tensor(ctx, &Qd.t, Q, one);
    if (!rc) rc = move_tensor(ctx, &Rdd.t, R, one);
    if (!rc) rc = check_info(ctx, dinfo, "qr_thin");
    return rc;
}

int tnb_svd_thin(tnb_ctx* ctx, const tnb_tensor* A, const int32_t* row_modes, int32_t nrow, int32_t virtual_mode,
                 const tnb_tensor* U, const tnb_tensor* S, const tnb_tensor* Vh) {
    if (!ctx || !A || !U || !S || !Vh || (nrow > 0 && !row_modes) || nrow < 0) return tnb_set_error(ctx, TNB_EINVAL, "svd_thin: bad arguments");
    if (A->dtype != U->dtype || A->dtype != Vh->dtype || A->dtype != S->dtype) return tnb_set_error(ctx, TNB_EINVAL, "svd_thin: dtype mismatch");
    std::string why;
    SolverApi* api = load_solver(&why);
    if (!api) return tnb_set_error(ctx, TNB_EUNSUPPORTED, "svd_thin: cannot load cuSOLVER: %s", why.c_str());
    Split sp;
    int rc = split_modes(ctx, A, row_modes, nrow, &sp);
    if (rc) return rc;
    const int dtype = A->dtype;
    const size_t esz = tnb_dtype_size(dtype);
    const bool cplx = tnb_dtype_complex(dtype);
    // gesvd wants m >= n: a wide matrix is factorised through its conjugate transpose, A^H = V S U^H
    const bool flip = sp.m < sp.n;
    const int64_t m = flip ? sp.n : sp.m, n = flip ? sp.m : sp.n, k = n;
    const std::vector<int32_t>& gm = flip ? sp.col_modes : sp.row_modes;      // modes along the rows of the gathered matrix
    const std::vector<int32_t>& gn = flip ? sp.row_modes : sp.col_modes;
    const std::vector<int64_t>& em = flip ? sp.col_ext : sp.row_ext;
    const std::vector<int64_t>& en = flip ? sp.row_ext : sp.col_ext;
    cudaSetDevice(ctx->device);
    Scratch sc(ctx);
    tnb_buf *W, *Ub, *Vb, *Sb, *Sw, *one, *dinfo, *dwork = nullptr;
    if ((rc = sc.get((size_t)m * n * esz, &W)) || (rc = sc.get((size_t)m * k * esz, &Ub)) || (rc = sc.get((size_t)k * n * esz, &Vb)) ||
        (rc = sc.get((size_t)k * 8, &Sb)) || (rc = sc.get((size_t)k * esz, &Sw)) || (rc = sc.get(sizeof(int), &dinfo)) ||
        (rc = make_one(ctx, sc, dtype, &one)))
        return rc;
    // gather: W = A (tall) or conj(A)^T (wide): the conj flag of the source descriptor flips for the wide case
    tnb_tensor Asrc = *A;
    if (flip && cplx) Asrc.conj = !A->conj;
    Dense Wd(W, dtype, cat(gm, gn), cat(em, en));
    if ((rc = move_tensor(ctx, &Asrc, &Wd.t, one))) return rc;
    {
        std::lock_guard<std::mutex> g(solver_mutex());
        api->SetStream(api->handle, ctx->stream);
        size_t wdev = 0, whost = 0;
        const cudaDataType ct = cuda_type(dtype), rt = cuda_real_type(dtype);
        if (api->XgesvdBuf(api->handle, api->params, 'S', 'S', m, n, ct, W->ptr, m, rt, Sb->ptr, ct, Ub->ptr, m, ct, Vb->ptr, k, ct, &wdev,
                           &whost) != CUSOLVER_STATUS_SUCCESS)
            return tnb_set_error(ctx, TNB_ECUDA, "svd_thin: cusolverDnXgesvd_bufferSize failed");
        std::vector<char> hwork(whost);
        if ((rc = sc.get(wdev, &dwork))) return rc;
        if (api->Xgesvd(api->handle, api->params, 'S', 'S', m, n, ct, W->ptr, m, rt, Sb->ptr, ct, Ub->ptr, m, ct, Vb->ptr, k, ct, dwork->ptr,
                        wdev, hwork.data(), whost, (int*)dinfo->ptr) != CUSOLVER_STATUS_SUCCESS)
            return tnb_set_error(ctx, TNB_ECUDA, "svd_thin: cusolverDnXgesvd failed");
        const unsigned blocks = (unsigned)std::min<int64_t>((k + 255) / 256, 148 * 8);
        switch (dtype) {
            case TNB_C128: widen_kernel<double, double2><<<blocks, 256, 0, ctx->stream>>>((const double*)Sb->ptr, (double2*)Sw->ptr, k); break;
            case TNB_C64: widen_kernel<float, float2><<<blocks, 256, 0, ctx->stream>>>((const float*)Sb->ptr, (float2*)Sw->ptr, k); break;
            case TNB_F64: cudaMemcpyAsync(Sw->ptr, Sb->ptr, (size_t)k * 8, cudaMemcpyDeviceToDevice, ctx->stream); break;
            default: cudaMemcpyAsync(Sw->ptr, Sb->ptr, (size_t)k * 4, cudaMemcpyDeviceToDevice, ctx->stream); break;
        }
        ctx->launches++;
    }
    // scatter.  Tall: U = Ub [rows.., k], Vh = Vb [k, cols..].  Wide: A = (Ub S Vb)^H  =>  U = conj(Vb)^T [rows.., k] with Vb [k, rows..],
    // Vh = conj(Ub)^T [k, cols..] with Ub [cols.., k]: the same dense buffers, read with swapped mode lists and the conj flag.
    const int cj = (flip && cplx) ? 1 : 0;
    Dense Ud(flip ? Vb : Ub, dtype, flip ? cat({virtual_mode}, sp.row_modes) : cat(sp.row_modes, {virtual_mode}),
             flip ? cat({k}, sp.row_ext) : cat(sp.row_ext, {k}), cj);
    Dense Vd(flip ? Ub : Vb, dtype, flip ? cat(sp.col_modes, {virtual_mode}) : cat({virtual_mode}, sp.col_modes),
             flip ? cat(sp.col_ext, {k}) : cat({k}, sp.col_ext), cj);
    Dense Sd(Sw, dtype, {virtual_mode}, {k});
    rc = move_tensor(ctx, &Ud.t, U, one);
    if (!rc) rc = move_tensor(ctx, &Vd.t, Vh, one);
    if (!rc) rc = move_tensor(ctx, &Sd.t, S, one);
    if (!rc) rc = check_info(ctx, dinfo, "svd_thin");
    return rc;
}

}  // extern "C"
