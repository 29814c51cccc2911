// api.cu — the two hot-path entry points of the C ABI: one pairwise einsum, and a whole sliced path.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <set>
#include "tnb_internal.h"

namespace {

PlanTensor to_plan_tensor(const tnb_tensor* T) {
    PlanTensor p;
    for (int r = 0; r < T->rank; r++) {
        p.modes.push_back(T->mode[r]);
        p.ext.push_back(T->extent[r]);
        p.stride.push_back(T->stride_elems[r]);
    }
    p.conj = T->conj;
    p.offset = T->offset_elems;
    return p;
}

// append every table of a step to a host blob (int64 entries), remembering positions
void pack_tables(StepSpec& S, std::vector<int64_t>& blob) {
    HostTable* tabs[9] = {&S.am, &S.ak, &S.al, &S.bn, &S.bk, &S.bl, &S.cm, &S.cn, &S.cl};
    for (HostTable* t : tabs) {
        t->lo_pos = blob.size();
        blob.insert(blob.end(), t->lo.begin(), t->lo.end());
        t->hi_pos = blob.size();
        blob.insert(blob.end(), t->hi.begin(), t->hi.end());
    }
    if (S.st_ok) {
        S.st_hi_pos = blob.size();  blob.insert(blob.end(), S.st_hi.begin(), S.st_hi.end());
        S.st_rel_pos = blob.size(); blob.insert(blob.end(), S.st_rel.begin(), S.st_rel.end());
        S.st_pos_pos = blob.size(); blob.insert(blob.end(), S.st_pos.begin(), S.st_pos.end());
    }
}

TabRef make_ref(const HostTable& t, const int64_t* dev_blob) {
    TabRef r;
    r.lo = dev_blob + t.lo_pos;
    r.hi = dev_blob + t.hi_pos;
    r.lo_size = (uint32_t)t.lo_size;
    r.affine = t.affine ? 1 : 0;
    r.stride = t.stride;
    return r;
}

void fill_args(const StepSpec& S, const int64_t* dev_blob, EinsumArgs* a) {
    memset(a, 0, sizeof *a);
    a->M = S.M; a->N = S.N; a->K = S.K; a->L = S.L;
    a->am = make_ref(S.am, dev_blob); a->ak = make_ref(S.ak, dev_blob); a->al = make_ref(S.al, dev_blob);
    a->bn = make_ref(S.bn, dev_blob); a->bk = make_ref(S.bk, dev_blob); a->bl = make_ref(S.bl, dev_blob);
    a->cm = make_ref(S.cm, dev_blob); a->cn = make_ref(S.cn, dev_blob); a->cl = make_ref(S.cl, dev_blob);
    a->conjA = S.conjA; a->conjB = S.conjB;
    a->a_kfast = S.a_kfast; a->b_kfast = S.b_kfast;
    a->alpha[0] = 1.0; a->alpha[1] = 0.0; a->beta[0] = 0.0; a->beta[1] = 0.0;
    a->splitk = 1; a->kchunk = S.K; a->ws = nullptr;
}

void read_scalar(int dtype, const void* p, double dflt, double out[2]) {
    out[0] = dflt; out[1] = 0.0;
    if (!p) return;
    switch (dtype) {
        case TNB_C128: out[0] = ((const double*)p)[0]; out[1] = ((const double*)p)[1]; break;
        case TNB_C64: out[0] = ((const float*)p)[0]; out[1] = ((const float*)p)[1]; break;
        case TNB_F64: out[0] = ((const double*)p)[0]; break;
        case TNB_F32: out[0] = ((const float*)p)[0]; break;
    }
}

// Pick the kernel for a planned step (sets kernel/splitk/kchunk/tc fields); returns workspace elements needed.
int64_t select_kernel(const tnb_ctx* ctx, int dtype, StepSpec& S) {
    int64_t kchunk = S.K, ws_elems = 0;
    S.kernel = TNB_KERNEL_GENERIC;
    S.splitk = 1;
    S.kchunk = S.K;
    S.tc_nt = 0;
    S.tc_swap = false;
    S.kred = false;
    const bool force_generic = ctx && ctx->force_generic;
    if (!force_generic) {
        int thin = tnb_choose_thin(ctx, S.M, S.N, S.K, S.L, &kchunk, &ws_elems);
        if (thin > 0) {
            S.kernel = TNB_KERNEL_STREAM; S.splitk = thin; S.kchunk = kchunk;
            return ws_elems;
        }
        int kred = tnb_choose_kred(ctx, dtype, S.M, S.N, S.K, S.L, S.a_mmajor, S.b_nmajor, &kchunk, &ws_elems);
        if (kred > 0) {
            S.kernel = TNB_KERNEL_STREAM; S.kred = true; S.splitk = kred; S.kchunk = kchunk;
            return ws_elems;
        }
        if (S.st_ok && !S.st_tc && (dtype == TNB_C64 || dtype == TNB_C128)) {
            S.kernel = TNB_KERNEL_STEM;
            return 0;
        }
        if (S.st_ok && S.st_tc && dtype == TNB_C64 && (!ctx || ctx->c64_mode != TNB_C64_SIMT)) {
            S.kernel = TNB_KERNEL_STEM_TC;
            return tnb_stem_tc_ws_elems(S.st_ncol, S.K, S.st_npass);
        }
        const bool tc_ok = dtype == TNB_C64 && (!ctx || ctx->c64_mode != TNB_C64_SIMT) &&
                           (double)S.M * (double)S.N * (double)S.K >= (double)(1ll << 18);
        if (tc_ok) {
            int nt = tnb_tc_c64_tile(S.M, S.N, S.K, S.L, S.a_mmajor, S.b_nmajor);
            int nt_sw = tnb_tc_c64_tile(S.N, S.M, S.K, S.L, S.b_nmajor, S.a_mmajor);
            const int64_t lda = S.K > 1 ? S.ak.stride : S.M, ldb = S.K > 1 ? S.bk.stride : S.N;
            if ((nt || nt_sw) && lda % 2 == 0 && ldb % 2 == 0) {
                S.kernel = TNB_KERNEL_C64_TF32;
                S.tc_swap = nt_sw > nt || (nt_sw == nt && S.N > S.M);   // longer free group on the 128-row side
                S.tc_nt = S.tc_swap ? nt_sw : nt;
                const int64_t m = S.tc_swap ? S.N : S.M, n = S.tc_swap ? S.M : S.N;
                S.splitk = tnb_tc_c64_splitk(ctx, m, n, S.K, &kchunk, &ws_elems);
                S.kchunk = kchunk;                                       // in k-blocks of 8 for this kernel
                return ws_elems;
            }
        }
    }
    if (!force_generic && dtype == TNB_C128 && S.M * S.N >= 1024 && S.K >= 4 &&
        (double)S.M * (double)S.N * (double)S.K >= (double)(1ll << 16)) {
        // FP64 tensor-core tile kernel (128 x 64 tile): put the longer free group on the 128 side
        S.kernel = TNB_KERNEL_C128_DMMA;
        S.tc_swap = S.N > S.M;
        const int64_t m = S.tc_swap ? S.N : S.M, n = S.tc_swap ? S.M : S.N;
        S.splitk = tnb_choose_splitk_dmma(ctx, m, n, S.K, S.L, &kchunk, &ws_elems, &S.dmma_small);
        S.kchunk = kchunk;
        return ws_elems;
    }
    S.splitk = tnb_choose_splitk(ctx, S.M, S.N, S.K, S.L, &kchunk, &ws_elems);
    S.kchunk = kchunk;
    if (S.splitk > 1) S.kernel = TNB_KERNEL_SPLITK;
    return ws_elems;
}

// launch one planned step. A/B/C are element pointers already offset to the operand origin.
int run_step(tnb_ctx* ctx, int dtype, const StepSpec& S, const int64_t* dev_blob, const void* A, const void* B,
             void* C, const double alpha[2], const double beta[2], void* ws) {
    EinsumArgs a;
    fill_args(S, dev_blob, &a);
    a.A = A; a.B = B; a.C = C;
    a.alpha[0] = alpha[0]; a.alpha[1] = alpha[1];
    a.beta[0] = beta[0]; a.beta[1] = beta[1];
    if (S.kernel == TNB_KERNEL_STEM || S.kernel == TNB_KERNEL_STEM_TC) {
        StemArgs t;
        memset(&t, 0, sizeof t);
        const bool sw = S.st_swap;
        t.A = sw ? B : A; t.B = sw ? A : B; t.C = (char*)C + 0;
        t.M = sw ? S.N : S.M; t.N = (int32_t)(sw ? S.M : S.N); t.K = (int32_t)S.K;
        t.lda = S.K > 1 ? (sw ? S.bk.stride : S.ak.stride) : t.M;
        t.TM = S.st_tm; t.contig = S.st_contig ? 1 : 0; t.run = S.st_run;
        t.additive = S.st_additive ? 1 : 0; t.even = S.st_even ? 1 : 0; t.direct = S.st_direct ? 1 : 0; t.pairs = S.st_pairs ? 1 : 0;
        t.conjA = sw ? S.conjB : S.conjA; t.conjB = sw ? S.conjA : S.conjB;
        t.bn = sw ? a.am : a.bn; t.bk = sw ? a.ak : a.bk;
        t.hi = dev_blob + S.st_hi_pos; t.rel = dev_blob + S.st_rel_pos; t.pos = dev_blob + S.st_pos_pos;
        t.alpha[0] = alpha[0]; t.alpha[1] = alpha[1]; t.beta[0] = beta[0]; t.beta[1] = beta[1];
        if (S.kernel == TNB_KERNEL_STEM) return tnb_launch_stem(ctx, dtype, t);
        const int64_t cnt = (int64_t)S.st_tm * S.st_ncol;
        // 128-column passes with a DENSE small operand: all passes in one launch of the CTA-pair kernel (cta_group::2 MMAs,
        // B tile by TMA) with the same sorted-pattern epilogue; otherwise (gathered small operand, misalignment,
        // TNB_OPT_GEMM_PAIR = 0) the 1-CTA stem kernel below
        if (S.st_ncol == 128 && S.st_tm == 128 && (sw ? S.a_mmajor : S.b_nmajor)) {
            const int64_t Ns = sw ? S.M : S.N;
            const int64_t ldb = S.K > 1 ? (sw ? S.ak.stride : S.bk.stride) : Ns;
            int rc = tnb_launch_c64_pair_staged(ctx, t, Ns, ldb, S.st_rel_small);
            if (rc != -1) return rc;
        }
        for (int ps = 0; ps < S.st_npass; ps++) {
            t.N = S.st_ncol; t.n0 = ps * S.st_ncol;
            t.rel = dev_blob + S.st_rel_pos + ps * cnt; t.pos = dev_blob + S.st_pos_pos + ps * cnt;
            int rc = tnb_launch_c64_stem_tc(ctx, t, ws, S.st_npass);
            if (rc == -1 && ps == 0)
                return tnb_launch_einsum_generic(ctx, dtype, a);  // misaligned big operand: exact-FP32 generic kernel
            if (rc) return rc;
        }
        return TNB_OK;
    }
    if (S.kernel == TNB_KERNEL_C64_TF32) {
        int64_t lda = S.K > 1 ? S.ak.stride : S.M, ldb = S.K > 1 ? S.bk.stride : S.N;
        if (S.tc_swap) {   // C^T = B * A^T : swap operand roles, the offset tables of C swap with them
            std::swap(a.A, a.B); std::swap(a.M, a.N); std::swap(a.cm, a.cn); std::swap(a.conjA, a.conjB);
            std::swap(lda, ldb);
        }
        if (S.splitk > 1 && ws) { a.splitk = S.splitk; a.kchunk = S.kchunk; a.ws = ws; }
        int rc = tnb_launch_c64_tc(ctx, a, S.tc_nt, lda, ldb, ctx->c64_mode != TNB_C64_TF32X3_FAST);
        if (rc == TNB_OK && a.splitk > 1) return tnb_launch_splitk_reduce(ctx, dtype, a);
        if (rc != -1) return rc;
        // operands not 16-byte aligned for the bulk copies (odd leaf offset): exact-FP32 generic kernel instead
        EinsumArgs g;
        fill_args(S, dev_blob, &g);
        g.A = A; g.B = B; g.C = C;
        g.alpha[0] = alpha[0]; g.alpha[1] = alpha[1]; g.beta[0] = beta[0]; g.beta[1] = beta[1];
        return tnb_launch_einsum_generic(ctx, dtype, g);
    }
    if (S.kernel == TNB_KERNEL_C128_DMMA) {
        a.pad_ = S.dmma_small;
        if (S.tc_swap) {   // C^T = B * A^T
            std::swap(a.A, a.B); std::swap(a.M, a.N); std::swap(a.conjA, a.conjB);
            std::swap(a.am, a.bn); std::swap(a.ak, a.bk); std::swap(a.al, a.bl); std::swap(a.cm, a.cn);
            std::swap(a.a_kfast, a.b_kfast);
        }
        if (S.splitk > 1 && ws) {
            a.splitk = S.splitk; a.kchunk = S.kchunk; a.ws = ws;
            int rc = tnb_launch_c128_dmma(ctx, a);
            if (rc) return rc;
            return tnb_launch_splitk_reduce(ctx, dtype, a);
        }
        return tnb_launch_c128_dmma(ctx, a);
    }
    if (S.kernel == TNB_KERNEL_STREAM && ws) {
        a.splitk = S.splitk; a.kchunk = S.kchunk; a.ws = ws;
        int rc = S.kred ? tnb_launch_einsum_kred(ctx, dtype, a) : tnb_launch_einsum_thin(ctx, dtype, a);
        if (rc == -1) {   // misaligned operands: same split through the generic kernel
            rc = tnb_launch_einsum_generic(ctx, dtype, a);
        }
        if (rc) return rc;
        return tnb_launch_splitk_reduce(ctx, dtype, a);
    }
    if (S.kernel == TNB_KERNEL_SPLITK && ws) {
        a.splitk = S.splitk; a.kchunk = S.kchunk; a.ws = ws;
        int rc = tnb_launch_einsum_generic(ctx, dtype, a);
        if (rc) return rc;
        return tnb_launch_splitk_reduce(ctx, dtype, a);
    }
    return tnb_launch_einsum_generic(ctx, dtype, a);
}

int check_tensor(tnb_ctx* ctx, const tnb_tensor* T, const char* name, bool need_buf) {
    if (!T) return tnb_set_error(ctx, TNB_EINVAL, "%s descriptor is NULL", name);
    if (T->rank < 0 || T->rank > TNB_MAX_RANK) return tnb_set_error(ctx, TNB_EUNSUPPORTED, "%s rank %d out of range", name, T->rank);
    if (T->rank > 0 && (!T->extent || !T->stride_elems || !T->mode)) return tnb_set_error(ctx, TNB_EINVAL, "%s has NULL extent/stride/mode", name);
    if (!tnb_dtype_size(T->dtype)) return tnb_set_error(ctx, TNB_EUNSUPPORTED, "%s has unsupported dtype %d", name, T->dtype);
    if (need_buf) {
        if (!T->buf) return tnb_set_error(ctx, TNB_EINVAL, "%s has no device buffer", name);
        // bounds: min/max reachable element
        int64_t lo = T->offset_elems, hi = T->offset_elems;
        for (int r = 0; r < T->rank; r++) {
            if (T->extent[r] < 1) return tnb_set_error(ctx, TNB_EINVAL, "%s extent[%d] < 1", name, r);
            int64_t span = (T->extent[r] - 1) * T->stride_elems[r];
            if (span > 0) hi += span; else lo += span;
        }
        size_t esz = tnb_dtype_size(T->dtype);
        if (lo < 0 || (size_t)(hi + 1) * esz > T->buf->cap)
            return tnb_set_error(ctx, TNB_EINVAL, "%s addresses elements [%lld,%lld] outside its %zu-byte buffer", name,
                                 (long long)lo, (long long)hi, T->buf->cap);
    }
    return TNB_OK;
}

}  // namespace

extern "C" {

int tnb_binary_einsum_result(const tnb_tensor* A, const tnb_tensor* B, const int32_t* sum_modes, int32_t nsum,
                             int32_t* out_rank, int32_t* out_modes, int64_t* out_extents) {
    if (!A || !B || !out_rank || !out_modes || !out_extents) return TNB_EINVAL;
    if (A->rank < 0 || A->rank > TNB_MAX_RANK || B->rank < 0 || B->rank > TNB_MAX_RANK) return TNB_EUNSUPPORTED;
    if ((A->rank > 0 && (!A->mode || !A->extent)) || (B->rank > 0 && (!B->mode || !B->extent))) return TNB_EINVAL;
    std::set<int32_t> sum(sum_modes, sum_modes + (sum_modes ? nsum : 0));
    std::set<int32_t> inA(A->mode, A->mode + A->rank), inB(B->mode, B->mode + B->rank);
    int n = 0;
    // out_modes / out_extents hold TNB_MAX_RANK entries: check before every store
    auto put = [&](int32_t mode, int64_t ext) {
        if (n >= TNB_MAX_RANK) return false;
        out_modes[n] = mode; out_extents[n++] = ext;
        return true;
    };
    // free(A) in A's order, free(B) in B's order, then batch modes in A's order
    for (int r = 0; r < A->rank; r++)
        if (!sum.count(A->mode[r]) && !inB.count(A->mode[r]) && !put(A->mode[r], A->extent[r])) return TNB_EUNSUPPORTED;
    for (int r = 0; r < B->rank; r++)
        if (!sum.count(B->mode[r]) && !inA.count(B->mode[r]) && !put(B->mode[r], B->extent[r])) return TNB_EUNSUPPORTED;
    for (int r = 0; r < A->rank; r++)
        if (!sum.count(A->mode[r]) && inB.count(A->mode[r]) && !put(A->mode[r], A->extent[r])) return TNB_EUNSUPPORTED;
    *out_rank = n;
    return TNB_OK;
}

int tnb_binary_einsum(tnb_ctx* ctx, const tnb_tensor* A, const tnb_tensor* B, const tnb_tensor* C,
                      const int32_t* sum_modes, int32_t nsum, const void* alpha, const void* beta) {
    if (!ctx) return TNB_EINVAL;
    int rc;
    if ((rc = check_tensor(ctx, A, "A", true))) return rc;
    if ((rc = check_tensor(ctx, B, "B", true))) return rc;
    if ((rc = check_tensor(ctx, C, "C", true))) return rc;
    if (A->dtype != B->dtype || A->dtype != C->dtype)
        return tnb_set_error(ctx, TNB_EINVAL, "operands must share one dtype (promote on the host): A=%d B=%d C=%d", A->dtype, B->dtype, C->dtype);
    if (nsum < 0 || (nsum > 0 && !sum_modes)) return tnb_set_error(ctx, TNB_EINVAL, "bad sum_modes");
    const int dtype = A->dtype;
    const size_t esz = tnb_dtype_size(dtype);
    StepSpec S;
    std::string msg;
    std::vector<int32_t> sum(sum_modes, sum_modes + nsum);
    PlanTensor Ct = to_plan_tensor(C);
    Ct.conj = 0;
    rc = tnb_build_step(to_plan_tensor(A), to_plan_tensor(B), Ct, sum, tnb_dtype_complex(dtype), esz, &S, &msg);
    if (rc) return tnb_set_error(ctx, rc, "binary_einsum: %s", msg.c_str());
    cudaSetDevice(ctx->device);

    std::vector<int64_t> blob;
    pack_tables(S, blob);
    tnb_buf* tb = nullptr;
    if ((rc = tnb_alloc(ctx, blob.size() * sizeof(int64_t), &tb))) return rc;
    cudaError_t e = cudaMemcpyAsync(tb->ptr, blob.data(), blob.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { tnb_free(ctx, tb); return tnb_set_error(ctx, TNB_ECUDA, "table upload: %s", cudaGetErrorString(e)); }
    // the blob is pageable host memory: cudaMemcpyAsync has staged it before returning

    tnb_buf* ws = nullptr;
    int64_t ws_elems = select_kernel(ctx, dtype, S);
    ctx->last_kernel = S.kernel;
    if (ws_elems > 0) {
        if ((rc = tnb_alloc(ctx, (size_t)ws_elems * esz, &ws))) { tnb_free(ctx, tb); return rc; }
    }
    double al[2], be[2];
    read_scalar(dtype, alpha, 1.0, al);
    read_scalar(dtype, beta, 0.0, be);
    const char* Ap = (const char*)A->buf->ptr + (size_t)A->offset_elems * esz;
    const char* Bp = (const char*)B->buf->ptr + (size_t)B->offset_elems * esz;
    char* Cp = (char*)C->buf->ptr + (size_t)C->offset_elems * esz;
    rc = run_step(ctx, dtype, S, (const int64_t*)tb->ptr, Ap, Bp, Cp, al, be, ws ? ws->ptr : nullptr);
    tnb_free(ctx, tb);
    if (ws) tnb_free(ctx, ws);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// plans
// ---------------------------------------------------------------------------------------------
static int plan_create_impl(tnb_ctx* ctx, const tnb_tensor* leaves, int32_t nleaves, const int32_t* steps,
                            int32_t nsteps, const int32_t* sliced_modes, int32_t nsliced, const tnb_tensor* out,
                            bool dry, tnb_plan** plan_out) {
    if (!plan_out) return tnb_set_error(ctx, TNB_EINVAL, "plan is NULL");
    *plan_out = nullptr;
    if (!leaves || nleaves < 1 || (nsteps > 0 && !steps) || nsliced < 0 || (nsliced > 0 && !sliced_modes))
        return tnb_set_error(ctx, TNB_EINVAL, "plan_create: bad arguments");
    int rc;
    for (int i = 0; i < nleaves; i++) {
        char name[32];
        snprintf(name, sizeof name, "leaf %d", i);
        if ((rc = check_tensor(ctx, &leaves[i], name, !dry))) return rc;
    }
    if ((rc = check_tensor(ctx, out, "out", !dry))) return rc;
    tnb_plan* P = new (std::nothrow) tnb_plan();
    if (!P) return tnb_set_error(ctx, TNB_ENOMEM, "host allocation failed");
    P->dry = dry;
    std::string msg;
    rc = tnb_plan_build(leaves, nleaves, steps, nsteps, sliced_modes, nsliced, out, P, &msg);
    if (rc) { delete P; return tnb_set_error(ctx, rc, "plan_create: %s", msg.c_str()); }
    // kernel selection + split-K workspace
    const size_t esz = tnb_dtype_size(P->dtype);
    int64_t ws_max = 0;
    for (StepSpec& S : P->steps) {
        int64_t ws_elems = select_kernel(ctx, P->dtype, S);
        ws_max = std::max(ws_max, ws_elems);
        pack_tables(S, P->table_blob);
    }
    P->ws_elems = ws_max;
    P->info.table_bytes = (int64_t)(P->table_blob.size() * sizeof(int64_t));
    if (!dry) {
        cudaSetDevice(ctx->device);
        if (P->arena_elems > 0 && (rc = tnb_alloc(ctx, (size_t)P->arena_elems * esz, &P->arena))) { delete P; return rc; }
        if ((rc = tnb_alloc(ctx, std::max<size_t>(8, P->table_blob.size() * sizeof(int64_t)), &P->tables))) {
            tnb_free(ctx, P->arena); delete P; return rc;
        }
        if (ws_max > 0 && (rc = tnb_alloc(ctx, (size_t)ws_max * esz, &P->ws))) {
            tnb_free(ctx, P->arena); tnb_free(ctx, P->tables); delete P; return rc;
        }
        cudaError_t e = cudaMemcpyAsync(P->tables->ptr, P->table_blob.data(), P->table_blob.size() * sizeof(int64_t),
                                        cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            tnb_free(ctx, P->arena); tnb_free(ctx, P->tables); tnb_free(ctx, P->ws); delete P;
            return tnb_set_error(ctx, TNB_ECUDA, "table upload: %s", cudaGetErrorString(e));
        }
    }
    *plan_out = P;
    return TNB_OK;
}

int tnb_plan_create(tnb_ctx* ctx, const tnb_tensor* leaves, int32_t nleaves, const int32_t* steps, int32_t nsteps,
                    const int32_t* sliced_modes, int32_t nsliced, const tnb_tensor* out, tnb_plan** plan) {
    if (!ctx) return TNB_EINVAL;
    return plan_create_impl(ctx, leaves, nleaves, steps, nsteps, sliced_modes, nsliced, out, false, plan);
}

// Planning only (no device, no buffers needed): lets host code and CPU tests inspect what the planner
// decided (GEMM views, offset tables, arena layout).  It cannot execute anything.
int tnb_plan_create_dry(const tnb_tensor* leaves, int32_t nleaves, const int32_t* steps, int32_t nsteps,
                        const int32_t* sliced_modes, int32_t nsliced, const tnb_tensor* out, tnb_plan** plan) {
    return plan_create_impl(nullptr, leaves, nleaves, steps, nsteps, sliced_modes, nsliced, out, true, plan);
}

int tnb_plan_destroy(tnb_ctx* ctx, tnb_plan* plan) {
    if (!plan) return TNB_OK;
    for (auto& e : plan->ev) cudaEventDestroy(e);
    for (auto& g : plan->gexec) if (g) cudaGraphExecDestroy(g);
    if (!plan->dry && ctx) {
        tnb_free(ctx, plan->arena);
        tnb_free(ctx, plan->tables);
        tnb_free(ctx, plan->ws);
    }
    delete plan;
    return TNB_OK;
}

// Per-step device timing: when enabled, every step launch is bracketed by CUDA events on the context
// stream and the host synchronises once per slice to read them (bench.py's roofline numbers).
int tnb_plan_profile(tnb_ctx* ctx, tnb_plan* plan, int32_t enable) {
    if (!ctx || !plan || plan->dry) return TNB_EINVAL;
    cudaSetDevice(ctx->device);
    if (enable && plan->ev.empty()) {
        plan->ev.resize(2 * plan->nsteps);
        for (auto& e : plan->ev) TNB_CUDA_CHECK(ctx, cudaEventCreate(&e));
    }
    plan->step_ms.assign(plan->nsteps, 0.0);
    plan->step_runs.assign(plan->nsteps, 0);
    plan->profiling = enable != 0;
    return TNB_OK;
}

int tnb_plan_get_step_time(const tnb_plan* plan, int32_t step, double* ms_total, int64_t* runs) {
    if (!plan || step < 0 || step >= plan->nsteps || plan->step_ms.empty()) return TNB_EINVAL;
    if (ms_total) *ms_total = plan->step_ms[step];
    if (runs) *runs = plan->step_runs[step];
    return TNB_OK;
}

int tnb_plan_get_info(const tnb_plan* plan, tnb_plan_info* info) {
    if (!plan || !info) return TNB_EINVAL;
    *info = plan->info;
    return TNB_OK;
}

int tnb_plan_get_step(const tnb_plan* plan, int32_t step, tnb_step_info* info) {
    if (!plan || !info || step < 0 || step >= plan->nsteps) return TNB_EINVAL;
    const StepSpec& S = plan->steps[step];
    info->M = S.M; info->N = S.N; info->K = S.K; info->L = S.L;
    info->kernel = S.kernel;
    info->hoisted = S.hoisted ? 1 : 0;
    info->flops = S.flops;
    info->bytes = S.bytes;
    return TNB_OK;
}

// Debug/test introspection: copy one offset table of a step, fully expanded (size entries), to host.
// which: 0..8 = am ak al bn bk bl cm cn cl.  Returns the table size; copies min(size, cap) entries.
int64_t tnb_plan_dump_table(const tnb_plan* plan, int32_t step, int32_t which, int64_t* dst, int64_t cap) {
    if (!plan || step < 0 || step >= plan->nsteps || which < 0 || which > 8) return -1;
    const StepSpec& S = plan->steps[step];
    const HostTable* tabs[9] = {&S.am, &S.ak, &S.al, &S.bn, &S.bk, &S.bl, &S.cm, &S.cn, &S.cl};
    const HostTable* t = tabs[which];
    if (dst)
        for (int64_t i = 0; i < t->size && i < cap; i++) dst[i] = t->at(i);
    return t->size;
}

// Debug/test introspection: operand placement of a step.  ids[3] = a,b,c node ids; base[3] = element
// offset of each operand inside its storage (leaf: user offset; intermediate: arena offset; root: out
// offset); kind[3] = 0 leaf, 1 arena, 2 out; slice_stride (nsliced entries per operand, row-major [3][nsliced]).
int tnb_plan_dump_step(const tnb_plan* plan, int32_t step, int32_t* ids, int64_t* base, int32_t* kind,
                       int64_t* slice_stride, int32_t* conj) {
    if (!plan || step < 0 || step >= plan->nsteps) return TNB_EINVAL;
    const StepSpec& S = plan->steps[step];
    const int root = plan->nleaves + plan->nsteps - 1;
    const int id3[3] = {S.a_id, S.b_id, S.c_id};
    const size_t ns = plan->sliced_modes.size();
    for (int i = 0; i < 3; i++) {
        const PlanNode& nd = plan->nodes[id3[i]];
        ids[i] = id3[i];
        if (nd.leaf) { kind[i] = 0; base[i] = plan->leaf_off[id3[i]]; }
        else if (id3[i] == root) { kind[i] = 2; base[i] = plan->out_off; }
        else { kind[i] = 1; base[i] = nd.arena_off; }
        for (size_t j = 0; j < ns; j++) slice_stride[i * ns + j] = nd.leaf ? nd.slice_stride[j] : 0;
    }
    conj[0] = S.conjA; conj[1] = S.conjB;
    return TNB_OK;
}

int tnb_plan_execute(tnb_ctx* ctx, tnb_plan* P, int64_t slice_begin, int64_t slice_step, int64_t slice_end,
                     int32_t accumulate) {
    if (!ctx || !P) return TNB_EINVAL;
    if (P->dry) return tnb_set_error(ctx, TNB_EINVAL, "a dry-run plan cannot be executed");
    if (slice_step < 1 || slice_begin < 0) return tnb_set_error(ctx, TNB_EINVAL, "bad slice range");
    if (slice_end > P->nslices) slice_end = P->nslices;
    cudaSetDevice(ctx->device);
    const int dtype = P->dtype;
    const size_t esz = tnb_dtype_size(dtype);
    const int root = P->nleaves + P->nsteps - 1;
    const int64_t* dev_blob = (const int64_t*)P->tables->ptr;
    const size_t ns = P->sliced_modes.size();
    std::vector<int64_t> digit(ns, 0);

    auto operand_ptr = [&](int id) -> char* {
        const PlanNode& nd = P->nodes[id];
        if (nd.leaf) {
            int64_t off = P->leaf_off[id];
            for (size_t j = 0; j < ns; j++) off += digit[j] * nd.slice_stride[j];
            return (char*)P->leaf_buf[id]->ptr + (size_t)off * esz;
        }
        if (id == root) return (char*)P->out_buf->ptr + (size_t)P->out_off * esz;
        return (char*)P->arena->ptr + (size_t)nd.arena_off * esz;
    };
    const double one[2] = {1.0, 0.0}, zero[2] = {0.0, 0.0};
    std::vector<int> timed;   // steps with events in flight
    auto run = [&](int s, bool acc) -> int {
        const StepSpec& S = P->steps[s];
        const double* beta = (S.c_id == root && acc) ? one : zero;
        if (P->profiling) { cudaEventRecord(P->ev[2 * s], ctx->stream); timed.push_back(s); }
        int r = run_step(ctx, dtype, S, dev_blob, operand_ptr(S.a_id), operand_ptr(S.b_id), operand_ptr(S.c_id), one,
                         beta, P->ws ? P->ws->ptr : nullptr);
        if (P->profiling) cudaEventRecord(P->ev[2 * s + 1], ctx->stream);
        return r;
    };
    auto collect = [&]() {
        if (!P->profiling || timed.empty()) return;
        cudaStreamSynchronize(ctx->stream);
        for (int s : timed) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, P->ev[2 * s], P->ev[2 * s + 1]) == cudaSuccess) {
                P->step_ms[s] += ms;
                P->step_runs[s]++;
            }
        }
        timed.clear();
    };
    int rc;
    bool acc = accumulate != 0;
    const bool root_hoisted = P->steps[P->nsteps - 1].hoisted;
    // accumulate = 0 means out = sum of the requested slices: an empty range (a rank that owns no slice) is zeros
    auto zero_out = [&]() -> int {
        const PlanTensor& ot = P->nodes[root].t;
        return tnb_launch_zero_strided(ctx, dtype, (char*)P->out_buf->ptr + (size_t)P->out_off * esz, (int)ot.ext.size(),
                                       ot.ext.data(), ot.stride.data());
    };
    if (root_hoisted) {
        // no slice dependence at all: one pass (a rank with an empty slice range contributes nothing)
        if (slice_begin != 0) return acc ? TNB_OK : zero_out();
        // Small un-sliced networks (configs[0]: 63 launches of a few microseconds each) are launch-overhead bound
        // (docs/src/manual/reactant.md:52-56 makes the same observation for the reference): the step sequence has the
        // same pointers on every execute, so it is captured once per accumulate flag and replayed as ONE graph launch.
        const int gi = acc ? 1 : 0;
        if (ctx->use_graphs && !P->profiling && !P->graph_failed) {
            if (!P->gexec[gi]) {
                cudaGraph_t graph = nullptr;
                const int64_t l0 = ctx->launches;
                bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                int crc = TNB_OK;
                if (ok) {
                    for (int s : P->order_hoisted)
                        if ((crc = run(s, acc))) break;
                    ok = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && graph != nullptr && crc == TNB_OK;
                }
                if (ok) ok = cudaGraphInstantiate(&P->gexec[gi], graph, 0) == cudaSuccess;
                if (graph) cudaGraphDestroy(graph);
                P->glaunches[gi] = ctx->launches - l0;
                ctx->launches = l0;
                if (!ok) {                      // not capturable on this driver / a launch failed: direct launches from now on
                    cudaGetLastError();
                    P->gexec[gi] = nullptr;
                    P->graph_failed = true;
                    if (crc) return crc;
                }
            }
            if (P->gexec[gi]) {
                TNB_CUDA_CHECK(ctx, cudaGraphLaunch(P->gexec[gi], ctx->stream));
                ctx->launches += P->glaunches[gi];
                return TNB_OK;
            }
        }
        for (int s : P->order_hoisted)
            if ((rc = run(s, acc))) return rc;
        collect();
        return TNB_OK;
    }
    if (slice_begin >= slice_end) return acc ? TNB_OK : zero_out();
    for (int s : P->order_hoisted)
        if ((rc = run(s, false))) return rc;
    collect();
    for (int64_t sl = slice_begin; sl < slice_end; sl += slice_step) {
        int64_t t = sl;
        for (size_t j = 0; j < ns; j++) { digit[j] = t % P->sliced_ext[j]; t /= P->sliced_ext[j]; }
        for (int s : P->order_dep)
            if ((rc = run(s, acc))) return rc;
        collect();
        acc = true;
    }
    return TNB_OK;
}

int tnb_contract_path(tnb_ctx* ctx, const tnb_tensor* leaves, int32_t nleaves, const int32_t* steps, int32_t nsteps,
                      const int32_t* sliced_modes, int32_t nsliced, int64_t slice_begin, int64_t slice_step,
                      int64_t slice_end, const tnb_tensor* out) {
    tnb_plan* P = nullptr;
    int rc = tnb_plan_create(ctx, leaves, nleaves, steps, nsteps, sliced_modes, nsliced, out, &P);
    if (rc) return rc;
    // out = sum of the requested slices: start from zero so that an empty slice range yields zeros
    int64_t lo = out->offset_elems, hi = out->offset_elems;
    for (int r = 0; r < out->rank; r++) {
        int64_t span = (out->extent[r] - 1) * out->stride_elems[r];
        if (span > 0) hi += span; else lo += span;
    }
    const size_t esz = tnb_dtype_size(out->dtype);
    bool dense = true;
    {
        int64_t n = 1;
        for (int r = 0; r < out->rank; r++) n *= out->extent[r];
        dense = (hi - lo + 1) == n;
    }
    int acc = 0;
    if (dense) {
        rc = tnb_memset_zero(ctx, out->buf, (size_t)lo * esz, (size_t)(hi - lo + 1) * esz);
        acc = 1;
    }
    if (!rc) rc = tnb_plan_execute(ctx, P, slice_begin, slice_step, slice_end, acc);
    tnb_plan_destroy(ctx, P);
    return rc;
}

}  // extern "C"
