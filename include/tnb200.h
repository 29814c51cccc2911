/*
 * tnb200.h — C ABI of libtnb200.so, the B200-native tensor-network contraction engine.
 *
 * This is the drop-in boundary for ONE hot path of bsc-quantic/Tenet.jl (v0.10.3):
 * "execute an EinExprs contraction path as a chain of pairwise binary einsums, summed over a set
 * of sliced indices".  Every entry point below names the reference interface it replaces.  The
 * definitions of those interfaces live in Muscle.jl / Tangles.jl / EinExprs.jl (registry deps, not
 * vendored: /root/reference/Project.toml:6-18,29-45), so citations are the *call sites* in the
 * reference tree that fix the semantics.
 *
 * Conventions
 *   - plain C, no C++ exception crosses this boundary, every function returns an int status;
 *     a human-readable message for the last failure is kept per context (tnb_last_error).
 *   - strides are in ELEMENTS (Julia column-major arrays of extents (d1..dN) have strides
 *     (1, d1, d1*d2, ...); SubArray views arrive as offset + non-unit strides).
 *   - "modes" are int32 labels (the host maps Muscle `Index` objects to ints).
 *   - all device work is stream-ordered on the context's single stream; only tnb_download and
 *     tnb_sync block the host.
 *   - there is NO CPU fallback: an unsupported request is an error, never a silent host compute.
 */
#ifndef TNB200_H
#define TNB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define TNB_OK            0
#define TNB_EINVAL        1  /* shape / mode / argument mismatch                        */
#define TNB_ENOMEM        2  /* device (or host) allocation failed                       */
#define TNB_ECUDA         3  /* CUDA runtime / driver error                              */
#define TNB_ENCCL         4  /* NCCL error (or libnccl could not be loaded)              */
#define TNB_EUNSUPPORTED  5  /* dtype / rank / feature outside what the engine implements */

/* element types (Tensor eltypes reaching binary_einsum in the reference: ComplexF64 everywhere,
 * ComplexF32 for circuits, Float64 e.g. test/unit/mps.jl:63-69 after Int promotion) */
#define TNB_C128 0
#define TNB_C64  1
#define TNB_F64  2
#define TNB_F32  3

/* precision policy for ComplexF32 / Float32 tensor-core GEMM steps (tnb_ctx_set_option) */
#define TNB_OPT_C64_MODE     1   /* value: TNB_C64_SIMT | TNB_C64_TF32X3 | TNB_C64_TF32X3_FAST */
#define TNB_OPT_FORCE_KERNEL 2   /* value: 0 auto, 1 generic table kernel only (debug/parity) */
#define TNB_OPT_GEMM_PAIR    3   /* value: 1 (default) c64 GEMM / wide stem steps on CTA pairs (cta_group::2), 0 the 1-CTA kernels */
#define TNB_OPT_CUDA_GRAPH   4   /* value: 1 (default) un-sliced plans are captured once and replayed as ONE CUDA graph, 0 direct launches */
#define TNB_C64_SIMT   0         /* exact FP32 FMA (BLAS-equivalent rounding)                  */
#define TNB_C64_TF32X3 1         /* tcgen05 kind::tf32, hi/lo split; TMEM chunks of 128 k drained into RN fp32 totals (default) */
#define TNB_C64_TF32X3_FAST 2    /* same, whole K chained in TMEM (RZ accumulate bias ~6e-8 * 0.75 K relative) */

#define TNB_MAX_RANK 64

typedef struct tnb_ctx  tnb_ctx;   /* one per (process, device); owns stream + allocator */
typedef struct tnb_buf  tnb_buf;   /* opaque device buffer handle                         */
typedef struct tnb_plan tnb_plan;  /* a planned contraction path (tables + arena layout)  */

/*
 * Tensor descriptor — replaces Muscle's `Tensor(array, inds)` (ctor call sites:
 * src/Components/MPS.jl:83, MPO.jl:120, PEPS.jl:46, ProductState.jl:48,98, src/Models/Ising.jl:29).
 * Borrowed for the duration of the call only.
 */
typedef struct tnb_tensor {
    tnb_buf*       buf;           /* device storage                                         */
    int64_t        offset_elems;  /* first element (views: compress.jl:46-58, evolve.jl:64-72) */
    int32_t        dtype;         /* TNB_C128 ...                                            */
    int32_t        rank;          /* 0 allowed (DMRG.jl:60-61)                               */
    const int64_t* extent;        /* [rank]                                                  */
    const int64_t* stride_elems;  /* [rank]                                                  */
    const int32_t* mode;          /* [rank] labels, unique within one tensor                 */
    int32_t        conj;          /* 1: use conj(element) — removes the materialised copy of */
                                  /*    overlap.jl:7,39 / DMRG.jl:80,96                       */
} tnb_tensor;

/* ---- context ------------------------------------------------------------------------------ */
int tnb_ctx_create(int device, tnb_ctx** out);
int tnb_ctx_destroy(tnb_ctx* ctx);
int tnb_ctx_set_option(tnb_ctx* ctx, int option, int64_t value);
const char* tnb_last_error(tnb_ctx* ctx);       /* ctx may be NULL: last error of ctx_create */
int tnb_sync(tnb_ctx* ctx);
/* raw cudaStream_t of the context (for callers that time with their own CUDA events) */
void* tnb_ctx_stream(tnb_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t tnb_ctx_launch_count(tnb_ctx* ctx);
/* TNB_KERNEL_* id selected by the most recent tnb_binary_einsum on this context (-1: none yet) */
int tnb_ctx_last_kernel(tnb_ctx* ctx);

/* ---- device memory: replaces the Array storage behind `parent(tensor)` ---------------------- */
int tnb_alloc(tnb_ctx* ctx, size_t bytes, tnb_buf** out);   /* stream-ordered caching allocator */
int tnb_free(tnb_ctx* ctx, tnb_buf* buf);                   /* safe from finalizer threads      */
int tnb_upload(tnb_ctx* ctx, tnb_buf* dst, size_t dst_offset_bytes, const void* host, size_t bytes);
int tnb_download(tnb_ctx* ctx, const tnb_buf* src, size_t src_offset_bytes, void* host, size_t bytes);
int tnb_memset_zero(tnb_ctx* ctx, tnb_buf* dst, size_t offset_bytes, size_t bytes);
void*  tnb_buf_ptr(const tnb_buf* buf);     /* raw device pointer (interop: __cuda_array_interface__) */
size_t tnb_buf_bytes(const tnb_buf* buf);
/* allocator statistics: bytes currently handed out, bytes cached, high-water mark */
int tnb_mem_stats(tnb_ctx* ctx, size_t* in_use, size_t* cached, size_t* peak);
int tnb_mem_trim(tnb_ctx* ctx);             /* return cached blocks to the driver */

/* ---- hot path 1: one pairwise contraction ---------------------------------------------------
 * Replaces Muscle.binary_einsum(a, b; dims, out) — 49 call sites, e.g. src/Operations/overlap.jl:42,46;
 * src/Algorithms/DMRG.jl:17; dims=Index[] (Hadamard/batch) at canonize.jl:44, absorb.jl:31.
 *
 *   C[modes(C)] = alpha * sum_{sum_modes} op(A)[modes(A)] * op(B)[modes(B)]  +  beta * C
 *
 * A mode present in A and B and listed in sum_modes is contracted; present in both and NOT listed is
 * a batch mode and must appear in C; a mode in exactly one operand must appear in C or in sum_modes
 * (then it is summed out).  alpha/beta point to one element of C's dtype (NULL: 1 and 0).
 */
int tnb_binary_einsum(tnb_ctx* ctx, const tnb_tensor* A, const tnb_tensor* B, const tnb_tensor* C,
                      const int32_t* sum_modes, int32_t nsum, const void* alpha, const void* beta);

/* Default result layout of binary_einsum when the caller has no preference: free(A) in A's order,
 * free(B) in B's order, then batch modes; writes rank / modes / extents (arrays of TNB_MAX_RANK). */
int tnb_binary_einsum_result(const tnb_tensor* A, const tnb_tensor* B, const int32_t* sum_modes,
                             int32_t nsum, int32_t* out_rank, int32_t* out_modes, int64_t* out_extents);

/* ---- hot path 2: a whole contraction path, summed over sliced modes --------------------------
 * Replaces Tangles.contract(tn; path) (call sites: src/Operations/overlap.jl:12, test/unit/mps.jl:89…520,
 * test/integration/itensormps.jl:45) driven by an EinExprs path (README.md:19-20).
 *
 * SSA path: step s contracts ids steps[2s], steps[2s+1] into id nleaves+s; ids < nleaves are leaves.
 * A mode is summed at the step where its last two carriers meet, unless it is a mode of `out`.
 * Slices are enumerated in mixed radix over sliced_modes (sliced_modes[0] fastest); slice ids
 * slice_begin, slice_begin+slice_step, ... < slice_end are contracted and ADDED into `out`
 * (out = beta_out*out + sum of slices; beta_out 0 or 1 via accumulate flag).
 */
int tnb_plan_create(tnb_ctx* ctx, const tnb_tensor* leaves, int32_t nleaves,
                    const int32_t* steps, int32_t nsteps,
                    const int32_t* sliced_modes, int32_t nsliced,
                    const tnb_tensor* out, tnb_plan** plan);
int tnb_plan_execute(tnb_ctx* ctx, tnb_plan* plan, int64_t slice_begin, int64_t slice_step,
                     int64_t slice_end, int32_t accumulate);
int tnb_plan_destroy(tnb_ctx* ctx, tnb_plan* plan);

/* plan introspection: what bench.py and the roofline accounting read */
typedef struct tnb_plan_info {
    int64_t nslices;              /* product of sliced extents                               */
    int64_t nsteps_per_slice;     /* pairwise steps that depend on the slice id              */
    int64_t nsteps_hoisted;       /* slice-invariant steps (run once per execute)            */
    double  flops_per_slice;      /* 8*MACs (complex) or 2*MACs (real) of slice-dependent steps */
    double  flops_hoisted;
    double  bytes_per_slice;      /* sizeof(T)*(|A|+|B|+|C|) summed over slice-dependent steps */
    double  bytes_hoisted;
    int64_t workspace_bytes;      /* arena size for intermediates                            */
    int64_t table_bytes;          /* device offset tables                                    */
    int64_t max_intermediate_elems;
} tnb_plan_info;
int tnb_plan_get_info(const tnb_plan* plan, tnb_plan_info* info);

/* per-step record for profiling (step index in SSA order) */
typedef struct tnb_step_info {
    int64_t M, N, K, L;           /* GEMM view extents                                       */
    int32_t kernel;               /* TNB_KERNEL_* actually selected                          */
    int32_t hoisted;
    double  flops, bytes;
} tnb_step_info;
#define TNB_KERNEL_GENERIC   0    /* table-driven SIMT tile kernel                           */
#define TNB_KERNEL_C64_TF32  1    /* tcgen05 3xTF32                                          */
#define TNB_KERNEL_C128_DMMA 2    /* mma.sync f64 tensor-core kernel                         */
#define TNB_KERNEL_STREAM    3    /* memory-bound streaming kernel (small K*N)               */
#define TNB_KERNEL_SPLITK    4    /* split-K reduction (M*N tiny, K huge)                    */
#define TNB_KERNEL_STEM      5    /* HBM-bound streaming kernel: huge dense operand x tiny operand */
#define TNB_KERNEL_STEM_TC   6    /* persistent tcgen05 kernels: huge dense operand x small operand (16..128 columns per pass; 128-column passes with a dense small operand run on CTA pairs) */
int tnb_plan_get_step(const tnb_plan* plan, int32_t step, tnb_step_info* info);

/* per-step device timing (CUDA events on the context stream; one host sync per slice while enabled) */
int tnb_plan_profile(tnb_ctx* ctx, tnb_plan* plan, int32_t enable);
int tnb_plan_get_step_time(const tnb_plan* plan, int32_t step, double* ms_total, int64_t* runs);

/* planning only — no device, no buffers (descriptor buf may be NULL): lets host code and CPU tests inspect the
 * planner's decisions.  A dry plan cannot be executed.  dump_table: which = 0..8 -> am ak al bn bk bl cm cn cl,
 * returns the table size and copies min(size, cap) fully expanded offsets.  dump_step: ids[3] = a,b,c node ids,
 * base[3] = element offset of each operand in its storage, kind[3] = 0 leaf / 1 arena / 2 out,
 * slice_stride[3][nsliced], conj[2]. */
int tnb_plan_create_dry(const tnb_tensor* leaves, int32_t nleaves, const int32_t* steps, int32_t nsteps,
                        const int32_t* sliced_modes, int32_t nsliced, const tnb_tensor* out, tnb_plan** plan);
int64_t tnb_plan_dump_table(const tnb_plan* plan, int32_t step, int32_t which, int64_t* dst, int64_t cap);
int tnb_plan_dump_step(const tnb_plan* plan, int32_t step, int32_t* ids, int64_t* base, int32_t* kind,
                       int64_t* slice_stride, int32_t* conj);

/* one-shot convenience = plan_create + (memset out) + plan_execute + plan_destroy */
int tnb_contract_path(tnb_ctx* ctx, const tnb_tensor* leaves, int32_t nleaves,
                      const int32_t* steps, int32_t nsteps,
                      const int32_t* sliced_modes, int32_t nsliced,
                      int64_t slice_begin, int64_t slice_step, int64_t slice_end,
                      const tnb_tensor* out);

/* ---- device-resident thin QR / SVD of a tensor viewed as a matrix -------------------------
 * Replaces Muscle.tensor_qr_thin(A; inds_q, inds_r, ind_virtual) / tensor_svd_thin(A; inds_u, inds_v, ind_s) behind
 * canonize! / compress! / evolve! / two-site DMRG (src/Operations/canonize.jl:41,58,97; evolve.jl:62,92;
 * src/Algorithms/DMRG.jl:338,437), so that the tensors stay on the device between the einsums of the hot path.
 * Rows of the matrix = row_modes (in that order), columns = the other modes of A in A's order, k = min(m, n):
 *   A = sum_k Q[row modes.., k] R[k, col modes..]              (Q^H Q = 1, R upper triangular in that matrix view)
 *   A = sum_k U[row modes.., k] S[k] Vh[k, col modes..]        (S real >= 0, descending, stored in A's dtype)
 * Q/R/U/S/Vh descriptors give the caller's layouts (any strides / mode order; extents must match; the new mode is
 * `virtual_mode`).  The factorisation itself is cuSOLVER (geqrf + orgqr / gesvd), dlopen'ed at first use: a missing
 * libcusolver is TNB_EUNSUPPORTED, never a CPU fallback.  All four dtypes. */
int tnb_qr_thin(tnb_ctx* ctx, const tnb_tensor* A, const int32_t* row_modes, int32_t nrow, int32_t virtual_mode,
                const tnb_tensor* Q, const tnb_tensor* R);
int tnb_svd_thin(tnb_ctx* ctx, const tnb_tensor* A, const int32_t* row_modes, int32_t nrow, int32_t virtual_mode,
                 const tnb_tensor* U, const tnb_tensor* S, const tnb_tensor* Vh);

/* ---- multi-GPU, single process: tnb_contract_path over `ngpus` devices of this box -------------------------
 * The reference calls `contract` from ONE Julia task, so this is the entry a drop-in `contract(tn; path,
 * ngpus)` binds: leaves and `out` live on ctx's device; the library runs one host thread + one context per additional
 * device (cached in ctx), broadcasts the leaves, deals slice s to rank s mod ngpus, sums the accumulators with one NCCL
 * all-reduce and leaves the result in `out` (which must be a dense view).  ngpus = 1 is tnb_contract_path.
 * Un-sliced paths do not shard (rank 0 contracts, the others contribute zeros). */
int tnb_multi_contract_path(tnb_ctx* ctx, const tnb_tensor* leaves, int32_t nleaves,
                            const int32_t* steps, int32_t nsteps,
                            const int32_t* sliced_modes, int32_t nsliced,
                            const tnb_tensor* out, int32_t ngpus);

/* ---- multi-GPU: one process per GPU, slices dealt round-robin, one all-reduce at the end -------
 * (the reference has no distributed code: README.md:22 only advertises it.)
 * NCCL is dlopen'ed at run time; the 128-byte unique id is created on rank 0 and shipped to the
 * other ranks by the host (torch.distributed store / MPI / a file). */
int tnb_comm_unique_id(void* id128);
int tnb_comm_init(tnb_ctx* ctx, const void* id128, int32_t rank, int32_t nranks);
int tnb_comm_allreduce_sum(tnb_ctx* ctx, tnb_buf* buf, size_t offset_bytes, int64_t count, int32_t dtype);
int32_t tnb_comm_size(const tnb_ctx* ctx);   /* ranks of the communicator bound to ctx; 1 when none */
int tnb_comm_destroy(tnb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TNB200_H */
