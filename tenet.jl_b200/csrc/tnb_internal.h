// tnb_internal.h — shared internal declarations of libtnb200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/tnb200.h"

// ------------------------------------------------------------------------------------------
// Offset tables.  Every operand of a pairwise step is addressed as
//     base + T_l(l) + T_x(m or n) + T_k(k)
// where each T maps a linear index of one index *group* (batch L, free M / N, contracted K) to an
// element offset.  Offsets are additive over modes, so a group's table factorises into a "lo"
// table over a prefix of its modes and a "hi" table over the rest:
//     T(i) = hi[i / lo_size] + lo[i % lo_size]
// which keeps tables small (<= a few thousand + size/lo_size entries) for 2^30-element operands.
// ------------------------------------------------------------------------------------------
struct TabRef {
    const int64_t* lo;
    const int64_t* hi;
    uint32_t lo_size;   // >= 1
    uint32_t affine;    // 1: T(i) == i*stride exactly (fast paths may skip the table)
    int64_t stride;     // valid when affine
};

struct EinsumArgs {
    const void* A;
    const void* B;
    void* C;
    int64_t M, N, K, L;
    TabRef am, ak, al;
    TabRef bn, bk, bl;
    TabRef cm, cn, cl;
    int32_t conjA, conjB;
    int32_t a_kfast, b_kfast;   // 1: consecutive k are (more) contiguous than consecutive m/n
    double alpha[2], beta[2];
    // split-K: gridDim.z = L*splitk; partials go to ws[((z*N + n)*M + m)] and are reduced later
    int32_t splitk;
    int32_t pad_;
    int64_t kchunk;
    void* ws;
};

// host-side mirror of one group table
struct HostTable {
    int64_t size = 1;
    int64_t lo_size = 1;
    std::vector<int64_t> lo, hi;
    bool affine = true;
    int64_t stride = 0;
    // byte offsets inside the plan's device table blob (filled when uploaded)
    size_t lo_pos = 0, hi_pos = 0;
    int64_t at(int64_t i) const { return hi[i / lo_size] + lo[i % lo_size]; }
};

struct StepSpec {
    int32_t a_id = -1, b_id = -1, c_id = -1;
    int64_t M = 1, N = 1, K = 1, L = 1;
    HostTable am, ak, al, bn, bk, bl, cm, cn, cl;
    int32_t conjA = 0, conjB = 0;
    int32_t a_kfast = 0, b_kfast = 0;
    int32_t kernel = TNB_KERNEL_GENERIC;
    int32_t splitk = 1;
    int64_t kchunk = 0;
    bool hoisted = false;
    double flops = 0, bytes = 0;
    int64_t a_elems = 0, b_elems = 0, c_elems = 0;
    // fast-kernel eligibility facts (filled by the planner)
    bool a_mmajor = false;   // A[m + M*k] exactly (dense, M fastest), no conj needed handled separately
    bool b_nmajor = false;
    // streaming "stem" kernel (huge dense operand x tiny operand): tile-invariant sorted output pattern
    bool st_ok = false, st_swap = false, st_contig = false, st_tc = false;
    int32_t st_npass = 1, st_ncol = 0;   // passes over the small operand's columns, columns per pass
    bool st_additive = false, st_even = false;   // see StemArgs
    bool st_direct = false;                      // stem kernels may store rows straight from registers (planner.cpp)
    bool st_pairs = false;                       // ... and rows 2i, 2i+1 are adjacent in the output (16-byte stores of two complex64)
    bool st_rel_small = false;                   // every entry of st_rel fits an int32 (run bases kept in shared memory)
    int32_t st_tm = 0, st_run = 1;   // tile rows; length of the contiguous output runs inside a tile (power of two)
    std::vector<int64_t> st_hi, st_rel, st_pos;
    size_t st_hi_pos = 0, st_rel_pos = 0, st_pos_pos = 0;
    bool kred = false;       // STREAM kernel variant: dense k-reduction kernel (small M x N, huge K)
    int32_t dmma_small = 0;  // FP64 tensor-core kernel: 64 x 32 tiles (the 128 x 64 grid cannot fill the machine)
    int32_t tc_nt = 0;       // tcgen05 kernel: N tile (256/128), 0 = not used
    bool tc_swap = false;    // tcgen05 kernel: operands swapped (C^T = B A^T)
};

struct tnb_buf {
    void* ptr = nullptr;
    size_t bytes = 0;      // requested
    size_t cap = 0;        // block capacity (size class)
};

struct NcclApi;  // dlopen'ed entry points (comm.cu)

struct tnb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::string err;
    // caching allocator
    std::multimap<size_t, void*> free_blocks;
    size_t in_use = 0, cached = 0, peak = 0;
    int64_t launches = 0;
    int sm_count = 148;
    int c64_mode = TNB_C64_TF32X3;
    int force_generic = 0;
    int use_graphs = 1;      // replay un-sliced plans as one CUDA graph (TNB_OPT_CUDA_GRAPH)
    int gemm_pair = 1;       // c64 GEMM steps on CTA pairs (TNB_OPT_GEMM_PAIR; env TNB_GEMM_PAIR=0 sets the default off)
    int last_kernel = -1;    // kernel id chosen by the most recent tnb_binary_einsum (introspection)
    // comm
    NcclApi* nccl = nullptr;
    void* comm = nullptr;
    int rank = 0, nranks = 1;
    void* multi = nullptr;   // tnb_multi: worker contexts + communicator of tnb_multi_contract_path (multi.cu)
};
void tnb_multi_release(tnb_ctx* ctx);

int tnb_set_error(tnb_ctx* ctx, int code, const char* fmt, ...);
#define TNB_CUDA_CHECK(ctx, expr)                                                              \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return tnb_set_error(ctx, e__ == cudaErrorMemoryAllocation ? TNB_ENOMEM : TNB_ECUDA, \
                                 "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
                                 __FILE__, __LINE__);                                           \
    } while (0)

static inline size_t tnb_dtype_size(int dtype) {
    switch (dtype) {
        case TNB_C128: return 16;
        case TNB_C64: return 8;
        case TNB_F64: return 8;
        case TNB_F32: return 4;
        default: return 0;
    }
}
static inline bool tnb_dtype_complex(int dtype) { return dtype == TNB_C128 || dtype == TNB_C64; }

// kernels_generic.cu
int tnb_launch_einsum_generic(tnb_ctx* ctx, int dtype, const EinsumArgs& args);
int tnb_launch_splitk_reduce(tnb_ctx* ctx, int dtype, const EinsumArgs& args);
// decide split-K factor for the generic kernel; returns splitk (>=1) and sets kchunk, ws elems needed
int tnb_choose_splitk(const tnb_ctx* ctx, int64_t M, int64_t N, int64_t K, int64_t L, int64_t* kchunk,
                      int64_t* ws_elems);

int tnb_choose_thin(const tnb_ctx* ctx, int64_t M, int64_t N, int64_t K, int64_t L, int64_t* kchunk, int64_t* ws_elems);
int tnb_launch_einsum_thin(tnb_ctx* ctx, int dtype, const EinsumArgs& args);
int tnb_choose_kred(const tnb_ctx* ctx, int dtype, int64_t M, int64_t N, int64_t K, int64_t L, bool a_mmajor, bool b_nmajor,
                    int64_t* kchunk, int64_t* ws_elems);
int tnb_launch_einsum_kred(tnb_ctx* ctx, int dtype, const EinsumArgs& args);   // -1: operands not 16-byte aligned
int tnb_launch_zero_strided(tnb_ctx* ctx, int dtype, void* base, int rank, const int64_t* ext, const int64_t* stride);
// kernels_c64_tc.cu
int tnb_tc_c64_tile(int64_t M, int64_t N, int64_t K, int64_t L, bool a_mmajor, bool b_nmajor);
int tnb_launch_c64_tc(tnb_ctx* ctx, const EinsumArgs& e, int nt, int64_t lda, int64_t ldb, bool chunked);
int tnb_tc_c64_splitk(const tnb_ctx* ctx, int64_t M, int64_t N, int64_t K, int64_t* kb_per_split, int64_t* ws_elems);

// kernels_stem.cu
struct StemArgs {
    const void* A;            // big operand, dense [K][M] (m fastest)
    const void* B;            // tiny operand, gathered through bn/bk
    void* C;
    int64_t M, lda;
    int32_t N, K, TM, contig, conjA, conjB;
    int32_t n0;               // first column of the small operand handled by this launch (multi-pass)
    int32_t run;              // contiguous run length of the sorted pattern (power of two): rel[j] = rel[j & ~(run-1)] + (j & (run-1))
    int32_t additive;         // pos[ml*N + n] == pos[ml*N] | (pos[n] - pos[0]), disjoint bits (rank separable in row and column)
    int32_t even;             // every tile base (hi) and every run base (rel[j*run]) is even: 16-byte aligned pairs
    int32_t direct;           // a warp's 32 rows of one column are whole 64-byte pieces of the output: no staging tile needed
    int32_t pairs;            // direct && rows 2i, 2i+1 adjacent in the output (SIMT complex64 form: 16-byte stores)
    TabRef bn, bk;
    const int64_t* hi;        // [M/TM] tile base offsets in C
    const int64_t* rel;       // [TM*N] ascending offsets inside a tile
    const int64_t* pos;       // [TM*N] (ml*N+n) -> rank in rel
    double alpha[2], beta[2];
};
int tnb_launch_stem(tnb_ctx* ctx, int dtype, const StemArgs& a);
bool tnb_stem_tc_shape_ok(int64_t Mbig, int64_t Nsmall, int64_t K);
int tnb_launch_c64_stem_tc(tnb_ctx* ctx, const StemArgs& e, void* ws, int npass);
int tnb_launch_c64_pair_staged(tnb_ctx* ctx, const StemArgs& e, int64_t Nsmall, int64_t ldb, bool rel_small);   // -1: not eligible
int64_t tnb_stem_tc_ws_elems(int64_t Nsmall, int64_t K, int64_t npass);

// kernels_c128_dmma.cu
int tnb_choose_splitk_dmma(const tnb_ctx* ctx, int64_t M, int64_t N, int64_t K, int64_t L, int64_t* kchunk, int64_t* ws_elems,
                           int32_t* small_tiles);
int tnb_launch_c128_dmma(tnb_ctx* ctx, const EinsumArgs& a);

// planner.cpp  (pure host code; also used by the dry-run plan that CPU tests inspect)
struct PlanTensor {
    std::vector<int32_t> modes;
    std::vector<int64_t> ext, stride;
    int32_t conj = 0;
    int64_t offset = 0;
};
// Build the GEMM view of one pairwise step from three fully specified operand layouts.
// sum: modes to be summed. Returns 0 or TNB_EINVAL with msg.
int tnb_build_step(const PlanTensor& A, const PlanTensor& B, const PlanTensor& C,
                   const std::vector<int32_t>& sum, bool cplx, size_t elem_size, StepSpec* out,
                   std::string* msg);

struct PlanNode {
    bool leaf = false;
    int a = -1, b = -1, consumer = -1;
    PlanTensor t;            // layout (leaves: user layout minus sliced modes)
    bool dep = false;        // depends on the slice id
    int64_t elems = 1;
    int64_t arena_off = -1;  // element offset inside the arena (intermediates)
    std::vector<int64_t> slice_stride;  // leaves: stride of each sliced mode (0 if absent)
};

struct tnb_plan {
    int dtype = 0;
    int nleaves = 0, nsteps = 0;
    std::vector<PlanNode> nodes;
    std::vector<StepSpec> steps;          // SSA order
    std::vector<int> order_hoisted, order_dep;
    std::vector<int32_t> sliced_modes;
    std::vector<int64_t> sliced_ext;
    int64_t nslices = 1;
    tnb_plan_info info{};
    // leaf / out storage captured at creation
    std::vector<tnb_buf*> leaf_buf;
    std::vector<int64_t> leaf_off;
    tnb_buf* out_buf = nullptr;
    int64_t out_off = 0;
    // device resources
    bool dry = false;
    tnb_buf* arena = nullptr;
    tnb_buf* tables = nullptr;
    tnb_buf* ws = nullptr;
    int64_t arena_elems = 0, ws_elems = 0;
    std::vector<int64_t> table_blob;      // host copy of all tables
    // optional per-step device timing (CUDA events on the context stream)
    bool profiling = false;
    std::vector<cudaEvent_t> ev;          // 2 per step
    std::vector<double> step_ms;          // accumulated
    std::vector<int64_t> step_runs;
    // CUDA graph of an UN-SLICED path (every step is slice-invariant: same pointers on every execute), one per
    // accumulate flag; captured on the first execute, replayed afterwards (tnb_plan_execute)
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    int64_t glaunches[2] = {0, 0};
    bool graph_failed = false;
};

int tnb_plan_build(const tnb_tensor* leaves, int32_t nleaves, const int32_t* steps, int32_t nsteps,
                   const int32_t* sliced_modes, int32_t nsliced, const tnb_tensor* out, tnb_plan* plan,
                   std::string* msg);
