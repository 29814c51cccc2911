// kernels_stem.cu — the HBM-bound streaming kernel for "stem" steps of a sliced contraction tree:
//
//   C[m, n] = alpha * sum_k A[m, k] * B[n, k] (+ beta * C)      M huge (>= 2^16), N <= 16, K <= 64
//
// i.e. a huge dense intermediate absorbing a tiny tensor (a gate, a small environment) — the shape that dominates the
// BYTES of a good contraction path (profiles/r1_steps_sycamore_opt.md: arithmetic intensity 2-16 flop/B, far below
// the ridge).  What it replaces: Muscle.binary_einsum's permutedims + gemm for these calls
// (/root/reference/src/Operations/overlap.jl:12 -> contract -> binary_einsum), where the permute
// passes alone triple the traffic.  Here each operand element is read once and each result element written once:
//
//   * A is dense [K][M] (m fastest, the planner's layout): a thread owns VM consecutive m and issues one 16-byte load
//     per k — K independent loads in flight per thread, fully coalesced across the warp;
//   * B (<= 1024 elements) is gathered once per CTA into shared memory and read by broadcast;
//   * the N results of each m go to a shared-memory tile at the RANK their address has inside the tile's output
//     pattern (pos table); the tile is then written to C in ascending address order (rel table, or plain
//     contiguous when the planner found the pattern to be one block) — consecutive lanes write consecutive
//     addresses whatever the consumer's layout is.
// Persistent grid (a few CTAs per SM, tiles strided over CTAs).  Algorithmic bytes per launch: sizeof(T)*(M*K + M*N).
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>
#include "tnb_internal.h"

namespace {

constexpr int ST_THREADS = 256;

__device__ __forceinline__ int64_t stab(const TabRef& t, uint32_t i) {
    uint32_t q = i / t.lo_size;
    uint32_t r = i - q * t.lo_size;
    return t.hi[q] + t.lo[r];
}

__device__ __forceinline__ void cmac(float2& c, float2 a, float2 b) {
    c.x = fmaf(a.x, b.x, c.x); c.x = fmaf(-a.y, b.y, c.x);
    c.y = fmaf(a.x, b.y, c.y); c.y = fmaf(a.y, b.x, c.y);
}
__device__ __forceinline__ void cmac(double2& c, double2 a, double2 b) {
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
__device__ __forceinline__ float2 cscale(const double* s, float2 a) {
    float sr = (float)s[0], si = (float)s[1];
    return make_float2(sr * a.x - si * a.y, sr * a.y + si * a.x);
}
__device__ __forceinline__ double2 cscale(const double* s, double2 a) {
    return make_double2(s[0] * a.x - s[1] * a.y, s[0] * a.y + s[1] * a.x);
}
__device__ __forceinline__ float2 czero(float2*) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ double2 czero(double2*) { return make_double2(0., 0.); }

// E = float2 (VM = 2: one 16-byte load covers two m) or double2 (VM = 1)
// Occupancy: the kernel is a latency machine (load phase, rank-ordered staging, write-out phase per tile), so the bytes in
// flight per SM are (resident CTAs) x 256 threads x (loads in flight per thread) x 16 B.  ncu r2 on configs[4]'s
// 1024^2 x 6 x 6 complex128 step: 104 registers -> 2 CTAs per SM, 21 % of DRAM throughput.  The accumulators need
// sizeof(E)/4 * VM * NMAX registers; with <= 16 of them the kernel is held to 64 registers (4 CTAs per SM), with <= 32 to 80 (3 CTAs).
template <typename E, int VM, int NMAX>
__global__ void __launch_bounds__(ST_THREADS, (sizeof(E) / 4 * VM * NMAX <= 16) ? 4 : ((sizeof(E) / 4 * VM * NMAX <= 32) ? 3 : 2)) stem_kernel(const StemArgs p) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int N = p.N, K = p.K, TM = p.TM;
    E* Bs = reinterpret_cast<E*>(st_smem);                                   // [K][N]
    E* tile = Bs + ((K * N + 1) & ~1);                                       // [TM*N] in output-rank order
    uint16_t* pos16 = reinterpret_cast<uint16_t*>(tile + TM * N);            // [TM*N]
    const int tid = threadIdx.x;

    for (int i = tid; i < K * N; i += (int)blockDim.x) {
        const int n = i % N, k = i / N;
        E v = reinterpret_cast<const E*>(p.B)[stab(p.bn, n) + stab(p.bk, k)];
        if (p.conjB) v.y = -v.y;
        Bs[k * N + n] = v;
    }
    // rank table transposed to [n][ml] (conflict-free for consecutive ml) and run bases of the sorted pattern
    for (int i = tid; i < TM * N; i += (int)blockDim.x) {
        const int ml = i / N, n = i - ml * N;
        pos16[n * TM + ml] = (uint16_t)p.pos[i];
    }
    int run_shift = 0;
    while ((1 << (run_shift + 1)) <= p.run) run_shift++;
    const int rmask = (1 << run_shift) - 1;
    const int nruns = (TM * N) >> run_shift;
    int64_t* runbase = reinterpret_cast<int64_t*>(pos16 + ((TM * N + 3) & ~3));     // [nruns] (launcher sized smem for it)
    for (int i = tid; i < nruns; i += (int)blockDim.x) runbase[i] = p.rel[(int64_t)i << run_shift];
    __syncthreads();

    const E* __restrict__ A = reinterpret_cast<const E*>(p.A);
    E* __restrict__ C = reinterpret_cast<E*>(p.C);
    const bool has_beta = (p.beta[0] != 0.0) || (p.beta[1] != 0.0);
    const bool unit_alpha = p.alpha[0] == 1.0 && p.alpha[1] == 0.0;
    const int64_t ntiles = p.M / TM;
    const int cnt = TM * N;
    struct __align__(16) Vec { E v[VM]; };

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t m0 = t * TM;
        for (int ml = tid * VM; ml < TM; ml += (int)blockDim.x * VM) {
            E acc[VM][NMAX];
#pragma unroll
            for (int v = 0; v < VM; v++)
#pragma unroll
                for (int n = 0; n < NMAX; n++) acc[v][n] = czero((E*)0);
            const E* src = A + m0 + ml;
#pragma unroll 4
            for (int k = 0; k < K; k++) {
                Vec a = *reinterpret_cast<const Vec*>(src + (int64_t)k * p.lda);
                if (p.conjA) {
#pragma unroll
                    for (int v = 0; v < VM; v++) a.v[v].y = -a.v[v].y;
                }
                const E* brow = Bs + k * N;
#pragma unroll
                for (int n = 0; n < NMAX; n++) {
                    if (n < N) {
                        const E b = brow[n];
#pragma unroll
                        for (int v = 0; v < VM; v++) cmac(acc[v][n], a.v[v], b);
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < VM; v++)
#pragma unroll
                for (int n = 0; n < NMAX; n++)
                    if (n < N) tile[pos16[n * TM + ml + v]] = acc[v][n];
        }
        __syncthreads();
        E* base = C + p.hi[t];
        for (int j = tid; j < cnt; j += (int)blockDim.x) {
            E* dst = base + runbase[j >> run_shift] + (j & rmask);
            E v = tile[j];
            if (!unit_alpha) v = cscale(p.alpha, v);
            if (has_beta) {
                E o = cscale(p.beta, *dst);
                v.x += o.x; v.y += o.y;
            }
            *dst = v;
        }
        __syncthreads();
    }
}

// Direct form (planner: st_direct — a warp's 32 consecutive rows of one column are whole 64-byte pieces of the output):
// no staging tile, no rank tables, no barrier after the set-up.  A thread loads its VM rows of A (K independent 16-byte
// loads), multiplies with the broadcast small operand and stores its N results straight to C at
// hi[tile] + rowaddr[ml] + coldelta[n]; every warp free-runs over its tiles, so loads, FMAs and stores of different warps
// overlap instead of alternating CTA-wide between a load phase and a write-out phase (ncu r2 on the complex128
// 1024^2 x 6 x 6 step of configs[4]: the staged form was latency-bound at 0.35 - 0.43 of HBM, 54 % of the stalls on the
// consumers of the global loads, 9 % on the two barriers per tile).  PAIR (VM = 2 only): rows 2i, 2i+1 are adjacent in the
// output, both results of a column go out as one 16-byte store.
template <typename E, int VM, int NMAX, bool PAIR>
__global__ void __launch_bounds__(ST_THREADS, (sizeof(E) / 4 * VM * NMAX <= 16) ? 4 : ((sizeof(E) / 4 * VM * NMAX <= 32) ? 3 : 2)) stem_direct_kernel(const StemArgs p) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int N = p.N, K = p.K, TM = p.TM;
    E* Bs = reinterpret_cast<E*>(st_smem);                                   // [K][N]
    int32_t* coldelta = reinterpret_cast<int32_t*>(Bs + ((K * N + 1) & ~1)); // [NMAX]
    int32_t* rowaddr = coldelta + NMAX;                                      // [TM]
    const int tid = threadIdx.x;
    for (int i = tid; i < K * N; i += (int)blockDim.x) {
        const int n = i % N, k = i / N;
        E v = reinterpret_cast<const E*>(p.B)[stab(p.bn, n) + stab(p.bk, k)];
        if (p.conjB) v.y = -v.y;
        Bs[k * N + n] = v;
    }
    {
        const int64_t a0 = p.rel[p.pos[0]];
        for (int i = tid; i < NMAX; i += (int)blockDim.x) coldelta[i] = i < N ? (int32_t)(p.rel[p.pos[i]] - a0) : 0;
        for (int i = tid; i < TM; i += (int)blockDim.x) rowaddr[i] = (int32_t)p.rel[p.pos[(int64_t)i * N]];
    }
    __syncthreads();

    const E* __restrict__ A = reinterpret_cast<const E*>(p.A);
    E* __restrict__ C = reinterpret_cast<E*>(p.C);
    const bool has_beta = (p.beta[0] != 0.0) || (p.beta[1] != 0.0);
    const bool unit_alpha = p.alpha[0] == 1.0 && p.alpha[1] == 0.0;
    const int64_t ntiles = p.M / TM;
    struct __align__(16) Vec { E v[VM]; };

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t m0 = t * TM;
        E* base = C + p.hi[t];
        for (int ml = tid * VM; ml < TM; ml += (int)blockDim.x * VM) {
            E acc[VM][NMAX];
#pragma unroll
            for (int v = 0; v < VM; v++)
#pragma unroll
                for (int n = 0; n < NMAX; n++) acc[v][n] = czero((E*)0);
            const E* src = A + m0 + ml;
            if (K <= 8) {
                // all K loads of the row in flight before the first FMA (the kernel is latency-bound: bytes in flight per
                // SM = resident warps x 32 x K x 16 B)
                Vec ar[8];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k < K) ar[k] = *reinterpret_cast<const Vec*>(src + (int64_t)k * p.lda);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (k < K) {
                        if (p.conjA) {
#pragma unroll
                            for (int v = 0; v < VM; v++) ar[k].v[v].y = -ar[k].v[v].y;
                        }
                        const E* brow = Bs + k * N;
#pragma unroll
                        for (int n = 0; n < NMAX; n++) {
                            if (n < N) {
                                const E b = brow[n];
#pragma unroll
                                for (int v = 0; v < VM; v++) cmac(acc[v][n], ar[k].v[v], b);
                            }
                        }
                    }
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < K; k++) {
                    Vec a = *reinterpret_cast<const Vec*>(src + (int64_t)k * p.lda);
                    if (p.conjA) {
#pragma unroll
                        for (int v = 0; v < VM; v++) a.v[v].y = -a.v[v].y;
                    }
                    const E* brow = Bs + k * N;
#pragma unroll
                    for (int n = 0; n < NMAX; n++) {
                        if (n < N) {
                            const E b = brow[n];
#pragma unroll
                            for (int v = 0; v < VM; v++) cmac(acc[v][n], a.v[v], b);
                        }
                    }
                }
            }
            int32_t ra[VM];
#pragma unroll
            for (int v = 0; v < VM; v++) ra[v] = rowaddr[ml + v];
#pragma unroll
            for (int n = 0; n < NMAX; n++) {
                if (n < N) {
                    const int32_t cd = coldelta[n];
                    if (PAIR) {
                        Vec* dst = reinterpret_cast<Vec*>(base + ra[0] + cd);
                        Vec o;
#pragma unroll
                        for (int v = 0; v < VM; v++) o.v[v] = unit_alpha ? acc[v][n] : cscale(p.alpha, acc[v][n]);
                        if (has_beta) {
                            const Vec old = *dst;
#pragma unroll
                            for (int v = 0; v < VM; v++) { const E q = cscale(p.beta, old.v[v]); o.v[v].x += q.x; o.v[v].y += q.y; }
                        }
                        *dst = o;
                    } else {
#pragma unroll
                        for (int v = 0; v < VM; v++) {
                            E* dst = base + ra[v] + cd;
                            E o = unit_alpha ? acc[v][n] : cscale(p.alpha, acc[v][n]);
                            if (has_beta) { const E q = cscale(p.beta, *dst); o.x += q.x; o.y += q.y; }
                            *dst = o;
                        }
                    }
                }
            }
        }
    }
}

template <typename E, int VM>
int launch(tnb_ctx* ctx, const StemArgs& a) {
    const size_t esz = sizeof(E);
    const size_t cnt = (size_t)a.TM * a.N;
    const size_t smem = (((size_t)a.K * a.N + 1) & ~(size_t)1) * esz + cnt * esz + ((cnt + 3) & ~(size_t)3) * 2 +
                        (cnt / (size_t)(a.run > 0 ? a.run : 1) + 1) * 8;
    const int64_t ntiles = a.M / a.TM;
    // every thread owns VM consecutive m of a tile: no idle lanes in the compute phase
    int threads = (int)(a.TM / VM);
    if (threads > ST_THREADS) threads = ST_THREADS;
    threads = (threads + 31) / 32 * 32;
    int64_t per_sm = smem > 0 ? (int64_t)(200 * 1024 / smem) : 8;
    const int64_t by_threads = 2048 / threads;
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    // persistent grid = what is actually RESIDENT (registers and shared memory decide): a grid sized from shared memory
    // alone left CTAs queued behind the resident ones and the kernel ran in 2.3 "waves" with an idle tail (r2)
#define ST_LAUNCH(NMAX)                                                                                            \
    do {                                                                                                           \
        if (smem > 48 * 1024)                                                                                      \
            TNB_CUDA_CHECK(ctx, cudaFuncSetAttribute(stem_kernel<E, VM, NMAX>,                                     \
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        int resident = 0;                                                                                          \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, stem_kernel<E, VM, NMAX>, threads, smem) ==   \
                cudaSuccess && resident >= 1 && resident < per_sm)                                                 \
            per_sm = resident;                                                                                     \
        int64_t grid = (int64_t)ctx->sm_count * per_sm;                                                            \
        if (grid > ntiles) grid = ntiles;                                                                          \
        if (grid < 1) return TNB_OK;                                                                               \
        stem_kernel<E, VM, NMAX><<<(unsigned)grid, threads, smem, ctx->stream>>>(a);                            \
    } while (0)
    // direct form: the set-up tables replace the staging tile (TNB_STEM_DIRECT is honoured at plan time; C must be 16-byte
    // aligned for the paired stores)
#define ST_LAUNCH_DIRECT(NMAX, PAIR_)                                                                              \
    do {                                                                                                           \
        const size_t dsmem = (((size_t)a.K * a.N + 1) & ~(size_t)1) * esz + ((size_t)NMAX + (size_t)a.TM) * 4;     \
        if (dsmem > 48 * 1024)                                                                                     \
            TNB_CUDA_CHECK(ctx, cudaFuncSetAttribute(stem_direct_kernel<E, VM, NMAX, PAIR_>,                       \
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));   \
        int resident = 0;                                                                                          \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, stem_direct_kernel<E, VM, NMAX, PAIR_>,       \
                                                          threads, dsmem) != cudaSuccess || resident < 1)          \
            resident = 1;                                                                                          \
        int64_t grid = (int64_t)ctx->sm_count * resident;                                                          \
        if (grid > ntiles) grid = ntiles;                                                                          \
        if (grid < 1) return TNB_OK;                                                                               \
        stem_direct_kernel<E, VM, NMAX, PAIR_><<<(unsigned)grid, threads, dsmem, ctx->stream>>>(a);             \
    } while (0)
    static const int simt_direct = [] { const char* e = getenv("TNB_STEM_SIMT_DIRECT"); return e ? atoi(e) : 1; }();   // 0: staged form only (comparison)
    // direct form only while the accumulators leave room for 3 CTAs per SM (<= 32 registers of them): with N = 16 complex128
    // it needs 128 registers and measured slower than the staged form (3.20 vs 3.96 TB/s, profiles/r2_simt_direct_probe.txt)
    const int nmax_ = a.N <= 4 ? 4 : (a.N <= 8 ? 8 : 16);
    if (a.direct && simt_direct && sizeof(E) / 4 * VM * nmax_ <= 32) {
        const bool pair = VM == 2 && a.pairs && ((uintptr_t)a.C % 16) == 0;
        if (pair) {
            if (a.N <= 4) ST_LAUNCH_DIRECT(4, (VM == 2));
            else if (a.N <= 8) ST_LAUNCH_DIRECT(8, (VM == 2));
            else ST_LAUNCH_DIRECT(16, (VM == 2));
        } else {
            if (a.N <= 4) ST_LAUNCH_DIRECT(4, false);
            else if (a.N <= 8) ST_LAUNCH_DIRECT(8, false);
            else ST_LAUNCH_DIRECT(16, false);
        }
        ctx->launches++;
        TNB_CUDA_CHECK(ctx, cudaGetLastError());
        return TNB_OK;
    }
#undef ST_LAUNCH_DIRECT
    if (a.N <= 4) ST_LAUNCH(4);
    else if (a.N <= 8) ST_LAUNCH(8);
    else ST_LAUNCH(16);
#undef ST_LAUNCH
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}

}  // namespace

int tnb_launch_stem(tnb_ctx* ctx, int dtype, const StemArgs& a) {
    if (dtype == TNB_C64) {
        // 16-byte loads need an even leading dimension and a 16-byte aligned base; otherwise one m per thread
        if ((a.lda % 2) == 0 && ((uintptr_t)a.A % 16) == 0) return launch<float2, 2>(ctx, a);
        return launch<float2, 1>(ctx, a);
    }
    if (dtype == TNB_C128) return launch<double2, 1>(ctx, a);
    return tnb_set_error(ctx, TNB_EUNSUPPORTED, "stem kernel: unsupported dtype %d", dtype);
}
