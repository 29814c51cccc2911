// kernels_generic.cu — the table-driven SIMT tile kernel: every pairwise step, every dtype.
//
// Replaces the arithmetic of Muscle.binary_einsum (call sites: /root/reference/src/Operations/overlap.jl:42,46,
// src/Algorithms/DMRG.jl:17, canonize.jl:44,61, absorb.jl:31 ...): the reference permutes both operands,
// reshapes and calls BLAS gemm; here operands are gathered straight from their original strides through
// the additive offset tables (tnb_internal.h), multiplied in a 64x64x8 shared-memory tile with a 4x4
// register micro-tile per thread, and scattered into the layout the consumer wants.  It is the
// always-correct path (any rank, strides, batch modes, conj flags, rank-0 operands, alpha/beta) and the
// one the specialised kernels are checked against.
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>
#include "tnb_internal.h"

namespace {

constexpr int TM = 64, TN = 64, BK = 8, NTHREADS = 256;
constexpr int PAD = 4;

template <typename R, bool CPLX> struct ElemT;
template <> struct ElemT<float, true> { typedef float2 type; };
template <> struct ElemT<double, true> { typedef double2 type; };
template <> struct ElemT<float, false> { typedef float type; };
template <> struct ElemT<double, false> { typedef double type; };

__device__ __forceinline__ float2 ezero(float2*) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ double2 ezero(double2*) { return make_double2(0., 0.); }
__device__ __forceinline__ float ezero(float*) { return 0.f; }
__device__ __forceinline__ double ezero(double*) { return 0.; }

__device__ __forceinline__ float2 econj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ double2 econj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ float econj(float a) { return a; }
__device__ __forceinline__ double econj(double a) { return a; }

__device__ __forceinline__ void emac(float2& c, float2 a, float2 b) {
    c.x = fmaf(a.x, b.x, c.x); c.x = fmaf(-a.y, b.y, c.x);
    c.y = fmaf(a.x, b.y, c.y); c.y = fmaf(a.y, b.x, c.y);
}
__device__ __forceinline__ void emac(double2& c, double2 a, double2 b) {
    c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
__device__ __forceinline__ void emac(float& c, float a, float b) { c = fmaf(a, b, c); }
__device__ __forceinline__ void emac(double& c, double a, double b) { c = fma(a, b, c); }

__device__ __forceinline__ float2 eadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 eadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float eadd(float a, float b) { return a + b; }
__device__ __forceinline__ double eadd(double a, double b) { return a + b; }

// scalar (alpha/beta given as double[2]) times element
__device__ __forceinline__ float2 escale(const double* s, float2 a) {
    float sr = (float)s[0], si = (float)s[1];
    return make_float2(sr * a.x - si * a.y, sr * a.y + si * a.x);
}
__device__ __forceinline__ double2 escale(const double* s, double2 a) {
    return make_double2(s[0] * a.x - s[1] * a.y, s[0] * a.y + s[1] * a.x);
}
__device__ __forceinline__ float escale(const double* s, float a) { return (float)s[0] * a; }
__device__ __forceinline__ double escale(const double* s, double a) { return s[0] * a; }

__device__ __forceinline__ int64_t tab(const TabRef& t, uint32_t i) {
    uint32_t q = i / t.lo_size;
    uint32_t r = i - q * t.lo_size;
    return t.hi[q] + t.lo[r];
}

template <typename E>
__device__ __forceinline__ void store_out(E* c, E acc, const double* alpha, const double* beta, bool has_beta) {
    E v = escale(alpha, acc);
    if (has_beta) v = eadd(v, escale(beta, *c));
    *c = v;
}

template <typename R, bool CPLX>
__global__ void __launch_bounds__(NTHREADS) einsum_tile_kernel(const EinsumArgs p) {
    typedef typename ElemT<R, CPLX>::type E;
    __shared__ __align__(16) E As[BK][TM + PAD];
    __shared__ __align__(16) E Bs[BK][TN + PAD];

    const int tid = threadIdx.x;
    const uint32_t tilesM = (uint32_t)((p.M + TM - 1) / TM);
    const uint32_t tilesN = (uint32_t)((p.N + TN - 1) / TN);
    uint32_t bid = blockIdx.x;
    const uint32_t bm = bid % tilesM; bid /= tilesM;
    const uint32_t bn = bid % tilesN; bid /= tilesN;
    const uint32_t ks = bid % (uint32_t)p.splitk;
    const uint32_t l = bid / (uint32_t)p.splitk;
    const uint32_t m0 = bm * TM, n0 = bn * TN;
    const uint32_t M = (uint32_t)p.M, N = (uint32_t)p.N, K = (uint32_t)p.K;
    uint32_t k_begin = 0, k_end = K;
    if (p.splitk > 1) {
        k_begin = (uint32_t)(ks * p.kchunk);
        uint64_t ke = (uint64_t)k_begin + (uint64_t)p.kchunk;
        k_end = ke < K ? (uint32_t)ke : K;
        if (k_begin > k_end) k_begin = k_end;
    }

    const E* __restrict__ A = (const E*)p.A + tab(p.al, l);
    const E* __restrict__ B = (const E*)p.B + tab(p.bl, l);

    // global -> register -> shared load mapping: 2 elements of A and 2 of B per thread and k-tile
    int a_ml[2], a_kl[2], b_nl[2], b_kl[2];
    int64_t a_off[2], b_off[2];
    bool a_ok[2], b_ok[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int e = tid + r * NTHREADS;
        if (p.a_kfast) { a_kl[r] = e % BK; a_ml[r] = e / BK; } else { a_ml[r] = e % TM; a_kl[r] = e / TM; }
        if (p.b_kfast) { b_kl[r] = e % BK; b_nl[r] = e / BK; } else { b_nl[r] = e % TN; b_kl[r] = e / TN; }
        uint32_t m = m0 + a_ml[r], n = n0 + b_nl[r];
        a_ok[r] = m < M; b_ok[r] = n < N;
        a_off[r] = a_ok[r] ? tab(p.am, m) : 0;
        b_off[r] = b_ok[r] ? tab(p.bn, n) : 0;
    }

    E acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = ezero((E*)0);

    const int tm = tid % 16, tn = tid / 16;
    E ra[2], rb[2];
    auto load_tile = [&](uint32_t k0) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            uint32_t ka = k0 + a_kl[r], kb = k0 + b_kl[r];
            E va = ezero((E*)0), vb = ezero((E*)0);
            if (a_ok[r] && ka < k_end) va = __ldg(A + a_off[r] + tab(p.ak, ka));
            if (b_ok[r] && kb < k_end) vb = __ldg(B + b_off[r] + tab(p.bk, kb));
            ra[r] = p.conjA ? econj(va) : va;
            rb[r] = p.conjB ? econj(vb) : vb;
        }
    };

    if (k_begin < k_end) load_tile(k_begin);
    for (uint32_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            As[a_kl[r]][a_ml[r]] = ra[r];
            Bs[b_kl[r]][b_nl[r]] = rb[r];
        }
        __syncthreads();
        if (k0 + BK < k_end) load_tile(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            E a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][i * 16 + tm];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tn * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) emac(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }

    if (p.splitk == 1) {
        E* __restrict__ C = (E*)p.C + tab(p.cl, l);
        const bool has_beta = (p.beta[0] != 0.0) || (p.beta[1] != 0.0);
        int64_t cn_off[4];
        bool n_ok[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t n = n0 + tn * 4 + j;
            n_ok[j] = n < N;
            cn_off[j] = n_ok[j] ? tab(p.cn, n) : 0;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t m = m0 + i * 16 + tm;
            if (m >= M) continue;
            int64_t cm_off = tab(p.cm, m);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (n_ok[j]) store_out(C + cm_off + cn_off[j], acc[i][j], p.alpha, p.beta, has_beta);
        }
    } else {
        E* __restrict__ W = (E*)p.ws;
        const uint64_t z = (uint64_t)l * (uint64_t)p.splitk + ks;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t m = m0 + i * 16 + tm;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t n = n0 + tn * 4 + j;
                if (n < N) W[(z * N + n) * M + m] = acc[i][j];
            }
        }
    }
}

// sums the split-K partials in a fixed order (deterministic) and applies alpha/beta + scatter
template <typename R, bool CPLX>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const EinsumArgs p) {
    typedef typename ElemT<R, CPLX>::type E;
    const uint64_t MN = (uint64_t)p.M * (uint64_t)p.N;
    const uint64_t total = MN * (uint64_t)p.L;
    const bool has_beta = (p.beta[0] != 0.0) || (p.beta[1] != 0.0);
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t l = (uint32_t)(idx / MN);
        uint64_t r = idx - (uint64_t)l * MN;
        uint32_t n = (uint32_t)(r / (uint64_t)p.M);
        uint32_t m = (uint32_t)(r - (uint64_t)n * (uint64_t)p.M);
        const E* W = (const E*)p.ws;
        E s = ezero((E*)0);
        for (int ks = 0; ks < p.splitk; ks++) s = eadd(s, W[((uint64_t)l * p.splitk + ks) * MN + r]);
        E* C = (E*)p.C + tab(p.cl, l) + tab(p.cm, m) + tab(p.cn, n);
        store_out(C, s, p.alpha, p.beta, has_beta);
    }
}

// ---------------------------------------------------------------------------------------------------------
// "thin" reduction kernel: M*N <= 16 (amplitude-closing dot products, tiny environments), K huge, L == 1.
// HBM-bound: every CTA streams a contiguous chunk of K with all threads along k (coalesced when the operand
// is k-fastest, which the planner's [free | contracted] layout guarantees when there are no free modes),
// keeps the MxN partial sums in registers, block-reduces them and writes one partial per CTA; the
// split-K reducer above then sums the partials in a fixed order (deterministic).
// Algorithmic bytes per launch: sizeof(T) * (M + N) * K.
// ---------------------------------------------------------------------------------------------------------
constexpr int THIN_MAX = 4;       // M, N <= 4
constexpr int THIN_THREADS = 256;
constexpr int THIN_UNROLL = 4;

template <typename R, bool CPLX>
__global__ void __launch_bounds__(THIN_THREADS) einsum_thin_kernel(const EinsumArgs p) {
    typedef typename ElemT<R, CPLX>::type E;
    const uint32_t M = (uint32_t)p.M, N = (uint32_t)p.N, K = (uint32_t)p.K;
    const E* __restrict__ A = (const E*)p.A;
    const E* __restrict__ B = (const E*)p.B;
    int64_t am[THIN_MAX], bn[THIN_MAX];
#pragma unroll
    for (int i = 0; i < THIN_MAX; i++) {
        am[i] = i < (int)M ? tab(p.am, i) : 0;
        bn[i] = i < (int)N ? tab(p.bn, i) : 0;
    }
    E acc[THIN_MAX][THIN_MAX];
#pragma unroll
    for (int i = 0; i < THIN_MAX; i++)
#pragma unroll
        for (int j = 0; j < THIN_MAX; j++) acc[i][j] = ezero((E*)0);

    const uint64_t kb = (uint64_t)blockIdx.x * (uint64_t)p.kchunk;
    uint64_t ke64 = kb + (uint64_t)p.kchunk;
    const uint32_t k_begin = kb < K ? (uint32_t)kb : K;
    const uint32_t k_end = ke64 < K ? (uint32_t)ke64 : K;
    const bool aff = p.ak.affine && p.bk.affine;
    uint32_t k_scalar = k_begin;
    // Fast path for the amplitude-closing dot product (M = N = 1, both operands contiguous in k): 32-byte loads
    // (4 complex64 / 2 complex128 per lane and instruction), 4 loads of each operand in flight per thread.
    if (M == 1 && N == 1 && aff && p.ak.stride == 1 && p.bk.stride == 1 &&
        (((uintptr_t)(A + am[0] + k_begin) | (uintptr_t)(B + bn[0] + k_begin)) & 31) == 0) {
        constexpr int VEC = 32 / (int)sizeof(E);
        struct __align__(32) Pack { E v[VEC]; };
        const Pack* __restrict__ Av = reinterpret_cast<const Pack*>(A + am[0] + k_begin);
        const Pack* __restrict__ Bv = reinterpret_cast<const Pack*>(B + bn[0] + k_begin);
        const uint32_t nvec = (k_end - k_begin) / VEC;
        E s0 = ezero((E*)0), s1 = ezero((E*)0);
        uint32_t i = threadIdx.x;
        for (; i + 3 * THIN_THREADS < nvec; i += 4 * THIN_THREADS) {
            Pack pa[4], pb[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { pa[u] = Av[i + u * THIN_THREADS]; pb[u] = Bv[i + u * THIN_THREADS]; }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    E av = p.conjA ? econj(pa[u].v[e]) : pa[u].v[e];
                    E bv = p.conjB ? econj(pb[u].v[e]) : pb[u].v[e];
                    emac((e & 1) ? s1 : s0, av, bv);
                }
        }
        for (; i < nvec; i += THIN_THREADS) {
            Pack pa = Av[i], pb = Bv[i];
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                E av = p.conjA ? econj(pa.v[e]) : pa.v[e];
                E bv = p.conjB ? econj(pb.v[e]) : pb.v[e];
                emac((e & 1) ? s1 : s0, av, bv);
            }
        }
        acc[0][0] = eadd(s0, s1);
        k_scalar = k_begin + nvec * VEC;          // ragged tail handled by the generic loop below
    }
    for (uint32_t k0 = k_scalar + threadIdx.x; k0 < k_end; k0 += THIN_THREADS * THIN_UNROLL) {
        E a[THIN_UNROLL][THIN_MAX], b[THIN_UNROLL][THIN_MAX];
#pragma unroll
        for (int u = 0; u < THIN_UNROLL; u++) {
            uint32_t k = k0 + u * THIN_THREADS;
            bool ok = k < k_end;
            int64_t oa = 0, ob = 0;
            if (ok) {
                if (aff) { oa = (int64_t)k * p.ak.stride; ob = (int64_t)k * p.bk.stride; }
                else { oa = tab(p.ak, k); ob = tab(p.bk, k); }
            }
#pragma unroll
            for (int i = 0; i < THIN_MAX; i++) {
                a[u][i] = (ok && i < (int)M) ? __ldg(A + am[i] + oa) : ezero((E*)0);
                b[u][i] = (ok && i < (int)N) ? __ldg(B + bn[i] + ob) : ezero((E*)0);
            }
        }
#pragma unroll
        for (int u = 0; u < THIN_UNROLL; u++)
#pragma unroll
            for (int i = 0; i < THIN_MAX; i++) {
                E av = p.conjA ? econj(a[u][i]) : a[u][i];
#pragma unroll
                for (int j = 0; j < THIN_MAX; j++) {
                    E bv = p.conjB ? econj(b[u][j]) : b[u][j];
                    emac(acc[i][j], av, bv);
                }
            }
    }
    // block reduction through shared memory, fixed order
    __shared__ E red[THIN_THREADS / 32][THIN_MAX * THIN_MAX];
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
    for (int i = 0; i < THIN_MAX; i++)
#pragma unroll
        for (int j = 0; j < THIN_MAX; j++) {
            E v = acc[i][j];
            if (i < (int)M && j < (int)N) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    if constexpr (CPLX) {
                        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
                        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
                    } else {
                        v += __shfl_xor_sync(0xffffffffu, v, o);
                    }
                }
                if (lane == 0) red[warp][i * THIN_MAX + j] = v;
            }
        }
    __syncthreads();
    if (threadIdx.x < THIN_MAX * THIN_MAX) {
        int i = threadIdx.x / THIN_MAX, j = threadIdx.x % THIN_MAX;
        if (i < (int)M && j < (int)N) {
            E s = ezero((E*)0);
            for (int w = 0; w < THIN_THREADS / 32; w++) s = eadd(s, red[w][threadIdx.x]);
            // partial layout expected by splitk_reduce_kernel: ws[(ks*N + n)*M + m]
            ((E*)p.ws)[((uint64_t)blockIdx.x * N + j) * M + i] = s;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// "k-reduction" kernel: small M x N (M*N <= 512), K huge, L == 1, complex, both operands dense [K][M] / [K][N]
// (free index fastest — what the planner's layouts give every intermediate).  HBM-bound (AI = MN/(M+N) <= ~12 flop/B):
// a CTA streams a contiguous K chunk tile by tile (KS k's = one contiguous block of KS*M resp. KS*N elements, fetched
// with 16-byte cp.async into a two-stage ring), threads are (k-slice, 2 x 4 micro-tile of C): per k a thread reads
// 2 A and 4 B elements (three LDS.128) for 8 complex MACs.  Partials of the k-slices are summed in shared memory in a
// fixed order, one partial per CTA goes to the split-K workspace, the deterministic reducer finishes.
// Algorithmic bytes per launch: sizeof(T) * (M + N) * K.
// ---------------------------------------------------------------------------------------------------------
constexpr int KRED_THREADS = 256;

struct KredArgs {
    const void* A; const void* B; void* ws;
    int64_t K, kchunk;
    int32_t M, N, KS, kslices, mt, nt, conjA, conjB;
};

__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <typename E, int TM>   // float2 / double2; TM x 4 micro-tile of C per thread (TM = 2 or 4)
__global__ void __launch_bounds__(KRED_THREADS) einsum_kred_kernel(const KredArgs p) {
    extern __shared__ __align__(16) unsigned char kred_smem[];
    constexpr int VEC = 16 / (int)sizeof(E);                       // elements per 16-byte copy
    const int M = p.M, N = p.N, KS = p.KS;
    const int tileA = KS * M, tileB = KS * N, stage_elems = tileA + tileB;
    E* sm = reinterpret_cast<E*>(kred_smem);
    const E* __restrict__ A = (const E*)p.A;
    const E* __restrict__ B = (const E*)p.B;
    const int64_t k_begin = (int64_t)blockIdx.x * p.kchunk;
    const int64_t k_end = (k_begin + p.kchunk) < p.K ? (k_begin + p.kchunk) : p.K;
    const int ntile = k_begin < k_end ? (int)((k_end - k_begin + KS - 1) / KS) : 0;
    const int tid = threadIdx.x;

    auto load = [&](int stage, int t) {
        const int64_t k0 = k_begin + (int64_t)t * KS;
        const int64_t kv = (k_end - k0) < KS ? (k_end - k0) : KS;  // valid k's in this tile
        E* dA = sm + stage * stage_elems;
        E* dB = dA + tileA;
        const E* gA = A + k0 * M;
        const E* gB = B + k0 * N;
        for (int i = tid * VEC; i < tileA; i += KRED_THREADS * VEC) cp_async16_zfill(dA + i, gA + i, i < kv * M);
        for (int i = tid * VEC; i < tileB; i += KRED_THREADS * VEC) cp_async16_zfill(dB + i, gB + i, i < kv * N);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int per = p.mt * p.nt;
    const bool active = tid < p.kslices * per;
    const int ks = tid / per, r = tid - ks * per;
    const int m0 = (r / p.nt) * TM, n0 = (r % p.nt) * 4;
    E acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = ezero((E*)0);

    if (ntile > 0) load(0, 0);
    for (int t = 0; t < ntile; t++) {
        if (t + 1 < ntile) { load((t + 1) & 1, t + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (active) {
            const E* sA = sm + (t & 1) * stage_elems;
            const E* sB = sA + tileA;
#pragma unroll 4
            for (int kk = ks; kk < KS; kk += p.kslices) {
                E a[TM], b[4];
#pragma unroll
                for (int i = 0; i < TM; i++) a[i] = sA[kk * M + m0 + i];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = sB[kk * N + n0 + j];
                if (p.conjA) {
#pragma unroll
                    for (int i = 0; i < TM; i++) a[i] = econj(a[i]);
                }
                if (p.conjB) {
#pragma unroll
                    for (int j = 0; j < 4; j++) b[j] = econj(b[j]);
                }
#pragma unroll
                for (int i = 0; i < TM; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) emac(acc[i][j], a[i], b[j]);
            }
        }
        __syncthreads();
    }
    // k-slice partials -> shared (reusing the ring), summed in slice order
    E* red = sm;
    if (active) {
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) red[(size_t)ks * M * N + (n0 + j) * M + m0 + i] = acc[i][j];
    }
    __syncthreads();
    for (int o = tid; o < M * N; o += KRED_THREADS) {
        E s = ezero((E*)0);
        for (int q = 0; q < p.kslices; q++) s = eadd(s, red[(size_t)q * M * N + o]);
        ((E*)p.ws)[(uint64_t)blockIdx.x * M * N + o] = s;           // ws[(cta*N + n)*M + m]
    }
}


// ---------------------------------------------------------------------------------------------------------
// k-reduction on the warp-level tensor-core path (complex64 only; M = 16*MT, N = 8*NT, MT*NT <= 4).
// With M*N = 512 the FP32-FMA kernel above sits on the ridge of the SIMT roofline (10.7 flop/B: the FMA pipe and HBM
// both need ~2 ms for the 16 x 32 x 2^25 step of the committed Sycamore path), so neither can be hidden behind the
// other.  Here the MACs go to `mma.sync.m16n8k8.tf32` (SASS HMMA.1688.F32.TF32) with the same error-free 3xTF32 split
// as the tcgen05 kernels (x = hi + lo, hi = x & 0xffffe000; hi*lo + lo*hi + hi*hi; -A_im folded into the A fragment),
// which leaves the FMA pipe idle and the kernel purely HBM-bound:
//   * no shared-memory staging: C = A^T B with A stored [K][M], B stored [K][N] is exactly the "row x col" fragment
//     layout of the instruction, so every lane loads its fragment elements straight from global memory (8-byte
//     loads; a warp-wide load covers whole 64-byte runs of 4 consecutive k rows: every fetched sector is fully used);
//   * a warp owns every 8th 8-k step of the CTA's K chunk (the 8 warps of a CTA walk one contiguous 3 KB * 8 window
//     per iteration) and prefetches the next step's 12 registers while the 12*NT*MT MMAs of the current one issue;
//   * the tensor core accumulates in FP32 with truncation, so a chain only lasts KRED_MMA_CHUNK steps
//     (48 MMAs per accumulator); the chunk is then added round-to-nearest into per-warp FP32 totals in shared memory;
//   * the 8 warps' totals are summed in a fixed order, one partial per CTA goes to the split-K workspace
//     (same layout as the kernel above) and the deterministic reducer finishes.
// Algorithmic bytes per launch: 8 * (M + N) * K.
// ---------------------------------------------------------------------------------------------------------
constexpr int KRED_MMA_CHUNK = 4;

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int MT, int NT>
__global__ void __launch_bounds__(KRED_THREADS, 2) einsum_kred_mma_kernel(const KredArgs p) {
    constexpr int M = 16 * MT, N = 8 * NT, R = MT * NT * 8;      // R accumulator registers per lane
    __shared__ float tot[KRED_THREADS / 32][R][32];
    const float2* __restrict__ A = (const float2*)p.A;
    const float2* __restrict__ B = (const float2*)p.B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t k_begin = (int64_t)blockIdx.x * p.kchunk;
    const int64_t k_end = (k_begin + p.kchunk) < p.K ? (k_begin + p.kchunk) : p.K;
    const int64_t nstep = k_begin < k_end ? (k_end - k_begin + 7) / 8 : 0;
    const float sa = p.conjA ? -1.f : 1.f, sb = p.conjB ? -1.f : 1.f;

#pragma unroll
    for (int r = 0; r < R; r++) tot[warp][r][lane] = 0.f;

    float2 ra[MT][4], rb[NT][2];                                  // raw fragment elements of one step
    auto fetch = [&](int64_t step) {
        const int64_t k0 = k_begin + step * 8 + t, k1 = k0 + 4;
        const bool v0 = k0 < k_end, v1 = k1 < k_end;
        const float2 z = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < MT; i++) {
            const float2* a0 = A + k0 * M + i * 16 + g;
            const float2* a1 = A + k1 * M + i * 16 + g;
            ra[i][0] = v0 ? __ldcs(a0) : z;       // (row g,     k = t)
            ra[i][1] = v0 ? __ldcs(a0 + 8) : z;   // (row g + 8, k = t)
            ra[i][2] = v1 ? __ldcs(a1) : z;       // (row g,     k = t + 4)
            ra[i][3] = v1 ? __ldcs(a1 + 8) : z;   // (row g + 8, k = t + 4)
        }
#pragma unroll
        for (int j = 0; j < NT; j++) {
            rb[j][0] = v0 ? __ldcs(B + k0 * N + j * 8 + g) : z;   // (k = t,     col g)
            rb[j][1] = v1 ? __ldcs(B + k1 * N + j * 8 + g) : z;   // (k = t + 4, col g)
        }
    };

    float acc[MT][NT][2][4];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[i][j][c][q] = 0.f;

    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++)
#pragma unroll
                for (int c = 0; c < 2; c++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int r = ((i * NT + j) * 2 + c) * 4 + q;
                        tot[warp][r][lane] += acc[i][j][c][q];
                        acc[i][j][c][q] = 0.f;
                    }
    };

    int in_chunk = 0;
    if (warp < nstep) fetch(warp);
    for (int64_t step = warp; step < nstep; step += KRED_THREADS / 32) {
        // split the raw elements (frees the raw registers for the prefetch of this warp's next step)
        uint32_t are_h[MT][4], are_l[MT][4], aim_h[MT][4], aim_l[MT][4], nim_h[MT][4], nim_l[MT][4];
        uint32_t bre_h[NT][2], bre_l[NT][2], bim_h[NT][2], bim_l[NT][2];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                split_tf32(ra[i][q].x, are_h[i][q], are_l[i][q]);
                split_tf32(sa * ra[i][q].y, aim_h[i][q], aim_l[i][q]);
                nim_h[i][q] = aim_h[i][q] ^ 0x80000000u;
                nim_l[i][q] = aim_l[i][q] ^ 0x80000000u;
            }
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int q = 0; q < 2; q++) {
                split_tf32(rb[j][q].x, bre_h[j][q], bre_l[j][q]);
                split_tf32(sb * rb[j][q].y, bim_h[j][q], bim_l[j][q]);
            }
        if (step + KRED_THREADS / 32 < nstep) fetch(step + KRED_THREADS / 32);
        // small terms first: hi*lo, lo*hi, then hi*hi.   Cre += Are Bre + (-Aim) Bim ;  Cim += Are Bim + Aim Bre
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                mma_tf32_16x8x8(acc[i][j][0], are_h[i], bre_l[j][0], bre_l[j][1]);
                mma_tf32_16x8x8(acc[i][j][1], are_h[i], bim_l[j][0], bim_l[j][1]);
            }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                mma_tf32_16x8x8(acc[i][j][0], nim_h[i], bim_l[j][0], bim_l[j][1]);
                mma_tf32_16x8x8(acc[i][j][1], aim_h[i], bre_l[j][0], bre_l[j][1]);
            }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                mma_tf32_16x8x8(acc[i][j][0], are_l[i], bre_h[j][0], bre_h[j][1]);
                mma_tf32_16x8x8(acc[i][j][1], are_l[i], bim_h[j][0], bim_h[j][1]);
            }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                mma_tf32_16x8x8(acc[i][j][0], nim_l[i], bim_h[j][0], bim_h[j][1]);
                mma_tf32_16x8x8(acc[i][j][1], aim_l[i], bre_h[j][0], bre_h[j][1]);
            }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                mma_tf32_16x8x8(acc[i][j][0], are_h[i], bre_h[j][0], bre_h[j][1]);
                mma_tf32_16x8x8(acc[i][j][1], are_h[i], bim_h[j][0], bim_h[j][1]);
            }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                mma_tf32_16x8x8(acc[i][j][0], nim_h[i], bim_h[j][0], bim_h[j][1]);
                mma_tf32_16x8x8(acc[i][j][1], aim_h[i], bre_h[j][0], bre_h[j][1]);
            }
        if (++in_chunk == KRED_MMA_CHUNK) { flush(); in_chunk = 0; }
    }
    flush();
    __syncthreads();
    // warp w sums registers r = w, w + 8, ... of all warps in warp order and writes them out.
    // accumulator fragment: q -> (row g + 8*(q>>1), col 2t + (q&1)) of the 16 x 8 tile (i, j); c = re / im
    float* ws = (float*)p.ws + (uint64_t)blockIdx.x * M * N * 2;
    for (int r = warp; r < R; r += KRED_THREADS / 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < KRED_THREADS / 32; w++) s += tot[w][r][lane];
        const int q = r & 3, c = (r >> 2) & 1, ij = r >> 3, i = ij / NT, j = ij - i * NT;
        const int m = i * 16 + g + 8 * (q >> 1), n = j * 8 + 2 * t + (q & 1);
        ws[((uint64_t)n * M + m) * 2 + c] = s;                      // ws[(cta*N + n)*M + m] as (re, im)
    }
}

}  // namespace

// thin path: returns number of CTAs (= splitk) or 0 if not applicable
int tnb_choose_thin(const tnb_ctx* ctx, int64_t M, int64_t N, int64_t K, int64_t L, int64_t* kchunk, int64_t* ws_elems) {
    if (L != 1 || M > THIN_MAX || N > THIN_MAX || K < 4096) return 0;
    const int64_t sms = ctx ? ctx->sm_count : 148;
    int64_t ctas = sms * 8;
    int64_t per = THIN_THREADS * THIN_UNROLL;
    int64_t kc = (K + ctas - 1) / ctas;
    kc = (kc + per - 1) / per * per;
    ctas = (K + kc - 1) / kc;
    *kchunk = kc;
    *ws_elems = ctas * M * N;
    return (int)ctas;
}

template <typename R, bool CPLX>
static int launch_thin(tnb_ctx* ctx, const EinsumArgs& a) {
    einsum_thin_kernel<R, CPLX><<<(unsigned)a.splitk, THIN_THREADS, 0, ctx->stream>>>(a);
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}

int tnb_launch_einsum_thin(tnb_ctx* ctx, int dtype, const EinsumArgs& args) {
    switch (dtype) {
        case TNB_C128: return launch_thin<double, true>(ctx, args);
        case TNB_C64: return launch_thin<float, true>(ctx, args);
        case TNB_F64: return launch_thin<double, false>(ctx, args);
        case TNB_F32: return launch_thin<float, false>(ctx, args);
    }
    return tnb_set_error(ctx, TNB_EUNSUPPORTED, "unsupported dtype %d", dtype);
}

int tnb_choose_splitk(const tnb_ctx* ctx, int64_t M, int64_t N, int64_t K, int64_t L, int64_t* kchunk,
                      int64_t* ws_elems) {
    *kchunk = K;
    *ws_elems = 0;
    int64_t tiles = ((M + TM - 1) / TM) * ((N + TN - 1) / TN) * L;
    const int64_t sms = ctx ? ctx->sm_count : 148;
    if (tiles >= sms || K < 128) return 1;
    int64_t want = (2 * sms + tiles - 1) / tiles;          // ~2 CTAs per SM
    int64_t maxs = K / 32;                                  // at least 32 k per split
    int64_t s = want < maxs ? want : maxs;
    const int64_t WS_MAX = (int64_t)1 << 24;                // elements
    while (s > 1 && s * M * N * L > WS_MAX) s--;
    if (s <= 1) return 1;
    int64_t kc = (K + s - 1) / s;
    kc = (kc + BK - 1) / BK * BK;
    s = (K + kc - 1) / kc;
    if (s <= 1) return 1;
    *kchunk = kc;
    *ws_elems = s * M * N * L;
    return (int)s;
}

template <typename R, bool CPLX>
static int launch_tile(tnb_ctx* ctx, const EinsumArgs& a) {
    int64_t tilesM = (a.M + TM - 1) / TM, tilesN = (a.N + TN - 1) / TN;
    int64_t blocks = tilesM * tilesN * a.L * a.splitk;
    if (blocks <= 0) return TNB_OK;
    if (blocks >= ((int64_t)1 << 31)) return tnb_set_error(ctx, TNB_EUNSUPPORTED, "grid too large (%lld tiles)", (long long)blocks);
    einsum_tile_kernel<R, CPLX><<<(unsigned)blocks, NTHREADS, 0, ctx->stream>>>(a);
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}

int tnb_launch_einsum_generic(tnb_ctx* ctx, int dtype, const EinsumArgs& args) {
    switch (dtype) {
        case TNB_C128: return launch_tile<double, true>(ctx, args);
        case TNB_C64: return launch_tile<float, true>(ctx, args);
        case TNB_F64: return launch_tile<double, false>(ctx, args);
        case TNB_F32: return launch_tile<float, false>(ctx, args);
    }
    return tnb_set_error(ctx, TNB_EUNSUPPORTED, "unsupported dtype %d", dtype);
}

template <typename R, bool CPLX>
static int launch_reduce(tnb_ctx* ctx, const EinsumArgs& a) {
    int64_t total = a.M * a.N * a.L;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    splitk_reduce_kernel<R, CPLX><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}

int tnb_launch_splitk_reduce(tnb_ctx* ctx, int dtype, const EinsumArgs& args) {
    switch (dtype) {
        case TNB_C128: return launch_reduce<double, true>(ctx, args);
        case TNB_C64: return launch_reduce<float, true>(ctx, args);
        case TNB_F64: return launch_reduce<double, false>(ctx, args);
        case TNB_F32: return launch_reduce<float, false>(ctx, args);
    }
    return tnb_set_error(ctx, TNB_EUNSUPPORTED, "unsupported dtype %d", dtype);
}

// k-reduction path (dense operands, small M x N, huge K): returns the number of CTAs (= splitk), 0 if not applicable
int tnb_choose_kred(const tnb_ctx* ctx, int dtype, int64_t M, int64_t N, int64_t K, int64_t L, bool a_mmajor, bool b_nmajor,
                    int64_t* kchunk, int64_t* ws_elems) {
    if (!tnb_dtype_complex(dtype) || L != 1 || !a_mmajor || !b_nmajor || K < 16384) return 0;
    if (M < 2 || N < 4 || (M % 2) || (N % 4) || M * N > 512 || M > 64 || N > 128) return 0;
    const int64_t sms = ctx ? ctx->sm_count : 148;
    int64_t ctas = sms * 4;
    const int64_t per = 128;                                   // multiple of every KS
    int64_t kc = (K + ctas - 1) / ctas;
    kc = (kc + per - 1) / per * per;
    ctas = (K + kc - 1) / kc;
    *kchunk = kc;
    *ws_elems = ctas * M * N;
    return (int)ctas;
}

// complex64 k-reductions with M = 16*MT, N = 8*NT (MT*NT <= 4) go to the mma.sync kernel (measured 2.0x the FP32-FMA
// kernel on the 16 x 32 x 2^25 step); TNB_KRED_MMA=0 keeps everything on the FMA kernel (debug / comparison).
static bool kred_mma_enabled() {
    static const int on = [] { const char* e = getenv("TNB_KRED_MMA"); return e ? atoi(e) : 1; }();
    return on != 0;
}

int tnb_launch_einsum_kred(tnb_ctx* ctx, int dtype, const EinsumArgs& a) {
    const size_t esz = tnb_dtype_size(dtype);
    KredArgs k;
    k.A = a.A; k.B = a.B; k.ws = a.ws; k.K = a.K; k.kchunk = a.kchunk;
    k.M = (int32_t)a.M; k.N = (int32_t)a.N; k.conjA = a.conjA; k.conjB = a.conjB;
    // TNB_C64_SIMT is the exact-FP32 policy: it must not see tensor-core rounding on the amplitude-closing steps either
    const bool simt_only = ctx && ctx->c64_mode == TNB_C64_SIMT;
    if (dtype == TNB_C64 && !simt_only && kred_mma_enabled() && k.M % 16 == 0 && k.N % 8 == 0 && !((uintptr_t)a.A % 8) && !((uintptr_t)a.B % 8)) {
        const int MT = k.M / 16, NT = k.N / 8;
        k.KS = k.kslices = k.mt = k.nt = 0;
        bool done = true;
#define TNB_KRED_MMA_LAUNCH(MT_, NT_) einsum_kred_mma_kernel<MT_, NT_><<<(unsigned)a.splitk, KRED_THREADS, 0, ctx->stream>>>(k)
        if (MT == 1 && NT == 4) TNB_KRED_MMA_LAUNCH(1, 4);
        else if (MT == 2 && NT == 2) TNB_KRED_MMA_LAUNCH(2, 2);
        else if (MT == 1 && NT == 2) TNB_KRED_MMA_LAUNCH(1, 2);
        else if (MT == 2 && NT == 1) TNB_KRED_MMA_LAUNCH(2, 1);
        else if (MT == 1 && NT == 1) TNB_KRED_MMA_LAUNCH(1, 1);
        else done = false;
#undef TNB_KRED_MMA_LAUNCH
        if (done) {
            ctx->launches++;
            TNB_CUDA_CHECK(ctx, cudaGetLastError());
            return TNB_OK;
        }
    }
    if (((uintptr_t)a.A % 16) || ((uintptr_t)a.B % 16)) return -1;
    const int TM = (k.M % 4 == 0 && k.M * k.N >= 256) ? 4 : 2;     // 4 x 4 micro-tiles halve the shared-memory reads per MAC
    k.mt = k.M / TM; k.nt = k.N / 4;
    int ks = 1;
    while (ks * 2 * k.mt * k.nt <= KRED_THREADS) ks *= 2;
    k.kslices = ks;
    int KS = 128;                                              // stage <= 24 KB
    while (KS > 16 && (size_t)KS * (k.M + k.N) * esz > 24 * 1024) KS /= 2;
    if (KS < ks) KS = ks;
    k.KS = KS;
    size_t smem = 2 * (size_t)KS * (k.M + k.N) * esz;
    const size_t red = (size_t)ks * k.M * k.N * esz;
    if (red > smem) smem = red;
#define TNB_KRED_LAUNCH(E, T)                                                                                              \
    do {                                                                                                                   \
        TNB_CUDA_CHECK(ctx, cudaFuncSetAttribute(einsum_kred_kernel<E, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
        einsum_kred_kernel<E, T><<<(unsigned)a.splitk, KRED_THREADS, smem, ctx->stream>>>(k);                              \
    } while (0)
    if (dtype == TNB_C64) { if (TM == 4) TNB_KRED_LAUNCH(float2, 4); else TNB_KRED_LAUNCH(float2, 2); }
    else { if (TM == 4) TNB_KRED_LAUNCH(double2, 4); else TNB_KRED_LAUNCH(double2, 2); }
#undef TNB_KRED_LAUNCH
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}

// Strided zero fill of an output view (tnb_plan_execute with accumulate = 0 and an empty slice range: the header
// contract "out = sum of the requested slices" means zeros, also for non-dense views).
struct ZeroArgs { int32_t rank; int64_t total; int64_t ext[TNB_MAX_RANK]; int64_t stride[TNB_MAX_RANK]; };
template <typename T>
__global__ void zero_strided_kernel(T* __restrict__ base, const ZeroArgs z) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < z.total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i, off = 0;
        for (int d = 0; d < z.rank; d++) { off += (r % z.ext[d]) * z.stride[d]; r /= z.ext[d]; }
        base[off] = T{};
    }
}
int tnb_launch_zero_strided(tnb_ctx* ctx, int dtype, void* base, int rank, const int64_t* ext, const int64_t* stride) {
    if (rank < 0 || rank > TNB_MAX_RANK) return tnb_set_error(ctx, TNB_EUNSUPPORTED, "zero fill: rank %d", rank);
    ZeroArgs z;
    z.rank = rank; z.total = 1;
    for (int d = 0; d < rank; d++) { z.ext[d] = ext[d]; z.stride[d] = stride[d]; z.total *= ext[d]; }
    int64_t blocks = (z.total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    switch (tnb_dtype_size(dtype)) {
        case 16: zero_strided_kernel<double2><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double2*)base, z); break;
        case 8: zero_strided_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)base, z); break;
        case 4: zero_strided_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)base, z); break;
        default: return tnb_set_error(ctx, TNB_EUNSUPPORTED, "zero fill: dtype %d", dtype);
    }
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}
