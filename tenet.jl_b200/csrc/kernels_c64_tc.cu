// kernels_c64_tc.cu — ComplexF32 steps on the 5th-gen tensor cores (tcgen05 + TMEM), 3xTF32 split.
//
// What it replaces: the BLAS cgemm (plus the permutedims passes around it) behind Muscle.binary_einsum for the large
// steps of a contraction path (/root/reference/src/Operations/overlap.jl:12 -> contract -> binary_einsum).
//
//   C[m,n] = sum_k A[m,k] * B[n,k]        A: M-fastest dense [K][M], B: N-fastest dense [K][N]  (planner layouts)
//
// Complex product from real sub-GEMMs:  Cre = Are.Bre - Aim.Bim,  Cim = Are.Bim + Aim.Bre  (minus sign = a_negate bit).
// Each real product a*b is evaluated as  a_hi*b_lo + a_lo*b_hi + a_hi*b_hi  with x_hi = x & 0xffffe000 (what the
// tensor core reads anyway), x_lo = x - x_hi (exact): 12 tcgen05.mma.kind::tf32 per 8 complex k per tile, FP32
// accumulation in TMEM.  Raw operand tiles arrive by 2-D TMA tile loads (a tensor map over the operand's own
// leading dimension) and are split by worker warps into four planes per operand (re_hi, re_lo, im_hi, im_lo): the A
// planes go to TENSOR MEMORY (tcgen05.st, ".ts" MMAs), the B planes to shared memory in the UMMA K-major no-swizzle
// core-matrix layout; the permuted / split operand never exists in global memory.
//
// Three kernels share these pieces:
//   c64_tf32x3_kernel<NT>      "fast mode": 128 x NT tile, whole K chained in TMEM (error grows with K), register prefetch,
//                              both operands in shared memory
//   c64_tf32x3_acc_kernel      default GEMM kernel: persistent CTAs over 128 x 128 tiles, TMA raw ring, A in TMEM, TMEM chunks
//                              of 128 k drained into round-to-nearest FP32 totals, ragged edges, split-K
//   c64_tf32x3_stem_kernel     persistent kernel for huge x small steps (small operand resident as planes or streamed,
//                              A in TMEM, sorted-pattern coalesced epilogue overlapped with the next tile)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>
#include "tnb_internal.h"

namespace {

constexpr int TC_BM = 128;          // tile rows (UMMA M, cta_group::1)
constexpr int TC_BK = 8;            // complex k per stage = UMMA K for tf32
constexpr int TC_PRODUCERS = 256;   // 8 warps
constexpr int TC_THREADS = TC_PRODUCERS + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// one lane of a CONVERGED warp: the MMA-issue loops run warp-uniform and only the tcgen05.mma / commit are predicated on
// this, so ptxas emits straight-line UTCHMMA code (an `if (lane == 0)` region makes it wrap every MMA in an
// ELECT / BRA.U.ANY serialisation loop: ~250 SASS instructions per k-block from one thread, which paced the kernels)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle: core matrix = 8 rows x 16 B (contiguous 128 B); the two core matrices along K (UMMA K = 8
// tf32 = 32 B) are LBO apart, consecutive 8-row groups are SBO apart.  Plane layout used here:
//   byte(row, k) = (row/8)*256 + (k/4)*128 + (row%8)*16 + (k%4)*4      =>  LBO = 128, SBO = 256
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(128u >> 4) << 16;   // leading-dimension byte offset (K direction)
    d |= (uint64_t)(256u >> 4) << 32;   // stride-dimension byte offset (M/N direction)
    d |= 1ull << 46;                    // descriptor version (Blackwell)
    return d;                           // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

template <int NT>
__host__ __device__ constexpr uint32_t make_idesc(bool neg_a) {
    return (1u << 4)                    // D format F32
           | (2u << 7) | (2u << 10)     // A, B format TF32
           | ((neg_a ? 1u : 0u) << 13)  // negate A
           | (0u << 15) | (0u << 16)    // A, B K-major
           | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, A operand in tensor memory (lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// x_hi = x with the low 13 mantissa bits cleared (what the tensor core would read anyway), x_lo = x - x_hi exactly.
// One LOP3 instead of the multi-instruction emulation of cvt.rna.tf32 on sm_100a; the split stays error-free and
// x_lo (<= 2^-10 |x|) is truncated to TF32 by the MMA, leaving a representation error <= 2^-21 |x|.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// L2-friendly tile order: groups of RASTER_GM m-tiles sweep all n-tiles, so the CTAs of one wave share A and B tiles.
constexpr uint32_t RASTER_GM = 8;
__device__ __forceinline__ void raster(uint32_t pid, uint32_t tilesM, uint32_t tilesN, uint32_t& bm, uint32_t& bn) {
    const uint32_t per_group = RASTER_GM * tilesN;
    const uint32_t group = pid / per_group, first_m = group * RASTER_GM;
    const uint32_t gsize = (tilesM - first_m) < RASTER_GM ? (tilesM - first_m) : RASTER_GM;
    const uint32_t r = pid - group * per_group;
    bm = first_m + r % gsize;
    bn = r / gsize;
}

template <int NT> struct TcSmem {
    static constexpr int A_PLANE = TC_BM * TC_BK * 4;    // 4 KB
    static constexpr int B_PLANE = NT * TC_BK * 4;
    static constexpr int STAGE = 4 * A_PLANE + 4 * B_PLANE;
    static constexpr int STAGES = (NT == 256) ? 4 : (NT == 128 ? 6 : 8);
    static constexpr int BAR_OFF = STAGES * STAGE;                 // full[STAGES], empty[STAGES], acc_full
    static constexpr int TMEM_OFF = BAR_OFF + (2 * STAGES + 1) * 8;
    static constexpr int CN_OFF = (TMEM_OFF + 4 + 15) / 16 * 16;   // int64 cn offsets [NT]
    static constexpr int TOTAL = CN_OFF + NT * 8;
};

struct TcArgs {
    const float2* A;
    const float2* B;
    float2* C;
    uint32_t M, N, K;
    int64_t lda, ldb;         // element stride between consecutive k
    TabRef cm, cn;
    int32_t conjA, conjB;
    float alpha[2], beta[2];
    // split-K (chunked kernel only): CTA z handles k-blocks [z*kb_per_split, ...) and writes raw partials to
    // ws[(z*N + n)*M + m]; the deterministic reducer of kernels_generic.cu finishes the job
    float2* ws;
    uint32_t splitk, kb_per_split;
    // pair kernel, STAGED epilogue (huge x small steps, stem tables of the planner): tile bases hi[M / 128], sorted tile
    // pattern rel / its inverse pos per 128-column pass [N / 128][128 * 128]
    const int64_t* st_hi;
    const int64_t* st_rel;
    const int64_t* st_pos;
    int32_t st_run_shift, st_vec2, st_rel_small;   // log2 run length; 16-byte pairs allowed; every rel entry fits int32
    int32_t st_direct;        // rows stored straight from the drain registers (planner.cpp st_direct): no staging tile
};

__device__ __forceinline__ int64_t tabc(const TabRef& t, uint32_t i) {
    uint32_t q = i / t.lo_size;
    uint32_t r = i - q * t.lo_size;
    return t.hi[q] + t.lo[r];
}

// one producer work unit: 4 consecutive k of one row -> four planes (re_hi, re_lo, im_hi, im_lo)
__device__ __forceinline__ void split_store(uint8_t* plane0, int plane_bytes, int row, int kc, const float2 v[4], int conj) {
    float rh[4], rl[4], ih[4], il[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float re = v[i].x, im = conj ? -v[i].y : v[i].y;
        rh[i] = tf32_hi(re); rl[i] = re - rh[i];
        ih[i] = tf32_hi(im); il[i] = im - ih[i];
    }
    const int off = (row >> 3) * 256 + kc * 128 + (row & 7) * 16;
    *reinterpret_cast<float4*>(plane0 + off) = make_float4(rh[0], rh[1], rh[2], rh[3]);
    *reinterpret_cast<float4*>(plane0 + plane_bytes + off) = make_float4(rl[0], rl[1], rl[2], rl[3]);
    *reinterpret_cast<float4*>(plane0 + 2 * plane_bytes + off) = make_float4(ih[0], ih[1], ih[2], ih[3]);
    *reinterpret_cast<float4*>(plane0 + 3 * plane_bytes + off) = make_float4(il[0], il[1], il[2], il[3]);
}

// 4 consecutive k of this thread's row (= its TMEM lane) -> planes rh | rl | ih | il, 8 columns apart, of a TMEM A stage
__device__ __forceinline__ void split_store_tmem(uint32_t taddr, const float2 v[4], int conj) {
    uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float re = v[i].x, im = conj ? -v[i].y : v[i].y;
        const float h = tf32_hi(re), g = tf32_hi(im);
        rh[i] = __float_as_uint(h); rl[i] = __float_as_uint(re - h);
        ih[i] = __float_as_uint(g); il[i] = __float_as_uint(im - g);
    }
    tmem_st4(taddr, rh[0], rh[1], rh[2], rh[3]);
    tmem_st4(taddr + 8, rl[0], rl[1], rl[2], rl[3]);
    tmem_st4(taddr + 16, ih[0], ih[1], ih[2], ih[3]);
    tmem_st4(taddr + 24, il[0], il[1], il[2], il[3]);
}

template <int NT>
__global__ void __launch_bounds__(TC_THREADS, 1) c64_tf32x3_kernel(const TcArgs p) {
    using S = TcSmem<NT>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint32_t bm, bn;
    raster(blockIdx.x, p.M / TC_BM, p.N / NT, bm, bn);
    const uint32_t m0 = bm * TC_BM, n0 = bn * NT;
    const uint32_t nkb = (p.K + TC_BK - 1) / TC_BK;

    const uint32_t bar0 = smem_u32(smem + S::BAR_OFF);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (S::STAGES + s); };
    const uint32_t acc_bar = bar0 + 8u * (2 * S::STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::TMEM_OFF);
    int64_t* cn_tab = reinterpret_cast<int64_t*>(smem + S::CN_OFF);

    if (tid == 0) {
        for (int s = 0; s < S::STAGES; s++) {
            mbar_init(full_bar(s), TC_PRODUCERS / 32);   // one arrive per producer warp
            mbar_init(empty_bar(s), 1);                  // tcgen05.commit
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 2 * NT);
    for (int i = tid; i < NT; i += TC_THREADS) cn_tab[i] = tabc(p.cn, n0 + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ===================== producers =====================
        constexpr int A_UNITS = TC_BM * 2 / TC_PRODUCERS;          // 1
        constexpr int B_UNITS = NT * 2 / TC_PRODUCERS;             // 2 (NT=256), 1 (NT=128)
        static_assert(A_UNITS >= 1 && B_UNITS >= 1, "tile too small for the producer mapping");
        float2 va[A_UNITS][4], vb[B_UNITS][4];
        auto load = [&](uint32_t kb) {
            const uint32_t k0 = kb * TC_BK;
#pragma unroll
            for (int u = 0; u < A_UNITS; u++) {
                int unit = tid + u * TC_PRODUCERS;
                int row = unit % TC_BM, kc = unit / TC_BM;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t k = k0 + kc * 4 + i;
                    va[u][i] = k < p.K ? __ldg(p.A + (int64_t)k * p.lda + (m0 + row)) : make_float2(0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < B_UNITS; u++) {
                int unit = tid + u * TC_PRODUCERS;
                int row = unit % NT, kc = unit / NT;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t k = k0 + kc * 4 + i;
                    vb[u][i] = k < p.K ? __ldg(p.B + (int64_t)k * p.ldb + (n0 + row)) : make_float2(0.f, 0.f);
                }
            }
        };
        if (nkb > 0) load(0);
        int stage = 0;
        uint32_t phase = 0;
        for (uint32_t kb = 0; kb < nkb; kb++) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            uint8_t* sa = smem + stage * S::STAGE;
            uint8_t* sb = sa + 4 * S::A_PLANE;
#pragma unroll
            for (int u = 0; u < A_UNITS; u++) {
                int unit = tid + u * TC_PRODUCERS;
                split_store(sa, S::A_PLANE, unit % TC_BM, unit / TC_BM, va[u], p.conjA);
            }
#pragma unroll
            for (int u = 0; u < B_UNITS; u++) {
                int unit = tid + u * TC_PRODUCERS;
                split_store(sb, S::B_PLANE, unit % NT, unit / NT, vb[u], p.conjB);
            }
            if (kb + 1 < nkb) load(kb + 1);        // next stage's global loads fly while the MMAs run
            fence_proxy_async_smem();              // generic-proxy stores -> visible to the tensor-core proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(stage));
            if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
        // ===================== epilogue =====================
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        const uint32_t row = q * 32 + lane;
        const int64_t cm_off = tabc(p.cm, m0 + row);
        float2* __restrict__ Crow = p.C + cm_off;
        const bool has_beta = p.beta[0] != 0.f || p.beta[1] != 0.f;
        const float ar = p.alpha[0], ai = p.alpha[1], br = p.beta[0], bi = p.beta[1];
        constexpr int COLS = NT / 2;
        for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += 16) {
            uint32_t re[16], im[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            tmem_ld16(taddr, re);
            tmem_ld16(taddr + NT, im);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float xr = __uint_as_float(re[j]), xi = __uint_as_float(im[j]);
                float2 v = make_float2(ar * xr - ai * xi, ar * xi + ai * xr);
                float2* dst = Crow + cn_tab[c0 + j];
                if (has_beta) {
                    float2 o = *dst;
                    v.x += br * o.x - bi * o.y;
                    v.y += br * o.y + bi * o.x;
                }
                *dst = v;
            }
        }
        tc_fence_before();
    } else {
        // ===================== MMA issuer (warp 8): warp-uniform loop, one elected lane issues =====================
        {
            constexpr uint32_t IDESC = make_idesc<NT>(false), IDESC_NEG = make_idesc<NT>(true);
            const uint32_t d_re = tmem_base, d_im = tmem_base + NT;
            int stage = 0;
            uint32_t phase = 0;
            for (uint32_t kb = 0; kb < nkb; kb++) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * S::STAGE);
                const uint32_t sb = sa + 4 * S::A_PLANE;
                const uint64_t a_rh = make_smem_desc(sa), a_rl = make_smem_desc(sa + S::A_PLANE),
                               a_ih = make_smem_desc(sa + 2 * S::A_PLANE), a_il = make_smem_desc(sa + 3 * S::A_PLANE);
                const uint64_t b_rh = make_smem_desc(sb), b_rl = make_smem_desc(sb + S::B_PLANE),
                               b_ih = make_smem_desc(sb + 2 * S::B_PLANE), b_il = make_smem_desc(sb + 3 * S::B_PLANE);
                const uint32_t acc = kb > 0 ? 1u : 0u;
                if (elect_one()) {
                    // Cre = Are.Bre - Aim.Bim     (small cross terms first, hi*hi last)
                    umma_tf32(d_re, a_rh, b_rl, IDESC, acc);
                    umma_tf32(d_re, a_rl, b_rh, IDESC, 1u);
                    umma_tf32(d_re, a_ih, b_il, IDESC_NEG, 1u);
                    umma_tf32(d_re, a_il, b_ih, IDESC_NEG, 1u);
                    umma_tf32(d_re, a_rh, b_rh, IDESC, 1u);
                    umma_tf32(d_re, a_ih, b_ih, IDESC_NEG, 1u);
                    // Cim = Are.Bim + Aim.Bre
                    umma_tf32(d_im, a_rh, b_il, IDESC, acc);
                    umma_tf32(d_im, a_rl, b_ih, IDESC, 1u);
                    umma_tf32(d_im, a_ih, b_rl, IDESC, 1u);
                    umma_tf32(d_im, a_il, b_rh, IDESC, 1u);
                    umma_tf32(d_im, a_rh, b_ih, IDESC, 1u);
                    umma_tf32(d_im, a_ih, b_rh, IDESC, 1u);
                    umma_commit(empty_bar(stage));     // implies tcgen05.fence::before_thread_sync
                    if (kb + 1 == nkb) umma_commit(acc_bar);
                }
                __syncwarp();
                if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * NT);
    }
}

template <int NT>
int launch_tc(tnb_ctx* ctx, const TcArgs& a) {
    using S = TcSmem<NT>;
    static bool configured[16] = {false};
    if (!configured[ctx->device & 15]) {
        TNB_CUDA_CHECK(ctx, cudaFuncSetAttribute(c64_tf32x3_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        configured[ctx->device & 15] = true;
    }
    const unsigned grid = (a.M / TC_BM) * (a.N / NT);
    c64_tf32x3_kernel<NT><<<grid, TC_THREADS, S::TOTAL, ctx->stream>>>(a);
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Chunked-accumulation variant (the default GEMM kernel): 128 x 128 tile; the TMEM accumulators only ever hold a
// CHUNK of KCB k-blocks.  The tensor core adds into its FP32 accumulator with round-toward-zero, a bias that grows
// linearly with the number of chained MMAs (measured: ~6e-8 * #MMAs relative, i.e. 1.5e-5 at K = 1024 when the whole
// K is chained).  After every chunk the 16 worker warps drain the accumulators with tcgen05.ld (measured 394 B/clk/SM:
// ~350 clk for the 128 KB, tools/probes/ldtm_probe.cu) and add them into per-thread FP32 totals with round-to-nearest.
// Chain length = 6*KCB = 96 MMAs => ~6e-6 relative.
// The A planes go to TENSOR MEMORY (tcgen05.st by the workers, ".ts" MMAs): with both operands in shared memory a
// 128x128x8 tf32 MMA reads 8 KB per 64 clk = all of the SM's 128 B/clk, and the split planes plus the raw ring pushed
// the kernel to 168 KB of shared-memory traffic per k-block (1312 clk) against 768 clk of MMA time — it was
// shared-memory bound at ~0.8 of the tensor peak.  With A in TMEM: 96 KB per k-block.  That leaves room for ONE
// accumulator set (256 columns) next to the 8-stage A ring (256 columns); the drain bubble is ~3 % of a chunk.
//
//   warps 0-15 workers: (a) producers, one (row, 4k) unit per thread and k-block (warps 0-7 feed A, 8-15 feed B),
//                       (b) every KCB k-blocks drain the finished chunk into 64 total registers
//                           (thread = row, 32 complex columns), (c) epilogue scatter of the totals.
//   warp 16    MMA issuer.
// ---------------------------------------------------------------------------------------------------------
constexpr int ACC_NT = 128;
constexpr int ACC_WORKERS = 512;              // warps 0-15
constexpr int ACC_THREADS = ACC_WORKERS + 128; // + warpgroup 4: warp 16 MMA issuer, warp 17 bulk-copy issuer, 18-19 idle
constexpr int ACC_KCB = 16;                   // k-blocks (of 8 complex k) per TMEM chunk: chain of 96 MMAs per accumulator
constexpr int ACC_RAW_STAGES = 6;             // raw (interleaved complex) operand tiles landed by cp.async.bulk
constexpr int ACC_PL_STAGES = 8;              // split planes consumed by tcgen05.mma (workers fill them two at a time):
                                              // B planes in shared memory, A planes in tensor memory (32 columns per stage)
constexpr uint32_t ACC_APL_COL0 = 256;        // TMEM: columns [0,256) accumulators (re | im), [256,512) the A plane ring

struct AccSmem {
    static constexpr int A_PLANE = TC_BM * TC_BK * 4;                  // 4 KB
    static constexpr int B_PLANE = ACC_NT * TC_BK * 4;
    static constexpr int PL_STAGE = 4 * B_PLANE;                       // 16 KB (the A planes live in tensor memory)
    static constexpr int RAW_HALF = TC_BK * TC_BM * 8;                 // 8 KB: [8 k][128 rows] float2
    static constexpr int RAW_STAGE = 2 * RAW_HALF;                     // A then B
    static constexpr int RAW_OFF = ACC_PL_STAGES * PL_STAGE;
    static constexpr int BAR_OFF = RAW_OFF + ACC_RAW_STAGES * RAW_STAGE;
    // raw_full[R], raw_empty[R], pl_full[P], pl_empty[P], accfull, accempty
    static constexpr int NBARS = 2 * ACC_RAW_STAGES + 2 * ACC_PL_STAGES + 2;
    static constexpr int TMEM_OFF = BAR_OFF + NBARS * 8;
    static constexpr int CN_OFF = (TMEM_OFF + 4 + 15) / 16 * 16;       // column offsets of the current / next work item
    static constexpr int TOTAL = CN_OFF + 2 * ACC_NT * 8;
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (UBLKCP), completion counted in bytes on an mbarrier; no tensor map needed
// because a tile row (128 consecutive m or n at fixed k) is contiguous in the planner's layouts.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// 2-D TMA tile load (UTMALDG): box {128 elements of 8 B along the free index, 8 k-rows} of an operand stored [K][M]
// with an arbitrary leading dimension, ONE request per operand and k-block instead of eight 1 KB bulk copies (the
// per-request service time of the copy engine, not bytes, was what starved the GEMM kernel's raw ring).  Rows / k
// beyond the tensor's extent are zero-filled by the hardware, which also covers ragged edge tiles.
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tm, uint32_t c0, uint32_t c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// host: tensor map over a dense-rows operand  X[k][m] = base[m + ld*k]  (8-byte elements), box = 128 x 8
typedef CUresult (*tnb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tnb_encode_tiled_fn get_encode_tiled() {
    static tnb_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (tnb_encode_tiled_fn)f;
    }
    return fn;
}
static bool make_operand_map(CUtensorMap* tm, const float2* base, uint64_t rows, uint64_t K, int64_t ld) {
    tnb_encode_tiled_fn enc = get_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = {rows, K};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};                // bytes between consecutive k (multiple of 16)
    const cuuint32_t box[2] = {TC_BM, TC_BK};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void __launch_bounds__(ACC_THREADS, 1) c64_tf32x3_acc_kernel(const TcArgs p, const __grid_constant__ CUtensorMap tmA,
                                                                        const __grid_constant__ CUtensorMap tmB) {
    using S = AccSmem;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t tilesM = (p.M + TC_BM - 1) / TC_BM, tilesN = (p.N + ACC_NT - 1) / ACC_NT;
    const uint32_t ntiles = tilesM * tilesN;
    const uint32_t nwork = ntiles * p.splitk;                         // persistent: CTA b takes work items b, b+grid, ...
    const uint32_t nkb_total = (p.K + TC_BK - 1) / TC_BK;
    // geometry of work item w (tile + K split): everything the three roles need, recomputed identically by each
    struct Work { uint32_t m0, n0, mv, nv, kb0, nkb, nchunks, zsplit; };
    auto work = [&](uint32_t w) {
        Work t;
        t.zsplit = w / ntiles;
        uint32_t bm, bn;
        raster(w - t.zsplit * ntiles, tilesM, tilesN, bm, bn);
        t.m0 = bm * TC_BM; t.n0 = bn * ACC_NT;
        t.mv = (p.M - t.m0) < TC_BM ? (p.M - t.m0) : TC_BM;           // valid rows / columns of a ragged edge tile
        t.nv = (p.N - t.n0) < ACC_NT ? (p.N - t.n0) : ACC_NT;
        t.kb0 = t.zsplit * p.kb_per_split;
        t.nkb = (nkb_total - t.kb0) < p.kb_per_split ? (nkb_total - t.kb0) : p.kb_per_split;
        t.nchunks = (t.nkb + ACC_KCB - 1) / ACC_KCB;
        return t;
    };

    const uint32_t bar0 = smem_u32(smem + S::BAR_OFF);
    auto raw_full = [&](int s) { return bar0 + 8u * s; };
    auto raw_empty = [&](int s) { return bar0 + 8u * (ACC_RAW_STAGES + s); };
    auto pl_full = [&](int s) { return bar0 + 8u * (2 * ACC_RAW_STAGES + s); };
    auto pl_empty = [&](int s) { return bar0 + 8u * (2 * ACC_RAW_STAGES + ACC_PL_STAGES + s); };
    const uint32_t accfull_bar = bar0 + 8u * (2 * ACC_RAW_STAGES + 2 * ACC_PL_STAGES);
    const uint32_t accempty_bar = accfull_bar + 8u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::TMEM_OFF);
    int64_t* cn_tab = reinterpret_cast<int64_t*>(smem + S::CN_OFF);

    if (tid == 0) {
        for (int s = 0; s < ACC_RAW_STAGES; s++) {
            mbar_init(raw_full(s), 1);                    // expect_tx arrive of the copy issuer
            mbar_init(raw_empty(s), ACC_WORKERS / 32);
        }
        for (int s = 0; s < ACC_PL_STAGES; s++) {
            mbar_init(pl_full(s), ACC_WORKERS / 32);
            mbar_init(pl_empty(s), 1);                    // tcgen05.commit
        }
        mbar_init(accfull_bar, 1);
        mbar_init(accempty_bar, ACC_WORKERS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 16) {
        // registers: launch gives every thread 96; warpgroup 4 shrinks to 32 and the 4 worker warpgroups grow to 112
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // ---- worker: raw tile -> split planes, chunk drain, epilogue ----
        const bool feeds_a = tid < 256;
        const int u = feeds_a ? tid : tid - 256;
        const int prow = u & 127, pkc = u >> 7;
        const int pconj = feeds_a ? p.conjA : p.conjB;
        // A feeders: TMEM lane = row, columns of this thread's 4 k inside a stage: plane*8 + pkc*4
        const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ACC_APL_COL0 + (uint32_t)(pkc * 4);
        const int raw_base = S::RAW_OFF + (feeds_a ? 0 : S::RAW_HALF) + (pkc * 4 * TC_BM + prow) * 8;
        const int q = warp & 3, g = warp >> 2;        // TMEM lane quarter, 32-column group
        const bool has_beta = p.beta[0] != 0.f || p.beta[1] != 0.f;
        const float ar = p.alpha[0], ai = p.alpha[1], br = p.beta[0], bi = p.beta[1];
        float tr[32], ti[32];
        uint32_t gdrained = 0;                         // chunks drained so far by this CTA (all work items)
        auto drain = [&]() {
            mbar_wait(accfull_bar, gdrained & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 32);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t r[16], im[16];
                tmem_ld16(taddr + h * 16, r);
                tmem_ld16(taddr + 128 + h * 16, im);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    tr[h * 16 + j] += __uint_as_float(r[j]);
                    ti[h * 16 + j] += __uint_as_float(im[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty_bar);
            gdrained++;
        };

        // Two k-blocks per iteration: one fence + one round of barrier traffic per 16 k, and twice the independent work
        // in flight per warp (the loop is latency- not issue-bound).  Stage counts are even.
        static_assert(ACC_RAW_STAGES % 2 == 0 && ACC_PL_STAGES % 2 == 0 && ACC_KCB % 2 == 0, "pairs of k-blocks");
        int rs = 0, ps = 0;
        uint32_t rphase = 0, pphase = 0;
        uint32_t item = 0;
        for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x, item++) {
            const Work t = work(w);
            const bool row_ok = (uint32_t)prow < (feeds_a ? t.mv : t.nv);
            // output column offsets of this work item -> shared (double buffered: the slowest warp may still be in the
            // previous item's epilogue); one worker-only barrier per item, long before the epilogue needs the table
            int64_t* cn_cur = cn_tab + (item & 1) * ACC_NT;
            if (p.splitk == 1) {
                if (tid < (int)t.nv) cn_cur[tid] = tabc(p.cn, t.n0 + tid);
                asm volatile("bar.sync 2, 512;" ::: "memory");
            }
#pragma unroll
            for (int j = 0; j < 32; j++) { tr[j] = 0.f; ti[j] = 0.f; }
            uint32_t drained = 0;                      // chunks of THIS work item drained
            for (uint32_t kb = 0; kb < t.nkb; kb += 2) {
                // single accumulator set: chunk c-1 must be drained before the MMAs of chunk c start.  Do it once the
                // plane ring is primed with the first stages of chunk c (they only need MMAs of chunk c-1 to retire),
                // so the tensor core restarts on ready stages right after the drain.
                if (kb >= ACC_KCB && kb % ACC_KCB == ACC_PL_STAGES) {
                    const uint32_t c = kb / ACC_KCB;
                    while (drained < c) { drain(); drained++; }
                }
                const bool two = kb + 1 < t.nkb;
                float2 v0[4], v1[4];
                mbar_wait(raw_full(rs), rphase);
                {
                    const uint8_t* raw = smem + rs * S::RAW_STAGE + raw_base;
                    const uint32_t kleft = p.K - (t.kb0 + kb) * TC_BK;   // >= 8 except in the last k-block of a ragged K
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        v0[i] = *reinterpret_cast<const float2*>(raw + i * TC_BM * 8);
                        if (!row_ok || (uint32_t)(pkc * 4 + i) >= kleft) v0[i] = make_float2(0.f, 0.f);   // stale smem beyond the edge
                    }
                }
                mbar_wait(raw_full(rs + 1), rphase);        // an odd tail k-block is a dummy stage on every role (phases stay aligned)
                if (two) {
                    const uint8_t* raw = smem + (rs + 1) * S::RAW_STAGE + raw_base;
                    const uint32_t kleft = p.K - (t.kb0 + kb + 1) * TC_BK;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        v1[i] = *reinterpret_cast<const float2*>(raw + i * TC_BM * 8);
                        if (!row_ok || (uint32_t)(pkc * 4 + i) >= kleft) v1[i] = make_float2(0.f, 0.f);
                    }
                }
                mbar_wait(pl_empty(ps), pphase ^ 1);
                if (feeds_a) { tc_fence_after(); split_store_tmem(a_lane + ps * 32u, v0, pconj); }
                else split_store(smem + ps * S::PL_STAGE, S::B_PLANE, prow, pkc, v0, pconj);
                mbar_wait(pl_empty(ps + 1), pphase ^ 1);
                if (two) {
                    if (feeds_a) { tc_fence_after(); split_store_tmem(a_lane + (ps + 1) * 32u, v1, pconj); }
                    else split_store(smem + (ps + 1) * S::PL_STAGE, S::B_PLANE, prow, pkc, v1, pconj);
                }
                if (feeds_a) { tmem_st_wait(); tc_fence_before(); }
                else fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(pl_full(ps));
                    mbar_arrive(raw_empty(rs));
                    mbar_arrive(pl_full(ps + 1));
                    mbar_arrive(raw_empty(rs + 1));
                }
                rs += 2; if (rs == ACC_RAW_STAGES) { rs = 0; rphase ^= 1; }
                ps += 2; if (ps == ACC_PL_STAGES) { ps = 0; pphase ^= 1; }
            }
            while (drained < t.nchunks) { drain(); drained++; }

            // ---- epilogue: totals -> alpha/beta -> scatter (the copy warp is already fetching the next work item) ----
            const uint32_t row = q * 32 + lane;
            if (row < t.mv) {
                if (p.splitk == 1) {
                    float2* __restrict__ Crow = p.C + tabc(p.cm, t.m0 + row);
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if ((uint32_t)(g * 32 + j) >= t.nv) break;
                        float2 o = make_float2(ar * tr[j] - ai * ti[j], ar * ti[j] + ai * tr[j]);
                        float2* dst = Crow + cn_cur[g * 32 + j];
                        if (has_beta) {
                            float2 old = *dst;
                            o.x += br * old.x - bi * old.y;
                            o.y += br * old.y + bi * old.x;
                        }
                        *dst = o;
                    }
                } else {
                    float2* __restrict__ W = p.ws + ((uint64_t)t.zsplit * p.N + t.n0) * p.M + (t.m0 + row);
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if ((uint32_t)(g * 32 + j) >= t.nv) break;
                        W[(uint64_t)(g * 32 + j) * p.M] = make_float2(tr[j], ti[j]);
                    }
                }
            }
        }
    } else if (warp == 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        // ---- MMA issuer: warp-uniform loop, one elected lane issues (see elect_one) ----
        constexpr uint32_t IDESC = make_idesc<ACC_NT>(false), IDESC_NEG = make_idesc<ACC_NT>(true);
        int ps = 0;
        uint32_t pphase = 0, gchunk = 0;               // chunks started so far by this CTA
        for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x) {
            const Work t = work(w);
            const uint32_t nkb2 = (t.nkb + 1) & ~1u;      // odd tail: one dummy stage (no MMAs), see the workers
            for (uint32_t kb = 0; kb < nkb2; kb++) {
                const bool first = (kb % ACC_KCB) == 0;
                if (first && kb < t.nkb) {
                    if (gchunk >= 1) mbar_wait(accempty_bar, (gchunk - 1) & 1);
                    gchunk++;
                }
                mbar_wait(pl_full(ps), pphase);
                tc_fence_after();
                if (kb >= t.nkb) {
                    if (elect_one()) umma_commit(pl_empty(ps));
                    __syncwarp();
                    if (++ps == ACC_PL_STAGES) { ps = 0; pphase ^= 1; }
                    continue;
                }
                const uint32_t d_re = tmem_base, d_im = d_re + 128u;
                const uint32_t a_rh = tmem_base + ACC_APL_COL0 + ps * 32u, a_rl = a_rh + 8, a_ih = a_rh + 16, a_il = a_rh + 24;
                const uint32_t sb = smem_u32(smem + ps * S::PL_STAGE);
                const uint64_t b_rh = make_smem_desc(sb), b_rl = make_smem_desc(sb + S::B_PLANE),
                               b_ih = make_smem_desc(sb + 2 * S::B_PLANE), b_il = make_smem_desc(sb + 3 * S::B_PLANE);
                const uint32_t acc = first ? 0u : 1u;
                if (elect_one()) {
                    umma_tf32_ts(d_re, a_rh, b_rl, IDESC, acc);
                    umma_tf32_ts(d_re, a_rl, b_rh, IDESC, 1u);
                    umma_tf32_ts(d_re, a_ih, b_il, IDESC_NEG, 1u);
                    umma_tf32_ts(d_re, a_il, b_ih, IDESC_NEG, 1u);
                    umma_tf32_ts(d_re, a_rh, b_rh, IDESC, 1u);
                    umma_tf32_ts(d_re, a_ih, b_ih, IDESC_NEG, 1u);
                    umma_tf32_ts(d_im, a_rh, b_il, IDESC, acc);
                    umma_tf32_ts(d_im, a_rl, b_ih, IDESC, 1u);
                    umma_tf32_ts(d_im, a_ih, b_rl, IDESC, 1u);
                    umma_tf32_ts(d_im, a_il, b_rh, IDESC, 1u);
                    umma_tf32_ts(d_im, a_rh, b_ih, IDESC, 1u);
                    umma_tf32_ts(d_im, a_ih, b_rh, IDESC, 1u);
                    umma_commit(pl_empty(ps));
                    if ((kb % ACC_KCB) == ACC_KCB - 1 || kb == t.nkb - 1) umma_commit(accfull_bar);
                }
                __syncwarp();
                if (++ps == ACC_PL_STAGES) { ps = 0; pphase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp > 17) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");   // idle warps: only there to complete warpgroup 4
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        // ---- copy issuer (warp 17): one 2-D TMA tile load per operand and k-block (lane 0: A, lane 1: B); it runs ahead
        // of the workers into the next work item by the depth of the raw ring ----
        int rs = 0;
        uint32_t rphase = 0;
        for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x) {
            const Work t = work(w);
            const uint32_t nkb2 = (t.nkb + 1) & ~1u;      // odd tail: one dummy stage (zero bytes)
            for (uint32_t kb = 0; kb < nkb2; kb++) {
                const uint32_t k0 = (t.kb0 + kb) * TC_BK;
                mbar_wait(raw_empty(rs), rphase ^ 1);
                if (lane == 0) mbar_expect_tx(raw_full(rs), kb < t.nkb ? (uint32_t)S::RAW_STAGE : 0u);   // full boxes: OOB is zero-filled
                __syncwarp();
                if (kb < t.nkb && lane < 2) {
                    const uint32_t dst = smem_u32(smem + S::RAW_OFF + rs * S::RAW_STAGE + (lane == 0 ? 0 : S::RAW_HALF));
                    tma_load_2d(dst, lane == 0 ? &tmA : &tmB, lane == 0 ? t.m0 : t.n0, k0, raw_full(rs));
                }
                if (++rs == ACC_RAW_STAGES) { rs = 0; rphase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int launch_tc_acc(tnb_ctx* ctx, const TcArgs& a) {
    static bool configured[16] = {false};
    if (!configured[ctx->device & 15]) {
        TNB_CUDA_CHECK(ctx, cudaFuncSetAttribute(c64_tf32x3_acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AccSmem::TOTAL));
        configured[ctx->device & 15] = true;
    }
    unsigned grid = ((a.M + TC_BM - 1) / TC_BM) * ((a.N + ACC_NT - 1) / ACC_NT) * a.splitk;
    if (grid > (unsigned)ctx->sm_count) grid = (unsigned)ctx->sm_count;   // persistent: one CTA per SM walks the work items
    CUtensorMap tmA, tmB;
    if (!make_operand_map(&tmA, a.A, a.M, a.K, a.lda) || !make_operand_map(&tmB, a.B, a.N, a.K, a.ldb))
        return tnb_set_error(ctx, TNB_ECUDA, "cuTensorMapEncodeTiled failed (M=%u N=%u K=%u lda=%lld ldb=%lld)", a.M, a.N, a.K,
                             (long long)a.lda, (long long)a.ldb);
    c64_tf32x3_acc_kernel<<<grid, ACC_THREADS, AccSmem::TOTAL, ctx->stream>>>(a, tmA, tmB);
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}

// bank swizzle of staging indices: a permutation inside aligned 16-element groups (reads of consecutive ranks stay
// conflict free, aligned pairs stay pairs) that spreads ranks which differ by a power-of-two stride over all banks
__device__ __forceinline__ uint32_t sk_swz(uint32_t r) { return r ^ ((r >> 4) & 15u) ^ ((r >> 8) & 15u); }

#include "pair_kernel.inc"

// ---------------------------------------------------------------------------------------------------------
// Persistent tensor-core STEM kernel: huge dense operand x small operand,  M huge, 16 <= N <= 128 per pass, K <= 128
// (K <= 512 in chunks for the 128-column form).
//
// These steps carry most of the BYTES of a good sliced path (arithmetic intensity 13-64 flop/B, around the ridge).
// Round-1 profiling (profiles/r1_summary.md) showed the first version of this kernel was bound by SHARED-MEMORY
// bandwidth, not by HBM or the tensor pipe: a 128 x 128 x 8 tf32 MMA with both operands in shared memory reads
// 8 KB per 64 clk = the whole 128 B/clk of the SM, and the 3xTF32 split re-reads every A plane up to twice.  So:
//   * one persistent CTA per SM walks 128-row tiles of the big operand; its A tile rows arrive by cp.async.bulk
//     (8 KB per 8 k) into a raw ring, every A byte is read from HBM exactly once;
//   * 8 worker warps read the raw tile once, split it (re/im, hi/lo) in registers and store the four A planes into
//     TENSOR MEMORY (tcgen05.st, lane = row, 8 columns per plane per k-block, 4-stage ring of 32 columns): the MMAs
//     take A from TMEM (".ts" form), so the A planes cost no shared-memory bandwidth at all and need no proxy fence;
//   * the small operand is split (hi/lo) ONCE per CTA for all of K into two resident UMMA planes P_hi, P_lo of 2N
//     rows each (rows [0,N) = B_re, rows [N,2N) = B_im) — or, when N*K*16 > 64 KB, once per launch into global memory
//     by a tiny pre-pass and streamed per k-block through a 4-stage ring (one 8 KB bulk copy, L2-resident source);
//   * MMA forms (a 128 x n x 8 tf32 MMA costs max(45.5, n/2) clk, SS or TS: tools/probes/mma_probe.cu, ts_probe.cu):
//       N <= 32: 6 MMAs of width 2N per k-block into F = A_re.[B_re|B_im], E = A_im.[B_re|B_im]  (4N TMEM columns per
//                set); the epilogue forms C_re = F[0:N] - E[N:2N], C_im = F[N:2N] + E[0:N];
//       N  = 64: 12 MMAs of width 64 into C_re | C_im directly (a_negate for the -A_im.B_im terms; 2N columns per
//                set) — F/E would need 2 x 256 accumulator columns and leave no room for the A ring;
//   * two accumulator sets: 8 epilogue warps drain one (tcgen05.ld) while the MMAs fill the other, drop each value
//     at the RANK of its output address inside the tile's (tile-invariant, planner-sorted) address pattern in a
//     shared staging tile, and write the tile to C in ascending address order, 16 B per thread — coalesced whatever
//     layout the consumer asked for.  When the rank is separable on disjoint bits, rank(row,col) = r(row) ^ c(col)
//     (always for power-of-two extents), it comes from one register and a broadcast column table.
//       N = 128: 12 MMAs of width 128 (64 clk each: the 45.5-clk floor of narrow MMAs no longer bites) into ONE set of
//                256 columns (re | im) next to the A ring; streamed-B mode only (16 KB B stages, 4-stage ring), 128 KB
//                staging tile, separable ranks only.  The big operand is read once per 128 columns and the TMEM drain
//                of a tile (not its write-out) is exposed (~10 % of a K = 128 tile).  Default for small operands whose
//                width is a multiple of 128 (planner.cpp); measured 205-216 TFLOP/s vs 170-176 for two 64-column passes;
//       WIDE64:  N = 64 with the 6-MMA form and one set (experiment, TNB_STEM_WIDE64=1: neutral);
//   * two accumulator sets (one for N = 128 / WIDE64): 8 epilogue warps drain one (tcgen05.ld) while the MMAs fill the
//     other (see above for the drain/write-out pipeline);
//   * K > 128 (N = 128 form, TNB_STEM_KMAX): the accumulators hold one chunk of SK_KCB = 16 k-blocks at a time; chunk 0 is
//     stored into the staging tile, later chunks are added to it round-to-nearest (same chunking as the GEMM kernel).
// Chain length in TMEM is <= 6*128/8 = 96 MMAs per accumulator (one chunk): round-toward-zero bias ~6e-6 relative.
// ---------------------------------------------------------------------------------------------------------
constexpr int SK_WORKERS = 256;                 // warps 0-7
constexpr int SK_EPI = 256;                     // warps 8-15: two per TMEM lane quarter, each takes half of the columns
constexpr int SK_THREADS = SK_WORKERS + SK_EPI + 64;   // + MMA warp 16, copy warp 17
constexpr int SK_PL_MAX = 4;                    // A plane stages in tensor memory (32 columns each)
constexpr int SK_RAW_MAX = 16;                  // raw A stages (8 KB each): as many as shared memory allows — the
                                                // bytes in flight per SM (>= 40 KB) are what saturates HBM
constexpr int SK_RUNS_MAX = 1024;               // run bases kept in shared memory (int32)
constexpr int SK_RAW_STAGE = TC_BK * TC_BM * 8;            // 8 KB
constexpr int SK_APL_COLS = 4 * TC_BK;                     // TMEM columns of one A stage: planes rh | rl | ih | il, 8 k each
constexpr int SK_BREP = 1;                      // streamed-B mode: replicas of the pre-split planes in global memory (CTA b reads
                                                // replica b % SK_BREP, so that 148 SMs do not hammer the same 64 L2 lines at once)
constexpr int SK_KCB = 16;                      // k-blocks per TMEM chunk (chain of <= 96 MMAs per accumulator), as in the GEMM kernel
constexpr int SK_BST = 8;                       // streamed-B mode: stages of the B plane ring (2 planes x 2*NT rows per k-block)
constexpr int SK_NBARS = 2 * SK_RAW_MAX + 2 * SK_PL_MAX + 4 + 2 * SK_BST;
constexpr int SK_BUDGET = 227 * 1024;

struct StemTcArgs {
    const float2* A;          // dense [K][M]
    const float2* B;          // small operand, gathered through bn/bk
    float2* C;
    int64_t M, lda;
    int32_t N, K, n0, conjA, conjB;
    int32_t raw_stages;       // chosen by the launcher from the shared-memory budget
    int32_t run_shift;        // log2 of the contiguous output run length; run bases are rel[j << run_shift]
    int32_t additive;         // pos[row*N + col] == pos[row*N] + pos[col] - pos[0]
    int32_t vec2;             // run >= 2, every run base even, C 16-byte aligned: the write-out moves pairs
    int32_t egroups;          // epilogue groups: 2 = two groups of 4 warps with one TMEM set + one staging tile each (two tiles in
                              // flight in the epilogue), 1 = one group of 8 warps
    int32_t direct;           // epilogue stores rows straight from registers (one column of 32 consecutive rows per warp
                              // instruction fills whole sectors — planner.cpp st_direct): no staging tile, no rank tables
    int32_t off_stg, off_tab, off_run, off_raw, off_bar;   // shared-memory map (bytes), B planes at 0
    int64_t brep_stride;      // streamed-B mode: byte distance between the SK_BREP replicas of the pre-split planes
    const uint8_t* bplanes;   // streamed-B mode: pre-split planes of this pass in global memory, [kb][hi|lo][2*NT rows x 32 B]
    TabRef bn, bk;
    const int64_t* hi;        // [M/128]
    const int64_t* rel;       // [128*N]
    const int64_t* pos;       // [128*N]
    float alpha[2], beta[2];
};

// shared-memory map: [ B planes (resident or 4-stage ring) | staging tile | rank table | run bases | raw ring | barriers, tmem slot ]
__host__ inline int sk_layout(int nt, StemTcArgs& a) {
    auto up = [](int x, int q) { return (x + q - 1) / q * q; };
    const int nkb = a.K / TC_BK;
    const int nruns = (TC_BM * a.N) >> a.run_shift;
    a.off_stg = up((a.bplanes ? (nt == 128 ? SK_BST / 2 : SK_BST) : nkb) * nt * 128, 1024);
    a.off_tab = a.off_stg + (a.direct ? 0 : a.egroups * up(TC_BM * a.N * 8, 1024));
    a.off_run = a.off_tab + up((a.additive || a.direct) ? nt * 4 : TC_BM * a.N * 2, 16);
    a.off_raw = up(a.off_run + ((nruns <= SK_RUNS_MAX && !a.direct) ? nruns * 4 : 0), 1024);
    int raw = (SK_BUDGET - (SK_NBARS * 8 + 16) - a.off_raw) / SK_RAW_STAGE;
    if (raw > SK_RAW_MAX) raw = SK_RAW_MAX;
    a.raw_stages = raw;
    a.off_bar = a.off_raw + (raw > 0 ? raw : 0) * SK_RAW_STAGE;
    return raw;
}

// small operand: 4 consecutive k of one column -> rows `row` (re) and `NT + row` (im) of the hi and lo planes
__device__ __forceinline__ void split_store_b(uint8_t* kb_base, int plane_bytes, int nt, int row, int kc, const float2 v[4], int conj) {
    float rh[4], rl[4], ih[4], il[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float re = v[i].x, im = conj ? -v[i].y : v[i].y;
        rh[i] = tf32_hi(re); rl[i] = re - rh[i];
        ih[i] = tf32_hi(im); il[i] = im - ih[i];
    }
    const int r2 = nt + row;
    const int off_re = (row >> 3) * 256 + kc * 128 + (row & 7) * 16;
    const int off_im = (r2 >> 3) * 256 + kc * 128 + (r2 & 7) * 16;
    *reinterpret_cast<float4*>(kb_base + off_re) = make_float4(rh[0], rh[1], rh[2], rh[3]);
    *reinterpret_cast<float4*>(kb_base + off_im) = make_float4(ih[0], ih[1], ih[2], ih[3]);
    *reinterpret_cast<float4*>(kb_base + plane_bytes + off_re) = make_float4(rl[0], rl[1], rl[2], rl[3]);
    *reinterpret_cast<float4*>(kb_base + plane_bytes + off_im) = make_float4(il[0], il[1], il[2], il[3]);
}


// BSTREAM: the small operand's planes do not fit in shared memory (N*K*16 > 64 KB): a tiny pre-pass
// (stem_bsplit_kernel) splits it once into global memory in the UMMA plane layout and lane 8 of the copy warp
// streams one 2*B_KB stage per k-block (one bulk copy, L2-resident source) next to the A rows.
// WIDE64 (NT = 64 only): the 6-MMA form with ONE accumulator set (F | E = 256 columns + the A ring) instead of the
// 12-MMA form with two sets: 6 x 64 clk instead of 12 x 48 clk per k-block; the price is that the TMEM drain of a
// tile (not its write-out) is no longer overlapped with the next tile's MMAs.
template <int NT, bool BSTREAM, bool WIDE64 = false>   // NT: columns of the small operand per launch, padded (16, 32, 64)
__global__ void __launch_bounds__(SK_THREADS, 1) c64_tf32x3_stem_kernel(const StemTcArgs p, const __grid_constant__ CUtensorMap tmA) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nkb = (uint32_t)p.K / TC_BK;
    const int64_t ntiles = p.M / TC_BM;
    static_assert(!WIDE64 || NT == 64, "WIDE64 is the NT = 64 variant");
    static_assert(NT != 128 || BSTREAM, "128-column passes exist in streamed-B mode only");
    constexpr bool F6 = NT <= 32 || WIDE64;                          // 6 wide MMAs into F/E, else 12 into re/im
    constexpr uint32_t NSETS = (WIDE64 || NT == 128) ? 1 : 2;        // accumulator sets in tensor memory
    constexpr int BST = NT == 128 ? SK_BST / 2 : SK_BST;             // B ring stages (16 KB each at NT = 128)
    constexpr int B_KB = 2 * NT * TC_BK * 4;                         // bytes of one B plane ([re | im] rows) per k-block
    const uint32_t b_plane = nkb * B_KB;                             // bytes of one B plane (all k), resident mode

    const int SK_RAW = p.raw_stages;
    uint16_t* tab16 = reinterpret_cast<uint16_t*>(smem + p.off_tab); // general: swizzled rank of (col, row)
    uint32_t* tab32 = reinterpret_cast<uint32_t*>(smem + p.off_tab); // additive: swizzled staging byte offset of column c [NT]
    int32_t* runbase = reinterpret_cast<int32_t*>(smem + p.off_run);
    const int nruns = (TC_BM * p.N) >> p.run_shift;
    bool runs_in_smem = nruns <= SK_RUNS_MAX;
    const uint32_t bar0 = smem_u32(smem + p.off_bar);
    auto raw_full = [&](int s) { return bar0 + 8u * s; };
    auto raw_empty = [&](int s) { return bar0 + 8u * (SK_RAW_MAX + s); };
    auto apl_full = [&](int s) { return bar0 + 8u * (2 * SK_RAW_MAX + s); };
    auto apl_empty = [&](int s) { return bar0 + 8u * (2 * SK_RAW_MAX + SK_PL_MAX + s); };
    auto accfull_bar = [&](int s) { return bar0 + 8u * (2 * SK_RAW_MAX + 2 * SK_PL_MAX + s); };
    auto accempty_bar = [&](int s) { return bar0 + 8u * (2 * SK_RAW_MAX + 2 * SK_PL_MAX + 2 + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (2 * SK_RAW_MAX + 2 * SK_PL_MAX + 4 + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (2 * SK_RAW_MAX + 2 * SK_PL_MAX + 4 + SK_BST + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.off_bar + SK_NBARS * 8);
    constexpr uint32_t SET_COLS = F6 ? 4 * NT : 2 * NT;              // F | E (2N each), or re | im (N each)
    constexpr uint32_t APL_COL0 = NSETS * SET_COLS;                  // A plane ring behind the accumulator set(s)
    constexpr uint32_t USED_COLS = APL_COL0 + SK_PL_MAX * SK_APL_COLS;
    constexpr uint32_t TMEM_COLS = USED_COLS <= 256 ? 256 : 512;

    if (tid == 0) {
        for (int s = 0; s < SK_RAW; s++) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), SK_WORKERS / 64); }
        for (int s = 0; s < SK_PL_MAX; s++) { mbar_init(apl_full(s), SK_WORKERS / 64); mbar_init(apl_empty(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(accfull_bar(s), 1); mbar_init(accempty_bar(s), SK_EPI / 32 / p.egroups); }
        for (int s = 0; s < BST; s++) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    // small operand -> resident planes (all k-blocks): plane q at q*b_plane, k-block kb at kb*B_KB
    for (uint32_t u = tid; !BSTREAM && u < (uint32_t)NT * nkb * 2; u += SK_THREADS) {
        const uint32_t row = u % NT, r = u / NT, kc = r & 1, kb = r >> 1;
        float2 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t k = kb * TC_BK + kc * 4 + i;
            v[i] = (int)row < p.N ? p.B[tabc(p.bn, p.n0 + row) + tabc(p.bk, k)] : make_float2(0.f, 0.f);
        }
        split_store_b(smem + kb * B_KB, (int)b_plane, NT, (int)row, (int)kc, v, p.conjB);
    }
    if (p.direct) {
        // column address deltas (the address of (row, col) is rel[pos[row*N]] + (rel[pos[col]] - rel[pos[0]]))
        int32_t* coltab = reinterpret_cast<int32_t*>(smem + p.off_tab);
        const int64_t a0 = p.rel[p.pos[0]];
        for (int i = tid; i < NT; i += SK_THREADS) coltab[i] = i < p.N ? (int32_t)(p.rel[p.pos[i]] - a0) : 0;
    } else if (p.additive) {
        // separable rank on disjoint bits: staging byte offset of (row, col) = rowoff ^ coloff (the swizzle is XOR-linear)
        const int64_t p0 = p.pos[0];
        for (int i = tid; i < NT; i += SK_THREADS) tab32[i] = i < p.N ? 8u * sk_swz((uint32_t)(p.pos[i] - p0)) : 0u;
    } else {
        // rank table transposed to [col][row]: the 32 lanes of an epilogue warp (consecutive rows) read consecutive entries
        for (int i = tid; i < TC_BM * p.N; i += SK_THREADS) {
            const int row = i / p.N, col = i - row * p.N;
            tab16[col * TC_BM + row] = (uint16_t)sk_swz((uint32_t)p.pos[i]);
        }
    }
    {
        int big = 0;
        if (runs_in_smem && !p.direct)
            for (int i = tid; i < nruns; i += SK_THREADS) {
                const int64_t v = p.rel[(int64_t)i << p.run_shift];
                if (v < 0 || v >= ((int64_t)1 << 31)) big = 1;
                runbase[i] = (int32_t)v;
            }
        fence_proxy_async_smem();
        tc_fence_before();
        if (__syncthreads_or(big)) runs_in_smem = false;             // pattern spans > 2^31 elements: bases from global
    }
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ---- workers: raw A tile -> four planes in TENSOR MEMORY.  Two groups of 4 warps take alternate k-blocks
        // (group = parity of the global k-block counter), so two latency chains (wait, LDS, split, tcgen05.st, arrive)
        // run concurrently; a thread owns one row (= its TMEM lane) and all 8 k of its k-block. ----
        const int group = warp >> 2, prow = tid & 127;
        const int raw_row = p.off_raw + prow * 8;
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + APL_COL0;
        int64_t my_tiles = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) my_tiles++;
        const uint64_t total_kb = (uint64_t)my_tiles * nkb;
        for (uint64_t g = group; g < total_kb; g += 2) {
            const int rs = (int)(g % (uint64_t)SK_RAW);
            const int ps = (int)(g % (uint64_t)SK_PL_MAX);      // group g%2 owns stages {group, group+2}
            const uint32_t rphase = (uint32_t)((g / (uint64_t)SK_RAW) & 1), pphase = (uint32_t)((g / (uint64_t)SK_PL_MAX) & 1);
            mbar_wait(raw_full(rs), rphase);
            const uint8_t* raw = smem + rs * SK_RAW_STAGE + raw_row;
            uint32_t pl[32];                                     // rh[8] | rl[8] | ih[8] | il[8]
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float2 v = *reinterpret_cast<const float2*>(raw + i * TC_BM * 8);
                const float re = v.x, im = p.conjA ? -v.y : v.y;
                const float rh = tf32_hi(re), ih = tf32_hi(im);
                pl[i] = __float_as_uint(rh); pl[8 + i] = __float_as_uint(re - rh);
                pl[16 + i] = __float_as_uint(ih); pl[24 + i] = __float_as_uint(im - ih);
            }
            mbar_wait(apl_empty(ps), pphase ^ 1);
            tc_fence_after();
            tmem_st32(lane_base + ps * SK_APL_COLS, pl);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(raw_empty(rs));                     // released after the tcgen05.st consumed the loaded values (an arrive
                mbar_arrive(apl_full(ps));                      // right after the ld.shared does not wait for them to return)
            }
        }
    } else if (warp < 16) {
        // ---- epilogue warps: TMEM -> (combine) -> staging (rank order) -> C (ascending addresses) ----
        // Two epilogue GROUPS (p.egroups == 2, forms with two TMEM sets): warps 8-11 take the even tiles of this CTA (set 0,
        // staging tile 0), warps 12-15 the odd ones, each warp draining ALL columns of its lane quarter — the drain + staging
        // of one tile overlaps the write-out of the previous one (one group: drain -> write-out was a serial chain per tile,
        // the limiter of the write-heavy small steps, r1 verdict).  One group (p.egroups == 1): 8 warps, two per lane quarter.
        const int G = p.egroups, EG = SK_EPI / G;                   // threads per group
        const int grp = G == 2 ? (warp - 8) >> 2 : 0;
        const int q = warp & 3, half = G == 2 ? 0 : (warp - 8) >> 2;
        const int etid = tid - SK_WORKERS - grp * EG;              // 0..EG-1
        const uint32_t row = q * 32 + lane;
        const int stg_bytes = (TC_BM * p.N * 8 + 1023) / 1024 * 1024;
        float2* stg = reinterpret_cast<float2*>(smem + p.off_stg + grp * stg_bytes);
        const bool has_beta = p.beta[0] != 0.f || p.beta[1] != 0.f;
        const float ar = p.alpha[0], ai = p.alpha[1], br = p.beta[0], bi = p.beta[1];
        const bool unit_alpha = ar == 1.f && ai == 0.f;
        const int cnt = TC_BM * p.N;
        const int COLS = (NT >= 32 && G == 1) ? NT / 2 : NT;       // columns per warp (NT = 16, one group: warps of half 1 skip the drain)
        const int cbeg = (NT >= 32 && G == 1) ? half * COLS : 0;
        const bool drains = NT >= 32 || half == 0;
        const bool additive = p.additive != 0;
        // additive: this thread's row contributes a fixed (swizzled) byte offset
        const uint32_t rowoff = additive ? 8u * sk_swz((uint32_t)p.pos[(int64_t)row * p.N]) : 0u;
        uint8_t* stg_b = smem + p.off_stg + grp * stg_bytes;
        const int rmask = (1 << p.run_shift) - 1;
        const int cend = (cbeg + COLS) < p.N ? (cbeg + COLS) : p.N;   // N is a multiple of 16 (eligibility)
        // write-out: pair j = 2*etid + 512*it; swz is XOR-linear and the two parts use disjoint bits
        const uint32_t swz_t = sk_swz((uint32_t)etid * 2u);
        uint32_t i = 0;
        uint32_t echunk0 = 0, echunk1 = 0;                          // chunks drained so far, per accumulator set
        const uint32_t nchunks = (nkb + SK_KCB - 1) / SK_KCB;
        const int64_t t0 = blockIdx.x + (int64_t)grp * gridDim.x, tstep = (int64_t)G * gridDim.x;
        int64_t hi_next = t0 < ntiles ? p.hi[t0] : 0;
        i = (uint32_t)grp;
        if (p.direct) {
            // ---- direct epilogue: TMEM -> registers -> C.  A thread owns one row; a warp instruction stores one column of
            // 32 consecutive rows, which the planner has checked to fill whole 32-byte sectors.  No staging tile, no group
            // barrier: the shared-memory pipe carries only the raw A tiles (it was the limiter of these steps: 64-72 % busy
            // with the staging round trip, profiles/r2_ncu_stalls.md) and each warp runs on its own. ----
            const int32_t* coltab = reinterpret_cast<const int32_t*>(smem + p.off_tab);
            const int64_t rowaddr = p.rel[p.pos[(int64_t)row * p.N]];
            for (int64_t t = t0; t < ntiles; t += tstep, i += (uint32_t)G) {
                const uint32_t set = i % NSETS;
                const int64_t hi_cur = hi_next;
                if (t + tstep < ntiles) hi_next = p.hi[t + tstep];
                mbar_wait(accfull_bar(set), (set ? echunk1 : echunk0) & 1);
                if (set) echunk1++; else echunk0++;
                tc_fence_after();
                if (drains) {
                    float2* rbase = p.C + hi_cur + rowaddr;
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * SET_COLS;
#pragma unroll 1
                    for (int c0 = cbeg; c0 < cend; c0 += 16) {
                        uint32_t xr[16], xi[16];
                        if (F6) {
                            uint32_t er[16], ei[16];
                            tmem_ld16(taddr + c0, xr);
                            tmem_ld16(taddr + NT + c0, xi);
                            tmem_ld16(taddr + 2 * NT + c0, er);
                            tmem_ld16(taddr + 3 * NT + c0, ei);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                xr[j] = __float_as_uint(__uint_as_float(xr[j]) - __uint_as_float(ei[j]));
                                xi[j] = __float_as_uint(__uint_as_float(xi[j]) + __uint_as_float(er[j]));
                            }
                        } else {
                            tmem_ld16(taddr + c0, xr);
                            tmem_ld16(taddr + NT + c0, xi);
                            tmem_ld_wait();
                        }
#pragma unroll
                        for (int j4 = 0; j4 < 16; j4 += 4) {
                            const int4 ct = *reinterpret_cast<const int4*>(coltab + c0 + j4);
                            const int32_t cd[4] = {ct.x, ct.y, ct.z, ct.w};
#pragma unroll
                            for (int jj = 0; jj < 4; jj++) {
                                const int j = j4 + jj;
                                float2 v = make_float2(__uint_as_float(xr[j]), __uint_as_float(xi[j]));
                                if (!unit_alpha) v = make_float2(ar * v.x - ai * v.y, ar * v.y + ai * v.x);
                                float2* dst = rbase + cd[jj];
                                if (has_beta) {
                                    const float2 old = *dst;
                                    v.x += br * old.x - bi * old.y;
                                    v.y += br * old.y + bi * old.x;
                                }
                                *dst = v;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(accempty_bar(set));
            }
        } else
        for (int64_t t = t0; t < ntiles; t += tstep, i += (uint32_t)G) {
            const uint32_t set = i % NSETS;
            const int64_t hi_cur = hi_next;
            if (t + tstep < ntiles) hi_next = p.hi[t + tstep];          // in flight while this tile drains
            // K > 128 (128-column form only): the accumulators hold one chunk of SK_KCB k-blocks at a time; chunk 0 is
            // stored into the staging tile, later chunks are added to it round-to-nearest (each thread owns its entries)
            for (uint32_t ch = 0; ch < nchunks; ch++) {
            const bool rmw = ch > 0;
            mbar_wait(accfull_bar(set), (set ? echunk1 : echunk0) & 1);
            if (set) echunk1++; else echunk0++;
            tc_fence_after();
            if (drains) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * SET_COLS;
#pragma unroll 1
                for (int c0 = cbeg; c0 < cend; c0 += 16) {
                    uint32_t xr[16], xi[16];
                    if (F6) {
                        uint32_t er[16], ei[16];
                        tmem_ld16(taddr + c0, xr);               // A_re.B_re
                        tmem_ld16(taddr + NT + c0, xi);          // A_re.B_im
                        tmem_ld16(taddr + 2 * NT + c0, er);      // A_im.B_re
                        tmem_ld16(taddr + 3 * NT + c0, ei);      // A_im.B_im
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            xr[j] = __float_as_uint(__uint_as_float(xr[j]) - __uint_as_float(ei[j]));
                            xi[j] = __float_as_uint(__uint_as_float(xi[j]) + __uint_as_float(er[j]));
                        }
                    } else {
                        tmem_ld16(taddr + c0, xr);
                        tmem_ld16(taddr + NT + c0, xi);
                        tmem_ld_wait();
                    }
                    if (additive) {
                        uint32_t co[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const uint4 c4 = *reinterpret_cast<const uint4*>(tab32 + c0 + j);
                            co[j] = c4.x; co[j + 1] = c4.y; co[j + 2] = c4.z; co[j + 3] = c4.w;
                        }
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float2* d = reinterpret_cast<float2*>(stg_b + (rowoff ^ co[j]));
                            float2 v = make_float2(__uint_as_float(xr[j]), __uint_as_float(xi[j]));
                            if (rmw) { const float2 o = *d; v.x += o.x; v.y += o.y; }
                            *d = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float2* d = stg + tab16[(c0 + j) * TC_BM + row];
                            float2 v = make_float2(__uint_as_float(xr[j]), __uint_as_float(xi[j]));
                            if (rmw) { const float2 o = *d; v.x += o.x; v.y += o.y; }
                            *d = v;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty_bar(set));          // TMEM set free: the next chunk / tile may start
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(EG) : "memory");   // staging tile complete (this group's warps only)
            float2* base = p.C + hi_cur;
            if (p.vec2) {
                constexpr int U = 4;                                 // pairs in flight per thread; cnt is a multiple of 2048
                for (int j0 = etid * 2; j0 < cnt; j0 += 2 * EG * U) {
                    float4 w[U];
                    bool od[U];                                      // pair stored in swapped order: bit 0 of the swizzled rank.  With
                                                                     // 256 threads it is a thread constant (offsets are multiples of
                                                                     // 512); with two groups of 128 the 256-offsets flip it
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const uint32_t s = swz_t ^ sk_swz((uint32_t)(j0 - etid * 2 + u * 2 * EG));
                        w[u] = *reinterpret_cast<const float4*>(stg + (s & ~1u));
                        od[u] = (s & 1u) != 0;
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int j = j0 + u * 2 * EG;
                        float2 v0 = od[u] ? make_float2(w[u].z, w[u].w) : make_float2(w[u].x, w[u].y);
                        float2 v1 = od[u] ? make_float2(w[u].x, w[u].y) : make_float2(w[u].z, w[u].w);
                        if (!unit_alpha) {
                            v0 = make_float2(ar * v0.x - ai * v0.y, ar * v0.y + ai * v0.x);
                            v1 = make_float2(ar * v1.x - ai * v1.y, ar * v1.y + ai * v1.x);
                        }
                        const int run = j >> p.run_shift;
                        const int64_t rb = runs_in_smem ? (int64_t)runbase[run] : p.rel[(int64_t)run << p.run_shift];
                        float4* dst = reinterpret_cast<float4*>(base + rb + (j & rmask));
                        if (has_beta) {
                            const float4 old = *dst;
                            v0.x += br * old.x - bi * old.y; v0.y += br * old.y + bi * old.x;
                            v1.x += br * old.z - bi * old.w; v1.y += br * old.w + bi * old.z;
                        }
                        *dst = make_float4(v0.x, v0.y, v1.x, v1.y);
                    }
                }
            } else {
#pragma unroll 4
                for (int j = etid; j < cnt; j += EG) {
                    float2 v = stg[sk_swz((uint32_t)j)];
                    float2 o = make_float2(ar * v.x - ai * v.y, ar * v.y + ai * v.x);
                    const int run = j >> p.run_shift;
                    const int64_t rb = runs_in_smem ? (int64_t)runbase[run] : p.rel[(int64_t)run << p.run_shift];
                    float2* dst = base + rb + (j & rmask);
                    if (has_beta) {
                        float2 old = *dst;
                        o.x += br * old.x - bi * old.y;
                        o.y += br * old.y + bi * old.x;
                    }
                    *dst = o;
                }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(EG) : "memory");   // staging tile free again
        }
    } else if (warp == 16) {
        // ---- MMA issuer: A planes from tensor memory, B planes from shared memory.  The whole warp runs the loop
        // (waits included); one elected lane issues the MMAs and commits. ----
        {
            constexpr uint32_t IDESC6 = make_idesc<2 * NT>(false);
            constexpr uint32_t IDESC = make_idesc<NT>(false), IDESC_NEG = make_idesc<NT>(true);
            const uint64_t bdesc0 = make_smem_desc(smem_u32(smem));
            const uint64_t bl_off = BSTREAM ? (uint64_t)(B_KB >> 4) : (uint64_t)(b_plane >> 4);
            constexpr uint64_t IM_OFF = (uint64_t)(NT * TC_BK * 4) >> 4;   // rows [N, 2N) of a plane = B_im
            const uint32_t a0 = tmem_base + APL_COL0;
            int ps = 0, bs = 0;
            uint32_t pphase = 0, bphase = 0, i = 0;
            uint32_t mchunk0 = 0, mchunk1 = 0;                       // chunks started so far, per accumulator set
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, i++) {
                const uint32_t set = i % NSETS;
                const uint32_t d0 = tmem_base + set * SET_COLS;
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    const bool cfirst = (kb % SK_KCB) == 0, clast = (kb % SK_KCB) == SK_KCB - 1 || kb + 1 == nkb;
                    if (cfirst) {                                    // the set must have been drained (previous chunk / tile)
                        const uint32_t mc = set ? mchunk1 : mchunk0;
                        if (mc >= 1) mbar_wait(accempty_bar(set), (mc - 1) & 1);
                        if (set) mchunk1++; else mchunk0++;
                    }
                    if (BSTREAM) mbar_wait(b_full(bs), bphase);
                    mbar_wait(apl_full(ps), pphase);
                    tc_fence_after();
                    const uint32_t a_rh = a0 + ps * SK_APL_COLS, a_rl = a_rh + 8, a_ih = a_rh + 16, a_il = a_rh + 24;
                    const uint64_t b_h = bdesc0 + (uint64_t)((BSTREAM ? 2u * bs : kb) * (B_KB >> 4)), b_l = b_h + bl_off;
                    const uint32_t acc = cfirst ? 0u : 1u;
                    if (elect_one()) {
                        if (F6) {
                            const uint32_t d_f = d0, d_e = d0 + 2 * NT;
                            umma_tf32_ts(d_f, a_rh, b_l, IDESC6, acc);
                            umma_tf32_ts(d_f, a_rl, b_h, IDESC6, 1u);
                            umma_tf32_ts(d_f, a_rh, b_h, IDESC6, 1u);
                            umma_tf32_ts(d_e, a_ih, b_l, IDESC6, acc);
                            umma_tf32_ts(d_e, a_il, b_h, IDESC6, 1u);
                            umma_tf32_ts(d_e, a_ih, b_h, IDESC6, 1u);
                        } else {
                            const uint32_t d_re = d0, d_im = d0 + NT;
                            const uint64_t b_rh = b_h, b_ih = b_h + IM_OFF, b_rl = b_l, b_il = b_l + IM_OFF;
                            umma_tf32_ts(d_re, a_rh, b_rl, IDESC, acc);
                            umma_tf32_ts(d_re, a_rl, b_rh, IDESC, 1u);
                            umma_tf32_ts(d_re, a_ih, b_il, IDESC_NEG, 1u);
                            umma_tf32_ts(d_re, a_il, b_ih, IDESC_NEG, 1u);
                            umma_tf32_ts(d_re, a_rh, b_rh, IDESC, 1u);
                            umma_tf32_ts(d_re, a_ih, b_ih, IDESC_NEG, 1u);
                            umma_tf32_ts(d_im, a_rh, b_il, IDESC, acc);
                            umma_tf32_ts(d_im, a_rl, b_ih, IDESC, 1u);
                            umma_tf32_ts(d_im, a_ih, b_rl, IDESC, 1u);
                            umma_tf32_ts(d_im, a_il, b_rh, IDESC, 1u);
                            umma_tf32_ts(d_im, a_rh, b_ih, IDESC, 1u);
                            umma_tf32_ts(d_im, a_ih, b_rh, IDESC, 1u);
                        }
                        umma_commit(apl_empty(ps));
                        if (BSTREAM) umma_commit(b_empty(bs));
                        if (clast) umma_commit(accfull_bar(set));
                    }
                    __syncwarp();
                    if (++ps == SK_PL_MAX) { ps = 0; pphase ^= 1; }
                    if (BSTREAM && ++bs == BST) { bs = 0; bphase ^= 1; }
                }
            }
        }
    } else {
        // ---- copy issuer: lane 0 fetches the A tile of every k-block with one 2-D TMA load; lane 8 streams the B planes.
        // The two loops are INDEPENDENT (divergent on purpose): the A ring runs ahead by its full depth. ----
        if (lane == 0) {
            // one 2-D TMA tile load (128 rows x 8 k) per k-block
            int rs = 0;
            uint32_t rphase = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    mbar_wait(raw_empty(rs), rphase ^ 1);
                    mbar_expect_tx(raw_full(rs), SK_RAW_STAGE);
                    tma_load_2d(smem_u32(smem + p.off_raw + rs * SK_RAW_STAGE), &tmA, (uint32_t)(t * TC_BM), kb * TC_BK, raw_full(rs));
                    if (++rs == SK_RAW) { rs = 0; rphase ^= 1; }
                }
            }
        } else if (BSTREAM && lane == 8) {
            int bs = 0;
            uint32_t bphase = 0;
            const uint8_t* bsrc = p.bplanes + (size_t)(blockIdx.x % SK_BREP) * p.brep_stride;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
                for (uint32_t kb = 0; kb < nkb; kb++) {
                    mbar_wait(b_empty(bs), bphase ^ 1);
                    mbar_expect_tx(b_full(bs), 2 * B_KB);
                    bulk_g2s(smem_u32(smem + bs * 2 * B_KB), bsrc + (size_t)kb * 2 * B_KB, 2 * B_KB, b_full(bs));
                    if (++bs == BST) { bs = 0; bphase ^= 1; }
                }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// pre-pass of the streamed-B mode: split the small operand once into global memory, [pass][kb][hi|lo][2*NT rows x 32 B]
template <int NT>
__global__ void stem_bsplit_kernel(const float2* __restrict__ B, TabRef bn, TabRef bk, int N, int conjB, uint8_t* __restrict__ out) {
    constexpr int B_KB = 2 * NT * TC_BK * 4;
    const uint32_t kb = blockIdx.x, pass = blockIdx.y, nkb = gridDim.x, npass = gridDim.y;
    uint8_t* dst = out + (((size_t)blockIdx.z * npass + pass) * nkb + kb) * 2 * B_KB;   // replica z
    for (uint32_t u = threadIdx.x; u < (uint32_t)NT * 2; u += blockDim.x) {
        const uint32_t row = u % NT, kc = u / NT;
        float2 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t k = kb * TC_BK + kc * 4 + i;
            v[i] = (int)row < N ? B[tabc(bn, pass * NT + row) + tabc(bk, k)] : make_float2(0.f, 0.f);
        }
        split_store_b(dst, B_KB, NT, (int)row, (int)kc, v, conjB);
    }
}

template <int NT, bool BSTREAM, bool WIDE64 = false>
int launch_stem_tc(tnb_ctx* ctx, StemTcArgs a) {
    // two epilogue groups whenever the form has two TMEM sets and the second staging tile leaves >= 6 raw stages
    // (TNB_STEM_EGROUPS=1: one group, comparison)
    static const int eg_env = [] { const char* e = getenv("TNB_STEM_EGROUPS"); return e ? atoi(e) : 2; }();
    static const int direct_env = [] { const char* e = getenv("TNB_STEM_DIRECT"); return e ? atoi(e) : 1; }();   // 0: always staged (comparison / tests)
    if (!direct_env || NT > 64 || a.K > 128) a.direct = 0;
    a.egroups = (NT <= 64 && !WIDE64 && eg_env == 2) ? 2 : 1;
    if (a.egroups == 2 && sk_layout(NT, a) < 6) a.egroups = 1;
    if (sk_layout(NT, a) < 3) return -1;
    const int smem = a.off_bar + SK_NBARS * 8 + 16;
    TNB_CUDA_CHECK(ctx, cudaFuncSetAttribute(c64_tf32x3_stem_kernel<NT, BSTREAM, WIDE64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_BUDGET));
    int64_t grid = a.M / TC_BM;
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    CUtensorMap tmA;
    if (!make_operand_map(&tmA, a.A, (uint64_t)a.M, (uint64_t)a.K, a.lda))
        return tnb_set_error(ctx, TNB_ECUDA, "cuTensorMapEncodeTiled failed (M=%lld K=%d lda=%lld)", (long long)a.M, a.K, (long long)a.lda);
    c64_tf32x3_stem_kernel<NT, BSTREAM, WIDE64><<<(unsigned)grid, SK_THREADS, smem, ctx->stream>>>(a, tmA);
    ctx->launches++;
    TNB_CUDA_CHECK(ctx, cudaGetLastError());
    return TNB_OK;
}

// cp.async.bulk needs 16-byte aligned rows: even M, N and leading dimensions, 16-byte aligned bases
bool tc_acc_ok(const TcArgs& a) {
    return (a.M % 2) == 0 && (a.N % 2) == 0 && (a.lda % 2) == 0 && (a.ldb % 2) == 0 &&
           ((uintptr_t)a.A % 16) == 0 && ((uintptr_t)a.B % 16) == 0;
}

}  // namespace

// Plan-time eligibility of the tcgen05 kernels: complex64, no batch, both operands dense with the free index fastest.
// Returns the tile class: 256 / 128 = aligned to the fast kernel's tiles as well, 1 = chunked kernel only (ragged), 0 = no.
int tnb_tc_c64_tile(int64_t M, int64_t N, int64_t K, int64_t L, bool a_mmajor, bool b_nmajor) {
    if (L != 1 || !a_mmajor || !b_nmajor) return 0;
    if (K < 16 || (M % 2) || (N % 2) || M < 32 || N < 32) return 0;
    if (M >= ((int64_t)1 << 27) || N >= ((int64_t)1 << 27)) return 0;
    if (M % TC_BM == 0 && N % 256 == 0) return 256;
    if (M % TC_BM == 0 && N % 128 == 0) return 128;
    return 1;
}

// split-K for the chunked kernel: few output tiles, long K.  Returns the split count; kb_per_split in k-blocks of 8.
int tnb_tc_c64_splitk(const tnb_ctx* ctx, int64_t M, int64_t N, int64_t K, int64_t* kb_per_split, int64_t* ws_elems) {
    const int64_t nkb = (K + TC_BK - 1) / TC_BK;
    *kb_per_split = nkb;
    *ws_elems = 0;
    const int64_t tiles = ((M + TC_BM - 1) / TC_BM) * ((N + ACC_NT - 1) / ACC_NT);
    const int64_t sms = ctx ? ctx->sm_count : 148;
    if (tiles * 2 > sms || nkb < 4 * ACC_KCB) return 1;
    int64_t s = (2 * sms) / tiles;
    const int64_t maxs = nkb / (2 * ACC_KCB);                 // at least 128 k per split
    if (s > maxs) s = maxs;
    while (s > 1 && s * M * N > ((int64_t)1 << 25)) s--;
    if (s <= 1) return 1;
    int64_t per = (nkb + s - 1) / s;
    per = (per + ACC_KCB - 1) / ACC_KCB * ACC_KCB;
    s = (nkb + per - 1) / per;
    if (s <= 1) return 1;
    *kb_per_split = per;
    *ws_elems = s * M * N;
    return (int)s;
}

// returns TNB_OK, or -1 if the operands are not 16-byte aligned for the bulk copies (caller falls back)
int tnb_launch_c64_tc(tnb_ctx* ctx, const EinsumArgs& e, int nt, int64_t lda, int64_t ldb, bool chunked) {
    TcArgs a;
    a.A = (const float2*)e.A; a.B = (const float2*)e.B; a.C = (float2*)e.C;
    a.M = (uint32_t)e.M; a.N = (uint32_t)e.N; a.K = (uint32_t)e.K;
    a.lda = lda; a.ldb = ldb;
    a.cm = e.cm; a.cn = e.cn;
    a.conjA = e.conjA; a.conjB = e.conjB;
    a.alpha[0] = (float)e.alpha[0]; a.alpha[1] = (float)e.alpha[1];
    a.beta[0] = (float)e.beta[0]; a.beta[1] = (float)e.beta[1];
    a.ws = (float2*)e.ws;
    a.st_hi = a.st_rel = a.st_pos = nullptr;
    a.st_run_shift = a.st_vec2 = a.st_rel_small = a.st_direct = 0;
    a.splitk = e.splitk > 1 ? (uint32_t)e.splitk : 1u;
    a.kb_per_split = e.splitk > 1 ? (uint32_t)e.kchunk : (uint32_t)((e.K + TC_BK - 1) / TC_BK);
    if (!chunked && a.splitk == 1 && nt >= 128) {
        if (nt == 256) return launch_tc<256>(ctx, a);
        return launch_tc<128>(ctx, a);
    }
    if (!tc_acc_ok(a)) return -1;
    if (ctx->gemm_pair && a.M > (uint32_t)TC_BM) return launch_tc_pair<false>(ctx, a);   // CTA pairs on 256 x 128 tiles
    return launch_tc_acc(ctx, a);
}

// Plan-time eligibility of the persistent tensor-core stem kernel (sizes only; density is checked by the planner).
// Small operand resident in shared memory (N*K*16 <= 64 KB), or streamed per k-block from a pre-split copy (64 columns
// per pass, K <= 128: the regime where the tile GEMM kernel's per-tile prologue/epilogue is not amortised).
bool tnb_stem_tc_shape_ok(int64_t Mbig, int64_t Nsmall, int64_t K) {
    if (Nsmall == 128) return Mbig >= 65536 && Mbig % TC_BM == 0 && K >= 64 && K <= 512 && K % TC_BK == 0;   // streamed, one set, TMEM chunks of 128 k
    if (!(Mbig >= 65536 && Mbig % TC_BM == 0 && Nsmall >= 16 && Nsmall <= 64 && Nsmall % 16 == 0 && K >= 8 && K % TC_BK == 0)) return false;
    if (K <= 128 && Nsmall * K * 16 <= 64 * 1024) return true;
    return Nsmall == 64 && K <= 128;
}
// workspace (in complex64 elements) the streamed-B mode needs for `npass` passes of `Nsmall` columns; 0 = resident mode
int64_t tnb_stem_tc_ws_elems(int64_t Nsmall, int64_t K, int64_t npass) {
    if (Nsmall * K * 16 <= 64 * 1024) return 0;
    return SK_BREP * npass * (K / TC_BK) * (128 * Nsmall) / 8;
}

// TNB_STEM_WIDE64=1: 64-column passes use the single-set 6-MMA form of the stem kernel (see WIDE64 above)
static bool stem_wide64() {
    static const int on = [] { const char* e = getenv("TNB_STEM_WIDE64"); return e ? atoi(e) : 0; }();
    return on != 0;
}

// returns TNB_OK, or -1 when the big operand is not 16-byte aligned / has an odd leading dimension (caller falls back).
// ws: streamed-B workspace for ALL passes (pass e.n0 / e.N uses its slice); the pre-split runs with the first pass.
int tnb_launch_c64_stem_tc(tnb_ctx* ctx, const StemArgs& e, void* ws, int npass) {
    if ((e.lda % 2) != 0 || ((uintptr_t)e.A % 16) != 0) return -1;
    StemTcArgs a;
    memset(&a, 0, sizeof a);
    a.A = (const float2*)e.A; a.B = (const float2*)e.B; a.C = (float2*)e.C;
    a.M = e.M; a.lda = e.lda; a.N = e.N; a.K = e.K; a.n0 = e.n0; a.conjA = e.conjA; a.conjB = e.conjB;
    a.bn = e.bn; a.bk = e.bk; a.hi = e.hi; a.rel = e.rel; a.pos = e.pos;
    a.run_shift = 0;
    while ((1 << (a.run_shift + 1)) <= e.run) a.run_shift++;
    a.additive = e.additive;
    a.direct = e.direct;
    a.vec2 = (e.run >= 2 && e.even && ((uintptr_t)e.C % 16) == 0) ? 1 : 0;
    a.alpha[0] = (float)e.alpha[0]; a.alpha[1] = (float)e.alpha[1];
    a.beta[0] = (float)e.beta[0]; a.beta[1] = (float)e.beta[1];
    const bool stream = tnb_stem_tc_ws_elems(e.N, e.K, 1) > 0;
    if (stream) {
        if (!ws || (e.N != 64 && e.N != 128)) return -1;
        const int nkb = e.K / TC_BK;
        const size_t pass_bytes = (size_t)nkb * 128 * e.N;
        if (e.n0 == 0) {
            if (e.N == 128)
                stem_bsplit_kernel<128><<<dim3(nkb, npass, SK_BREP), 128, 0, ctx->stream>>>(a.B, a.bn, a.bk, e.N, e.conjB, (uint8_t*)ws);
            else
                stem_bsplit_kernel<64><<<dim3(nkb, npass, SK_BREP), 128, 0, ctx->stream>>>(a.B, a.bn, a.bk, e.N, e.conjB, (uint8_t*)ws);
            ctx->launches++;
            TNB_CUDA_CHECK(ctx, cudaGetLastError());
        }
        a.bplanes = (const uint8_t*)ws + (size_t)(e.n0 / e.N) * pass_bytes;
        a.brep_stride = (int64_t)npass * pass_bytes;
        if (e.N == 128) return launch_stem_tc<128, true>(ctx, a);
        return stem_wide64() ? launch_stem_tc<64, true, true>(ctx, a) : launch_stem_tc<64, true>(ctx, a);
    }
    if (e.N <= 16) return launch_stem_tc<16, false>(ctx, a);
    if (e.N <= 32) return launch_stem_tc<32, false>(ctx, a);
    return stem_wide64() ? launch_stem_tc<64, false, true>(ctx, a) : launch_stem_tc<64, false>(ctx, a);
}

// Wide stem steps (huge dense operand x small DENSE operand, 128 columns per pass) on the CTA-pair kernel with the
// stem kernel's sorted-pattern epilogue: one launch covers all passes.  e: big operand as A (dense [K][M], M % 128 == 0),
// small operand as e.B with leading dimension ldb (dense [K][N], N % 128 == 0).  Returns -1 when the operands do not meet
// the alignment rules of the bulk copies (caller falls back to the 1-CTA stem kernel).
int tnb_launch_c64_pair_staged(tnb_ctx* ctx, const StemArgs& e, int64_t Nsmall, int64_t ldb, bool rel_small) {
    if (!ctx->gemm_pair || !e.additive || (e.M % TC_BM) != 0 || (Nsmall % PR_NT) != 0 || e.M <= TC_BM) return -1;
    TcArgs a;
    memset(&a, 0, sizeof a);
    a.A = (const float2*)e.A; a.B = (const float2*)e.B; a.C = (float2*)e.C;
    a.M = (uint32_t)e.M; a.N = (uint32_t)Nsmall; a.K = (uint32_t)e.K;
    a.lda = e.lda; a.ldb = ldb;
    a.conjA = e.conjA; a.conjB = e.conjB;
    a.alpha[0] = (float)e.alpha[0]; a.alpha[1] = (float)e.alpha[1];
    a.beta[0] = (float)e.beta[0]; a.beta[1] = (float)e.beta[1];
    a.splitk = 1;
    a.kb_per_split = (uint32_t)((e.K + TC_BK - 1) / TC_BK);
    a.st_hi = e.hi; a.st_rel = e.rel; a.st_pos = e.pos;
    a.st_run_shift = 0;
    while ((1 << (a.st_run_shift + 1)) <= e.run) a.st_run_shift++;
    a.st_vec2 = (e.run >= 2 && e.even && ((uintptr_t)e.C % 16) == 0) ? 1 : 0;
    a.st_rel_small = rel_small ? 1 : 0;
    a.st_direct = (e.direct && rel_small) ? 1 : 0;
    if (!tc_acc_ok(a)) return -1;
    return launch_tc_pair<true>(ctx, a);
}
