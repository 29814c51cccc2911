// Micro-probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (TS mode).
// (1) correctness of the assumed A layout: lane = row m (0..127), 8 consecutive 32-bit columns = k 0..7, written with
//     tcgen05.st.32x32b; B in shared memory (K-major no-swizzle core matrices); D = A * B^T checked on the host.
// (2) issue rate in clk/MMA for N = 16..256.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(128u >> 4) << 16;
    d |= (uint64_t)(256u >> 4) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni D;\n\tbra.uni W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int N>
__global__ void probe(const float* Ag, const float* Bg, float* Dg, long long* clk, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // B[n][k] -> core-matrix layout
    for (int i = tid; i < N * 8; i += blockDim.x) {
        const int n = i / 8, k = i % 8;
        *reinterpret_cast<float*>(smem + (n >> 3) * 256 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4) = Bg[n * 8 + k];
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    // A: thread = row, columns 256..263
    {
        uint32_t r[8];
        for (int k = 0; k < 8; k++) r[k] = __float_as_uint(Ag[tid * 8 + k]);
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + 256;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        mma_ts(tb, tb + 256, make_desc(smem_u32(smem)), IDESC, 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (blockIdx.x == 0) for (int j = 0; j < 8; j++) Dg[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // timing
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t bd = make_desc(smem_u32(smem));
        long long t0 = clock64();
        for (int i = 0; i < iters; i++) mma_ts(tb + (uint32_t)((i & 1) * N), tb + 256 + (uint32_t)((i & 3) * 8), bd, IDESC, 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 1);
        long long t1 = clock64();
        if (blockIdx.x == 0) clk[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
template <int N>
void run() {
    float hA[128 * 8], hB[N * 8];
    static float hD[128 * 256];
    for (int i = 0; i < 128 * 8; i++) hA[i] = (float)((i * 7 + 3) % 17 - 8);          // small integers: exact in tf32
    for (int i = 0; i < N * 8; i++) hB[i] = (float)((i * 5 + 1) % 13 - 6);
    float *dA, *dB, *dD; long long* dc;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    const int iters = 4096;
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    probe<N><<<148, 128, 32 * 1024>>>(dA, dB, dD, dc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
            float ref = 0;
            for (int k = 0; k < 8; k++) ref += hA[m * 8 + k] * hB[n * 8 + k];
            if (fabsf(ref - hD[m * N + n]) > 1e-3f) { if (bad < 4) printf("  mismatch m=%d n=%d got %g want %g\n", m, n, hD[m * N + n], ref); bad++; }
        }
    printf("TS N=%3d: %s (%d mismatches), %.1f clk/MMA  [%s]\n", N, bad ? "WRONG" : "ok", bad, (double)c / iters, cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc);
}
int main() { run<16>(); run<32>(); run<64>(); run<128>(); run<256>(); return 0; }
