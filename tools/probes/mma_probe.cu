// Micro-probe: issue rate of tcgen05.mma.kind::tf32 128 x N x 8 with both operands in shared memory (SS mode),
// for N = 16..256, one CTA per SM.  Prints clocks per MMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
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
template <int N, int NA>   // NA = number of distinct A planes cycled through
__global__ void probe(long long* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
        long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const uint64_t ad = make_desc(sa + (i % NA) * 4096), bd = make_desc(sb);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb + (uint32_t)((i & 1) * N)), "l"(ad), "l"(bd), "r"(IDESC), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
template <int N, int NA>
void run(long long* d) {
    const int iters = 8192;
    cudaFuncSetAttribute(probe<N, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe<N, NA><<<148, 128, 64 * 1024>>>(d, iters);
    probe<N, NA><<<148, 128, 64 * 1024>>>(d, iters);
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("N=%3d A-planes=%d : %.1f clk/MMA (%s)\n", N, NA, (double)h / iters, cudaGetErrorString(e));
}
int main() {
    long long* d;
    cudaMalloc(&d, 8);
    run<16, 1>(d); run<16, 4>(d); run<32, 1>(d); run<32, 4>(d); run<64, 1>(d); run<64, 4>(d);
    run<128, 1>(d); run<128, 4>(d); run<256, 1>(d); run<256, 4>(d);
    return 0;
}
