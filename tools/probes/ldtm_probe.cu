// Micro-probe: tensor-memory read bandwidth. W warps (W = 4, 8, 16) each read 64 columns of their lane quarter
// (4 x tcgen05.ld.32x32b.x16 + wait) per repetition; prints clk per repetition and bytes/clk/SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
__global__ void probe(long long* out, float* sink, int reps) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    const uint32_t taddr = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        uint32_t a[16], b[16], c[16], d[16];
        ld16(taddr, a); ld16(taddr + 16, b); ld16(taddr + 32, c); ld16(taddr + 48, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) acc += __uint_as_float(a[j]) + __uint_as_float(b[j]) + __uint_as_float(c[j]) + __uint_as_float(d[j]);
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
int main() {
    long long* d; float* s;
    cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
    for (int warps : {4, 8, 16}) {
        const int reps = 2048;
        probe<<<148, warps * 32>>>(d, s, reps);
        probe<<<148, warps * 32>>>(d, s, reps);
        long long h = 0;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double clk = (double)h / reps, bytes = warps * 32.0 * 64 * 4;
        printf("warps=%2d: %.1f clk per 64-column read per warp set, %.1f B/clk/SM (%s)\n", warps, clk, bytes / clk, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
