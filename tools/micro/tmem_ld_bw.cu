// tmem_ld_bw.cu -- how fast can epilogue warps read TMEM?  (development microbenchmark, not product)
// 1 CTA per SM, W warps (4..16), each warp loads 32 lanes x NCOL columns per iteration from its lane quarter.
// Prints cycles per iteration and bytes/clk/SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define R8(a, o) "=r"(a[o + 0]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), "=r"(a[o + 6]), "=r"(a[o + 7])

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        : R8(r, 0), R8(r, 8)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: ld x32 + wait per iteration; mode 1: two x16 loads back to back + one wait; mode 2: x32, wait only every 2nd (two in flight)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, int warps, long long *out, uint32_t *sink)
{
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_ptr;
    const int lq = warp & 3, cq = warp >> 2;
    const uint32_t addr = base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(cq * 32);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < warps) {
        __syncwarp();
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t a = addr + (uint32_t)((i & 3) * 128);
            if (MODE == 0) {
                uint32_t r[32];
                ld32(a, r);
                wait_ld();
#pragma unroll
                for (int q = 0; q < 32; q += 8) acc ^= r[q];
            } else if (MODE == 1) {
                uint32_t r0[16], r1[16];
                ld16(a, r0);
                ld16(a + 16, r1);
                wait_ld();
#pragma unroll
                for (int q = 0; q < 16; q += 8) acc ^= r0[q] ^ r1[q];
            } else {
                uint32_t r0[32], r1[32];
                ld32(a, r0);
                ld32(a ^ 256u, r1);
                wait_ld();
#pragma unroll
                for (int q = 0; q < 32; q += 8) acc ^= r0[q] ^ r1[q];
            }
        }
        t1 = clock64();
    }
    if (lane == 0 && warp < warps) out[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 0x12345u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

int main()
{
    long long *out;
    uint32_t *sink;
    cudaMalloc(&out, 148 * 16 * sizeof(long long));
    cudaMalloc(&sink, 4);
    const int iters = 4000;
    long long h[148 * 16];
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {1, 4, 8, 12, 16}) {
            cudaMemset(out, 0, sizeof(h));
            if (mode == 0) k<0><<<148, 512>>>(iters, warps, out, sink);
            if (mode == 1) k<1><<<148, 512>>>(iters, warps, out, sink);
            if (mode == 2) k<2><<<148, 512>>>(iters, warps, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
            const double cyc = (double)mx / iters;
            const double bytes = (double)warps * 32 * 32 * 4 * (mode == 2 ? 2 : 1);
            printf("mode %d warps %2d: %.1f cycles/iter, %.1f B/clk/SM (%.0f B per iter)\n", mode, warps, cyc, bytes / cyc, bytes);
        }
    return 0;
}
