// Micro-benchmark: how are the 13 warps of a 416-thread CTA spread over the four SM sub-partitions when two such
// CTAs are resident, and what does it cost?  (a) prints %warpid (hardware warp slot) of every warp of the CTAs on
// SM 0; (b) times a pure FFMA2 loop for 12 / 13 / 16 warps per CTA (2 CTAs per SM) and reports lane-ops/clk/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_slots warp_slots.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__global__ void slots(int* out)
{
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    extern __shared__ int dummy[];   // 72 KB dynamic: two CTAs per SM like the row kernels
    dummy[threadIdx.x] = 1;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        int* o = out + (blockIdx.x * 32 + (threadIdx.x >> 5)) * 3;
        o[0] = smid; o[1] = wid; o[2] = dummy[threadIdx.x];
    }
    // keep the CTA alive for a while so that co-resident CTAs overlap
    long long t0 = clock64();
    while (clock64() - t0 < 200000) {}
}
__global__ void __launch_bounds__(512, 2) spin(float* out, int iters, float s, int active_warps, int sync_every)
{
    extern __shared__ int pad[];
    pad[threadIdx.x] = 0;
    u64 p[8];
    for (int i = 0; i < 8; i++) p[i] = ((u64)__float_as_uint(threadIdx.x * 0.001f + i) << 32) | __float_as_uint(1.0f + i);
    const u64 ss = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    const bool act = (threadIdx.x >> 5) < active_warps;
    for (int it = 0; it < iters; it++) {
        if (act) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], ss, ss);
        }
        if (sync_every && (it % sync_every) == sync_every - 1) __syncthreads();
    }
    float r = 0;
    for (int i = 0; i < 8; i++) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + pad[threadIdx.x];
}
int main()
{
    int* d; cudaMalloc(&d, 296 * 32 * 3 * sizeof(int)); cudaMemset(d, 0xff, 296 * 32 * 3 * sizeof(int));
    cudaFuncSetAttribute(slots, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    slots<<<296, 416, 72 * 1024>>>(d);
    cudaDeviceSynchronize();
    static int h[296 * 32 * 3];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int sm = 0; sm < 2; sm++) {
        printf("SM %d:", sm);
        for (int b = 0; b < 296; b++)
            if (h[b * 96] == sm) {
                printf("  cta %d slots", b);
                for (int w = 0; w < 13; w++) printf(" %d", h[(b * 32 + w) * 3 + 1]);
            }
        printf("\n");
    }
    float* o; cudaMalloc(&o, 296 * 512 * sizeof(float));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4000;
    for (int sync_every = 0; sync_every <= 4; sync_every += 4)
        for (int threads : {384, 416, 512}) {
            for (int aw : {12, 13, 16}) {
                if (aw * 32 > threads) continue;
                spin<<<296, threads, 72 * 1024>>>(o, 10, 1.0001f, aw, sync_every);
                cudaEventRecord(a);
                spin<<<296, threads, 72 * 1024>>>(o, iters, 1.0001f, aw, sync_every);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                double laneops = 296.0 * aw * 32 * 2 * 64.0 * iters;   // per FFMA2: 2 lanes-ops per thread
                printf("threads %d active warps %d sync_every %d: %.3f ms  %.1f lane-ops/clk/SM (1.965 GHz)  err=%d\n", threads, aw,
                       sync_every, ms, laneops / (ms * 1e-3) / 1.965e9 / 148, (int)cudaGetLastError());
            }
        }
    return 0;
}
