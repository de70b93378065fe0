// Micro-benchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s)
{
    // 8 independent chains per thread
    float a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    u64 p[8];
    for (int i = 0; i < 8; i++) p[i] = ((u64)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const u64 ss = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], s, s);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], ss, ss);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = a[i] + s;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fadd2(p[i], ss);
        }
    }
    float r = 0;
    for (int i = 0; i < 16; i++) r += a[i];
    for (int i = 0; i < 8; i++) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* d, int warps_per_sm)
{
    int nsm = 148, iters = 20000;
    int blocks = nsm * warps_per_sm / 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, 256>>>(d, 100, 1.0001f);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(d, iters, 1.0001f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double lane_ops = (double)blocks * 256 * iters * 16;    // scalar-equivalent ops
    printf("%-8s warps/SM %2d: %.3f ms, %.2f T scalar-ops/s (%.1f per SM per clk @1.965GHz)\n", name, warps_per_sm, ms,
           lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 64 * 32 * 4 * 4);
    for (int w : {8, 16, 32}) {
        run<0>("FFMA", d, w); run<1>("FFMA2", d, w); run<2>("FADD", d, w); run<3>("FADD2", d, w);
    }
    return 0;
}
