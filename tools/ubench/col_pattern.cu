// Memory-system ceiling of the COLUMN-pass access pattern, without any arithmetic: a CTA owns a tile of TS adjacent
// columns of an [L = 625][N2 = 8000] matrix, every thread fetches 25 elements a row apart (runs of TS elements per
// row), then stores 25 elements the same way.  Modes:
//   0  read int16 pairs (4 B), write cf32 (8 B)   = col_fwd_raw   (12 B per point)
//   1  read cf32, no write (one dummy word per CTA) = col_inv_argmax (8 B per point)
//   2  read cf32, write cf32                        (16 B per point)
//   3  contiguous copy of the same byte counts as mode 0 (yardstick)
// Prints GB/s per (mode, TS, CTAs per SM); compare with the measured copy peak of MEASURED_PEAKS.json.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o col_pattern col_pattern.cu && ./col_pattern
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int L = 625, N2 = 8000, PER = 25;

template <int TS, int MODE> __global__ void __launch_bounds__(PER * TS) kern(const void* src, float2* dst, int lanes, float* sink)
{
    extern __shared__ float2 sm[];
    const int ntx = N2 / TS;
    const int tile = blockIdx.x % ntx, lane = blockIdx.x / ntx;
    const int s = threadIdx.x % TS, j = threadIdx.x / TS;
    float2 v[PER];
    if (MODE == 0) {
        const short2* g = reinterpret_cast<const short2*>(src) + (size_t)lane * L * N2 + (size_t)j * N2 + tile * TS + s;
#pragma unroll
        for (int m = 0; m < PER; m++) { const short2 r = g[(size_t)m * PER * N2]; v[m] = make_float2((float)r.x, (float)r.y); }
    } else if (MODE == 3) {
        const short2* g = reinterpret_cast<const short2*>(src) + ((size_t)blockIdx.x * PER) * (PER * TS) + threadIdx.x;
#pragma unroll
        for (int m = 0; m < PER; m++) { const short2 r = g[(size_t)m * PER * TS]; v[m] = make_float2((float)r.x, (float)r.y); }
    } else {
        const float2* g = reinterpret_cast<const float2*>(src) + (size_t)lane * L * N2 + (size_t)j * N2 + tile * TS + s;
#pragma unroll
        for (int m = 0; m < PER; m++) v[m] = g[(size_t)m * PER * N2];
    }
    // one shared-memory exchange and barrier, like the transform (keeps the load and store phases apart)
#pragma unroll
    for (int m = 0; m < PER; m++) sm[(m * PER + j) * TS + s] = v[m];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < PER; m++) v[m] = sm[(j * PER + m) * TS + s];
    if (MODE == 1) {
        float a = 0.f;
#pragma unroll
        for (int m = 0; m < PER; m++) a += v[m].x * v[m].y;
        if (a == 12345.678f) sink[blockIdx.x] = a;
    } else if (MODE == 3) {
        float2* d = dst + ((size_t)blockIdx.x * PER) * (PER * TS) + threadIdx.x;
#pragma unroll
        for (int m = 0; m < PER; m++) d[(size_t)m * PER * TS] = v[m];
    } else {
        float2* d = dst + (size_t)lane * L * N2 + (size_t)j * N2 + tile * TS + s;
#pragma unroll
        for (int m = 0; m < PER; m++) d[(size_t)m * PER * N2] = v[m];
    }
}

template <int TS, int MODE> void run(const void* src, float2* dst, float* sink, int lanes, int ctas_per_sm)
{
    // occupancy through the dynamic shared-memory size (the exchange needs L*TS*8 bytes; more is requested to cap)
    size_t smem = (size_t)L * TS * sizeof(float2);
    const size_t cap = (size_t)(227 * 1024 / ctas_per_sm) - 1024;
    if (cap < smem) { return; }
    // keep exactly ctas_per_sm: request just over 227K/(ctas_per_sm+1)
    const size_t floor_ = (size_t)(227 * 1024 / (ctas_per_sm + 1)) + 1024;
    if (smem < floor_) smem = floor_ < cap ? floor_ : cap;
    cudaFuncSetAttribute(kern<TS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern<TS, MODE>, PER * TS, smem);
    const int grid = (N2 / TS) * lanes;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; w++) kern<TS, MODE><<<grid, PER * TS, smem>>>(src, dst, lanes, sink);
    cudaEventRecord(a);
    const int reps = 5;
    for (int r = 0; r < reps; r++) kern<TS, MODE><<<grid, PER * TS, smem>>>(src, dst, lanes, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    ms /= reps;
    const double per_point = MODE == 0 || MODE == 3 ? 12.0 : MODE == 1 ? 8.0 : 16.0;
    const double bytes = per_point * (double)L * N2 * lanes;
    const cudaError_t e = cudaGetLastError();
    printf("mode %d TS %2d threads %3d ctas/sm %d (occ %d) : %.3f ms  %.0f GB/s %s\n", MODE, TS, PER * TS, ctas_per_sm, occ, ms,
           bytes / ms * 1e-6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char** argv)
{
    const int lanes = argc > 1 ? atoi(argv[1]) : 32;
    void* src; float2* dst; float* sink;
    cudaMalloc(&src, (size_t)lanes * L * N2 * 8);
    cudaMalloc(&dst, (size_t)lanes * L * N2 * 8);
    cudaMalloc(&sink, 1 << 22);
    cudaMemset(src, 1, (size_t)lanes * L * N2 * 8);
    for (int c = 1; c <= 4; c++) {
        run<8, 0>(src, dst, sink, lanes, c);  run<16, 0>(src, dst, sink, lanes, c);  run<32, 0>(src, dst, sink, lanes, c);
    }
    for (int c = 1; c <= 4; c++) {
        run<8, 1>(src, dst, sink, lanes, c);  run<16, 1>(src, dst, sink, lanes, c);  run<32, 1>(src, dst, sink, lanes, c);
    }
    for (int c = 1; c <= 4; c++) {
        run<8, 2>(src, dst, sink, lanes, c);  run<16, 2>(src, dst, sink, lanes, c);  run<32, 2>(src, dst, sink, lanes, c);
    }
    for (int c = 1; c <= 4; c++) { run<16, 3>(src, dst, sink, lanes, c); run<32, 3>(src, dst, sink, lanes, c); }
    return 0;
}
