// Yardstick only (never linked into the product): cuFFT batched C2C fp32 at the record length,
// to place the hand-written four-step transform.  Build: nvcc -O3 -o cufft_yardstick cufft_yardstick.cu -lcufft
#include <cstdio>
#include <cufft.h>
#include <cuda_runtime.h>
int main()
{
    const int sizes[] = {5000000, 8000, 625};
    for (int si = 0; si < 3; si++) {
        const int n = sizes[si];
        const int batch = si == 0 ? 16 : (si == 1 ? 16 * 625 : 16 * 8000);
        cufftComplex* d;
        cudaMalloc(&d, sizeof(cufftComplex) * (size_t)n * batch);
        cudaMemset(d, 0, sizeof(cufftComplex) * (size_t)n * batch);
        cufftHandle h;
        if (cufftPlan1d(&h, n, CUFFT_C2C, batch) != CUFFT_SUCCESS) { printf("plan failed\n"); return 1; }
        size_t ws = 0; cufftGetSize(h, &ws);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int i = 0; i < 3; i++) cufftExecC2C(h, d, d, CUFFT_FORWARD);
        cudaEventRecord(a);
        const int reps = 10;
        for (int i = 0; i < reps; i++) cufftExecC2C(h, d, d, CUFFT_FORWARD);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double per = ms / reps;
        const double bytes = 16.0 * n * batch;      // one read + one write of the data
        printf("cuFFT C2C n=%d batch=%d: %.3f ms per batch, %.1f us per transform, %.0f GB/s for ONE read+write of the data (workspace %.0f MB)\n",
               n, batch, per, 1e3 * per / batch, bytes / per / 1e6, ws / 1e6);
        cufftDestroy(h); cudaFree(d);
    }
    return 0;
}
