// Host->device copy yardstick for the end-to-end path (40 MB records from pinned memory):
//   nvcc -O2 -o h2d h2d.cu && ./h2d
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    const size_t rec = 40000000, nrec = 8;
    void *h, *hw, *d;
    cudaHostAlloc(&h, rec * nrec, cudaHostAllocDefault);
    cudaHostAlloc(&hw, rec * nrec, cudaHostAllocWriteCombined);
    cudaMalloc(&d, rec * nrec);
    cudaStream_t s[2];
    cudaStreamCreate(&s[0]); cudaStreamCreate(&s[1]);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 4; mode++) {
        const char* name[] = {"pinned, one 320 MB copy", "pinned, 8 x 40 MB on one stream", "pinned, 8 x 40 MB on two streams", "write-combined, one copy"};
        float best = 1e9f;
        for (int rep = 0; rep < 6; rep++) {
            cudaDeviceSynchronize();
            cudaEventRecord(a, 0);
            if (mode == 0) cudaMemcpyAsync(d, h, rec * nrec, cudaMemcpyHostToDevice, 0);
            else if (mode == 3) cudaMemcpyAsync(d, hw, rec * nrec, cudaMemcpyHostToDevice, 0);
            else {
                for (size_t i = 0; i < nrec; i++)
                    cudaMemcpyAsync((char*)d + i * rec, (char*)h + i * rec, rec, cudaMemcpyHostToDevice, mode == 1 ? s[0] : s[i & 1]);
                cudaStreamSynchronize(s[0]); cudaStreamSynchronize(s[1]);
            }
            cudaEventRecord(b, 0);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        printf("%-36s %.3f ms  %.1f GB/s  -> %.0f records/s\n", name[mode], best, rec * nrec / best / 1e6, nrec / (best * 1e-3));
    }
    return 0;
}
