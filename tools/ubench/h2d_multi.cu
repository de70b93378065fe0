// Host->device ceiling of one box with G GPUs copying AT THE SAME TIME (one host thread per GPU), for the two kinds of
// pinned memory: what bounds the end-to-end rate of the correlator (40 MB of int16 IQ per record).
//   nvcc -O2 -o h2d_multi h2d_multi.cu -lpthread && ./h2d_multi [G]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>
#include <cuda_runtime.h>

int main(int argc, char** argv)
{
    int G = 0;
    cudaGetDeviceCount(&G);
    if (argc > 1 && atoi(argv[1]) > 0 && atoi(argv[1]) < G) G = atoi(argv[1]);
    const size_t bytes = (size_t)40000000 * 16;      // 16 records per copy burst
    for (int wc = 0; wc < 2; wc++)
        for (int g_active = 1; g_active <= G; g_active *= 2) {
            std::vector<void*> h(g_active), d(g_active);
            for (int g = 0; g < g_active; g++) {
                cudaSetDevice(g);
                cudaHostAlloc(&h[g], bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
                cudaMalloc(&d[g], bytes);
            }
            std::atomic<int> ready{0};
            std::atomic<bool> go{false};
            std::vector<std::thread> th;
            for (int g = 0; g < g_active; g++)
                th.emplace_back([&, g] {
                    cudaSetDevice(g);
                    cudaMemcpy(d[g], h[g], bytes, cudaMemcpyHostToDevice);       // warm-up
                    ready++;
                    while (!go.load()) {}
                    for (int rep = 0; rep < 4; rep++) cudaMemcpy(d[g], h[g], bytes, cudaMemcpyHostToDevice);
                });
            while (ready.load() < g_active) {}
            const auto t0 = std::chrono::steady_clock::now();
            go = true;
            for (auto& t : th) t.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("%-15s %d GPU(s) at once: %.1f GB/s aggregate, %.1f GB/s per GPU -> %.0f records/s\n",
                   wc ? "write-combined" : "pinned", g_active, 4.0 * bytes * g_active / dt / 1e9, 4.0 * bytes / dt / 1e9,
                   4.0 * bytes * g_active / dt / 4e7);
            for (int g = 0; g < g_active; g++) { cudaSetDevice(g); cudaFreeHost(h[g]); cudaFree(d[g]); }
        }
    return 0;
}
