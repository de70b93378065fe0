// TEST INFRASTRUCTURE ONLY (never compiled into the product library).
// Host emulation backend: runs the __host__ __device__ kernel phases of amaranth_twstft_b200/csrc
// thread by thread, with barriers/reductions between phases performed by this driver, so that the
// CPU test-suite can validate the CUDA kernels' index arithmetic without a GPU.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace twstft {

template <class K, int PH> struct EmuPhases {
    static void run(const typename K::Params& p, const Blk& b, std::vector<typename K::Regs>& regs, unsigned char* smem,
                    int it, int nit)
    {
        for (int tid = 0; tid < K::NT; tid++) {
            if constexpr (K::PERSISTENT) K::template phase<PH>(p, b, tid, regs[tid], smem, it, nit);
            else K::template phase<PH>(p, b, tid, regs[tid], smem);
        }
        if constexpr (PH + 1 < K::NPH) {
            constexpr int kind = K::sync_after(PH);
            if constexpr (kind == SYNC_MAXKEY || kind == SYNC_MAXKEY_SUMD) {
                u64 best = 0;
                for (int tid = 0; tid < K::NT; tid++) best = regs[tid].key > best ? regs[tid].key : best;
                for (int tid = 0; tid < K::NT; tid++) regs[tid].key = best;
            }
            if constexpr (kind == SYNC_SUMD || kind == SYNC_MAXKEY_SUMD) {
                for (int i = 0; i < K::NSUM; i++) {
                    double t = 0.0;
                    for (int tid = 0; tid < K::NT; tid++) t += regs[tid].dsum[i];
                    regs[0].dsum[i] = t;
                }
            }
            if constexpr (kind == SYNC_MAXKEY_ATOMIC) {
                u64 best = 0;
                for (int tid = 0; tid < K::NT; tid++) best = regs[tid].key > best ? regs[tid].key : best;
                u64* slot = K::key_slot(p, b, it);
                if (best > *slot) *slot = best;
            }
            if constexpr (kind == SYNC_SUMF16) {
                double t[16];
                for (int i = 0; i < 16; i++) {
                    t[i] = 0.0;
                    for (int tid = 0; tid < K::NT; tid++) t[i] += (double)regs[tid].fsum[i];
                }
                for (int i = 0; i < 16; i++) regs[i].dsum[0] = t[i];
            }
            EmuPhases<K, PH + 1>::run(p, b, regs, smem, it, nit);
        }
    }
};

struct Backend {
    static constexpr bool is_emulation = true;
    int device = 0;
    int num_sms = 3;     // small on purpose: persistent kernels then loop several times per CTA
    int err = 0;
    long launches = 0;
    bool ok() const { return err == 0; }
    const char* err_string() const { return "emulation backend error"; }
    int init(int) { return 0; }
    void shutdown() {}
    bool oom = false;
    void use() { oom = false; }
    void* alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
    void free(void* p) { ::free(p); }
    void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
    void h2d(void* d, const void* h, size_t bytes) { memcpy(d, h, bytes); }
    void d2h(void* h, const void* d, size_t bytes) { memcpy(h, d, bytes); }
    void sync() {}
    void stage_copy(int, void* d, const void* h, size_t bytes) { memcpy(d, h, bytes); }
    void stage_release(int) {}
    struct ProfSlot { const char* name; double ms; long count; };
    bool profiling = false;
    std::vector<ProfSlot> prof;
    void prof_collect() {}
    void timer_start() {}
    float timer_stop() { return 0.f; }
    template <class K> void launch(dim3 grid, const typename K::Params& p)
    {
        std::vector<unsigned char> smem(K::smem_bytes() + 64);
        std::vector<typename K::Regs> regs(K::NT);
        for (unsigned z = 0; z < grid.z; z++)
            for (unsigned y = 0; y < grid.y; y++)
                for (unsigned x = 0; x < grid.x; x++) {
                    Blk b{(int)x, (int)y, (int)z, (int)grid.x};
                    if constexpr (HasPrologue<K>::value)
                        for (int tid = 0; tid < K::NT; tid++) K::prologue(p, tid, smem.data());
                    if constexpr (IsDynamic<K>::value) {
                        // dynamic work distribution: on the host the chunks are dealt out to the CTAs in turn
                        const int total = K::total_items(p);
                        int k = 0;
                        for (int base = 0; base < total; base += K::CHUNK, k++) {
                            if (k % (int)grid.x != (int)x) continue;
                            const int end = base + K::CHUNK < total ? base + K::CHUNK : total;
                            for (int it = base; it < end; it++) EmuPhases<K, 0>::run(p, b, regs, smem.data(), it, end);
                        }
                    } else if constexpr (K::PERSISTENT) {
                        const int nit = K::niter(p, (int)x, (int)grid.x);
                        for (int it = 0; it < nit; it++) EmuPhases<K, 0>::run(p, b, regs, smem.data(), it, nit);
                    } else {
                        EmuPhases<K, 0>::run(p, b, regs, smem.data(), 0, 1);
                    }
                }
        launches++;
    }
};

}  // namespace twstft
