// Memory + launch backend used by pipeline.cuh.
//
// PRODUCT build (default): CUDA only -- cudaMalloc / cudaMemcpyAsync / <<<>>> launches on the
// plan's stream.  There is NO CPU code path in the product library.
//
// TEST build (-DTWSTFT_EMU, only tests/emu/Makefile does that, output tests/emu/_build/): the very
// same kernel phase functions are executed thread by thread on the host so that the CPU test-suite
// can check every kernel's index arithmetic against the oracle without a GPU.  That library is
// test infrastructure; amaranth_twstft_b200 never loads it.
#pragma once
#include "launch.cuh"
#ifdef TWSTFT_EMU
#include "../../tests/emu/emu_backend.h"
#else
#include <stdio.h>
#include <string.h>
#include <vector>

namespace twstft {

struct Backend {
    typedef DeviceLauncher Launcher;
    static constexpr bool is_emulation = false;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    int device = 0;
    int num_sms = 148;
    cudaError_t err = cudaSuccess;
    long launches = 0;

    // `err` keeps the first CUDA error of the plan (kernel faults and the like: the context is gone, the plan is
    // dead).  A failed allocation is NOT such an error: it is remembered in `oom` for the current ABI call only
    // (use() clears it), so an oversized request returns TWSTFT_E_NOMEM and the plan stays usable.
    bool oom = false;
    bool ok() const { return err == cudaSuccess && !oom; }
    void note(cudaError_t e) { if (e != cudaSuccess && err == cudaSuccess) err = e; }
    const char* err_string() const { return (oom && err == cudaSuccess) ? "out of device memory" : cudaGetErrorString(err); }

    int init(int dev)
    {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) { err = (e != cudaSuccess) ? e : cudaErrorNoDevice; return -1; }
        if (dev < 0 || dev >= n) { err = cudaErrorInvalidDevice; return -1; }
        device = dev;
        note(cudaSetDevice(dev));
        note(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        note(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        note(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        return ok() ? 0 : -1;
    }
    void shutdown()
    {
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        for (int s = 0; s < 2; s++) {
            if (ev_copied[s]) cudaEventDestroy(ev_copied[s]);
            if (ev_consumed[s]) cudaEventDestroy(ev_consumed[s]);
            ev_copied[s] = ev_consumed[s] = nullptr;
        }
        stream = copy_stream = nullptr;
    }
    void use() { cudaSetDevice(device); oom = false; }
    void* alloc(size_t bytes)
    {
        void* p = nullptr;
        const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e == cudaErrorMemoryAllocation) {
            oom = true;
            cudaGetLastError();      // clear the (non-sticky) runtime error state
            return nullptr;
        }
        note(e);
        return p;
    }
    void free(void* p) { if (p) cudaFree(p); }
    void zero(void* p, size_t bytes) { note(cudaMemsetAsync(p, 0, bytes, stream)); }
    void h2d(void* d, const void* h, size_t bytes) { note(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream)); }
    void d2h(void* h, const void* d, size_t bytes) { note(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream)); }
    void sync() { note(cudaStreamSynchronize(stream)); }
    // Double-buffered input staging: host->device copies run on copy_stream and overlap the kernels of
    // the previous chunk on `stream`; slot s is reused only after the kernels that read it have finished.
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    void stage_copy(int slot, void* d, const void* h, size_t bytes)
    {
        if (!ev_copied[slot]) {
            note(cudaEventCreateWithFlags(&ev_copied[slot], cudaEventDisableTiming));
            note(cudaEventCreateWithFlags(&ev_consumed[slot], cudaEventDisableTiming));
        }
        if (slot_used[slot]) note(cudaStreamWaitEvent(copy_stream, ev_consumed[slot], 0));
        note(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, copy_stream));
        note(cudaEventRecord(ev_copied[slot], copy_stream));
        note(cudaStreamWaitEvent(stream, ev_copied[slot], 0));
    }
    void stage_release(int slot)
    {
        note(cudaEventRecord(ev_consumed[slot], stream));
        slot_used[slot] = true;
    }
    // optional per-kernel CUDA-event profiling (bench.py roofline leg): one event pair per launch
    struct ProfSlot { const char* name; double ms; long count; };
    bool profiling = false;
    std::vector<ProfSlot> prof;
    struct Pending { int slot; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    cudaEvent_t t0 = nullptr, t1 = nullptr;

    int prof_slot(const char* name)
    {
        for (size_t i = 0; i < prof.size(); i++)
            if (!strcmp(prof[i].name, name)) return (int)i;
        prof.push_back(ProfSlot{name, 0.0, 0});
        return (int)prof.size() - 1;
    }
    void prof_collect()
    {
        sync();
        for (auto& pe : pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) { prof[pe.slot].ms += ms; prof[pe.slot].count++; }
            cudaEventDestroy(pe.a);
            cudaEventDestroy(pe.b);
        }
        pending.clear();
    }
    void timer_start()
    {
        if (!t0) { cudaEventCreate(&t0); cudaEventCreate(&t1); }
        note(cudaEventRecord(t0, stream));
    }
    float timer_stop()
    {
        float ms = -1.f;
        note(cudaEventRecord(t1, stream));
        note(cudaEventSynchronize(t1));
        note(cudaEventElapsedTime(&ms, t0, t1));
        return ms;
    }
    template <class K> void launch(dim3 grid, const typename K::Params& p)
    {
        DeviceLauncher l{stream};
        if (profiling) {
            Pending pe;
            pe.slot = prof_slot(K::name());
            cudaEventCreate(&pe.a);
            cudaEventCreate(&pe.b);
            cudaEventRecord(pe.a, stream);
            l.launch<K>(grid, p);
            cudaEventRecord(pe.b, stream);
            pending.push_back(pe);
        } else {
            l.launch<K>(grid, p);
        }
        note(l.err);
        launches++;
    }
};

}  // namespace twstft
#endif
