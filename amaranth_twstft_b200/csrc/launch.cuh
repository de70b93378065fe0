// Generic __global__ entry for phase-structured kernels and the block-wide reductions that separate
// their phases.  A kernel struct K provides:
//   K::Params, K::Regs, K::NT (threads per CTA), K::MINB (min CTAs/SM), K::NPH (phase count),
//   K::smem_bytes(), template<int PH> K::phase(p, blk, tid, regs, smem), K::sync_after(ph)
// SYNC_MAXKEY  : regs.key   <- max over the CTA (valid in every thread afterwards)
// SYNC_SUMD    : regs.dsum[i] <- sum over the CTA for i < K::NSUM (valid in thread 0)
// SYNC_SUMF16  : regs.dsum[0] of thread i < 16 <- sum over the CTA of regs.fsum[i] (fp32 inside a warp, fp64 across
//                the warps): 16 shuffles per thread instead of the 160 of sixteen fp64 sums
#pragma once
#include "common.cuh"

namespace twstft {

constexpr int RED_SCRATCH_BYTES = 32 * 8 * 20;   // 32 warps x up to 20 doubles

#ifdef __CUDACC__
// Warp maximum of a 64-bit key with two REDUX instructions (value word, then the index word among the lanes that hold
// the maximum value) instead of five shuffle rounds on register pairs; a key's low word is never 0 for a valid entry
// (0xFFFFFFFF - index), an empty key is 0 in both words.
__device__ __forceinline__ u64 warp_max_key(u64 k)
{
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((u64)mhi << 32) | (u64)mlo;
}

template <class K> __device__ __forceinline__ void block_max_key(typename K::Regs& r, unsigned char* scratch)
{
    u64 k = r.key;                   // (shuffle rounds: the REDUX form of warp_max_key measured 5 % slower at the end of
#pragma unroll                       //  the short-lived inverse-column CTAs, 1.6 % faster in the resident row pass)
    for (int o = 16; o > 0; o >>= 1) {
        u64 other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other > k ? other : k;
    }
    u64* s = reinterpret_cast<u64*>(scratch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = (K::NT + 31) / 32;
    __syncthreads();                 // scratch may still be read from a previous reduction
    if (lane == 0) s[warp] = k;
    __syncthreads();
    u64 best = s[0];
    if constexpr (NW <= 8) {         // few warps: all partials requested at once (the rolled loop is a chain of NW
#pragma unroll                       // shared-memory round trips at the very end of a short-lived CTA)
        for (int w = 1; w < NW; w++) best = s[w] > best ? s[w] : best;
    } else {
#pragma unroll 1
        for (int w = 1; w < NW; w++) best = s[w] > best ? s[w] : best;
    }
    r.key = best;
}

// SYNC_MAXKEY_ATOMIC: no shared scratch, no barrier -- the result is only read by a later kernel
template <class K> __device__ __forceinline__ void block_max_key_atomic(typename K::Regs& r, u64* slot)
{
    const u64 k = warp_max_key(r.key);
    if ((threadIdx.x & 31) == 0 && k != 0) atomicMax(reinterpret_cast<unsigned long long*>(slot), (unsigned long long)k);
}

template <class K> __device__ __forceinline__ void block_sum_d(typename K::Regs& r, unsigned char* scratch)
{
    static_assert(K::NSUM * 32 * 8 <= RED_SCRATCH_BYTES, "reduction scratch too small for this kernel's sums");
    double* s = reinterpret_cast<double*>(scratch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = (K::NT + 31) / 32;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < K::NSUM; i++) {
        double x = r.dsum[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s[i * 32 + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < K::NSUM; i++) {
            double t = 0.0;
            for (int w = 0; w < NW; w++) t += s[i * 32 + w];
            r.dsum[i] = t;
        }
    }
}

// Sixteen fp32 sums per thread: transposing butterfly (each exchange halves the number of values a lane carries, so
// the whole warp reduction is 8+4+2+1+1 shuffles), lanes 2k and 2k+1 end with the warp total of value k; the warp
// totals are added in fp64 by threads 0..15.
template <class K> __device__ __forceinline__ void block_sum_f16(typename K::Regs& r, unsigned char* scratch)
{
    static_assert(16 * 32 * 8 <= RED_SCRATCH_BYTES, "reduction scratch too small");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = (K::NT + 31) / 32;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = r.fsum[i];
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {                // offsets 16, 8, 4, 2 keep h values each
        const bool up = (lane & (2 * h)) != 0;
#pragma unroll
        for (int i = 0; i < h; i++) {
            const float send = up ? v[i] : v[i + h];
            const float keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * h);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    double* s = reinterpret_cast<double*>(scratch);
    __syncthreads();
    if ((lane & 1) == 0) s[warp * 16 + (lane >> 1)] = (double)v[0];
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
        for (int w = 0; w < NW; w++) t += s[w * 16 + threadIdx.x];
        r.dsum[0] = t;
    }
}

// Persistent kernels (K::PERSISTENT) loop over work items: phase<PH>(..., it, nit) is run for
// it = 0..nit-1 with nit = K::niter(p, blockIdx.x, gridDim.x).
template <class K, int PH>
__device__ __forceinline__ void run_phases(const typename K::Params& p, const Blk& b, int tid,
                                           typename K::Regs& r, unsigned char* smem, int it, int nit)
{
    if constexpr (K::PERSISTENT) K::template phase<PH>(p, b, tid, r, smem, it, nit);
    else K::template phase<PH>(p, b, tid, r, smem);
    if constexpr (PH + 1 < K::NPH) {
        constexpr int kind = K::sync_after(PH);
        unsigned char* scratch = smem + K::smem_bytes() - RED_SCRATCH_BYTES;
        if constexpr (kind == SYNC_BARRIER) __syncthreads();
        if constexpr (kind == SYNC_MAXKEY || kind == SYNC_MAXKEY_SUMD) block_max_key<K>(r, scratch);
        if constexpr (kind == SYNC_SUMD || kind == SYNC_MAXKEY_SUMD) block_sum_d<K>(r, scratch);
        if constexpr (kind == SYNC_SUMF16) block_sum_f16<K>(r, scratch);
        if constexpr (kind == SYNC_MAXKEY_ATOMIC) block_max_key_atomic<K>(r, K::key_slot(p, b, it));
        run_phases<K, PH + 1>(p, b, tid, r, smem, it, nit);
    }
}

// Kernels that park per-thread data in tensor memory declare K::TMEM_COLS (a power of two >= 32) and a
// `unsigned tmem` member in K::Regs; warp 0 allocates the columns once per CTA and frees them at the end.
template <class K, class = void> struct TmemCols { static constexpr int value = 0; };
template <class K> struct TmemCols<K, decltype((void)K::TMEM_COLS)> { static constexpr int value = K::TMEM_COLS; };

// Kernels with per-CTA constant tables provide  static void prologue(p, tid, smem): it runs once per CTA,
// followed by a block barrier, before the first work item.
template <class K, class = void> struct HasPrologue { static constexpr bool value = false; };
template <class K> struct HasPrologue<K, decltype((void)&K::prologue)> { static constexpr bool value = true; };

// Resident kernels with DYNAMIC work distribution declare K::CHUNK (items per grab), K::total_items(p) and
// K::counter(p) (two device ints, zero before the first launch; the kernel leaves them at zero): a CTA takes the
// next CHUNK items with one atomicAdd and `it` is then the GLOBAL item index.  SMs do not run at the same speed (L2
// distance, co-resident CTAs), so a static split of the items makes every launch as slow as its slowest SM.
template <class K, class = void> struct IsDynamic { static constexpr bool value = false; };
template <class K> struct IsDynamic<K, decltype((void)K::CHUNK)> { static constexpr bool value = true; };

template <class K> __global__ void __launch_bounds__(K::NT, K::MINB) kernel_entry(const typename K::Params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    typename K::Regs r;
    Blk b{(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)gridDim.x};
    if constexpr (HasPrologue<K>::value) {
        K::prologue(p, (int)threadIdx.x, smem);
        __syncthreads();
    }
    constexpr int TC = TmemCols<K>::value;
    if constexpr (TC > 0) {
        __shared__ unsigned tmem_slot;
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             (unsigned)__cvta_generic_to_shared(&tmem_slot)), "r"(TC) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        r.tmem = tmem_slot;
    }
    if constexpr (IsDynamic<K>::value) {
        __shared__ int s_base;
        const int total = K::total_items(p);
        for (;;) {
            __syncthreads();                 // the previous chunk is done: s_base and the shared buffers are free
            if (threadIdx.x == 0) s_base = atomicAdd(K::counter(p), (int)K::CHUNK);
            __syncthreads();
            const int base = s_base;
            if (base >= total) break;
            const int end = base + K::CHUNK < total ? base + K::CHUNK : total;
            for (int it = base; it < end; it++) {
                if (it > base) __syncthreads();
                run_phases<K, 0>(p, b, (int)threadIdx.x, r, smem, it, end);
            }
        }
        // counter[1] counts the CTAs that are done; the last one leaves both counters at zero for the next launch
        // on this stream (no memset between launches)
        if (threadIdx.x == 0) {
            int* c = K::counter(p);
            __threadfence();
            if (atomicAdd(c + 1, 1) == (int)gridDim.x - 1) { c[0] = 0; c[1] = 0; }
        }
    } else if constexpr (K::PERSISTENT) {
        const int nit = K::niter(p, (int)blockIdx.x, (int)gridDim.x);
        for (int it = 0; it < nit; it++) {
            if (it > 0) __syncthreads();     // shared buffers of the previous item are free again
            run_phases<K, 0>(p, b, (int)threadIdx.x, r, smem, it, nit);
        }
    } else {
        run_phases<K, 0>(p, b, (int)threadIdx.x, r, smem, 0, 1);
    }
    if constexpr (TC > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(r.tmem), "r"(TC) : "memory");
    }
}

// Device launcher: used by the product library.
struct DeviceLauncher {
    cudaStream_t stream;
    mutable cudaError_t err = cudaSuccess;
    mutable long launches = 0;
    template <class K> void launch(dim3 grid, const typename K::Params& p) const
    {
        static bool configured[64] = {};    // per kernel type and device
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev & 63]) {
            cudaFuncSetAttribute(kernel_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::smem_bytes());
            configured[dev & 63] = true;
        }
        kernel_entry<K><<<grid, K::NT, K::smem_bytes(), stream>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        launches++;
    }
};
#endif  // __CUDACC__

}  // namespace twstft
