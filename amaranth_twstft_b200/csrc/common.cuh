// Common definitions for the B200 PRN correlator kernels.
//
// Every kernel is written as a struct of per-thread "phases" (__host__ __device__), separated by
// block-wide barriers or block reductions.  On the GPU the phases run inside one __global__
// kernel (launch.cuh).  The same phase functions are also compiled for the host by the TEST-ONLY
// emulator (tests/emu), which runs them thread by thread so the index arithmetic of every kernel
// can be checked against the oracle in the CPU test-suite.  The emulator is never part of the
// product library.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <cuda_runtime.h>

#define HD __host__ __device__ __forceinline__

namespace twstft {

typedef unsigned long long u64;
typedef long long i64;

struct Blk { int x, y, z, gx; };   // block index and gridDim.x

enum SyncKind { SYNC_BARRIER = 0, SYNC_MAXKEY = 1, SYNC_SUMD = 2, SYNC_MAXKEY_SUMD = 3, SYNC_SUMF16 = 4,
                SYNC_MAXKEY_ATOMIC = 5 };   // regs.key maximised over the CTA straight into *K::key_slot(p, blk, it) (zeroed
                                            // by the host before the launch): warp shuffles + one atomic per warp, no barrier

// ------------------------------------------------------------------ complex fp32 helpers
// On sm_100a a complex value lives in an aligned 64-bit register pair and the arithmetic uses the
// packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2): same FP32 pipe throughput as the scalar forms
// (measured, tools/ubench/fp32x2.cu) but half the issue slots, which is what bounds the FFT passes.
// pk()/upk() are register renames for ptxas.  The host build (test emulation) uses plain floats.
HD float2 cf(float x, float y) { return make_float2(x, y); }
#ifdef __CUDA_ARCH__
__device__ __forceinline__ u64 pk(float2 a) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ u64 pk(float x, float y) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#endif
HD float2 cadd(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    return upk(add2(pk(a), pk(b)));
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
HD float2 csub(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    return upk(sub2(pk(a), pk(b)));
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// s * a (real scalar)
HD float2 cscale(float s, float2 a)
{
#ifdef __CUDA_ARCH__
    return upk(mul2(pk(s, s), pk(a)));
#else
    return make_float2(s * a.x, s * a.y);
#endif
}
// s * a + b (real scalar s)
HD float2 caxpy(float s, float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    return upk(fma2(pk(s, s), pk(a), pk(b)));
#else
    return make_float2(s * a.x + b.x, s * a.y + b.y);
#endif
}
// a*b in TWO packed instructions: t = b*(ax,ax); r = swap(b)*(-ay,ay) + t.  ptxas folds the half swap and
// the single-half negation into operand modifiers of FFMA2 (SASS `-R.F32x2.LO_HI.NP, R.F32`), so no scalar
// FFMA and no register moves remain (checked with cuobjdump; the scalar form cost three issue slots).
HD float2 cmul(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    const u64 t = mul2(pk(a.x, a.x), pk(b));
    return upk(fma2(pk(b.y, b.x), pk(-a.y, a.y), t));
#else
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
// a*conj(b)
HD float2 cmulc(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    const u64 t = mul2(pk(b.x, b.x), pk(a));
    return upk(fma2(pk(a.y, a.x), pk(b.y, -b.y), t));
#else
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
#endif
}
// a*(c - i s), c and s compile-time constants of an in-register butterfly
HD float2 cmul_const(float2 a, float c, float s)
{
#ifdef __CUDA_ARCH__
    return upk(fma2(pk(a.y, a.x), pk(s, -s), mul2(pk(c, c), pk(a))));
#else
    return make_float2(c * a.x + a.y * s, c * a.y - a.x * s);
#endif
}
// -i*a and +i*a: register renames plus one sign, folded by ptxas into the consuming packed instruction
HD float2 rot_mi(float2 a) { return make_float2(a.y, -a.x); }
HD float2 rot_pi(float2 a) { return make_float2(-a.y, a.x); }
HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
HD float2 cswap(float2 a) { return make_float2(a.y, a.x); }
HD float cnorm(float2 a) { return a.x * a.x + a.y * a.y; }

// exp(2*pi*i*cycles) in fp32, phase reduced in fp64 so that the fp32 argument is < 1/8 turn.
HD float2 cis_cycles(double c)
{
    c -= rint(c);
    double c4 = c * 4.0;
    double q = rint(c4);
    float r = (float)((c4 - q) * 0.25);  // |r| <= 1/8 cycle
    int qi = ((int)q) & 3;
    float s, co;
#ifdef __CUDA_ARCH__
    sincospif(2.0f * r, &s, &co);
#else
    s = (float)sin(6.283185307179586476925287 * (double)r);
    co = (float)cos(6.283185307179586476925287 * (double)r);
#endif
    float2 w;
    if (qi == 0) w = make_float2(co, s);
    else if (qi == 1) w = make_float2(-s, co);
    else if (qi == 2) w = make_float2(-co, -s);
    else w = make_float2(s, -co);
    return w;
}

// exp(2*pi*i*cycles) in fp64.
HD double2 cis_cycles_d(double c)
{
    c -= rint(c);
    double s, co;
#ifdef __CUDA_ARCH__
    sincospi(2.0 * c, &s, &co);
#else
    s = sin(6.283185307179586476925287 * c);
    co = cos(6.283185307179586476925287 * c);
#endif
    return make_double2(co, s);
}

HD double2 dmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// Pull one 128-byte line towards L2 ahead of the loads of the next work item (no-op on the host).
HD void prefetch_l2(const void* g)
{
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
#endif
}

// Order-preserving key for "first maximum": larger value wins, then SMALLER index.
HD u64 make_key(float v, unsigned idx)
{
#ifdef __CUDA_ARCH__
    unsigned b = __float_as_uint(v);
#else
    union { float f; unsigned u; } cv; cv.f = v; unsigned b = cv.u;
#endif
    return ((u64)b << 32) | (u64)(0xFFFFFFFFu - idx);   // v >= 0 so the bit pattern is monotone
}
HD unsigned key_index(u64 k) { return 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull); }
HD float key_value(u64 k)
{
    unsigned b = (unsigned)(k >> 32);
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; unsigned u; } cv; cv.u = b; return cv.f;
#endif
}

}  // namespace twstft
