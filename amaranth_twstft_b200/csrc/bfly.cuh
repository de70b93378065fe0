// In-register forward DFT butterflies (W = exp(-2*pi*i/R)) for radices 2,4,5 and the composites
// 8,10,16,20,25 built from them (Cooley-Tukey inside the register file, constant twiddles).
// Inverse transforms reuse these through the swap(re,im) identity, see fft_core.cuh.
#pragma once
#include "common.cuh"
#include "unit_roots.inc"

namespace twstft {

template <int R> struct Bfly;

template <> struct Bfly<1> { static HD void run(float2*) {} };

template <> struct Bfly<2> {
    static HD void run(float2* v)
    {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <> struct Bfly<4> {
    static HD void run(float2* v)
    {
        const float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        const float2 t2 = cadd(v[1], v[3]);
        const float2 t3r = rot_mi(csub(v[1], v[3]));                        // -i*(v1 - v3)
        v[0] = cadd(t0, t2);
        v[2] = csub(t0, t2);
        v[1] = cadd(t1, t3r);   // t1 - i*t3
        v[3] = csub(t1, t3r);   // t1 + i*t3
    }
};

template <> struct Bfly<5> {
    static HD void run(float2* v)
    {
        const float c1 = 0.30901699437494745f, c2 = -0.8090169943749473f;
        const float s1 = 0.9510565162951535f, s2 = 0.5877852522924732f;
        const float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
        const float2 t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        const float2 a1 = caxpy(c2, t2, caxpy(c1, t1, v[0]));
        const float2 a2 = caxpy(c1, t2, caxpy(c2, t1, v[0]));
        const float2 b1 = rot_mi(caxpy(s2, t4, cscale(s1, t3)));    // -i*(s1*t3 + s2*t4)
        const float2 b2 = rot_mi(caxpy(-s1, t4, cscale(s2, t3)));   // -i*(s2*t3 - s1*t4)
        v[0] = cadd(v[0], cadd(t1, t2));
        v[1] = cadd(a1, b1);
        v[4] = csub(a1, b1);
        v[2] = cadd(a2, b2);
        v[3] = csub(a2, b2);
    }
};

// multiply by W_R^t with the trivial rotations resolved at compile time (t is a constant after unrolling)
template <int R> HD float2 mul_root(float2 a, int t)
{
    t = t % R;
    if (t == 0) return a;
    if ((R % 4 == 0) && t == R / 4) return make_float2(a.y, -a.x);       // * (-i)
    if ((R % 2 == 0) && t == R / 2) return make_float2(-a.x, -a.y);
    if ((R % 4 == 0) && t == 3 * R / 4) return make_float2(-a.y, a.x);   // * (+i)
    float c, s;
    UnitRoots<R>::get(t, c, s);
    return cmul_const(a, c, s);                                          // a * (c - i s)
}

// R = RA*RB: n = na*RB + nb, k = ka + RA*kb
template <int RA, int RB> struct BflyCT {
    static HD void run(float2* v)
    {
        constexpr int R = RA * RB;
#pragma unroll
        for (int nb = 0; nb < RB; nb++) {
            float2 t[RA];
#pragma unroll
            for (int na = 0; na < RA; na++) t[na] = v[na * RB + nb];
            Bfly<RA>::run(t);
#pragma unroll
            for (int ka = 0; ka < RA; ka++) v[ka * RB + nb] = mul_root<R>(t[ka], nb * ka);
        }
#pragma unroll
        for (int ka = 0; ka < RA; ka++) Bfly<RB>::run(v + ka * RB);
        float2 o[R];
#pragma unroll
        for (int ka = 0; ka < RA; ka++)
#pragma unroll
            for (int kb = 0; kb < RB; kb++) o[ka + RA * kb] = v[RB * ka + kb];
#pragma unroll
        for (int i = 0; i < R; i++) v[i] = o[i];
    }
};

// R = RA*RB with gcd(RA,RB)=1: Good-Thomas prime-factor algorithm, no internal twiddles.
//   input  n = (RB*n1 + RA*n2) mod R,  output k = (RB*IB*k1 + RA*IA*k2) mod R
//   IB = RB^-1 mod RA, IA = RA^-1 mod RB  (index maps are compile-time register renames)
template <int RA, int RB, int IB, int IA> struct BflyPFA {
    static HD void run(float2* v)
    {
        constexpr int R = RA * RB;
        static_assert((RB * IB) % RA == 1 % RA && (RA * IA) % RB == 1 % RB, "CRT inverses");
        float2 t[RA][RB];
#pragma unroll
        for (int n1 = 0; n1 < RA; n1++)
#pragma unroll
            for (int n2 = 0; n2 < RB; n2++) t[n1][n2] = v[(RB * n1 + RA * n2) % R];
#pragma unroll
        for (int n2 = 0; n2 < RB; n2++) {
            float2 c[RA];
#pragma unroll
            for (int n1 = 0; n1 < RA; n1++) c[n1] = t[n1][n2];
            Bfly<RA>::run(c);
#pragma unroll
            for (int k1 = 0; k1 < RA; k1++) t[k1][n2] = c[k1];
        }
#pragma unroll
        for (int k1 = 0; k1 < RA; k1++) {
            Bfly<RB>::run(t[k1]);
#pragma unroll
            for (int k2 = 0; k2 < RB; k2++) v[(RB * IB * k1 + RA * IA * k2) % R] = t[k1][k2];
        }
    }
};

template <> struct Bfly<8> { static HD void run(float2* v) { BflyCT<2, 4>::run(v); } };
template <> struct Bfly<10> { static HD void run(float2* v) { BflyPFA<2, 5, 1, 3>::run(v); } };
template <> struct Bfly<16> { static HD void run(float2* v) { BflyCT<4, 4>::run(v); } };
template <> struct Bfly<20> { static HD void run(float2* v) { BflyPFA<4, 5, 1, 4>::run(v); } };
template <> struct Bfly<25> { static HD void run(float2* v) { BflyCT<5, 5>::run(v); } };

}  // namespace twstft
