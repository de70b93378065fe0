// Shared-memory FFT building blocks: in-place decimation-in-frequency, radices 2/4/5 and composites.
//
// A forward transform of length L = R0*R1*R2 is run by NTS = L/max(R) threads, each owning
// E = max(R) elements in registers.  With the index written in mixed radix
//     i = i0*(R1*R2) + i1*R2 + i2,        k = k0 + R0*k1 + R0*R1*k2,
// stage 0 transforms over i0 (stride R1*R2) and multiplies output k0 by W_L^{k0*(i1*R2+i2)},
// stage 1 transforms over i1 (stride R2)    and multiplies output k1 by W_{R1*R2}^{k1*i2},
// stage 2 transforms over i2 (stride 1).
// Every butterfly writes its outputs to the very slots it read (digit i_s replaced by k_s), so a
// stage needs only ONE block barrier (before the next stage's reads) and its stores overlap its
// arithmetic.  The first stage reads global memory directly (index jb + m*L/R0: coalesced), the last
// stage is visited as (output index jb + m*L/RLAST, value): natural order, coalesced again.
//
// Stage twiddles are powers w^m of one per-butterfly root w.  Only the seeds w^1..w^(B-1) and
// w^B, w^2B, ... (B = 5 or 4) come from tables; the rest are products of two seeds (error ~2e-7).
// Tables are laid out [seed][k] so that consecutive threads read consecutive words; when the
// root index k exceeds 32 values it is split k = KD*kh + kl and the seed is TA[.][kh]*TB[.][kl].
// The tables ("blob") are computed in fp64 on the host once per plan.
//
// Inverse transforms use the same code through IDFT(x) = swap(DFT(swap(x))), swap = (re<->im).
#pragma once
#include "bfly.cuh"

namespace twstft {

constexpr int split_kd(int ns)
{
    if (ns <= 32) return ns;
    // a half-warp (16 threads, 64-bit words) reads TB[.][jb % KD]: bank-conflict free only if the wrap of
    // jb % KD falls on a multiple of 16 (ncu: KD = 25 cost 37 % extra shared wavefronts on these loads)
    if (ns % 32 == 0) return 32;
    if (ns % 16 == 0) return 16;
    for (int d = 32; d >= 2; d--)
        if (ns % d == 0) return d;
    return 1;
}
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// seeds of a radix-R stage: w^1..w^(B-1), w^B, w^2B, ..., w^(R-B).  FULL: every power comes from the
// table (B = R), used where the table is small (root index <= 32 values): no products at all.
template <int R, bool FULL = false> struct Seeds {
    static constexpr int B = (FULL || R <= 5) ? R : ((R % 5 == 0) ? 5 : 4);
    static constexpr int NLOW = B - 1;                 // w^1..w^(B-1)
    static constexpr int NHIGH = R / B - 1;            // w^B..w^(R-B)
    static constexpr int N = NLOW + NHIGH;
    static constexpr int exponent(int row) { return row < NLOW ? row + 1 : B * (row - NLOW + 1); }
};

template <int R0_, int R1_ = 1, int R2_ = 1> struct FFTSpec {
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
    static_assert(R1 > 1 || R2 == 1, "stages must be packed to the front");
    static constexpr int L = R0 * R1 * R2;
    static constexpr int NSTAGE = 1 + (R1 > 1) + (R2 > 1);
    static constexpr int RMAX = cmax(R0, cmax(R1, R2));
    static constexpr int NTS = L / RMAX;          // threads per sequence
    static constexpr int E = RMAX;                // elements per thread
    static_assert((L / R0) % NTS == 0 && (L / R1) % NTS == 0 && (L / R2) % NTS == 0,
                  "thread count must divide the butterfly count of every stage");
    static constexpr int RLAST = R2 > 1 ? R2 : (R1 > 1 ? R1 : R0);
    static constexpr int NSLAST = L / RLAST;      // output index of the last stage: jb + m*NSLAST
    static constexpr int SP0 = R1 * R2;           // stride of digit i0 (= butterflies in stage 0)
    // stage-0 twiddle tables: root index jb in [0, SP0)
    static constexpr int KD0 = split_kd(SP0), KH0 = SP0 / KD0;
    typedef Seeds<R0, false> SD0;          // stage-0 twiddles w = W_L^jb, jb in [0, SP0)
#ifndef TWSTFT_SD1_FULL
#define TWSTFT_SD1_FULL 1
#endif
    // middle-stage twiddles w = W_{R1 R2}^i2, i2 in [0, R2): the table is small (R1*R2 entries), so every
    // power is tabulated and no seed products are formed (12 complex multiplies less per radix-20 butterfly)
    typedef Seeds<R1, TWSTFT_SD1_FULL != 0> SD1;
    static constexpr int NSEED0 = SD0::N, NSEED1 = SD1::N;
    static constexpr int OFF_B0 = 0;
    static constexpr int OFF_A0 = OFF_B0 + (NSTAGE > 1 ? NSEED0 * KD0 : 0);
    // middle-stage twiddle table (3 stages only): root index i2 in [0, R2), R2 <= 32
    static constexpr int OFF_B1 = OFF_A0 + ((NSTAGE > 1 && KH0 > 1) ? NSEED0 * KH0 : 0);
    static constexpr int BLOB = OFF_B1 + (NSTAGE == 3 ? NSEED1 * R2 : 0);     // float2 entries
    static constexpr int BLOB_PAD = (BLOB + 1) & ~1;
    static_assert(NSTAGE < 3 || R2 <= 32, "middle-stage root table is unsplit");
};

// Host: fill the twiddle blob of a spec (fp64 -> fp32).
template <class S> void fill_blob(float2* out)
{
    const double TWO_PI = 6.283185307179586476925287;
    if (S::NSTAGE > 1) {
        for (int row = 0; row < S::NSEED0; row++) {
            const int m = S::SD0::exponent(row);
            for (int kl = 0; kl < S::KD0; kl++) {
                double a = -TWO_PI * (double)((long long)m * kl) / (double)S::L;
                out[S::OFF_B0 + row * S::KD0 + kl] = make_float2((float)cos(a), (float)sin(a));
            }
            if (S::KH0 > 1)
                for (int kh = 0; kh < S::KH0; kh++) {
                    double a = -TWO_PI * (double)((long long)m * S::KD0 * kh) / (double)S::L;
                    out[S::OFF_A0 + row * S::KH0 + kh] = make_float2((float)cos(a), (float)sin(a));
                }
        }
    }
    if (S::NSTAGE == 3) {
        for (int row = 0; row < S::NSEED1; row++) {
            const int m = S::SD1::exponent(row);
            for (int i2 = 0; i2 < S::R2; i2++) {
                double a = -TWO_PI * (double)(m * i2) / (double)(S::R1 * S::R2);
                out[S::OFF_B1 + row * S::R2 + i2] = make_float2((float)cos(a), (float)sin(a));
            }
        }
    }
}

// Multiply outputs v[1..R-1] of a butterfly by w^m, the seeds read through `seed(row)`.
template <int R, class SD, class SeedFn> HD void apply_powers(float2* v, SeedFn seed)
{
    float2 lo[SD::NLOW > 0 ? SD::NLOW : 1];
#pragma unroll
    for (int b = 1; b < SD::B; b++) {
        lo[b - 1] = seed(b - 1);
        v[b] = cmul(v[b], lo[b - 1]);
    }
#pragma unroll
    for (int a = 1; a <= SD::NHIGH; a++) {
        const float2 hi = seed(SD::NLOW + a - 1);
        v[a * SD::B] = cmul(v[a * SD::B], hi);
#pragma unroll
        for (int b = 1; b < SD::B; b++) v[a * SD::B + b] = cmul(v[a * SD::B + b], cmul(hi, lo[b - 1]));
    }
}

// Exchange-buffer layouts.  at(block, within): `block` = most significant digit (0..R0-1),
// `within` in [0, SP0).  unit() = pointer stride of one `within` step.
//   RowAcc: one sequence, each block padded so that the block stride is odd (conflict-free last stage)
//   ColAcc: TS interleaved sequences, element p of sequence s at data[p*TS + s]
template <int SP0> struct RowAcc {
    static constexpr int BSTRIDE = SP0 + ((SP0 % 2 == 0) ? 5 : 4);
    float2* data;
    HD float2* at(int block, int within) const { return data + block * BSTRIDE + within; }
    static constexpr int unit() { return 1; }
    static constexpr int bstride() { return BSTRIDE; }
};
template <int SP0, int TS> struct ColAcc {
    float2* data;   // already offset by s
    HD float2* at(int block, int within) const { return data + (block * SP0 + within) * TS; }
    static constexpr int unit() { return TS; }
    static constexpr int bstride() { return SP0 * TS; }
};

// ---- stage 0: butterflies over the most significant digit; v[b*R0+m] = x[jb + m*SP0] on entry ----
template <class S, class Acc> HD void dif_stage0(int j, float2* v, const float2* tw, const Acc& acc)
{
    constexpr int R = S::R0, NB = S::SP0 / S::NTS;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int jb = j + S::NTS * b;
        float2* vb = v + b * R;
        Bfly<R>::run(vb);
        if (S::NSTAGE > 1) {
            const float2* tb = tw + S::OFF_B0 + jb % S::KD0;
            const float2* ta = tw + S::OFF_A0 + jb / S::KD0;
            apply_powers<R, typename S::SD0>(vb, [&](int row) {
                float2 w = tb[row * S::KD0];
                if (S::KH0 > 1) w = cmul(w, ta[row * S::KH0]);
                return w;
            });
            float2* p = acc.at(0, jb);
#pragma unroll
            for (int m = 0; m < R; m++) p[m * Acc::bstride()] = vb[m];
        }
    }
}

// Middle-stage butterflies are indexed by (k0, i2) and touch the words k0*bstride + i2 + m*R2.  With
// jb -> (jb / R2, jb % R2) three half-warps out of four straddle two blocks k0 and, the block stride being
// odd for the sake of the last stage, collide in one bank (ncu: 37.5 % excess wavefronts on every load and
// store of this stage).  For the 20 x 20 grid of the length-8000 rows the butterflies are dealt out in
// tiles of 16 instead: 20 tiles (one k0, i2 = 0..15), 4 tiles (one i2 >= 16, k0 = 0..15: odd stride ->
// 16 distinct banks) and one 4 x 4 tile in the corner (the only one left with a 2-way conflict).
template <class S> HD void mid_map(int jb, int& k0, int& i2)
{
    if (S::R0 == 20 && S::R2 == 20) {
        if (jb < 320) { k0 = jb >> 4; i2 = jb & 15; }
        else if (jb < 384) { const int u = jb - 320; i2 = 16 + (u >> 4); k0 = u & 15; }
        else { const int u = jb - 384; k0 = 16 + (u >> 2); i2 = 16 + (u & 3); }
    } else {
        k0 = jb / S::R2;
        i2 = jb % S::R2;
    }
}

// ---- middle stage (3-stage transforms): digit i1, in place ------------------------------------
template <class S, class Acc> HD void dif_stage_mid(int j, float2* v, const float2* tw, const Acc& acc)
{
    constexpr int R = S::R1, NB = (S::L / R) / S::NTS;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int jb = j + S::NTS * b;
        int k0, i2;
        mid_map<S>(jb, k0, i2);
        float2* vb = v + b * R;
        float2* p = acc.at(k0, i2);
#pragma unroll
        for (int m = 0; m < R; m++) vb[m] = p[m * S::R2 * Acc::unit()];
        Bfly<R>::run(vb);
        const float2* tb = tw + S::OFF_B1 + i2;
        apply_powers<R, typename S::SD1>(vb, [&](int row) { return tb[row * S::R2]; });
#pragma unroll
        for (int m = 0; m < R; m++) p[m * S::R2 * Acc::unit()] = vb[m];
    }
}

// ---- middle stage, OUTPUT-PRUNED (radix 20 = 4 x 5 prime-factor form): only the outputs k1 whose bit is set in
// `mask` are formed, twiddled and stored.  The estimator row passes need the few bins of a row that fall inside the
// carrier window, i.e. a handful of (k0 + 20 k1) groups: the five radix-4 butterflies of the first PFA half are
// shared, each wanted output then costs one five-term dot product with constant roots instead of its share of four
// radix-5 butterflies, 19 twiddles and 20 stores.  The mask is uniform over the CTA (no divergence).
// CMASK != 0: the mask is known at compile time (the default carrier window of the 1-s records), no tests remain.
template <class S, class Acc, unsigned CMASK = 0u>
HD void dif_stage_mid_pruned(int j, float2* v, const float2* tw, const Acc& acc, unsigned mask_rt)
{
    const unsigned mask = CMASK ? CMASK : mask_rt;
    static_assert(S::R1 == 20 && S::SD1::B == S::R1, "pruned middle stage: radix 20, every twiddle power tabulated");
    static_assert((S::L / S::R1) / S::NTS == 1, "one middle-stage butterfly per thread");
    int k0, i2;
    mid_map<S>(j, k0, i2);
    float2* p = acc.at(k0, i2);
    float2 t[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; n2++) {
        float2 c[4];
#pragma unroll
        for (int n1 = 0; n1 < 4; n1++) c[n1] = p[((5 * n1 + 4 * n2) % 20) * S::R2 * Acc::unit()];
        Bfly<4>::run(c);
#pragma unroll
        for (int ka = 0; ka < 4; ka++) t[ka][n2] = c[ka];
    }
    const float2* tb = tw + S::OFF_B1 + i2;
#pragma unroll
    for (int k1 = 0; k1 < 20; k1++) {
        if (mask & (1u << k1)) {
            const int ka = k1 % 4, kb = k1 % 5;       // output k of the 4 x 5 map: (k mod 4, k mod 5)
            float2 y = t[ka][0];
#pragma unroll
            for (int n2 = 1; n2 < 5; n2++) y = cadd(y, mul_root<20>(t[ka][n2], 4 * n2 * kb));      // W_5^(n2 kb) = W_20^(4 n2 kb)
            if (k1 > 0) y = cmul(y, tb[(k1 - 1) * S::R2]);
            p[k1 * S::R2 * Acc::unit()] = y;
        }
    }
    (void)v;
}

// ---- last stage (NSTAGE >= 2): digit i_last (stride 1); results stay in registers ---------------
template <class S, class Acc> HD void dif_stage_last(int j, float2* v, const Acc& acc)
{
    constexpr int R = S::RLAST, NB = (S::L / R) / S::NTS;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int jb = j + S::NTS * b;
        float2* vb = v + b * R;
        const float2* p = acc.at(jb % S::R0, (jb / S::R0) * R);
#pragma unroll
        for (int m = 0; m < R; m++) vb[m] = p[m * Acc::unit()];
        Bfly<R>::run(vb);
    }
}

// Visit the outputs of the last stage: (output index jb + m*NSLAST, value, b, m).
template <class S, class Dst> HD void last_stage_visit(int j, const float2* v, Dst dst)
{
    constexpr int R = S::RLAST;
    constexpr int NB = (S::L / R) / S::NTS;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int jb = j + S::NTS * b;
#pragma unroll
        for (int m = 0; m < R; m++) dst(jb + m * S::NSLAST, v[b * R + m], b, m);
    }
}

}  // namespace twstft
