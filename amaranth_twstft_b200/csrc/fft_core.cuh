// Shared-memory Stockham FFT building blocks (forward transform, radices 2/4/5 and composites).
//
// A transform of length L = R0*R1*R2 is run by NTS = L/max(R) threads, each owning E = max(R)
// elements in registers.  Stage s (radix R, Ns = product of the earlier radices) does, for
// butterfly index jb in [0, L/R):
//     v[m]  = X[jb + m*L/R] * W_{Ns*R}^{(jb mod Ns)*m}        m = 0..R-1
//     v     = DFT_R(v)
//     Y[(jb/Ns)*Ns*R + (jb mod Ns) + m*Ns] = v[m]
// (autosort: natural order in, natural order out, no digit reversal).  The first stage reads global
// memory directly (coalesced: consecutive jb), the last stage writes global memory directly
// (index jb + m*L/R, coalesced again); only the exchanges in between go through shared memory.
//
// Stage twiddles come from small 2-D tables laid out [m][k] so that consecutive threads read
// consecutive words (no bank conflicts); for Ns > 32, k = KD*kh + kl and the twiddle is the product
// TA[m][kh] * TB[m][kl] (two correctly-rounded factors, error ~1e-7).  The tables ("blob") are
// computed in fp64 on the host once per plan and copied into shared memory by every CTA.
//
// Inverse transforms use the same code through IDFT(x) = swap(DFT(swap(x))), swap = (re<->im).
#pragma once
#include "bfly.cuh"

namespace twstft {

constexpr int split_kd(int ns)
{
    if (ns <= 32) return ns;
    for (int d = 32; d >= 2; d--)
        if (ns % d == 0) return d;
    return 1;
}
constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int R0_, int R1_ = 1, int R2_ = 1> struct FFTSpec {
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
    static_assert(R1 > 1 || R2 == 1, "stages must be packed to the front");
    static constexpr int L = R0 * R1 * R2;
    static constexpr int NSTAGE = 1 + (R1 > 1) + (R2 > 1);
    static constexpr int RMAX = cmax(R0, cmax(R1, R2));
    static constexpr int NTS = L / RMAX;          // threads per sequence
    static constexpr int E = RMAX;                // elements per thread
    static_assert((L / R0) % NTS == 0 && (L / R1) % NTS == 0 && (L / R2) % NTS == 0 || R1 == 1,
                  "thread count must divide the butterfly count of every stage");
    static constexpr int RLAST = R2 > 1 ? R2 : (R1 > 1 ? R1 : R0);
    static constexpr int NSLAST = L / RLAST;      // output index of the last stage: jb + m*NSLAST
    // stage 1 tables
    static constexpr int NS1 = R0, KD1 = split_kd(NS1), KH1 = NS1 / KD1;
    static constexpr int NS2 = R0 * R1, KD2 = split_kd(NS2), KH2 = NS2 / KD2;
    static constexpr int OFF_B1 = 0;
    static constexpr int OFF_A1 = OFF_B1 + (R1 > 1 ? R1 * KD1 : 0);
    static constexpr int OFF_B2 = OFF_A1 + ((R1 > 1 && KH1 > 1) ? R1 * KH1 : 0);
    static constexpr int OFF_A2 = OFF_B2 + (R2 > 1 ? R2 * KD2 : 0);
    static constexpr int BLOB = OFF_A2 + ((R2 > 1 && KH2 > 1) ? R2 * KH2 : 0);   // float2 entries
    static constexpr int BLOB_PAD = (BLOB + 1) & ~1;
};

template <class S, int STAGE> struct StageInfo;
template <class S> struct StageInfo<S, 0> {
    static constexpr int R = S::R0, NS = 1, KD = 1, KH = 1, OFF_B = 0, OFF_A = 0;
};
template <class S> struct StageInfo<S, 1> {
    static constexpr int R = S::R1, NS = S::NS1, KD = S::KD1, KH = S::KH1, OFF_B = S::OFF_B1, OFF_A = S::OFF_A1;
};
template <class S> struct StageInfo<S, 2> {
    static constexpr int R = S::R2, NS = S::NS2, KD = S::KD2, KH = S::KH2, OFF_B = S::OFF_B2, OFF_A = S::OFF_A2;
};

// Host: fill the twiddle blob of a spec (fp64 -> fp32).
template <class S> void fill_blob(float2* out)
{
    const double TWO_PI = 6.283185307179586476925287;
    auto fill = [&](int R, int NS, int KD, int KH, int offB, int offA) {
        if (R <= 1) return;
        double M = (double)NS * R;
        for (int m = 0; m < R; m++)
            for (int kl = 0; kl < KD; kl++) {
                double a = -TWO_PI * (double)(kl * m) / M;
                out[offB + m * KD + kl] = make_float2((float)cos(a), (float)sin(a));
            }
        if (KH > 1)
            for (int m = 0; m < R; m++)
                for (int kh = 0; kh < KH; kh++) {
                    double a = -TWO_PI * (double)((long long)KD * kh * m) / M;
                    out[offA + m * KH + kh] = make_float2((float)cos(a), (float)sin(a));
                }
    };
    fill(S::R1, S::NS1, S::KD1, S::KH1, S::OFF_B1, S::OFF_A1);
    fill(S::R2, S::NS2, S::KD2, S::KH2, S::OFF_B2, S::OFF_A2);
}

// One stage, first half: gather R inputs per butterfly through `src(index)`, apply the stage
// twiddles from the (shared-memory) blob `tw`, run the butterflies.  j = thread index in [0,NTS).
template <class S, int STAGE, class Src> HD void stage_compute(int j, float2* v, const float2* tw, Src src)
{
    typedef StageInfo<S, STAGE> I;
    constexpr int R = I::R, NS = I::NS, KD = I::KD, KH = I::KH;
    constexpr int NBF = S::L / R;            // butterflies in this stage
    constexpr int NB = NBF / S::NTS;         // butterflies per thread
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int jb = j + S::NTS * b;
        float2* vb = v + b * R;
#pragma unroll
        for (int m = 0; m < R; m++) vb[m] = src(jb + m * NBF);
        if (NS > 1) {
            const int k = jb % NS;
            const int kl = k % KD, kh = k / KD;
#pragma unroll
            for (int m = 1; m < R; m++) {
                float2 w = tw[I::OFF_B + m * KD + kl];
                if (KH > 1) w = cmul(w, tw[I::OFF_A + m * KH + kh]);
                vb[m] = cmul(vb[m], w);
            }
        }
        Bfly<R>::run(vb);
    }
}

// One stage, second half: scatter the butterfly outputs through `dst(index, value)`.
template <class S, int STAGE, class Dst> HD void stage_scatter(int j, const float2* v, Dst dst)
{
    typedef StageInfo<S, STAGE> I;
    constexpr int R = I::R, NS = I::NS;
    constexpr int NBF = S::L / R;
    constexpr int NB = NBF / S::NTS;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int jb = j + S::NTS * b;
        const int base = (jb / NS) * NS * R + (jb % NS);
#pragma unroll
        for (int m = 0; m < R; m++) dst(base + m * NS, v[b * R + m]);
    }
}

// For the LAST stage the output index is jb + m*NSLAST; this visits (out_index, value, b, m).
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
