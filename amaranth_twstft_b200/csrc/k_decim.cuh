// Whole-file carrier estimator of the C++ program (GoRanging::df, processing/CPP/main.cpp:363-450):
// one sample instant out of `dec` over the whole recording, derotated by foffset, raw mean removed, squared, ONE
// transform of L = instants/dec points per channel, arg-max of the fftshift'ed magnitude (channel 1 inside
// [kmin,kmax), channel 2 everywhere).  L is whatever the file length gives (1.2e8 for ten minutes), so the
// transform is split L = s*M with M a length the plan's four-step kernels support:
//   n = s*n2 + n1, k = kk + M*k1:
//   X[kk + M*k1] = sum_{n1<s} W_s^{n1 k1} * ( W_L^{n1 kk} * sum_{n2<M} y[s n2 + n1] W_M^{n2 kk} )
// KDecSquare builds the s decimated sub-sequences, the plan's column/row kernels transform them (M points each),
// KDecTwiddle applies W_L^{n1 kk}, KDecCombine evaluates the s-point transforms directly (s is the number of
// M-sample chunks of the decimated file, a few hundred) and reduces |X|^2 to first-maximum keys.
#pragma once
#include "launch.cuh"

namespace twstft {

struct DecSquareParams {
    const short2* g;       // [L][nch] decimated complex-int16 samples
    int nch, ch;
    i64 L;
    int s;                 // sub-sequences
    i64 M;                 // points per sub-sequence
    double mean_re, mean_im;   // raw mean of the decimated samples (main.cpp:381,407: NOT derotated)
    double cyc_per_sample;     // foffset*dec/fs: lo[n] = exp(-j 2 pi cyc n)
    float2* dst;           // [n1][n2] = ((g[s n2 + n1]*lo - mean))^2
};
struct KDecSquare {
    typedef DecSquareParams Params;
    static const char* name() { return "dec_square"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 1, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        const i64 n = (i64)blk.x * NT + tid;
        if (n >= p.L) return;
        const short2 raw = p.g[n * p.nch + p.ch];
        double xr = (double)raw.x, xi = (double)raw.y;
        if (p.cyc_per_sample != 0.0) {
            const double2 lo = cis_cycles_d(-p.cyc_per_sample * (double)n);
            const double tr = xr * lo.x - xi * lo.y;
            xi = xr * lo.y + xi * lo.x;
            xr = tr;
        }
        xr -= p.mean_re;
        xi -= p.mean_im;
        p.dst[(n % p.s) * p.M + n / p.s] = make_float2((float)(xr * xr - xi * xi), (float)(2.0 * xr * xi));
    }
};

// spectrum position e = k1c*N2 + k2c of an M-point four-step transform holds bin kk = k1c + L1*k2c
struct DecTwiddleParams {
    float2* G;             // [n1][M] spectra of the sub-sequences (transposed bin order), in place
    int s;
    i64 M, L;
    int L1, N2;
};
struct KDecTwiddle {
    typedef DecTwiddleParams Params;
    static const char* name() { return "dec_twiddle"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 1, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        const i64 e = (i64)blk.x * NT + tid;
        const i64 n1 = blk.y;
        if (e >= p.M || n1 == 0) return;
        const i64 kk = e / p.N2 + (i64)p.L1 * (e % p.N2);
        const float2 w = cis_cycles(-(double)((n1 * kk) % p.L) / (double)p.L);
        float2* x = p.G + n1 * p.M + e;
        *x = cmul(*x, w);
    }
};

struct DecCombineParams {
    const float2* G;       // [n1][M], twiddled
    const float2* wtab;    // [s] W_s^j = exp(-2 pi i j/s)
    int s;
    i64 M, L;
    int L1, N2;
    i64 win_lo, win_hi;    // inclusive window on the shifted axis ([kmin,kmax-1] or [0,L-1])
    u64* part_key;         // [gridDim.y][gridDim.x]
};
struct KDecCombine {
    typedef DecCombineParams Params;
    static const char* name() { return "dec_combine"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 2, NPH = 2, NSUM = 1, KT = 8;   // KT outputs k1 per thread
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_MAXKEY; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        if (PH == 1) {
            if (tid == 0) p.part_key[(i64)blk.y * blk.gx + blk.x] = r.key;
            return;
        }
        r.key = 0;
        const i64 e = (i64)blk.x * NT + tid;
        if (e >= p.M) return;
        const int k10 = blk.y * KT;
        float2 acc[KT];
        int idx[KT];
#pragma unroll
        for (int i = 0; i < KT; i++) { acc[i] = make_float2(0.f, 0.f); idx[i] = 0; }
        const float2* g = p.G + e;
        for (int n1 = 0; n1 < p.s; n1++) {
            const float2 x = g[(i64)n1 * p.M];
#pragma unroll
            for (int i = 0; i < KT; i++) {
                const float2 w = p.wtab[idx[i]];          // W_s^{n1*k1 mod s}
                acc[i] = cadd(acc[i], cmul(x, w));
                idx[i] += (k10 + i < p.s) ? k10 + i : 0;
                if (idx[i] >= p.s) idx[i] -= p.s;
            }
        }
        const i64 kk = e / p.N2 + (i64)p.L1 * (e % p.N2);
        const i64 h = p.L / 2;
        u64 best = 0;
#pragma unroll
        for (int i = 0; i < KT; i++) {
            if (k10 + i >= p.s) continue;
            const i64 k = kk + p.M * (i64)(k10 + i);
            // out = [f[h:2h], f[:h], f[2h:]]  (main.cpp:413-414: two memcpy of L/2 elements)
            const i64 sidx = k < h ? k + h : (k < 2 * h ? k - h : k);
            if (sidx < p.win_lo || sidx > p.win_hi) continue;
            const u64 key = make_key(cnorm(acc[i]), (unsigned)sidx);
            best = key > best ? key : best;
        }
        r.key = best;
    }
};

}  // namespace twstft
