// Fast peak path: the x3-interpolated correlation maximum from ONE inverse transform instead of three.
//
// The reference takes  [~,indice]=max(abs(ifft([P(1:N/2) zeros(2N) P(N/2+1:N)])))  over all 3N interpolated samples
// (processing/Octave/godual_ranging.m:26-29).  Sample 3n of that sequence is the plain N-point inverse x1[n] (one
// row + one column pass, k_rowcorr.cuh NQ = 1, k_col.cuh COL_INV_ARGMAX_K2); the others are its periodic-sinc
// interpolation
//     x3[3n+q] = sum_m x1[m] h_q(n-m),   |h_q(d)| = sin(pi q/3) / (N |sin(pi (d+q/3)/N)|) =: hb(|d+q/3|),  q = 1,2.
// Around the 1x maximum n* the interpolated lags 3n*-6 .. 3n*+6 are evaluated EXACTLY (directly from the stored
// product spectrum P, KLagP; the 1x samples of the row segment S around n* from the inverse row pass, KPeakEval).
// Every other interpolated sample is BOUNDED:
//     |x3[3n+q]| <= sum_{m in S} |x1[m]| hb(|n+q/3-m|)  +  A2 * H  +  sqrt(E0) * T
// with A2 the largest 1x magnitude outside S (tile maxima + second maxima of the column pass), H = sum_{|d|<=K} hb,
// T^2 = sum_{|d|>K} hb^2 (Cauchy-Schwarz against the total energy E0 = sum |x1|^2, Parseval on P).  If every bound is
// below the largest evaluated lag (with a 2e-3 margin for the fp32 transforms) the arg-max is proven to be that lag
// and KPeakFinal finishes exactly as in the three-polyphase path; otherwise the lane is flagged (status bit 2) and
// the host re-runs the record through the full path.  Results are therefore identical either way.
#pragma once
#include "launch.cuh"
#ifdef TWSTFT_EMU
#include <stdio.h>
#include <stdlib.h>
#endif

namespace twstft {

constexpr int FAST_NLAG = 8;                  // lags 3n*+j, j = +-1, +-2, +-4, +-5 (j = 0, +-3, +-6 are 1x samples)
// slot s of the lag arrays: j = {1,2,4,5,-1,-2,-4,-5}
HD int fast_lag_j(int s) { const int a[4] = {1, 2, 4, 5}; return s < 4 ? a[s] : -a[s - 4]; }

struct FastPrepParams {
    const u64* key1;       // [lane][ntile] tile maxima of the 1x correlation (index = n)
    const u64* key2;       // KFastA2: [lane][3] maxima of the three tiles of S over the rows other than their maximum's
    int ntile, tsi;        // tiles per row, columns per tile
    int N2, nrow;          // row length, rows (N = N2 * nrow)
    i64* nstar;            // [lane] 1x arg-max
    i64* base;             // [lane] first sample of S = the three column tiles around n* in linear time (may span two rows)
    int* tiles;            // [lane][3] column tiles of S (work list of the second-maximum launch)
    double* a2;            // [lane] largest |z| outside S (unnormalised units of the inverse transform)
    int* fail;             // [lane] 1: S would wrap around the array ends -> full path
};
// 1x arg-max, the segment S and its three column tiles
struct KFastPrep {
    typedef FastPrepParams Params;
    static const char* name() { return "fast_prep"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 2, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_MAXKEY; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        const u64* k1 = p.key1 + lane * p.ntile;
        if (PH == 0) {
            u64 best = 0;
            for (int t = tid; t < p.ntile; t += NT) best = k1[t] > best ? k1[t] : best;
            r.key = best;
        } else if (tid == 0) {
            const i64 ns = (i64)key_index(r.key);
            const i64 n = (i64)p.N2 * p.nrow;
            const i64 base = (ns / p.tsi) * p.tsi - p.tsi;           // S = [base, base + 3 tsi) in linear time
            p.nstar[lane] = ns;
            p.base[lane] = base < 0 ? 0 : base;
            p.fail[lane] = (base < 0 || base + 3 * (i64)p.tsi > n) ? 1 : 0;   // S would wrap the array
            for (int c = 0; c < 3; c++) {
                i64 b = base + (i64)c * p.tsi;
                b = b < 0 ? 0 : (b >= n ? n - 1 : b);
                p.tiles[lane * 3 + c] = (int)((b % p.N2) / p.tsi);
            }
        }
    }
};
// A2: the largest 1x magnitude outside S -- tile maxima elsewhere; in the three (row, tile) cells of S the maximum
// over the OTHER rows of that tile (COL_INV_ARGMAX_K2 run on those three tiles only)
struct KFastA2 {
    typedef FastPrepParams Params;
    static const char* name() { return "fast_a2"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 2, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_MAXKEY; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        const u64* k1 = p.key1 + lane * p.ntile;
        if (PH == 0) {
            const i64 ns = p.nstar[lane];
            const i64 base = (ns / p.tsi) * p.tsi - p.tsi;
            u64 best = 0;
            for (int t = tid; t < p.ntile; t += NT) {
                u64 k = k1[t];
                for (int c = 0; c < 3; c++) {
                    const i64 b = base + (i64)c * p.tsi;
                    if (b >= 0 && t == (int)((b % p.N2) / p.tsi) && (i64)key_index(k) / p.N2 == b / p.N2)
                        k = p.key2[lane * 3 + c];
                }
                const u64 v = (k >> 32) << 32;              // compare magnitudes only
                best = v > best ? v : best;
            }
            r.key = best;
        } else if (tid == 0) {
            p.a2[lane] = sqrt((double)key_value(r.key));
        }
    }
};

// per lane: R2[k2] = exp(2 pi i (k2 n* mod N2)/N2)
struct LagTabParams {
    const i64* nstar;
    int N2;
    float2* R2;            // [lane][N2]
};
struct KLagTables {
    typedef LagTabParams Params;
    static const char* name() { return "fast_lag_tables"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 1, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        const i64 lane = blk.y;
        const int k2 = blk.x * NT + tid;
        if (k2 >= p.N2) return;
        const i64 c = p.nstar[lane] % p.N2;
        p.R2[lane * p.N2 + k2] = cis_cycles((double)(((i64)k2 * c) % p.N2) / (double)p.N2);
    }
};

// Z[3n*+j] = sum_k P[k] exp(2 pi i ktilde (3n*+j)/3N), k = k1 + L1*k2, ktilde = k - N for k2 >= N2/2:
//   = sum_k1 R1[k1] u1[k1]^j  sum_k2 (P[k1][k2] R2[k2]) U2[k2]^j
// one CTA per (row k1, lane): the inner sums for the eight j in fp32 per thread, fp64 across the CTA.
struct LagPParams {
    const float2* P;       // [lane][L1][N2]
    const float2* R2;      // [lane][N2]
    const float2* U2;      // [N2] exp(2 pi i (L1 k2 - N [k2 >= N2/2]) / 3N)
    const i64* nstar;
    i64 N;
    int L1, N2;
    double* part;          // [lane][L1][2*FAST_NLAG]
};
struct KLagP {
    typedef LagPParams Params;
    static const char* name() { return "fast_lag_p"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 3, NPH = 2, NSUM = 1;
    static_assert(2 * FAST_NLAG == 16, "the sixteen partial sums go through SYNC_SUMF16");
    struct Regs { u64 key; double dsum[NSUM]; float fsum[16]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_SUMF16; }
    // one spectrum sample: a = P*R2, then a*w^j and a*conj(w^j) for j = 1, 2, 4, 5 share their products: with
    // w^j = c + i s,  a w = a c + i (a s),  a conj(w) = a c - i (a s); accumulate Sr = sum a c and Ss = sum a s (one
    // packed FMA each) and combine once at the end
    static HD void accumulate(float2 pv, float2 r2, float2 u, float2* sr, float2* ss)
    {
        const float2 a = cmul(pv, r2);
        float2 w[4];
        w[0] = u;
        w[1] = cmul(w[0], w[0]);
        w[2] = cmul(w[1], w[1]);
        w[3] = cmul(w[2], w[0]);
#pragma unroll
        for (int s = 0; s < 4; s++) {
            sr[s] = caxpy(w[s].x, a, sr[s]);
            ss[s] = caxpy(w[s].y, a, ss[s]);
        }
    }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        const int k1 = blk.x;
        if (PH == 0) {
            float2 sr[4], ss[4];
#pragma unroll
            for (int s = 0; s < 4; s++) sr[s] = ss[s] = make_float2(0.f, 0.f);
            // two spectrum samples per 16-byte load, four loads of each stream in flight per thread: the row of P
            // comes from HBM exactly once, the two tables (64 KB each) stay in L1/L2
            const float4* row = reinterpret_cast<const float4*>(p.P + (lane * p.L1 + k1) * (i64)p.N2);
            const float4* r2 = reinterpret_cast<const float4*>(p.R2 + lane * p.N2);
            const float4* u2 = reinterpret_cast<const float4*>(p.U2);
            const int npair = p.N2 / 2;
#ifndef TWSTFT_LAGP_U
#define TWSTFT_LAGP_U 8
#endif
            constexpr int U = TWSTFT_LAGP_U;
            // every pass requests U row pieces per thread before it uses the first; the last pass is predicated
            // instead of falling back to one load at a time (pieces past the end contribute zero)
            for (int i = tid; i < npair; i += U * NT) {
                float4 a[U], b[U], c[U];
#pragma unroll
                for (int u = 0; u < U; u++) a[u] = (i + u * NT < npair) ? row[i + u * NT] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int k = (i + u * NT < npair) ? i + u * NT : tid;
                    b[u] = r2[k]; c[u] = u2[k];
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    accumulate(make_float2(a[u].x, a[u].y), make_float2(b[u].x, b[u].y), make_float2(c[u].x, c[u].y), sr, ss);
                    accumulate(make_float2(a[u].z, a[u].w), make_float2(b[u].z, b[u].w), make_float2(c[u].z, c[u].w), sr, ss);
                }
            }
            if ((p.N2 & 1) && tid == 0) {            // odd row length: the last sample alone
                const int k2 = p.N2 - 1;
                accumulate(p.P[(lane * p.L1 + k1) * (i64)p.N2 + k2], p.R2[lane * p.N2 + k2], p.U2[k2], sr, ss);
            }
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const float2 lo = cadd(sr[s], rot_pi(ss[s])), hi = csub(sr[s], rot_pi(ss[s]));
                r.fsum[2 * s] = lo.x; r.fsum[2 * s + 1] = lo.y;
                r.fsum[2 * (s + 4)] = hi.x; r.fsum[2 * (s + 4) + 1] = hi.y;
            }
        } else if (tid < 2 * FAST_NLAG) {
            // raw row sums; the row factor R1[k1] u1[k1]^j is applied by the reduction in KFastCert (one thread per row)
            p.part[(lane * p.L1 + k1) * (2 * FAST_NLAG) + tid] = r.dsum[0];
        }
    }
};

// Certificate + hand-over to KPeakFinal: part7 = the seven lags pc-4 .. pc+2 (unnormalised sums), p0 = pc.
struct FastCertParams {
    const double* lagpart; // [lane][L1][2*FAST_NLAG]
    int L1;
    const double* spart;   // [lane][ns][2]  1x samples of S (unnormalised sums), S[i] = x1[base + i]
    int ns;
    const i64* nstar;
    const i64* base;
    const double* a2;
    const double* penergy; // [lane][nen] sum |P|^2 partials
    int nen;
    i64 N;
    double H, T;           // sum_{|d|<=K} hb, sqrt(sum_{|d|>K} hb^2)   (host, fp64)
    const double* hbtab;   // [2][2*ns+5]: hb(|dd + q/3|) for dd = -(ns+2) .. ns+2, q = 1, 2   (host, fp64)
    double* work;          // [lane][8] scratch between the phases
    int* fail;             // in: 1 = S leaves the row; out: 1 = not certified
    i64* p0;               // [lane] decided 0-based peak on the interpolated grid
    double* part7;         // [lane][7][2]
};
struct KFastCert {
    typedef FastCertParams Params;
    static const char* name() { return "fast_cert"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 4, NSUM = 2 * FAST_NLAG + 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int ph) { return ph == 0 ? SYNC_SUMD : ph == 1 ? SYNC_BARRIER : SYNC_MAXKEY; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        double* wk = p.work + lane * 8;                   // mstar, noise, worst0, ic, bad
        const double* sp = p.spart + lane * p.ns * 2;
        if (PH == 0) {
#pragma unroll
            for (int i = 0; i < NSUM; i++) r.dsum[i] = 0.0;
            const i64 nst = p.nstar[lane];
            for (int k1 = tid; k1 < p.L1; k1 += NT) {
                const double* s = p.lagpart + (lane * p.L1 + k1) * (2 * FAST_NLAG);
                const double c0 = (double)(((i64)k1 * (nst % p.N)) % p.N) / (double)p.N;     // R1[k1]
#pragma unroll
                for (int i = 0; i < FAST_NLAG; i++) {
                    const double2 w = cis_cycles_d(c0 + (double)k1 * (double)fast_lag_j(i) / (3.0 * (double)p.N));
                    const double2 z = dmul(make_double2(s[2 * i], s[2 * i + 1]), w);
                    r.dsum[2 * i] += z.x;
                    r.dsum[2 * i + 1] += z.y;
                }
            }
            for (int b = tid; b < p.nen; b += NT) r.dsum[2 * FAST_NLAG] += p.penergy[lane * p.nen + b];
        } else if (PH == 1) {
            r.key = 0;
            if (tid != 0) return;
            const double n = (double)p.N;
            const i64 ns = p.nstar[lane], base = p.base[lane];
            const int ic = (int)(ns - base);                 // position of n* inside S
            int bad = p.fail[lane];
            if (ic < 2 || ic > p.ns - 3) bad = 1;
            double zr[13], zi[13], a2[13];                   // lags j = -6 .. 6 around 3 n*
            for (int j = -6; j <= 6; j++) {
                double re, im;
                if (j % 3 == 0) {
                    int i = ic + j / 3;
                    if (i < 0) i = 0;
                    if (i > p.ns - 1) i = p.ns - 1;
                    re = sp[i * 2];
                    im = sp[i * 2 + 1];
                } else {
                    const int aj = j < 0 ? -j : j;
                    const int s = (aj == 1 ? 0 : aj == 2 ? 1 : aj == 4 ? 2 : 3) + (j < 0 ? 4 : 0);
                    re = r.dsum[2 * s];
                    im = r.dsum[2 * s + 1];
                }
                zr[j + 6] = re; zi[j + 6] = im; a2[j + 6] = re * re + im * im;
            }
            int jc = 0;                                      // arg-max among the candidates j = -2 .. 2
            for (int j = -2; j <= 2; j++)
                if (a2[j + 6] > a2[jc + 6]) jc = j;
            for (int j = -6; j <= 6; j++)
                if (j != jc && a2[j + 6] >= a2[jc + 6]) bad = 1;             // a tie or a larger lag further out
            double worst = p.a2[lane];                       // 1x samples that are not candidates: outside S, and in S
            for (int i = 0; i < p.ns; i++)
                if (i < ic - 2 || i > ic + 2) {
                    const double m = sqrt(sp[i * 2] * sp[i * 2] + sp[i * 2 + 1] * sp[i * 2 + 1]);
                    worst = m > worst ? m : worst;
                }
            wk[0] = sqrt(a2[jc + 6]);
            wk[1] = p.a2[lane] * p.H + sqrt(n * r.dsum[2 * FAST_NLAG]) * p.T;
            wk[2] = worst;
            wk[3] = (double)ic;
            wk[4] = (double)bad;
            i64 pc = 3 * ns + jc;
            const i64 M = 3 * p.N;
            if (pc < 0) pc += M;
            if (pc >= M) pc -= M;
            p.p0[lane] = pc;
            for (int d = 0; d < 7; d++) {                    // pc-4 .. pc+2  (PEAK_DLO = -4, PEAK_NPTS = 7)
                int j = jc - 4 + d;
                if (j < -6) j = -6;
                if (j > 6) j = 6;
                p.part7[(lane * 7 + d) * 2] = zr[j + 6];
                p.part7[(lane * 7 + d) * 2 + 1] = zi[j + 6];
            }
        } else if (PH == 2) {
            // bound of every interpolated position t = pos + q/3, pos = -2 .. ns (relative to S) that was not
            // evaluated; positions further out have a smaller S term and the same noise term
            const int ic = (int)wk[3];
            const int ntab = 2 * p.ns + 5;
            double worst = 0.0;
            for (int idx = tid; idx < 2 * (p.ns + 3); idx += NT) {
                const int pos = -2 + idx / 2, q = 1 + idx % 2;
                const int j = 3 * (pos - ic) + q;
                if (j >= -6 && j <= 6) continue;
                const double* hb = p.hbtab + (q - 1) * ntab + (p.ns + 2);
                double ssum = 0.0;
                for (int i = 0; i < p.ns; i++)
                    ssum += sqrt(sp[i * 2] * sp[i * 2] + sp[i * 2 + 1] * sp[i * 2 + 1]) * hb[pos - i];
                worst = ssum > worst ? ssum : worst;
            }
            r.key = make_key((float)(worst * (1.0 + 1e-6)), 0u);
        } else if (tid == 0) {
            const double worst_s = (double)key_value(r.key) + wk[1];
            const double worst = worst_s > wk[2] ? worst_s : wk[2];
            int bad = (int)wk[4];
            if (!(worst < (1.0 - 2e-3) * wk[0])) bad = 1;
#ifdef TWSTFT_EMU
            if (getenv("TWSTFT_FAST_DEBUG"))
                printf("fast lane %lld: n*=%lld ic=%d M*=%.4g A2=%.4g noise=%.4g worst=%.4g bad=%d\n", (long long)lane,
                       (long long)p.nstar[lane], (int)wk[3], wk[0], p.a2[lane], wk[1], worst, bad);
#endif
            p.fail[lane] = bad;
        }
    }
};

}  // namespace twstft
