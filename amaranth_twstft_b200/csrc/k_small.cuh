// Small kernels around the FFT passes: window mean, carrier-offset finalisation (+ derotation
// tables), peak refinement (exact fp64 evaluation of the interpolated correlation around the coarse
// maximum), parabolic fine delay, phase and the "Claudio" SNR, and the LFSR PRN generator.
#pragma once
#include "fft_core.cuh"

namespace twstft {

HD void atomic_add_d(double* p, double v)
{
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
HD double mul_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
HD double add_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

// ------------------------------------------------------------------------------------------------
// KCvtF64: complex double -> complex float, for callers that hold their windows in floating point already (the
// Octave seam processing(d,k), godual_ranging.m:12: `d` is a complex double column).
struct CvtParams {
    const double2* src;
    float2* dst;
    i64 n;
};
struct KCvtF64 {
    typedef CvtParams Params;
    static const char* name() { return "cvt_f64"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 1, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        for (i64 i = (i64)blk.x * NT + tid; i < p.n; i += (i64)blk.gx * NT) {
            const double2 v = p.src[i];
            p.dst[i] = make_float2((float)v.x, (float)v.y);
        }
    }
};

// ------------------------------------------------------------------------------------------------
// KMean: exact sums of I and Q over each window (godual_ranging.m:80 d1=d1-mean(d1); the means are
// applied inside the column pass).  Sums of int16 over <= 1e7 samples stay below 2^53, so double
// accumulation (and double atomics, in any order) is exact.
struct MeanParams {
    const short2* src;
    i64 nelem;             // complex-int16 elements in the window, all channels (= N * nch)
    int nch;
    i64 rec_stride;
    const i64* rec_offset;
    double* sums;          // [record*nch + ch][2], zeroed by the caller
    int nblk;
};
struct KMean {
    typedef MeanParams Params;
    static const char* name() { return "mean"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 2, NSUM = 4;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_SUMD; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        if (PH == 0) {
            const i64 base = p.rec_offset ? p.rec_offset[blk.y] : (i64)blk.y * p.rec_stride;
            const short2* s = p.src + base;
            const i64 stride = (i64)p.nblk * NT;      // even, so a thread always sees the same channel
            const i64 e0 = (i64)blk.x * NT + tid;
            i64 a0x = 0, a0y = 0, a1x = 0, a1y = 0;   // channel 0 / channel 1 (nch == 2) accumulators
            if ((base & 3) == 0 && (p.nelem & 3) == 0 && ((size_t)p.src & 15) == 0) {
                // 16-byte loads: 4 complex-int16 elements = [ch0 ch1 ch0 ch1] (nch 2) or 4 x ch0 (nch 1)
                const int4* s4 = reinterpret_cast<const int4*>(s);
                const i64 nquad = p.nelem >> 2;
                int q0x = 0, q0y = 0, q1x = 0, q1y = 0;
                auto add = [&](const int4 v) {
                    q0x += (int)(short)(v.x & 0xffff) + (int)(short)(v.z & 0xffff);
                    q0y += (v.x >> 16) + (v.z >> 16);
                    q1x += (int)(short)(v.y & 0xffff) + (int)(short)(v.w & 0xffff);
                    q1y += (v.y >> 16) + (v.w >> 16);
                };
                i64 e = e0;
                constexpr int UB = 8;                         // eight independent 16-byte loads in flight per thread
                for (; e + (UB - 1) * stride < nquad; e += UB * stride) {
                    int4 v[UB];
#pragma unroll
                    for (int u = 0; u < UB; u++) v[u] = s4[e + u * stride];
#pragma unroll
                    for (int u = 0; u < UB; u++) add(v[u]);
                }
                for (; e < nquad; e += stride) add(s4[e]);
                if (p.nch == 2) { a0x = q0x; a0y = q0y; a1x = q1x; a1y = q1y; }
                else { a0x = (i64)q0x + q1x; a0y = (i64)q0y + q1y; }
            } else {
                i64 sx = 0, sy = 0;
#pragma unroll 4
                for (i64 e = e0; e < p.nelem; e += stride) {
                    const short2 v = s[e];
                    sx += v.x;
                    sy += v.y;
                }
                if ((int)(e0 % p.nch) == 0) { a0x = sx; a0y = sy; } else { a1x = sx; a1y = sy; }
            }
            r.dsum[0] = (double)a0x;
            r.dsum[1] = (double)a0y;
            r.dsum[2] = (double)a1x;
            r.dsum[3] = (double)a1y;
        } else if (tid == 0) {
            double* o = p.sums + (i64)blk.y * p.nch * 2;
            atomic_add_d(o + 0, r.dsum[0]);
            atomic_add_d(o + 1, r.dsum[1]);
            if (p.nch > 1) {
                atomic_add_d(o + 2, r.dsum[2]);
                atomic_add_d(o + 3, r.dsum[3]);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// KBinsDFT: a few ADJACENT bins of fft(x.^p) over a buffer of ANY length, evaluated directly in fp64.
// The re/ranging production scripts re-estimate the carrier on d=[dold ; block] -- a transform whose
// length changes from block to block and is not 2/5-smooth -- but look at seven bins only:
//   d2=fftshift(abs(fft(d(1:end).^2))); [~,df]=max(d2(kbon-3:kbon+3))
//   acquisition/claudio_aligned_code_re_separate.m:167-168 (= ..._ranging_separate.m)
// X[k0+i] = sum_n x[n]^p W^{(k0+i) n}, W = exp(-2 pi i/L): per sample one exactly reduced phase
// (k0*n mod L in 64-bit integers), one step phase and nb-1 complex multiplies; one pass over the buffer.
struct BinsParams {
    const short2* src;     // [ninst][nch] complex int16
    i64 ninst;             // transform length L (sample instants)
    int nch;
    int power;             // 2 or 4
    i64 k0;                // first bin, 0 <= k0 < L; bins (k0+i) mod L
    int nb;                // bins, <= KBinsDFT::MAXB
    double* part;          // [ch][nblk][2*MAXB] partial sums (re, im per bin)
    int nblk;
};
struct KBinsDFT {
    typedef BinsParams Params;
    static const char* name() { return "bins_dft"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int MAXB = 8;
    static constexpr int NT = 256, MINB = 2, NPH = 2, NSUM = 2 * MAXB;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_SUMD; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        if (PH == 0) {
            double ar[MAXB], ai[MAXB];
#pragma unroll
            for (int i = 0; i < MAXB; i++) ar[i] = ai[i] = 0.0;
            const int ch = blk.y;
            const double invL = 1.0 / (double)p.ninst;
            const i64 stride = (i64)p.nblk * NT;
            for (i64 n = (i64)blk.x * NT + tid; n < p.ninst; n += stride) {
                const short2 v = p.src[n * p.nch + ch];
                const double xr = (double)v.x, xi = (double)v.y;
                double zr = xr * xr - xi * xi, zi = 2.0 * xr * xi;          // x^2 (exact: |x| < 2^15)
                if (p.power == 4) { const double t = zr * zr - zi * zi; zi = 2.0 * zr * zi; zr = t; }
                const i64 ph = (i64)(((u64)p.k0 * (u64)n) % (u64)p.ninst);  // k0*n < 2^63 for L < 2^31
                double2 w = cis_cycles_d(-(double)ph * invL);
                const double2 ws = cis_cycles_d(-(double)n * invL);
#pragma unroll
                for (int i = 0; i < MAXB; i++) {
                    if (i < p.nb) {
                        ar[i] += zr * w.x - zi * w.y;
                        ai[i] += zr * w.y + zi * w.x;
                        w = dmul(w, ws);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < MAXB; i++) { r.dsum[2 * i] = ar[i]; r.dsum[2 * i + 1] = ai[i]; }
        } else if (tid == 0) {
            double* o = p.part + ((i64)blk.y * p.nblk + blk.x) * (2 * MAXB);
#pragma unroll
            for (int i = 0; i < 2 * MAXB; i++) o[i] = r.dsum[i];
        }
    }
};

// ------------------------------------------------------------------------------------------------
// KDfFinal: reduce the per-row window maxima of the |fft(x^p)| spectrum to the carrier bin,
// df = freq(s)/p with the reference's inclusive linspace axis (godual_ranging.m:15,73 -- the
// "+0.25 Hz" grid quirk; claudio_aligned_code_lo_separate.m:103,129), then build the two-level
// derotation tables exp(-j2pi df n/fs), n = n1*N2 + n2  (godual_ranging.m:17).
struct DfParams {
    const u64* part_key;   // [lane*key_stride + i], i < npart
    int npart, key_stride;
    const double* df_in;   // optional [lane]: skip the estimator (variant K processing(d,k,df))
    double f0, fstep, flast;   // linspace(start, stop, n): start + s*step, last point = stop exactly
    i64 naxis;
    double power;          // 2 (BPSK) or 4 (QPSK)
    double fs;
    int L1, N2;
    double* df_out;        // [lane]
    i64* s_out;            // [lane] index into the fftshift'ed axis (-1 when df_in is used)
    float2* lrow;          // [lane][L1]
    float2* lcol;          // [lane][N2]
    int nsplit;
};
struct KDfFinal {
    typedef DfParams Params;
    static const char* name() { return "df_final"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 2, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_MAXKEY; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        if (PH == 0) {
            u64 best = 0;
            if (!p.df_in)
                for (int i = tid; i < p.npart; i += NT) {
                    const u64 k = p.part_key[lane * p.key_stride + i];
                    best = k > best ? k : best;
                }
            r.key = best;
            r.dsum[0] = 0.0;
        } else {
            double df;
            i64 s = -1;
            if (p.df_in) df = p.df_in[lane];
            else {
                s = (i64)key_index(r.key);
                const double f = (s == p.naxis - 1) ? p.flast : add_rn(mul_rn((double)s, p.fstep), p.f0);
                df = f / p.power;
            }
            if (blk.x == 0 && tid == 0) {
                p.df_out[lane] = df;
                if (p.s_out) p.s_out[lane] = s;
            }
            const int tot = p.L1 + p.N2;
            for (int e = blk.x * NT + tid; e < tot; e += p.nsplit * NT) {
                if (e < p.L1) p.lrow[lane * p.L1 + e] = cis_cycles(-df * ((double)((i64)e * p.N2) / p.fs));
                else p.lcol[lane * p.N2 + (e - p.L1)] = cis_cycles(-df * ((double)(e - p.L1) / p.fs));
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Fine-frequency refinement: experiments/221219_twoway/processing/godual_ranging.py:26-30
// (= the block commented out at processing/Octave/godual_ranging.m:19-24):
//     a = polyfit([1:10:fs/3]/fs, conv(angle(y(1:10:fs/3)), ones(100,1)/100)(50:end-50), 1)
//     dfleftover = a(1)/2/pi ;  y = y .* exp(-j*2*pi*dfleftover*temps) ;  df += dfleftover
// KFineAngle evaluates angle(y) of every `step`-th derotated, mean-removed sample in fp64 (the angles are
// NOT unwrapped by the reference, so a sample that lands on the other side of +/-pi would move the fit:
// fp64 keeps that as unlikely as in numpy); KFineFit forms the centred least-squares numerator
// sum (i - c) s_i of the 100-tap smoothed angles s_i (the regressor is equally spaced, so the slope is
// fs/step * sum (i-c) s_i / (M (M^2-1)/12)); KFineFinal adds the slope/2pi to df.
struct FineParams {
    const short2* src;
    int nch;
    i64 rec_stride;
    const i64* rec_offset;
    const double* sums;    // [lane][2] window sums (means = sums * invN)
    double invN;
    double fs;
    int M, step, taps;     // M samples at stride `step`, moving average of `taps` (100)
    double* ang;           // [lane][M]
    double* part;          // [lane][nsplit]
    int nsplit;
    double* df;            // [lane] in: coarse df, out (KFineFinal): df + dfleftover
};
struct KFineAngle {
    typedef FineParams Params;
    static const char* name() { return "fine_angle"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 1, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        const i64 lane = blk.y;
        const int rec = (int)(lane / p.nch), ch = (int)(lane % p.nch);
        const i64 i = (i64)blk.x * NT + tid;
        if (i >= p.M) return;
        const i64 base = p.rec_offset ? p.rec_offset[rec] : (i64)rec * p.rec_stride;
        const i64 n = i * p.step;
        const short2 raw = p.src[base + n * p.nch + ch];
        const double x = (double)raw.x - p.sums[lane * 2] * p.invN, y = (double)raw.y - p.sums[lane * 2 + 1] * p.invN;
        const double2 lo = cis_cycles_d(-p.df[lane] * ((double)n / p.fs));
        p.ang[lane * p.M + i] = atan2(x * lo.y + y * lo.x, x * lo.x - y * lo.y);
    }
};
struct KFineFit {
    typedef FineParams Params;
    static const char* name() { return "fine_fit"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 2, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_SUMD; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        if (PH == 0) {
            // thread t of CTA x owns the contiguous run [i0, i1) of smoothed samples: one window sum, then slide
            const i64 nthr = (i64)p.nsplit * NT, per = (p.M + nthr - 1) / nthr;
            const i64 i0 = ((i64)blk.x * NT + tid) * per;
            i64 i1 = i0 + per;
            if (i1 > p.M) i1 = p.M;
            const double* a = p.ang + lane * p.M;
            const int before = p.taps / 2, after = p.taps - before - 1;     // window [i-50, i+49]
            const double c = 0.5 * (double)(p.M - 1), w = 1.0 / (double)p.taps;
            double acc = 0.0, win = 0.0;
            for (i64 i = i0; i < i1; i++) {
                if (i == i0 || ((i - i0) & 63) == 0) {                      // fresh sum now and then: no drift
                    win = 0.0;
                    const i64 lo = i - before < 0 ? 0 : i - before, hi = i + after > p.M - 1 ? p.M - 1 : i + after;
                    for (i64 l = lo; l <= hi; l++) win += a[l];
                } else {
                    if (i + after <= p.M - 1) win += a[i + after];
                    if (i - before - 1 >= 0) win -= a[i - before - 1];
                }
                acc += ((double)i - c) * (win * w);
            }
            r.key = 0;
            r.dsum[0] = acc;
        } else if (tid == 0) {
            p.part[lane * p.nsplit + blk.x] = r.dsum[0];
        }
    }
};
struct KFineFinal {
    typedef FineParams Params;
    static const char* name() { return "fine_final"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 32, MINB = 4, NPH = 1, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        if (tid != 0) return;
        const i64 lane = blk.y;
        double num = 0.0;
        for (int i = 0; i < p.nsplit; i++) num += p.part[lane * p.nsplit + i];
        const double M = (double)p.M;
        const double slope = (p.fs / (double)p.step) * num / (M * (M * M - 1.0) / 12.0);   // rad / s
        p.df[lane] += slope / 6.283185307179586476925287;
    }
};

// ------------------------------------------------------------------------------------------------
// Peak refinement.  With P the (transposed-order) product spectrum, the interpolated correlation is
//   prnmap[p] = (1/M) sum_k P[k] exp(2 pi i ktilde p / M),  M = nq*N, ktilde = k - N for k >= N/2
// (godual_ranging.m:26-28).  Around the coarse maximum p0 found by the fp32 inverse FFT we re-evaluate
// NPTS points p0+DLO .. p0+DLO+NPTS-1 from the inverse row-pass output with fp64 accumulation.
constexpr int PEAK_NPTS = 7;
constexpr int PEAK_DLO = -4;

struct PeakEvalParams {
    const u64* part_key;   // [lane][npart] first-max keys of the column-pass tiles
    int npart;
    const float2* B;       // [lane][nq][L1][N2] inverse row-pass output
    i64 N;
    int nq;
    int L1, N2;
    i64* p0_out;           // [lane]
    double* part;          // [lane][NPTS][2]
    const i64* p0_in;      // optional [lane]: evaluate around this index instead of the tiles' maximum
    int npts_rt, dlo_rt;   // npts_rt > 0: grid.x = npts_rt lags starting at p0 + dlo_rt (instead of NPTS / DLO)
};
// One CTA per (lag, lane): prnmap[p] = sum_k1 B[q][k1][n2] exp(+2 pi i k1 n1 / L1) with
// p = nq*(n1*N2 + n2) + q, i.e. one output of the inverse column transform evaluated directly
// (fp64 twiddles and accumulation) -- a few thousand terms instead of a pass over the spectrum.
template <int NPTS, int DLO> struct KPeakEvalT {
    typedef PeakEvalParams Params;
    static const char* name() { return "peak_eval"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 128, MINB = 4, NPH = 3, NSUM = 2;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int ph) { return ph == 0 ? SYNC_MAXKEY : SYNC_SUMD; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        const int d = blk.x;
        if (PH == 0) {
            u64 best = 0;
            for (int i = tid; i < p.npart; i += NT) {
                const u64 k = p.part_key[lane * p.npart + i];
                best = k > best ? k : best;
            }
            r.key = best;
            r.dsum[0] = r.dsum[1] = 0.0;
        } else if (PH == 1) {
            const i64 M = (i64)p.nq * p.N;
            const i64 p0 = p.p0_in ? p.p0_in[lane] : (i64)key_index(r.key);
            i64 pp = (p0 + (p.npts_rt ? p.dlo_rt : DLO) + d) % M;
            if (pp < 0) pp += M;
            const int q = (int)(pp % p.nq);
            const i64 m = pp / p.nq;
            const i64 n1 = m / p.N2, n2 = m % p.N2;
            const float2* src = p.B + ((lane * p.nq + q) * p.L1) * (i64)p.N2 + n2;
            double sr = 0.0, si = 0.0;
            for (int k1 = tid; k1 < p.L1; k1 += NT) {
                const float2 x = src[(i64)k1 * p.N2];
                const double2 w = cis_cycles_d((double)((k1 * n1) % p.L1) / (double)p.L1);
                sr += (double)x.x * w.x - (double)x.y * w.y;
                si += (double)x.x * w.y + (double)x.y * w.x;
            }
            r.dsum[0] = sr;
            r.dsum[1] = si;
        } else if (tid == 0) {
            const int np = p.npts_rt ? p.npts_rt : NPTS;
            p.part[(lane * np + d) * 2] = r.dsum[0];
            p.part[(lane * np + d) * 2 + 1] = r.dsum[1];
            if (d == 0 && !p.p0_in) p.p0_out[lane] = (i64)key_index(r.key);
        }
    }
};

typedef KPeakEvalT<PEAK_NPTS, PEAK_DLO> KPeakEval;

// One result row per lane = one row of the reference's .mat variables.
struct Result {
    double indice;          // 1-based index on the interpolated grid (Octave convention)
    double correction;      // parabolic vertex offset, interpolated-grid units
    double xval_re, xval_im, xvalm1_re, xvalm1_im, xvalp1_re, xvalp1_im;
    double SNRr, SNRi;
    double df;
    double puissance, puissancecode, puissancenoise;
    double carrier_bin;     // 0-based index into the fftshift'ed estimator axis (-1: df was given)
    double status;          // bit0: peak at an edge (reference would index out of range), bit1: refined peak moved
};
constexpr int RESULT_NDOUBLE = sizeof(Result) / sizeof(double);

struct PeakFinParams {
    const double* part;        // [lane][nblk][2*NPTS]
    int nblk;
    const double* part_energy; // [lane][nen]
    int nen;
    const float2* y0;          // [lane]
    const i64* p0;             // [lane]
    const double* df;          // [lane]
    const i64* sbin;           // [lane]
    i64 N;
    int nq;
    int variant;               // 0 = G (godual_ranging.m), 1 = K (claudio_aligned_code_*_separate.m)
    int ddof;                  // 1 = Octave var, 0 = numpy var
    double code_pow;           // mean |code|^2 (sum|yincode|^2 = code_pow * sum|yint|^2)
    Result* out;               // [lane]
    i64* pc_out;               // optional [lane]: the decided 0-based peak index
    const int* fail;           // optional [lane]: fast peak path not certified -> status bit 2 (the host re-runs the record)
    int snr_only;              // 1: the lags are those of the UNWINDOWED correlation around the already decided
                               //    peak p0 (C++ flavour, processing/CPP/main.cpp:319-357): only the SNR / power
                               //    fields of `out` are rewritten
};
struct KPeakFinal {
    typedef PeakFinParams Params;
    static const char* name() { return "peak_final"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 2, NSUM = 2 * PEAK_NPTS + 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_SUMD; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        if (PH == 0) {
#pragma unroll
            for (int i = 0; i < NSUM; i++) r.dsum[i] = 0.0;
            for (int b = tid; b < p.nblk; b += NT) {
                const double* s = p.part + (lane * p.nblk + b) * (2 * PEAK_NPTS);
#pragma unroll
                for (int i = 0; i < 2 * PEAK_NPTS; i++) r.dsum[i] += s[i];
            }
            for (int b = tid; b < p.nen; b += NT) r.dsum[2 * PEAK_NPTS] += p.part_energy[lane * p.nen + b];
            return;
        }
        if (tid != 0) return;
        const i64 M = (i64)p.nq * p.N;
        const double invM = 1.0 / (double)M;
        double sr[PEAK_NPTS], si[PEAK_NPTS], a2[PEAK_NPTS];
#pragma unroll
        for (int d = 0; d < PEAK_NPTS; d++) {
            sr[d] = r.dsum[2 * d] * invM;
            si[d] = r.dsum[2 * d + 1] * invM;
            a2[d] = sr[d] * sr[d] + si[d] * si[d];
        }
        const i64 p0 = p.p0[lane];
        // fp64 re-decision among p0-1, p0, p0+1 with first-max semantics on the (wrapped) index
        const int c0 = -PEAK_DLO;
        int c = c0;
        i64 pc = p0;
        for (int t = -1; t <= 1 && !p.snr_only; t += 2) {
            i64 pt = (p0 + t) % M;
            if (pt < 0) pt += M;
            const int ct = c0 + t;
            if (a2[ct] > a2[c] || (a2[ct] == a2[c] && pt < pc)) { c = ct; pc = pt; }
        }
        double status = (c != c0) ? 2.0 : 0.0;
        const double am1 = sqrt(a2[c - 1]), a0 = sqrt(a2[c]);
        int cp = c + 1;
        if (p.variant == 1 && pc == M - 2) cp = c + 2;     // lo_separate.m:51-55 strict '<' quirk
        if (cp >= PEAK_NPTS) { cp = PEAK_NPTS - 1; status += 1.0; }
        if (p.variant == 0 && (pc == 0 || pc == M - 1)) status += 1.0;   // Octave would index out of range
        const double ap1 = sqrt(a2[cp]);
        Result o;
        if (p.snr_only) o = p.out[lane];
        if (p.pc_out) p.pc_out[lane] = pc;
        if (!p.snr_only) {
        o.indice = (double)(pc + 1);
        o.correction = (am1 - ap1) / (am1 + ap1 - 2.0 * a0) / 2.0;
        o.xval_re = sr[c]; o.xval_im = si[c];
        o.xvalm1_re = sr[c - 1]; o.xvalm1_im = si[c - 1];
        o.xvalp1_re = sr[cp]; o.xvalp1_im = si[cp];
        o.status = status;
        }
        // "Claudio" SNR through the spectral identity: the mean of yincode is the sum of
        // three adjacent correlation lags / M; sum|yint|^2 follows from Parseval.
        double mr, mi;
        if (p.variant == 0) {            // godual_ranging.m:43 rotate the signal by indice-2
            mr = (sr[c - 1] + sr[c] + sr[c + 1 < PEAK_NPTS ? c + 1 : c]) * invM;
            mi = (si[c - 1] + si[c] + si[c + 1 < PEAK_NPTS ? c + 1 : c]) * invM;
        } else {                         // lo_separate.m:65-69 rotate the code by indice-2 when indice>2
            const int a = (pc == 0) ? c : c - 1;
            mr = (sr[a] + sr[a - 1] + sr[a - 2]) * invM;
            mi = -(si[a] + si[a - 1] + si[a - 2]) * invM;
        }
        const double E = r.dsum[2 * PEAK_NPTS];
        const double n = (double)p.N, m = (double)M;
        const double var_c = (p.code_pow * E / m - m * (mr * mr + mi * mi)) / (m - (double)p.ddof);
        const float2 y0 = p.y0[lane];
        const double y0n = (double)y0.x * (double)y0.x + (double)y0.y * (double)y0.y;
        o.SNRr = mr * mr / var_c;
        o.SNRi = mi * mi / var_c;
        o.df = p.df[lane];
        o.puissance = p.snr_only ? E : (E / n - y0n / n) / (n - (double)p.ddof);   // C++: sum |Y|^2 (main.cpp:295)
        o.puissancecode = mr * mr + mi * mi;
        o.puissancenoise = var_c;
        o.carrier_bin = (double)p.sbin[lane];
        if (p.fail && p.fail[lane]) o.status += 4.0;
        p.out[lane] = o;
    }
};

// ------------------------------------------------------------------------------------------------
// KCodeWindow: the C++ program multiplies the conjugated code spectrum by hamming(N) over the BIN index
// (processing/CPP/main.cpp:717-719; SigPack: 0.54 - 0.46 cos(2 pi k/(N-1))).  Position (k1,k2) of the transposed
// spectrum holds bin k = k1 + N1*k2.
struct CodeWinParams {
    const float2* src;
    float2* dst;
    i64 N;
    int N1, L;
};
struct KCodeWindow {
    typedef CodeWinParams Params;
    static const char* name() { return "code_window"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 1, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        const i64 e = (i64)blk.x * NT + tid;
        if (e >= p.N) return;
        const i64 k1 = e / p.L, k2 = e % p.L;
        const i64 k = k1 + (i64)p.N1 * k2;
        const double2 c = cis_cycles_d((double)k / (double)(p.N - 1));
        const float w = (float)(0.54 - 0.46 * c.x);
        const float2 x = p.src[e];
        p.dst[e] = make_float2(x.x * w, x.y * w);
    }
};

// ------------------------------------------------------------------------------------------------
// KBankFinal: per-bin metric of search_df / the Doppler bank (claudio_aligned_code_re_separate.m:33-42):
//   prnmap=abs(ifft(fcode.*conj(fft(y)))); [prnsig,b]=max(prnmap); prnmap(b-5:b+5)=0; snr=prnsig^2/var(prnmap)
// from the tile sums of |z|, |z|^2 (inverse column pass) and the 11 lags around the maximum (KPeakEvalT).
constexpr int BANK_NPTS = 11;
constexpr int BANK_DLO = -5;
struct BankRow { double b0, prnsig, metric, status; };
struct BankFinParams {
    const double* part_stat;   // [lane][npart][2]
    int npart;
    const double* lags;        // [lane][BANK_NPTS][2]  unnormalised correlation around the maximum
    const i64* p0;             // [lane]
    i64 N;
    BankRow* out;              // [lane]
};
struct KBankFinal {
    typedef BankFinParams Params;
    static const char* name() { return "bank_final"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 2, NSUM = 2;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_SUMD; }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char*)
    {
        const i64 lane = blk.y;
        if (PH == 0) {
            double s1 = 0.0, s2 = 0.0;
            for (int i = tid; i < p.npart; i += NT) {
                s1 += p.part_stat[(lane * p.npart + i) * 2];
                s2 += p.part_stat[(lane * p.npart + i) * 2 + 1];
            }
            r.dsum[0] = s1;
            r.dsum[1] = s2;
            return;
        }
        if (tid != 0) return;
        const double n = (double)p.N;
        double s1 = r.dsum[0], s2 = r.dsum[1], peak = 0.0;
        for (int d = 0; d < BANK_NPTS; d++) {           // the zeroed samples b-5..b+5 leave both sums
            const double re = p.lags[(lane * BANK_NPTS + d) * 2], im = p.lags[(lane * BANK_NPTS + d) * 2 + 1];
            const double a2 = re * re + im * im;
            s1 -= sqrt(a2);
            s2 -= a2;
            if (d == -BANK_DLO) peak = sqrt(a2);
        }
        s1 /= n;                                        // ifft scale 1/N on the magnitudes
        s2 /= n * n;
        peak /= n;
        const double mu = s1 / n;
        const double var = (s2 - n * mu * mu) / (n - 1.0);
        BankRow o;
        const i64 b0 = p.p0[lane];
        o.b0 = (double)b0;
        o.prnsig = peak;
        o.metric = peak * peak / var;
        o.status = (b0 < 5 || b0 > p.N - 6) ? 1.0 : 0.0;   // Octave's b-5:b+5 would leave the array
        p.out[lane] = o;
    }
};

// ------------------------------------------------------------------------------------------------
// KPrnGen: LFSR PRN replica (amaranth_twstft/common.py:23-30,59-73; prn.py:43-51).  Every thread
// jumps to chip t0 = gid*CH with the precomputed GF(2) powers M^(2^j) of the one-step matrix
// (columns as bit masks), then steps CH chips emitting the LSB before each step.
struct PrnParams {
    unsigned jump[24][24];   // jump[j][i] = column i of M^(2^j)
    int bitlen;
    unsigned taps, seed;
    i64 nchips;
    unsigned char* out;
    int ostride, ooffset;    // out[ostride*i + ooffset]  (QPSK: stride 2, offset 0/1)
};
struct KPrnGen {
    typedef PrnParams Params;
    static const char* name() { return "prn_gen"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 128, MINB = 4, NPH = 1, NSUM = 1, CH = 64;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    static HD unsigned parity(unsigned x)
    {
        x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
        return x & 1u;
    }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        const i64 t0 = ((i64)blk.x * NT + tid) * CH;
        if (t0 >= p.nchips) return;
        unsigned s = p.seed;
        for (int j = 0; j < 24; j++)
            if ((t0 >> j) & 1) {
                unsigned ns = 0;
                for (int i = 0; i < p.bitlen; i++)
                    if ((s >> i) & 1) ns ^= p.jump[j][i];
                s = ns;
            }
        const i64 tend = (t0 + CH < p.nchips) ? t0 + CH : p.nchips;
        for (i64 t = t0; t < tend; t++) {
            p.out[t * p.ostride + p.ooffset] = (unsigned char)(s & 1u);
            const unsigned bit = parity(s & p.taps);
            s = (s >> 1) | (bit << (p.bitlen - 1));
        }
    }
};

}  // namespace twstft
