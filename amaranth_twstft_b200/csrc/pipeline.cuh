// Launch sequence of the per-record correlator (one "lane" = one channel of one record/window).
//
//   KMean                       exact window means                          godual_ranging.m:80,94
//   KCol<RAW,FWD>(x^p) + KRow<FWD_ARGMAX> + KDfFinal   carrier offset       godual_ranging.m:14-15
//   KCol<RAW,FWD>(derotate) + KRowCorr                 fft(y).*fcode, row half of the x3 ifft  godual_ranging.m:17-28
//   KCol<CF32,INV_ARGMAX>                              column half of the x3 ifft + coarse max  godual_ranging.m:27-29
//   KPeakEval + KPeakFinal                             parabola, phase, SNR godual_ranging.m:30-48
//
// The same code drives the CUDA backend (product) and the host emulation backend (tests only).
#pragma once
#include <vector>
#include <math.h>
#include <stdlib.h>
#include "backend.cuh"
#include "k_row.cuh"
#include "k_decim.cuh"
#include "k_fast.cuh"
#include "k_col.cuh"
#include "k_rowcorr.cuh"
#include "k_small.cuh"
#include "k_synth.cuh"

namespace twstft {

struct EstimatorCfg {
    double f0, fstep, flast;   // frequency axis linspace(start, stop, naxis)
    i64 naxis;
    i64 win_lo, win_hi;        // inclusive window on the fftshift'ed axis
    int power;                 // 2 or 4
    int sub_mean;              // 1: square the mean-removed window (G), 0: raw (K)
};

struct RunArgs {
    const float2* d_cf = nullptr;   // complex-float windows instead of int16 records (one channel, mean already
                               // removed by the caller): rec_stride / d_rec_offset count samples
    const short2* d_raw;
    int nrec;
    int nch;
    i64 rec_stride;            // complex-int16 elements between records
    const i64* d_rec_offset;   // optional device array [nrec]
    int estimate;              // run the x^p carrier estimator
    const double* d_df_in;     // [lanes] when !estimate
    int variant;               // 0 G, 1 K
    int ddof;
    double fs;
    Result* d_out;             // [lanes]
    int allow_fast;            // the caller re-runs rows flagged status bit 2 through the full path (k_fast.cuh)
};

class IPipe {
public:
    virtual ~IPipe() {}
    virtual i64 size() const = 0;
    virtual int col_len() const = 0;
    virtual int row_len() const = 0;
    virtual size_t bytes_per_lane() const = 0;
    virtual int setup(Backend* be, int max_lanes, int nq) = 0;
    virtual int build_code(const unsigned char* d_bits, int rep, int qpsk, float2 mean, double code_pow) = 0;
    virtual void set_code_window(int w) = 0;
    virtual int set_fast_peak(int on) = 0;     // fast peak path (k_fast.cuh): 0 ok, <0 allocation failure
    // forward transform of `lanes` cf32 sequences of size() points, [lane][n] -> [lane][k1][k2] (bin k1 + col_len*k2);
    // in place allowed (d_out == d_in); lanes <= max_lanes
    virtual int spectrum_cf32(const float2* d_in, float2* d_out, int lanes) = 0;   // 1: Hamming-windowed code spectrum (C++ program); before build_code
    virtual int estimate(const RunArgs& a, const EstimatorCfg& e) = 0;   // -> d_df/d_sbin (+ derotation tables)
    virtual int estimate_block2(const RunArgs& a, const EstimatorCfg& e) = 0;
    virtual int fine_frequency(const RunArgs& a) = 0;   // A10: d_df += dfleftover, derotation tables rebuilt
    virtual int correlate(const RunArgs& a) = 0;
    virtual int window_spectrum(const RunArgs& a, const EstimatorCfg& e, float* d_win) = 0;
    virtual int window_spectrum_block2(const RunArgs& a, const EstimatorCfg& e, float* d_win) = 0;
    virtual int bank(const RunArgs& a, int sub_mean, BankRow* d_out) = 0;
    virtual int debug_fetch(int what, int lane, void* host_out, size_t bytes) = 0;
};

// TS: column-tile width of the forward (raw / code) column passes, TSI: of the inverse column pass
template <class CS, class RS, int RPC, int TS, int TSI = TS> class Pipe : public IPipe {
public:
    static constexpr i64 N = (i64)CS::L * RS::L;
    static constexpr int L1 = CS::L, N2 = RS::L;
    static_assert(N2 % TS == 0 && (2 * N2) % TS == 0 && N2 % TSI == 0, "column tile must divide the row length");
    static_assert(L1 % RPC == 0, "rows per CTA must divide the column length");
    static_assert(N2 % 2 == 0, "negative-frequency split needs an even row length");

    Backend* be = nullptr;
    int max_lanes = 0, nq = 3;
    bool have_means = false;   // estimate() already left the window means in d_sums
    float2 *d_blob_c = nullptr, *d_blob_r = nullptr, *d_code = nullptr, *d_qconst = nullptr;
    float2 *d_A = nullptr, *d_B = nullptr;
    double *d_sums = nullptr, *d_energy = nullptr, *d_df = nullptr, *d_peakpart = nullptr, *d_stat = nullptr, *d_bankpart = nullptr;
    u64 *d_keyrow = nullptr, *d_keycol = nullptr;
    int* d_counter = nullptr;  // work counter of the dynamically scheduled kernels
    i64 *d_sbin = nullptr, *d_p0 = nullptr;
    float2 *d_y0 = nullptr, *d_lrow = nullptr, *d_lcol = nullptr;
    float2 *d_lrow2 = nullptr, *d_lcol2 = nullptr;     // W_2N^n tables of the 2N-sample block estimator (block2_passes)
    static constexpr int NROWBLK = L1 / RPC;
    static constexpr int NCOLTILE = N2 / TS;
    static constexpr int NCOLTILE_I = N2 / TSI;

    ~Pipe() override
    {
        if (!be) return;
        void* all[] = {d_counter, d_blob_c, d_blob_r, d_code, d_qconst, d_A, d_B, d_sums, d_energy, d_df, d_peakpart,
                       d_keyrow, d_keycol, d_sbin, d_p0, d_y0, d_lrow, d_lcol, d_stat, d_bankpart, d_ang, d_finepart,
                       d_code_win, d_P, d_R2, d_U2, d_keycol2, d_fn, d_fbase, d_fa2, d_ffail, d_penergy, d_spart,
                       d_lagpart, d_hbtab, d_fwork, d_ftiles, d_fk1, d_lrow2, d_lcol2};
        for (void* p : all) be->free(p);
    }
    i64 size() const override { return N; }
    int col_len() const override { return L1; }
    int row_len() const override { return N2; }
    size_t bytes_per_lane() const override { return (size_t)N * sizeof(float2) * (1 + nq); }

    template <class T> T* dalloc(size_t n) { return reinterpret_cast<T*>(be->alloc(n * sizeof(T))); }

    // persistent kernels: one wave of resident CTAs, each looping over work items
    bool persistent = false;   // false: one CTA per work item; true: one resident wave looping over items
    // estimator row passes: a few resident waves, each CTA looping over ~a dozen rows and pulling its next row towards
    // L2 while it transforms the current one (measured: 8 waves -4 % against one CTA per row, one wave +0 %)
    template <class K> dim3 rgrid(i64 total) const
    {
        i64 g = (i64)be->num_sms * K::MINB * 8;
        if (persistent) g = (i64)be->num_sms * K::MINB;
        if (g > total) g = total;
        if (g < 1) g = 1;
        return dim3((unsigned)g, 1, 1);
    }
    // column passes (one CTA per tile): every CTA pulls the tile of the CTA that starts one resident wave later towards
    // L2.  Measured (tools/kbench, 16 records): inverse arg-max pass 0.382 -> 0.373 ms at one wave, worse at two or
    // three; forward raw pass unchanged at one wave, worse beyond -> on for the inverse passes only.
    int col_pf = -1;           // TWSTFT_COL_PF: waves for every column pass (experiments); -1: the defaults above
    template <class K> int pf_dist() const
    {
        const int waves = col_pf >= 0 ? col_pf : (K::INVERSE ? 1 : 0);
        return persistent ? 0 : waves * be->num_sms * K::MINB;
    }
    template <class K> dim3 pgrid(i64 total) const
    {
        i64 g = persistent ? (i64)be->num_sms * K::MINB : total;
        if (g > total) g = total;
        if (g < 1) g = 1;
        return dim3((unsigned)g, 1, 1);
    }

    int setup(Backend* b, int lanes, int nq_) override
    {
        be = b;
        persistent = false;
        if (const char* e = getenv("TWSTFT_PERSISTENT")) persistent = atoi(e) != 0;
        if (const char* e = getenv("TWSTFT_COL_PF")) col_pf = atoi(e);
        max_lanes = lanes;
        nq = nq_;
        std::vector<float2> blob(CS::BLOB + RS::BLOB + 4);
        d_blob_c = dalloc<float2>(CS::BLOB + 1);
        d_blob_r = dalloc<float2>(RS::BLOB + 1);
        fill_blob<CS>(blob.data());
        be->h2d(d_blob_c, blob.data(), CS::BLOB * sizeof(float2));
        be->sync();
        fill_blob<RS>(blob.data());
        be->h2d(d_blob_r, blob.data(), RS::BLOB * sizeof(float2));
        be->sync();
        // constant part of the polyphase twist exp(2 pi i q ktilde/(nq N)) for element m of a stage-0
        // butterfly (k_rowcorr.cuh): exp(2 pi i q m/(nq R0)), times exp(-2 pi i q/nq) on the negative side
        std::vector<float2> qc((size_t)nq * RS::R0);
        const double TWO_PI = 6.283185307179586476925287;
        for (int q = 0; q < nq; q++)
            for (int m = 0; m < RS::R0; m++) {
                double c = (double)(q * m) / ((double)nq * RS::R0) - (m >= RS::R0 / 2 ? (double)q / nq : 0.0);
                c -= floor(c);
                qc[(size_t)q * RS::R0 + m] = make_float2((float)cos(TWO_PI * c), (float)sin(TWO_PI * c));
            }
        d_qconst = dalloc<float2>(qc.size());
        be->h2d(d_qconst, qc.data(), qc.size() * sizeof(float2));
        be->sync();
        d_code = dalloc<float2>(N);
        d_A = dalloc<float2>((size_t)lanes * N);
        d_B = dalloc<float2>((size_t)lanes * nq * N);
        d_sums = dalloc<double>((size_t)lanes * 2);
        d_energy = dalloc<double>((size_t)lanes * NROWBLK);
        d_df = dalloc<double>(lanes);
        d_peakpart = dalloc<double>((size_t)lanes * 2 * PEAK_NPTS);
        d_stat = dalloc<double>((size_t)lanes * NCOLTILE_I * 2);
        d_bankpart = dalloc<double>((size_t)lanes * 2 * BANK_NPTS);
        d_keyrow = dalloc<u64>((size_t)lanes * NROWBLK * 2);
        d_counter = dalloc<int>(4);
        if (d_counter) be->zero(d_counter, 4 * sizeof(int));
        d_keycol = dalloc<u64>((size_t)lanes * nq * NCOLTILE_I);
        d_sbin = dalloc<i64>(lanes);
        d_p0 = dalloc<i64>(lanes);
        d_y0 = dalloc<float2>(lanes);
        d_lrow = dalloc<float2>((size_t)lanes * L1);
        d_lcol = dalloc<float2>((size_t)lanes * N2);
        return be->ok() ? 0 : -1;
    }

    // ---- fast peak path (k_fast.cuh): one inverse transform + exact local lags + a bound on everything else
    int fast_peak = 0;
    double fast_H = 0.0, fast_T = 0.0;
    float2 *d_P = nullptr, *d_R2 = nullptr, *d_U2 = nullptr;
    u64* d_keycol2 = nullptr;
    i64 *d_fn = nullptr, *d_fbase = nullptr;
    double *d_fa2 = nullptr, *d_penergy = nullptr, *d_spart = nullptr, *d_lagpart = nullptr, *d_hbtab = nullptr, *d_fwork = nullptr;
    int *d_ffail = nullptr, *d_ftiles = nullptr;
    u64* d_fk1 = nullptr;
    static double fast_hb(double t, double n)
    {
        const double s = fabs(sin(3.14159265358979323846 * t / n));
        return 0.86602540378443865 / (n * s);
    }
    int set_fast_peak(int on) override
    {
        if (!on || nq != 3) { fast_peak = 0; return 0; }
        if (!d_P) {
            const double n = (double)N;
            const i64 K = N / 4 < 100000 ? N / 4 : 100000;
            for (int q = 1; q <= 2; q++) {
                double h = 0.0, t2 = 0.0;
                for (i64 d = -(N / 2); d < N - N / 2; d++) {
                    const double v = fast_hb(fabs((double)d + q / 3.0), n);
                    if (d >= -K && d <= K) h += v; else t2 += v * v;
                }
                if (h > fast_H) fast_H = h;
                if (sqrt(t2) > fast_T) fast_T = sqrt(t2);
            }
            std::vector<float2> u2(N2);
            for (int k2 = 0; k2 < N2; k2++) {
                const double c = ((double)L1 * k2 - (k2 >= N2 / 2 ? n : 0.0)) / (3.0 * n);
                const double a = 6.283185307179586476925287 * (c - floor(c));
                u2[k2] = make_float2((float)cos(a), (float)sin(a));
            }
            const int ns = 3 * TSI, ntab = 2 * ns + 5;
            std::vector<double> hbt(2 * (size_t)ntab);
            for (int q = 1; q <= 2; q++)
                for (int dd = -(ns + 2); dd <= ns + 2; dd++)
                    hbt[(size_t)(q - 1) * ntab + dd + ns + 2] = fast_hb(fabs((double)dd + q / 3.0), n);
            d_hbtab = dalloc<double>(hbt.size());
            d_fwork = dalloc<double>((size_t)max_lanes * 8);
            d_P = dalloc<float2>((size_t)max_lanes * N);
            d_R2 = dalloc<float2>((size_t)max_lanes * N2);
            d_U2 = dalloc<float2>(N2);
            d_keycol2 = dalloc<u64>((size_t)max_lanes * 3);
            d_fk1 = dalloc<u64>((size_t)max_lanes * 3);
            d_ftiles = dalloc<int>((size_t)max_lanes * 3);
            d_fn = dalloc<i64>(max_lanes);
            d_fbase = dalloc<i64>(max_lanes);
            d_fa2 = dalloc<double>(max_lanes);
            d_ffail = dalloc<int>(max_lanes);
            d_penergy = dalloc<double>((size_t)max_lanes * NROWBLK);
            d_spart = dalloc<double>((size_t)max_lanes * 3 * TSI * 2);
            d_lagpart = dalloc<double>((size_t)max_lanes * L1 * 2 * FAST_NLAG);
            if (!be->ok()) return -1;
            be->h2d(d_U2, u2.data(), sizeof(float2) * N2);
            be->h2d(d_hbtab, hbt.data(), sizeof(double) * hbt.size());
            be->sync();
        }
        fast_peak = 1;
        return be->ok() ? 0 : -1;
    }

    int code_window = 0;       // 1: correlate against hamming(N) .* conj(fft(code)) (processing/CPP/main.cpp:717-719)
    float2* d_code_win = nullptr;
    void set_code_window(int w) override { code_window = w; }
    double code_pow = 1.0;     // mean |code|^2: 1 for the +/-1 BPSK replica, ~0.5 for the QPSK one
    int build_code(const unsigned char* d_bits, int rep, int qpsk, float2 mean, double cpow) override
    {
        code_pow = cpow;
        ColParams c{};
        c.src = d_bits; c.dst = d_A; c.blob = d_blob_c; c.N = N; c.N2 = N2; c.invN = 1.0 / (double)N; c.nch = 1;
        c.rep = rep; c.qpsk = qpsk; c.code_mean = mean; c.nq = 1; c.ntx = NCOLTILE; c.ny = 1;
        typedef KCol<CS, TS, SRC_BITS, COL_FWD, 0, N2, 1> KC;
        c.pf_dist = pf_dist<KC>();
        be->template launch<KC>(pgrid<KC>(NCOLTILE), c);
        RowParams r{};
        r.src = d_A; r.dst = d_code; r.blob = d_blob_r; r.N = N; r.N1 = L1; r.invN = 1.0 / (double)N; r.invM = 1.0 / ((double)nq * (double)N); r.nblk = NROWBLK;
        r.nlanes = 1;
        typedef KRow<RS, RPC, ROW_FWD_STORE> KR;
        be->template launch<KR>(pgrid<KR>(NROWBLK), r);
        if (code_window) {
            if (!d_code_win) d_code_win = dalloc<float2>(N);
            CodeWinParams w{};
            w.src = d_code; w.dst = d_code_win; w.N = N; w.N1 = L1; w.L = N2;
            be->template launch<KCodeWindow>(dim3((unsigned)((N + KCodeWindow::NT - 1) / KCodeWindow::NT), 1, 1), w);
        }
        return be->ok() ? 0 : -1;
    }

    int spectrum_cf32(const float2* d_in, float2* d_out, int lanes) override
    {
        if (lanes > max_lanes) return -2;
        ColParams c{};
        c.src = d_in; c.dst = d_A; c.blob = d_blob_c; c.N = N; c.N2 = N2; c.invN = 1.0 / (double)N; c.nch = 1;
        c.nq = 1; c.ntx = NCOLTILE; c.ny = lanes;
        typedef KCol<CS, TS, SRC_CF32, COL_FWD, 0, N2, 1> KC;
        c.pf_dist = pf_dist<KC>();
        be->template launch<KC>(pgrid<KC>((i64)NCOLTILE * lanes), c);
        RowParams r{};
        r.src = d_A; r.dst = d_out; r.blob = d_blob_r; r.N = N; r.N1 = L1; r.invN = 1.0 / (double)N;
        r.invM = 1.0 / ((double)nq * (double)N); r.nblk = NROWBLK; r.nlanes = lanes;
        typedef KRow<RS, RPC, ROW_FWD_STORE> KR;
        be->template launch<KR>(pgrid<KR>((i64)NROWBLK * lanes), r);
        return be->ok() ? 0 : -1;
    }

    void means(const RunArgs& a)
    {
        const int lanes = a.nrec * a.nch;
        be->zero(d_sums, (size_t)lanes * 2 * sizeof(double));
        if (a.d_cf) return;        // complex-float windows arrive mean-removed (godual_ranging.m:80,94): sums stay 0
        MeanParams m{};
        m.src = a.d_raw; m.nelem = N * a.nch; m.nch = a.nch; m.rec_stride = a.rec_stride;
        m.rec_offset = a.d_rec_offset; m.sums = d_sums;
        i64 per = (m.nelem + KMean::NT - 1) / KMean::NT;
        // two resident waves (148 SMs x 4 CTAs) over the whole batch: CTAs live long enough to hide their launch
        const i64 cap = (2 * 592 + a.nrec - 1) / a.nrec;
        int nblk = (int)(per < cap ? (per < 1 ? 1 : per) : cap);
        m.nblk = nblk;
        be->template launch<KMean>(dim3(nblk, a.nrec, 1), m);
    }

    template <int OP> void col_raw(const RunArgs& a, int sub_mean, i64 half_stride = 0)
    {
        ColParams c{};
        c.half_stride = half_stride;
        c.src = a.d_raw; c.dst = d_A; c.blob = d_blob_c; c.N = N; c.N2 = N2; c.invN = 1.0 / (double)N; c.nch = a.nch;
        c.sub_mean = sub_mean; c.sums = d_sums; c.lrow = d_lrow; c.lcol = d_lcol;
        if (OP >= 10) { c.lrow = d_lrow2; c.lcol = d_lcol2; }        // radix-2 split of a 2N block: the plan's W_2N^n tables
        c.rec_stride = a.rec_stride; c.rec_offset = a.d_rec_offset; c.nq = 1;
        c.ntx = NCOLTILE * a.nch; c.ny = a.nrec;
        if (a.d_cf) {
            c.src = a.d_cf; c.sub_mean = 0; c.nch = 1; c.ntx = NCOLTILE;
            typedef KCol<CS, TS, SRC_CF32, COL_FWD, OP, N2, 1> KC;
            c.pf_dist = pf_dist<KC>();
            be->template launch<KC>(pgrid<KC>((i64)c.ntx * c.ny), c);
        } else if (a.nch == 2) {
            typedef KCol<CS, TS, SRC_RAW, COL_FWD, OP, N2, 2> KC;
            c.pf_dist = pf_dist<KC>();
            be->template launch<KC>(pgrid<KC>((i64)c.ntx * c.ny), c);
        } else {
            typedef KCol<CS, TS, SRC_RAW, COL_FWD, OP, N2, 1> KC;
            c.pf_dist = pf_dist<KC>();
            be->template launch<KC>(pgrid<KC>((i64)c.ntx * c.ny), c);
        }
    }

    void df_final(const RunArgs& a, const EstimatorCfg* e, int npart = NROWBLK, bool keep_sbin = false)
    {
        const int lanes = a.nrec * a.nch;
        DfParams d{};
        d.part_key = d_keyrow; d.npart = npart; d.key_stride = 2 * NROWBLK; d.df_in = e ? nullptr : a.d_df_in;
        if (e) { d.f0 = e->f0; d.fstep = e->fstep; d.flast = e->flast; d.naxis = e->naxis; d.power = (double)e->power; }
        else { d.power = 1.0; d.naxis = 0; }
        d.fs = a.fs; d.L1 = L1; d.N2 = N2; d.df_out = d_df; d.s_out = keep_sbin ? nullptr : d_sbin; d.lrow = d_lrow; d.lcol = d_lcol;
        d.nsplit = 8;
        be->template launch<KDfFinal>(dim3(d.nsplit, lanes, 1), d);
    }

    // row-pass parameters of the estimator window; (bmul, badd, nbig) map this transform's bins onto the
    // axis the window lives on (identity, or one half of the radix-2 split of a 2N-sample block)
    RowParams estimator_row(const RunArgs& a, const EstimatorCfg& e, int bmul, i64 badd, i64 nbig)
    {
        const int lanes = a.nrec * a.nch;
        RowParams r{};
        r.src = d_A; r.blob = d_blob_r; r.N = N; r.N1 = L1; r.invN = 1.0 / (double)N; r.invM = 1.0 / ((double)nq * (double)N); r.win_lo = e.win_lo; r.win_hi = e.win_hi;
        r.part_key = d_keyrow; r.key_stride = 2 * NROWBLK; r.key_off = 0; r.nblk = NROWBLK;
        r.bmul = bmul; r.badd = badd; r.nbig = nbig;
        {   // k2 ranges whose bins k = k1 + L1*k2 can intersect the window (bins as signed frequencies)
            auto fdiv = [](i64 x, i64 d) { i64 q = x / d; if ((x % d) && ((x < 0) != (d < 0))) q--; return q; };
            // signed range of this transform's bins: bmul*k + badd in [win_lo - nbig/2, win_hi - nbig/2]
            const i64 ka = fdiv(e.win_lo - nbig / 2 - badd, bmul) - 1, kb = fdiv(e.win_hi - nbig / 2 - badd, bmul) + 1;
            auto lo_of = [&](i64 k) { return fdiv(k - (L1 - 1), L1); };
            r.oa1 = 0; r.ob1 = -1; r.oa2 = 0; r.ob2 = -1;
            if (ka >= 0) { r.oa1 = (int)lo_of(ka); r.ob1 = (int)(kb / L1); }
            else if (kb < 0) { r.oa1 = (int)lo_of(ka + N); r.ob1 = (int)((kb + N) / L1); }
            else { r.oa1 = 0; r.ob1 = (int)(kb / L1); r.oa2 = (int)lo_of(ka + N); r.ob2 = N2 - 1; }
            if (r.oa1 < 0) r.oa1 = 0;
            if (r.oa2 < 0) r.oa2 = 0;
            if (r.ob1 > N2 - 1) r.ob1 = N2 - 1;
        }
        r.nlanes = lanes;
        // output-pruned row pass (k_row.cuh PRUNABLE): which middle-stage outputs k1 feed a bin of those k2 ranges
        r.k1mask = 0u;
        if (RS::NSTAGE == 3 && RS::R0 == 20 && RS::R1 == 20 && RS::R2 == 20 && RPC == 1 && !getenv("TWSTFT_NO_PRUNE")) {
            unsigned mask = 0u;
            long long nout = 0;
            const int lo[2] = {r.oa1, r.oa2}, hi[2] = {r.ob1, r.ob2};
            for (int g = 0; g < 2; g++)
                for (int o = lo[g]; o <= hi[g]; o++) { mask |= 1u << ((o % RS::NSLAST) / RS::R0); nout++; }
            int bits = 0;
            for (int b = 0; b < 20; b++) bits += (mask >> b) & 1u;
            if (bits > 0 && bits <= 8 && nout <= 2 * RS::NTS) r.k1mask = mask;       // wide windows: the full transform
        }
        return r;
    }

    // forward column transform of x^p and the row-pass parameters of the estimator window
    RowParams estimator_front(const RunArgs& a, const EstimatorCfg& e)
    {
        if (e.sub_mean) means(a);
        have_means = e.sub_mean != 0;
        if (e.power == 4) col_raw<4>(a, e.sub_mean);
        else col_raw<2>(a, e.sub_mean);
        return estimator_row(a, e, 1, 0, N);
    }

    // Carrier estimator on blocks of 2N samples (the 2-s blocks of claudio_aligned_code_lo_separate.m:120-129):
    // first radix-2 split fused into the column-pass loads, two N-point transforms, one arg-max over both.
    // two N-point transforms of the even/odd radix-2 halves of every 2N-sample block, row pass in MODE
    template <int MODE> int block2_passes(const RunArgs& a, const EstimatorCfg& e, float* d_win)
    {
        const int lanes = a.nrec * a.nch;
        if (lanes > max_lanes) return -2;
        have_means = false;
        if (!d_lrow2) {   // W_2N^n = exp(-2 pi i n/(2N)) as a "derotation" by fs/(2N) through the two-level tables: the same
            // for every lane and every call, built once per plan (it used to cost a blocking copy and a launch per block)
            d_lrow2 = dalloc<float2>((size_t)max_lanes * L1);
            d_lcol2 = dalloc<float2>((size_t)max_lanes * N2);
            if (!d_lrow2 || !d_lcol2 || !be->ok()) return -1;
            std::vector<double> half(max_lanes, a.fs / (2.0 * (double)N));
            be->h2d(d_df, half.data(), sizeof(double) * max_lanes);
            be->sync();
            DfParams d{};
            d.part_key = d_keyrow; d.npart = NROWBLK; d.key_stride = 2 * NROWBLK; d.df_in = d_df; d.power = 1.0; d.naxis = 0;
            d.fs = a.fs; d.L1 = L1; d.N2 = N2; d.df_out = d_df; d.s_out = d_sbin; d.lrow = d_lrow2; d.lcol = d_lcol2;
            d.nsplit = 8;
            be->template launch<KDfFinal>(dim3(d.nsplit, max_lanes, 1), d);
        }
        typedef KRow<RS, RPC, MODE> KR;
        if (MODE == ROW_FWD_ARGMAX) be->zero(d_keyrow, (size_t)max_lanes * NROWBLK * 2 * sizeof(u64));   // atomic maxima land here
        for (int par = 0; par < 2; par++) {
            if (e.power == 4) { if (par) col_raw<24>(a, 0, N * a.nch); else col_raw<14>(a, 0, N * a.nch); }
            else { if (par) col_raw<22>(a, 0, N * a.nch); else col_raw<12>(a, 0, N * a.nch); }
            RowParams r = estimator_row(a, e, 2, par, 2 * N);
            r.key_off = par * NROWBLK;
            r.winmag = d_win;
            be->template launch<KR>(rgrid<KR>((i64)NROWBLK * lanes), r);
        }
        return 0;
    }
    int estimate_block2(const RunArgs& a, const EstimatorCfg& e) override
    {
        const int rc = block2_passes<ROW_FWD_ARGMAX>(a, e, nullptr);
        if (rc != 0) return rc;
        df_final(a, &e, 2 * NROWBLK);
        return be->ok() ? 0 : -1;
    }
    // |fft(x^p)|^2 of every window bin of the 2N-point block spectrum -> d_win[lane][win_hi-win_lo+1]
    // (search_df runs on the 2-s block: claudio_aligned_code_re_separate.m:28-29,150)
    int window_spectrum_block2(const RunArgs& a, const EstimatorCfg& e, float* d_win) override
    {
        const int rc = block2_passes<ROW_FWD_WINMAG>(a, e, d_win);
        if (rc != 0) return rc;
        return be->ok() ? 0 : -1;
    }

    int estimate(const RunArgs& a, const EstimatorCfg& e) override
    {
        const int lanes = a.nrec * a.nch;
        if (lanes > max_lanes) return -2;
        RowParams r = estimator_front(a, e);
        typedef KRow<RS, RPC, ROW_FWD_ARGMAX> KR;
        be->zero(d_keyrow, (size_t)lanes * NROWBLK * 2 * sizeof(u64));       // the rows' atomic maxima land here
        be->template launch<KR>(rgrid<KR>((i64)NROWBLK * lanes), r);
        df_final(a, &e);
        return be->ok() ? 0 : -1;
    }

    // Fine-frequency refinement after estimate(): needs the window means (left in d_sums by estimate() when the
    // estimator removes them, computed here otherwise) and the coarse df in d_df.
    double* d_ang = nullptr;
    double* d_finepart = nullptr;
    int fine_M = 0;
    static constexpr int FINE_SPLIT = 8;
    int fine_frequency(const RunArgs& a) override
    {
        const int lanes = a.nrec * a.nch;
        if (lanes > max_lanes) return -2;
        if (a.d_cf) return -4;                                     // the angle kernel reads the int16 records
        const i64 third = (i64)floor(a.fs) / 3;                    // fs//3
        const int step = 10;
        const i64 My = ((third < N ? third : N) + step - 1) / step;   // len(y[0:fs//3:10])
        const i64 Mt = (third - 1 + step - 1) / step;                 // len(arange(1, fs//3, 10))
        if (My != Mt || My < 2) return -3;                            // numpy raises in the reference
        if (!d_ang || fine_M != (int)My) {
            be->free(d_ang); be->free(d_finepart);
            d_ang = dalloc<double>((size_t)max_lanes * My);
            d_finepart = dalloc<double>((size_t)max_lanes * FINE_SPLIT);
            fine_M = (int)My;
        }
        if (!have_means) { means(a); have_means = true; }
        FineParams f{};
        f.src = a.d_raw; f.nch = a.nch; f.rec_stride = a.rec_stride; f.rec_offset = a.d_rec_offset; f.sums = d_sums;
        f.invN = 1.0 / (double)N; f.fs = a.fs; f.M = (int)My; f.step = step; f.taps = 100; f.ang = d_ang;
        f.part = d_finepart; f.nsplit = FINE_SPLIT; f.df = d_df;
        be->template launch<KFineAngle>(dim3((unsigned)((My + KFineAngle::NT - 1) / KFineAngle::NT), lanes, 1), f);
        be->template launch<KFineFit>(dim3(FINE_SPLIT, lanes, 1), f);
        be->template launch<KFineFinal>(dim3(1, lanes, 1), f);
        RunArgs t = a;
        t.d_df_in = d_df;
        df_final(t, nullptr, NROWBLK, true);                          // tables for df + dfleftover, carrier bin kept
        return be->ok() ? 0 : -1;
    }

    // |fft(x^p)|^2 of every bin of the estimator window -> d_win[lane][win_hi-win_lo+1]  (search_df, A4)
    int window_spectrum(const RunArgs& a, const EstimatorCfg& e, float* d_win) override
    {
        const int lanes = a.nrec * a.nch;
        if (lanes > max_lanes) return -2;
        RowParams r = estimator_front(a, e);
        have_means = false;
        r.winmag = d_win;
        typedef KRow<RS, RPC, ROW_FWD_WINMAG> KR;
        be->template launch<KR>(rgrid<KR>((i64)NROWBLK * lanes), r);
        return be->ok() ? 0 : -1;
    }

    // 1x correlation of every lane against the code, conj(Y)*C as in search_df, and the per-lane metric
    // max^2/var with b-5..b+5 zeroed (claudio_aligned_code_re_separate.m:33-42) -> d_out[lane]
    int bank(const RunArgs& a, int sub_mean, BankRow* d_out) override
    {
        const int lanes = a.nrec * a.nch;
        if (lanes > max_lanes) return -2;
        if (sub_mean) means(a);
        have_means = false;
        df_final(a, nullptr);
        col_raw<1>(a, sub_mean);
        RowCorrParams r{};
        r.src = d_A; r.dst = d_B; r.blob = d_blob_r; r.code = d_code; r.qconst = d_qconst; r.N = N; r.N1 = L1;
        r.invN = 1.0 / (double)N; r.invM = 1.0 / (double)N; r.nlanes = lanes; r.nblk = NROWBLK;
        r.conj_mode = 1; r.part_energy = d_energy; r.y0 = d_y0;
        typedef KRowCorr<RS, RPC, 1> K1;
        i64 g = (i64)NROWBLK * lanes;
        if (g > (i64)be->num_sms * K1::MINB) g = (i64)be->num_sms * K1::MINB;
        r.counter = d_counter;
        be->template launch<K1>(dim3((unsigned)g, 1, 1), r);
        ColParams c{};
        c.src = d_B; c.blob = d_blob_c; c.N = N; c.N2 = N2; c.invN = 1.0 / (double)N; c.nch = 1; c.nq = 1;
        c.part_key = d_keycol; c.part_stat = d_stat; c.ntx = NCOLTILE_I; c.ny = lanes;
        typedef KCol<CS, TSI, SRC_CF32, COL_INV_ARGMAX_STATS, 0, N2, 1> KCS;
        c.pf_dist = pf_dist<KCS>();
        be->template launch<KCS>(pgrid<KCS>((i64)c.ntx * c.ny), c);
        PeakEvalParams pe{};
        pe.part_key = d_keycol; pe.npart = NCOLTILE_I; pe.B = d_B; pe.N = N; pe.nq = 1; pe.L1 = L1; pe.N2 = N2;
        pe.p0_out = d_p0; pe.part = d_bankpart;
        be->template launch<KPeakEvalT<BANK_NPTS, BANK_DLO>>(dim3(BANK_NPTS, lanes, 1), pe);
        BankFinParams bf{};
        bf.part_stat = d_stat; bf.npart = NCOLTILE_I; bf.lags = d_bankpart; bf.p0 = d_p0; bf.N = N; bf.out = d_out;
        be->template launch<KBankFinal>(dim3(1, lanes, 1), bf);
        return be->ok() ? 0 : -1;
    }

    // correlate windows with df taken from d_df (left there by estimate()) or from a.d_df_in
    int correlate(const RunArgs& a) override
    {
        const int lanes = a.nrec * a.nch;
        if (lanes > max_lanes) return -2;
        if (!(a.estimate && have_means)) means(a);
        have_means = false;
        if (!a.estimate) df_final(a, nullptr);
        col_raw<1>(a, 1);
        RowCorrParams r{};
        r.src = d_A; r.dst = d_B; r.blob = d_blob_r; r.code = d_code; r.qconst = d_qconst; r.N = N; r.N1 = L1;
        r.invN = 1.0 / (double)N; r.invM = 1.0 / ((double)nq * (double)N); r.nlanes = lanes; r.nblk = NROWBLK;
        r.conj_mode = a.variant; r.part_energy = d_energy; r.y0 = d_y0;
        if (code_window && d_code_win) r.code = d_code_win;
        if (fast_peak && a.allow_fast && nq == 3 && !(code_window && d_code_win)) return correlate_fast(a, r);
        auto row_corr = [&]() {
            i64 g = (i64)NROWBLK * lanes;
            typedef KRowCorr<RS, RPC, 3> K3;
            typedef KRowCorr<RS, RPC, 1> K1;
            const i64 wave = (i64)be->num_sms * K3::MINB;      // resident CTAs: TMEM columns are per CTA
            if (g > wave) g = wave;
            r.counter = d_counter;
            if (nq == 3) be->template launch<K3>(dim3((unsigned)g, 1, 1), r);
            else be->template launch<K1>(dim3((unsigned)g, 1, 1), r);
        };
        row_corr();
        ColParams c{};
        c.src = d_B; c.blob = d_blob_c; c.N = N; c.N2 = N2; c.invN = 1.0 / (double)N; c.nch = 1; c.nq = nq; c.part_key = d_keycol;
        c.ntx = NCOLTILE_I; c.ny = lanes * nq;
        typedef KCol<CS, TSI, SRC_CF32, COL_INV_ARGMAX, 0, N2, 1> KCI;
        c.pf_dist = pf_dist<KCI>();
        be->template launch<KCI>(pgrid<KCI>((i64)c.ntx * c.ny), c);
        PeakEvalParams pe{};
        pe.part_key = d_keycol; pe.npart = nq * NCOLTILE_I; pe.B = d_B; pe.N = N; pe.nq = nq; pe.L1 = L1; pe.N2 = N2;
        pe.p0_out = d_p0; pe.part = d_peakpart;
        be->template launch<KPeakEval>(dim3(PEAK_NPTS, lanes, 1), pe);
        PeakFinParams pf{};
        pf.part = d_peakpart; pf.nblk = 1; pf.part_energy = d_energy; pf.nen = NROWBLK; pf.y0 = d_y0;
        pf.code_pow = code_pow; pf.p0 = d_p0; pf.df = d_df; pf.sbin = d_sbin; pf.N = N; pf.nq = nq; pf.variant = a.variant; pf.ddof = a.ddof;
        pf.out = a.d_out;
        pf.pc_out = (code_window && d_code_win) ? d_p0 : nullptr;
        be->template launch<KPeakFinal>(dim3(1, lanes, 1), pf);
        if (code_window && d_code_win) {
            // C++ flavour: the SNR despreads yint with the PLAIN code (main.cpp:319-357), i.e. it needs the lags of the
            // unwindowed correlation around the peak just decided: second row pass against d_code, no column pass
            r.code = d_code;
            row_corr();
            pe.p0_in = d_p0;
            be->template launch<KPeakEval>(dim3(PEAK_NPTS, lanes, 1), pe);
            pf.snr_only = 1; pf.pc_out = nullptr;
            be->template launch<KPeakFinal>(dim3(1, lanes, 1), pf);
        }
        return be->ok() ? 0 : -1;
    }

    // fast peak path (k_fast.cuh); `r` prepared by correlate()
    int correlate_fast(const RunArgs& a, RowCorrParams r)
    {
        const int lanes = a.nrec * a.nch;
        typedef KRowCorr<RS, RPC, 1> K1;
        r.pdst = d_P; r.part_penergy = d_penergy;
        i64 g = (i64)NROWBLK * lanes;
        if (g > (i64)be->num_sms * K1::MINB) g = (i64)be->num_sms * K1::MINB;
        r.counter = d_counter;
        be->template launch<K1>(dim3((unsigned)g, 1, 1), r);
        ColParams c{};
        c.src = d_B; c.blob = d_blob_c; c.N = N; c.N2 = N2; c.invN = 1.0 / (double)N; c.nch = 1; c.nq = 1;
        c.part_key = d_keycol; c.ntx = NCOLTILE_I; c.ny = lanes;
        typedef KCol<CS, TSI, SRC_CF32, COL_INV_ARGMAX, 0, N2, 1> KCI;
        c.pf_dist = pf_dist<KCI>();
        be->template launch<KCI>(pgrid<KCI>((i64)c.ntx * c.ny), c);
        FastPrepParams fp{};
        fp.key1 = d_keycol; fp.key2 = d_keycol2; fp.ntile = NCOLTILE_I; fp.tsi = TSI; fp.N2 = N2; fp.nrow = L1;
        fp.nstar = d_fn; fp.base = d_fbase; fp.tiles = d_ftiles; fp.a2 = d_fa2; fp.fail = d_ffail;
        be->template launch<KFastPrep>(dim3(1, lanes, 1), fp);
        // second maxima (rows other than the tile maximum's) of the three column tiles of S only
        ColParams c2 = c;
        c2.part_key = d_fk1; c2.part_key2 = d_keycol2; c2.ntx = 3; c2.tile_list = d_ftiles;
        typedef KCol<CS, TSI, SRC_CF32, COL_INV_ARGMAX_K2, 0, N2, 1> KC2;
        be->template launch<KC2>(pgrid<KC2>((i64)c2.ntx * c2.ny), c2);
        be->template launch<KFastA2>(dim3(1, lanes, 1), fp);
        PeakEvalParams pe{};
        pe.part_key = d_keycol; pe.npart = 0; pe.B = d_B; pe.N = N; pe.nq = 1; pe.L1 = L1; pe.N2 = N2;
        pe.p0_out = d_p0; pe.part = d_spart; pe.p0_in = d_fbase; pe.npts_rt = 3 * TSI; pe.dlo_rt = 0;
        be->template launch<KPeakEval>(dim3(3 * TSI, lanes, 1), pe);
        LagTabParams lt{};
        lt.nstar = d_fn; lt.N2 = N2; lt.R2 = d_R2;
        be->template launch<KLagTables>(dim3((N2 + KLagTables::NT - 1) / KLagTables::NT, lanes, 1), lt);
        LagPParams lp{};
        lp.P = d_P; lp.R2 = d_R2; lp.U2 = d_U2; lp.nstar = d_fn; lp.N = N; lp.L1 = L1; lp.N2 = N2; lp.part = d_lagpart;
        be->template launch<KLagP>(dim3(L1, lanes, 1), lp);
        FastCertParams fc{};
        fc.lagpart = d_lagpart; fc.L1 = L1; fc.spart = d_spart; fc.ns = 3 * TSI; fc.nstar = d_fn; fc.base = d_fbase;
        fc.a2 = d_fa2; fc.penergy = d_penergy; fc.nen = NROWBLK; fc.N = N; fc.H = fast_H; fc.T = fast_T;
        fc.fail = d_ffail; fc.p0 = d_p0; fc.part7 = d_peakpart; fc.hbtab = d_hbtab; fc.work = d_fwork;
        be->template launch<KFastCert>(dim3(1, lanes, 1), fc);
        PeakFinParams pf{};
        pf.part = d_peakpart; pf.nblk = 1; pf.part_energy = d_energy; pf.nen = NROWBLK; pf.y0 = d_y0;
        pf.code_pow = code_pow; pf.p0 = d_p0; pf.df = d_df; pf.sbin = d_sbin; pf.N = N; pf.nq = nq; pf.variant = a.variant;
        pf.ddof = a.ddof; pf.out = a.d_out; pf.fail = d_ffail;
        be->template launch<KPeakFinal>(dim3(1, lanes, 1), pf);
        return be->ok() ? 0 : -1;
    }

    // test hooks: copy an intermediate of one lane to the host
    int debug_fetch(int what, int lane, void* host_out, size_t bytes) override
    {
        const void* src = nullptr;
        switch (what) {
        case 0: src = d_code; break;
        case 1: src = d_A + (size_t)lane * N; break;
        case 3: src = d_B + (size_t)lane * nq * N; break;
        case 4: src = d_sums + (size_t)lane * 2; break;
        case 5: src = d_df + lane; break;
        case 6: src = d_p0 + lane; break;
        case 7: src = d_sbin + lane; break;
        default: return -1;
        }
        be->d2h(host_out, src, bytes);
        be->sync();
        return be->ok() ? 0 : -1;
    }
};

// Supported window lengths N = (column length) x (row length).
inline IPipe* make_pipe(i64 n)
{
    typedef FFTSpec<20, 20, 20> R8000;
    typedef FFTSpec<20, 2, 20> R800;       // first radix == last radix, as the fused correlator row pass needs
    typedef FFTSpec<20, 20> R400;
    switch (n) {
#ifndef TWSTFT_TS_FWD
#define TWSTFT_TS_FWD 16
#endif
    case 5000000: return new Pipe<FFTSpec<25, 25>, R8000, 1, TWSTFT_TS_FWD, 8>();    // 1-s window, 2.5 Mchip/s code
#ifndef TWSTFT_ONLY_5E6   /* kernel-variant builds of tools/kbench compile the 1-s window only */
    case 1000000: return new Pipe<FFTSpec<5, 25>, R8000, 1, 16>();     // 500 kchip codes
    case 200000: return new Pipe<FFTSpec<25>, R8000, 1, 64>();         // 100 kchip/s code windows
    case 250000: return new Pipe<FFTSpec<25, 25>, R400, 5, 16>();      // test size (same column kernel as 5e6)
    case 10000: return new Pipe<FFTSpec<25>, R400, 5, 16>();           // 5 kchip codes / test size
    // the other code lengths shipped under experiments/221207_twoway_codes/codes (x2 samples per chip)
    case 500000: return new Pipe<FFTSpec<25, 25>, R800, 5, 16, 8>();   // 250 kchip codes (noiselen250000_bitlen18_*)
    case 50000: return new Pipe<FFTSpec<5, 25>, R400, 5, 16>();        // 25 kchip codes  (noiselen25000_bitlen17_*)
    case 20000: return new Pipe<FFTSpec<25>, R800, 5, 16>();           // 10 kchip codes  (noiselen10000_bitlen14_*)
#endif
    default: return nullptr;
    }
}

}  // namespace twstft
