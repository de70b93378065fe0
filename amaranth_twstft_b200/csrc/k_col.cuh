// Column pass of the four-step FFT: length-L transforms (L = S::L, e.g. 625) over the slow index of
// an [L][ncols] matrix, TS adjacent columns per CTA so that every global access is a run of TS
// consecutive elements.  Fuses the IQ reader/unpacker of the reference into the forward transform:
//   processing/Octave/godual_ranging.m:76-80,93-94 (int16 -> complex, channel split, mean removal),
//   :14 (d.^2), :17-18 (derotation by exp(-j2pi df t)), and the code replica build (:62-66).
// and the peak search into the inverse one (:29  [~,indice]=max(abs(prnmap))).
//
// Forward (COL_FWD): x[n1*N2 + n2] -> A[k1][n2] * W_N^{n2 k1}   (four-step twiddle; for the derotated
//   path the per-column factor exp(-j2pi df n2/fs) commutes with the column DFT and is folded into it).
// Inverse (COL_INV_*): B[k1][n2] -> z[n1*N2 + n2], either reduced to the first-max of |z|^2 with index
//   p = nq*(n1*N2+n2) + q (polyphase q of the x3 interpolated correlation) or stored.
//
// Work items (column tile, matrix) are walked by whatever grid is launched (one CTA per item or a
// resident wave looping over items), like the row pass.
#pragma once
#include "fft_core.cuh"

#ifndef TWSTFT_ABLATE
#define TWSTFT_ABLATE 0      // timing experiments only (results are wrong), bit mask: 1 no stores, 2 no loads, 4 no butterflies
#endif

namespace twstft {

enum ColSrc { SRC_RAW = 0, SRC_CF32 = 1, SRC_BITS = 2 };
enum ColMode { COL_FWD = 0, COL_INV_ARGMAX = 1, COL_INV_STORE = 2, COL_INV_ARGMAX_STATS = 3, COL_INV_ARGMAX_K2 = 4 };
// COL_INV_ARGMAX_K2: besides the tile's first maximum, the maximum over the tile's rows OTHER than the row of that
// maximum (part_key2) -- the "everything else" bound of the fast peak path (k_fast.cuh).

struct ColParams {
    const void* src;
    float2* dst;
    const float2* blob;
    i64 N;                 // L * N2
    int N2;
    double invN;           // 1/N
    int ntx;               // column tiles per matrix
    int ny;                // matrices in this launch: records (SRC_RAW), lanes (CF32 fwd), lanes*nq (inverse)
    // SRC_RAW
    int nch;               // channels interleaved per sample instant (2: [I1 Q1 I2 Q2], 1: [I Q])
    // (the pre-transform operation is the kernel's OP template parameter:
    //  0 none, 1 derotate, 2 square, 4 fourth power;
    //  12/14 and 22/24: first radix-2 split of a 2N-sample block for the x^2/x^4 estimator --
    //  even bins from x^p[n] + x^p[n+N], odd bins from (x^p[n] - x^p[n+N]) * W_2N^n, W from lrow/lcol)
    i64 half_stride;       // elements between the two halves of a 2N-sample block (OP >= 10)
    int sub_mean;
    const double* sums;    // [lane][2] sum(I), sum(Q) over the window (exact integers)
    const float2* lrow;    // [lane][L]  exp(-j2pi df N2 n1/fs)
    const float2* lcol;    // [lane][N2] exp(-j2pi df n2/fs)
    i64 rec_stride;        // complex-int16 elements between records (all channels counted)
    const i64* rec_offset; // optional [record] start element (sample instants * nch); overrides rec_stride
    // SRC_BITS
    int rep;               // samples per chip (fs/fchip = 2)
    int qpsk;              // bits interleaved a,b -> (a + j b) - code_mean ; else 2*bit-1
    float2 code_mean;
    // COL_INV_*
    int nq;
    u64* part_key;         // [y][ntx]
    double* part_stat;     // COL_INV_ARGMAX_STATS: [y][ntx][2]  sum|z|, sum|z|^2 of the tile
    u64* part_key2;        // COL_INV_ARGMAX_K2: [y][ntx]
    const int* tile_list;  // COL_INV_ARGMAX_K2, optional: [y][ntx] column tile of work item (x, y) (a few tiles per lane)
    int pf_dist;           // > 0: every CTA pulls the rows of work item (its own + pf_dist) towards L2 before it starts
};

// N2C / NCH: row length and channel count as compile-time constants (0 = take them from Params), so that
// the strided global accesses of a column become immediate offsets from one base register.
template <class S, int TS, int SRC, int MODE, int OP = 0, int N2C = 0, int NCH = 0> struct KCol {
    typedef ColParams Params;
    static const char* name()
    {
        return MODE == COL_INV_ARGMAX ? "col_inv_argmax" : MODE == COL_INV_ARGMAX_STATS ? "col_inv_argmax_stats"
             : MODE == COL_INV_ARGMAX_K2 ? "col_inv_argmax_k2"
             : MODE == COL_INV_STORE ? "col_inv_store"
             : SRC == SRC_RAW ? ((OP % 10) >= 2 ? "col_fwd_raw_pow" : "col_fwd_raw") : SRC == SRC_BITS ? "col_fwd_bits" : OP != 0 ? "col_fwd_cf32_op" : "col_fwd_cf32";
    }
    static constexpr bool PERSISTENT = true;
    static constexpr bool INVERSE = (MODE != COL_FWD);
    static constexpr int NTA = S::NTS * TS;
    static constexpr int NT = (NTA + 31) / 32 * 32;
    static constexpr int MINB = (S::L * TS >= 8000) ? 2 : 4;
    typedef ColAcc<S::SP0, TS> Acc;
    static constexpr int NG = (MODE == COL_FWD) ? TS * S::RLAST : 0;
    static constexpr int OFF_G = S::BLOB_PAD;
    static constexpr int OFF_DATA = OFF_G + ((NG + 1) & ~1);
    static constexpr int NCOMP = S::NSTAGE == 1 ? 2 : S::NSTAGE;
    static constexpr bool STATS = (MODE == COL_INV_ARGMAX_STATS);
    static constexpr bool K2 = (MODE == COL_INV_ARGMAX_K2);
    static constexpr bool HASFIN = (MODE == COL_INV_ARGMAX || STATS || K2);
    static constexpr int NPH = NCOMP + (HASFIN ? 1 : 0) + (K2 ? 1 : 0);
    static constexpr int NSUM = STATS ? 2 : 1;
    static constexpr int NBL = (S::L / S::RLAST) / S::NTS;
    static constexpr int NB0 = S::SP0 / S::NTS;
    struct Regs {
        float2 v[S::E];
        u64 key;
        u64 key1;             // K2: the tile maximum while `key` is reused for the second search
        double dsum[NSUM];
    };
    static HD constexpr size_t smem_bytes()
    {
        return (size_t)(OFF_DATA + (S::NSTAGE > 1 ? S::L * TS : 0)) * sizeof(float2) + RED_SCRATCH_BYTES;
    }
    static HD constexpr int sync_after(int ph)
    {
        if (K2 && ph == NCOMP) return SYNC_MAXKEY;
        return (ph == NCOMP - 1 && HASFIN) ? (STATS ? SYNC_MAXKEY_SUMD : SYNC_MAXKEY) : SYNC_BARRIER;
    }
    static HD int n2of(const Params& p) { return N2C > 0 ? N2C : n2of(p); }
    static HD int nchof(const Params& p) { return NCH > 0 ? NCH : nchof(p); }
    static HD int niter(const Params& p, int bx, int gdim)
    {
        const i64 total = (i64)p.ntx * p.ny;
        return bx < total ? (int)((total - bx + gdim - 1) / gdim) : 0;
    }
    struct Work { int x, y, tile; };     // x: slot of the result arrays, tile: column tile (= x unless a tile list is given)
    static HD Work work_of(const Params& p, const Blk& blk, int it)
    {
        const i64 w = (i64)blk.x + (i64)it * blk.gx;
        Work k;
        k.x = (int)(w % p.ntx);
        k.y = (int)(w / p.ntx);
        k.tile = (MODE == COL_INV_ARGMAX_K2 && p.tile_list) ? p.tile_list[w] : k.x;
        return k;
    }

    // Column s of tile `tile` -> (channel, n2, element offset inside a raw row).  Interleaved two-channel records
    // (CHSPLIT): a tile holds TS sample instants of ONE channel (even tiles channel 0, odd tiles channel 1), so that the
    // stores of a row are full TS*8-byte runs; the two CTAs of a pair read the same raw sectors (the second from L2).
#ifndef TWSTFT_CHSPLIT
#define TWSTFT_CHSPLIT 1
#endif
    static constexpr bool CHSPLIT = (TWSTFT_CHSPLIT != 0) && SRC == SRC_RAW && NCH == 2;
    struct ColIdx { int ch, n2, rawcol; };
    static HD ColIdx colidx(const Params& p, int tile, int s)
    {
        ColIdx c;
        if (SRC != SRC_RAW) { c.ch = 0; c.n2 = tile * TS + s; c.rawcol = c.n2; }
        else if (CHSPLIT) { c.ch = tile & 1; c.n2 = (tile >> 1) * TS + s; c.rawcol = c.n2 * 2 + c.ch; }
        else { const int col = tile * TS + s; c.ch = col % nchof(p); c.n2 = col / nchof(p); c.rawcol = col; }
        return c;
    }

    static constexpr int POW = OP % 10;                 // 1 derotate, 2 / 4 power
    static constexpr int SPLIT = OP / 10;               // 0 none, 1 even half-sum, 2 odd half-difference
    static constexpr bool USE_LROW = (OP == 1 || SPLIT == 2);
    static HD float2 power_op(float2 x)
    {
        if (POW >= 2) {       // x*x as one packed complex multiply (two instructions instead of four scalar ones)
            x = cmul(x, x);
            if (POW == 4) x = cmul(x, x);
        }
        return x;
    }
    static HD float2 pre_op(float2 x, float2 lr)
    {
        if (OP == 1) return cmul(x, lr);
        return power_op(x);
    }

    // stage-0 operands straight from global memory (rows j + NTS*b + m*SP0 of column `col`),
    // with the int16 unpack / mean removal / pre-transform operation fused in
    static HD void load0(const Params& p, const Work& wk, int s, int j, Regs& r)
    {
        const ColIdx ci = colidx(p, wk.tile, s);
        const int col = ci.rawcol;
        if (SRC == SRC_RAW) {
            const i64 lane = (i64)wk.y * nchof(p) + ci.ch;
            float mx = 0.f, my = 0.f;
            if (p.sub_mean) {
                mx = (float)(p.sums[lane * 2] * p.invN);
                my = (float)(p.sums[lane * 2 + 1] * p.invN);
            }
            const i64 rowlen = (i64)nchof(p) * n2of(p);
            const i64 rbase = p.rec_offset ? p.rec_offset[wk.y] : (i64)wk.y * p.rec_stride;
            const short2* gsrc = reinterpret_cast<const short2*>(p.src) + rbase + (i64)j * rowlen + col;
            const i64 bstride = (i64)S::NTS * rowlen, mstride = (i64)S::SP0 * rowlen;
            const float2* lr = p.lrow + lane * S::L + j;
#pragma unroll
            for (int b = 0; b < NB0; b++)
#pragma unroll
                for (int m = 0; m < S::R0; m++) {
                    const short2 raw = ((TWSTFT_ABLATE & 2) && p.ny > 0) ? make_short2((short)(j + b), (short)(m + s))
                                                                        : gsrc[b * bstride + m * mstride];
                    const float2 x = make_float2((float)raw.x - mx, (float)raw.y - my);
                    if (SPLIT == 0) {
                        r.v[b * S::R0 + m] = pre_op(x, OP == 1 ? lr[S::NTS * b + m * S::SP0] : x);
                    } else {
                        const short2 raw2 = gsrc[p.half_stride + b * bstride + m * mstride];
                        const float2 x2 = make_float2((float)raw2.x - mx, (float)raw2.y - my);
                        const float2 pa = power_op(x), pb = power_op(x2);
                        r.v[b * S::R0 + m] = SPLIT == 1 ? cadd(pa, pb)
                                                        : cmul(csub(pa, pb), lr[S::NTS * b + m * S::SP0]);
                    }
                }
        } else if (SRC == SRC_CF32 && MODE == COL_FWD && OP != 0) {
            // complex-float windows handed over by a caller that already holds them in floating point (the Octave
            // seam processing(d,k): `d` is a complex column, godual_ranging.m:12): same pre-transform operations as
            // the int16 path, window base from rec_offset / rec_stride (in samples), no mean handling (the caller
            // of processing() removes the mean before the call, godual_ranging.m:80,94)
            const i64 rbase = p.rec_offset ? p.rec_offset[wk.y] : (i64)wk.y * p.rec_stride;
            const float2* gsrc = reinterpret_cast<const float2*>(p.src) + rbase + (i64)j * n2of(p) + col;
            const i64 bstride = (i64)S::NTS * n2of(p), mstride = (i64)S::SP0 * n2of(p);
            const float2* lr = p.lrow + (i64)wk.y * S::L + j;
#pragma unroll
            for (int b = 0; b < NB0; b++)
#pragma unroll
                for (int m = 0; m < S::R0; m++) {
                    const float2 x = gsrc[b * bstride + m * mstride];
                    if (SPLIT == 0) {
                        r.v[b * S::R0 + m] = pre_op(x, OP == 1 ? lr[S::NTS * b + m * S::SP0] : x);
                    } else {
                        const float2 x2 = gsrc[p.half_stride + b * bstride + m * mstride];
                        const float2 pa = power_op(x), pb = power_op(x2);
                        r.v[b * S::R0 + m] = SPLIT == 1 ? cadd(pa, pb)
                                                        : cmul(csub(pa, pb), lr[S::NTS * b + m * S::SP0]);
                    }
                }
        } else if (SRC == SRC_CF32) {
            const float2* gsrc = reinterpret_cast<const float2*>(p.src) + (i64)wk.y * p.N + (i64)j * n2of(p) + col;
            const i64 bstride = (i64)S::NTS * n2of(p), mstride = (i64)S::SP0 * n2of(p);
#pragma unroll
            for (int b = 0; b < NB0; b++)
#pragma unroll
                for (int m = 0; m < S::R0; m++) {
                    const float2 x = ((TWSTFT_ABLATE & 2) && p.ny > 0) ? make_float2((float)(j + b), (float)(m + s))
                                                                      : gsrc[b * bstride + m * mstride];
                    r.v[b * S::R0 + m] = (MODE == COL_FWD) ? x : cswap(x);
                }
        } else {  // SRC_BITS: sample n = n1*N2 + col holds chip n/rep
            const unsigned char* bits = reinterpret_cast<const unsigned char*>(p.src);
#pragma unroll
            for (int b = 0; b < NB0; b++)
#pragma unroll
                for (int m = 0; m < S::R0; m++) {
                    const i64 n = (i64)(j + S::NTS * b + m * S::SP0) * n2of(p) + col;
                    const i64 chip = n / p.rep;
                    r.v[b * S::R0 + m] = p.qpsk ? make_float2((float)bits[2 * chip] - p.code_mean.x,
                                                              (float)bits[2 * chip + 1] - p.code_mean.y)
                                                : make_float2(2.f * (float)bits[chip] - 1.f, 0.f);
                }
        }
    }

    // The CTA that will run work item w2 starts about one resident wave after this one: its TS-element row pieces
    // (L of them, a row apart) are requested now so that its stage-0 loads find them in L2.
    static HD void prefetch_item(const Params& p, i64 w2, int tid)
    {
        if (w2 >= (i64)p.ntx * p.ny) return;
        const int x2 = (int)(w2 % p.ntx), y2 = (int)(w2 / p.ntx);
        const char* base;
        size_t rstride;
        constexpr size_t ELEM = (SRC == SRC_RAW) ? sizeof(short2) : sizeof(float2);
        if (SRC == SRC_RAW || (SRC == SRC_CF32 && MODE == COL_FWD && OP != 0)) {
            const int nch = (SRC == SRC_RAW) ? nchof(p) : 1;
            const i64 rbase = p.rec_offset ? p.rec_offset[y2] : (i64)y2 * p.rec_stride;
            base = reinterpret_cast<const char*>(p.src) + (size_t)(rbase + (i64)(CHSPLIT ? (x2 >> 1) * 2 : x2) * TS) * ELEM;
            rstride = (size_t)nch * n2of(p) * ELEM;
        } else {
            base = reinterpret_cast<const char*>(p.src) + (size_t)((i64)y2 * p.N + (i64)x2 * TS) * ELEM;
            rstride = (size_t)n2of(p) * ELEM;
        }
        constexpr int PER = (int)(((CHSPLIT ? 2 : 1) * TS * ELEM + 127) / 128);
        for (int l = tid; l < S::L * PER; l += NTA) prefetch_l2(base + (size_t)(l / PER) * rstride + (size_t)(l % PER) * 128);
    }

    static HD void epilogue(const Params& p, const Work& wk, int s, int j, Regs& r, const float2* g)
    {
        const ColIdx ci = colidx(p, wk.tile, s);
        const int ch = ci.ch, n2 = ci.n2;
        if (MODE == COL_FWD) {
            const i64 lane = (SRC == SRC_RAW) ? (i64)wk.y * nchof(p) + ch : (i64)wk.y;
            float2* d = p.dst + lane * p.N + n2;
            float2 base[NBL];
#pragma unroll
            for (int b = 0; b < NBL; b++) {
                const int jb = j + S::NTS * b;
                base[b] = cis_cycles(-(double)((i64)n2 * jb) * p.invN);
                if ((SRC == SRC_RAW || SRC == SRC_CF32) && USE_LROW) base[b] = cmul(base[b], p.lcol[lane * n2of(p) + n2]);
            }
            const i64 bstr = (i64)S::NTS * n2of(p), mstr = (i64)S::NSLAST * n2of(p);
            float2* dj = d + (i64)j * n2of(p);
            float2 ablsum = make_float2(0.f, 0.f);
            last_stage_visit<S>(j, r.v, [&](int, float2 x, int b, int m) {
                const float2 y = cmul(x, cmul(base[b], g[m * TS + s]));
                if (TWSTFT_ABLATE & 1) ablsum = cadd(ablsum, y);     // experiment: one store per thread instead of 25
                else dj[b * bstr + m * mstr] = y;
            });
            if (TWSTFT_ABLATE & 1) dj[0] = ablsum;
        } else if (MODE == COL_INV_STORE) {
            float2* dj = p.dst + (i64)wk.y * p.N + n2 + (i64)j * n2of(p);
            const i64 bstr = (i64)S::NTS * n2of(p), mstr = (i64)S::NSLAST * n2of(p);
            last_stage_visit<S>(j, r.v, [&](int, float2 x, int b, int m) { dj[b * bstr + m * mstr] = cswap(x); });
        } else {  // COL_INV_ARGMAX: outputs are visited in increasing index, so '>' keeps the first maximum
            const int q = wk.y % p.nq;
            float bestv = -1.f;
            int bestc = 0;                  // output index = j + bestc; b and m are compile-time after unrolling
            float s1 = 0.f, s2 = 0.f;
            last_stage_visit<S>(j, r.v, [&](int, float2 x, int b, int m) {
                const float v = cnorm(x);
                if (STATS) { s1 += sqrtf(v); s2 += v; }
                const int c = S::NTS * b + m * S::NSLAST;
                if (v > bestv || (NBL > 1 && v == bestv && c < bestc)) { bestv = v; bestc = c; }
            });
            const unsigned besti = (unsigned)p.nq * (unsigned)n2 + (unsigned)q +
                                   (unsigned)p.nq * (unsigned)n2of(p) * (unsigned)(j + bestc);
            r.key = make_key(bestv, besti);
            if (STATS) { r.dsum[0] = (double)s1; r.dsum[1] = (double)s2; }
        }
    }

    static HD void prologue(const Params& p, int tid, unsigned char* smem)
    {
        float2* sm = reinterpret_cast<float2*>(smem);
        for (int i = tid; i < S::BLOB; i += NT) sm[i] = p.blob[i];
    }

    template <int PH>
    static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char* smem, int it, int nit)
    {
        float2* sm = reinterpret_cast<float2*>(smem);
        if (tid >= NTA) {
            r.key = 0;
            for (int i = 0; i < NSUM; i++) r.dsum[i] = 0.0;
            return;
        }
        const int s = tid % TS, j = tid / TS;
        const Acc acc{sm + OFF_DATA + s};
        const Work wk = work_of(p, blk, it);
        float2* gtab = sm + OFF_G;

        if (PH == 0) {
            r.key = 0;
            for (int i = 0; i < NSUM; i++) r.dsum[i] = 0.0;
            if (p.pf_dist > 0 && !K2 && SRC != SRC_BITS) prefetch_item(p, (i64)blk.x + (i64)it * blk.gx + p.pf_dist, tid);
            if (MODE == COL_FWD) {
                for (int t = tid; t < NG; t += NTA) {
                    const int s2 = t % TS, m = t / TS;
                    const i64 n2b = colidx(p, wk.tile, s2).n2;
                    gtab[m * TS + s2] = cis_cycles(-(double)(n2b * S::NSLAST * m) * p.invN);
                }
            }
        }
        if (PH == (S::NSTAGE == 1 ? 1 : 0)) {
            load0(p, wk, s, j, r);
            if (!(TWSTFT_ABLATE & 4) || p.ny < 0) dif_stage0<S>(j, r.v, sm, acc);
            if (S::NSTAGE == 1) epilogue(p, wk, s, j, r, gtab);
        }
        if (S::NSTAGE == 3 && PH == 1) dif_stage_mid<S>(j, r.v, sm, acc);
        if (S::NSTAGE >= 2 && PH == S::NSTAGE - 1) {
            if (!(TWSTFT_ABLATE & 4) || p.ny < 0) dif_stage_last<S>(j, r.v, acc);
            epilogue(p, wk, s, j, r, gtab);
        }
        if (K2 && PH == NCOMP) {
            // second search: this thread's outputs outside the row of the tile maximum (values are still in r.v)
            r.key1 = r.key;
            const unsigned rowlen = (unsigned)p.nq * (unsigned)n2of(p);
            const int row1 = (int)(key_index(r.key) / rowlen);
            const int q = wk.y % p.nq;
            const int col = wk.tile * TS + s;
            u64 best = 0;
            last_stage_visit<S>(j, r.v, [&](int, float2 x, int b, int m) {
                const int c = S::NTS * b + m * S::NSLAST;
                if (j + c != row1) {
                    const unsigned idx = (unsigned)p.nq * (unsigned)col + (unsigned)q + rowlen * (unsigned)(j + c);
                    const u64 k = make_key(cnorm(x), idx);
                    best = k > best ? k : best;
                }
            });
            r.key = best;
        }
        if (K2 && PH == NCOMP + 1) {
            if (tid == 0) {
                p.part_key[(i64)wk.y * p.ntx + wk.x] = r.key1;
                p.part_key2[(i64)wk.y * p.ntx + wk.x] = r.key;
            }
        }
        if (!K2 && HASFIN && PH == NCOMP) {
            if (tid == 0) {
                p.part_key[(i64)wk.y * p.ntx + wk.x] = r.key;
                if (STATS) {
                    p.part_stat[((i64)wk.y * p.ntx + wk.x) * 2] = r.dsum[0];
                    p.part_stat[((i64)wk.y * p.ntx + wk.x) * 2 + 1] = r.dsum[1];
                }
            }
        }
    }
};

}  // namespace twstft
