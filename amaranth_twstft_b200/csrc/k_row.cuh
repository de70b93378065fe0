// Row pass of the four-step FFT: one contiguous length-L transform per row (L = S::L, e.g. 8000),
// data layout [lane][k1 = row][L].  Replaces the FFT calls of the reference at
//   processing/Octave/godual_ranging.m:14 (spectrum of d.^2), :25 (ffty), :26-28 (first half of the
//   3N ifft) and acquisition/claudio_aligned_code_lo_separate.m:29,34-36.
//
// Bin bookkeeping: after the column pass (k_col.cuh) row k1 holds A[k1][n2]; this pass transforms
// over n2 -> k2 and position (k1,k2) holds bin k = k1 + N1*k2 ("transposed" spectrum order, never
// un-transposed: the code spectrum, the pad/twist tables and the windows use the same order).
//
// Work items (row group, lane[, polyphase q]) are walked by whatever grid is launched: one CTA per
// item, or a resident wave of CTAs looping over items (the launcher inserts a barrier between items).
//
// MODES
//   ROW_FWD_STORE   plain store                               (code spectrum C = fft(code))
//   ROW_FWD_ARGMAX  |X|^2 first-max inside a window of the fftshift'ed axis (carrier estimator A3)
//   ROW_FWD_WINMAG  |X|^2 of every bin inside the window, stored densely (search_df candidates, A4)
// The correlator's own row pass (forward, code product, inverse polyphases) is k_rowcorr.cuh.
#pragma once
#include "fft_core.cuh"

namespace twstft {

enum RowMode { ROW_FWD_STORE = 0, ROW_FWD_ARGMAX = 1, ROW_FWD_WINMAG = 2 };

struct RowParams {
    const float2* src;
    float2* dst;
    const float2* blob;      // S::BLOB twiddle entries
    i64 N;                   // N1 * L
    int N1;
    double invN, invM;       // 1/N, 1/(nq*N)
    int nlanes;              // lanes in this launch
    int nblk;                // row groups per lane (N1 / RPC)
    // ROW_FWD_ARGMAX
    i64 win_lo, win_hi;      // inclusive window, index s of fftshift(X): s = (k + N/2) mod N
    u64* part_key;           // [lane*key_stride + key_off + rowblk]
    int key_stride, key_off;
    // bin k of this transform is bin bmul*k + badd of a transform of nbig points (radix-2 split of a
    // 2N-sample block, k_col.cuh); the window and the stored index live on that axis.  Default 1, 0, N.
    int bmul;
    i64 badd, nbig;
    int oa1, ob1, oa2, ob2;  // k2 ranges that can intersect the window (superset; exact test follows)
    unsigned k1mask;         // != 0: output-pruned form (KRow::PRUNABLE): bit k1 set <=> some k2 = k0 + R0 k1 + R0 R1 m of
                             // those ranges needs the middle-stage output k1; the last stage is then one dot product per bin
    // ROW_FWD_WINMAG
    float* winmag;           // [lane][win_hi - win_lo + 1]  |X|^2 at shifted index s = win_lo + i
};

template <class S, int RPC, int MODE> struct KRow {
    typedef RowParams Params;
    static const char* name()
    {
        return MODE == ROW_FWD_STORE ? "row_fwd_store" : MODE == ROW_FWD_ARGMAX ? "row_fwd_argmax" : "row_fwd_winmag";
    }
    static constexpr bool PERSISTENT = true;
    static constexpr int NTA = S::NTS * RPC;
    static constexpr int NT = (NTA + 31) / 32 * 32;
    static constexpr int MINB = (S::L * RPC >= 8000) ? 2 : 4;
    typedef RowAcc<S::SP0> Acc;
    static constexpr int ROWSM = (S::R0 * Acc::BSTRIDE + 1) & ~1;      // padded row (float2 units)
    // estimator modes on the 20.20.20 rows: middle and last stage output-pruned when the window is narrow (k1mask)
    static constexpr bool PRUNABLE = (MODE != ROW_FWD_STORE) && S::NSTAGE == 3 && S::R0 == 20 && S::R1 == 20 && S::R2 == 20 &&
                                     RPC == 1;
    static constexpr int NG = PRUNABLE ? S::RLAST : 0;          // W_20^i, i = 0..19 (last-stage dot products)
    static constexpr int OFF_G = S::BLOB_PAD;
    static constexpr int OFF_DATA = OFF_G + ((NG + 1) & ~1);
    static constexpr int NCOMP = S::NSTAGE == 1 ? 2 : S::NSTAGE;       // 1-stage: phase 0 = tables only
    static constexpr bool HASFIN = (MODE == ROW_FWD_ARGMAX);
    static constexpr int NPH = NCOMP + (HASFIN ? 1 : 0);
    static constexpr int NSUM = 1;
    static constexpr int NBL = (S::L / S::RLAST) / S::NTS;     // butterflies per thread in the last stage
    static constexpr int NB0 = S::SP0 / S::NTS;                // stage-0 butterflies per thread
    struct Regs {
        float2 v[S::E];
        u64 key;
        double dsum[NSUM];
    };
    static HD constexpr size_t smem_bytes()
    {
        return (size_t)(OFF_DATA + (S::NSTAGE > 1 ? RPC * ROWSM : 0)) * sizeof(float2) + RED_SCRATCH_BYTES;
    }
    static HD constexpr int sync_after(int ph)
    {
        return (ph == NCOMP - 1 && MODE == ROW_FWD_ARGMAX) ? SYNC_MAXKEY_ATOMIC : SYNC_BARRIER;
    }
    // ROW_FWD_ARGMAX: the row's first-maximum key goes straight into its slot of part_key (zeroed by the host before
    // the launch) with one atomic per warp that holds a candidate -- no block reduction, no barrier (it used to be two
    // barriers and ~85 instructions per thread and row)
    static HD u64* key_slot(const Params& p, const Blk& blk, int it)
    {
        const Work wk = work_of(p, blk, it);
        return p.part_key + wk.lane * p.key_stride + p.key_off + wk.rowblk;
    }
    // work decomposition: item w -> (rowblk = w % nblk, lane = w / nblk)
    static HD int gx(const Params& p) { return p.nblk; }
    static HD i64 total_work(const Params& p) { return (i64)gx(p) * p.nlanes; }
    static HD int niter(const Params& p, int bx, int gdim)
    {
        const i64 total = total_work(p);
        return bx < total ? (int)((total - bx + gdim - 1) / gdim) : 0;
    }
    struct Work { int rowblk, q; i64 lane; };
    static HD Work work_of(const Params& p, const Blk& blk, int it)
    {
        const i64 w = (i64)blk.x + (i64)it * blk.gx;
        const int g = gx(p);
        const int x = (int)(w % g);
        Work k;
        k.lane = w / g;
        k.rowblk = x;
        k.q = 0;
        return k;
    }

    // bits m (0..RLAST-1) for which the output index jb + m*NSLAST lies in [oa, ob]
    static HD unsigned slot_mask(int oa, int ob, int jb)
    {
        if (ob < oa) return 0u;
        const int lo = oa - jb, hi = ob - jb;
        if (hi < 0) return 0u;
        const int mlo = lo <= 0 ? 0 : (lo + S::NSLAST - 1) / S::NSLAST;
        int mhi = hi / S::NSLAST;
        if (mhi > S::RLAST - 1) mhi = S::RLAST - 1;
        if (mlo > mhi) return 0u;
        return ((2u << mhi) - 1u) & ~((1u << mlo) - 1u);
    }

    // stage-0 operands straight from global memory: element jb + m*SP0 of this thread's row
    static HD void load0(const Params& p, const Work& wk, int rl, int j, Regs& r)
    {
        const float2* xj = p.src + wk.lane * p.N + (i64)(wk.rowblk * RPC + rl) * S::L + j;
        {
#pragma unroll
            for (int b = 0; b < NB0; b++)
#pragma unroll
                for (int m = 0; m < S::R0; m++) r.v[b * S::R0 + m] = xj[S::NTS * b + m * S::SP0];
        }
    }

    // one spectrum bin of row `row`, output index o: exact window test on the fftshift'ed axis, then key / magnitude
    static HD void emit_bin(const Params& p, i64 lane, int row, int o, float2 x, u64& best)
    {
        const i64 half = p.nbig / 2;
        i64 s = ((i64)row + (i64)p.N1 * o) * p.bmul + p.badd + half;
        if (s >= p.nbig) s -= p.nbig;
        if (s >= p.win_lo && s <= p.win_hi) {
            if (MODE == ROW_FWD_ARGMAX) {
                const u64 k = make_key(cnorm(x), (unsigned)s);
                best = k > best ? k : best;
            } else if (MODE == ROW_FWD_WINMAG) {
                p.winmag[lane * (p.win_hi - p.win_lo + 1) + (s - p.win_lo)] = cnorm(x);
            }
        }
    }

    // output-pruned last stage: thread t of the row forms the t-th wanted bin as a 20-term dot product over i2
    static HD void pruned_last(const Params& p, const Work& wk, int j, Regs& r, const Acc& acc, const float2* w20)
    {
        const int n1 = p.ob1 >= p.oa1 ? p.ob1 - p.oa1 + 1 : 0, n2 = p.ob2 >= p.oa2 ? p.ob2 - p.oa2 + 1 : 0;
        u64 best = 0;
        for (int t = j; t < n1 + n2; t += S::NTS) {
            const int o = t < n1 ? p.oa1 + t : p.oa2 + (t - n1);
            const int jb = o % S::NSLAST, m = o / S::NSLAST;
            const float2* q = acc.at(jb % S::R0, (jb / S::R0) * S::RLAST);
            float2 x = q[0];
            int e = 0;
#pragma unroll
            for (int i2 = 1; i2 < S::RLAST; i2++) {
                e += m;
                if (e >= S::RLAST) e -= S::RLAST;
                x = cadd(x, cmul(q[i2 * Acc::unit()], w20[e]));
            }
            emit_bin(p, wk.lane, wk.rowblk * RPC, o, x, best);
        }
        if (MODE == ROW_FWD_ARGMAX) r.key = best;
    }

    static HD void epilogue(const Params& p, const Work& wk, int rl, int j, Regs& r, const float2* g)
    {
        const int row = wk.rowblk * RPC + rl;
        const i64 lane = wk.lane;
        const i64 rowoff = (i64)row * S::L;
        if (MODE == ROW_FWD_STORE) {
            float2* d = p.dst + lane * p.N + rowoff;
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int, int) { d[o] = x; });
        } else if (MODE == ROW_FWD_ARGMAX) {
            // Cheap reject first: only k2 in [oa1,ob1] u [oa2,ob2] (host-computed superset) can fall in
            // the window; turn that into a per-thread bit mask over the output slots m, then run the
            // exact test on the fftshift'ed index s for those few only.
            u64 best = 0;
            const i64 half = p.nbig / 2;
            unsigned mask[NBL];
#pragma unroll
            for (int b = 0; b < NBL; b++) {
                const int jb = j + S::NTS * b;
                mask[b] = slot_mask(p.oa1, p.ob1, jb) | slot_mask(p.oa2, p.ob2, jb);
            }
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int b, int m) {
                if (mask[b] & (1u << m)) {
                    i64 s = ((i64)row + (i64)p.N1 * o) * p.bmul + p.badd + half;
                    if (s >= p.nbig) s -= p.nbig;
                    if (s >= p.win_lo && s <= p.win_hi) {
                        u64 k = make_key(cnorm(x), (unsigned)s);
                        best = k > best ? k : best;
                    }
                }
            });
            r.key = best;
        } else {  // ROW_FWD_WINMAG: same window test as the arg-max, values stored instead of reduced
            const i64 half = p.nbig / 2;
            const i64 wlen = p.win_hi - p.win_lo + 1;
            float* wm = p.winmag + lane * wlen;
            unsigned mask[NBL];
#pragma unroll
            for (int b = 0; b < NBL; b++) {
                const int jb = j + S::NTS * b;
                mask[b] = slot_mask(p.oa1, p.ob1, jb) | slot_mask(p.oa2, p.ob2, jb);
            }
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int b, int m) {
                if (mask[b] & (1u << m)) {
                    i64 s = ((i64)row + (i64)p.N1 * o) * p.bmul + p.badd + half;
                    if (s >= p.nbig) s -= p.nbig;
                    if (s >= p.win_lo && s <= p.win_hi) wm[s - p.win_lo] = cnorm(x);
                }
            });
        }
    }

    static HD void prologue(const Params& p, int tid, unsigned char* smem)
    {
        float2* sm = reinterpret_cast<float2*>(smem);
        for (int i = tid; i < S::BLOB; i += NT) sm[i] = p.blob[i];
        if (PRUNABLE && tid < S::RLAST) sm[OFF_G + tid] = cis_cycles(-(double)tid / (double)S::RLAST);
    }

    template <int PH>
    static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char* smem, int it, int nit)
    {
        float2* sm = reinterpret_cast<float2*>(smem);
        if (tid >= NTA) {
            r.key = 0;
            r.dsum[0] = 0.0;
            return;
        }
        const int rl = tid / S::NTS, j = tid % S::NTS;
        const Acc acc{sm + OFF_DATA + rl * ROWSM};
        const Work wk = work_of(p, blk, it);
        float2* gtab = sm + OFF_G;

        if (PH == 0) {
            r.key = 0;
            r.dsum[0] = 0.0;
        }
        if (PH == (S::NSTAGE == 1 ? 1 : 0)) {
            load0(p, wk, rl, j, r);
            dif_stage0<S>(j, r.v, sm, acc);
            if (S::NSTAGE == 1) epilogue(p, wk, rl, j, r, gtab);
        }
        if (S::NSTAGE >= 2 && PH == 1 && it + 1 < nit) {   // next item's rows towards L2
            const Work nx = work_of(p, blk, it + 1);
            const char* a = reinterpret_cast<const char*>(p.src + nx.lane * p.N + (i64)nx.rowblk * RPC * S::L);
            constexpr int NLINE = (int)((size_t)RPC * S::L * sizeof(float2) / 128);
            for (int l = tid; l < NLINE; l += NTA) prefetch_l2(a + (size_t)l * 128);
        }
        if (PRUNABLE && p.k1mask != 0u) {
            if constexpr (PRUNABLE) {
                if (PH == 1) {
                    constexpr unsigned LOCAL = 0xC0003u;       // k1 = 0, 1, 18, 19: |f| < 20 kHz on the 625 x 8000 split
                    if (p.k1mask == LOCAL) dif_stage_mid_pruned<S, Acc, LOCAL>(j, r.v, sm, acc, LOCAL);
                    else dif_stage_mid_pruned<S, Acc>(j, r.v, sm, acc, p.k1mask);
                }
                if (PH == 2) pruned_last(p, wk, j, r, acc, gtab);
            }
        } else if (S::NSTAGE == 3 && PH == 1) dif_stage_mid<S>(j, r.v, sm, acc);
        if (S::NSTAGE >= 2 && PH == S::NSTAGE - 1 && !(PRUNABLE && p.k1mask != 0u)) {
            // estimator modes: a thread whose last-stage outputs all fall outside the window (most of them: the
            // window is a few dozen of the L bins of a row) skips its last butterfly; whole warps drop out
            bool need = true;
            if (MODE != ROW_FWD_STORE) {
                need = false;
#pragma unroll
                for (int b = 0; b < NBL; b++) {
                    const int jb = j + S::NTS * b;
                    need = need || (slot_mask(p.oa1, p.ob1, jb) | slot_mask(p.oa2, p.ob2, jb)) != 0u;
                }
            }
            if (need) {
                dif_stage_last<S>(j, r.v, acc);
                epilogue(p, wk, rl, j, r, gtab);
            }
        }
        // (phase NCOMP of ROW_FWD_ARGMAX is empty: the key was delivered by the SYNC_MAXKEY_ATOMIC step before it)
    }
};

}  // namespace twstft
