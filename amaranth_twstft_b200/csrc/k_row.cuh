// Row pass of the four-step FFT: one contiguous length-L transform per row (L = S::L, e.g. 8000),
// data layout [lane][k1 = row][L].  Replaces the FFT calls of the reference at
//   processing/Octave/godual_ranging.m:14 (spectrum of d.^2), :25 (ffty), :26-28 (first half of the
//   3N ifft) and acquisition/claudio_aligned_code_lo_separate.m:29,34-36.
//
// Bin bookkeeping: after the column pass (k_col.cuh) row k1 holds A[k1][n2]; this pass transforms
// over n2 -> k2 and position (k1,k2) holds bin k = k1 + N1*k2 ("transposed" spectrum order, never
// un-transposed: the code spectrum, the pad/twist tables and the windows use the same order).
//
// MODES
//   ROW_FWD_STORE   plain store                               (code spectrum C = fft(code))
//   ROW_FWD_ARGMAX  |X|^2 first-max inside a window of the fftshift'ed axis (carrier estimator A3)
//   ROW_FWD_MULCODE P = Y*conj(C) (G) or conj(Y)*C (K), sum|Y|^2, Y[0]   (A5/A6 front half)
//   ROW_INV_TWIST   inverse over k2 of P*twist_q (polyphase q of the x3 interpolation), times the
//                   conjugate four-step twiddle; output [lane][q][k1][n2]         (A6 back half)
#pragma once
#include "fft_core.cuh"

namespace twstft {

enum RowMode { ROW_FWD_STORE = 0, ROW_FWD_ARGMAX = 1, ROW_FWD_MULCODE = 2, ROW_INV_TWIST = 3 };

struct RowParams {
    const float2* src;
    float2* dst;
    const float2* blob;      // S::BLOB twiddle entries
    i64 N;                   // N1 * L
    int N1;
    // ROW_FWD_ARGMAX
    i64 win_lo, win_hi;      // inclusive window, index s of fftshift(X): s = (k + N/2) mod N
    u64* part_key;           // [lane][gridDim.x]
    // ROW_FWD_MULCODE
    const float2* code;      // C in the transposed order [k1][k2]
    int conj_mode;           // 0: Y*conj(C) (variant G)   1: conj(Y)*C (variant K)
    double* part_energy;     // [lane][gridDim.x]  sum |Y|^2
    float2* y0;              // [lane] Y[0]
    // ROW_INV_TWIST
    const float2* twq;       // [nq][L] per-k2 twist (includes the negative-side factor), row 0 unused
    int nq;                  // 3 (x3 interpolation) or 1
    int nblk;                // gridDim.x (for partial indexing)
};

template <class S, int RPC, int MODE> struct KRow {
    typedef RowParams Params;
    static const char* name()
    {
        return MODE == ROW_FWD_STORE ? "row_fwd_store" : MODE == ROW_FWD_ARGMAX ? "row_fwd_argmax"
             : MODE == ROW_FWD_MULCODE ? "row_fwd_mulcode" : "row_inv_twist";
    }
    static constexpr int NTA = S::NTS * RPC;
    static constexpr int NT = (NTA + 31) / 32 * 32;
    static constexpr int MINB = (S::L * RPC >= 8000) ? 2 : 4;
    static constexpr int ROWSM = S::L + S::L / S::R0 + 2;      // padded row (float2 units)
    static constexpr int NG = (MODE == ROW_INV_TWIST) ? RPC * S::RLAST : 0;
    static constexpr int OFF_G = S::BLOB_PAD;
    static constexpr int OFF_DATA = OFF_G + ((NG + 1) & ~1);
    static constexpr int NCOMP = S::NSTAGE == 3 ? 4 : 2;   // single-stage: phase 0 = tables only
    static constexpr int P_S0 = S::NSTAGE == 1 ? 1 : 0;       // phase that runs stage 0
    static constexpr bool HASFIN = (MODE == ROW_FWD_ARGMAX || MODE == ROW_FWD_MULCODE);
    static constexpr int NPH = NCOMP + (HASFIN ? 1 : 0);
    static constexpr int NSUM = 1;
    static constexpr int NBL = (S::L / S::RLAST) / S::NTS;     // butterflies per thread in the last stage
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
        return (ph == NCOMP - 1) ? (MODE == ROW_FWD_ARGMAX ? SYNC_MAXKEY : (MODE == ROW_FWD_MULCODE ? SYNC_SUMD : SYNC_BARRIER))
                                 : SYNC_BARRIER;
    }
    static HD int pad(int i) { return i + i / S::R0; }

    static HD void epilogue(const Params& p, const Blk& blk, int rl, int j, Regs& r, const float2* sm)
    {
        const int row = blk.x * RPC + rl;
        const i64 lane = blk.y;
        const i64 rowoff = (i64)row * S::L;
        if (MODE == ROW_FWD_STORE) {
            float2* d = p.dst + lane * p.N + rowoff;
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int, int) { d[o] = x; });
        } else if (MODE == ROW_FWD_ARGMAX) {
            u64 best = 0;
            const i64 half = p.N / 2;
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int, int) {
                i64 s = (i64)row + (i64)p.N1 * o + half;
                if (s >= p.N) s -= p.N;
                if (s >= p.win_lo && s <= p.win_hi) {
                    u64 k = make_key(cnorm(x), (unsigned)s);
                    best = k > best ? k : best;
                }
            });
            r.key = best;
        } else if (MODE == ROW_FWD_MULCODE) {
            float2* d = p.dst + lane * p.N + rowoff;
            const float2* c = p.code + rowoff;
            float e = 0.f;
            const int cm = p.conj_mode;
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int, int) {
                float2 cc = c[o];
                e += cnorm(x);
                d[o] = cm ? cmul(cconj(x), cc) : cmulc(x, cc);
                if (row == 0 && o == 0) p.y0[lane] = x;
            });
            r.dsum[0] = (double)e;
        } else {  // ROW_INV_TWIST
            const int q = blk.z;
            float2* d = p.dst + (lane * p.nq + q) * p.N + rowoff;
            const float2* g = sm + OFF_G + rl * S::RLAST;
            float2 base[NBL];
#pragma unroll
            for (int b = 0; b < NBL; b++) {
                const int jb = j + S::NTS * b;
                base[b] = cis_cycles((double)((i64)row * jb) / (double)p.N +
                                     (double)((i64)q * row) / ((double)p.nq * (double)p.N));
            }
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int b, int m) {
                d[o] = cmul(cswap(x), cmul(base[b], g[m]));
            });
        }
    }

    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char* smem)
    {
        float2* sm = reinterpret_cast<float2*>(smem);
        if (tid >= NTA) {
            r.key = 0;
            r.dsum[0] = 0.0;
            return;
        }
        const int rl = tid / S::NTS, j = tid % S::NTS;
        const int row = blk.x * RPC + rl;
        const i64 lane = blk.y;
        float2* data = sm + OFF_DATA + rl * ROWSM;
        auto sm_src = [&](int i) { return data[pad(i)]; };
        auto sm_dst = [&](int i, float2 x) { data[pad(i)] = x; };
        if (PH == 0) {
            r.key = 0;
            r.dsum[0] = 0.0;
            for (int i = tid; i < S::BLOB; i += NTA) sm[i] = p.blob[i];
            if (MODE == ROW_INV_TWIST) {
                for (int t = tid; t < NG; t += NTA) {
                    const int rl2 = t / S::RLAST, m = t % S::RLAST;
                    const i64 row2 = (i64)blk.x * RPC + rl2;
                    sm[OFF_G + t] = cis_cycles((double)(row2 * S::NSLAST * m) / (double)p.N);
                }
            }
        }
        if (PH == P_S0) {
            const float2* s = p.src + lane * p.N + (i64)row * S::L;
            if (MODE == ROW_INV_TWIST) {
                const int q = blk.z;
                const float2* tq = p.twq + (i64)q * S::L;
                stage_compute<S, 0>(j, r.v, sm, [&](int i) {
                    float2 z = s[i];
                    if (q > 0) z = cmul(z, tq[i]);
                    return cswap(z);
                });
            } else {
                stage_compute<S, 0>(j, r.v, sm, [&](int i) { return s[i]; });
            }
            if (S::NSTAGE == 1) epilogue(p, blk, rl, j, r, sm);
            else stage_scatter<S, 0>(j, r.v, sm_dst);
        }
        if (S::NSTAGE >= 2 && PH == 1) {
            stage_compute<S, 1>(j, r.v, sm, sm_src);
            if (S::NSTAGE == 2) epilogue(p, blk, rl, j, r, sm);
        }
        if (S::NSTAGE == 3 && PH == 2) stage_scatter<S, 1>(j, r.v, sm_dst);
        if (S::NSTAGE == 3 && PH == 3) {
            stage_compute<S, 2>(j, r.v, sm, sm_src);
            epilogue(p, blk, rl, j, r, sm);
        }
        if (HASFIN && PH == NCOMP) {
            if (tid == 0) {
                if (MODE == ROW_FWD_ARGMAX) p.part_key[lane * p.nblk + blk.x] = r.key;
                if (MODE == ROW_FWD_MULCODE) p.part_energy[lane * p.nblk + blk.x] = r.dsum[0];
            }
        }
    }
};

}  // namespace twstft
