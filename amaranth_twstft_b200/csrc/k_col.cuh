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
#pragma once
#include "fft_core.cuh"

namespace twstft {

enum ColSrc { SRC_RAW = 0, SRC_CF32 = 1, SRC_BITS = 2 };
enum ColMode { COL_FWD = 0, COL_INV_ARGMAX = 1, COL_INV_STORE = 2 };

struct ColParams {
    const void* src;
    float2* dst;
    const float2* blob;
    i64 N;                 // L * N2
    int N2;
    // SRC_RAW
    int nch;               // channels interleaved per sample instant (2: [I1 Q1 I2 Q2], 1: [I Q])
    int op;                // 0 none, 1 derotate, 2 square, 4 fourth power
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
    u64* part_key;         // [blk.y][ntile]
    int ntile;             // gridDim.x
};

template <class S, int TS, int SRC, int MODE> struct KCol {
    typedef ColParams Params;
    static const char* name()
    {
        return MODE == COL_INV_ARGMAX ? "col_inv_argmax" : MODE == COL_INV_STORE ? "col_inv_store"
             : SRC == SRC_RAW ? "col_fwd_raw" : SRC == SRC_BITS ? "col_fwd_bits" : "col_fwd_cf32";
    }
    static constexpr int NTA = S::NTS * TS;
    static constexpr int NT = (NTA + 31) / 32 * 32;
    static constexpr int MINB = (S::L * TS >= 8000) ? 2 : 4;
    static constexpr int NG = (MODE == COL_FWD) ? TS * S::RLAST : 0;
    static constexpr int OFF_G = S::BLOB_PAD;
    static constexpr int OFF_DATA = OFF_G + ((NG + 1) & ~1);
    static constexpr int NCOMP = S::NSTAGE == 3 ? 4 : 2;   // single-stage: phase 0 = tables only
    static constexpr int P_S0 = S::NSTAGE == 1 ? 1 : 0;       // phase that runs stage 0
    static constexpr bool HASFIN = (MODE == COL_INV_ARGMAX);
    static constexpr int NPH = NCOMP + (HASFIN ? 1 : 0);
    static constexpr int NSUM = 1;
    static constexpr int NBL = (S::L / S::RLAST) / S::NTS;
    struct Regs {
        float2 v[S::E];
        u64 key;
        double dsum[NSUM];
    };
    static HD constexpr size_t smem_bytes()
    {
        return (size_t)(OFF_DATA + (S::NSTAGE > 1 ? S::L * TS : 0)) * sizeof(float2) + RED_SCRATCH_BYTES;
    }
    static HD constexpr int sync_after(int ph)
    {
        return (ph == NCOMP - 1 && MODE == COL_INV_ARGMAX) ? SYNC_MAXKEY : SYNC_BARRIER;
    }

    // ---- element fetch for stage 0 (n1 = idx) ------------------------------------------------
    static HD float2 fetch(const Params& p, const Blk& blk, int col, int idx, float mx, float my, i64 lane)
    {
        if (SRC == SRC_RAW) {
            const i64 rowlen = (i64)p.nch * p.N2;
            const i64 base = p.rec_offset ? p.rec_offset[blk.y] : (i64)blk.y * p.rec_stride;
            const short2 raw = reinterpret_cast<const short2*>(p.src)[base + (i64)idx * rowlen + col];
            float2 x = make_float2((float)raw.x - mx, (float)raw.y - my);
            if (p.op == 1) x = cmul(x, p.lrow[lane * S::L + idx]);
            else if (p.op >= 2) {
                x = make_float2(x.x * x.x - x.y * x.y, 2.f * x.x * x.y);
                if (p.op == 4) x = make_float2(x.x * x.x - x.y * x.y, 2.f * x.x * x.y);
            }
            return x;
        } else if (SRC == SRC_CF32) {
            const float2 x = reinterpret_cast<const float2*>(p.src)[(i64)blk.y * p.N + (i64)idx * p.N2 + col];
            return (MODE == COL_FWD) ? x : cswap(x);
        } else {  // SRC_BITS: sample n = idx*N2 + col holds chip n/rep
            const i64 n = (i64)idx * p.N2 + col;
            const i64 chip = n / p.rep;
            const unsigned char* bits = reinterpret_cast<const unsigned char*>(p.src);
            if (p.qpsk) return make_float2((float)bits[2 * chip] - p.code_mean.x, (float)bits[2 * chip + 1] - p.code_mean.y);
            return make_float2(2.f * (float)bits[chip] - 1.f, 0.f);
        }
    }

    static HD void epilogue(const Params& p, const Blk& blk, int s, int j, int col, Regs& r, const float2* sm)
    {
        const int ch = (SRC == SRC_RAW) ? col % p.nch : 0;
        const int n2 = (SRC == SRC_RAW) ? col / p.nch : col;
        if (MODE == COL_FWD) {
            const i64 lane = (SRC == SRC_RAW) ? (i64)blk.y * p.nch + ch : (i64)blk.y;
            float2* d = p.dst + lane * p.N + n2;
            const float2* g = sm + OFF_G;
            float2 base[NBL];
#pragma unroll
            for (int b = 0; b < NBL; b++) {
                const int jb = j + S::NTS * b;
                base[b] = cis_cycles(-(double)((i64)n2 * jb) / (double)p.N);
                if (SRC == SRC_RAW && p.op == 1) base[b] = cmul(base[b], p.lcol[lane * p.N2 + n2]);
            }
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int b, int m) {
                d[(i64)o * p.N2] = cmul(x, cmul(base[b], g[m * TS + s]));
            });
        } else if (MODE == COL_INV_STORE) {
            float2* d = p.dst + (i64)blk.y * p.N + n2;
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int, int) { d[(i64)o * p.N2] = cswap(x); });
        } else {  // COL_INV_ARGMAX
            const int q = blk.y % p.nq;
            u64 best = 0;
            last_stage_visit<S>(j, r.v, [&](int o, float2 x, int, int) {
                const i64 pidx = (i64)p.nq * ((i64)o * p.N2 + n2) + q;
                u64 k = make_key(cnorm(x), (unsigned)pidx);
                best = k > best ? k : best;
            });
            r.key = best;
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
        const int s = tid % TS, j = tid / TS;
        const int col = blk.x * TS + s;
        float2* data = sm + OFF_DATA;
        auto sm_src = [&](int i) { return data[i * TS + s]; };
        auto sm_dst = [&](int i, float2 x) { data[i * TS + s] = x; };
        if (PH == 0) {
            r.key = 0;
            r.dsum[0] = 0.0;
            for (int i = tid; i < S::BLOB; i += NTA) sm[i] = p.blob[i];
            if (MODE == COL_FWD) {
                for (int t = tid; t < NG; t += NTA) {
                    const int s2 = t % TS, m = t / TS;
                    const int col2 = blk.x * TS + s2;
                    const i64 n2b = (SRC == SRC_RAW) ? col2 / p.nch : col2;
                    sm[OFF_G + m * TS + s2] = cis_cycles(-(double)(n2b * S::NSLAST * m) / (double)p.N);
                }
            }
        }
        if (PH == P_S0) {
            float mx = 0.f, my = 0.f;
            i64 lane = 0;
            if (SRC == SRC_RAW) {
                lane = (i64)blk.y * p.nch + col % p.nch;
                if (p.sub_mean) {
                    const double invn = 1.0 / (double)p.N;
                    mx = (float)(p.sums[lane * 2] * invn);
                    my = (float)(p.sums[lane * 2 + 1] * invn);
                }
            }
            stage_compute<S, 0>(j, r.v, sm, [&](int i) { return fetch(p, blk, col, i, mx, my, lane); });
            if (S::NSTAGE == 1) epilogue(p, blk, s, j, col, r, sm);
            else stage_scatter<S, 0>(j, r.v, sm_dst);
        }
        if (S::NSTAGE >= 2 && PH == 1) {
            stage_compute<S, 1>(j, r.v, sm, sm_src);
            if (S::NSTAGE == 2) epilogue(p, blk, s, j, col, r, sm);
        }
        if (S::NSTAGE == 3 && PH == 2) stage_scatter<S, 1>(j, r.v, sm_dst);
        if (S::NSTAGE == 3 && PH == 3) {
            stage_compute<S, 2>(j, r.v, sm, sm_src);
            epilogue(p, blk, s, j, col, r, sm);
        }
        if (HASFIN && PH == NCOMP) {
            if (tid == 0) p.part_key[(i64)blk.y * p.ntile + blk.x] = r.key;
        }
    }
};

}  // namespace twstft
