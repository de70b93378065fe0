// Fused correlator row pass: for one row k1 of the (transposed-order) spectrum and one lane,
//   forward row FFT  ->  P = Y*conj(C) (G) / conj(Y)*C (K), sum|Y|^2, Y[0]
//   -> for every polyphase q of the x3 interpolation: twist, inverse row FFT, conjugate four-step
//      twiddle, store B[lane][q][k1][n2].
// Replaces  ffty=fft(y); prnmap=ifft([P(1:N/2) zeros(2N) P(N/2+1:N)])  of the reference
//   processing/Octave/godual_ranging.m:25-28, acquisition/claudio_aligned_code_lo_separate.m:29,34-36
// for the row half of the four-step transforms.
//
// The product spectrum P never goes to HBM.  The last forward stage leaves thread j with the bins
// k2 = j + m*SP0, which are exactly the stage-0 operands of the inverse transforms (R0 == RLAST), so
// P is a per-thread quantity.  It has to survive three inverse transforms, but neither the register
// file (two 416-thread CTAs per SM leave 78 registers per thread) nor shared memory (the exchange
// buffers fill it) has room for it: it is parked in Blackwell's TENSOR MEMORY with tcgen05.st and
// fetched back with tcgen05.ld (12-cycle latency, per-CTA private) -- TMEM used as a software-managed
// register spill space, no tensor-core instruction involved.
#pragma once
#include "fft_core.cuh"

#ifndef TWSTFT_ABLATE_RC
#define TWSTFT_ABLATE_RC 0   // timing experiments only (results are wrong), bit mask: 1 one store per thread, 2 no loads of A
#endif

namespace twstft {

// ---- TMEM park/fetch of a thread's E complex values (device) / plain array (test emulation) ----
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void tmem_st8(unsigned a, const unsigned* s)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(s[0]), "r"(s[1]),
                 "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned a, unsigned* d)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
                 : "r"(a) : "memory");
}
// this warp's window: lanes 32*(warp%4).., columns (warp/4)*NCOL..
template <int NCOL> __device__ __forceinline__ unsigned tmem_window(unsigned base)
{
    const unsigned warp = threadIdx.x >> 5;
    return base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * NCOL;
}
template <int E> __device__ __forceinline__ void tmem_park(unsigned base, const float2* v)
{
    static_assert((2 * E) % 8 == 0, "TMEM transfers move 8 columns at a time");
    unsigned w[2 * E];
#pragma unroll
    for (int i = 0; i < E; i++) { w[2 * i] = __float_as_uint(v[i].x); w[2 * i + 1] = __float_as_uint(v[i].y); }
    const unsigned a = tmem_window<2 * E>(base);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 2 * E; c += 8) tmem_st8(a + c, w + c);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
template <int E> __device__ __forceinline__ void tmem_fetch(unsigned base, float2* v)
{
    unsigned w[2 * E];
    const unsigned a = tmem_window<2 * E>(base);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 2 * E; c += 8) tmem_ld8(a + c, w + c);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < E; i++) v[i] = make_float2(__uint_as_float(w[2 * i]), __uint_as_float(w[2 * i + 1]));
}
#endif

// chunk c (8 columns = 4 complex values) of this thread's TMEM window
#if defined(__CUDA_ARCH__)
template <int NCOL> __device__ __forceinline__ void tmem_fetch4(unsigned base, int c, float2* v4)
{
    unsigned w[8];
    tmem_ld8(tmem_window<NCOL>(base) + 8 * c, w);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; i++) v4[i] = make_float2(__uint_as_float(w[2 * i]), __uint_as_float(w[2 * i + 1]));
}
#endif

constexpr int tmem_pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

struct RowCorrParams {
    const float2* src;       // A [lane][N1][L]
    float2* dst;             // B [lane][nq][N1][L]
    const float2* blob;      // S::BLOB twiddle entries
    const float2* code;      // C [N1][L] (transposed order)
    const float2* qconst;    // [nq][R0]: exp(2 pi i q m/(nq R0)) * (m >= R0/2 ? exp(-2 pi i q/nq) : 1)
    i64 N;                   // N1 * L
    int N1;
    double invN, invM;       // 1/N, 1/(nq*N)
    int nlanes;
    int nblk;                // row groups per lane (N1 / RPC)
    int conj_mode;           // 0: Y*conj(C) (variant G)   1: conj(Y)*C (variant K)
    double* part_energy;     // [lane][nblk]  sum |Y|^2
    float2* y0;              // [lane] Y[0]
    int* counter;            // dynamic work distribution (launch.cuh): zeroed by the host before the launch
    float2* pdst;            // optional [lane][N1][L]: the product spectrum P (fast peak path, k_fast.cuh)
    double* part_penergy;    // with pdst: [lane][nblk] sum |P|^2
};

#ifndef TWSTFT_CODE_TMEM
#define TWSTFT_CODE_TMEM 1
#endif
#ifndef TWSTFT_CODE_AHEAD
#define TWSTFT_CODE_AHEAD 0     /* 1: read-only code loads issued one chunk ahead of their use (no gain measured, more spills) */
#endif

template <class S, int RPC, int NQ> struct KRowCorr {
    typedef RowCorrParams Params;
    static const char* name() { return "row_corr"; }
    // The 1x kernel (fast peak path, Doppler bank) has no spectrum to park, so tensor memory holds this thread's
    // share of the CODE row instead: a CTA walks a contiguous run of (row, lane) items, i.e. the same row for many
    // lanes in a row, and re-reads the code from global memory only when the row changes.  (With the code fetched
    // per item the 20 dependent L2 round trips of the product loop were a third of all stall samples, ncu r01 g.)
    static constexpr bool CODE_TMEM = (NQ == 1) && (TWSTFT_CODE_TMEM != 0);
    static_assert(S::NSTAGE >= 2, "row transforms have at least two stages");
    static_assert(S::R0 == S::RLAST && S::SP0 == S::NSLAST, "forward outputs must be the inverse stage-0 operands");
    static_assert(S::R0 % 2 == 0, "negative-frequency split needs an even first radix");
    static constexpr bool PERSISTENT = true;
    static constexpr int NST = S::NSTAGE;
    static constexpr int NTA = S::NTS * RPC;
    static constexpr int NT = (NTA + 31) / 32 * 32;
    static constexpr int MINB = (S::L * RPC >= 8000) ? 2 : 4;
    // TMEM: 2*E columns per warp slot, four warps (one per lane quarter) share a slot
    static constexpr int TMEM_COLS = tmem_pow2_cols(((NT / 32 + 3) / 4) * 2 * S::E);
    static_assert(((NT / 32 + 3) / 4) * 2 * S::E <= 512 / MINB, "parked spectrum does not fit the CTA's share of TMEM");
    typedef RowAcc<S::SP0> Acc;
    static constexpr int ROWSM = (S::R0 * Acc::BSTRIDE + 1) & ~1;
    static constexpr int NQC = (NQ * S::R0 + 1) & ~1;
    static constexpr int NG = RPC * S::RLAST;
    static constexpr int OFF_QC = S::BLOB_PAD;
    static constexpr int OFF_G = OFF_QC + NQC;
    static constexpr int OFF_DATA = OFF_G + ((NG + 1) & ~1);
    static constexpr int NPH = NST * (1 + NQ);
    static constexpr bool PSTORE = (NQ == 1);     // the product spectrum can be stored by the 1x kernel only
    static constexpr int NSUM = PSTORE ? 2 : 1;
    static constexpr int NBL = (S::L / S::RLAST) / S::NTS;
    static constexpr int NB0 = S::SP0 / S::NTS;
    static_assert(NBL == NB0, "butterflies per thread must match between the first and the last stage");
    struct Regs {
        float2 v[S::E];
        u64 key;
        double dsum[NSUM];
        unsigned tmem;       // TMEM base address of this CTA (device)
        float2 obase[(S::L / S::RLAST) / S::NTS];   // four-step twiddle base of this thread's outputs, per item
        int crow = -1;       // CODE_TMEM: row whose code values sit in this thread's TMEM window (-1: none)
#ifdef TWSTFT_EMU
        float2 P[S::E];      // host emulation: the parked spectrum / code row
#endif
    };
    static HD constexpr size_t smem_bytes()
    {
        return (size_t)(OFF_DATA + RPC * ROWSM) * sizeof(float2) + RED_SCRATCH_BYTES;
    }
    static HD constexpr int sync_after(int ph) { return ph == NST - 1 ? SYNC_SUMD : SYNC_BARRIER; }

    // work item w = (rowblk = w / nlanes, lane = w % nlanes): CTAs that run together share the code row in L2
#ifndef TWSTFT_ROWCORR_DYNAMIC
#define TWSTFT_ROWCORR_DYNAMIC 1
#endif
    // items in (row, lane) order; a CTA grabs CHUNK consecutive ones at a time (launch.cuh, dynamic distribution): the
    // lanes of one row stay together, so the code row in tensor memory is re-read once per chunk at most
#ifndef TWSTFT_ROWCORR_CHUNK
#define TWSTFT_ROWCORR_CHUNK 8
#endif
    static constexpr int CHUNK = TWSTFT_ROWCORR_CHUNK;
    static HD int total_items(const Params& p) { return p.nblk * p.nlanes; }
    static HD int* counter(const Params& p) { return p.counter; }
    static HD int niter(const Params&, int, int) { return 0; }
    struct Work { int rowblk; i64 lane; };
    static HD Work work_of(const Params& p, const Blk&, int it)
    {
        const i64 w = it;
        Work k;
        k.rowblk = (int)(w / p.nlanes);
        k.lane = w % p.nlanes;
        return k;
    }

    static HD void park(Regs& r)
    {
#if defined(__CUDA_ARCH__)
        tmem_park<S::E>(r.tmem, r.v);
#elif defined(TWSTFT_EMU)
        for (int i = 0; i < S::E; i++) r.P[i] = r.v[i];
#endif
    }
    static HD void fetch(Regs& r)
    {
#if defined(__CUDA_ARCH__)
        tmem_fetch<S::E>(r.tmem, r.v);
#elif defined(TWSTFT_EMU)
        for (int i = 0; i < S::E; i++) r.v[i] = r.P[i];
#endif
    }

    // four code values (elements 4c .. 4c+3 of this thread) out of tensor memory; warp-collective on the device
    static HD void code4(const Regs& r, int c, float2* cc)
    {
#if defined(__CUDA_ARCH__)
        tmem_fetch4<2 * S::E>(r.tmem, c, cc);
#elif defined(TWSTFT_EMU)
        for (int i = 0; i < 4; i++) cc[i] = r.P[4 * c + i];
#endif
    }
    static HD float2 ldg_ro(const float2* g)
    {
#if defined(__CUDA_ARCH__)
        return __ldg(g);
#else
        return *g;
#endif
    }

    static HD void prologue(const Params& p, int tid, unsigned char* smem)
    {
        float2* sm = reinterpret_cast<float2*>(smem);
        for (int i = tid; i < S::BLOB; i += NT) sm[i] = p.blob[i];
        for (int i = tid; i < NQ * S::R0; i += NT) sm[OFF_QC + i] = p.qconst[i];
    }

    template <int PH>
    static HD void phase(const Params& p, const Blk& blk, int tid, Regs& r, unsigned char* smem, int it, int nit)
    {
        constexpr int F = PH / NST, ST = PH % NST;     // F = 0 forward, F = q+1 inverse
        float2* sm = reinterpret_cast<float2*>(smem);
        const bool active = tid < NTA;                 // the TMEM transfers below are warp-wide: idle threads take part
        const int rl = active ? tid / S::NTS : 0, j = tid % S::NTS;
        const Acc acc{sm + OFF_DATA + rl * ROWSM};
        const Work wk = work_of(p, blk, it);
        const int row = wk.rowblk * RPC + rl;
        float2* gtab = sm + OFF_G;

        if (PH == 0) {
            r.key = 0;
            r.dsum[0] = 0.0;
            if (PSTORE) r.dsum[NSUM - 1] = 0.0;
            if (CODE_TMEM) {
                if (r.crow != row) {                      // uniform over the CTA's rows: every thread of a warp takes part
                    if (active) {
                        const float2* c = p.code + (i64)row * S::L + j;
#pragma unroll
                        for (int b = 0; b < NB0; b++)
#pragma unroll
                            for (int m = 0; m < S::R0; m++) r.v[b * S::R0 + m] = ldg_ro(c + S::NTS * b + m * S::SP0);
                    }
                    park(r);
                    r.crow = row;
                }
            }
            if (active) {
                for (int t = tid; t < NG; t += NTA) {      // conjugate four-step twiddle steps of this item's rows
                    const int rl2 = t / S::RLAST, m = t % S::RLAST;
                    const i64 row2 = (i64)wk.rowblk * RPC + rl2;
                    gtab[t] = cis_cycles((double)(row2 * S::NSLAST * m) * p.invN);
                }
                const float2* xj = p.src + wk.lane * p.N + (i64)row * S::L + j;
#pragma unroll
                for (int b = 0; b < NB0; b++)
#pragma unroll
                    for (int m = 0; m < S::R0; m++)
                        r.v[b * S::R0 + m] = ((TWSTFT_ABLATE_RC & 2) && p.nlanes > 0) ? make_float2((float)(j + m), (float)(row - m))
                                                                                    : xj[S::NTS * b + m * S::SP0];
                dif_stage0<S>(j, r.v, sm, acc);
            }
        }
        if (F == (NQ > 1 ? 2 : 1) && ST == 0 && it + 1 < nit) {
            // next item's spectrum rows and code rows towards L2 while this one is still being transformed
            const Work nx = work_of(p, blk, it + 1);
            const char* a = reinterpret_cast<const char*>(p.src + nx.lane * p.N + (i64)nx.rowblk * RPC * S::L);
            const char* c = reinterpret_cast<const char*>(p.code + (i64)nx.rowblk * RPC * S::L);
            constexpr int NLINE = (int)((size_t)RPC * S::L * sizeof(float2) / 128);
            for (int l = tid; l < NLINE; l += NT) {
                prefetch_l2(a + (size_t)l * 128);
                if (!CODE_TMEM) prefetch_l2(c + (size_t)l * 128);
            }
        }
        if (F > 0 && ST == 0) {
            constexpr int q = F - 1;
            if (q > 0) fetch(r);                          // q == 0: P is still in v
            if (active) {
                if (q == 0 && tid == 0) {
                    p.part_energy[wk.lane * p.nblk + wk.rowblk] = r.dsum[0];
                    if (PSTORE && p.pdst) p.part_penergy[wk.lane * p.nblk + wk.rowblk] = r.dsum[NSUM - 1];
                }
#pragma unroll
                for (int b = 0; b < NB0; b++) {
                    if (q == 0) {
#pragma unroll
                        for (int m = 0; m < S::R0; m++) r.v[b * S::R0 + m] = cswap(r.v[b * S::R0 + m]);
                    } else {
                        const int jb = j + S::NTS * b;
                        const float2 base = cis_cycles((double)((i64)q * ((i64)row + (i64)p.N1 * jb)) * p.invM);
                        const float2* qc = sm + OFF_QC + q * S::R0;
#pragma unroll
                        for (int m = 0; m < S::R0; m++)
                            r.v[b * S::R0 + m] = cswap(cmul(r.v[b * S::R0 + m], cmul(base, qc[m])));
                    }
                }
                dif_stage0<S>(j, r.v, sm, acc);
            }
        }
        if (NST == 3 && ST == 1 && active) dif_stage_mid<S>(j, r.v, sm, acc);
        if (ST == NST - 1 && F == 0) {
            // forward last stage, then the code product four outputs at a time: the code values come out of tensor
            // memory (CODE_TMEM) or from read-only loads issued a chunk ahead of their use
            static_assert(S::E % 4 == 0, "code product walks the outputs in chunks of four");
            if (active) {
                dif_stage_last<S>(j, r.v, acc);
                if (row == 0 && j == 0) p.y0[wk.lane] = r.v[0];
            }
            float e = 0.f, e2 = 0.f;
            const int cm = p.conj_mode;
            float2* pd = (PSTORE && p.pdst) ? p.pdst + wk.lane * p.N + (i64)row * S::L : nullptr;
            const float2* c = p.code + (i64)row * S::L;
            auto out_index = [&](int el) { return j + S::NTS * (el / S::RLAST) + (el % S::RLAST) * S::NSLAST; };
            constexpr bool AHEAD = TWSTFT_CODE_AHEAD != 0;
            float2 cn[4];
            if (!CODE_TMEM && AHEAD && active) {
#pragma unroll
                for (int i = 0; i < 4; i++) cn[i] = ldg_ro(c + out_index(i));
            }
#pragma unroll
            for (int c4 = 0; c4 < S::E / 4; c4++) {
                float2 cc[4];
                if (CODE_TMEM) {
                    code4(r, c4, cc);
                } else if (AHEAD) {
#pragma unroll
                    for (int i = 0; i < 4; i++) cc[i] = cn[i];
                    if (active && c4 + 1 < S::E / 4) {
#pragma unroll
                        for (int i = 0; i < 4; i++) cn[i] = ldg_ro(c + out_index(4 * (c4 + 1) + i));
                    }
                } else if (active) {
#pragma unroll
                    for (int i = 0; i < 4; i++) cc[i] = ldg_ro(c + out_index(4 * c4 + i));
                }
                if (active) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int el = 4 * c4 + i;
                        const float2 x = r.v[el];
                        e += cnorm(x);
                        const float2 pr = cm ? cmul(cconj(x), cc[i]) : cmulc(x, cc[i]);
                        r.v[el] = pr;
                        if (PSTORE && pd) {
                            if (TWSTFT_ABLATE_RC & 1) { if (el == 0) pd[out_index(el)] = pr; }
                            else pd[out_index(el)] = pr;
                            e2 += cnorm(pr);
                        }
                    }
                }
            }
            if (active) {
                r.dsum[0] = (double)e;
                if (PSTORE) r.dsum[NSUM - 1] = (double)e2;
            }
            if (NQ > 1) park(r);                          // P -> tensor memory for the polyphases q >= 1
        }
        if (ST == NST - 1 && F > 0) {
            if (active) {
                dif_stage_last<S>(j, r.v, acc);
                {
                    constexpr int q = F - 1;
                    float2* d = p.dst + (wk.lane * NQ + q) * p.N + (i64)row * S::L;
                    const float2* gr = gtab + rl * S::RLAST;
                    if (q == 0) {                         // the same for every polyphase of the item: computed once
#pragma unroll
                        for (int b = 0; b < NBL; b++)
                            r.obase[b] = cis_cycles((double)((i64)row * (j + S::NTS * b)) * p.invN);
                    }
                    float2 ablsum = make_float2(0.f, 0.f);
                    last_stage_visit<S>(j, r.v, [&](int o, float2 x, int b, int m) {
                        const float2 y = cmul(cswap(x), cmul(r.obase[b], gr[m]));
                        if (TWSTFT_ABLATE_RC & 1) ablsum = cadd(ablsum, y);
                        else d[o] = y;
                    });
                    if (TWSTFT_ABLATE_RC & 1) d[j] = ablsum;
                }
            }
        }
    }
};

}  // namespace twstft
