// Synthetic-input tooling (recipe of experiments/221127_SNR/simu_snr.m: code + AWGN at a
// chosen SNR): PRN-modulated noisy int16 records generated ON THE DEVICE, so that the benchmark can hold a pool of
// many distinct records (and a 5.5-min single-channel file) without minutes of host-side numpy.  Not on the hot
// path; the parity tests use the host generator (amaranth_twstft_b200/synth.py) and the oracle.
//   x[n] = amp * code[(n - delay) mod N] * exp(j 2 pi (phase + dphi n)) + sigma * (g1 + j g2) + dc   -> round, saturate
// Noise: Philox4x32-10 keyed by the record seed, counter = (sample, lane), Box-Muller.
#pragma once
#include "launch.cuh"

namespace twstft {

struct SynthLane {
    i64 delay;             // integer samples, 0 <= delay < period
    double dphi, phase;    // cycles per sample, cycles
    float amp, sigma, dc_re, dc_im;
    unsigned key0, key1;
};

struct SynthParams {
    const unsigned char* bits;   // chips, 0/1
    i64 period;                  // code period in samples (= nchips * rep)
    int rep;
    int nch;
    i64 nsamp;                   // sample instants per record
    const SynthLane* lanes;      // [record*nch + ch]
    short2* out;                 // [record][nsamp][nch]
};

HD void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* o)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const u64 p0 = (u64)0xD2511F53u * c0, p1 = (u64)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

struct KSynth {
    typedef SynthParams Params;
    static const char* name() { return "synth"; }
    static constexpr bool PERSISTENT = false;
    static constexpr int NT = 256, MINB = 4, NPH = 1, NSUM = 1;
    struct Regs { u64 key; double dsum[NSUM]; };
    static HD constexpr size_t smem_bytes() { return RED_SCRATCH_BYTES; }
    static HD constexpr int sync_after(int) { return SYNC_BARRIER; }
    static HD short sat16(float v)
    {
        v = rintf(v);
        v = v > 32767.f ? 32767.f : (v < -32768.f ? -32768.f : v);
        return (short)v;
    }
    template <int PH> static HD void phase(const Params& p, const Blk& blk, int tid, Regs&, unsigned char*)
    {
        const i64 lane = blk.y;                               // record*nch + ch
        const SynthLane L = p.lanes[lane];
        const i64 rec = lane / p.nch;
        const int ch = (int)(lane % p.nch);
        for (i64 n = (i64)blk.x * NT + tid; n < p.nsamp; n += (i64)blk.gx * NT) {
            i64 m = (n - L.delay) % p.period;
            if (m < 0) m += p.period;
            const float c = p.bits[m / p.rep] ? 1.f : -1.f;
            const float2 w = cis_cycles(L.phase + L.dphi * (double)n);
            unsigned u[4];
            philox4x32_10((unsigned)n, (unsigned)((u64)n >> 32), (unsigned)lane, 0u, L.key0, L.key1, u);
            const float u1 = ((float)(u[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float u2 = ((float)(u[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float rr = L.sigma * sqrtf(-2.0f * logf(u1));
            float sn, cs;
#ifdef __CUDA_ARCH__
            sincospif(2.0f * u2, &sn, &cs);
#else
            sn = sinf(6.2831853f * u2); cs = cosf(6.2831853f * u2);
#endif
            short2 o;
            o.x = sat16(L.amp * c * w.x + rr * cs + L.dc_re);
            o.y = sat16(L.amp * c * w.y + rr * sn + L.dc_im);
            p.out[(rec * p.nsamp + n) * p.nch + ch] = o;
        }
    }
};

}  // namespace twstft
