// C ABI (include/twstft.h) over the kernel pipeline.  Host-side logic only: configuration,
// window arithmetic identical to the reference's linspace/find, batching, copies.
#include "../../include/twstft.h"
#include "pipeline.cuh"
#include <string>
#include <vector>
#include <algorithm>
#include <thread>
#include <chrono>
#include <stdio.h>
#ifndef TWSTFT_EMU
#include <sched.h>
#endif

using namespace twstft;

static_assert(sizeof(twstft_result) == sizeof(Result), "public result row must match the device row");
static_assert(sizeof(twstft_bank_result) == sizeof(BankRow), "public bank row must match the device row");

struct twstft_plan {
    twstft_config cfg;
    Backend be;
    IPipe* pipe = nullptr;
    EstimatorCfg est;
    EstimatorCfg est2;             // the same estimator on blocks of 2N samples (ls = 2)
    int nch = 2;
    int nq = 3;
    int max_lanes = 0;
    bool have_code = false;
    int code_qpsk = 0;
    size_t code_nchips = 0;
    short2* d_raw = nullptr;       // staging for host inputs: two slots of max_batch records
    size_t raw_elems_per_rec = 0;
    Result* d_res = nullptr;
    size_t res_cap = 0;            // rows d_res can hold
    double* d_dfin = nullptr;
    i64* d_off = nullptr;
    unsigned char* d_bits = nullptr;
    size_t bits_cap = 0;
    long long fast_records = 0, fast_rerun = 0;   // records through the fast peak path / re-run through the full path
    // governor of the fast peak path: an uncertified record costs a fast attempt PLUS the full path, so the path only
    // pays while less than about a quarter of the records fall back.  Over every window of FAST_WIN records the
    // fall-back rate is measured; above FAST_MAX_RATE the plan runs the full path for the next FAST_OFF records and
    // then probes again.  Results are identical either way.
    long long fast_win_seen = 0, fast_win_rerun = 0, fast_off_left = 0, fast_suspended = 0;
    long long fast_win = 64, fast_off = 1024;     // TWSTFT_FAST_WIN / TWSTFT_FAST_OFF override them (tests)
    short2* d_winbuf = nullptr;    // staging of twstft_process_windows' host buffers (grows, never shrinks)
    size_t win_cap = 0;
    double2* d_cf64 = nullptr;     // upload staging of the complex-double entry points (one chunk)
    float2* d_cf32 = nullptr;      // their windows as complex float
    size_t cf64_cap = 0, cf32_cap = 0;   // samples
    std::string err;
};

static int fail(twstft_plan* p, int code, const std::string& msg)
{
    if (p) p->err = msg;
    return code;
}

static double axis_value(const EstimatorCfg& e, i64 s)
{
    if (s == e.naxis - 1) return e.flast;
    volatile double prod = (double)s * e.fstep;      // two roundings, like numpy/Octave linspace
    volatile double v = prod + e.f0;
    return v;
}

// k=find((freq<hi)&(freq>lo)) on freq=linspace(...): returns the inclusive index range
static void window_bins(const EstimatorCfg& e, double lo, double hi, i64* slo, i64* shi)
{
    i64 a = (i64)floor((lo - e.f0) / e.fstep) - 3;
    if (a < 0) a = 0;
    while (a < e.naxis && !(axis_value(e, a) > lo)) a++;
    i64 b = (i64)ceil((hi - e.f0) / e.fstep) + 3;
    if (b > e.naxis - 1) b = e.naxis - 1;
    while (b >= 0 && !(axis_value(e, b) < hi)) b--;
    *slo = a;
    *shi = b;
}

extern "C" {

const char* twstft_backend(void) { return Backend::is_emulation ? "emu" : "cuda"; }

void twstft_default_config(twstft_config* c)
{
    memset(c, 0, sizeof(*c));
    c->fs = 5e6;
    c->nsamples = 5000000;
    c->samples_per_chip = 2;
    c->variant = TWSTFT_VARIANT_G;
    c->modulation = TWSTFT_BPSK;
    c->nint = 1;
    c->layout = TWSTFT_LAYOUT_INTERLEAVED2;
    c->freq_axis = TWSTFT_AXIS_G;
    c->win_lo_hz = -20000.0;
    c->win_hi_hz = 20000.0;
    c->est_sub_mean = 1;
    c->ddof = 1;
    c->max_batch = 4;
    c->device_id = 0;
    c->fine_freq = 0;
    c->code_window = TWSTFT_CODEWIN_NONE;
    c->fast_peak = 0;
}

int twstft_plan_create(const twstft_config* cfg, twstft_plan** out)
{
    if (!cfg || !out) return TWSTFT_E_ARG;
    *out = nullptr;
    twstft_plan* p = new twstft_plan();
    p->cfg = *cfg;
    *out = p;   // returned even on failure so that twstft_last_error can be read; destroy it either way
    if (cfg->nsamples <= 0 || cfg->max_batch <= 0 || cfg->fs <= 0 || cfg->samples_per_chip <= 0)
        return fail(p, TWSTFT_E_ARG, "bad configuration");
    if (cfg->code_window != TWSTFT_CODEWIN_NONE && cfg->code_window != TWSTFT_CODEWIN_HAMMING)
        return fail(p, TWSTFT_E_ARG, "unknown code_window");
    if (cfg->nint != 0 && cfg->nint != 1)
        return fail(p, TWSTFT_E_ARG, "only Nint=1 (x3) and Nint=0 are supported (the reference uses Nint=1)");
    if (p->be.init(cfg->device_id) != 0)
        return fail(p, TWSTFT_E_NODEVICE, std::string("no usable CUDA device: ") + p->be.err_string() +
                                              " (this library has no CPU fallback)");
    p->pipe = make_pipe(cfg->nsamples);
    if (!p->pipe) return fail(p, TWSTFT_E_SIZE, "unsupported window length (see make_pipe in pipeline.cuh)");
    p->nch = cfg->layout == TWSTFT_LAYOUT_INTERLEAVED2 ? 2 : 1;
    p->nq = cfg->nint ? 3 : 1;
    p->max_lanes = cfg->max_batch * p->nch;
    if (p->pipe->setup(&p->be, p->max_lanes, p->nq) != 0)
        return fail(p, TWSTFT_E_NOMEM, std::string("device allocation failed: ") + p->be.err_string());
    if (getenv("TWSTFT_PROFILE_DUMP")) p->be.profiling = true;   // CLIs: print per-kernel totals when the plan is destroyed
    if (const char* e = getenv("TWSTFT_FAST_WIN")) { if (atoll(e) > 0) p->fast_win = atoll(e); }
    if (const char* e = getenv("TWSTFT_FAST_OFF")) { if (atoll(e) > 0) p->fast_off = atoll(e); }
    if (cfg->fast_peak && p->pipe->set_fast_peak(1) != 0)
        return fail(p, TWSTFT_E_NOMEM, std::string("device allocation failed (fast peak path): ") + p->be.err_string());
    const i64 N = cfg->nsamples;
    p->raw_elems_per_rec = (size_t)N * p->nch;
    p->d_res = (Result*)p->be.alloc(sizeof(Result) * p->max_lanes);
    p->res_cap = (size_t)p->max_lanes;
    p->d_dfin = (double*)p->be.alloc(sizeof(double) * p->max_lanes);
    p->d_off = (i64*)p->be.alloc(sizeof(i64) * cfg->max_batch);
    // frequency axis of the estimator: length = window length
    EstimatorCfg& e = p->est;
    e.naxis = N;
    e.f0 = -cfg->fs / 2;
    e.flast = cfg->freq_axis == TWSTFT_AXIS_G ? cfg->fs / 2 : cfg->fs / 2 - 1.0;
    e.fstep = (e.flast - e.f0) / (double)(N - 1);
    e.power = cfg->modulation == TWSTFT_QPSK ? 4 : 2;
    e.sub_mean = cfg->est_sub_mean;
    window_bins(e, cfg->win_lo_hz, cfg->win_hi_hz, &e.win_lo, &e.win_hi);
    if (e.win_lo > e.win_hi) return fail(p, TWSTFT_E_ARG, "empty estimator window");
    EstimatorCfg& e2 = p->est2;     // freq=linspace(-fs/2,fs/2-fs/fs,fs*ls), ls=2 (lo_separate.m:103)
    e2 = e;
    e2.naxis = 2 * N;
    e2.fstep = (e2.flast - e2.f0) / (double)(2 * N - 1);
    window_bins(e2, cfg->win_lo_hz, cfg->win_hi_hz, &e2.win_lo, &e2.win_hi);
    if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    return TWSTFT_OK;
}

void twstft_plan_destroy(twstft_plan* p)
{
    if (!p) return;
    p->be.use();
    if (p->be.profiling && getenv("TWSTFT_PROFILE_DUMP")) {     // per-kernel CUDA-event totals of this plan's lifetime
        p->be.prof_collect();
        fprintf(stderr, "twstft_profile N=%lld {", (long long)p->cfg.nsamples);
        for (size_t i = 0; i < p->be.prof.size(); i++)
            fprintf(stderr, "%s\"%s\": [%.4f, %ld]", i ? ", " : "", p->be.prof[i].name, p->be.prof[i].ms, p->be.prof[i].count);
        fprintf(stderr, "}\n");
    }
    delete p->pipe;
    p->be.free(p->d_raw);
    p->be.free(p->d_res);
    p->be.free(p->d_dfin);
    p->be.free(p->d_off);
    p->be.free(p->d_bits);
    p->be.free(p->d_cf64);
    p->be.free(p->d_cf32);
    p->be.free(p->d_winbuf);
    p->be.shutdown();
    delete p;
}

const char* twstft_last_error(const twstft_plan* p) { return p ? p->err.c_str() : "null plan"; }

int twstft_fast_counters(const twstft_plan* p, int64_t* records, int64_t* rerun)
{
    if (!p) return TWSTFT_E_ARG;
    if (records) *records = p->fast_records;
    if (rerun) *rerun = p->fast_rerun;
    return TWSTFT_OK;
}

int64_t twstft_fast_suspended(const twstft_plan* p) { return p ? p->fast_suspended : 0; }

int64_t twstft_kernel_launches(const twstft_plan* p) { return p ? p->be.launches : 0; }

int twstft_window_bins(const twstft_plan* p, int64_t* lo, int64_t* hi)
{
    if (!p) return TWSTFT_E_ARG;
    if (lo) *lo = p->est.win_lo;
    if (hi) *hi = p->est.win_hi;
    return TWSTFT_OK;
}

static int ensure_bits(twstft_plan* p, size_t n)
{
    if (p->bits_cap >= n) return 0;
    p->be.free(p->d_bits);
    p->d_bits = (unsigned char*)p->be.alloc(n);
    p->bits_cap = p->d_bits ? n : 0;
    return (p->d_bits && p->be.ok()) ? 0 : -1;
}

static int install_code(twstft_plan* p, const unsigned char* host_bits, size_t nbytes, int qpsk)
{
    const size_t nchips = qpsk ? nbytes / 2 : nbytes;
    if ((i64)nchips * p->cfg.samples_per_chip != p->cfg.nsamples)
        return fail(p, TWSTFT_E_ARG, "code length * samples_per_chip != nsamples");
    float2 mean = make_float2(0.f, 0.f);
    double code_pow = 1.0;
    if (qpsk) {   // goqpsk.m:11-12 codec=codei+j*codeq; codec=codec-mean(codec)
        double sa = 0, sb = 0;
        for (size_t i = 0; i < nchips; i++) { sa += host_bits[2 * i]; sb += host_bits[2 * i + 1]; }
        const double ma = sa / (double)nchips, mb = sb / (double)nchips;
        mean = make_float2((float)ma, (float)mb);
        // mean |codec|^2 with bits in {0,1}: E[(a-ma)^2] + E[(b-mb)^2] = ma(1-ma) + mb(1-mb)
        code_pow = ma * (1.0 - ma) + mb * (1.0 - mb);
    }
    p->pipe->set_code_window(p->cfg.code_window == TWSTFT_CODEWIN_HAMMING);
    if (p->pipe->build_code(p->d_bits, p->cfg.samples_per_chip, qpsk, mean, code_pow) != 0)
        return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    p->be.sync();
    if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    p->have_code = true;
    p->code_qpsk = qpsk;
    p->code_nchips = nchips;
    return TWSTFT_OK;
}

int twstft_set_code(twstft_plan* p, const uint8_t* chips, size_t nbytes, int qpsk)
{
    if (!p || !chips || !p->pipe) return TWSTFT_E_ARG;
    p->be.use();
    if (ensure_bits(p, nbytes) != 0) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    p->be.h2d(p->d_bits, chips, nbytes);
    p->be.sync();
    return install_code(p, chips, nbytes, qpsk);
}

// GF(2) jump matrices of the LFSR step: column i of M^(2^j)
static void lfsr_jump_tables(int bitlen, unsigned taps, unsigned jump[24][24])
{
    auto step = [&](unsigned s) {
        unsigned x = s & taps;
        x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
        return (s >> 1) | ((x & 1u) << (bitlen - 1));
    };
    memset(jump, 0, sizeof(unsigned) * 24 * 24);
    for (int i = 0; i < bitlen; i++) jump[0][i] = step(1u << i);
    for (int j = 1; j < 24; j++)
        for (int i = 0; i < bitlen; i++) {
            unsigned v = jump[j - 1][i], r = 0;
            for (int b = 0; b < bitlen; b++)
                if ((v >> b) & 1) r ^= jump[j - 1][b];
            jump[j][i] = r;
        }
}

static int prn_on_device(Backend& be, unsigned char* d_out, int bitlen, unsigned taps, unsigned seed,
                         size_t noiselen, int ostride, int ooffset)
{
    PrnParams pp;
    lfsr_jump_tables(bitlen, taps, pp.jump);
    pp.bitlen = bitlen; pp.taps = taps; pp.seed = seed; pp.nchips = (i64)noiselen; pp.out = d_out;
    pp.ostride = ostride; pp.ooffset = ooffset;
    const i64 nthreads = ((i64)noiselen + KPrnGen::CH - 1) / KPrnGen::CH;
    const int nblk = (int)((nthreads + KPrnGen::NT - 1) / KPrnGen::NT);
    be.launch<KPrnGen>(dim3(nblk, 1, 1), pp);
    return be.ok() ? 0 : -1;
}

int twstft_gen_code(twstft_plan* p, int bitlen, uint32_t taps_a, uint32_t taps_b, uint32_t seed,
                    size_t noiselen, uint8_t* chips_out)
{
    if (!p || !p->pipe || bitlen < 2 || bitlen > 24 || noiselen == 0 || noiselen >= (1u << 24))
        return fail(p, TWSTFT_E_ARG, "bad LFSR parameters");
    p->be.use();
    const int qpsk = taps_b != 0;
    const size_t nbytes = noiselen * (qpsk ? 2 : 1);
    if (ensure_bits(p, nbytes) != 0) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    prn_on_device(p->be, p->d_bits, bitlen, taps_a, seed, noiselen, qpsk ? 2 : 1, 0);
    if (qpsk) prn_on_device(p->be, p->d_bits, bitlen, taps_b, seed, noiselen, 2, 1);
    std::vector<unsigned char> host(nbytes);
    p->be.d2h(host.data(), p->d_bits, nbytes);
    p->be.sync();
    if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    if (chips_out) memcpy(chips_out, host.data(), nbytes);
    return install_code(p, host.data(), nbytes, qpsk);
}

int twstft_prn_generate(int device_id, int bitlen, uint32_t taps_a, uint32_t taps_b, uint32_t seed,
                        size_t noiselen, uint8_t* out)
{
    if (!out || bitlen < 2 || bitlen > 24 || noiselen == 0 || noiselen >= (1u << 24)) return TWSTFT_E_ARG;
    Backend be;
    if (be.init(device_id) != 0) return TWSTFT_E_NODEVICE;
    const int qpsk = taps_b != 0;
    const size_t nbytes = noiselen * (qpsk ? 2 : 1);
    unsigned char* d = (unsigned char*)be.alloc(nbytes);
    prn_on_device(be, d, bitlen, taps_a, seed, noiselen, qpsk ? 2 : 1, 0);
    if (qpsk) prn_on_device(be, d, bitlen, taps_b, seed, noiselen, 2, 1);
    be.d2h(out, d, nbytes);
    be.sync();
    be.free(d);
    const bool ok = be.ok();
    be.shutdown();
    return ok ? TWSTFT_OK : TWSTFT_E_CUDA;
}

static int ensure_raw(twstft_plan* p)
{
    if (p->d_raw) return 0;
    p->d_raw = (short2*)p->be.alloc(sizeof(short2) * p->raw_elems_per_rec * p->cfg.max_batch * 2);
    return (p->d_raw && p->be.ok()) ? 0 : -1;
}

static constexpr double FAST_MAX_RATE = 0.25;

static bool fast_on(const twstft_plan* p) { return p->cfg.fast_peak && p->fast_off_left <= 0; }

static void fast_account(twstft_plan* p, size_t n, size_t nrerun, bool was_fast)
{
    if (!p->cfg.fast_peak) return;
    if (!was_fast) { p->fast_off_left -= (long long)n; p->fast_suspended += (long long)n; return; }
    p->fast_records += (long long)n;
    p->fast_rerun += (long long)nrerun;
    p->fast_win_seen += (long long)n;
    p->fast_win_rerun += (long long)nrerun;
    if (p->fast_win_seen >= p->fast_win) {
        if ((double)p->fast_win_rerun > FAST_MAX_RATE * (double)p->fast_win_seen) p->fast_off_left = p->fast_off;
        p->fast_win_seen = p->fast_win_rerun = 0;
    }
}

// Fast peak path (k_fast.cuh): rows whose certificate failed carry status bit 2; those records go through the full
// three-polyphase path here, ALL FLAGGED RECORDS OF THE CALL IN BATCHES (one launch sequence per max_batch of them,
// addressed through the window-offset table when they sit in device memory, copied into the staging slots when they
// sit in host memory).  `src` host or device; returns the number of re-run records through *nrerun.
static int rerun_uncertified(twstft_plan* p, const short2* src, bool on_device, size_t nrecords, twstft_result* results,
                             size_t* nrerun)
{
    std::vector<size_t> flagged;
    for (size_t r = 0; r < nrecords; r++) {
        bool f = false;
        for (int c = 0; c < p->nch; c++) f = f || (((int)results[r * p->nch + c].status) & 4);
        if (f) flagged.push_back(r);
    }
    if (nrerun) *nrerun = flagged.size();
    const size_t B = (size_t)p->cfg.max_batch;
    std::vector<i64> off(B);
    std::vector<twstft_result> rows(B * p->nch);
    for (size_t f0 = 0; f0 < flagged.size(); f0 += B) {
        const int nb = (int)std::min(B, flagged.size() - f0);
        RunArgs a{};
        if (on_device) {
            for (int i = 0; i < nb; i++) off[i] = (i64)(flagged[f0 + i] * p->raw_elems_per_rec);
            p->be.h2d(p->d_off, off.data(), sizeof(i64) * nb);
            a.d_raw = src; a.rec_stride = 0; a.d_rec_offset = p->d_off;
        } else {
            if (ensure_raw(p) != 0) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
            for (int i = 0; i < nb; i++)
                p->be.h2d(p->d_raw + (size_t)i * p->raw_elems_per_rec, src + flagged[f0 + i] * p->raw_elems_per_rec,
                          sizeof(short2) * p->raw_elems_per_rec);
            a.d_raw = p->d_raw; a.rec_stride = (i64)p->raw_elems_per_rec; a.d_rec_offset = nullptr;
        }
        a.nrec = nb; a.nch = p->nch;
        a.estimate = 1; a.variant = p->cfg.variant; a.ddof = p->cfg.ddof; a.fs = p->cfg.fs; a.d_out = p->d_res;
        a.allow_fast = 0;
        if (p->pipe->estimate(a, p->est) != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        if (p->cfg.fine_freq && p->pipe->fine_frequency(a) != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        if (p->pipe->correlate(a) != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        p->be.d2h(rows.data(), p->d_res, sizeof(Result) * nb * p->nch);
        p->be.sync();
        if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        for (int i = 0; i < nb; i++)
            memcpy(results + flagged[f0 + i] * p->nch, rows.data() + (size_t)i * p->nch, sizeof(twstft_result) * p->nch);
    }
    return TWSTFT_OK;
}

// records resident on the device, contiguous
static int process_dev(twstft_plan* p, const short2* d_iq, size_t nrecords, twstft_result* results)
{
    const int B = p->cfg.max_batch;
    const bool fast = fast_on(p);
    for (size_t r0 = 0; r0 < nrecords; r0 += B) {
        const int nb = (int)((nrecords - r0) < (size_t)B ? (nrecords - r0) : B);
        RunArgs a{};
        a.d_raw = d_iq + r0 * p->raw_elems_per_rec;
        a.nrec = nb; a.nch = p->nch; a.rec_stride = (i64)p->raw_elems_per_rec; a.d_rec_offset = nullptr;
        a.estimate = 1; a.variant = p->cfg.variant; a.ddof = p->cfg.ddof; a.fs = p->cfg.fs; a.d_out = p->d_res;
        a.allow_fast = fast;
        if (p->pipe->estimate(a, p->est) != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        if (p->cfg.fine_freq) {
            const int frc = p->pipe->fine_frequency(a);
            if (frc == -3) return fail(p, TWSTFT_E_ARG, "fine_freq: the window is shorter than fs/3 (the reference's polyfit raises)");
            if (frc != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        }
        if (p->pipe->correlate(a) != 0)
            return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        p->be.d2h(results + r0 * p->nch, p->d_res, sizeof(Result) * nb * p->nch);
        p->be.sync();
        if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    }
    size_t nrerun = 0;
    if (fast) {
        const int rc = rerun_uncertified(p, d_iq, true, nrecords, results, &nrerun);
        if (rc != TWSTFT_OK) return rc;
    }
    fast_account(p, nrecords, nrerun, fast);
    return TWSTFT_OK;
}

int twstft_process_device(twstft_plan* p, const int16_t* iq_dev, size_t nrecords, twstft_result* results)
{
    if (!p || !p->pipe || !iq_dev || !results) return TWSTFT_E_ARG;
    if (!p->have_code) return fail(p, TWSTFT_E_NOCODE, "no PRN code installed (twstft_set_code / twstft_gen_code)");
    p->be.use();
    return process_dev(p, reinterpret_cast<const short2*>(iq_dev), nrecords, results);
}

static int ensure_results(twstft_plan* p, size_t rows)
{
    if (p->res_cap >= rows) return 0;
    p->be.sync();
    p->be.free(p->d_res);
    p->d_res = (Result*)p->be.alloc(sizeof(Result) * rows);
    p->res_cap = p->d_res ? rows : 0;
    return (p->d_res && p->be.ok()) ? 0 : -1;
}

// Host records: the input is streamed in chunks through two staging slots so that the PCIe copy of
// chunk c+1 overlaps the kernels of chunk c; one result copy at the end.
int twstft_process(twstft_plan* p, const int16_t* iq, size_t nrecords, twstft_result* results)
{
    if (!p || !p->pipe || !iq || !results) return TWSTFT_E_ARG;
    if (!p->have_code) return fail(p, TWSTFT_E_NOCODE, "no PRN code installed (twstft_set_code / twstft_gen_code)");
    if (nrecords == 0) return TWSTFT_OK;
    p->be.use();
    if (ensure_raw(p) != 0 || ensure_results(p, nrecords * p->nch) != 0)
        return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    const size_t B = (size_t)p->cfg.max_batch;
    const bool fast = fast_on(p);
    // One record per chunk: the copy of a record (0.72 ms at 55 GB/s) hides the kernels of the previous one
    // (0.33 ms), so the only time not covered by the PCIe transfer is the last chunk's compute -- keep it small.
    const size_t chunk = 1;
    const short2* h = reinterpret_cast<const short2*>(iq);
    int slot = 0;
    for (size_t r0 = 0; r0 < nrecords; r0 += chunk, slot ^= 1) {
        const size_t nb = (nrecords - r0) < chunk ? (nrecords - r0) : chunk;
        short2* d = p->d_raw + (size_t)slot * B * p->raw_elems_per_rec;
        p->be.stage_copy(slot, d, h + r0 * p->raw_elems_per_rec, sizeof(short2) * p->raw_elems_per_rec * nb);
        RunArgs a{};
        a.d_raw = d; a.nrec = (int)nb; a.nch = p->nch; a.rec_stride = (i64)p->raw_elems_per_rec; a.d_rec_offset = nullptr;
        a.estimate = 1; a.variant = p->cfg.variant; a.ddof = p->cfg.ddof; a.fs = p->cfg.fs;
        a.d_out = p->d_res + r0 * p->nch;
        a.allow_fast = fast;
        if (p->pipe->estimate(a, p->est) != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        if (p->cfg.fine_freq) {
            const int frc = p->pipe->fine_frequency(a);
            if (frc == -3) return fail(p, TWSTFT_E_ARG, "fine_freq: the window is shorter than fs/3 (the reference's polyfit raises)");
            if (frc != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        }
        if (p->pipe->correlate(a) != 0)
            return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        p->be.stage_release(slot);
    }
    p->be.d2h(results, p->d_res, sizeof(Result) * nrecords * p->nch);
    p->be.sync();
    if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    size_t nrerun = 0;
    if (fast) {
        const int rc = rerun_uncertified(p, h, false, nrecords, results, &nrerun);
        if (rc != TWSTFT_OK) return rc;
    }
    fast_account(p, nrecords, nrerun, fast);
    return TWSTFT_OK;
}

int twstft_process_windows(twstft_plan* p, const int16_t* iq, int on_device, size_t total_instants,
                           const int64_t* offsets, const double* df, size_t nwin, twstft_result* results)
{
    if (!p || !p->pipe || !iq || !offsets || !df || !results) return TWSTFT_E_ARG;
    if (!p->have_code) return fail(p, TWSTFT_E_NOCODE, "no PRN code installed");
    p->be.use();
    const i64 N = p->cfg.nsamples;
    for (size_t i = 0; i < nwin; i++)
        if (offsets[i] < 0 || (size_t)(offsets[i] + N) > total_instants)
            return fail(p, TWSTFT_E_ARG, "window outside the buffer");
    const short2* d_iq = reinterpret_cast<const short2*>(iq);
    if (!on_device) {   // host buffer: staged in a plan-owned buffer that only ever grows (no allocation per call)
        const size_t need = total_instants * p->nch;
        if (p->win_cap < need) {
            p->be.sync();
            p->be.free(p->d_winbuf);
            p->d_winbuf = (short2*)p->be.alloc(sizeof(short2) * need);
            p->win_cap = p->d_winbuf ? need : 0;
            if (!p->d_winbuf || !p->be.ok()) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
        }
        p->be.h2d(p->d_winbuf, iq, sizeof(short2) * need);
        d_iq = p->d_winbuf;
    }
    const int B = p->cfg.max_batch;
    const bool fast = fast_on(p);
    int rc = TWSTFT_OK;
    std::vector<i64> off(B), off2(B);
    std::vector<double> df2((size_t)B * p->nch);
    std::vector<twstft_result> rows((size_t)B * p->nch);
    size_t nrerun = 0;
    for (size_t w0 = 0; w0 < nwin && rc == TWSTFT_OK; w0 += B) {
        const int nb = (int)((nwin - w0) < (size_t)B ? (nwin - w0) : B);
        for (int i = 0; i < nb; i++) off[i] = offsets[w0 + i] * p->nch;
        p->be.h2d(p->d_off, off.data(), sizeof(i64) * nb);
        p->be.h2d(p->d_dfin, df + w0 * p->nch, sizeof(double) * nb * p->nch);
        RunArgs a{};
        a.d_raw = d_iq; a.nrec = nb; a.nch = p->nch; a.rec_stride = 0; a.d_rec_offset = p->d_off;
        a.estimate = 0; a.d_df_in = p->d_dfin; a.variant = p->cfg.variant; a.ddof = p->cfg.ddof;
        a.fs = p->cfg.fs; a.d_out = p->d_res;
        a.allow_fast = fast;
        if (p->pipe->correlate(a) != 0) { rc = fail(p, TWSTFT_E_CUDA, p->be.err_string()); break; }
        p->be.d2h(results + w0 * p->nch, p->d_res, sizeof(Result) * nb * p->nch);
        p->be.sync();
        if (!p->be.ok()) { rc = fail(p, TWSTFT_E_CUDA, p->be.err_string()); break; }
        if (!fast) continue;
        // fast peak path: the windows of this batch whose certificate failed (status bit 2) go through the full path
        // together, in one launch sequence
        std::vector<int> flagged;
        for (int i = 0; i < nb; i++) {
            bool f = false;
            for (int c = 0; c < p->nch; c++) f = f || (((int)results[(w0 + i) * p->nch + c].status) & 4);
            if (f) flagged.push_back(i);
        }
        if (flagged.empty()) continue;
        nrerun += flagged.size();
        const int nf = (int)flagged.size();
        for (int i = 0; i < nf; i++) {
            off2[i] = off[flagged[i]];
            for (int c = 0; c < p->nch; c++) df2[(size_t)i * p->nch + c] = df[(w0 + flagged[i]) * p->nch + c];
        }
        p->be.h2d(p->d_off, off2.data(), sizeof(i64) * nf);
        p->be.h2d(p->d_dfin, df2.data(), sizeof(double) * nf * p->nch);
        RunArgs b2 = a;
        b2.nrec = nf; b2.allow_fast = 0;
        if (p->pipe->correlate(b2) != 0) { rc = fail(p, TWSTFT_E_CUDA, p->be.err_string()); break; }
        p->be.d2h(rows.data(), p->d_res, sizeof(Result) * nf * p->nch);
        p->be.sync();
        if (!p->be.ok()) { rc = fail(p, TWSTFT_E_CUDA, p->be.err_string()); break; }
        for (int i = 0; i < nf; i++)
            memcpy(results + (w0 + flagged[i]) * p->nch, rows.data() + (size_t)i * p->nch, sizeof(twstft_result) * p->nch);
    }
    if (rc == TWSTFT_OK) fast_account(p, nwin, nrerun, fast);
    return rc;
}

// carrier estimator on nblocks blocks of ls*N sample instants (ls = 1: the window length; ls = 2: the 2-s blocks
// of the K scripts, N = fs)
int twstft_estimate_carrier_block(twstft_plan* p, const int16_t* iq, int on_device, size_t nblocks, int ls,
                                  double* df, int64_t* bins)
{
    if (!p || !p->pipe || !iq || (ls != 1 && ls != 2)) return TWSTFT_E_ARG;
    if (ls == 2 && p->cfg.est_sub_mean) return fail(p, TWSTFT_E_ARG, "ls=2 blocks are estimated on the raw samples (est_sub_mean=0)");
    p->be.use();
    if (!on_device && ensure_raw(p) != 0) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    const int B = p->cfg.max_batch;
    const size_t blk_elems = p->raw_elems_per_rec * ls;
    const short2* src = reinterpret_cast<const short2*>(iq);
    const EstimatorCfg& e = ls == 2 ? p->est2 : p->est;
    for (size_t r0 = 0; r0 < nblocks; r0 += B) {
        const int nb = (int)((nblocks - r0) < (size_t)B ? (nblocks - r0) : B);
        const short2* d = src + r0 * blk_elems;
        if (!on_device) {
            p->be.h2d(p->d_raw, d, sizeof(short2) * blk_elems * nb);
            d = p->d_raw;
        }
        RunArgs a{};
        a.d_raw = d; a.nrec = nb; a.nch = p->nch; a.rec_stride = (i64)blk_elems;
        a.estimate = 1; a.variant = p->cfg.variant; a.ddof = p->cfg.ddof; a.fs = p->cfg.fs; a.d_out = p->d_res;
        const int rc = ls == 2 ? p->pipe->estimate_block2(a, e) : p->pipe->estimate(a, e);
        if (rc != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        for (int l = 0; l < nb * p->nch; l++) {
            if (df) p->pipe->debug_fetch(5, l, df + r0 * p->nch + l, sizeof(double));
            if (bins) p->pipe->debug_fetch(7, l, bins + r0 * p->nch + l, sizeof(i64));
        }
        p->be.sync();
        if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    }
    return TWSTFT_OK;
}

int twstft_estimate_carrier(twstft_plan* p, const int16_t* iq, int on_device, size_t nrecords,
                            double* df, int64_t* bins)
{
    return twstft_estimate_carrier_block(p, iq, on_device, nrecords, 1, df, bins);
}

// ---------------------------------------------------------------------------------------------
// Doppler bank (A12) and search_df (A4)

// stage one window (all channels) on the device unless it is already there
static const short2* stage_window(twstft_plan* p, const int16_t* iq, int on_device, int64_t offset)
{
    const short2* src = reinterpret_cast<const short2*>(iq) + (size_t)offset * p->nch;
    if (on_device) return src;
    if (ensure_raw(p) != 0) return nullptr;
    p->be.h2d(p->d_raw, src, sizeof(short2) * p->raw_elems_per_rec);
    return p->d_raw;
}

static int run_bank(twstft_plan* p, const short2* d_win, const double* dfs, size_t nbins, int sub_mean,
                    twstft_bank_result* results, const float2* cf_win = nullptr)
{
    const int nch_src = cf_win ? 1 : p->nch;
    const int B = p->cfg.max_batch;
    std::vector<i64> off(B, 0);
    std::vector<double> dfl((size_t)B * p->nch);
    p->be.h2d(p->d_off, off.data(), sizeof(i64) * B);
    p->be.sync();          // `off` is a local
    BankRow* d_rows = (BankRow*)p->d_res;                // the result staging is large enough: Result > BankRow
    for (size_t b0 = 0; b0 < nbins; b0 += B) {
        const int nb = (int)((nbins - b0) < (size_t)B ? (nbins - b0) : B);
        for (int i = 0; i < nb; i++)
            for (int c = 0; c < nch_src; c++) dfl[(size_t)i * nch_src + c] = dfs[b0 + i];
        p->be.h2d(p->d_dfin, dfl.data(), sizeof(double) * nb * nch_src);
        RunArgs a{};
        a.d_raw = d_win; a.d_cf = cf_win; a.nrec = nb; a.nch = nch_src; a.rec_stride = 0; a.d_rec_offset = p->d_off;
        a.estimate = 0; a.d_df_in = p->d_dfin; a.variant = 1; a.ddof = 1; a.fs = p->cfg.fs;
        if (p->pipe->bank(a, cf_win ? 0 : sub_mean, d_rows) != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        p->be.d2h(results + b0 * nch_src, d_rows, sizeof(BankRow) * nb * nch_src);
        p->be.sync();
        if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    }
    return TWSTFT_OK;
}

int twstft_doppler_bank(twstft_plan* p, const int16_t* iq, int iq_on_device, size_t total_instants, int64_t offset,
                        const double* dfs, size_t nbins, int sub_mean, twstft_bank_result* results)
{
    if (!p || !p->pipe || !iq || !dfs || !results) return TWSTFT_E_ARG;
    if (!p->have_code) return fail(p, TWSTFT_E_NOCODE, "no PRN code installed");
    if (offset < 0 || (size_t)(offset + p->cfg.nsamples) > total_instants)
        return fail(p, TWSTFT_E_ARG, "window outside the buffer");
    p->be.use();
    const short2* d_win = stage_window(p, iq, iq_on_device, offset);
    if (!d_win) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    return run_bank(p, d_win, dfs, nbins, sub_mean, results);
}

// Doppler / code-phase acquisition with a step-halving frequency grid: the search loop of
// experiments/231001_DLL_PLL/rxcomplex.cpp:522-568 over the bank kernels.  Every grid point is one bank bin (derotate,
// correlate at 1x, peak magnitude); the best point so far is replaced only by a strictly larger peak, in grid order;
// after each pass the step is halved, the range becomes the new step around the best carrier, until the step is < 1 Hz.
int twstft_acquire(twstft_plan* p, const int16_t* iq, int iq_on_device, size_t total_instants, int64_t offset,
                   double fc_init, double range, double step, double min_step, int sub_mean, int channel,
                   double* fc_out, twstft_bank_result* best_out, int32_t* nbins_out)
{
    if (!p || !p->pipe || !iq || !fc_out || !(step > 0) || !(range >= 0) || channel < 0 || channel >= p->nch) return TWSTFT_E_ARG;
    if (!p->have_code) return fail(p, TWSTFT_E_NOCODE, "no PRN code installed");
    if (offset < 0 || (size_t)(offset + p->cfg.nsamples) > total_instants)
        return fail(p, TWSTFT_E_ARG, "window outside the buffer");
    if (!(min_step > 0)) min_step = 1.0;                     // rxcomplex.cpp:557 "if (fstep < 1.0) break"
    p->be.use();
    const short2* d_win = stage_window(p, iq, iq_on_device, offset);
    if (!d_win) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    double fc = fc_init, frange = range, fstep = step, pk = 0.0;
    twstft_bank_result best{};
    int32_t nb = 0;
    std::vector<double> grid;
    std::vector<twstft_bank_result> rows;
    for (;;) {
        const double flow = fc - frange, fhigh = fc + frange;
        grid.clear();
        for (double fcc = flow; fcc <= fhigh; fcc += fstep) grid.push_back(fcc);      // :527, accumulated like the reference
        rows.resize(grid.size() * p->nch);
        if (!grid.empty()) {
            const int rc = run_bank(p, d_win, grid.data(), grid.size(), sub_mean, rows.data());
            if (rc != TWSTFT_OK) return rc;
        }
        for (size_t i = 0; i < grid.size(); i++) {
            const twstft_bank_result& r = rows[i * p->nch + channel];
            if (r.prnsig > pk) { pk = r.prnsig; fc = grid[i]; best = r; }            // :545-550
        }
        nb += (int32_t)grid.size();
        fstep = fstep / 2.0;                                                          // :554-557
        frange = fstep;
        if (fstep < min_step) break;
    }
    *fc_out = fc;
    if (best_out) *best_out = best;
    if (nbins_out) *nbins_out = nb;
    return TWSTFT_OK;
}

// Body of search_df (claudio_aligned_code_re_separate.m:26-46).  `sp` owns the x^p spectrum (its estimator window
// on the ls-block axis), `cp` owns the code and runs the per-candidate correlation on the window at d_win.
static int search_df_core(twstft_plan* sp, const short2* d_blk, int ls, twstft_plan* cp, const short2* d_win,
                          double df_threshold, int64_t* kbon, double* df_found, int32_t* ncand,
                          const float2* cf_blk = nullptr, const float2* cf_win = nullptr, const EstimatorCfg* e_over = nullptr)
{
    const EstimatorCfg& e = e_over ? *e_over : (ls == 2 ? sp->est2 : sp->est);
    const size_t wlen = (size_t)(e.win_hi - e.win_lo + 1);
    const int nch_src = cf_blk ? 1 : sp->nch;
    sp->be.use();
    float* d_mag = (float*)sp->be.alloc(sizeof(float) * wlen * nch_src);
    if (!sp->be.ok()) return fail(sp, TWSTFT_E_NOMEM, "device allocation failed");
    i64 zero = 0;
    sp->be.h2d(sp->d_off, &zero, sizeof(i64));
    RunArgs a{};
    a.d_raw = d_blk; a.d_cf = cf_blk; a.nrec = 1; a.nch = nch_src; a.rec_stride = 0; a.d_rec_offset = sp->d_off;
    a.estimate = 1; a.variant = sp->cfg.variant; a.ddof = sp->cfg.ddof; a.fs = sp->cfg.fs;
    int rc = TWSTFT_OK;
    std::vector<float> mag(wlen * nch_src);
    const int src = ls == 2 ? sp->pipe->window_spectrum_block2(a, e, d_mag) : sp->pipe->window_spectrum(a, e, d_mag);
    if (src != 0) rc = fail(sp, TWSTFT_E_CUDA, sp->be.err_string());
    if (rc == TWSTFT_OK) {
        sp->be.d2h(mag.data(), d_mag, sizeof(float) * wlen * nch_src);
        sp->be.sync();
        if (!sp->be.ok()) rc = fail(sp, TWSTFT_E_CUDA, sp->be.err_string());
    }
    sp->be.free(d_mag);
    if (rc != TWSTFT_OK) return rc;
    cp->be.use();
    for (int c = 0; c < nch_src; c++) {
        // ktmp=find(d2(k) > median(d2(k))*df_threshold)   re_separate.m:29
        std::vector<double> v(wlen);
        for (size_t i = 0; i < wlen; i++) v[i] = sqrt((double)mag[(size_t)c * wlen + i]);
        std::vector<double> s(v);
        std::nth_element(s.begin(), s.begin() + wlen / 2, s.end());
        double med = s[wlen / 2];
        if (wlen % 2 == 0) med = 0.5 * (med + *std::max_element(s.begin(), s.begin() + wlen / 2));
        std::vector<i64> cand;
        for (size_t i = 0; i < wlen; i++)
            if (v[i] > med * df_threshold) cand.push_back(e.win_lo + (i64)i);
        kbon[c] = 0;
        if (df_found) df_found[c] = 0.0;
        if (ncand) ncand[c] = (int32_t)cand.size();
        if (cand.empty() || cand.size() >= 100) continue;          // re_separate.m:31
        std::vector<double> dfs(cand.size());
        for (size_t i = 0; i < cand.size(); i++) dfs[i] = axis_value(e, cand[i]) / (double)e.power;
        const int nch_b = cf_win ? 1 : cp->nch;
        std::vector<twstft_bank_result> rows(cand.size() * nch_b);
        rc = run_bank(cp, d_win, dfs.data(), dfs.size(), e.sub_mean, rows.data(), cf_win);   // y=d(1:length(fcode)).*lo (:35)
        if (rc != TWSTFT_OK) return rc;
        for (size_t i = 0; i < cand.size(); i++)
            if (rows[i * nch_b + c].metric > 100.0) {            // "SNR > 20 dB", last accepted wins (:42)
                kbon[c] = cand[i] + 1;
                if (df_found) df_found[c] = dfs[i];
            }
    }
    return TWSTFT_OK;
}

int twstft_search_df(twstft_plan* p, const int16_t* iq, int iq_on_device, size_t total_instants, int64_t offset,
                     double df_threshold, int64_t* kbon, double* df_found, int32_t* ncand)
{
    if (!p || !p->pipe || !iq || !kbon) return TWSTFT_E_ARG;
    if (!p->have_code) return fail(p, TWSTFT_E_NOCODE, "no PRN code installed");
    if (offset < 0 || (size_t)(offset + p->cfg.nsamples) > total_instants)
        return fail(p, TWSTFT_E_ARG, "window outside the buffer");
    p->be.use();
    const short2* d_win = stage_window(p, iq, iq_on_device, offset);
    if (!d_win) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    return search_df_core(p, d_win, 1, p, d_win, df_threshold, kbon, df_found, ncand);
}

int twstft_search_df_block(twstft_plan* est, twstft_plan* cor, const int16_t* iq, int iq_on_device, int ls,
                           double df_threshold, int64_t* kbon, double* df_found, int32_t* ncand)
{
    if (!est || !cor || !est->pipe || !cor->pipe || !iq || !kbon || (ls != 1 && ls != 2)) return TWSTFT_E_ARG;
    if (!cor->have_code) return fail(cor, TWSTFT_E_NOCODE, "no PRN code installed");
    if (est->nch != cor->nch || est->be.device != cor->be.device)
        return fail(est, TWSTFT_E_ARG, "estimator and correlator plans must share the layout and the device");
    if (cor->cfg.nsamples > est->cfg.nsamples * ls)
        return fail(est, TWSTFT_E_ARG, "the correlation window is longer than the block");
    if (ls == 2 && est->cfg.est_sub_mean)
        return fail(est, TWSTFT_E_ARG, "ls=2 blocks are estimated on the raw samples (est_sub_mean=0)");
    est->be.use();
    const short2* d_blk = reinterpret_cast<const short2*>(iq);
    if (!iq_on_device) {
        if (ensure_raw(est) != 0) return fail(est, TWSTFT_E_NOMEM, "device allocation failed");
        est->be.h2d(est->d_raw, iq, sizeof(short2) * est->raw_elems_per_rec * ls);
        est->be.sync();                      // the correlator plan reads the same staging on its own stream
        d_blk = est->d_raw;
    }
    const int rc = search_df_core(est, d_blk, ls, cor, d_blk, df_threshold, kbon, df_found, ncand);
    if (rc != TWSTFT_OK && cor != est && cor->err.size() && est->err.empty()) est->err = cor->err;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Octave function seam: the windows arrive as complex doubles, mean already removed by the caller
// (godual_ranging.m:12,80,94; claudio_aligned_code_lo_separate.m:24,132; ..._re_separate.m:26,48)

// grow the complex-float window buffer / the complex-double upload staging to `n32` / `n64` samples
static int ensure_cf(twstft_plan* p, size_t n32, size_t n64)
{
    if (p->cf32_cap < n32) {
        p->be.sync();
        p->be.free(p->d_cf32);
        p->d_cf32 = (float2*)p->be.alloc(sizeof(float2) * n32);
        p->cf32_cap = p->d_cf32 ? n32 : 0;
    }
    if (p->cf64_cap < n64) {
        p->be.sync();
        p->be.free(p->d_cf64);
        p->d_cf64 = (double2*)p->be.alloc(sizeof(double2) * n64);
        p->cf64_cap = p->d_cf64 ? n64 : 0;
    }
    return (p->cf32_cap >= n32 && p->cf64_cap >= n64 && p->be.ok()) ? 0 : -1;
}

// host complex doubles [n] -> device complex floats at dst (through the upload staging, `chunk` samples at a time)
static void upload_cf64(twstft_plan* p, const double* host, size_t n, float2* dst)
{
    const size_t chunk = p->cf64_cap;
    for (size_t i0 = 0; i0 < n; i0 += chunk) {
        const size_t m = (n - i0) < chunk ? (n - i0) : chunk;
        p->be.h2d(p->d_cf64, host + 2 * i0, sizeof(double2) * m);
        CvtParams c{};
        c.src = p->d_cf64; c.dst = dst + i0; c.n = (i64)m;
        i64 nblk = ((i64)m + KCvtF64::NT - 1) / KCvtF64::NT;
        const i64 wave = (i64)p->be.num_sms * 8;
        if (nblk > wave) nblk = wave;
        p->be.launch<KCvtF64>(dim3((unsigned)nblk, 1, 1), c);
    }
}

static int processing_cf_core(twstft_plan* p, const double* d, size_t nwin, const EstimatorCfg* est, const double* df,
                              twstft_result* results)
{
    if (!p->have_code) return fail(p, TWSTFT_E_NOCODE, "no PRN code installed (twstft_set_code / twstft_gen_code)");
    if (p->cfg.fine_freq) return fail(p, TWSTFT_E_ARG, "fine_freq needs the int16 records (twstft_process)");
    const size_t N = (size_t)p->cfg.nsamples, B = (size_t)p->cfg.max_batch;
    const size_t perbatch = B * (size_t)p->nch;                 // the plan has room for max_batch * nch lanes
    if (ensure_cf(p, N * perbatch, N) != 0) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    for (size_t w0 = 0; w0 < nwin; w0 += perbatch) {
        const int nb = (int)((nwin - w0) < perbatch ? (nwin - w0) : perbatch);
        upload_cf64(p, d + 2 * w0 * N, (size_t)nb * N, p->d_cf32);
        RunArgs a{};
        a.d_cf = p->d_cf32; a.d_raw = nullptr; a.nrec = nb; a.nch = 1; a.rec_stride = (i64)N; a.d_rec_offset = nullptr;
        a.estimate = est ? 1 : 0; a.variant = p->cfg.variant; a.ddof = p->cfg.ddof; a.fs = p->cfg.fs; a.d_out = p->d_res;
        a.allow_fast = 0;
        if (est) {
            if (p->pipe->estimate(a, *est) != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        } else {
            p->be.h2d(p->d_dfin, df + w0, sizeof(double) * nb);
            a.d_df_in = p->d_dfin;
        }
        if (p->pipe->correlate(a) != 0) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
        p->be.d2h(results + w0, p->d_res, sizeof(Result) * nb);
        p->be.sync();
        if (!p->be.ok()) return fail(p, TWSTFT_E_CUDA, p->be.err_string());
    }
    return TWSTFT_OK;
}

int twstft_processing_cf64(twstft_plan* p, const double* d, size_t nwin, int64_t k_first, int64_t k_last,
                           twstft_result* results)
{
    if (!p || !p->pipe || !d || !results) return TWSTFT_E_ARG;
    p->be.use();
    EstimatorCfg e = p->est;
    e.sub_mean = 0;                        // the caller removed the mean (godual_ranging.m:80,94)
    if (k_first > 0) {                     // k(1) .. k(end), 1-based on the fftshift'ed axis
        if (k_last < k_first || k_last > e.naxis) return fail(p, TWSTFT_E_ARG, "estimator window outside the spectrum");
        e.win_lo = k_first - 1;
        e.win_hi = k_last - 1;
    }
    return processing_cf_core(p, d, nwin, &e, nullptr, results);
}

int twstft_processing_cf64_df(twstft_plan* p, const double* d, size_t nwin, const double* df, twstft_result* results)
{
    if (!p || !p->pipe || !d || !df || !results) return TWSTFT_E_ARG;
    p->be.use();
    return processing_cf_core(p, d, nwin, nullptr, df, results);
}

int twstft_search_df_cf64(twstft_plan* est, twstft_plan* cor, const double* d, int ls, int64_t k_first, int64_t k_last,
                          double df_threshold, int64_t* kbon, double* df_found, int32_t* ncand)
{
    if (!est || !cor || !est->pipe || !cor->pipe || !d || !kbon || (ls != 1 && ls != 2)) return TWSTFT_E_ARG;
    if (!cor->have_code) return fail(cor, TWSTFT_E_NOCODE, "no PRN code installed");
    if (est->be.device != cor->be.device) return fail(est, TWSTFT_E_ARG, "estimator and correlator plans must share the device");
    if (cor->cfg.nsamples > est->cfg.nsamples * ls)
        return fail(est, TWSTFT_E_ARG, "the correlation window is longer than the block");
    est->be.use();
    EstimatorCfg e = ls == 2 ? est->est2 : est->est;
    e.sub_mean = 0;                        // search_df squares the block as it is (re_separate.m:28)
    if (k_first > 0) {
        if (k_last < k_first || k_last > e.naxis) return fail(est, TWSTFT_E_ARG, "estimator window outside the spectrum");
        e.win_lo = k_first - 1;
        e.win_hi = k_last - 1;
    }
    const size_t nblk = (size_t)est->cfg.nsamples * ls;
    if (ensure_cf(est, nblk, (size_t)est->cfg.nsamples) != 0) return fail(est, TWSTFT_E_NOMEM, "device allocation failed");
    upload_cf64(est, d, nblk, est->d_cf32);
    est->be.sync();                        // the correlator plan reads the same buffer on its own stream
    const int rc = search_df_core(est, nullptr, ls, cor, nullptr, df_threshold, kbon, df_found, ncand,
                                  est->d_cf32, est->d_cf32, &e);
    if (rc != TWSTFT_OK && cor != est && cor->err.size() && est->err.empty()) est->err = cor->err;
    return rc;
}

// |fft(x.^p)| at `nbins` adjacent positions of the fftshift'ed spectrum of a buffer of ANY length
int twstft_estimate_bins(twstft_plan* p, const int16_t* iq, int iq_on_device, size_t ninstants, int64_t s_first,
                         int nbins, double* mag)
{
    if (!p || !p->pipe || !iq || !mag || nbins < 1 || nbins > KBinsDFT::MAXB) return TWSTFT_E_ARG;
    const i64 L = (i64)ninstants;
    if (L < 1 || L >= ((i64)1 << 31) || s_first < 0 || s_first + nbins > L)
        return fail(p, TWSTFT_E_ARG, "bins outside the spectrum");
    p->be.use();
    const short2* d = reinterpret_cast<const short2*>(iq);
    short2* tmp = nullptr;
    if (!iq_on_device) {
        tmp = (short2*)p->be.alloc(sizeof(short2) * ninstants * p->nch);
        if (!p->be.ok()) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
        p->be.h2d(tmp, iq, sizeof(short2) * ninstants * p->nch);
        d = tmp;
    }
    BinsParams b{};
    b.src = d; b.ninst = L; b.nch = p->nch; b.power = p->est.power; b.nb = nbins;
    b.k0 = (s_first + (L + 1) / 2) % L;      // fftshift: shifted[s] = X[(s + ceil(L/2)) mod L]
    i64 nblk = (L + KBinsDFT::NT - 1) / KBinsDFT::NT;
    const i64 wave = (i64)p->be.num_sms * 4;
    if (nblk > wave) nblk = wave;
    b.nblk = (int)nblk;
    const size_t npart = (size_t)p->nch * nblk * 2 * KBinsDFT::MAXB;
    b.part = (double*)p->be.alloc(sizeof(double) * npart);
    int rc = TWSTFT_OK;
    if (!p->be.ok()) rc = fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    if (rc == TWSTFT_OK) {
        p->be.launch<KBinsDFT>(dim3((unsigned)nblk, p->nch, 1), b);
        std::vector<double> part(npart);
        p->be.d2h(part.data(), b.part, sizeof(double) * npart);
        p->be.sync();
        if (!p->be.ok()) rc = fail(p, TWSTFT_E_CUDA, p->be.err_string());
        else
            for (int c = 0; c < p->nch; c++)
                for (int i = 0; i < nbins; i++) {
                    double re = 0.0, im = 0.0;
                    for (i64 k = 0; k < nblk; k++) {
                        const double* o = &part[((size_t)c * nblk + k) * 2 * KBinsDFT::MAXB];
                        re += o[2 * i];
                        im += o[2 * i + 1];
                    }
                    mag[(size_t)c * nbins + i] = hypot(re, im);
                }
    }
    p->be.free(b.part);
    if (tmp) p->be.free(tmp);
    return rc;
}

// GoRanging::df (processing/CPP/main.cpp:363-450); kernels and the L = s*M decomposition in k_decim.cuh.
int twstft_estimate_decimated(twstft_plan* p, const int16_t* iq, size_t total_instants, int dec, double foffset,
                              int remote, double* df, int64_t* pos)
{
    if (!p || !p->pipe || !iq || !df || dec < 1) return TWSTFT_E_ARG;
    const int nch = p->nch;
    const i64 M = p->pipe->size();
    const i64 L = (i64)(total_instants / (size_t)dec);          // ftell/(sizeof(int16)*4*N), main.cpp:377
    if (L < M || L % M != 0 || L / M > 32768 || L >= ((i64)1 << 32))
        return fail(p, TWSTFT_E_SIZE, "decimated length must be a multiple (<= 32768) of the plan's window length");
    const int s = (int)(L / M);
    const int nuse = (nch == 2 && !remote) ? 2 : 1;
    const double fs = p->cfg.fs;
    p->be.use();
    // reader: one sample instant out of `dec` (fread 4 int16, fseek the rest, main.cpp:381-397) and the exact raw sums
    std::vector<short2> g((size_t)L * nch);
    long long sum[2][2] = {{0, 0}, {0, 0}};
    const short2* src = reinterpret_cast<const short2*>(iq);
    for (i64 n = 0; n < L; n++)
        for (int c = 0; c < nch; c++) {
            const short2 v = src[(size_t)n * dec * nch + c];
            g[(size_t)n * nch + c] = v;
            sum[c][0] += v.x;
            sum[c][1] += v.y;
        }
    // freq=linspace(-fs/2/N, fs/2/N, L); kmax = last i with freq<16000, kmin = last i with freq<=-16000 (:393-398)
    const double f0 = -fs / 2 / dec, f1 = fs / 2 / dec, delta = L > 1 ? (f1 - f0) / (double)(L - 1) : 0.0;
    auto freq_at = [&](i64 i) { return i == L - 1 ? f1 : f0 + delta * (double)i; };
    const double frange = 8000.0;
    i64 kmin = 0, kmax = 0;
    for (i64 i = 0; i < L; i++) {
        const double f = freq_at(i);
        if (f < 2 * frange) kmax = i;
        if (f <= -2 * frange) kmin = i;
    }
    if (kmax <= kmin) return fail(p, TWSTFT_E_ARG, "empty estimator window");
    std::vector<float2> wt(s);
    for (int j = 0; j < s; j++) {
        const double a = -6.283185307179586476925287 * (double)j / (double)s;
        wt[j] = make_float2((float)cos(a), (float)sin(a));
    }
    short2* d_g = (short2*)p->be.alloc(sizeof(short2) * g.size());
    float2* d_G = (float2*)p->be.alloc(sizeof(float2) * (size_t)L);
    float2* d_wt = (float2*)p->be.alloc(sizeof(float2) * s);
    const unsigned gx = (unsigned)((M + KDecCombine::NT - 1) / KDecCombine::NT);
    const unsigned gy = (unsigned)((s + KDecCombine::KT - 1) / KDecCombine::KT);
    u64* d_key = (u64*)p->be.alloc(sizeof(u64) * (size_t)gx * gy);
    int rc = TWSTFT_OK;
    if (!p->be.ok()) rc = fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    if (rc == TWSTFT_OK) {
        p->be.h2d(d_g, g.data(), sizeof(short2) * g.size());
        p->be.h2d(d_wt, wt.data(), sizeof(float2) * s);
    }
    std::vector<u64> keys((size_t)gx * gy);
    for (int c = 0; c < nuse && rc == TWSTFT_OK; c++) {
        DecSquareParams q{};
        q.g = d_g; q.nch = nch; q.ch = c; q.L = L; q.s = s; q.M = M;
        q.mean_re = (double)sum[c][0] / (double)L; q.mean_im = (double)sum[c][1] / (double)L;
        q.cyc_per_sample = foffset * (double)dec / fs;
        q.dst = d_G;
        p->be.launch<KDecSquare>(dim3((unsigned)((L + KDecSquare::NT - 1) / KDecSquare::NT), 1, 1), q);
        for (int l0 = 0; l0 < s && rc == TWSTFT_OK; l0 += p->max_lanes) {
            const int nl = std::min(p->max_lanes, s - l0);
            if (p->pipe->spectrum_cf32(d_G + (size_t)l0 * M, d_G + (size_t)l0 * M, nl) != 0)
                rc = fail(p, TWSTFT_E_CUDA, p->be.err_string());
        }
        DecTwiddleParams t{};
        t.G = d_G; t.s = s; t.M = M; t.L = L; t.L1 = p->pipe->col_len(); t.N2 = p->pipe->row_len();
        p->be.launch<KDecTwiddle>(dim3((unsigned)((M + KDecTwiddle::NT - 1) / KDecTwiddle::NT), (unsigned)s, 1), t);
        DecCombineParams k{};
        k.G = d_G; k.wtab = d_wt; k.s = s; k.M = M; k.L = L; k.L1 = t.L1; k.N2 = t.N2;
        k.win_lo = c == 0 ? kmin : 0;                 // channel 1: out[kmin:kmax), channel 2: every bin (:427-445)
        k.win_hi = c == 0 ? kmax - 1 : L - 1;
        k.part_key = d_key;
        p->be.launch<KDecCombine>(dim3(gx, gy, 1), k);
        p->be.d2h(keys.data(), d_key, sizeof(u64) * keys.size());
        p->be.sync();
        if (!p->be.ok()) { rc = fail(p, TWSTFT_E_CUDA, p->be.err_string()); break; }
        u64 best = 0;
        for (u64 v : keys) best = v > best ? v : best;
        const i64 ps = (i64)key_index(best);
        df[c] = freq_at(ps) / 2.0 + foffset;
        if (pos) pos[c] = ps;
    }
    p->be.free(d_key);
    p->be.free(d_wt);
    p->be.free(d_G);
    p->be.free(d_g);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Synthetic records on the device (bench / test tooling, k_synth.cuh)
static inline uint64_t splitmix64(uint64_t& x)
{
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double unit(uint64_t& x) { return (double)(splitmix64(x) >> 11) * (1.0 / 9007199254740992.0); }

static SynthLane synth_lane(uint64_t seed, int ch, i64 period, double delay, double df, double phase_cyc, double snr_db,
                            double fs)
{
    SynthLane L{};
    const double ps = pow(10.0, snr_db / 10.0);
    const double sigma = 3000.0 / sqrt(1.0 + ps);            // per-component noise std: total I std ~ 3000 LSB
    L.delay = ((i64)floor(delay) % period + period) % period;
    L.dphi = df / fs;
    L.phase = phase_cyc;
    L.amp = (float)(sigma * sqrt(2.0 * ps));
    L.sigma = (float)sigma;
    L.dc_re = ch == 0 ? 11.3f : -5.2f;
    L.dc_im = ch == 0 ? -7.9f : 3.4f;
    uint64_t k = seed * 2 + (uint64_t)ch + 0x1234567ull;
    const uint64_t kk = splitmix64(k);
    L.key0 = (unsigned)kk;
    L.key1 = (unsigned)(kk >> 32);
    return L;
}

static int run_synth(twstft_plan* p, const std::vector<SynthLane>& lanes, int nch, i64 nsamp, short2* d_out)
{
    SynthLane* d_l = (SynthLane*)p->be.alloc(sizeof(SynthLane) * lanes.size());
    if (!d_l || !p->be.ok()) return fail(p, TWSTFT_E_NOMEM, "device allocation failed");
    p->be.h2d(d_l, lanes.data(), sizeof(SynthLane) * lanes.size());
    SynthParams sp{};
    sp.bits = p->d_bits; sp.period = (i64)p->code_nchips * p->cfg.samples_per_chip; sp.rep = p->cfg.samples_per_chip;
    sp.nch = nch; sp.nsamp = nsamp; sp.lanes = d_l; sp.out = d_out;
    i64 nblk = (nsamp + KSynth::NT - 1) / KSynth::NT;
    const i64 cap = (i64)p->be.num_sms * 8;
    if (nblk > cap) nblk = cap;
    const size_t MAXY = 32768;
    for (size_t l0 = 0; l0 < lanes.size(); l0 += MAXY) {           // gridDim.y limit
        SynthParams q = sp;
        q.lanes = d_l + l0;
        q.out = d_out + (l0 / nch) * (size_t)nsamp * nch;
        const size_t ny = std::min(MAXY, lanes.size() - l0);
        p->be.launch<KSynth>(dim3((unsigned)nblk, (unsigned)ny, 1), q);
    }
    p->be.sync();
    p->be.free(d_l);
    return p->be.ok() ? TWSTFT_OK : fail(p, TWSTFT_E_CUDA, p->be.err_string());
}

int twstft_synth_records(twstft_plan* p, uint64_t seed0, size_t nrecords, double snr1_db, double snr2_db, int16_t* dev_out)
{
    if (!p || !p->pipe || !dev_out) return TWSTFT_E_ARG;
    if (!p->have_code || p->code_qpsk) return fail(p, TWSTFT_E_NOCODE, "synthetic records need an installed BPSK code");
    p->be.use();
    const i64 N = p->cfg.nsamples;
    std::vector<SynthLane> lanes(nrecords * p->nch);
    for (size_t r = 0; r < nrecords; r++) {
        uint64_t s = seed0 + r;
        // ch1 delay 0.25..0.99 s, df +-9 kHz off grid, SNR -5 dB; ch2 loopback delay ~0.7253 s, df 0, +0.6 dB
        const double d1 = (0.25 + 0.74 * unit(s)) * (double)N, f1 = (2.0 * unit(s) - 1.0) * 9000.0, p1 = unit(s);
        const double d2 = 0.7253 * (double)N + (2.0 * unit(s) - 1.0), p2 = unit(s);
        lanes[r * p->nch] = synth_lane(seed0 + r, 0, N, d1, f1, p1, snr1_db, p->cfg.fs);
        if (p->nch == 2) lanes[r * 2 + 1] = synth_lane(seed0 + r, 1, N, d2, 0.0, p2, snr2_db, p->cfg.fs);
    }
    return run_synth(p, lanes, p->nch, N, reinterpret_cast<short2*>(dev_out));
}

int twstft_synth_stream(twstft_plan* p, uint64_t seed, size_t ninstants, double delay, double df, double snr_db,
                        int16_t* dev_out)
{
    if (!p || !p->pipe || !dev_out || p->nch != 1) return TWSTFT_E_ARG;
    if (!p->have_code || p->code_qpsk) return fail(p, TWSTFT_E_NOCODE, "synthetic records need an installed BPSK code");
    p->be.use();
    std::vector<SynthLane> lanes(1);
    lanes[0] = synth_lane(seed, 0, (i64)p->code_nchips * p->cfg.samples_per_chip, delay, df, 0.3 / 6.283185307179586, snr_db, p->cfg.fs);
    return run_synth(p, lanes, 1, (i64)ninstants, reinterpret_cast<short2*>(dev_out));
}

// ---------------------------------------------------------------------------------------------
// A long run of records drawn cyclically from a pool (the day-of-data driver) and the multi-GPU sharder

int twstft_process_cycled(twstft_plan* p, const int16_t* pool, int pool_on_device, size_t pool_records,
                          size_t first_record, size_t nrecords, twstft_result* results)
{
    if (!p || !p->pipe || !pool || !results || pool_records == 0) return TWSTFT_E_ARG;
    const size_t rec_i16 = p->raw_elems_per_rec * 2;
    const size_t B = (size_t)p->cfg.max_batch;
    for (size_t r = first_record; r < first_record + nrecords;) {
        const size_t slot = r % pool_records;
        size_t nb = std::min(B, pool_records - slot);
        nb = std::min(nb, first_record + nrecords - r);
        const int16_t* src = pool + slot * rec_i16;
        twstft_result* out = results + (r - first_record) * p->nch;
        const int rc = pool_on_device ? twstft_process_device(p, src, nb, out) : twstft_process(p, src, nb, out);
        if (rc != TWSTFT_OK) return rc;
        r += nb;
    }
    return TWSTFT_OK;
}

struct twstft_shard {
    std::vector<twstft_plan*> plans;
    std::vector<int> devices;
    std::string err;
};

#ifndef TWSTFT_EMU
// bind the calling thread to the CPUs the kernel lists as local to the GPU (sysfs), so that the pinned buffers it
// allocates next are first-touched on that socket; a no-op when the topology is not exposed
static void bind_thread_to_gpu(int device)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return;
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) return;
    char line[1024] = {0};
    const bool got = fgets(line, sizeof line, f) != nullptr;
    fclose(f);
    if (!got) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); n++; }
    }
    if (n > 0) sched_setaffinity(0, sizeof set, &set);
}
#else
static void bind_thread_to_gpu(int) {}
#endif

int twstft_shard_create(const twstft_config* cfg, const int* devices, int ngpus, twstft_shard** out)
{
    if (!cfg || !out || ngpus < 1) return TWSTFT_E_ARG;
    twstft_shard* s = new twstft_shard();
    *out = s;
    for (int g = 0; g < ngpus; g++) {
        twstft_config c = *cfg;
        c.device_id = devices ? devices[g] : g;
        twstft_plan* p = nullptr;
        const int rc = twstft_plan_create(&c, &p);
        s->plans.push_back(p);
        s->devices.push_back(c.device_id);
        if (rc != TWSTFT_OK) { s->err = std::string("GPU ") + std::to_string(c.device_id) + ": " + twstft_last_error(p); return rc; }
    }
    return TWSTFT_OK;
}
void twstft_shard_destroy(twstft_shard* s)
{
    if (!s) return;
    for (twstft_plan* p : s->plans) twstft_plan_destroy(p);
    delete s;
}
int twstft_shard_ngpus(const twstft_shard* s) { return s ? (int)s->plans.size() : 0; }
twstft_plan* twstft_shard_plan(twstft_shard* s, int g) { return (s && g >= 0 && g < (int)s->plans.size()) ? s->plans[g] : nullptr; }
const char* twstft_shard_last_error(const twstft_shard* s) { return s ? s->err.c_str() : "null shard"; }

int twstft_shard_set_code(twstft_shard* s, const uint8_t* chips, size_t nbytes, int qpsk)
{
    if (!s) return TWSTFT_E_ARG;
    for (twstft_plan* p : s->plans) {
        const int rc = twstft_set_code(p, chips, nbytes, qpsk);
        if (rc != TWSTFT_OK) { s->err = twstft_last_error(p); return rc; }
    }
    return TWSTFT_OK;
}

}  // extern "C"
// run fn(g) on one host thread per GPU; first error wins
template <class F> static int shard_parallel(twstft_shard* s, F fn)
{
    const int G = (int)s->plans.size();
    std::vector<int> rc(G, TWSTFT_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) th.emplace_back([&, g] { rc[g] = fn(g); });
    for (auto& t : th) t.join();
    for (int g = 0; g < G; g++)
        if (rc[g] != TWSTFT_OK) { s->err = std::string("GPU ") + std::to_string(s->devices[g]) + ": " + twstft_last_error(s->plans[g]); return rc[g]; }
    return TWSTFT_OK;
}
extern "C" {

int twstft_shard_process(twstft_shard* s, const int16_t* iq, size_t nrecords, twstft_result* results)
{
    if (!s || !iq || !results) return TWSTFT_E_ARG;
    const size_t G = s->plans.size();
    return shard_parallel(s, [&](int g) {
        twstft_plan* p = s->plans[g];
        const size_t lo = nrecords * (size_t)g / G, hi = nrecords * (size_t)(g + 1) / G;      // SURVEY 8e record ranges
        if (hi == lo) return (int)TWSTFT_OK;
        bind_thread_to_gpu(s->devices[g]);
        return twstft_process(p, iq + lo * p->raw_elems_per_rec * 2, hi - lo, results + lo * p->nch);
    });
}

int twstft_shard_process_cycled(twstft_shard* s, const int16_t* pool, size_t pool_records, size_t nrecords, int resident,
                                twstft_result* results, double* seconds_wall, double* seconds_dev)
{
    if (!s || !pool || !results || pool_records == 0) return TWSTFT_E_ARG;
    const size_t G = s->plans.size();
    std::vector<int16_t*> local(G, nullptr);       // per-GPU copy of the pool: device memory (resident) or pinned host memory
    const size_t pool_bytes = pool_records * s->plans[0]->raw_elems_per_rec * sizeof(short2);
    int rc = shard_parallel(s, [&](int g) {        // set-up, not timed
        twstft_plan* p = s->plans[g];
        bind_thread_to_gpu(s->devices[g]);
        p->be.use();
        if (resident) {
            void* d = nullptr;
            int r = twstft_device_alloc(p, pool_bytes, &d);
            if (r != TWSTFT_OK) return r;
            local[g] = (int16_t*)d;
            return twstft_copy_to_device(p, d, pool, pool_bytes);
        }
        local[g] = (int16_t*)twstft_host_alloc(pool_bytes);       // first-touched by the thread bound to this GPU's CPUs
        if (!local[g]) return fail(p, TWSTFT_E_NOMEM, "pinned host allocation failed");
        memcpy(local[g], pool, pool_bytes);
        return (int)TWSTFT_OK;
    });
    std::vector<float> dev_ms(G, 0.f);
    const auto t0 = std::chrono::steady_clock::now();
    if (rc == TWSTFT_OK)
        rc = shard_parallel(s, [&](int g) {
            twstft_plan* p = s->plans[g];
            const size_t lo = nrecords * (size_t)g / G, hi = nrecords * (size_t)(g + 1) / G;
            twstft_timer_start(p);
            const int r = twstft_process_cycled(p, local[g], resident, pool_records, lo, hi - lo, results + lo * p->nch);
            twstft_timer_stop(p, &dev_ms[g]);
            return r;
        });
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_wall) *seconds_wall = std::chrono::duration<double>(t1 - t0).count();
    if (seconds_dev) for (size_t g = 0; g < G; g++) seconds_dev[g] = 1e-3 * dev_ms[g];
    for (size_t g = 0; g < G; g++) {
        if (!local[g]) continue;
        if (resident) twstft_device_free(s->plans[g], local[g]);
        else twstft_host_free(local[g]);
    }
    return rc;
}

void* twstft_host_alloc(size_t bytes)
{
#ifdef TWSTFT_EMU
    return malloc(bytes);
#else
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
#endif
}
void twstft_host_free(void* p)
{
#ifdef TWSTFT_EMU
    free(p);
#else
    if (p) cudaFreeHost(p);
#endif
}
int twstft_device_alloc(twstft_plan* p, size_t bytes, void** out)
{
    if (!p || !out) return TWSTFT_E_ARG;
    p->be.use();
    *out = p->be.alloc(bytes);
    return (*out && p->be.ok()) ? TWSTFT_OK : fail(p, TWSTFT_E_NOMEM, "device allocation failed");
}
int twstft_device_free(twstft_plan* p, void* d)
{
    if (!p) return TWSTFT_E_ARG;
    p->be.use();
    p->be.sync();
    p->be.free(d);
    return TWSTFT_OK;
}
int twstft_copy_to_device(twstft_plan* p, void* dst, const void* src, size_t bytes)
{
    if (!p) return TWSTFT_E_ARG;
    p->be.use();
    p->be.h2d(dst, src, bytes);
    p->be.sync();
    return p->be.ok() ? TWSTFT_OK : fail(p, TWSTFT_E_CUDA, p->be.err_string());
}
int twstft_copy_to_host(twstft_plan* p, void* dst, const void* src, size_t bytes)
{
    if (!p) return TWSTFT_E_ARG;
    p->be.use();
    p->be.d2h(dst, src, bytes);
    p->be.sync();
    return p->be.ok() ? TWSTFT_OK : fail(p, TWSTFT_E_CUDA, p->be.err_string());
}
int twstft_synchronize(twstft_plan* p)
{
    if (!p) return TWSTFT_E_ARG;
    p->be.use();
    p->be.sync();
    return p->be.ok() ? TWSTFT_OK : fail(p, TWSTFT_E_CUDA, p->be.err_string());
}
int twstft_timer_start(twstft_plan* p)
{
    if (!p) return TWSTFT_E_ARG;
    p->be.use();
    p->be.timer_start();
    return TWSTFT_OK;
}
int twstft_timer_stop(twstft_plan* p, float* ms)
{
    if (!p || !ms) return TWSTFT_E_ARG;
    p->be.use();
    *ms = p->be.timer_stop();
    return p->be.ok() ? TWSTFT_OK : fail(p, TWSTFT_E_CUDA, p->be.err_string());
}
int twstft_profile_enable(twstft_plan* p, int on)
{
    if (!p) return TWSTFT_E_ARG;
    p->be.use();
    if (!on) p->be.prof_collect();
    p->be.profiling = on != 0;
    return TWSTFT_OK;
}
int twstft_profile_read(twstft_plan* p, int slot, char* name, size_t name_cap, double* total_ms, int64_t* count)
{
    if (!p) return TWSTFT_E_ARG;
    p->be.use();
    p->be.prof_collect();
    if (slot < 0 || slot >= (int)p->be.prof.size()) return 1;   /* end of list */
    if (name && name_cap) { strncpy(name, p->be.prof[slot].name, name_cap - 1); name[name_cap - 1] = 0; }
    if (total_ms) *total_ms = p->be.prof[slot].ms;
    if (count) *count = p->be.prof[slot].count;
    return TWSTFT_OK;
}
int twstft_debug_fetch(twstft_plan* p, int what, int lane, void* host_out, size_t bytes)
{
    if (!p || !p->pipe) return TWSTFT_E_ARG;
    p->be.use();
    return p->pipe->debug_fetch(what, lane, host_out, bytes);
}

}  // extern "C"
