/*
 * twstft.h -- C ABI of the B200-native amaranth_twstft post-processing correlator.
 *
 * The reference has no FFI; this ABI is drawn at its narrowest seams.
 * Each entry point names the reference interface it replaces (paths relative to the reference root):
 *
 *   twstft_plan_create   script set-up: fs/Nint/windows/code spectrum
 *                        processing/Octave/godual_ranging.m:3-10,62-73
 *                        acquisition/claudio_aligned_code_lo_separate.m:7-22,92-113
 *                        processing/CPP/main.cpp:93-189 (GoRanging ctor)
 *   twstft_set_code      code=fread(f,inf,'int8') ... fcode=conj(fft(code'))      godual_ranging.m:62-66
 *   twstft_gen_code      write_prn_seq / nextstate                                amaranth_twstft/common.py:23-30,59-73
 *   twstft_process       the record loop + processing(d,k) of variant G           godual_ranging.m:12-49,75-102
 *                        (= GoRanging::compute + _process_method,                 processing/CPP/main.cpp:224-361,452-519)
 *   twstft_process_windows  processing(d,k,df) / processing(d,df) of variant K   claudio_aligned_code_lo_separate.m:24-77
 *                                                                                 claudio_aligned_code_re_separate.m:48-101
 *   twstft_estimate_carrier  d2=fftshift(abs(fft(d.^2))); [~,df]=max(d2(k))      godual_ranging.m:14-15, lo_separate.m:126-129
 *   twstft_estimate_carrier_block  the estimator on the 2-s blocks of the K scripts                lo_separate.m:120-129
 *   twstft_doppler_bank  per-bin body of search_df over a grid of offsets (cross-ambiguity)
 *                                                                                 claudio_aligned_code_re_separate.m:33-42
 *   twstft_acquire       coarse grid + step halving around the best bin            experiments/231001_DLL_PLL/rxcomplex.cpp:522-568
 *   twstft_search_df     k=search_df(d,k,df_threshold)                            claudio_aligned_code_re_separate.m:26-46
 *   twstft_search_df_block  the same on the 2-s block the scripts hand it                 claudio_aligned_code_re_separate.m:150
 *   twstft_estimate_bins d2=fftshift(abs(fft(d.^2))); max(d2(kbon-3:kbon+3)), d=[dold;block]  claudio_aligned_code_re_separate.m:166-168
 *   twstft_estimate_decimated  GoRanging::df: whole-file estimator on one instant out of 25            processing/CPP/main.cpp:363-450
 *   twstft_processing_cf64     [indice,correction,SNRr,SNRi,df,puissance,...]=processing(d,k)   godual_ranging.m:12
 *   twstft_processing_cf64_df  [xval,indice,correction,...]=processing(d,k,df) / processing(d,df)
 *                                                     claudio_aligned_code_lo_separate.m:24, ..._re_separate.m:48
 *   twstft_search_df_cf64      k=search_df(d,k,df_threshold)                      claudio_aligned_code_re_separate.m:26
 *                        (the three take the complex double column `d` the Octave functions receive, mean already
 *                         removed by the caller, and the index vector k as its 1-based first/last element: what
 *                         processing/C/twstft_mex.c marshals)
 *   twstft_result        one row of the .mat variables                            godual_ranging.m:128-130, lo_separate.m:166
 *
 * Conventions: plain C, caller owns every host buffer, the plan owns device memory and streams,
 * int return codes (0 = ok, <0 = error, text via twstft_last_error), never an exception or a
 * torch type across the ABI.  One plan per GPU; a plan is not re-entrant; different plans may be
 * driven from different host threads.  There is NO CPU fallback: plan creation fails without a
 * CUDA device.
 */
#ifndef TWSTFT_H
#define TWSTFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TWSTFT_VARIANT_G 0 /* godual_ranging.m: fcode=conj(fft(code)), P=ffty.*fcode           */
#define TWSTFT_VARIANT_K 1 /* claudio_aligned_code_*: fcode=fft(code), P=fcode.*conj(ffty)     */

#define TWSTFT_BPSK 0 /* carrier from fft(x.^2), df=freq/2 */
#define TWSTFT_QPSK 1 /* carrier from fft(x.^4), df=freq/4 (Doc/07_Demodulation.md:60-63) */

#define TWSTFT_LAYOUT_INTERLEAVED2 0 /* [I1 Q1 I2 Q2] int16 per sample instant (b210_file.py:60-65)   */
#define TWSTFT_LAYOUT_SINGLE 1       /* [I Q] int16, one channel per file (rx_multi_samples.cpp:214)  */

#define TWSTFT_AXIS_G 0 /* freq=linspace(-fs/2,fs/2,n)      godual_ranging.m:73   */
#define TWSTFT_AXIS_K 1 /* freq=linspace(-fs/2,fs/2-1,n)    lo_separate.m:103     */

#define TWSTFT_CODEWIN_NONE 0    /* fcode = conj(fft(code))                      godual_ranging.m:66          */
#define TWSTFT_CODEWIN_HAMMING 1 /* fcode = conj(fft(code)) .* hamming(N) over bins processing/CPP/main.cpp:717-719 */

#define TWSTFT_OK 0
#define TWSTFT_E_ARG (-1)
#define TWSTFT_E_NODEVICE (-2)
#define TWSTFT_E_SIZE (-3)   /* unsupported window length */
#define TWSTFT_E_CUDA (-4)
#define TWSTFT_E_NOCODE (-5)
#define TWSTFT_E_NOMEM (-6)

typedef struct twstft_plan twstft_plan;

typedef struct twstft_config {
    double fs;          /* sample rate, 5e6 */
    int64_t nsamples;   /* N: samples per correlation window = samples_per_chip * code length */
    int samples_per_chip; /* fs/fchip, 2 */
    int variant;        /* TWSTFT_VARIANT_* */
    int modulation;     /* TWSTFT_BPSK / TWSTFT_QPSK */
    int nint;           /* Nint: 1 -> x3 interpolation (the only value the reference uses), 0 -> none */
    int layout;         /* TWSTFT_LAYOUT_* */
    int freq_axis;      /* TWSTFT_AXIS_* */
    double win_lo_hz;   /* estimator window k=find((freq<hi)&(freq>lo)), strict */
    double win_hi_hz;
    int est_sub_mean;   /* 1: estimator squares the mean-removed window (G), 0: raw (K) */
    int ddof;           /* 1: Octave var (n-1), 0: numpy/C++ var (n) */
    int max_batch;      /* records (or windows) per device batch */
    int device_id;
    int fine_freq;      /* 1: fine-frequency refinement before the correlation (twstft_process only): slope of the
                           100-tap smoothed angle(y(1:10:fs/3)), df += slope/2pi --
                           experiments/221219_twoway/processing/godual_ranging.py:26-30 (commented out at
                           processing/Octave/godual_ranging.m:19-24); 0 (default): off */
    int code_window;    /* TWSTFT_CODEWIN_*: the C++ program correlates against a Hamming-windowed code spectrum but
                           despreads the SNR estimate with the plain code (processing/CPP/main.cpp:319-357,717-719);
                           with HAMMING the SNR/power fields therefore come from a second, unwindowed row pass around
                           the peak decided on the windowed correlation, and `puissance` is sum|fft(y)|^2
                           (main.cpp:295) instead of var(y).  Scales stay Octave's (ifft 1/n): FFTW's unnormalised
                           values are 3N times xval*, and the stored correction1 is indice-1+correction */
    int fast_peak;      /* 1: twstft_process / twstft_process_device / twstft_process_windows find the interpolated maximum from ONE inverse
                           transform plus exactly evaluated lags around its maximum, guarded by a rigorous bound on
                           every other interpolated sample; records that fail the bound (no correlation peak) are
                           re-run through the full three-transform path, so results are identical (k_fast.cuh).
                           Needs nint = 1 and code_window = NONE; +8N bytes of device memory per lane */
} twstft_config;

/* One row of the reference's result variables, per channel and record (all doubles). */
typedef struct twstft_result {
    double indice;        /* 1-based index of the maximum on the interpolated grid */
    double correction;    /* parabolic vertex offset in interpolated samples */
    double xval_re, xval_im;
    double xvalm1_re, xvalm1_im;
    double xvalp1_re, xvalp1_im;
    double SNRr, SNRi;
    double df;            /* carrier offset used for the derotation, Hz */
    double puissance, puissancecode, puissancenoise;
    double carrier_bin;   /* 0-based index of the estimator maximum on the fftshift'ed axis, -1 if df was given */
    double status;        /* bit0: peak at the array edge (Octave would fail), bit1: fp64 refinement moved the peak,
                             bit2: fast peak path could not certify the maximum (only ever visible inside the library:
                             twstft_process* re-run such records through the full path before returning) */
} twstft_result;

void twstft_default_config(twstft_config* cfg);
int twstft_plan_create(const twstft_config* cfg, twstft_plan** out);
void twstft_plan_destroy(twstft_plan* plan);
const char* twstft_last_error(const twstft_plan* plan);
const char* twstft_backend(void); /* "cuda" for the product library */

/* PRN replica. chips: 0/1 bytes; qpsk!=0 -> generator A and B bytes interleaved (common.py:66-71). */
int twstft_set_code(twstft_plan* plan, const uint8_t* chips, size_t nbytes, int qpsk);
/* Generate the LFSR code(s) on the GPU, install the replica, optionally return the chip bytes. */
int twstft_gen_code(twstft_plan* plan, int bitlen, uint32_t taps_a, uint32_t taps_b, uint32_t seed,
                    size_t noiselen, uint8_t* chips_out);
/* Stand-alone GPU LFSR (no plan state): out[noiselen] (taps_b==0) or out[2*noiselen] interleaved. */
int twstft_prn_generate(int device_id, int bitlen, uint32_t taps_a, uint32_t taps_b, uint32_t seed,
                        size_t noiselen, uint8_t* out);

/* Variant-G record loop: nrecords consecutive records of nch*N complex-int16 samples in HOST memory;
 * results[record*nch + ch].  Estimates the carrier per record and channel, then correlates. */
int twstft_process(twstft_plan* plan, const int16_t* iq, size_t nrecords, twstft_result* results);
/* Same with the records already resident in DEVICE memory. */
int twstft_process_device(twstft_plan* plan, const int16_t* iq_dev, size_t nrecords, twstft_result* results);

/* Variant-K style: nwin windows of N samples starting at sample instants offsets[i] of a (host or
 * device) buffer holding `total_instants` sample instants, derotated by the given df[i*nch+ch]. */
int twstft_process_windows(twstft_plan* plan, const int16_t* iq, int iq_on_device, size_t total_instants,
                           const int64_t* offsets, const double* df, size_t nwin, twstft_result* results);

/* The Octave FUNCTION seam.  `d`: nwin windows of N complex doubles (re, im interleaved: Octave's complex column),
 * already mean-removed by the caller exactly as the scripts do before the call (godual_ranging.m:80,94;
 * lo_separate.m:132); one channel per window, results[window].  The "globals" the Octave functions read
 * (temps freq fcode code fs Nint) are the plan's configuration and installed code.
 *   twstft_processing_cf64     variant G processing(d,k): k = k_first..k_last, the 1-based index range into the
 *                              fftshift'ed spectrum (k_first <= 0: the plan's own window); estimates df, correlates
 *   twstft_processing_cf64_df  variant K processing(d,k,df) / processing(d,df): df[window] given by the caller
 * (fine_freq and the fast peak path are not offered here.) */
int twstft_processing_cf64(twstft_plan* plan, const double* d, size_t nwin, int64_t k_first, int64_t k_last,
                           twstft_result* results);
int twstft_processing_cf64_df(twstft_plan* plan, const double* d, size_t nwin, const double* df, twstft_result* results);
/* k=search_df(d,k,df_threshold) (claudio_aligned_code_re_separate.m:26-46,150): `d` = the block of ls*est->nsamples
 * complex doubles as the script holds it (raw, not mean-removed), k_first..k_last 1-based on the block's fftshift'ed
 * axis (<= 0: est's window); outputs as twstft_search_df_block, one channel. */
int twstft_search_df_cf64(twstft_plan* est, twstft_plan* cor, const double* d, int ls, int64_t k_first, int64_t k_last,
                          double df_threshold, int64_t* kbon, double* df_found, int32_t* ncand);

/* Carrier estimator alone (A3): df[record*nch+ch], bins[...] = 0-based index on the shifted axis. */
int twstft_estimate_carrier(twstft_plan* plan, const int16_t* iq, int iq_on_device, size_t nrecords,
                            double* df, int64_t* bins);

/* One Doppler bin: [prnsig,b]=max(abs(ifft(fcode.*conj(fft(y))))) at 1x, metric = prnsig^2/var with b-5..b+5 zeroed. */
typedef struct twstft_bank_result {
    double b0;      /* 0-based index of the maximum (Octave b = b0+1) */
    double prnsig;  /* magnitude of the maximum (ifft scale 1/N) */
    double metric;  /* prnsig^2 / var(prnmap), the "snr" of search_df */
    double status;  /* bit0: b within 5 samples of the array edge (Octave's b-5:b+5 would fail; the zeroing wraps) */
} twstft_bank_result;

/* Cross-ambiguity search: the window of N sample instants starting at `offset` of a (host or device) buffer
 * is derotated by every dfs[i] and correlated against the code; results[i*nch + ch].
 * sub_mean != 0 removes the window mean first (the caller of processing() does, search_df does not). */
int twstft_doppler_bank(twstft_plan* plan, const int16_t* iq, int iq_on_device, size_t total_instants, int64_t offset,
                        const double* dfs, size_t nbins, int sub_mean, twstft_bank_result* results);

/* Doppler / code-phase acquisition with the step-halving grid of experiments/231001_DLL_PLL/rxcomplex.cpp:522-568
 * on top of the bank: pass 1 evaluates fc_init-range .. fc_init+range in steps of `step` (grid accumulated as
 * fcc += step), the best bin so far is replaced only by a strictly larger peak magnitude; then step /= 2, the range
 * becomes the new step around the best carrier, and so on until step < min_step (1 Hz in the reference).
 * fc_out: the carrier of the best bin, best_out: its bank row (peak index, magnitude, metric) for `channel`,
 * nbins_out: bins evaluated. */
int twstft_acquire(twstft_plan* plan, const int16_t* iq, int iq_on_device, size_t total_instants, int64_t offset,
                   double fc_init, double range, double step, double min_step, int sub_mean, int channel,
                   double* fc_out, twstft_bank_result* best_out, int32_t* nbins_out);

/* search_df on one window: candidates = bins of the plan's estimator window whose |fft(x^p)| exceeds
 * df_threshold*median; if there are 1..99 each is tested with the bank body and the LAST one whose metric
 * exceeds 100 wins.  kbon[ch] = its 1-based index on the fftshift'ed axis (0 = none), df_found[ch] = freq/p,
 * ncand[ch] = number of candidates.  (The reference runs it on 2-s blocks; here the block is the plan window.) */
int twstft_search_df(twstft_plan* plan, const int16_t* iq, int iq_on_device, size_t total_instants, int64_t offset,
                     double df_threshold, int64_t* kbon, double* df_found, int32_t* ncand);

/* search_df exactly as the re/ranging scripts call it (claudio_aligned_code_re_separate.m:26-46,150): the x^p
 * spectrum is taken over the whole block of ls*est->nsamples sample instants (ls = 2, est->nsamples = fs: the 2-s
 * block; window and freq axis of `est` on that block), every candidate is tested on the first cor->nsamples
 * instants of the block against cor's code.  `est` and `cor` may be the same plan (ls = 1).  Outputs as above,
 * kbon 1-based on the block axis. */
int twstft_search_df_block(twstft_plan* est, twstft_plan* cor, const int16_t* iq, int iq_on_device, int ls,
                           double df_threshold, int64_t* kbon, double* df_found, int32_t* ncand);

/* d2=fftshift(abs(fft(d.^p))) evaluated at nbins (<= 8) adjacent 0-based positions s_first.. of the shifted axis,
 * for a buffer of ANY length ninstants (the re/ranging scripts take it over d=[dold ; block] and look at
 * kbon-3:kbon+3 only, claudio_aligned_code_re_separate.m:166-168): direct fp64 evaluation, one pass over the
 * buffer.  mag[ch*nbins + i]. */
int twstft_estimate_bins(twstft_plan* plan, const int16_t* iq, int iq_on_device, size_t ninstants, int64_t s_first,
                         int nbins, double* mag);

/* The same estimator on nblocks blocks of ls*N sample instants.  ls = 2 with N = fs is the 2-s block of
 * acquisition/claudio_aligned_code_lo_separate.m:120-129 (freq=linspace(-fs/2,fs/2-fs/fs,fs*ls), raw samples
 * squared, window on that axis): the first radix-2 split of the 2N-point transform is fused into the loads. */
int twstft_estimate_carrier_block(twstft_plan* plan, const int16_t* iq, int iq_on_device, size_t nblocks, int ls,
                                  double* df, int64_t* bins);

/* Whole-file carrier estimator of the C++ program (GoRanging::df, processing/CPP/main.cpp:363-450) on a HOST
 * buffer of total_instants sample instants in the plan's layout: one instant out of `dec` (25), derotated by
 * foffset, raw mean removed, squared, ONE transform of L = total_instants/dec points per channel;
 * freq = linspace(-fs/2/dec, fs/2/dec, L); channel 1 searches [kmin,kmax) (|f| < 16 kHz), channel 2 (remote == 0)
 * every bin; df[ch] = freq[pos]/2 + foffset, pos[ch] = index on the fftshift'ed axis (may be NULL).
 * The plan's window length M must divide L (L = s*M: s transforms of M points on the plan's kernels plus a direct
 * s-point stage); TWSTFT_E_SIZE otherwise. */
int twstft_estimate_decimated(twstft_plan* plan, const int16_t* iq, size_t total_instants, int dec, double foffset,
                              int remote, double* df, int64_t* pos);

/* ---- multi-GPU: records are independent (godual_ranging.m:75-102 keeps no state between records), so
 * a run of records is split by index, GPU g of G taking [g*R/G, (g+1)*R/G): one host thread + one plan per GPU, the
 * code spectrum replicated, rows written at their record index of the caller's array.  No collective, no NCCL. ---- */
typedef struct twstft_shard twstft_shard;
int twstft_shard_create(const twstft_config* cfg, const int* devices /* NULL: 0..ngpus-1 */, int ngpus, twstft_shard** out);
void twstft_shard_destroy(twstft_shard* s);
int twstft_shard_ngpus(const twstft_shard* s);
twstft_plan* twstft_shard_plan(twstft_shard* s, int g);          /* e.g. to twstft_gen_code on plan 0 and read the chips */
const char* twstft_shard_last_error(const twstft_shard* s);
int twstft_shard_set_code(twstft_shard* s, const uint8_t* chips, size_t nbytes, int qpsk);
/* the record loop of godual_ranging.m:75-102 / processing/CPP/main.cpp:452-519 over nrecords HOST records */
int twstft_shard_process(twstft_shard* s, const int16_t* iq, size_t nrecords, twstft_result* results);
/* One plan: nrecords records first_record.. drawn cyclically from a pool (record r = pool[r % pool_records]), pool in
 * host or device memory.  The day-of-data driver (BASELINE configs[4]) when a day does not fit host memory. */
int twstft_process_cycled(twstft_plan* plan, const int16_t* pool, int pool_on_device, size_t pool_records,
                          size_t first_record, size_t nrecords, twstft_result* results);
/* The same over all GPUs of a shard.  resident != 0: every GPU first uploads the pool (untimed) and correlates from
 * HBM; 0: every record is streamed from a per-GPU pinned copy of the pool allocated by the thread bound to that GPU's
 * CPUs.  seconds_wall: host wall clock of the parallel region; seconds_dev[g]: CUDA-event time on GPU g's stream. */
int twstft_shard_process_cycled(twstft_shard* s, const int16_t* pool, size_t pool_records, size_t nrecords, int resident,
                                twstft_result* results, double* seconds_wall, double* seconds_dev);

/* ---- synthetic inputs generated on the device (bench / test tooling; SURVEY.md 8d recipe: delayed code * carrier
 * + AWGN at a chosen SNR, round + saturate to int16; Philox noise keyed by the record seed). ----
 * twstft_synth_records: nrecords records in the plan's layout into DEVICE memory; record r uses seed0 + r; channel 1
 *   delay 0.25..0.99 of the window, df +-9 kHz, SNR snr1_db; channel 2 (two-channel layout) the loopback: delay
 *   ~0.7253 of the window, df 0, SNR snr2_db.
 * twstft_synth_stream: one single-channel recording of ninstants sample instants, the installed code repeating back
 *   to back (rx_multi_samples.cpp:214-215 files of the K scripts). */
int twstft_synth_records(twstft_plan* plan, uint64_t seed0, size_t nrecords, double snr1_db, double snr2_db,
                         int16_t* dev_out);
int twstft_synth_stream(twstft_plan* plan, uint64_t seed, size_t ninstants, double delay, double df, double snr_db,
                        int16_t* dev_out);

/* Pinned host memory for the end-to-end path, device memory helpers for device-resident inputs. */
void* twstft_host_alloc(size_t bytes);
void twstft_host_free(void* p);
int twstft_device_alloc(twstft_plan* plan, size_t bytes, void** out);
int twstft_device_free(twstft_plan* plan, void* p);
int twstft_copy_to_device(twstft_plan* plan, void* dst_dev, const void* src_host, size_t bytes);
int twstft_copy_to_host(twstft_plan* plan, void* dst_host, const void* src_dev, size_t bytes);
int twstft_synchronize(twstft_plan* plan);

/* Device-side timing on the plan's own stream (CUDA events), used by bench.py.
 * twstft_profile_*: one event pair per kernel launch, accumulated per kernel type. */
int twstft_timer_start(twstft_plan* plan);
int twstft_timer_stop(twstft_plan* plan, float* ms);
int twstft_profile_enable(twstft_plan* plan, int on);
int twstft_profile_read(twstft_plan* plan, int slot, char* name, size_t name_cap, double* total_ms, int64_t* count);

/* Introspection used by tests and the benchmark. */
int twstft_fast_counters(const twstft_plan* plan, int64_t* records, int64_t* rerun); /* fast peak path: records seen / re-run in full */
/* The fast peak path only pays while few records fall back (an uncertified record costs the fast attempt plus the full
 * path): when more than a quarter of the last 64 records fell back the plan runs the next 1024 through the full path
 * and then probes again.  Returns how many records went through the full path for that reason.  Rows never change. */
int64_t twstft_fast_suspended(const twstft_plan* plan);
int64_t twstft_kernel_launches(const twstft_plan* plan);   /* kernels launched so far by this plan */
int twstft_debug_fetch(twstft_plan* plan, int what, int lane, void* host_out, size_t bytes);
int twstft_window_bins(const twstft_plan* plan, int64_t* lo, int64_t* hi);

#ifdef __cplusplus
}
#endif
#endif /* TWSTFT_H */
