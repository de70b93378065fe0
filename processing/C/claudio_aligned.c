/* claudio_aligned -- plain-C front end over libtwstft (include/twstft.h): the production windowing / tracking
 * loops of the three scripts acquisition/goprocess.sh:9-11 starts, on ONE single-channel X310 recording:
 *
 *   claudio_aligned MODE data_N.bin code.bin [OP=0] [outdir=.] [fs=5e6] [skip_s=30] [batch=32]
 *
 *   MODE  lo         acquisition/claudio_aligned_code_lo_separate.m:79-172, remote=0   -> localclaudio<name>.mat
 *         lo-remote  the same script with remote=1 (:107-111)                          -> remoteclaudio<name>.mat
 *         re         acquisition/claudio_aligned_code_re_separate.m:103-219, remote=1  -> remoteclaudio<name>.mat
 *         ranging    acquisition/claudio_aligned_code_ranging_separate.m (ranging=1)   -> rangingclaudio<name>.mat
 *   data  int16 LE [I Q] per sample instant (acquisition/rx_multi_samples.cpp:214-215), read in ls = 2 s blocks
 *   code  one byte (0/1) per chip (amaranth_twstft/common.py:59-73); window N = 2 * chips
 *   OP    1 selects the negative-frequency estimator window (lo :108, re :136)
 *   fs, skip_s  the scripts hard-code 5e6 and fseek(f,30*fs*2*2) (re :127); the tests run scaled-down recordings
 *
 * lo:  df(block) = first max of |fft(block.^2)| inside the window of freq=linspace(-fs/2,fs/2-1,fs*ls) (:103,126-129)
 *      windows of d=[dold;block] at dindex, indice1=floor(i/3), MOVED re-slice (:130-153), tail -> dold (:155-159)
 * re/ranging:  kbon=search_df(block at 30 s) (:148-151), re-open, then per block df = max over kbon-3:kbon+3 of
 *      |fft(d.^2)|, d=[dold;block] (:165-167); indice1=i/3 un-floored, slices at round(dindex) (:170-185)
 * Output: the stdout table of the scripts (header :132, one row per window :192, "MOVED" lines :176) and the MAT-v5
 * file with the variables of `save -mat nomout corr* df indic* SNR* code puissan* xval* moved*` (lo :166, re :204)
 * in the order and shapes of the reference's own result files (experiments/240527/op/ *.mat.gz).
 * Windows go to the GPU in speculative batches (moves are rare); a batch is cut at the first MOVED window.
 * Deviations, all reported on stderr: a re-slice that leaves the block is an Octave index error in the scripts, here
 * the block is abandoned; the re script never ends while the carrier is not found (it re-opens the file on every
 * iteration, :152-154), here the search goes on with blocks 1, 2, ... and exits with status 2 if none holds it.
 * Exit status 0 on success, 1 on any error, 2 when re/ranging finds no carrier.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/twstft.h"
#include "matv5.h"

typedef struct { double* v; size_t n, cap; } vec;
static void push(vec* a, double x)
{
    if (a->n == a->cap) {
        a->cap = a->cap ? 2 * a->cap : 1024;
        a->v = (double*)realloc(a->v, sizeof(double) * a->cap);
    }
    a->v[a->n++] = x;
}
static long long oround(double x) { return (long long)floor(x + 0.5); }     /* Octave round(), positive arguments */

static int die(twstft_plan* p, const char* what)
{
    fprintf(stderr, "%s: %s\n", what, p ? twstft_last_error(p) : "");
    return 1;
}

/* Octave's printf("%d") prints a non-integer value in %f/%g style */
static void print_moved(double v)
{
    if (v == floor(v)) printf("MOVED %lld\n", (long long)v);
    else printf("MOVED %g\n", v);
}

int main(int argc, char** argv)
{
    if (argc < 4) {
        fprintf(stderr, "%s lo|lo-remote|re|ranging data.bin code.bin [OP=0] [outdir=.] [fs=5e6] [skip_s=30] [batch=32]\n", argv[0]);
        return 1;
    }
    const char* mode = argv[1];
    const int OP = argc > 4 ? atoi(argv[4]) : 0;
    const char* outdir = argc > 5 ? argv[5] : ".";
    const double fs = argc > 6 ? atof(argv[6]) : 5e6;
    const double skip_s = argc > 7 ? atof(argv[7]) : 30.0;
    const int batch = argc > 8 ? atoi(argv[8]) : 32;
    const int ls = 2;
    int search = 0;
    double wlo, whi;
    const char* prefix;
    if (!strcmp(mode, "lo")) { wlo = -20000; whi = 20000; prefix = "localclaudio"; }
    else if (!strcmp(mode, "lo-remote")) {
        if (OP == 1) { wlo = -120000; whi = -80000; } else { wlo = 100000; whi = 120000; }
        prefix = "remoteclaudio";
    } else if (!strcmp(mode, "re")) {
        if (OP == 1) { wlo = -108000; whi = -92000; } else { wlo = 92000; whi = 108000; }
        prefix = "remoteclaudio"; search = 1;
    } else if (!strcmp(mode, "ranging")) { wlo = -8000; whi = 8000; prefix = "rangingclaudio"; search = 1; }
    else { fprintf(stderr, "unknown mode %s\n", mode); return 1; }

    /* code=fread(f,inf,'int8'); repelems x2; 2*code-1 (lo :92-95) */
    FILE* cf = fopen(argv[3], "rb");
    if (!cf) { fprintf(stderr, "cannot read code file %s\n", argv[3]); return 1; }
    fseek(cf, 0, SEEK_END);
    const long nchips = ftell(cf);
    fseek(cf, 0, SEEK_SET);
    uint8_t* chips = (uint8_t*)malloc(nchips > 0 ? (size_t)nchips : 1);
    if (nchips <= 0 || fread(chips, 1, (size_t)nchips, cf) != (size_t)nchips) { fprintf(stderr, "cannot read code file %s\n", argv[3]); return 1; }
    fclose(cf);

    twstft_config ce, cc;
    twstft_default_config(&ce);
    ce.fs = fs; ce.nsamples = (int64_t)fs; ce.variant = TWSTFT_VARIANT_K; ce.layout = TWSTFT_LAYOUT_SINGLE;
    ce.freq_axis = TWSTFT_AXIS_K; ce.win_lo_hz = wlo; ce.win_hi_hz = whi; ce.est_sub_mean = 0; ce.max_batch = 1;
    cc = ce;
    cc.nsamples = (int64_t)nchips * cc.samples_per_chip; cc.max_batch = batch > 0 ? batch : 32;
    cc.win_lo_hz = -20000; cc.win_hi_hz = 20000;
    cc.fast_peak = 0;      /* the tracking loop parks the correlation peak ~20 samples into the window (:142), where
                              the fast path's local segment would wrap around the array: every window would fall back
                              (measured: the governor suspends it after 64 windows), so the full path is used directly */
    twstft_plan *est = NULL, *cor = NULL;
    if (twstft_plan_create(&ce, &est) != TWSTFT_OK) return die(est, "twstft_plan_create (block estimator)");
    if (twstft_plan_create(&cc, &cor) != TWSTFT_OK) return die(cor, "twstft_plan_create (correlator)");
    if (twstft_set_code(cor, chips, (size_t)nchips, 0) != TWSTFT_OK) return die(cor, "twstft_set_code");

    const long long N = cc.nsamples;                       /* length(fcode) */
    const long long blk_inst = (long long)fs * ls;         /* sample instants per block */
    const size_t blk_bytes = (size_t)blk_inst * 4;
    FILE* f = fopen(argv[2], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
    /* d holds [dold ; block]: dold is shorter than one window plus the MOVED slack */
    const size_t dcap_inst = (size_t)blk_inst + 2 * (size_t)N + 64;
    int16_t* d = (int16_t*)twstft_host_alloc(dcap_inst * 4);
    if (!d) { fprintf(stderr, "pinned allocation failed\n"); return 1; }
    void* dev = NULL;
    if (twstft_device_alloc(cor, dcap_inst * 4, &dev) != TWSTFT_OK) return die(cor, "twstft_device_alloc");

    /* frequency axis freq=linspace(-fs/2,fs/2-fs/fs,fs*ls) (:103): value of 0-based index s */
    const long long nax = blk_inst;
    const double f0 = -fs / 2, flast = fs / 2 - 1.0, fstep = (flast - f0) / (double)(nax - 1);

    int64_t kbon = 0;
    if (search) {     /* re :127,148-154 */
        const long long skip_bytes = (long long)(skip_s * fs) * 4;      /* fseek(f,30*fs*2*2) */
        for (long long b = 0;; b++) {                                   /* b = 0: the block at skip_s; then blocks 1, 2, ... */
            const long long start = b == 0 ? skip_bytes : b * (long long)blk_bytes;
            if (fseek(f, (long)start, SEEK_SET) != 0 || fread(d, 1, blk_bytes, f) != blk_bytes) {
                if (b == 0) continue;                                   /* recording shorter than skip_s + one block */
                break;                                                  /* ran out of blocks */
            }
            double dff = 0;
            int32_t nc = 0;
            if (twstft_search_df_block(est, cor, d, 0, ls, 20.0, &kbon, &dff, &nc) != TWSTFT_OK)
                return die(est, "twstft_search_df_block");
            fprintf(stderr, "search_df at byte %lld: %d candidates, kbon=%lld (%.3f Hz)\n", start, (int)nc, (long long)kbon, dff);
            if (kbon) break;
        }
        if (!kbon) { fprintf(stderr, "carrier not found in any block\n"); return 2; }
        fseek(f, 0, SEEK_SET);
    }

    vec xr = {0}, xi = {0}, ind = {0}, cor1 = {0}, snrr = {0}, snri = {0}, pui = {0}, dfv = {0}, moved = {0}, movedval = {0};
    double pcode = NAN, pnoise = NAN;
    const int B = cc.max_batch;
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * B);
    double* dfl = (double*)malloc(sizeof(double) * B);
    twstft_result* res = (twstft_result*)malloc(sizeof(twstft_result) * B);
    size_t nold = 0;                                       /* sample instants carried in d[0..) */
    printf("n\tdt1\tdf1\tP1\tSNR1\tdt2\tdf2\tP2\tSNR2\r\n");
    for (;;) {
        if (fread(d + 2 * nold, 1, blk_bytes, f) != blk_bytes) break;     /* whole blocks only (:119) */
        const long long L = (long long)nold + blk_inst;
        double df;
        if (twstft_copy_to_device(cor, dev, d, (size_t)L * 4) != TWSTFT_OK) return die(cor, "twstft_copy_to_device");
        if (search) {      /* re :166-167 on [dold ; block] */
            double mag[7];
            if (twstft_estimate_bins(est, (const int16_t*)dev, 1, (size_t)L, kbon - 4, 7, mag) != TWSTFT_OK)
                return die(est, "twstft_estimate_bins");
            int best = 0;
            for (int i = 1; i < 7; i++)
                if (mag[i] > mag[best]) best = i;
            const long long s = kbon - 4 + best;
            df = (s == nax - 1 ? flast : f0 + (double)s * fstep) / 2;
        } else {           /* lo :126,129 on the raw block, before dold is prepended */
            int64_t bin;
            if (twstft_estimate_carrier_block(est, (const int16_t*)dev + 2 * nold, 1, 1, ls, &df, &bin) != TWSTFT_OK)
                return die(est, "twstft_estimate_carrier_block");
        }
        push(&dfv, df);
        double dindex = 1.0;
        int ok = 1;
        while (ok) {
            long long nwin = (long long)floor(((double)(L - N + 1) - dindex) / (double)N) + 1;
            if (nwin > B) nwin = B;
            if (nwin < 1) nwin = 1;
            for (long long t = 0; t < nwin; t++) { offs[t] = oround(dindex + (double)(t * N)) - 1; dfl[t] = df; }
            while (nwin > 1 && offs[nwin - 1] + N > L) nwin--;
            if (offs[0] + N > L) { fprintf(stderr, "block %zu shorter than one window\n", dfv.n); break; }
            if (twstft_process_windows(cor, (const int16_t*)dev, 1, (size_t)L, offs, dfl, (size_t)nwin, res) != TWSTFT_OK)
                return die(cor, "twstft_process_windows");
            for (long long t = 0; t < nwin; t++) {
                const twstft_result* r = &res[t];
                twstft_result r2;
                double i1 = r->indice / 3.0;                                   /* re :173 */
                if (!search) i1 = floor(i1);                                   /* lo :134 */
                const double snr = 10 * log10(r->SNRi + r->SNRr);
                double cur = dindex + (double)(t * N);
                int was_moved = 0;
                if (snr > -30 && ((i1 > 43 && i1 < N / 2.0) || (i1 < N - 2 && i1 > N / 2.0))) {   /* :135-136 */
                    print_moved(i1);
                    push(&moved, (double)(ind.n + 1));
                    push(&movedval, i1 + 1);
                    if (cur - i1 + 1 < 0) cur += (double)N;
                    cur = cur - i1 + 21;                                       /* :142 (230802) */
                    const int64_t o2 = oround(cur) - 1;
                    if (o2 < 0 || o2 + N > L) {
                        fprintf(stderr, "window %zu: re-slice leaves the block, block abandoned\n", ind.n + 1);
                        ok = 0;
                        break;
                    }
                    if (twstft_process_windows(cor, (const int16_t*)dev, 1, (size_t)L, &o2, &df, 1, &r2) != TWSTFT_OK)
                        return die(cor, "twstft_process_windows");
                    r = &r2;
                    i1 = r2.indice;                                            /* :145 raw index kept */
                    was_moved = 1;
                }
                push(&xr, r->xval_re); push(&xi, r->xval_im); push(&ind, i1); push(&cor1, r->correction);
                push(&snrr, r->SNRr); push(&snri, r->SNRi); push(&pui, r->puissance);
                pcode = r->puissancecode; pnoise = r->puissancenoise;
                printf("%zu\t%.12f\t%.3f\t%.1f\t%.1f\r\n", ind.n, (i1 - 1 + r->correction) / fs / 3, df,
                       10 * log10(r->puissance), 10 * log10(r->SNRi + r->SNRr));                  /* :150 */
                if (was_moved || t == nwin - 1) { dindex = cur + (double)N; break; }
            }
            if (dindex + (double)N - 1 > (double)L) break;                     /* :153 */
        }
        if (dindex < (double)L) {                                              /* :155-159 */
            const long long a = oround(dindex) - 1;
            nold = (size_t)(L - a);
            memmove(d, d + 2 * a, nold * 4);
        } else nold = 0;
        if (nold + (size_t)blk_inst > dcap_inst) { fprintf(stderr, "carry-over longer than expected\n"); return 1; }
    }
    fclose(f);

    /* nom=strrep(name,'.bin','.mat'); nomout=[prefix,nom] (lo :84-90) */
    char out[4096];
    const char* slash = strrchr(argv[2], '/');
    const char* base = slash ? slash + 1 : argv[2];
    snprintf(out, sizeof out, "%s/%s%s", outdir, prefix, base);
    char* dot = strrchr(out, '.');
    if (dot && !strcmp(dot, ".bin")) strcpy(dot, ".mat");
    else strncat(out, ".mat", sizeof out - strlen(out) - 1);
    FILE* m = fopen(out, "wb");
    if (!m || mat5_open(m, "variant K result record") != 0) { fprintf(stderr, "cannot write %s\n", out); return 1; }
    mat5_vector(m, "correction1", cor1.v, NULL, cor1.n);
    mat5_vector(m, "df", dfv.v, NULL, dfv.n);
    mat5_vector(m, "indice1", ind.v, NULL, ind.n);
    mat5_vector(m, "SNR1i", snri.v, NULL, snri.n);
    mat5_vector(m, "SNR1r", snrr.v, NULL, snrr.n);
    double* code = (double*)malloc(sizeof(double) * (size_t)N);
    for (long long i = 0; i < N; i++) code[i] = 2.0 * chips[i / cc.samples_per_chip] - 1.0;
    mat5_vector(m, "code", code, NULL, (size_t)N);
    mat5_vector(m, "puissance1", pui.v, NULL, pui.n);
    mat5_vector(m, "puissancecode", &pcode, NULL, 1);
    mat5_vector(m, "puissancenoise", &pnoise, NULL, 1);
    mat5_vector(m, "xval1", xr.v, xi.v, xr.n);
    mat5_vector(m, "moved", moved.v, NULL, moved.n);
    mat5_vector(m, "movedval", movedval.v, NULL, movedval.n);
    fclose(m);
    fprintf(stderr, "%zu windows, %zu blocks, %zu moved -> %s\n", ind.n, dfv.n, moved.n, out);
    twstft_device_free(cor, dev);
    twstft_host_free(d);
    twstft_plan_destroy(cor);
    twstft_plan_destroy(est);
    return 0;
}
