/* godual_cpp -- plain-C front end over libtwstft with the behaviour of the reference's C++ program
 * processing/CPP/main.cpp (GoRanging): same command line, same stdout rows, same <name>C.mat variables.
 *
 *   godual_cpp data.bin code.bin [remote=0] [foffset=0.]                      main.cpp:773-808
 *
 * data.bin : int16 LE [I1 Q1 I2 Q2] per sample instant; code.bin : one byte (0/1) per chip, N = 2 * chips.
 * 1. GoRanging::df (:363-450): whole-file carrier estimate on one instant out of 25 -> df1 (and df2 unless remote);
 *    on the GPU through twstft_estimate_decimated.  The decimated length must be a multiple of a transform length
 *    the library supports (whole 1-s records at 5 MS/s always are: 200 000 points per record).
 * 2. GoRanging::compute / _process_method (:224-361,452-519): per record and channel mean removal, derotation by
 *    the file-wide df, correlation against conj(fft(code)).*hamming, x3 interpolation, 0-based peak + parabola,
 *    SNR despread with the plain code, var/n.  twstft_process_windows on a plan with code_window = HAMMING, ddof = 0.
 * 3. GoRanging::save (:521-656): <dir>/[remote]<name>C.mat with [P x 1] variables correction1 (= indice0+corr), SNR1
 *    (dB), df1, puissance1 (sum|fft(y)|^2), puissance1code (dB), xval1, xval1m1, xval1p1 (FFTW-unnormalised: x 3N),
 *    and the same with 2 unless remote.
 * A short trailing record prints "No more data" and stops (:467-470).  Exit status: 0, or 1 on an error (the C++
 * throws std::runtime_error on open failures, :98-110).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/twstft.h"
#include "matv5.h"

static void* slurp(const char* path, size_t* bytes)
{
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* p = malloc(n > 0 ? (size_t)n : 1);
    if (p && fread(p, 1, (size_t)n, f) != (size_t)n) { free(p); p = NULL; }
    fclose(f);
    *bytes = (size_t)n;
    return p;
}

int main(int argc, char** argv)
{
    const double fs = 5e6;
    const int dec = 25;                       /* main.cpp:776: 5 MS/s decimated by 25 = 200 kS/s */
    printf("%s data.bin code.bin [remote=0] [foffset=0.]\n", argv[0]);
    if (argc < 3) return 1;
    const int remote = argc >= 4 ? atoi(argv[3]) : 0;
    const double foffset = argc >= 5 ? atof(argv[4]) : 0.0;
    size_t nchips = 0, nbytes = 0;
    uint8_t* chips = (uint8_t*)slurp(argv[2], &nchips);
    if (!chips || nchips == 0) { fprintf(stderr, "cannot read code file %s\n", argv[2]); return 1; }
    int16_t* data = (int16_t*)slurp(argv[1], &nbytes);
    if (!data) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    const size_t N = 2 * nchips;
    const size_t instants = nbytes / (4 * sizeof(int16_t));
    printf("file size : %zu %zu\n", N, N);    /* fill_fcode, main.cpp:714 */

    /* 1. whole-file estimator on a plan whose window length divides the decimated length */
    static const int64_t sizes[] = {5000000, 1000000, 250000, 200000, 10000};
    const size_t L = instants / dec;
    int64_t M = 0;
    for (size_t i = 0; i < sizeof sizes / sizeof sizes[0]; i++)
        if (L >= (size_t)sizes[i] && L % (size_t)sizes[i] == 0 && L / (size_t)sizes[i] <= 32768) { M = sizes[i]; break; }
    if (!M) { fprintf(stderr, "decimated length %zu is not a multiple of a supported transform length\n", L); return 1; }
    twstft_config ecfg;
    twstft_default_config(&ecfg);
    ecfg.fs = fs; ecfg.nsamples = M; ecfg.max_batch = M >= 1000000 ? 4 : 32;
    twstft_plan* est = NULL;
    if (twstft_plan_create(&ecfg, &est) != TWSTFT_OK) {
        fprintf(stderr, "twstft_plan_create: %s\n", twstft_last_error(est));
        return 1;
    }
    double df[2] = {0.0, 0.0};
    if (twstft_estimate_decimated(est, data, instants, dec, foffset, remote, df, NULL) != TWSTFT_OK) {
        fprintf(stderr, "twstft_estimate_decimated: %s\n", twstft_last_error(est));
        return 1;
    }
    twstft_plan_destroy(est);
    printf("df1=%.3f\n", df[0]);
    if (!remote) printf(" df2=%.3f\n", df[1]); else printf("\n");

    /* 2. record loop */
    twstft_config cfg;
    twstft_default_config(&cfg);
    cfg.fs = fs; cfg.nsamples = (int64_t)N; cfg.ddof = 0; cfg.code_window = TWSTFT_CODEWIN_HAMMING;
    cfg.max_batch = N >= 5000000 ? 8 : 32;
    twstft_plan* plan = NULL;
    if (twstft_plan_create(&cfg, &plan) != TWSTFT_OK || twstft_set_code(plan, chips, nchips, 0) != TWSTFT_OK) {
        fprintf(stderr, "plan: %s\n", twstft_last_error(plan));
        return 1;
    }
    const size_t nrec = instants / N;
    const int nuse = remote ? 1 : 2;
    double* col[2][11];   /* correction SNR df puissance puissancecode xval(re,im) xvalm1(re,im) xvalp1(re,im) */
    for (int c = 0; c < 2; c++)
        for (int v = 0; v < 11; v++) col[c][v] = (double*)malloc(sizeof(double) * (nrec ? nrec : 1));
    twstft_result* res = (twstft_result*)malloc(sizeof(twstft_result) * 2 * cfg.max_batch);
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * cfg.max_batch);
    double* dfw = (double*)malloc(sizeof(double) * 2 * cfg.max_batch);
    const double m3 = 3.0 * (double)N;
    size_t np = 0;
    for (size_t r0 = 0; r0 < nrec; r0 += cfg.max_batch) {
        const size_t nb = nrec - r0 < (size_t)cfg.max_batch ? nrec - r0 : (size_t)cfg.max_batch;
        for (size_t i = 0; i < nb; i++) { off[i] = (int64_t)((r0 + i) * N); dfw[2 * i] = df[0]; dfw[2 * i + 1] = df[1]; }
        if (twstft_process_windows(plan, data, 0, instants, off, dfw, nb, res) != TWSTFT_OK) {
            fprintf(stderr, "twstft_process_windows: %s\n", twstft_last_error(plan));
            return 1;
        }
        for (size_t i = 0; i < nb; i++, np++) {
            for (int c = 0; c < nuse; c++) {
                const twstft_result* x = &res[2 * i + c];
                const double corr = x->indice - 1.0 + x->correction;            /* indice1 + corr1, main.cpp:310 */
                const double snr = 10.0 * log10(x->SNRr + x->SNRi);
                const double vals[11] = {corr, snr, x->df, x->puissance, 10.0 * log10(x->puissancecode * m3 * m3),
                                         x->xval_re * m3, x->xval_im * m3, x->xvalm1_re * m3, x->xvalm1_im * m3,
                                         x->xvalp1_re * m3, x->xvalp1_im * m3};
                for (int v = 0; v < 11; v++) col[c][v][np] = vals[v];
                printf("%zu/%d %0.12lf\t%.3f\t", np, c, corr / fs / 3.0, x->df);   /* :315 */
                printf("%.1lf\t%0.1lf\t", x->puissance, snr);                     /* :353 */
            }
            printf("\n");
        }
    }
    printf("No more data\n");

    /* 3. <dir>/[remote]<name>C.mat  (main.cpp:786-798: ".bin" -> "C.mat") */
    if (np == 0) { printf("Nothing to save\n"); return 0; }
    char out[4096];
    const char* slash = strrchr(argv[1], '/');
    const size_t dirlen = slash ? (size_t)(slash - argv[1] + 1) : 0;
    snprintf(out, sizeof out, "%.*s%s%s", (int)dirlen, argv[1], remote ? "remote" : "", argv[1] + dirlen);
    char* dot = strrchr(out, '.');
    if (dot && !strcmp(dot, ".bin")) strcpy(dot, "C.mat");
    else strncat(out, "C.mat", sizeof out - strlen(out) - 1);
    FILE* m = fopen(out, "wb");
    if (!m || mat5_open(m, "C++ program result record") != 0) { printf("mat file creation: FAIL\n"); return 1; }
    static const char* real_names[2][5] = {{"correction1", "SNR1", "df1", "puissance1", "puissance1code"},
                                           {"correction2", "SNR2", "df2", "puissance2", "puissance2code"}};
    static const char* cplx_names[2][3] = {{"xval1", "xval1m1", "xval1p1"}, {"xval2", "xval2m1", "xval2p1"}};
    for (int c = 0; c < nuse; c++) {
        for (int v = 0; v < 5; v++) mat5_matrix(m, real_names[c][v], col[c][v], NULL, np, 1);
        for (int v = 0; v < 3; v++) mat5_matrix(m, cplx_names[c][v], col[c][5 + 2 * v], col[c][6 + 2 * v], np, 1);
    }
    fclose(m);
    fprintf(stderr, "%zu records -> %s\n", np, out);
    twstft_plan_destroy(plan);
    return 0;
}
