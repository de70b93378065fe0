/* godual -- plain-C front end over libtwstft (include/twstft.h): the record loop of
 * processing/Octave/godual_ranging.m:57-130 (variant G) on a B210 two-channel recording.
 *
 *   godual data.bin code.bin [remote=0] [batch=8]
 *
 * data.bin : int16 LE [I1 Q1 I2 Q2] per sample instant (processing/Python/b210_file.py:60-65)
 * code.bin : one byte (0/1) per chip (amaranth_twstft/common.py:59-73); N = 2 * chips
 * remote   : 0 = local file, both channels, carrier window |f| < 20 kHz          (godual_ranging.m:82-83)
 *            1 = remote file, channel 1 only is reported, window 80..120 kHz     (:84-90, OP != 1)
 * Output   : one stdout row per record in the format of godual_ranging.m:96-98 and a MAT-v5 file
 *            <dir>/[remote]<name>.mat with the variables of :128-130 (corr*, df*, indic*, SNR*, code, puissan*).
 * The command-line shape (positional data/code/remote) is that of processing/CPP/main.cpp:779-798.
 * A trailing partial record is ignored (:81,102).  Exit status 0 on success, 1 on any error.
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
    if (argc < 3) {
        fprintf(stderr, "%s data.bin code.bin [remote=0] [batch=8]\n", argv[0]);
        return 1;
    }
    const int remote = argc >= 4 ? atoi(argv[3]) : 0;
    const int batch = argc >= 5 ? atoi(argv[4]) : 8;
    size_t nchips = 0;
    uint8_t* chips = (uint8_t*)slurp(argv[2], &nchips);
    if (!chips || nchips == 0) { fprintf(stderr, "cannot read code file %s\n", argv[2]); return 1; }

    twstft_config cfg;
    twstft_default_config(&cfg);
    cfg.nsamples = (int64_t)nchips * cfg.samples_per_chip;
    cfg.max_batch = batch > 0 ? batch : 8;
    cfg.fast_peak = 1;     /* one inverse transform + certificate; uncertified records are re-run in full: same rows */
    if (remote) { cfg.win_lo_hz = 80000.0; cfg.win_hi_hz = 120000.0; }
    twstft_plan* plan = NULL;
    int rc = twstft_plan_create(&cfg, &plan);
    if (rc != TWSTFT_OK) {
        fprintf(stderr, "twstft_plan_create: %s\n", twstft_last_error(plan));
        twstft_plan_destroy(plan);
        return 1;
    }
    if (twstft_set_code(plan, chips, nchips, 0) != TWSTFT_OK) {
        fprintf(stderr, "twstft_set_code: %s\n", twstft_last_error(plan));
        twstft_plan_destroy(plan);
        return 1;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); twstft_plan_destroy(plan); return 1; }
    const size_t N = (size_t)cfg.nsamples, rec_i16 = 4 * N;
    int16_t* buf = (int16_t*)twstft_host_alloc(sizeof(int16_t) * rec_i16 * cfg.max_batch);
    twstft_result* res = (twstft_result*)malloc(sizeof(twstft_result) * 2 * cfg.max_batch);
    size_t cap = 1024, np = 0;
    double* col[2][8];                       /* indice correction SNRr SNRi df puissance puissancecode puissancenoise */
    for (int c = 0; c < 2; c++)
        for (int v = 0; v < 8; v++) col[c][v] = (double*)malloc(sizeof(double) * cap);
    const double fs = cfg.fs;
    for (;;) {
        const size_t got = fread(buf, sizeof(int16_t) * rec_i16, cfg.max_batch, f);   /* whole records only */
        if (got == 0) break;
        if (twstft_process(plan, buf, got, res) != TWSTFT_OK) {
            fprintf(stderr, "twstft_process: %s\n", twstft_last_error(plan));
            return 1;
        }
        for (size_t r = 0; r < got; r++, np++) {
            if (np == cap) {
                cap *= 2;
                for (int c = 0; c < 2; c++)
                    for (int v = 0; v < 8; v++) col[c][v] = (double*)realloc(col[c][v], sizeof(double) * cap);
            }
            for (int c = 0; c < 2; c++) {
                const twstft_result* x = &res[r * 2 + c];
                const double vals[8] = {x->indice, x->correction, x->SNRr, x->SNRi, x->df,
                                        x->puissance, x->puissancecode, x->puissancenoise};
                for (int v = 0; v < 8; v++) col[c][v][np] = vals[v];
            }
            const twstft_result *a = &res[r * 2], *b = &res[r * 2 + 1];
            printf("%zu\t%.12f\t%.3f\t%.1f\t%.1f", np + 1, (a->indice - 1 + a->correction) / fs / 3, a->df,
                   10 * log10(a->puissance), 10 * log10(a->SNRi + a->SNRr));
            if (!remote)      /* the reference prints -correction2 here (godual_ranging.m:96); values kept as stored */
                printf("\t%.12f\t%.3f\t%.1f\t%.1f", (b->indice - 1 - b->correction) / fs / 3, b->df,
                       10 * log10(b->puissance), 10 * log10(b->SNRi + b->SNRr));
            printf("\r\n");
        }
        if (got < (size_t)cfg.max_batch) break;
    }
    fclose(f);

    /* <dir>/[remote]<name>.mat */
    char out[4096];
    const char* slash = strrchr(argv[1], '/');
    const size_t dirlen = slash ? (size_t)(slash - argv[1] + 1) : 0;
    snprintf(out, sizeof out, "%.*s%s%s", (int)dirlen, argv[1], remote ? "remote" : "", argv[1] + dirlen);
    char* dot = strrchr(out, '.');
    if (dot && !strcmp(dot, ".bin")) strcpy(dot, ".mat");
    else strncat(out, ".mat", sizeof out - strlen(out) - 1);
    FILE* m = fopen(out, "wb");
    if (!m || mat5_open(m, "variant G result record") != 0) { fprintf(stderr, "cannot write %s\n", out); return 1; }
    static const char* names[2][8] = {
        {"indice1", "correction1", "SNR1r", "SNR1i", "df1", "puissance1", "puissance1code", "puissance1noise"},
        {"indice2", "correction2", "SNR2r", "SNR2i", "df2", "puissance2", "puissance2code", "puissance2noise"}};
    for (int c = 0; c < (remote ? 1 : 2); c++)
        for (int v = 0; v < 8; v++) mat5_vector(m, names[c][v], col[c][v], NULL, np);
    double* code = (double*)malloc(sizeof(double) * N);
    for (size_t i = 0; i < N; i++) code[i] = 2.0 * chips[i / cfg.samples_per_chip] - 1.0;   /* godual_ranging.m:64-65 */
    mat5_vector(m, "code", code, NULL, N);
    fclose(m);
    fprintf(stderr, "%zu records -> %s\n", np, out);
    twstft_host_free(buf);
    twstft_plan_destroy(plan);
    return 0;
}
