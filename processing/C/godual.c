/* godual -- plain-C front end over libtwstft (include/twstft.h): the record loop of
 * processing/Octave/godual_ranging.m:57-130 (variant G) on a B210 two-channel recording.
 *
 *   godual data.bin code.bin [remote=0] [batch=8] [gpus=1]
 *
 * data.bin : int16 LE [I1 Q1 I2 Q2] per sample instant (processing/Python/b210_file.py:60-65)
 * code.bin : one byte (0/1) per chip (amaranth_twstft/common.py:59-73); N = 2 * chips
 * remote   : 0 = local file, both channels, carrier window |f| < 20 kHz          (godual_ranging.m:82-83)
 *            1 = remote file, channel 1 only is reported, window 80..120 kHz     (:84-90, OP != 1)
 * Output   : one stdout row per record in the format of godual_ranging.m:96-98 and a MAT-v5 file
 *            <dir>/[remote]<name>.mat with the variables of :128-130 (corr*, df*, indic*, SNR*, code, puissan*).
 * The command-line shape (positional data/code/remote) is that of processing/CPP/main.cpp:779-798.
 * A trailing partial record is ignored (:81,102).  Exit status 0 on success, 1 on any error.
 * gpus     : the records of every chunk are split by index over that many GPUs (twstft_shard_*: one host thread and
 *            one plan per GPU, no collective); rows come back in file order.
 * Ingest   : a reader thread fills the second pinned chunk while the GPUs work on the first, as the C++ program
 *            prefetches the next record while its workers run (processing/CPP/main.cpp:488-497).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <time.h>
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

/* two-slot hand-over between the reader thread and the processing loop */
typedef struct {
    FILE* f;
    int16_t* buf[2];
    size_t got[2];
    int full[2];
    size_t rec_bytes, chunk;
    pthread_mutex_t mu;
    pthread_cond_t cv;
} ingest;

static void* reader_main(void* arg)
{
    ingest* in = (ingest*)arg;
    for (int s = 0;; s ^= 1) {
        pthread_mutex_lock(&in->mu);
        while (in->full[s]) pthread_cond_wait(&in->cv, &in->mu);
        pthread_mutex_unlock(&in->mu);
        const size_t got = fread(in->buf[s], in->rec_bytes, in->chunk, in->f);   /* whole records only */
        pthread_mutex_lock(&in->mu);
        in->got[s] = got;
        in->full[s] = 1;
        pthread_cond_broadcast(&in->cv);
        pthread_mutex_unlock(&in->mu);
        if (got < in->chunk) return NULL;                                        /* short read ends the loop (:81,102) */
    }
}

static double now_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "%s data.bin code.bin [remote=0] [batch=8] [gpus=1]\n", argv[0]);
        return 1;
    }
    const int remote = argc >= 4 ? atoi(argv[3]) : 0;
    const int batch = argc >= 5 ? atoi(argv[4]) : 8;
    const int gpus = argc >= 6 && atoi(argv[5]) > 0 ? atoi(argv[5]) : 1;
    size_t nchips = 0;
    uint8_t* chips = (uint8_t*)slurp(argv[2], &nchips);
    if (!chips || nchips == 0) { fprintf(stderr, "cannot read code file %s\n", argv[2]); return 1; }

    twstft_config cfg;
    twstft_default_config(&cfg);
    cfg.nsamples = (int64_t)nchips * cfg.samples_per_chip;
    cfg.max_batch = batch > 0 ? batch : 8;
    cfg.fast_peak = 1;     /* one inverse transform + certificate; uncertified records are re-run in full: same rows */
    if (remote) { cfg.win_lo_hz = 80000.0; cfg.win_hi_hz = 120000.0; }
    twstft_shard* sh = NULL;
    int rc = twstft_shard_create(&cfg, NULL, gpus, &sh);
    if (rc != TWSTFT_OK) {
        fprintf(stderr, "twstft_shard_create: %s\n", twstft_shard_last_error(sh));
        twstft_shard_destroy(sh);
        return 1;
    }
    if (twstft_shard_set_code(sh, chips, nchips, 0) != TWSTFT_OK) {
        fprintf(stderr, "twstft_set_code: %s\n", twstft_shard_last_error(sh));
        twstft_shard_destroy(sh);
        return 1;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); twstft_shard_destroy(sh); return 1; }
    const size_t N = (size_t)cfg.nsamples, rec_i16 = 4 * N;
    const size_t chunk = (size_t)cfg.max_batch * (size_t)gpus;             /* records per hand-over */
    ingest in;
    memset(&in, 0, sizeof in);
    in.f = f; in.rec_bytes = sizeof(int16_t) * rec_i16; in.chunk = chunk;
    pthread_mutex_init(&in.mu, NULL);
    pthread_cond_init(&in.cv, NULL);
    for (int s2 = 0; s2 < 2; s2++) {
        in.buf[s2] = (int16_t*)twstft_host_alloc(in.rec_bytes * chunk);    /* pinned: the H2D copies read it directly */
        if (!in.buf[s2]) { fprintf(stderr, "pinned host allocation failed\n"); return 1; }
    }
    twstft_result* res = (twstft_result*)malloc(sizeof(twstft_result) * 2 * chunk);
    pthread_t reader;
    if (pthread_create(&reader, NULL, reader_main, &in) != 0) { fprintf(stderr, "cannot start the reader thread\n"); return 1; }
    const double t_start = now_s();
    size_t cap = 1024, np = 0;
    double* col[2][8];                       /* indice correction SNRr SNRi df puissance puissancecode puissancenoise */
    for (int c = 0; c < 2; c++)
        for (int v = 0; v < 8; v++) col[c][v] = (double*)malloc(sizeof(double) * cap);
    const double fs = cfg.fs;
    for (int slot = 0;; slot ^= 1) {
        pthread_mutex_lock(&in.mu);
        while (!in.full[slot]) pthread_cond_wait(&in.cv, &in.mu);
        pthread_mutex_unlock(&in.mu);
        const size_t got = in.got[slot];
        if (got == 0) break;
        if (twstft_shard_process(sh, in.buf[slot], got, res) != TWSTFT_OK) {
            fprintf(stderr, "twstft_process: %s\n", twstft_shard_last_error(sh));
            return 1;
        }
        pthread_mutex_lock(&in.mu);           /* rows are in `res`: the reader may refill this slot */
        in.full[slot] = 0;
        pthread_cond_broadcast(&in.cv);
        pthread_mutex_unlock(&in.mu);
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
        if (got < chunk) break;
    }
    pthread_join(reader, NULL);
    const double t_run = now_s() - t_start;
    fclose(f);
    fprintf(stderr, "file -> rows: %zu records in %.3f s = %.1f records/s (%.2f GB/s of raw IQ), %d GPU(s)\n", np, t_run,
            t_run > 0 ? (double)np / t_run : 0.0, t_run > 0 ? (double)np * (double)in.rec_bytes / t_run / 1e9 : 0.0, gpus);

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
    twstft_host_free(in.buf[0]);
    twstft_host_free(in.buf[1]);
    twstft_shard_destroy(sh);
    return 0;
}
