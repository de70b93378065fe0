/* Kernel-variant bench over the C ABI: loads a given build of libtwstft.so (dlopen), synthesises a few
 * two-channel 1-s records on the host (PRN from the library's own LFSR kernel, delay, carrier, AWGN),
 * runs twstft_process_device with per-kernel CUDA-event profiling and prints one JSON line with the
 * per-kernel average launch durations and a digest of the result rows -- so that several builds can be
 * compared (time AND results) in one GPU call.
 *
 *   kbench <lib.so> [records_per_step=16] [steps=6] [fast_peak=1] [distinct=4] [snr_db=-5]
 *
 * Development tool only (not part of the product path, not a benchmark of record: bench.py is).
 */
#define _POSIX_C_SOURCE 200809L
#include <dlfcn.h>
#include <time.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/twstft.h"

#define SYM(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(h, #name); \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

static uint64_t rng_s = 88172645463325252ull;
static inline uint64_t rng(void) { rng_s ^= rng_s << 13; rng_s ^= rng_s >> 7; rng_s ^= rng_s << 17; return rng_s; }
static inline double uni(void) { return ((rng() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

static int16_t sat16(double v) { v = rint(v); if (v > 32767) v = 32767; if (v < -32768) v = -32768; return (int16_t)v; }

/* dual-plan experiment: two plans on the same GPU driven by two host threads, each over half the records */
#include <pthread.h>
typedef struct { int (*proc)(twstft_plan*, const int16_t*, size_t, twstft_result*); twstft_plan* plan; const int16_t* dev; size_t n; twstft_result* res; int steps; } job;
static void* job_main(void* a)
{
    job* j = (job*)a;
    for (int s = 0; s < j->steps; s++) j->proc(j->plan, j->dev, j->n, j->res);
    return NULL;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: kbench lib.so [R] [steps] [fast] [distinct] [snr_db]\n"); return 2; }
    const int R = argc > 2 ? atoi(argv[2]) : 16, steps = argc > 3 ? atoi(argv[3]) : 6;
    const int fast = argc > 4 ? atoi(argv[4]) : 1, distinct = argc > 5 ? atoi(argv[5]) : 4;
    const double snr_db = argc > 6 ? atof(argv[6]) : -5.0;
    const int dual = argc > 7 ? atoi(argv[7]) : 0;
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    SYM(twstft_default_config) SYM(twstft_plan_create) SYM(twstft_plan_destroy) SYM(twstft_last_error)
    SYM(twstft_gen_code) SYM(twstft_process_device) SYM(twstft_device_alloc) SYM(twstft_copy_to_device)
    SYM(twstft_synchronize) SYM(twstft_timer_start) SYM(twstft_timer_stop) SYM(twstft_profile_enable)
    SYM(twstft_profile_read) SYM(twstft_fast_counters) SYM(twstft_host_alloc) SYM(twstft_kernel_launches)

    const int64_t N = 5000000;
    const double fs = 5e6;
    twstft_config cfg;
    p_twstft_default_config(&cfg);
    cfg.max_batch = R;
    cfg.fast_peak = fast;
    twstft_plan* plan = NULL;
    int rc = p_twstft_plan_create(&cfg, &plan);
    if (rc) { fprintf(stderr, "plan_create %d: %s\n", rc, p_twstft_last_error(plan)); return 1; }
    uint8_t* chips = (uint8_t*)malloc(N / 2);
    rc = p_twstft_gen_code(plan, 22, 3, 0, 1, N / 2, chips);
    if (rc) { fprintf(stderr, "gen_code %d: %s\n", rc, p_twstft_last_error(plan)); return 1; }

    /* records: ch1 SNR snr_db, random delay/df/phase; ch2 loopback +0.6 dB, df 0 */
    const size_t rec_i16 = (size_t)N * 4;
    int16_t* host = (int16_t*)p_twstft_host_alloc(rec_i16 * sizeof(int16_t) * distinct);
    for (int r = 0; r < distinct; r++) {
        int16_t* o = host + (size_t)r * rec_i16;
        for (int ch = 0; ch < 2; ch++) {
            const double snr = pow(10.0, (ch == 0 ? snr_db : 0.6) / 10.0);
            const double sigma = 3000.0 / sqrt(1.0 + snr), amp = sigma * sqrt(2.0 * snr);   /* per-component noise std, signal amplitude */
            const int64_t delay = (int64_t)(uni() * N);
            const double df = ch == 0 ? (uni() * 2.0 - 1.0) * 9000.0 : 0.0, ph = uni() * 6.283185307179586;
            const double dcr = 7.0, dci = -3.0;
            for (int64_t n = 0; n < N; n += 1) {
                int64_t m = n - delay; if (m < 0) m += N;
                const double c = chips[m >> 1] ? 1.0 : -1.0;
                const double a = ph + 6.283185307179586 * fmod(df * (double)n / fs, 1.0);
                const double u1 = uni(), u2 = uni();
                const double rr = sigma * sqrt(-2.0 * log(u1));
                o[4 * n + 2 * ch] = sat16(amp * c * cos(a) + rr * cos(6.283185307179586 * u2) + dcr);
                o[4 * n + 2 * ch + 1] = sat16(amp * c * sin(a) + rr * sin(6.283185307179586 * u2) + dci);
            }
        }
    }
    void* dev = NULL;
    rc = p_twstft_device_alloc(plan, rec_i16 * sizeof(int16_t) * R, &dev);
    if (rc) { fprintf(stderr, "device_alloc %d\n", rc); return 1; }
    for (int r = 0; r < R; r++)
        p_twstft_copy_to_device(plan, (char*)dev + (size_t)r * rec_i16 * 2, host + (size_t)(r % distinct) * rec_i16, rec_i16 * 2);
    twstft_result* res = (twstft_result*)calloc((size_t)R * 2, sizeof(twstft_result));
    for (int w = 0; w < 3; w++) {
        rc = p_twstft_process_device(plan, (const int16_t*)dev, R, res);
        if (rc) { fprintf(stderr, "process %d: %s\n", rc, p_twstft_last_error(plan)); return 1; }
    }
    p_twstft_synchronize(plan);
    const long long l0 = p_twstft_kernel_launches(plan);
    p_twstft_timer_start(plan);
    for (int s = 0; s < steps; s++) p_twstft_process_device(plan, (const int16_t*)dev, R, res);
    float ms = 0;
    p_twstft_timer_stop(plan, &ms);
    const long long launches = p_twstft_kernel_launches(plan) - l0;
    if (dual) {     /* second plan, same device: the two halves of the batch run concurrently on two streams */
        twstft_plan* plan2 = NULL;
        cfg.max_batch = R / 2;
        if (p_twstft_plan_create(&cfg, &plan2) || p_twstft_gen_code(plan2, 22, 3, 0, 1, N / 2, chips)) { fprintf(stderr, "plan2 failed\n"); return 1; }
        twstft_result* res2 = (twstft_result*)calloc((size_t)R, sizeof(twstft_result));
        job jb[2] = {{p_twstft_process_device, plan, (const int16_t*)dev, (size_t)R / 2, res, 3},
                     {p_twstft_process_device, plan2, (const int16_t*)((char*)dev + (size_t)(R / 2) * rec_i16 * 2), (size_t)R / 2, res2, 3}};
        pthread_t th[2];
        for (int t = 0; t < 2; t++) pthread_create(&th[t], NULL, job_main, &jb[t]);
        for (int t = 0; t < 2; t++) pthread_join(th[t], NULL);
        jb[0].steps = jb[1].steps = steps;
        struct timespec a, b;
        clock_gettime(CLOCK_MONOTONIC, &a);
        for (int t = 0; t < 2; t++) pthread_create(&th[t], NULL, job_main, &jb[t]);
        for (int t = 0; t < 2; t++) pthread_join(th[t], NULL);
        clock_gettime(CLOCK_MONOTONIC, &b);
        const double dt = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
        printf("{\"dual_plans\": true, \"R\": %d, \"ms_per_step_wall\": %.4f, \"records_per_s\": %.1f}\n", R, 1e3 * dt / steps, R * steps / dt);
        p_twstft_plan_destroy(plan2);
    }
    p_twstft_profile_enable(plan, 1);
    for (int s = 0; s < 2; s++) p_twstft_process_device(plan, (const int16_t*)dev, R, res);
    p_twstft_profile_enable(plan, 0);
    int64_t frec = 0, frerun = 0;
    p_twstft_fast_counters(plan, &frec, &frerun);
    printf("{\"lib\": \"%s\", \"R\": %d, \"steps\": %d, \"ms_per_step\": %.4f, \"records_per_s\": %.1f, \"us_per_lane\": %.3f, "
           "\"launches_per_step\": %.1f, \"fast\": [%lld, %lld], \"kernels\": {",
           argv[1], R, steps, ms / steps, R * steps / (ms * 1e-3), 1e3 * ms / steps / (2.0 * R), (double)launches / steps,
           (long long)frec, (long long)frerun);
    int first = 1;
    for (int slot = 0;; slot++) {
        char name[64]; double tot; int64_t cnt;
        if (p_twstft_profile_read(plan, slot, name, sizeof name, &tot, &cnt) != 0) break;
        if (cnt == 0) continue;
        printf("%s\"%s\": %.4f", first ? "" : ", ", name, tot / (double)cnt);
        first = 0;
    }
    printf("}, \"rows\": [");
    for (int i = 0; i < 2 * distinct && i < 2 * R; i++)
        printf("%s[%.0f, %.9f, %.6f, %.4f, %.0f, %.0f]", i ? ", " : "", res[i].indice, res[i].correction,
               atan2(res[i].xval_im, res[i].xval_re), 10.0 * log10(res[i].SNRr + res[i].SNRi), res[i].carrier_bin, res[i].status);
    printf("]}\n");
    p_twstft_plan_destroy(plan);
    return 0;
}
