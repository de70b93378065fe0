/* MEX entry point over libtwstft: drop-in replacements for the Octave FUNCTIONS of the reference, with their
 * exact signatures and their use of the script-level globals.  One source, installed under two names
 * (mexFunctionName() selects the personality):
 *
 *   processing.mex
 *     [indice,correction,SNRr,SNRi,df,puissance,puissancecode,puissancenoise] = processing(d,k)
 *                                                           processing/Octave/godual_ranging.m:12-49   (variant G)
 *     [xval,indice,correction,SNRr,SNRi,puissance,puissancecode,puissancenoise] = processing(d,k,df)
 *                                                           acquisition/claudio_aligned_code_lo_separate.m:24-77
 *     [xval,indice,correction,SNRr,SNRi,puissance,puissancecode,puissancenoise] = processing(d,df)
 *                                                           acquisition/claudio_aligned_code_re_separate.m:48-101
 *   search_df.mex
 *     k = search_df(d,k,df_threshold)                       acquisition/claudio_aligned_code_re_separate.m:26-46
 *
 * d: complex double column (mean already removed by the caller for processing, raw block for search_df);
 * k: vector of 1-based indices into the fftshift'ed spectrum (the contiguous range k=find(...) yields).
 * Like the Octave functions the MEX reads `global temps freq fcode code fs Nint` (godual_ranging.m:3,13;
 * lo_separate.m:7,25): fs, Nint, the +-1 replica `code` (x2 sample-and-hold of the chips, godual_ranging.m:64-65)
 * and the frequency axis `freq` (inclusive linspace = godual_ranging.m:73, or ...-fs/fs = lo_separate.m:103).
 * `temps` and `fcode` are implied by N, fs and code.  Which personality of processing() is meant follows from the
 * arguments exactly as in the scripts: (d,k) with a vector k -> G; three arguments, or a scalar second one -> K.
 *
 * Build where mex.h exists:  mkoctfile --mex -o processing.mex twstft_mex.c -ltwstft && cp processing.mex search_df.mex
 * This image has no Octave/MATLAB: the file is compiled and RUN in the test-suite against a minimal stand-in for the
 * MEX runtime (tests/mex_stub/), which calls mexFunction with the arguments and globals the scripts would pass. */
#if defined(HAVE_MEX_H) || defined(MATLAB_MEX_FILE)
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "mex.h"
#include "../../include/twstft.h"

typedef struct {
    twstft_plan* plan;
    int64_t n;
    double fs;
    int variant, axis;
    uint64_t code_hash;
    int have_code;
} slot;
static slot g_slot[3];          /* 0: processing G, 1: processing K / search_df correlator, 2: search_df estimator */
static int g_exit_registered = 0;

static void cleanup(void)
{
    for (int i = 0; i < 3; i++) {
        if (g_slot[i].plan) twstft_plan_destroy(g_slot[i].plan);
        memset(&g_slot[i], 0, sizeof g_slot[i]);
    }
}

static const mxArray* need_global(const char* name)
{
    const mxArray* a = mexGetVariablePtr("global", name);
    if (!a) {
        char msg[128];
        snprintf(msg, sizeof msg, "global variable `%s` is not set (the scripts define temps freq fcode code fs Nint)", name);
        mexErrMsgTxt(msg);
    }
    return a;
}

/* complex double argument -> interleaved (re,im) buffer; *owned says whether the caller must mxFree it */
static const double* complex_column(const mxArray* a, size_t* n, int* owned)
{
    if (!mxIsDouble(a)) mexErrMsgTxt("d must be a double (complex) vector");
    *n = mxGetNumberOfElements(a);
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
    if (mxIsComplex(a)) { *owned = 0; return (const double*)mxGetComplexDoubles(a); }
#endif
    double* buf = (double*)mxMalloc(sizeof(double) * 2 * (*n ? *n : 1));
    const double* re = mxGetPr(a);
    const double* im = mxIsComplex(a) ? mxGetPi(a) : NULL;
    for (size_t i = 0; i < *n; i++) { buf[2 * i] = re[i]; buf[2 * i + 1] = im ? im[i] : 0.0; }
    *owned = 1;
    return buf;
}

static uint64_t hash_code(const double* c, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= (uint64_t)(c[i] > 0.0) + 2 * (i & 1); h *= 1099511628211ull; }
    return h ^ n;
}

/* plan of slot `si` for (n, fs, variant, axis); the code comes from the global `code` unless with_code == 0 */
static twstft_plan* plan_for(int si, int64_t n, double fs, int variant, int axis, int with_code)
{
    slot* s = &g_slot[si];
    if (!g_exit_registered) { mexAtExit(cleanup); g_exit_registered = 1; }
    if (mxGetScalar(need_global("Nint")) != 1.0) mexErrMsgTxt("only Nint = 1 is supported (the value every script uses)");
    if (!s->plan || s->n != n || s->fs != fs || s->variant != variant || s->axis != axis) {
        if (s->plan) twstft_plan_destroy(s->plan);
        memset(s, 0, sizeof *s);
        twstft_config cfg;
        twstft_default_config(&cfg);
        cfg.fs = fs; cfg.nsamples = n; cfg.variant = variant; cfg.freq_axis = axis;
        cfg.layout = TWSTFT_LAYOUT_SINGLE; cfg.max_batch = with_code ? 8 : 1; cfg.est_sub_mean = 0; cfg.ddof = 1;
        if (twstft_plan_create(&cfg, &s->plan) != TWSTFT_OK) {
            static char msg[512];
            snprintf(msg, sizeof msg, "twstft_plan_create: %s", twstft_last_error(s->plan));
            twstft_plan_destroy(s->plan);
            s->plan = NULL;
            mexErrMsgTxt(msg);
        }
        s->n = n; s->fs = fs; s->variant = variant; s->axis = axis;
    }
    if (with_code) {
        const mxArray* code = need_global("code");
        const size_t nc = mxGetNumberOfElements(code);
        if (mxIsComplex(code) || !mxIsDouble(code)) mexErrMsgTxt("global `code` must be the real +-1 replica (BPSK)");
        if ((int64_t)nc != n || (nc & 1)) mexErrMsgTxt("length(d) must equal length(code) (one code period per call)");
        const double* c = mxGetPr(code);
        const uint64_t h = hash_code(c, nc);
        if (!s->have_code || h != s->code_hash) {
            uint8_t* chips = (uint8_t*)mxMalloc(nc / 2);
            for (size_t i = 0; i < nc / 2; i++) {
                if (c[2 * i] != c[2 * i + 1]) mexErrMsgTxt("global `code` is not a x2 sample-and-hold replica");
                chips[i] = c[2 * i] > 0.0;
            }
            const int rc = twstft_set_code(s->plan, chips, nc / 2, 0);
            mxFree(chips);
            if (rc != TWSTFT_OK) mexErrMsgTxt(twstft_last_error(s->plan));
            s->code_hash = h;
            s->have_code = 1;
        }
    }
    return s->plan;
}

/* inclusive linspace(-fs/2,fs/2,n) (godual_ranging.m:73) or linspace(-fs/2,fs/2-1,n) (lo_separate.m:103) */
static int axis_of_freq(double fs)
{
    const mxArray* f = need_global("freq");
    const size_t n = mxGetNumberOfElements(f);
    if (n < 2) mexErrMsgTxt("global `freq` is empty");
    return mxGetPr(f)[n - 1] == fs / 2 ? TWSTFT_AXIS_G : TWSTFT_AXIS_K;
}

static void index_range(const mxArray* k, int64_t* first, int64_t* last)
{
    const size_t nk = mxGetNumberOfElements(k);
    if (nk == 0) mexErrMsgTxt("k is empty");
    const double* v = mxGetPr(k);
    *first = (int64_t)v[0];
    *last = (int64_t)v[nk - 1];
    if (*last - *first + 1 != (int64_t)nk) mexErrMsgTxt("k must be a contiguous index range (k=find(...) of a band)");
}

static mxArray* scalar(double v)
{
    mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
    mxGetPr(a)[0] = v;
    return a;
}

static void do_processing(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    if (nrhs < 2 || nrhs > 3) mexErrMsgTxt("usage: processing(d,k) | processing(d,k,df) | processing(d,df)");
    const double fs = mxGetScalar(need_global("fs"));
    size_t n;
    int owned;
    const double* d = complex_column(prhs[0], &n, &owned);
    const int is_k = nrhs == 3 || mxGetNumberOfElements(prhs[1]) == 1;
    twstft_result r;
    int rc;
    if (!is_k) {                                            /* processing(d,k): variant G */
        twstft_plan* p = plan_for(0, (int64_t)n, fs, TWSTFT_VARIANT_G, axis_of_freq(fs), 1);
        int64_t k1, k2;
        index_range(prhs[1], &k1, &k2);
        rc = twstft_processing_cf64(p, d, 1, k1, k2, &r);
        if (rc != TWSTFT_OK) mexErrMsgTxt(twstft_last_error(p));
        const double v[8] = {r.indice, r.correction, r.SNRr, r.SNRi, r.df, r.puissance, r.puissancecode, r.puissancenoise};
        for (int o = 0; o < (nlhs > 1 ? nlhs : 1) && o < 8; o++) plhs[o] = scalar(v[o]);
    } else {                                                /* processing(d,k,df) / processing(d,df): variant K */
        twstft_plan* p = plan_for(1, (int64_t)n, fs, TWSTFT_VARIANT_K, TWSTFT_AXIS_K, 1);
        const double df = mxGetScalar(prhs[nrhs - 1]);
        rc = twstft_processing_cf64_df(p, d, 1, &df, &r);
        if (rc != TWSTFT_OK) mexErrMsgTxt(twstft_last_error(p));
        plhs[0] = mxCreateDoubleMatrix(1, 1, mxCOMPLEX);
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
        mxGetComplexDoubles(plhs[0])[0].real = r.xval_re;
        mxGetComplexDoubles(plhs[0])[0].imag = r.xval_im;
#else
        mxGetPr(plhs[0])[0] = r.xval_re;
        mxGetPi(plhs[0])[0] = r.xval_im;
#endif
        const double v[7] = {r.indice, r.correction, r.SNRr, r.SNRi, r.puissance, r.puissancecode, r.puissancenoise};
        for (int o = 1; o < nlhs && o < 8; o++) plhs[o] = scalar(v[o - 1]);
    }
    if (owned) mxFree((void*)d);
}

static void do_search_df(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    (void)nlhs;
    if (nrhs != 3) mexErrMsgTxt("usage: k = search_df(d,k,df_threshold)");
    const double fs = mxGetScalar(need_global("fs"));
    size_t n;
    int owned;
    const double* d = complex_column(prhs[0], &n, &owned);
    const size_t nfreq = mxGetNumberOfElements(need_global("freq"));
    if (nfreq != n) mexErrMsgTxt("length(d) must equal length(freq) (the block the spectrum is taken over)");
    const size_t ncode = mxGetNumberOfElements(need_global("code"));
    /* the block is ls windows of the estimator plan: ls = 2 for the 2-s blocks (N = fs), 1 otherwise */
    const int ls = (n % 2 == 0 && (double)(n / 2) == fs) ? 2 : 1;
    twstft_plan* cor = plan_for(1, (int64_t)ncode, fs, TWSTFT_VARIANT_K, TWSTFT_AXIS_K, 1);
    twstft_plan* est = plan_for(2, (int64_t)(n / ls), fs, TWSTFT_VARIANT_K, axis_of_freq(fs), 0);   /* candidates: freq(k)/2 */
    int64_t k1, k2, kbon = 0;
    index_range(prhs[1], &k1, &k2);
    const int rc = twstft_search_df_cf64(est, cor, d, ls, k1, k2, mxGetScalar(prhs[2]), &kbon, NULL, NULL);
    if (owned) mxFree((void*)d);
    if (rc != TWSTFT_OK) mexErrMsgTxt(twstft_last_error(est));
    plhs[0] = scalar((double)kbon);
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    const char* name = mexFunctionName();
    if (name && strstr(name, "search_df")) do_search_df(nlhs, plhs, nrhs, prhs);
    else do_processing(nlhs, plhs, nrhs, prhs);
}
#else
/* mex.h not available: nothing to build. */
typedef int twstft_mex_translation_unit_is_empty;
#endif
