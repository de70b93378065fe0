/* MEX / oct-file entry point over libtwstft: a drop-in for the Octave function
 *   [indice,correction,SNRr,SNRi,df,puissance,puissancecode,puissancenoise] = processing(d,k)
 * of processing/Octave/godual_ranging.m:12-49 (variant G), called as
 *   [...] = twstft_processing(raw_int16, chips, [win_lo_hz win_hi_hz])
 * with raw_int16 the int16 vector read from the file for ONE record ([I1 Q1 I2 Q2]..., 4N values) and
 * chips the uint8 code file contents; outputs are 1x2 (channel 1, channel 2).
 * Compiled only where mex.h exists (mkoctfile --mex / mex); this image has neither, so the file is
 * shipped as source and exercised through the same C ABI by processing/C/godual.c and the ctypes tests. */
#if defined(HAVE_MEX_H) || defined(MATLAB_MEX_FILE)
#include <string.h>
#include "mex.h"
#include "../../include/twstft.h"

static twstft_plan* g_plan = NULL;
static size_t g_nchips = 0;

static void cleanup(void)
{
    if (g_plan) twstft_plan_destroy(g_plan);
    g_plan = NULL;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    if (nrhs < 2 || !mxIsInt16(prhs[0]) || !mxIsUint8(prhs[1]))
        mexErrMsgTxt("usage: [indice,correction,SNRr,SNRi,df,puissance,puissancecode,puissancenoise] = "
                     "twstft_processing(int16 record, uint8 chips, [lo hi])");
    const size_t nchips = mxGetNumberOfElements(prhs[1]);
    if (!g_plan || nchips != g_nchips) {
        cleanup();
        twstft_config cfg;
        twstft_default_config(&cfg);
        cfg.nsamples = (int64_t)nchips * cfg.samples_per_chip;
        cfg.max_batch = 1;
        if (nrhs >= 3 && mxGetNumberOfElements(prhs[2]) == 2) {
            cfg.win_lo_hz = mxGetPr(prhs[2])[0];
            cfg.win_hi_hz = mxGetPr(prhs[2])[1];
        }
        if (twstft_plan_create(&cfg, &g_plan) != TWSTFT_OK) {
            const char* msg = twstft_last_error(g_plan);
            mexErrMsgTxt(msg);
        }
        if (twstft_set_code(g_plan, (const uint8_t*)mxGetData(prhs[1]), nchips, 0) != TWSTFT_OK)
            mexErrMsgTxt(twstft_last_error(g_plan));
        g_nchips = nchips;
        mexAtExit(cleanup);
    }
    if (mxGetNumberOfElements(prhs[0]) != 8 * nchips) mexErrMsgTxt("record must hold 4*N int16 values, N = 2*chips");
    twstft_result r[2];
    if (twstft_process(g_plan, (const int16_t*)mxGetData(prhs[0]), 1, r) != TWSTFT_OK)
        mexErrMsgTxt(twstft_last_error(g_plan));
    for (int o = 0; o < nlhs && o < 8; o++) {
        plhs[o] = mxCreateDoubleMatrix(1, 2, mxREAL);
        double* d = mxGetPr(plhs[o]);
        for (int c = 0; c < 2; c++) {
            const double v[8] = {r[c].indice, r[c].correction, r[c].SNRr, r[c].SNRi, r[c].df,
                                 r[c].puissance, r[c].puissancecode, r[c].puissancenoise};
            d[c] = v[o];
        }
    }
}
#else
/* mex.h not available: nothing to build. */
typedef int twstft_mex_translation_unit_is_empty;
#endif
