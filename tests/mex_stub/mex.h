/* TEST INFRASTRUCTURE: a minimal stand-in for the MEX runtime (Octave's / MATLAB's mex.h), just the calls
 * processing/C/twstft_mex.c makes, with the separate real/imaginary storage Octave uses.  It lets the test-suite
 * compile AND run the marshaller (tests/mex_stub/mex_host.c calls mexFunction the way the interpreter would);
 * this image has no Octave, MATLAB, mkoctfile or real mex.h. */
#ifndef TWSTFT_TEST_MEX_H
#define TWSTFT_TEST_MEX_H
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef struct mxArray_tag {
    size_t m, n;
    int is_complex;
    double *pr, *pi;
} mxArray;
mxArray* mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c);
double* mxGetPr(const mxArray* a);
double* mxGetPi(const mxArray* a);
int mxIsComplex(const mxArray* a);
int mxIsDouble(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
double mxGetScalar(const mxArray* a);
void* mxMalloc(size_t n);
void mxFree(void* p);
void mexErrMsgTxt(const char* msg);                                /* does not return */
const mxArray* mexGetVariablePtr(const char* workspace, const char* name);
int mexAtExit(void (*fn)(void));
const char* mexFunctionName(void);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
