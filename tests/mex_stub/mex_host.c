/* TEST INFRASTRUCTURE: plays the interpreter for processing/C/twstft_mex.c.
 *   mex_host <function name> <nlhs> <blob>
 * The blob holds the right-hand-side arguments and the global variables as records
 *   name[32] | is_global int32 | is_complex int32 | n int64 | re[n] double | im[n] double (if complex)
 * in call order.  Calls mexFunction once and prints every output as "re im" lines, one block per output. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"

#define MAXV 32
static struct { char name[32]; mxArray* a; } g_globals[MAXV];
static int g_nglobals = 0;
static const char* g_fname = "processing";
static void (*g_atexit)(void) = NULL;

mxArray* mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c)
{
    mxArray* a = (mxArray*)calloc(1, sizeof *a);
    a->m = m; a->n = n; a->is_complex = c == mxCOMPLEX;
    a->pr = (double*)calloc((m * n) > 0 ? m * n : 1, sizeof(double));
    a->pi = a->is_complex ? (double*)calloc((m * n) > 0 ? m * n : 1, sizeof(double)) : NULL;
    return a;
}
double* mxGetPr(const mxArray* a) { return a->pr; }
double* mxGetPi(const mxArray* a) { return a->pi; }
int mxIsComplex(const mxArray* a) { return a->is_complex; }
int mxIsDouble(const mxArray* a) { (void)a; return 1; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
double mxGetScalar(const mxArray* a) { return a->pr[0]; }
void* mxMalloc(size_t n) { return malloc(n ? n : 1); }
void mxFree(void* p) { free(p); }
void mexErrMsgTxt(const char* msg)
{
    fprintf(stderr, "error: %s\n", msg);
    if (g_atexit) g_atexit();
    exit(3);
}
const mxArray* mexGetVariablePtr(const char* workspace, const char* name)
{
    if (strcmp(workspace, "global")) return NULL;
    for (int i = 0; i < g_nglobals; i++)
        if (!strcmp(g_globals[i].name, name)) return g_globals[i].a;
    return NULL;
}
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }
const char* mexFunctionName(void) { return g_fname; }

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: mex_host <function> <nlhs> <blob>\n"); return 2; }
    g_fname = argv[1];
    const int nlhs = atoi(argv[2]);
    FILE* f = fopen(argv[3], "rb");
    if (!f) { perror(argv[3]); return 2; }
    mxArray* rhs[MAXV];
    int nrhs = 0;
    for (;;) {
        char name[32];
        int32_t isg, isc;
        int64_t n;
        if (fread(name, 1, 32, f) != 32) break;
        if (fread(&isg, 4, 1, f) != 1 || fread(&isc, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) return 2;
        mxArray* a = mxCreateDoubleMatrix((size_t)n, 1, isc ? mxCOMPLEX : mxREAL);
        if (fread(a->pr, 8, (size_t)n, f) != (size_t)n) return 2;
        if (isc && fread(a->pi, 8, (size_t)n, f) != (size_t)n) return 2;
        if (isg) { memcpy(g_globals[g_nglobals].name, name, 32); g_globals[g_nglobals++].a = a; }
        else rhs[nrhs++] = a;
    }
    fclose(f);
    mxArray* lhs[16] = {0};
    mexFunction(nlhs, lhs, nrhs, (const mxArray**)rhs);
    for (int o = 0; o < (nlhs > 1 ? nlhs : 1); o++) {
        if (!lhs[o]) { printf("output %d missing\n", o); continue; }
        printf("output %d %zu\n", o, mxGetNumberOfElements(lhs[o]));
        for (size_t i = 0; i < mxGetNumberOfElements(lhs[o]); i++)
            printf("%.17g %.17g\n", lhs[o]->pr[i], lhs[o]->pi ? lhs[o]->pi[i] : 0.0);
    }
    if (g_atexit) g_atexit();
    return 0;
}
