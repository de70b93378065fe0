/* go_1s -- the per-second reducer of acquisition/go_1s.m for ONE session, in plain C (host arithmetic only;
 * nothing here touches the GPU):
 *
 *   go_1s OPlocal.mat OPremote.mat LTFBlocal.mat LTFBremote.mat MJD [out.1s]
 *
 * The four inputs are the K-variant result files written by claudio_aligned (or by the Octave scripts:
 * `*claudio*.mat`, plain or whole-file gzip'ed).  Statement by statement after go_1s.m:
 *   :80-87   k=find(abs(xval1)>max(abs(xval1))/2); cut at the first gap; drop the first 10 windows and the last
 *   :88-90   delay = (indice1+correction1/(2N+1))/fs*1e9  [ns]
 *   :93-100  cut the OP loopback at the first jump > 2 ns ("sample loss")
 *   :101     sessions with <= 102 valid windows are skipped
 *   :106-123 OP remote on the same windows, truncated at the first amplitude gap; median C/N0
 *   :139-172 the same for the LTFB pair (:165-169 "truncated2")
 *   :175-181 equal lengths
 *   :191-193 res = 0.5*((opre-oplo)-(ltre-ltlo)), outliers (> 5 ns from the median) -> NaN
 *   :208-211 the +-200/(2N+1) ns unwrap, exactly as written (NaN-skipping median: the script loads pkg nan)
 *   :251-266 the `.1s` file: header, then for k=1:25:length-25 the 25-point linear fit of each of the four series
 *            evaluated at its 13th point, "%f\t%f\t%f\t%f\t%f\n" with MJD advancing by 1/86400
 * stdout: the four lengths (:173), then mean and std of res (:226,231) and the two median C/N0 values.
 * The MJD comes from the caller (the script derives it from the LTFB file name, :131-134).
 * Exit status 0, 1 on errors, 3 when the session is too short (:101).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "matv5_read.h"

typedef struct { const mat5_var *xval1, *indice1, *correction1, *snrr, *snri; mat5_file f; } kfile;

static int load_k(const char* path, kfile* k)
{
    if (mat5_load(path, &k->f) != 0) { fprintf(stderr, "cannot read %s\n", path); return -1; }
    k->xval1 = mat5_find(&k->f, "xval1");
    k->indice1 = mat5_find(&k->f, "indice1");
    k->correction1 = mat5_find(&k->f, "correction1");
    k->snrr = mat5_find(&k->f, "SNR1r");
    k->snri = mat5_find(&k->f, "SNR1i");
    if (!k->xval1 || !k->indice1 || !k->correction1 || !k->snrr || !k->snri || !k->xval1->im ||
        k->indice1->n != k->xval1->n || k->correction1->n != k->xval1->n) {
        fprintf(stderr, "%s: not a K-variant result file (xval1/indice1/correction1/SNR1r/SNR1i)\n", path);
        return -1;
    }
    return 0;
}
static double xabs(const kfile* k, size_t i) { return hypot(k->xval1->re[i], k->xval1->im[i]); }

/* :80-87 -> indices (0-based) into the file; returns the count */
static size_t select_valid(const kfile* k, size_t** out)
{
    const size_t n = k->xval1->n;
    double mx = 0;
    for (size_t i = 0; i < n; i++) if (xabs(k, i) > mx) mx = xabs(k, i);
    size_t* idx = (size_t*)malloc(sizeof(size_t) * (n ? n : 1));
    size_t m = 0;
    for (size_t i = 0; i < n; i++) if (xabs(k, i) > mx / 2) idx[m++] = i;
    size_t gap = 0;                                        /* kk(1), 1-based position in k of the last before the gap */
    for (size_t i = 0; i + 1 < m; i++) if (idx[i + 1] - idx[i] > 1) { gap = i + 1; break; }
    size_t last = gap ? gap : (m ? m - 1 : 0);             /* k(11:kk(1)) or k(11:end-1): 1-based inclusive end */
    size_t cnt = last > 10 ? last - 10 : 0;
    memmove(idx, idx + 10 < idx + m ? idx + 10 : idx, sizeof(size_t) * cnt);
    *out = idx;
    return cnt;
}
static void to_ns(const kfile* k, const size_t* idx, size_t n, double fs, double* out)
{
    for (size_t i = 0; i < n; i++) out[i] = (k->indice1->re[idx[i]] + k->correction1->re[idx[i]] / 3.0) / fs * 1e9;
}
/* xval1=xval1(k); kkk=find(abs(xval1)>max/2); kkkk=find(diff(kkk)>1): returns kkkk(1) (0 = none), *nkkk, *kkk_end */
static size_t amplitude_gap(const kfile* k, const size_t* idx, size_t n, size_t* nkkk, size_t* kkk_end)
{
    double mx = 0;
    for (size_t i = 0; i < n; i++) if (xabs(k, idx[i]) > mx) mx = xabs(k, idx[i]);
    size_t prev = 0, cnt = 0, gap = 0, end = 0;
    for (size_t i = 0; i < n; i++)
        if (xabs(k, idx[i]) > mx / 2) {
            if (cnt && !gap && i - prev > 1) gap = cnt;    /* position (1-based) in kkk of the last before the gap */
            prev = i;
            cnt++;
            end = i + 1;                                   /* kkk(end), 1-based */
        }
    *nkkk = cnt;
    *kkk_end = end;
    return gap;
}
static int cmpd(const void* a, const void* b) { const double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }
static double nanmedian(const double* v, size_t n)
{
    double* t = (double*)malloc(sizeof(double) * (n ? n : 1));
    size_t m = 0;
    for (size_t i = 0; i < n; i++) if (!isnan(v[i])) t[m++] = v[i];
    double r = NAN;
    if (m) { qsort(t, m, sizeof(double), cmpd); r = m % 2 ? t[m / 2] : 0.5 * (t[m / 2 - 1] + t[m / 2]); }
    free(t);
    return r;
}
static double median_cn0(const kfile* k, const size_t* idx, size_t n, double fs)
{
    double* t = (double*)malloc(sizeof(double) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) t[i] = 10 * log10(fabs(k->snrr->re[idx[i]] + k->snri->re[idx[i]]) * fs);   /* :123 */
    const double r = nanmedian(t, n);
    free(t);
    return r;
}
/* [a,b]=polyfit(x,y,1); b.yf(13) over 25 points x=(k-1:k+23)/25 */
static double fit13(const double* y, size_t k0)
{
    double sx = 0, sy = 0, sxx = 0, sxy = 0;
    for (int i = 0; i < 25; i++) {
        const double x = (double)(k0 + i) / 25.0;
        sx += x; sy += y[k0 + i]; sxx += x * x; sxy += x * y[k0 + i];
    }
    const double mx = sx / 25, my = sy / 25;
    double cxx = 0, cxy = 0;
    for (int i = 0; i < 25; i++) {
        const double x = (double)(k0 + i) / 25.0 - mx;
        cxx += x * x; cxy += x * (y[k0 + i] - my);
    }
    (void)sxx; (void)sxy;
    const double slope = cxy / cxx;
    return my + slope * ((double)(k0 + 12) / 25.0 - mx);
}

int main(int argc, char** argv)
{
    if (argc < 6) {
        fprintf(stderr, "%s OPlocal.mat OPremote.mat LTFBlocal.mat LTFBremote.mat MJD [out.1s] [fs=5e6]\n", argv[0]);
        return 1;
    }
    const double mjd = atof(argv[5]);
    const double fs = argc > 7 ? atof(argv[7]) : 5e6;
    kfile oplo_f, opre_f, ltlo_f, ltre_f;
    if (load_k(argv[1], &oplo_f) || load_k(argv[2], &opre_f) || load_k(argv[3], &ltlo_f) || load_k(argv[4], &ltre_f)) return 1;

    size_t* k;
    size_t nk = select_valid(&oplo_f, &k);
    double* oplo = (double*)malloc(sizeof(double) * (nk + 1));
    to_ns(&oplo_f, k, nk, fs, oplo);
    const size_t loplo = nk;
    size_t n_oplo = nk;
    for (size_t i = 0; i + 1 < n_oplo; i++)                /* :93-100 */
        if (fabs(oplo[i + 1] - oplo[i]) > 2) {
            printf("Sample loss at %zu\n", i + 1);
            if (i + 1 > 1) n_oplo = i;                     /* oplo(1:kk-1) */
            break;
        }
    if (n_oplo <= 102) { fprintf(stderr, "session too short: %zu valid loopback windows\n", n_oplo); return 3; }
    for (size_t i = 0; i < nk; i++)
        if (k[i] >= opre_f.xval1->n) { fprintf(stderr, "OP remote file shorter than the local one\n"); return 1; }
    size_t nkkk, kend;
    size_t gap = amplitude_gap(&opre_f, k, nk, &nkkk, &kend);      /* :106-117 */
    if (gap) {
        nk = gap;
        printf(" *** truncated *** ");
        if (gap < n_oplo) n_oplo = gap;
    }
    double* opre = (double*)malloc(sizeof(double) * (nk + 1));
    to_ns(&opre_f, k, nk, fs, opre);
    const size_t lopre = nk;
    size_t n_opre = nk < n_oplo ? nk : n_oplo;             /* opre(1:length(oplo)) */
    const double snrop = median_cn0(&opre_f, k, nk, fs);
    free(k);

    nk = select_valid(&ltlo_f, &k);                        /* :139-150 */
    double* ltlo = (double*)malloc(sizeof(double) * (nk + 1));
    to_ns(&ltlo_f, k, nk, fs, ltlo);
    const size_t lltlo = nk;
    size_t n_ltlo = nk;
    for (size_t i = 0; i < nk; i++)
        if (k[i] >= ltre_f.xval1->n) { fprintf(stderr, "LTFB remote file shorter than the local one\n"); return 1; }
    gap = amplitude_gap(&ltre_f, k, nk, &nkkk, &kend);     /* :157-169 */
    if (gap) { nk = gap; n_ltlo = gap; printf(" *** truncated *** "); }
    if (nkkk < n_ltlo) { nk = kend < nk ? kend : nk; n_ltlo = kend < n_ltlo ? kend : n_ltlo; printf(" *** truncated2 *** "); }
    double* ltre = (double*)malloc(sizeof(double) * (nk + 1));
    to_ns(&ltre_f, k, nk, fs, ltre);
    const size_t lltre = nk;
    size_t n_ltre = nk;
    const double snrlt = median_cn0(&ltre_f, k, nk, fs);
    printf("%zu %zu %zu %zu\n", loplo, lopre, lltlo, lltre);       /* :173 */
    if (n_oplo > n_ltlo) { n_oplo = n_ltlo; if (n_opre > n_ltlo) n_opre = n_ltlo; }   /* :175-181 */
    else { n_ltlo = n_oplo; if (n_ltre > n_oplo) n_ltre = n_oplo; }
    size_t n = n_oplo;
    if (n_opre < n) n = n_opre;
    if (n_ltlo < n) n = n_ltlo;
    if (n_ltre < n) n = n_ltre;

    double* res = (double*)malloc(sizeof(double) * (n + 1));
    for (size_t i = 0; i < n; i++) res[i] = 0.5 * ((opre[i] - oplo[i]) - (ltre[i] - ltlo[i]));     /* :191 */
    double med = nanmedian(res, n);
    for (size_t i = 0; i < n; i++) if (fabs(res[i] - med) > 5) res[i] = NAN;                        /* :192-193 */
    const double step = 200.0 / 3.0;
    med = nanmedian(res, n);
    for (size_t i = 0; i < n; i++) if (res[i] > med + 10) res[i] -= step;                           /* :208-209 */
    med = nanmedian(res, n);
    for (size_t i = 0; i < n; i++) if (res[i] > med - 10) res[i] += step;                           /* :210-211 as written */
    double s = 0, ss = 0;
    size_t m = 0;
    for (size_t i = 0; i < n; i++) if (!isnan(res[i])) { s += res[i]; m++; }
    const double mean = m ? s / (double)m : NAN;
    for (size_t i = 0; i < n; i++) if (!isnan(res[i])) ss += (res[i] - mean) * (res[i] - mean);
    const double sd = m > 1 ? sqrt(ss / (double)(m - 1)) : NAN;
    printf("resmean=%.6f resstd=%.6f n=%zu nan=%zu snrop=%.3f snrlt=%.3f\n", mean, sd, n, n - m, snrop, snrlt);

    char outname[4096];
    if (argc > 6 && argv[6][0]) snprintf(outname, sizeof outname, "%s", argv[6]);
    else {   /* [num2str(ladate(p)),'.1s']: num2str of a non-integer prints max(floor(log10(|x|))+5, 5) significant digits */
        int digits = (mjd != 0.0) ? (int)floor(log10(fabs(mjd))) + 5 : 5;
        if (digits < 5) digits = 5;
        if (digits > 16) digits = 16;
        if (mjd == floor(mjd)) snprintf(outname, sizeof outname, "%.0f.1s", mjd);
        else snprintf(outname, sizeof outname, "%.*g.1s", digits, mjd);
    }
    FILE* fo = fopen(outname, "w");
    if (!fo) { fprintf(stderr, "cannot write %s\n", outname); return 1; }
    fprintf(fo, "# MJD\t\tOPlocal\tOPremote\tLTFBlocal\tLTBBremote\n");   /* :253 */
    size_t cpt = 0;
    for (size_t k0 = 0; k0 + 25 < n_opre && k0 + 25 <= n; k0 += 25, cpt++)   /* for k=1:25:length(opre)-25 */
        fprintf(fo, "%f\t%f\t%f\t%f\t%f\n", mjd + (double)cpt / 86400.0, fit13(oplo, k0), fit13(opre, k0),
                fit13(ltlo, k0), fit13(ltre, k0));
    fclose(fo);
    fprintf(stderr, "%zu one-second rows -> %s\n", cpt, outname);
    mat5_free(&oplo_f.f); mat5_free(&opre_f.f); mat5_free(&ltlo_f.f); mat5_free(&ltre_f.f);
    return 0;
}
