/* gofinal -- host-only text reducers of the reference, in C (no Octave, no GPU library):
 *
 *   gofinal op   local.mat[.gz] remote.mat[.gz] end_epoch out.txt      processing/Octave/gofinal_op.m:24-81
 *   gofinal ltfb local.mat[.gz] remote.mat[.gz] end_epoch out.txt      processing/Octave/gofinal_ltfb.m:27-93
 *   gofinal tw   op.txt ltfb.txt                                       processing/Octave/tw.m:11-50
 *
 * op / ltfb: one station's local (two-channel) and remote result files -> the per-second text table
 *   "% y  m  d  h  m  s<TAB>delay<TAB><TAB>df1<TAB>SNR1<TAB>df2<TAB>SNR2<TAB>delayrem<TAB><TAB>df1rem<TAB>SNR1rem\r\n"
 *   rows "%Y %m %d %H %M %S\t%.12f\t%.3f\t%.1f\t%.12f\t%.3f\t%.1f\t%.12f\t%.3f\t%.1f\n"   (gofinal_ltfb.m:84-90)
 *   A file whose PATH contains a 'C' is taken as the C++ program's flavour (correction = indice0+corr, SNR in dB,
 *   gofinal_ltfb.m:35-45); otherwise indice+correction and 10*log10(SNRr+SNRi).  `op` accepts the older variable
 *   name correction1_a (gofinal_op.m:30,38-39).  The scripts find the files and the end date by walking site
 *   directories; here they are arguments.  Times are localtime() like Octave's strftime(localtime()).
 *   Selection (both scripts): records with SNR1 > max-10 dB, quadratic fit of the local two-channel difference, and
 *   when its residual std exceeds 1 us the larger of the two clusters around res(30:40) / res(140:150).  `op` then
 *   writes EVERY record (:74), `ltfb` only the selected ones (:87) -- including the reference's quirk that after the
 *   cluster step the kept numbers index the already filtered list.
 *   stdout: "stdOP=%f\t<OP>=%f\n" / "stdLTFB=%f\t<LTFB>=%f\n".
 * tw: pairs the two tables as tw.m does (its date arithmetic literally), applies the two SNR criteria, the +1 s
 *   unwrap, res = 0.5*((ltfb10-ltfb13)-(op10-op13)), keeps |res| < 1 us and prints "diff=%d\n", std and mean in ns.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "matv5_read.h"

static double mean_of(const double* x, size_t n)
{
    long double s = 0;
    for (size_t i = 0; i < n; i++) s += x[i];
    return n ? (double)(s / n) : NAN;
}
static double std_of(const double* x, size_t n)       /* Octave std: n-1 */
{
    if (n < 2) return 0.0;
    const double m = mean_of(x, n);
    long double s = 0;
    for (size_t i = 0; i < n; i++) s += ((long double)x[i] - m) * ((long double)x[i] - m);
    return (double)sqrtl(s / (n - 1));
}
/* residual of the least-squares quadratic through (1..n, y): y - b.yf of [a,b]=polyfit(1:n,y,2) */
static void quad_residual(const double* y, size_t n, double* r)
{
    long double S[5] = {0}, T[3] = {0};
    const long double xc = (n + 1) / 2.0L, ys = mean_of(y, n);
    for (size_t i = 0; i < n; i++) {
        const long double x = (long double)(i + 1) - xc, v = (long double)y[i] - ys;
        long double p = 1;
        for (int k = 0; k < 5; k++) { S[k] += p; if (k < 3) T[k] += p * v; p *= x; }
    }
    long double A[3][4] = {{S[0], S[1], S[2], T[0]}, {S[1], S[2], S[3], T[1]}, {S[2], S[3], S[4], T[2]}};
    for (int c = 0; c < 3; c++) {                   /* Gauss-Jordan with partial pivoting */
        int piv = c;
        for (int q = c + 1; q < 3; q++) if (fabsl(A[q][c]) > fabsl(A[piv][c])) piv = q;
        for (int k = 0; k < 4; k++) { long double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        if (A[c][c] == 0) continue;
        for (int q = 0; q < 3; q++) {
            if (q == c) continue;
            const long double f = A[q][c] / A[c][c];
            for (int k = c; k < 4; k++) A[q][k] -= f * A[c][k];
        }
    }
    long double co[3];
    for (int c = 0; c < 3; c++) co[c] = A[c][c] != 0 ? A[c][3] / A[c][c] : 0;
    for (size_t i = 0; i < n; i++) {
        const long double x = (long double)(i + 1) - xc;
        r[i] = (double)((long double)y[i] - ys - (co[0] + co[1] * x + co[2] * x * x));
    }
}

typedef struct { size_t n; double *sol1, *sol2, *snr1, *snr2, *df1, *df2; size_t codelen; } station;

static const double* need(const mat5_file* mf, const char* a, const char* b, size_t* n)
{
    const mat5_var* v = mat5_find(mf, a);
    if (!v && b) v = mat5_find(mf, b);
    if (!v) { fprintf(stderr, "variable %s missing\n", a); exit(1); }
    if (*n == 0 || v->n < *n) *n = v->n;
    return v->re;
}
static void load_station(const char* path, int two, int ltfb, station* s)
{
    mat5_file mf;
    if (mat5_load(path, &mf) != 0) { fprintf(stderr, "cannot read %s\n", path); exit(1); }
    /* "a C in the name = written by the C++ program" exists in gofinal_ltfb.m:35 only (gofinal_op.m always takes the
     * Octave variables); the reference tests the whole path string, here the FILE name, so that a directory such as
     * /home/Claudio/ does not flip the flavour */
    const char* base = strrchr(path, '/');
    base = base ? base + 1 : path;
    const int cflavour = ltfb && strchr(base, 'C') != NULL;
    memset(s, 0, sizeof *s);
    size_t n = 0;
    const double *c1 = need(&mf, "correction1_a", "correction1", &n), *d1 = need(&mf, "df1", NULL, &n);
    const double *c2 = NULL, *d2 = NULL, *i1 = NULL, *i2 = NULL, *r1 = NULL, *q1 = NULL, *r2 = NULL, *q2 = NULL, *s1 = NULL, *s2 = NULL;
    if (two) { c2 = need(&mf, "correction2_a", "correction2", &n); d2 = need(&mf, "df2", NULL, &n); }
    if (cflavour) {
        s1 = need(&mf, "SNR1", NULL, &n);
        if (two) s2 = need(&mf, "SNR2", NULL, &n);
    } else {
        i1 = need(&mf, "indice1", NULL, &n); r1 = need(&mf, "SNR1r", NULL, &n); q1 = need(&mf, "SNR1i", NULL, &n);
        if (two) { i2 = need(&mf, "indice2", NULL, &n); r2 = need(&mf, "SNR2r", NULL, &n); q2 = need(&mf, "SNR2i", NULL, &n); }
    }
    s->n = n;
    double** out[6] = {&s->sol1, &s->sol2, &s->snr1, &s->snr2, &s->df1, &s->df2};
    for (int k = 0; k < 6; k++) *out[k] = (double*)calloc(n ? n : 1, sizeof(double));
    for (size_t p = 0; p < n; p++) {
        s->sol1[p] = cflavour ? c1[p] : i1[p] + c1[p];
        s->snr1[p] = cflavour ? s1[p] : 10 * log10(r1[p] + q1[p]);
        s->df1[p] = d1[p];
        if (two) {
            s->sol2[p] = cflavour ? c2[p] : i2[p] + c2[p];
            s->snr2[p] = cflavour ? s2[p] : 10 * log10(r2[p] + q2[p]);
            s->df2[p] = d2[p];
        }
    }
    const mat5_var* code = mat5_find(&mf, "code");
    s->codelen = code ? code->n : 5000000;        /* global code=ones(fs,1) when the file has none, gofinal_ltfb.m:6 */
    mat5_free(&mf);
}

static int reduce_station(int ltfb, const char* local, const char* remote, double end_epoch, const char* outpath)
{
    const double fs = 5e6, nint3 = 3.0;
    station lo, re;
    load_station(local, 1, ltfb, &lo);
    load_station(remote, 0, ltfb, &re);
    const size_t n = lo.n;
    if (n == 0) { fprintf(stderr, "empty result file\n"); return 1; }
    /* gofinal_ltfb.m:48 dateini=dateltfb-length(indice1) of the LOCAL file; gofinal_op.m:36-37 takes length(indice1)
     * while the REMOTE file is the one loaded */
    const double dateini = end_epoch - (double)(ltfb ? n : re.n);
    double mx = -INFINITY;
    for (size_t p = 0; p < n; p++) if (lo.snr1[p] > mx) mx = lo.snr1[p];
    size_t* ind = (size_t*)malloc(sizeof(size_t) * n);
    double* res = (double*)malloc(sizeof(double) * n);
    double* rr = (double*)malloc(sizeof(double) * n);
    size_t m = 0;
    for (size_t p = 0; p < n; p++)
        if (lo.snr1[p] > mx - 10) { ind[m] = p; res[m] = (lo.sol1[p] - lo.sol2[p]) / fs / nint3; m++; }
    quad_residual(res, m, rr);
    if (std_of(rr, m) > 1e-6 && m >= 150) {
        const double m1 = mean_of(res + 29, 11), m2 = mean_of(res + 139, 11);
        size_t n1 = 0, n2 = 0;
        for (size_t i = 0; i < m; i++) { n1 += fabs(res[i] - m1) < 1e-5; n2 += fabs(res[i] - m2) < 1e-5; }
        if (n1 && n2) {
            const double mm = n1 > n2 ? m1 : m2;
            size_t k = 0;
            for (size_t i = 0; i < m; i++)
                if (fabs(res[i] - mm) < 1e-5) { ind[k] = i; res[k] = res[i]; k++; }   /* indices into the filtered list */
            m = k;
            quad_residual(res, m, rr);
            printf("keep %zu samples ", m);
        }
    }
    for (size_t i = 0; i < m; i++) rr[i] *= 1e9;
    double moy = mean_of(res, m);
    if (moy < 0) moy += 1;
    printf(ltfb ? "stdLTFB=%f\t" : "stdOP=%f\t", std_of(rr, m));
    printf(ltfb ? "<LTFB>=%f\n" : "<OP>=%f\n", moy);

    FILE* f = fopen(outpath, "w");
    if (!f) { fprintf(stderr, "cannot write %s\n", outpath); return 1; }
    fprintf(f, "%% y  m  d  h  m  s\tdelay\t\tdf1\tSNR1\tdf2\tSNR2\tdelayrem\t\tdf1rem\tSNR1rem\r\n");
    const double step = ((double)lo.codelen / 2) / 2.5e6;
    const size_t nrows = ltfb ? m : n;
    for (size_t i = 0; i < nrows; i++) {
        const size_t p = ltfb ? ind[i] : i;
        if (p >= re.n) break;
        const time_t t = (time_t)floor(dateini + (double)p * step);
        struct tm tmv;
        localtime_r(&t, &tmv);
        char stamp[64];
        strftime(stamp, sizeof stamp, "%Y %m %d %H %M %S", &tmv);
        fprintf(f, "%s\t%.12f\t%.3f\t%.1f\t%.12f\t%.3f\t%.1f\t%.12f\t%.3f\t%.1f\n", stamp, lo.sol1[p] / nint3 / fs,
                lo.df1[p], lo.snr1[p], lo.sol2[p] / nint3 / fs, lo.df2[p], lo.snr2[p], re.sol1[p] / nint3 / fs, re.df1[p],
                re.snr1[p]);
    }
    fclose(f);
    return 0;
}

/* numeric table with '%' comment lines: rows of 15 columns */
typedef struct { size_t n; double (*r)[15]; } table;
static int load_table(const char* path, table* t)
{
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    size_t cap = 256;
    t->n = 0;
    t->r = malloc(sizeof(*t->r) * cap);
    char line[1024];
    while (fgets(line, sizeof line, f)) {
        if (line[0] == '%' || line[0] == '#') continue;
        double v[15];
        char* s = line;
        int k = 0;
        for (; k < 15; k++) { char* e; v[k] = strtod(s, &e); if (e == s) break; s = e; }
        if (k < 15) continue;
        if (t->n == cap) { cap *= 2; t->r = realloc(t->r, sizeof(*t->r) * cap); }
        memcpy(t->r[t->n++], v, sizeof v);
    }
    fclose(f);
    return 0;
}
static void keep_snr(table* a, table* b)            /* k=find(a(:,9)>max(a(:,9)-10)); k=[k(1)+1:k(end)-1]   tw.m:31-34 */
{
    double mx = -INFINITY;
    for (size_t i = 0; i < a->n; i++) if (a->r[i][8] > mx) mx = a->r[i][8];
    long first = -1, last = -1;
    for (size_t i = 0; i < a->n; i++) if (a->r[i][8] > mx - 10) { if (first < 0) first = (long)i; last = (long)i; }
    const long lo = first + 1, hi = last - 1;        /* inclusive, 0-based */
    size_t k = 0;
    for (long i = lo; i <= hi && first >= 0; i++, k++) {
        memcpy(a->r[k], a->r[i], sizeof a->r[0]);
        memcpy(b->r[k], b->r[i], sizeof b->r[0]);
    }
    a->n = b->n = k;
}
static int two_way(const char* opf, const char* ltf)
{
    table op, lt;
    if (load_table(opf, &op) || load_table(ltf, &lt) || !op.n || !lt.n) { fprintf(stderr, "cannot read the tables\n"); return 1; }
#define STAMP(t) (((((t).r[0][1] * 31 + (t).r[0][2]) * 24) + (t).r[0][3] * 60) + (t).r[0][4] * 60) + (t).r[0][5]   /* tw.m:16-17 */
    const double top = STAMP(op), tlt = STAMP(lt);
    long diff;
    table *late = top > tlt ? &op : &lt, *early = top > tlt ? &lt : &op;   /* the earlier table loses its head */
    diff = (long)fabs(top - tlt);
    if ((size_t)diff >= early->n) { printf(" diff=%ld\n", diff); return 0; }
    memmove(early->r, early->r + diff, sizeof(*early->r) * (early->n - diff));
    early->n -= diff;
    if (early->n > late->n) early->n = late->n;
    late->n = early->n;
    printf(" diff=%ld\n", diff);
    if (diff >= 180) return 0;
    keep_snr(&lt, &op);
    keep_snr(&op, &lt);
    const size_t n = op.n;
    if (!n) return 0;
    double *a = malloc(sizeof(double) * n), *b = malloc(sizeof(double) * n), *res = malloc(sizeof(double) * n);
    for (size_t i = 0; i < n; i++) { a[i] = lt.r[i][9] - lt.r[i][12]; b[i] = op.r[i][9] - op.r[i][12]; }
    const double addl = mean_of(a, n) < 0 ? 1.0 : 0.0, addo = mean_of(b, n) < 0 ? 1.0 : 0.0;   /* tw.m:35-36 */
    size_t k = 0;
    double all = 0;
    for (size_t i = 0; i < n; i++) {
        const double r = 0.5 * ((a[i] + addl) - (b[i] + addo));
        all += r;
        if (fabs(r) < 1e-6) res[k++] = r * 1e9;
    }
    if (!k) { printf("%.12f\n", all / n * 1e9); return 0; }
    printf("std=%.6f ns\nmean=%.6f ns\nn=%zu\n", std_of(res, k), mean_of(res, k), k);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc == 6 && (!strcmp(argv[1], "op") || !strcmp(argv[1], "ltfb")))
        return reduce_station(!strcmp(argv[1], "ltfb"), argv[2], argv[3], atof(argv[4]), argv[5]);
    if (argc == 4 && !strcmp(argv[1], "tw")) return two_way(argv[2], argv[3]);
    fprintf(stderr, "%s op|ltfb local.mat[.gz] remote.mat[.gz] end_epoch out.txt\n%s tw op.txt ltfb.txt\n", argv[0], argv[0]);
    return 1;
}
