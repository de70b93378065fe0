/* Minimal MAT-file level 5 reader: named real / complex numeric vectors of an uncompressed (or whole-file gzip'ed,
 * as the reference stores its results: experiments/240527/op/ *.mat.gz) little-endian MAT-v5 file, values returned as
 * doubles.  Reads what Octave's `save -mat`, scipy.io.savemat and matv5.h write.  Compressed elements
 * (miCOMPRESSED, MATLAB's own `save`) are inflated with zlib.  matio is not available in this image. */
#ifndef TWSTFT_MATV5_READ_H
#define TWSTFT_MATV5_READ_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct { char name[64]; size_t n; double* re; double* im; } mat5_var;
typedef struct { mat5_var* v; size_t n; } mat5_file;

static size_t mat5_type_size(uint32_t t)
{
    switch (t) {
    case 1: case 2: return 1;            /* miINT8 miUINT8 */
    case 3: case 4: return 2;            /* miINT16 miUINT16 */
    case 5: case 6: case 7: return 4;    /* miINT32 miUINT32 miSINGLE */
    case 9: case 12: case 13: return 8;  /* miDOUBLE miINT64 miUINT64 */
    default: return 0;
    }
}
static double mat5_value(uint32_t t, const unsigned char* p)
{
    switch (t) {
    case 1: return (double)*(const int8_t*)p;
    case 2: return (double)*(const uint8_t*)p;
    case 3: { int16_t v; memcpy(&v, p, 2); return v; }
    case 4: { uint16_t v; memcpy(&v, p, 2); return v; }
    case 5: { int32_t v; memcpy(&v, p, 4); return v; }
    case 6: { uint32_t v; memcpy(&v, p, 4); return v; }
    case 7: { float v; memcpy(&v, p, 4); return v; }
    case 9: { double v; memcpy(&v, p, 8); return v; }
    case 12: { int64_t v; memcpy(&v, p, 8); return (double)v; }
    case 13: { uint64_t v; memcpy(&v, p, 8); return (double)v; }
    default: return 0.0;
    }
}
/* one data element at p: type, payload pointer and length; returns the offset of the next element */
static size_t mat5_element(const unsigned char* p, size_t avail, uint32_t* type, const unsigned char** data, size_t* len)
{
    if (avail < 8) return 0;
    uint32_t w0, w1;
    memcpy(&w0, p, 4);
    memcpy(&w1, p + 4, 4);
    if (w0 >> 16) {                       /* small data element: bytes in the upper half, data in the second word */
        *type = w0 & 0xffff;
        *len = w0 >> 16;
        *data = p + 4;
        return 8;
    }
    *type = w0;
    *len = w1;
    *data = p + 8;
    if (8 + (size_t)w1 > avail) return 0;
    return 8 + (w0 == 15 ? (size_t)w1 : (((size_t)w1 + 7) & ~(size_t)7));   /* compressed elements are not padded */
}

static double* mat5_numeric(const unsigned char* p, size_t avail, size_t count, size_t* used)
{
    uint32_t t;
    const unsigned char* d;
    size_t len;
    const size_t adv = mat5_element(p, avail, &t, &d, &len);
    const size_t sz = mat5_type_size(t);
    if (!adv || !sz || len / sz < count) return NULL;
    double* out = (double*)malloc(sizeof(double) * (count ? count : 1));
    for (size_t i = 0; i < count; i++) out[i] = mat5_value(t, d + i * sz);
    *used = adv;
    return out;
}

static int mat5_parse_matrix(mat5_file* mf, const unsigned char* p, size_t len)
{
    uint32_t t;
    const unsigned char* d;
    size_t l, off = 0, adv;
    if (!(adv = mat5_element(p, len, &t, &d, &l)) || l < 8) return -1;      /* array flags */
    uint32_t flags;
    memcpy(&flags, d, 4);
    const uint32_t cls = flags & 0xff;
    const int is_complex = (flags & 0x0800) != 0;
    off += adv;
    if (!(adv = mat5_element(p + off, len - off, &t, &d, &l))) return -1;   /* dimensions */
    size_t count = 1;
    for (size_t i = 0; i + 4 <= l; i += 4) { int32_t dim; memcpy(&dim, d + i, 4); count *= (size_t)(dim < 0 ? 0 : dim); }
    off += adv;
    if (!(adv = mat5_element(p + off, len - off, &t, &d, &l))) return -1;   /* name */
    mat5_var v;
    memset(&v, 0, sizeof v);
    memcpy(v.name, d, l < sizeof v.name - 1 ? l : sizeof v.name - 1);
    off += adv;
    if (cls < 6 || cls > 15) return 0;                                       /* cells, structs, chars, sparse: skipped */
    size_t used = 0;
    v.n = count;
    v.re = mat5_numeric(p + off, len - off, count, &used);
    if (!v.re) return -1;
    off += used;
    if (is_complex) {
        v.im = mat5_numeric(p + off, len - off, count, &used);
        if (!v.im) { free(v.re); return -1; }
    }
    mf->v = (mat5_var*)realloc(mf->v, sizeof(mat5_var) * (mf->n + 1));
    mf->v[mf->n++] = v;
    return 0;
}

/* Load every numeric variable of the file; returns 0 on success. */
static int mat5_load(const char* path, mat5_file* mf)
{
    memset(mf, 0, sizeof *mf);
    gzFile g = gzopen(path, "rb");        /* transparent for plain files */
    if (!g) return -1;
    size_t cap = 1 << 20, n = 0;
    unsigned char* buf = (unsigned char*)malloc(cap);
    for (;;) {
        if (n == cap) { cap *= 2; buf = (unsigned char*)realloc(buf, cap); }
        const int got = gzread(g, buf + n, (unsigned)(cap - n > (1u << 30) ? (1u << 30) : cap - n));
        if (got <= 0) break;
        n += (size_t)got;
    }
    gzclose(g);
    int rc = 0;
    if (n < 128 || memcmp(buf, "MATLAB 5.0 MAT-file", 19) != 0 || buf[126] != 'I' || buf[127] != 'M') rc = -2;
    size_t off = 128;
    while (rc == 0 && off + 8 <= n) {
        uint32_t t;
        const unsigned char* d;
        size_t l;
        const size_t adv = mat5_element(buf + off, n - off, &t, &d, &l);
        if (!adv) { rc = -3; break; }
        if (t == 14) rc = mat5_parse_matrix(mf, d, l);
        else if (t == 15) {               /* miCOMPRESSED: one inflated element, itself a miMATRIX */
            uLongf ucap = (uLongf)l * 8 + 4096;
            unsigned char* u = NULL;
            int zr;
            do {
                u = (unsigned char*)realloc(u, ucap);
                uLongf ulen = ucap;
                zr = uncompress(u, &ulen, d, (uLong)l);
                if (zr == Z_OK) {
                    uint32_t t2;
                    const unsigned char* d2;
                    size_t l2;
                    if (mat5_element(u, ulen, &t2, &d2, &l2) && t2 == 14) rc = mat5_parse_matrix(mf, d2, l2);
                }
                ucap *= 4;
            } while (zr == Z_BUF_ERROR);
            free(u);
            if (zr != Z_OK) rc = -4;
        }
        off += adv;
    }
    free(buf);
    return rc;
}

static const mat5_var* mat5_find(const mat5_file* mf, const char* name)
{
    for (size_t i = 0; i < mf->n; i++)
        if (!strcmp(mf->v[i].name, name)) return &mf->v[i];
    return NULL;
}

static void mat5_free(mat5_file* mf)
{
    for (size_t i = 0; i < mf->n; i++) { free(mf->v[i].re); free(mf->v[i].im); }
    free(mf->v);
    mf->v = NULL;
    mf->n = 0;
}
#endif
