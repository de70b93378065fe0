/* Minimal MAT-file level 5 writer (uncompressed): real / complex double row vectors.
 * Enough for the result records the reference scripts save with `save -mat`
 * (processing/Octave/godual_ranging.m:128-130, acquisition/claudio_aligned_code_lo_separate.m:166);
 * scipy.io.loadmat / Octave `load` read the files.  matio is not available in this image. */
#ifndef TWSTFT_MATV5_H
#define TWSTFT_MATV5_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>

enum { MI_INT8 = 1, MI_INT32 = 5, MI_UINT32 = 6, MI_DOUBLE = 9, MI_MATRIX = 14, MX_DOUBLE_CLASS = 6 };

static inline int mat5_open(FILE* f, const char* note)
{
    char hdr[128];
    memset(hdr, ' ', sizeof hdr);
    const char* base = "MATLAB 5.0 MAT-file, written by twstft godual: ";
    size_t n = strlen(base);
    memcpy(hdr, base, n);
    size_t m = strlen(note);
    if (m > 116 - n) m = 116 - n;
    memcpy(hdr + n, note, m);
    memset(hdr + 116, 0, 8);                 /* subsystem data offset */
    const uint16_t version = 0x0100, endian = 0x4D49;   /* "IM" read back as "MI" => little endian */
    memcpy(hdr + 124, &version, 2);
    memcpy(hdr + 126, &endian, 2);
    return fwrite(hdr, 1, 128, f) == 128 ? 0 : -1;
}

static inline void mat5_tag(FILE* f, uint32_t type, uint32_t bytes)
{
    fwrite(&type, 4, 1, f);
    fwrite(&bytes, 4, 1, f);
}
static inline void mat5_pad(FILE* f, size_t bytes)
{
    static const char z[8] = {0};
    if (bytes % 8) fwrite(z, 1, 8 - bytes % 8, f);
}

/* rows x cols double matrix (column-major data), im may be NULL */
static inline int mat5_matrix(FILE* f, const char* name, const double* re, const double* im, size_t rows, size_t cols)
{
    const size_t n = rows * cols;
    const size_t nl = strlen(name), nlp = (nl + 7) / 8 * 8;
    const size_t data = 8 + n * 8;           /* n*8 is already a multiple of 8 */
    const size_t total = 16 + 16 + 8 + nlp + data * (im ? 2 : 1);
    mat5_tag(f, MI_MATRIX, (uint32_t)total);
    mat5_tag(f, MI_UINT32, 8);
    const uint32_t flags[2] = {(uint32_t)MX_DOUBLE_CLASS | (im ? 0x0800u : 0u), 0};
    fwrite(flags, 4, 2, f);
    mat5_tag(f, MI_INT32, 8);
    const int32_t dims[2] = {(int32_t)rows, (int32_t)cols};
    fwrite(dims, 4, 2, f);
    mat5_tag(f, MI_INT8, (uint32_t)nl);
    fwrite(name, 1, nl, f);
    mat5_pad(f, nl);
    mat5_tag(f, MI_DOUBLE, (uint32_t)(n * 8));
    fwrite(re, 8, n, f);
    if (im) {
        mat5_tag(f, MI_DOUBLE, (uint32_t)(n * 8));
        fwrite(im, 8, n, f);
    }
    return ferror(f) ? -1 : 0;
}
/* 1 x n double vector, im may be NULL */
static inline int mat5_vector(FILE* f, const char* name, const double* re, const double* im, size_t n)
{
    return mat5_matrix(f, name, re, im, 1, n);
}
#endif
