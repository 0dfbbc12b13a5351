/*
 * TEST INFRASTRUCTURE — not part of the shipped product path.
 *
 * FFTW-API shim: the seven fftwf_* symbols the reference's CPU path links
 * against (mfcccpu.cpp:109-116 alloc+plan, :189 execute, :170-173 free/destroy/cleanup).
 * FFTW 3.3.x single precision (libfftw3f-3.lib, OpenCLProject3.vcxproj:127) is NOT
 * installed in this image and cannot be fetched, so the r2c transform itself is
 * restated here from its published definition
 *     X[k] = sum_j x[j] * exp(-2*pi*i*j*k/n),  k = 0..n/2, unnormalised
 * (FFTW manual, "The 1d Real-data DFT"; declared at include/fftw3.h:184).
 * Parity of this file is therefore pinned by the DFT definition
 * (tests/test_oracle_cpu.py checks it against a float64 direct DFT), not by an FFTW golden vector.
 *
 * Algorithm: half-length complex FFT (radix-2 DIT, tables computed in double) +
 * real split post-pass, vectorised ACROSS the `howmany` rows (8 rows per SIMD op)
 * so the CPU baseline is not handicapped by a naive DFT.
 * -DSHIM_DOUBLE computes in double and rounds once: the "most accurate float FFT" anchor.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifdef SHIM_DOUBLE
typedef double real;
#define VL 4
typedef double vreal __attribute__((vector_size(32)));
#else
typedef float real;
#define VL 8
typedef float vreal __attribute__((vector_size(32)));
#endif

typedef float fftwf_complex[2];

struct fftwf_plan_s {
    int n, m, logm, howmany, idist, odist;
    float *in;
    fftwf_complex *out;
    int *bitrev;      /* [m]                       */
    real *tw_re;      /* [m/2]  exp(-2 pi i j / m) */
    real *tw_im;
    real *pw_re;      /* [m/2+1] exp(-2 pi i k / n) */
    real *pw_im;
};
typedef struct fftwf_plan_s *fftwf_plan;

void *fftwf_malloc_aligned(size_t bytes)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) return NULL;
    return p;
}

float *fftwf_alloc_real(size_t n) { return (float *)fftwf_malloc_aligned(n * sizeof(float)); }
fftwf_complex *fftwf_alloc_complex(size_t n) { return (fftwf_complex *)fftwf_malloc_aligned(n * sizeof(fftwf_complex)); }
void fftwf_free(void *p) { free(p); }
void fftwf_cleanup(void) {}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    free(p->bitrev); free(p->tw_re); free(p->tw_im); free(p->pw_re); free(p->pw_im);
    free(p);
}

fftwf_plan fftwf_plan_many_dft_r2c(int rank, const int *n, int howmany,
                                   float *in, const int *inembed, int istride, int idist,
                                   fftwf_complex *out, const int *onembed, int ostride, int odist,
                                   unsigned flags)
{
    (void)inembed; (void)onembed; (void)flags;
    if (rank != 1 || istride != 1 || ostride != 1 || !n) return NULL;
    int N = n[0];
    if (N < 4 || (N & (N - 1))) return NULL; /* power of two only (the reference always passes ceil2(W)) */
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->n = N; p->m = N / 2; p->howmany = howmany; p->idist = idist; p->odist = odist;
    p->in = in; p->out = out;
    int m = p->m, logm = 0;
    while ((1 << logm) < m) logm++;
    p->logm = logm;
    p->bitrev = (int *)malloc(sizeof(int) * m);
    p->tw_re = (real *)malloc(sizeof(real) * (m / 2 + 1));
    p->tw_im = (real *)malloc(sizeof(real) * (m / 2 + 1));
    p->pw_re = (real *)malloc(sizeof(real) * (m / 2 + 1));
    p->pw_im = (real *)malloc(sizeof(real) * (m / 2 + 1));
    for (int i = 0; i < m; i++) {
        int r = 0;
        for (int b = 0; b < logm; b++) if (i & (1 << b)) r |= 1 << (logm - 1 - b);
        p->bitrev[i] = r;
    }
    for (int j = 0; j < m / 2; j++) {
        double a = -2.0 * M_PI * (double)j / (double)m;
        p->tw_re[j] = (real)cos(a); p->tw_im[j] = (real)sin(a);
    }
    for (int k = 0; k <= m / 2; k++) {
        double a = -2.0 * M_PI * (double)k / (double)N;
        p->pw_re[k] = (real)cos(a); p->pw_im[k] = (real)sin(a);
    }
    return p;
}

/* transform VL consecutive rows starting at row r0 (rows >= howmany are skipped on load/store) */
__attribute__((target_clones("avx2", "default")))
static void exec_group(const struct fftwf_plan_s *p, int r0, vreal *zr, vreal *zi)
{
    const int m = p->m, N = p->n;
    int nrows = p->howmany - r0; if (nrows > VL) nrows = VL;
    /* load + bit-reverse: z[j] = x[2j] + i x[2j+1] */
    for (int j = 0; j < m; j++) {
        vreal a, b;
        for (int l = 0; l < VL; l++) {
            const float *row = p->in + (size_t)(r0 + (l < nrows ? l : 0)) * p->idist;
            a[l] = row[2 * j]; b[l] = row[2 * j + 1];
        }
        zr[p->bitrev[j]] = a; zi[p->bitrev[j]] = b;
    }
    /* radix-2 DIT passes */
    for (int half = 1; half < m; half <<= 1) {
        int step = m / (2 * half);
        for (int k = 0; k < m; k += 2 * half) {
            for (int j = 0; j < half; j++) {
                real wr = p->tw_re[j * step], wi = p->tw_im[j * step];
                vreal xr = zr[k + j + half], xi = zi[k + j + half];
                vreal tr = xr * wr - xi * wi;
                vreal ti = xr * wi + xi * wr;
                vreal ur = zr[k + j], ui = zi[k + j];
                zr[k + j] = ur + tr; zi[k + j] = ui + ti;
                zr[k + j + half] = ur - tr; zi[k + j + half] = ui - ti;
            }
        }
    }
    /* real split: X[k] = (Z[k]+conj Z[m-k])/2 - i/2 * W_N^k * (Z[k]-conj Z[m-k]) */
    for (int k = 0; k <= m / 2; k++) {
        int k2 = (m - k) & (m - 1);
        vreal ar = zr[k], ai = zi[k], br = zr[k2], bi = zi[k2];
        vreal er = (ar + br) * (real)0.5, ei = (ai - bi) * (real)0.5;
        vreal orr = (ar - br) * (real)0.5, oi = (ai + bi) * (real)0.5;
        /* t = -i * W * o ; W = (wr, wi): W*o = (wr*or - wi*oi, wr*oi + wi*or); -i*(x+iy) = (y, -x) */
        real wr = p->pw_re[k], wi = p->pw_im[k];
        vreal pr = orr * wr - oi * wi, pi = oi * wr + orr * wi;
        vreal tr = pi, ti = -pr;
        vreal x1r = er + tr, x1i = ei + ti;       /* X[k]     */
        vreal x2r = er - tr, x2i = -(ei - ti);    /* X[m-k] = conj(E - T) */
        for (int l = 0; l < nrows; l++) {
            fftwf_complex *row = p->out + (size_t)(r0 + l) * p->odist;
            if (k == 0) {
                row[0][0] = (float)(ar[l] + ai[l]); row[0][1] = 0.0f;
                row[m][0] = (float)(ar[l] - ai[l]); row[m][1] = 0.0f;
            } else {
                row[k][0] = (float)x1r[l]; row[k][1] = (float)x1i[l];
                row[m - k][0] = (float)x2r[l]; row[m - k][1] = (float)x2i[l];
            }
        }
    }
    (void)N;
}

void fftwf_execute(const fftwf_plan p)
{
    vreal *zr = (vreal *)fftwf_malloc_aligned(sizeof(vreal) * p->m * 2);
    vreal *zi = zr + p->m;
    for (int r0 = 0; r0 < p->howmany; r0 += VL) exec_group(p, r0, zr, zi);
    free(zr);
}
