/*
 * TEST INFRASTRUCTURE — not part of the shipped product path. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load liboracle_port.so.
 *
 * Plain-C restatement ("port") of the reference's CPU MFCC path, written from the behaviour of
 *   parambase.cpp:4-19, mfccbase.cpp:3-43, segmentercpu.cpp:17-106, mfcccpu.cpp:24-60,73-160,187-444,
 *   deltacpu.cpp:16-30, normalizercpu.cpp:22-89 and the driver loop ASR_OCL.cpp:149-152,227-301
 * (all paths relative to /root/reference). It keeps the reference's float/double expression order
 * (float libm as MSVC / `g++ -include math.h` bind it) so that it can be pinned
 * bit-for-bit against oracle/_ref/libref_mfcc.so (tests/test_oracle_cpu.py) and against the committed
 * golden vectors (tests/golden/golden_v1.npz, generated from that library).
 * The r2c FFT is oracle/fftw_shim.c (FFTW itself is not in the image; see that file's header).
 *
 * State is one struct per stream; buffers are sized like the reference sizes them
 * (window_limit = input_window_limit + 2 [+ 3*(l1+l2)], mfcccpu.cpp:95-103).
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef float fftwf_complex[2];
typedef struct fftwf_plan_s *fftwf_plan;
float *fftwf_alloc_real(size_t n);
fftwf_complex *fftwf_alloc_complex(size_t n);
fftwf_plan fftwf_plan_many_dft_r2c(int, const int *, int, float *, const int *, int, int, fftwf_complex *,
                                   const int *, int, int, unsigned);
void fftwf_execute(const fftwf_plan);
void fftwf_free(void *);
void fftwf_destroy_plan(fftwf_plan);

enum { NORM_NONE, NORM_CMN, NORM_CVN, NORM_MINMAX }; /* normalizer.h:5 */
enum { DYN_NONE, DYN_DELTA, DYN_ACC };               /* parambase.h:9  */

static __thread char g_err[256];
const char *port_last_error(void) { return g_err; }
static int fail(const char *m) { snprintf(g_err, sizeof g_err, "%s", m); return -1; }

static unsigned pow2_at_least(unsigned v) { unsigned p = 1; while (p < v) p <<= 1; return p; } /* mfcccpu.cpp:10-20 */

/* floor(float(samples - (W - S)) / S), evaluated in float like parambase.cpp:16-19 */
static int est_windows(int samples, int W, int S) { return (int)floorf((float)(samples - (W - S)) / S); }

/* ------------------------------------------------------------------ per-column normaliser (normalizercpu.cpp:22-89) */
typedef struct { int type, dim; float *mean, *scale; } port_norm;

static void norm_init(port_norm *n, int type, int dim)
{
    n->type = type; n->dim = dim;
    n->mean = (float *)calloc(dim > 0 ? dim : 1, sizeof(float));
    n->scale = (float *)calloc(dim > 0 ? dim : 1, sizeof(float));
}
static void norm_free(port_norm *n) { free(n->mean); free(n->scale); n->mean = n->scale = NULL; }

static void norm_apply(port_norm *n, float *x, int rows, int reuse_stats)
{
    const int d = n->dim;
    if (n->type == NORM_NONE) return;
    if (!reuse_stats) {
        for (int c = 0; c < d; c++) {
            double s = 0, s2 = 0;
            float lo = FLT_MAX, hi = -FLT_MAX;
            for (int r = 0; r < rows; r++) {
                float v = x[d * r + c];
                s += v;
                s2 += v * v;                 /* float product widened, as normalizercpu.cpp:45 */
                lo = fminf(lo, v); hi = fmaxf(hi, v);
            }
            n->mean[c] = (float)(s / rows);
            if (n->type == NORM_CVN)       /* unbiased inverse std, normalizercpu.cpp:48-49 */
                n->scale[c] = (float)sqrt((rows - 1) / (s2 - s * (s / rows)));
            else if (n->type == NORM_MINMAX) /* normalizercpu.cpp:65-66 */
                n->scale[c] = 1.f / fmaxf(fabsf(lo - n->mean[c]), fabsf(hi - n->mean[c]));
        }
    }
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < d; c++) {
            float v = x[d * r + c] - n->mean[c];
            x[d * r + c] = n->type == NORM_CMN ? v : v * n->scale[c];
        }
}

/* ------------------------------------------------------------------ regression deltas (deltacpu.cpp:16-30) */
/* in: [rows + 2L][dim], out: [rows][dim] */
static void delta_rows(const float *in, float *out, int dim, int rows, int L)
{
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < dim; c++) {
            float num = 0, den = 0;
            for (int l = 1; l <= L; l++) {
                num += l * (in[dim * (r + L + l) + c] - in[dim * (r + L - l) + c]);
                den += l * l;
            }
            out[dim * r + c] = num / (2 * den);
        }
}
void port_delta_apply(const float *in, float *out, int dim, int window_count, int delta_size)
{
    delta_rows(in, out, dim, window_count, delta_size);
}

/* ------------------------------------------------------------------ segmenter state machine (segmentercpu.cpp) */
typedef struct {
    int W, W2, S, D, remaining, samples, flushed, last_calc_flushed;
    size_t cap;
    short *carry;
    float *window;
} port_seg;

static void seg_init(port_seg *s, int W, int S, int window_limit, int D)
{
    memset(s, 0, sizeof *s);
    s->W = W; s->S = S; s->D = D; s->W2 = (int)pow2_at_least((unsigned)W);
    s->flushed = 1; s->last_calc_flushed = 0;
    s->cap = (size_t)window_limit * S + W - S;
    s->carry = (short *)malloc(sizeof(short) * (s->cap + 8));
    s->window = (float *)calloc(W, sizeof(float));
}
static void seg_free(port_seg *s) { free(s->carry); free(s->window); }

static void seg_frames(const port_seg *s, float *out, int frames) /* segmentercpu.cpp:17-28 */
{
    for (int f = 0; f < frames; f++)
        for (int j = 0; j < s->W; j++)
            out[(size_t)s->W2 * f + j] = s->window[j] * s->carry[f * s->S + j];
}

static int seg_set_input(port_seg *s, const short *in, float *out, int samples, int *wc, int *wc_nd)
{
    s->last_calc_flushed = s->flushed;
    /* The reference copies without a bound check (segmentercpu.cpp:61,78): carry-over + block can exceed the buffer of
     * window_limit * S + W - S samples (e.g. W > 2 S without deltas) and it overruns its heap. The port refuses instead. */
    if ((size_t)samples + (s->flushed ? 0 : (size_t)s->remaining) > s->cap)
        return fail("Can't process data, buffer is too small (the reference overruns its buffer here)");
    if (s->flushed) { /* first block of a stream: segmentercpu.cpp:59-75 */
        memcpy(s->carry, in, sizeof(short) * samples);
        *wc_nd = est_windows(samples, s->W, s->S);
        *wc = *wc_nd - s->D;
        if (*wc <= 0) return fail("Can't process data, window count is too small");
        seg_frames(s, out, *wc_nd);
        int used = (*wc - s->D) * s->S + s->W - s->S;
        if (used <= 0) return fail("Processed samples <= 0, this should never happen");
        s->remaining = samples - used + s->W - s->S;
        memmove(s->carry, s->carry + samples - s->remaining, sizeof(short) * s->remaining);
        s->flushed = 0;
    } else {          /* later block: append to the carry-over, segmentercpu.cpp:76-93 */
        memcpy(s->carry + s->remaining, in, sizeof(short) * samples);
        samples += s->remaining;
        *wc_nd = est_windows(samples, s->W, s->S);
        *wc = *wc_nd - 2 * s->D;
        if (*wc > 0) seg_frames(s, out, *wc_nd); else *wc = 0;
        int used = *wc * s->S + s->W - s->S;
        s->remaining = samples - used + s->W - s->S;
        memmove(s->carry, s->carry + samples - s->remaining, sizeof(short) * s->remaining);
    }
    s->samples = samples;
    return 0;
}

static void seg_flush(port_seg *s, float *out, int *wc, int *wc_nd) /* segmentercpu.cpp:97-106 */
{
    s->flushed = 1;
    *wc_nd = est_windows(s->remaining, s->W, s->S);
    *wc = *wc_nd - s->D;
    if (*wc > 0) seg_frames(s, out, *wc_nd);
}

void *port_segmenter_create(int W, int S, int window_limit, int D)
{
    port_seg *s = (port_seg *)malloc(sizeof *s);
    seg_init(s, W, S, window_limit, D);
    return s;
}
void port_segmenter_destroy(void *h) { seg_free((port_seg *)h); free(h); }
void port_segmenter_set_window(void *h, const float *w) { memcpy(((port_seg *)h)->window, w, sizeof(float) * ((port_seg *)h)->W); }
int port_segmenter_set_input(void *h, const short *in, float *out, int samples, int *wc, int *wc_nd)
{ return seg_set_input((port_seg *)h, in, out, samples, wc, wc_nd); }
int port_segmenter_flush(void *h, float *out, int *wc, int *wc_nd) { seg_flush((port_seg *)h, out, wc, wc_nd); return 0; }
int port_segmenter_remaining_samples(void *h) { return ((port_seg *)h)->remaining; }
int port_segmenter_samples(void *h) { return ((port_seg *)h)->samples; }
int port_segmenter_is_flushed(void *h) { return ((port_seg *)h)->flushed; }
int port_segmenter_was_flushed(void *h) { return ((port_seg *)h)->last_calc_flushed; }

void *port_normalizer_create(int norm, int dim) { port_norm *n = (port_norm *)malloc(sizeof *n); norm_init(n, norm, dim); return n; }
void port_normalizer_destroy(void *h) { norm_free((port_norm *)h); free(h); }
void port_normalizer_normalize(void *h, float *data, int wc, int use_last) { norm_apply((port_norm *)h, data, wc, use_last); }

/* ------------------------------------------------------------------ the MFCC stream object (mfcccpu.cpp) */
typedef struct {
    /* parameters (mfccbase.h:21-35) */
    int in_cap, in_frames_cap, W, S, W2, nb, C, ncep, l1, l2, D, norm, dyn, norm_after_dyn, want_c0, cols;
    float sr, lo, hi, lift, alpha;
    int last_block, frame_cap;
    float *frames, *mel, *cep, *dctm, *fbank, *dpad, *d1, *d2;
    int *edge;
    fftwf_complex *spec;
    fftwf_plan plan;
    port_seg seg;
    port_norm n0, n1, n2;
} port_mfcc;

static float hz_to_mel(float f) { return 1127 * logf(f / 700 + 1); }      /* mfcccpu.cpp:21 */
static float mel_to_hz(float m) { return 700 * (expf(m / 1127) - 1); }    /* mfcccpu.cpp:22 */

/* mel filterbank tables (mfcccpu.cpp:24-60): edge bins b_i and two interleaved weight rows (even / odd filters) */
static void build_filters(port_mfcc *m)
{
    const int nb = m->nb, N2 = m->W2;
    float *cent = (float *)malloc(sizeof(float) * (nb + 2));
    memset(m->fbank, 0, sizeof(float) * 2 * N2);
    float mlo = hz_to_mel(m->lo), mhi = hz_to_mel(m->hi);
    for (int i = 0; i < nb + 2; i++) {
        float f = mel_to_hz(i / (float)(nb + 1) * (mhi - mlo) + mlo);
        float o = 2 * (float)M_PI * f / m->sr;
        o = o + 2 * atanf(((1 - m->alpha) * sinf(o)) / (1 - (1 - m->alpha) * cosf(o))); /* VTLN bilinear warp */
        cent[i] = m->sr * o / (2 * (float)M_PI);
        m->edge[i] = (int)floor(cent[i] * N2 / m->sr + 0.5);
    }
    for (int i = 0; i < nb; i++) {
        float cl = cent[i], cc = cent[i + 1], cr = cent[i + 2];
        int il = (int)floor(N2 * cl / m->sr + 0.5), ir = (int)floor(N2 * cr / m->sr + 0.5);
        for (int j = il; j < ir; j++) {
            float up = (j * m->sr / (N2) - cl) / (cc - cl);
            float dn = (j * m->sr / (N2) - cr) / (cc - cr);
            m->fbank[(i % 2) * N2 + j] = fmaxf(0.0f, fminf(up, dn));
        }
    }
    free(cent);
}

void *port_mfcc_create(int input_buffer_size, int W, int S, int nb, float sr, float lo, float hi, int C,
                       int want_c0, float lift, int norm, int dyn, int l1, int l2, int norm_after_dyn)
{
    port_mfcc *m = (port_mfcc *)calloc(1, sizeof *m);
    m->W = W; m->S = S; m->nb = nb; m->sr = sr; m->lo = lo; m->hi = hi; m->C = C; m->want_c0 = want_c0 != 0;
    m->lift = lift; m->norm = norm; m->dyn = dyn; m->norm_after_dyn = norm_after_dyn != 0; m->alpha = 1;
    m->l1 = dyn != DYN_NONE ? l1 : 0;            /* mfccbase.cpp:26-27 */
    m->l2 = dyn == DYN_ACC ? l2 : 0;
    m->D = m->l1 + m->l2;
    m->ncep = want_c0 ? C + 1 : C;               /* m_dct_len, mfccbase.cpp:28 */
    m->cols = C > 0 ? m->ncep : nb;
    m->in_frames_cap = est_windows(input_buffer_size, W, S);   /* parambase.cpp:12-13 */
    m->in_cap = m->in_frames_cap * S + W - S;
    m->W2 = (int)pow2_at_least((unsigned)W);
    m->frame_cap = m->in_frames_cap + 2 + (dyn != DYN_NONE ? 3 * m->D : 0);
    seg_init(&m->seg, W, S, m->frame_cap, m->D);
    size_t n = (size_t)m->frame_cap * m->W2;
    m->frames = fftwf_alloc_real(n);
    memset(m->frames, 0, n * sizeof(float));
    m->spec = fftwf_alloc_complex(n);
    m->mel = (float *)malloc(sizeof(float) * nb * m->frame_cap);
    m->plan = fftwf_plan_many_dft_r2c(1, &m->W2, m->frame_cap, m->frames, NULL, 1, m->W2, m->spec, NULL, 1, m->W2, 0);
    if (!m->plan) { fail("Can't create FFTW plan."); return NULL; }
    if (C > 0) {                                 /* DCT-II + lifter, c0 in the LAST column (mfcccpu.cpp:118-136) */
        m->cep = (float *)malloc(sizeof(float) * m->ncep * m->frame_cap);
        m->dctm = (float *)calloc((size_t)nb * m->ncep, sizeof(float));
        float nf = (float)sqrt(2.0 / nb);
        for (int k = 0; k < nb; k++)
            for (int i = 1; i <= C; i++) {
                float lifter = (1 + lift / 2 * sinf((float)M_PI * (float)i / lift));
                m->dctm[m->ncep * k + i - 1] = lifter * nf * cosf((float)M_PI * i * (k + 0.5f) / nb);
            }
        if (want_c0) for (int k = 0; k < nb; k++) m->dctm[m->ncep * k + C] = nf;
    }
    norm_init(&m->n0, norm, m->cols); norm_init(&m->n1, norm, m->cols); norm_init(&m->n2, norm, m->cols);
    if (dyn != DYN_NONE) {
        m->dpad = (float *)malloc(sizeof(float) * m->cols * (m->frame_cap + 2 * m->D));
        m->d1 = (float *)malloc(sizeof(float) * m->cols * (m->frame_cap + 2 * m->l2));
        m->d2 = (float *)malloc(sizeof(float) * m->cols * m->frame_cap);
    }
    m->fbank = (float *)malloc(sizeof(float) * 2 * m->W2);
    m->edge = (int *)malloc(sizeof(int) * (nb + 3));
    m->edge[nb + 2] = -1;                         /* the reference reads one past the end here (Q4); keep it inert */
    build_filters(m);
    return m;
}

void port_mfcc_destroy(void *h)
{
    port_mfcc *m = (port_mfcc *)h;
    if (!m) return;
    seg_free(&m->seg); norm_free(&m->n0); norm_free(&m->n1); norm_free(&m->n2);
    fftwf_free(m->frames); fftwf_free(m->spec); fftwf_destroy_plan(m->plan);
    free(m->mel); free(m->cep); free(m->dctm); free(m->fbank); free(m->edge); free(m->dpad); free(m->d1); free(m->d2);
    free(m);
}
void port_mfcc_set_window(void *h, const float *w) { port_mfcc *m = (port_mfcc *)h; memcpy(m->seg.window, w, sizeof(float) * m->W); }
void port_mfcc_set_alpha(void *h, float a) { ((port_mfcc *)h)->alpha = a; }
int port_mfcc_input_buffer_size(void *h) { return ((port_mfcc *)h)->in_cap; }
int port_mfcc_estimated_window_count(void *h, int n) { port_mfcc *m = (port_mfcc *)h; return est_windows(n, m->W, m->S); }
int port_mfcc_width(void *h) { port_mfcc *m = (port_mfcc *)h; return m->cols * (m->dyn == DYN_ACC ? 3 : m->dyn == DYN_DELTA ? 2 : 1); }

int port_mfcc_set_input(void *h, const short *pcm, int samples) /* mfcccpu.cpp:338-346 */
{
    port_mfcc *m = (port_mfcc *)h;
    if (samples > m->in_cap) return fail("Can't process data, buffer is too small");
    int wc, wc_nd;
    if (seg_set_input(&m->seg, pcm, m->frames, samples, &wc, &wc_nd)) return -1;
    if (wc <= 0) return 0;
    fftwf_execute(m->plan);
    return wc;
}

int port_mfcc_flush(void *h) /* mfcccpu.cpp:348-369 */
{
    port_mfcc *m = (port_mfcc *)h;
    if (m->last_block) return 0;
    m->last_block = 1;
    int wc, wc_nd;
    seg_flush(&m->seg, m->frames, &wc, &wc_nd);
    if (wc <= 0) return 0;
    fftwf_execute(m->plan);
    return wc;
}

/* |X|/N2 -> triangular mel bank -> log (mfcccpu.cpp:192-220). Two running sums serve even and odd filters. */
static void mel_log(port_mfcc *m, int frames)
{
    build_filters(m);
    const int nb = m->nb, N2 = m->W2;
    for (int t = 0; t < frames; t++) {
        float acc[2] = {0, 0};
        int open = 0, last = m->edge[nb + 1];
        for (int j = m->edge[0]; j <= last; j++) {
            const float *c = m->spec[(size_t)N2 * t + j];
            float v = sqrtf(c[0] * c[0] + c[1] * c[1]) / N2;
            while (j == m->edge[open + 1]) {
                open++;
                if (open >= 2) {
                    m->mel[nb * t + open - 2] = logf(fmaxf(acc[open % 2], 1e-30f));
                    acc[open % 2] = 0;
                }
            }
            acc[0] += m->fbank[j] * v;
            acc[1] += m->fbank[N2 + j] * v;
        }
    }
}

static void cepstra(port_mfcc *m, int frames) /* mfcccpu.cpp:222-232 */
{
    for (int t = 0; t < frames; t++)
        for (int j = 0; j < m->ncep; j++) {
            float s = 0;
            for (int k = 0; k < m->nb; k++) s += m->mel[m->nb * t + k] * m->dctm[m->ncep * k + j];
            m->cep[m->ncep * t + j] = s;
        }
}

static float *static_rows(port_mfcc *m) { return m->C > 0 ? m->cep : m->mel; }

/* edge replication + delta + delta-delta (mfcccpu.cpp:234-263) */
static void dynamics(port_mfcc *m, int wc, int first, int last)
{
    if (m->dyn == DYN_NONE || (first && last)) return;
    const int d = m->cols, D = m->D;
    const float *src = static_rows(m);
    if (first) {
        memcpy(m->dpad + d * D, src, sizeof(float) * d * (wc + D));
        for (int i = 0; i < D; i++) memcpy(m->dpad + d * i, src, sizeof(float) * d);
    } else if (last) {
        memcpy(m->dpad, src, sizeof(float) * d * (wc + D));
        for (int i = 0; i < D; i++) memcpy(m->dpad + d * (i + wc + D), src + d * (wc + D - 1), sizeof(float) * d);
    } else
        memcpy(m->dpad, src, sizeof(float) * d * (wc + 2 * D));
    delta_rows(m->dpad, m->d1, d, wc + 2 * m->l2, m->l1);
    if (m->dyn == DYN_ACC) delta_rows(m->d1, m->d2, d, wc, m->l2);
}

static void normalise(port_mfcc *m, int wc, int reuse) /* mfcccpu.cpp:265-282 */
{
    if (m->norm == NORM_NONE) return;
    float *src = static_rows(m);
    if (m->norm_after_dyn) {
        norm_apply(&m->n0, m->seg.last_calc_flushed ? src : src + m->D * m->cols, wc, reuse);
        if (m->dyn != DYN_NONE) norm_apply(&m->n1, m->d1 + m->l2 * m->cols, wc, reuse);
        if (m->dyn == DYN_ACC) norm_apply(&m->n2, m->d2, wc, reuse);
    } else
        norm_apply(&m->n0, src, wc, reuse);
}

int port_mfcc_apply(void *h) /* the three-way case analysis of mfcccpu.cpp:371-425 */
{
    port_mfcc *m = (port_mfcc *)h;
    int wc_nd, wc, first = 0, last = 0, reuse = 0;
    if (m->last_block) {
        wc_nd = est_windows(m->seg.remaining, m->W, m->S); wc = wc_nd - m->D; last = 1; reuse = 1;
        if (wc <= 0) return 0;
    } else if (m->seg.last_calc_flushed) {
        wc_nd = est_windows(m->seg.samples, m->W, m->S); wc = wc_nd - m->D; first = 1;
        if (wc <= 0) return fail("Can't process data, window count is too small");
    } else {
        wc_nd = est_windows(m->seg.samples, m->W, m->S); wc = wc_nd - 2 * m->D;
        if (wc <= 0) return 0;
    }
    mel_log(m, wc_nd);
    if (m->C > 0) cepstra(m, wc_nd);
    if (!m->norm_after_dyn && m->norm != NORM_NONE) normalise(m, wc_nd, reuse);
    if (m->dyn != DYN_NONE) dynamics(m, wc, first, last);
    if (m->norm_after_dyn && m->norm != NORM_NONE) normalise(m, wc, reuse);
    return 0;
}

int port_mfcc_get_output(void *h, float *out, int wc) /* mfcccpu.cpp:427-444 */
{
    port_mfcc *m = (port_mfcc *)h;
    if (wc > m->frame_cap) return fail("Window count too high");
    const int d = m->cols, pitch = port_mfcc_width(h);
    const float *s0 = static_rows(m);
    if (!m->seg.last_calc_flushed) s0 += m->D * d;       /* Q1: still "flushed" after a single set_input */
    for (int r = 0; r < wc; r++) {
        memcpy(out + (size_t)r * pitch, s0 + (size_t)r * d, sizeof(float) * d);
        if (m->dyn != DYN_NONE) memcpy(out + (size_t)r * pitch + d, m->d1 + (size_t)(r + m->l2) * d, sizeof(float) * d);
        if (m->dyn == DYN_ACC) memcpy(out + (size_t)r * pitch + 2 * d, m->d2 + (size_t)r * d, sizeof(float) * d);
    }
    return 0;
}

/* ------------------------------------------------------------------ driver loop (ASR_OCL.cpp:149-152, 227-301) */
void port_make_window(float *w, int W)
{
    for (int i = 0; i < W; i++) w[i] = (float)(0.56f - 0.46f * cos((2.0f * M_PI * i) / W)) / 32768.f;
}

typedef struct {
    int window_size, shift, num_banks;
    float sample_rate, low_freq, high_freq;
    int ceps_len, want_c0;
    float lift_coef;
    int norm, dyn, delta_l1, delta_l2, norm_after_dyn;
    float alpha;
} port_params;

int port_extract(const port_params *p, const short *pcm, const long long *offsets, int n_utts, float *out,
                 const long long *out_offsets, int sample_limit, int n_threads, long long *frames_out, double *seconds)
{
    (void)n_threads; /* the port is a scalar single-thread checker */
    float *win = (float *)malloc(sizeof(float) * p->window_size);
    port_make_window(win, p->window_size);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int rc = 0;
    for (int u = 0; u < n_utts && rc == 0; u++) {
        long n = (long)(offsets[u + 1] - offsets[u]);
        void *m = port_mfcc_create(sample_limit > 0 ? sample_limit : (int)n, p->window_size, p->shift, p->num_banks,
                                   p->sample_rate, p->low_freq, p->high_freq, p->ceps_len, p->want_c0, p->lift_coef,
                                   p->norm, p->dyn, p->delta_l1, p->delta_l2, p->norm_after_dyn);
        if (!m) { rc = -1; break; }
        port_mfcc_set_window(m, win);
        port_mfcc_set_alpha(m, p->alpha);
        int width = port_mfcc_width(m), cap_in = port_mfcc_input_buffer_size(m);
        long cap = (long)(out_offsets[u + 1] - out_offsets[u]), done = 0, pos = 0;
        float *o = out + out_offsets[u] * width;
        const short *x = pcm + offsets[u];
        while (pos < n && rc == 0) {
            int s = (int)((n - pos) < cap_in ? (n - pos) : cap_in);
            int wc = port_mfcc_set_input(m, x + pos, s);
            if (wc < 0 || port_mfcc_apply(m)) { rc = -1; break; }
            if (wc > 0) {
                if (done + wc > cap) { rc = fail("port_extract: output too small"); break; }
                port_mfcc_get_output(m, o + done * width, wc);
                done += wc;
            }
            pos += s;
        }
        if (rc == 0) {
            int wc = port_mfcc_flush(m);
            if (wc > 0) {
                if (port_mfcc_apply(m)) rc = -1;
                else if (done + wc > cap) rc = fail("port_extract: output too small");
                else { port_mfcc_get_output(m, o + done * width, wc); done += wc; }
            }
        }
        if (frames_out) frames_out[u] = rc == 0 ? done : -1;
        port_mfcc_destroy(m);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    free(win);
    return rc;
}
