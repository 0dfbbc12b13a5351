// TEST INFRASTRUCTURE — not part of the shipped product path.
//
// C API over the REFERENCE'S OWN CPU classes, compiled from the sources where they lie
// under /root/reference (see oracle/Makefile; nothing is copied into this repo).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// the resulting oracle/_ref/libref_mfcc.so.
//
// Wrapped types: MfccCpu (mfcccpu.h:14-67), SegmenterCPU (segmentercpu.h:4-37),
// DeltaCPU (deltacpu.h:3-17), NormalizerCPU (normalizercpu.h:4-18).
// ref_extract() restates the reference driver's block loop (ASR_OCL.cpp:149-152 window,
// :156-161 buffer sizing, :227-301 set_input/apply/get_output_data/flush order) without
// libsndfile or text IO so it can be timed as the CPU baseline.
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>
#include <stdexcept>
#include <algorithm>
#include <malloc.h>

#include "mfcccpu.h"

static thread_local std::string g_err;

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// ---------------------------------------------------------------- MfccCpu
void *ref_mfcc_create(int input_buffer_size, int window_size, int shift, int num_banks,
                      float sample_rate, float low_freq, float high_freq, int ceps_len,
                      int want_c0, float lift_coef, int norm, int dyn, int delta_l1,
                      int delta_l2, int norm_after_dyn)
{
    try {
        return new MfccCpu(input_buffer_size, window_size, shift, num_banks, sample_rate, low_freq,
                           high_freq, ceps_len, want_c0 != 0, lift_coef, (Normalizer::norm_t)norm,
                           (ParamBase::dyn_t)dyn, delta_l1, delta_l2, norm_after_dyn != 0);
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void ref_mfcc_destroy(void *h) { delete (MfccCpu *)h; }
void ref_mfcc_set_window(void *h, const float *w) { ((MfccCpu *)h)->set_window(w); }
void ref_mfcc_set_alpha(void *h, float a) { ((MfccCpu *)h)->set_alpha(a); }
int ref_mfcc_input_buffer_size(void *h) { return ((MfccCpu *)h)->get_input_buffer_size(); }
int ref_mfcc_estimated_window_count(void *h, int samples) { return ((MfccCpu *)h)->estimated_window_count(samples); }
int ref_mfcc_width(void *h) { return ((MfccCpu *)h)->get_output_data_width(); }
int ref_mfcc_set_input(void *h, const short *d, int samples)
{
    try { return ((MfccCpu *)h)->set_input(d, samples); }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int ref_mfcc_flush(void *h)
{
    try { return ((MfccCpu *)h)->flush(); }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int ref_mfcc_apply(void *h)
{
    try { ((MfccCpu *)h)->apply(); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int ref_mfcc_get_output(void *h, float *out, int wc)
{
    try { ((MfccCpu *)h)->get_output_data(out, wc); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// ---------------------------------------------------------------- SegmenterCPU
void *ref_segmenter_create(int window_size, int shift, int window_limit, int deltasize)
{
    SegmenterCPU *s = new SegmenterCPU();
    s->init(window_size, shift, window_limit, deltasize);
    return s;
}
void ref_segmenter_destroy(void *h) { ((SegmenterCPU *)h)->cleanup(); delete (SegmenterCPU *)h; }
void ref_segmenter_set_window(void *h, const float *w) { ((SegmenterCPU *)h)->set_window(w); }
int ref_segmenter_set_input(void *h, const short *in, float *out, int samples, int *wc, int *wc_nd)
{
    try { ((SegmenterCPU *)h)->set_input(in, out, samples, *wc, *wc_nd); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int ref_segmenter_flush(void *h, float *out, int *wc, int *wc_nd)
{
    try { ((SegmenterCPU *)h)->flush(out, *wc, *wc_nd); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int ref_segmenter_remaining_samples(void *h) { return ((SegmenterCPU *)h)->get_remaining_samples(); }
int ref_segmenter_samples(void *h) { return ((SegmenterCPU *)h)->get_samples(); }
int ref_segmenter_is_flushed(void *h) { return ((SegmenterCPU *)h)->is_flushed(); }
int ref_segmenter_was_flushed(void *h) { return ((SegmenterCPU *)h)->was_flushed(); }

// ---------------------------------------------------------------- DeltaCPU
// in: [window_count + 2*delta_size][dim]  out: [window_count][dim]   (deltacpu.cpp:16-30)
void ref_delta_apply(const float *in, float *out, int dim, int window_count, int delta_size)
{
    DeltaCPU d;
    d.init(dim, window_count > 0 ? window_count : 1, delta_size);
    d.apply(in, window_count);
    memcpy(out, d.get_output_buffer(), sizeof(float) * (size_t)dim * window_count);
    d.cleanup();
}

// ---------------------------------------------------------------- NormalizerCPU
void *ref_normalizer_create(int norm, int dim)
{
    NormalizerCPU *n = new NormalizerCPU();
    n->init((Normalizer::norm_t)norm, dim);
    return n;
}
void ref_normalizer_destroy(void *h) { ((NormalizerCPU *)h)->cleanup(); delete (NormalizerCPU *)h; }
void ref_normalizer_normalize(void *h, float *data, int wc, int use_last_stats)
{
    ((NormalizerCPU *)h)->normalize(data, wc, use_last_stats != 0);
}

// ---------------------------------------------------------------- driver loop (timed CPU baseline)
// ASR_OCL.cpp:149-152
void ref_make_window(float *w, int window_size)
{
    for (int i = 0; i < window_size; i++)
        w[i] = (float)(0.56f - 0.46f * cos((2.0f * M_PI * i) / window_size)) / 32768.f;
}

struct ref_params {
    int window_size, shift, num_banks;
    float sample_rate, low_freq, high_freq;
    int ceps_len, want_c0;
    float lift_coef;
    int norm, dyn, delta_l1, delta_l2, norm_after_dyn;
    float alpha;
};

// One utterance through the reference driver loop. sample_limit <= 0 means "the utterance length"
// (Q5: the FFT plan always transforms window_limit rows, so the baseline sizes it to the utterance).
// Returns frames written to out (row-major, pitch = width) or -1.
static long extract_one(const ref_params &p, const float *window, const short *pcm, long n,
                        int sample_limit, float *out, long out_capacity_frames)
{
    int limit = sample_limit > 0 ? sample_limit : (int)n;
    MfccCpu m(limit, p.window_size, p.shift, p.num_banks, p.sample_rate, p.low_freq, p.high_freq,
              p.ceps_len, p.want_c0 != 0, p.lift_coef, (Normalizer::norm_t)p.norm,
              (ParamBase::dyn_t)p.dyn, p.delta_l1, p.delta_l2, p.norm_after_dyn != 0);
    m.set_window(window);
    m.set_alpha(p.alpha);
    const int width = m.get_output_data_width();
    const int in_limit = m.get_input_buffer_size();
    long done = 0, pos = 0;
    while (pos < n) {
        int s = (int)std::min<long>(n - pos, in_limit);
        int wc = m.set_input(pcm + pos, s);
        m.apply();
        if (wc > 0) {
            if (done + wc > out_capacity_frames) throw std::runtime_error("ref_extract: output too small");
            m.get_output_data(out + done * width, wc);
            done += wc;
        }
        pos += s;
    }
    int wc = m.flush();
    if (wc > 0) {
        m.apply();
        if (done + wc > out_capacity_frames) throw std::runtime_error("ref_extract: output too small");
        m.get_output_data(out + done * width, wc);
        done += wc;
    }
    return done;
}

// offsets: [n_utts+1] sample offsets into pcm; out_offsets: [n_utts+1] FRAME offsets into out.
// frames_out (optional): [n_utts] frames produced. seconds: wall time of the extraction loop only.
int ref_extract(const ref_params *p, const short *pcm, const long long *offsets, int n_utts,
                float *out, const long long *out_offsets, int sample_limit, int n_threads,
                long long *frames_out, double *seconds)
{
    // The reference needs a fresh MfccCpu per utterance (Q3). Keep its MB-sized buffers on the malloc heap instead of
    // mmap/munmap per object, otherwise worker threads serialise on the kernel's mm lock and the baseline stops scaling.
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
    std::vector<float> window(p->window_size);
    ref_make_window(window.data(), p->window_size);
    int cols = p->ceps_len > 0 ? p->ceps_len + (p->want_c0 ? 1 : 0) : p->num_banks;
    int width = cols * (p->dyn == 2 ? 3 : p->dyn == 1 ? 2 : 1);
    std::atomic<int> next(0);
    std::atomic<int> failed(0);
    std::string err;
    if (n_threads < 1) n_threads = 1;
    auto worker = [&]() {
        for (;;) {
            int u = next.fetch_add(1);
            if (u >= n_utts) break;
            try {
                long cap = (long)(out_offsets[u + 1] - out_offsets[u]);
                long got = extract_one(*p, window.data(), pcm + offsets[u], (long)(offsets[u + 1] - offsets[u]),
                                       sample_limit, out + out_offsets[u] * width, cap);
                if (frames_out) frames_out[u] = got;
            } catch (const std::exception &e) {
                if (!failed.exchange(1)) err = e.what();
                if (frames_out) frames_out[u] = -1;
            }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (n_threads == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; i++) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (failed.load()) { g_err = err; return -1; }
    return 0;
}

} // extern "C"
