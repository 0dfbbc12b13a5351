// Streaming objects of the C ABI: afe_mfcc (MfccOpenCL replacement), afe_segmenter, afe_delta, afe_normalizer.
// Host-side state machines follow the reference CPU classes (segmentercpu.cpp:56-106, mfcccpu.cpp:338-444); all
// arithmetic runs in the CUDA kernels of afe_stages.cu. One CUDA stream per object; every verb returns after the
// data it hands back is valid on the host (like the blocking reads of the OpenCL classes, mfccopencl.cpp:78-90).
#include <algorithm>
#include <cstring>
#include <memory>

#include "afe_internal.h"

using namespace afe;

namespace {

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count)
    {
        release();
        n = count;
        AFE_CUDA(cudaMalloc(&p, sizeof(T) * std::max<size_t>(count, 1)));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

// Carry-over bookkeeping shared by afe_segmenter and afe_mfcc (segmentercpu.cpp). PCM lives on the device in two
// ping-pong buffers: `cur` holds [carry-over | new block]; after the frames are cut the tail is copied to the front
// of the other buffer.
struct SegState {
    int W, S, N2, D, frame_cap;
    int remaining = 0, samples = 0;
    bool flushed = true, last_calc_flushed = false;
    size_t cap = 0;
    DevBuf<int16_t> buf[2];
    int16_t *h_pin = nullptr; // pinned staging so the H2D copy is truly asynchronous and the caller may reuse its buffer
    size_t pin_cap = 0;
    int cur = 0;

    void init(int W_, int S_, int frame_cap_, int D_)
    {
        W = W_; S = S_; D = D_; frame_cap = frame_cap_;
        N2 = afe_fft_size(W);
        cap = (size_t)frame_cap * S + W - S;                 // segmentercpu.cpp:40
        for (auto &b : buf) {
            b.alloc(cap + N2 + 16);                          // + slack: the FFT kernel reads whole N2-sample words
            AFE_CUDA(cudaMemset(b.p, 0, sizeof(int16_t) * (cap + N2 + 16)));
        }
        pin_cap = cap;
        AFE_CUDA(cudaMallocHost(&h_pin, sizeof(int16_t) * std::max<size_t>(pin_cap, 1)));
    }
    void release()
    {
        buf[0].release(); buf[1].release();
        if (h_pin) cudaFreeHost(h_pin);
        h_pin = nullptr;
    }
    void reset() { remaining = samples = 0; flushed = true; last_calc_flushed = false; }
    int est(int n) const { return afe_estimated_window_count(n, W, S); }

    // Returns the device pointer holding this block's contiguous PCM (valid until the next call).
    const int16_t *set_input(const int16_t *in, int n, int &wc, int &wc_nd, cudaStream_t st)
    {
        if ((size_t)n + (flushed ? 0 : remaining) > cap) throw Error("Can't process data, buffer is too small");
        last_calc_flushed = flushed;
        AFE_CUDA(cudaStreamSynchronize(st));                 // h_pin may still feed the previous copy
        memcpy(h_pin, in, sizeof(int16_t) * n);
        int16_t *dst = buf[cur].p;
        int total;
        if (flushed) {                                       // first block of a stream (segmentercpu.cpp:59-75)
            AFE_CUDA(cudaMemcpyAsync(dst, h_pin, sizeof(int16_t) * n, cudaMemcpyHostToDevice, st));
            total = n;
            wc_nd = est(total);
            wc = wc_nd - D;
            if (wc <= 0) throw Error("Can't process data, window count is too small");
            const int used = (wc - D) * S + W - S;
            if (used <= 0) throw Error("Processed samples <= 0, this should never happen");
            remaining = total - used + W - S;
            flushed = false;
        } else {                                             // append to the carry-over (segmentercpu.cpp:76-93)
            AFE_CUDA(cudaMemcpyAsync(dst + remaining, h_pin, sizeof(int16_t) * n, cudaMemcpyHostToDevice, st));
            total = n + remaining;
            wc_nd = est(total);
            wc = wc_nd - 2 * D;
            if (wc < 0) wc = 0;
            const int used = wc * S + W - S;
            remaining = total - used + W - S;
        }
        samples = total;
        return dst;
    }
    // after the frames of the current block were consumed: move the tail to the front of the other buffer
    void carry(cudaStream_t st)
    {
        const int16_t *src = buf[cur].p + samples - remaining;
        AFE_CUDA(cudaMemcpyAsync(buf[cur ^ 1].p, src, sizeof(int16_t) * remaining, cudaMemcpyDeviceToDevice, st));
        cur ^= 1;
    }
    const int16_t *flush(int &wc, int &wc_nd)                // segmentercpu.cpp:97-106
    {
        flushed = true;
        wc_nd = est(remaining);
        wc = wc_nd - D;
        return buf[cur].p;
    }
};

struct NormState {
    int type = AFE_NORM_NONE, dim = 0;
    DevBuf<float> mean, scale;
    void init(int t, int d) { type = t; dim = d; mean.alloc(d); scale.alloc(d); }
    void normalize(float *d_x, int rows, bool use_last, cudaStream_t st)   // normalizercpu.cpp:22-89
    {
        if (type == AFE_NORM_NONE || rows <= 0) return;
        if (!use_last) launch_colstats(d_x, rows, dim, type, mean.p, scale.p, st);
        launch_affine(d_x, rows, dim, type, mean.p, scale.p, st);
    }
};

} // namespace

// ================================================================================================== afe_mfcc
struct afe_mfcc {
    Derived d;
    int device;
    cudaStream_t st = nullptr;
    FftTables fft;
    MelTables mel;
    SegState seg;
    NormState n0, n1, n2;
    float alpha = 1.f;
    bool window_set = false, last_block = false, fix_q1 = false;
    DevBuf<float> mag, melv, cep, dpad, d1, d2, outb;
    float *h_out = nullptr; size_t h_out_cap = 0;
    afe_mfcc(const afe_params &p, int dev) : d(p), device(dev) {}
    float *statics() { return d.C > 0 ? cep.p : melv.p; }
    // Row offset of this block's first OUTPUT static: the reference keys it on was_flushed() (mfcccpu.cpp:274,439),
    // which is still true when flushing after a single set_input (quirk Q1) unless the fix is requested.
    int static_row_offset() const { return (!seg.last_calc_flushed || (fix_q1 && last_block)) ? d.D : 0; }
};

extern "C" {

int afe_mfcc_create(const afe_params *p, int cuda_device, afe_mfcc **out)
{
    *out = nullptr;
    return guarded([&] {
        if (afe_device_count() <= cuda_device) throw Error("no usable CUDA device " + std::to_string(cuda_device) + " (the product has no CPU fallback)");
        std::unique_ptr<afe_mfcc> h(new afe_mfcc(*p, cuda_device));
        const Derived &d = h->d;
        if (d.in_frames_cap < 1) throw Error("input_buffer_size is smaller than one window");
        DeviceGuard g(cuda_device);
        AFE_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
        if (d.N2 == 512 || d.N2 == 256) h->fft.build(d.N2);
        h->seg.init(d.W, d.S, d.frame_cap, d.D);
        const size_t fc = d.frame_cap;
        h->mag.alloc(fc * d.bins);
        h->melv.alloc(fc * d.nb);
        if (d.C > 0) h->cep.alloc(fc * d.dct_len);
        if (d.p.norm != AFE_NORM_NONE) { h->n0.init(d.p.norm, d.cols); h->n1.init(d.p.norm, d.cols); h->n2.init(d.p.norm, d.cols); }
        if (d.p.dyn != AFE_DYN_NONE) {
            h->dpad.alloc((fc + 2 * d.D) * d.cols);          // m_delta_in, mfcccpu.cpp:148-156
            h->d1.alloc((fc + 2 * d.l2) * d.cols);
            h->d2.alloc(fc * d.cols);
        }
        h->outb.alloc(fc * d.width);
        h->h_out_cap = fc * d.width;
        AFE_CUDA(cudaMallocHost(&h->h_out, sizeof(float) * h->h_out_cap));
        *out = h.release();
    });
}

void afe_mfcc_destroy(afe_mfcc *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    h->fft.release(); h->mel.release(); h->seg.release();
    if (h->h_out) cudaFreeHost(h->h_out);
    cudaStreamDestroy(h->st);
    delete h;
}

int afe_mfcc_set_window(afe_mfcc *h, const float *window)
{
    return guarded([&] { DeviceGuard g(h->device); upload_window(h->d, window, h->mel, h->st); h->window_set = true; });
}
int afe_mfcc_set_alpha(afe_mfcc *h, float alpha) { h->alpha = alpha; return 0; }
int afe_mfcc_input_buffer_size(const afe_mfcc *h) { return h->d.in_cap; }
int afe_mfcc_estimated_window_count(const afe_mfcc *h, int samples) { return afe_estimated_window_count(samples, h->d.W, h->d.S); }
int afe_mfcc_output_width(const afe_mfcc *h) { return h->d.width; }
int afe_mfcc_set_option(afe_mfcc *h, int option, int value)
{
    if (option == AFE_OPT_FIX_FLUSH_STATICS) { h->fix_q1 = value != 0; return 0; }
    return fail("unknown option");
}
int afe_mfcc_reset(afe_mfcc *h)
{
    return guarded([&] { DeviceGuard g(h->device); AFE_CUDA(cudaStreamSynchronize(h->st)); h->seg.reset(); h->last_block = false; });
}

int afe_mfcc_set_input(afe_mfcc *h, const int16_t *data, int samples, int *frames)
{
    *frames = 0;
    return guarded([&] {
        if (!h->window_set) throw Error("set_window must be called before set_input");
        if (samples > h->d.in_cap) throw Error("Can't process data, buffer is too small");   // mfcccpu.cpp:338-339
        DeviceGuard g(h->device);
        int wc, wc_nd;
        const int16_t *pcm = h->seg.set_input(data, samples, wc, wc_nd, h->st);
        if (wc > 0) launch_fft_mag(h->d, h->fft, h->mel, pcm, h->mag.p, wc_nd, h->st);      // segment + fft fused
        h->seg.carry(h->st);
        *frames = wc > 0 ? wc : 0;
    });
}

int afe_mfcc_flush(afe_mfcc *h, int *frames)
{
    *frames = 0;
    return guarded([&] {
        if (h->last_block) return;                             // nothing to flush (mfcccpu.cpp:350-351)
        h->last_block = true;
        DeviceGuard g(h->device);
        int wc, wc_nd;
        const int16_t *pcm = h->seg.flush(wc, wc_nd);
        if (wc <= 0) return;
        launch_fft_mag(h->d, h->fft, h->mel, pcm, h->mag.p, wc_nd, h->st);
        *frames = wc;
    });
}

// do_delta (mfcccpu.cpp:234-263)
static void dynamics(afe_mfcc *h, int wc, bool first, bool last)
{
    const Derived &d = h->d;
    if (d.p.dyn == AFE_DYN_NONE || (first && last)) return;
    const int D = d.D;
    const float *src = h->statics();
    if (first) launch_pad_rows(src, h->dpad.p, wc + D, d.cols, D, 0, h->st);
    else if (last) launch_pad_rows(src, h->dpad.p, wc + D, d.cols, 0, D, h->st);
    else launch_pad_rows(src, h->dpad.p, wc + 2 * D, d.cols, 0, 0, h->st);
    launch_delta(h->dpad.p, h->d1.p, wc + 2 * d.l2, d.cols, d.l1, h->st);
    if (d.p.dyn == AFE_DYN_ACC) launch_delta(h->d1.p, h->d2.p, wc, d.cols, d.l2, h->st);
}

// MfccCpu::normalize (mfcccpu.cpp:265-282)
static void normalise(afe_mfcc *h, int wc, bool use_last)
{
    const Derived &d = h->d;
    if (d.p.norm == AFE_NORM_NONE) return;
    float *src = h->statics();
    if (d.p.norm_after_dyn) {
        h->n0.normalize(src + h->static_row_offset() * d.cols, wc, use_last, h->st);
        if (d.p.dyn != AFE_DYN_NONE) h->n1.normalize(h->d1.p + d.l2 * d.cols, wc, use_last, h->st);
        if (d.p.dyn == AFE_DYN_ACC) h->n2.normalize(h->d2.p, wc, use_last, h->st);
    } else
        h->n0.normalize(src, wc, use_last, h->st);
}

int afe_mfcc_apply(afe_mfcc *h)
{
    return guarded([&] {
        const Derived &d = h->d;
        DeviceGuard g(h->device);
        int wc_nd, wc;
        bool first = false, last = false, use_last = false;
        if (h->last_block) {                                   // mfcccpu.cpp:373-390
            wc_nd = h->seg.est(h->seg.remaining); wc = wc_nd - d.D; last = true; use_last = true;
            if (wc <= 0) return;
        } else if (h->seg.last_calc_flushed) {                 // :391-407
            wc_nd = h->seg.est(h->seg.samples); wc = wc_nd - d.D; first = true;
            if (wc <= 0) throw Error("Can't process data, window count is too small");
        } else {                                               // :408-424
            wc_nd = h->seg.est(h->seg.samples); wc = wc_nd - 2 * d.D;
            if (wc <= 0) return;
        }
        if (h->mel.alpha_built != h->alpha) upload_mel_tables(d, h->alpha, h->mel, h->st);  // refresh_filters per alpha
        launch_mel_dct(d, h->mel, h->mag.p, h->melv.p, h->cep.p, wc_nd, h->st);
        const bool norm = d.p.norm != AFE_NORM_NONE;
        if (!d.p.norm_after_dyn && norm) normalise(h, wc_nd, use_last);
        if (d.p.dyn != AFE_DYN_NONE) dynamics(h, wc, first, last);
        if (d.p.norm_after_dyn && norm) normalise(h, wc, use_last);
    });
}

int afe_mfcc_get_output(afe_mfcc *h, float *out, int frames)
{
    return guarded([&] {
        const Derived &d = h->d;
        if (frames > d.frame_cap) throw Error("Window count too high");   // mfcccpu.cpp:429-430
        if (frames <= 0) return;
        DeviceGuard g(h->device);
        const float *s0 = h->statics() + h->static_row_offset() * d.cols;
        const int ns = d.width / d.cols;
        launch_pack(s0, ns > 1 ? h->d1.p + d.l2 * d.cols : nullptr, ns > 2 ? h->d2.p : nullptr, h->outb.p, frames, d.cols, ns, h->st);
        const size_t n = (size_t)frames * d.width;
        AFE_CUDA(cudaMemcpyAsync(h->h_out, h->outb.p, sizeof(float) * n, cudaMemcpyDeviceToHost, h->st));
        AFE_CUDA(cudaStreamSynchronize(h->st));
        memcpy(out, h->h_out, sizeof(float) * n);
    });
}

} // extern "C"

// ================================================================================================== stage objects
struct afe_segmenter {
    int device;
    cudaStream_t st = nullptr;
    SegState seg;
    DevBuf<float> window;
    bool window_set = false;
};
struct afe_delta {
    int device, dim, window_limit, L;
    cudaStream_t st = nullptr;
    DevBuf<float> out;
};
struct afe_normalizer {
    int device;
    cudaStream_t st = nullptr;
    NormState ns;
};

extern "C" {

int afe_segmenter_create(int W, int S, int window_limit, int deltasize, int cuda_device, afe_segmenter **out)
{
    *out = nullptr;
    return guarded([&] {
        if (afe_device_count() <= cuda_device) throw Error("no usable CUDA device (the product has no CPU fallback)");
        std::unique_ptr<afe_segmenter> s(new afe_segmenter());
        s->device = cuda_device;
        DeviceGuard g(cuda_device);
        AFE_CUDA(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
        s->seg.init(W, S, window_limit, deltasize);
        s->window.alloc(W);
        *out = s.release();
    });
}
void afe_segmenter_destroy(afe_segmenter *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->st);
    s->seg.release();
    cudaStreamDestroy(s->st);
    delete s;
}
int afe_segmenter_set_window(afe_segmenter *s, const float *w)
{
    return guarded([&] {
        DeviceGuard g(s->device);
        AFE_CUDA(cudaMemcpy(s->window.p, w, sizeof(float) * s->seg.W, cudaMemcpyHostToDevice));
        s->window_set = true;
    });
}
int afe_segmenter_set_input(afe_segmenter *s, const int16_t *in, float *d_out, int samples, int *wc, int *wc_nd)
{
    return guarded([&] {
        if (!s->window_set) throw Error("set_window must be called before set_input");
        DeviceGuard g(s->device);
        const int16_t *pcm = s->seg.set_input(in, samples, *wc, *wc_nd, s->st);
        if (*wc > 0) launch_segment(pcm, s->window.p, d_out, *wc_nd, s->seg.W, s->seg.S, s->seg.N2, s->st);
        s->seg.carry(s->st);
        AFE_CUDA(cudaStreamSynchronize(s->st));
    });
}
int afe_segmenter_flush(afe_segmenter *s, float *d_out, int *wc, int *wc_nd)
{
    return guarded([&] {
        DeviceGuard g(s->device);
        const int16_t *pcm = s->seg.flush(*wc, *wc_nd);
        if (*wc > 0) launch_segment(pcm, s->window.p, d_out, *wc_nd, s->seg.W, s->seg.S, s->seg.N2, s->st);
        AFE_CUDA(cudaStreamSynchronize(s->st));
    });
}
int afe_segmenter_remaining_samples(const afe_segmenter *s) { return s->seg.remaining; }
int afe_segmenter_samples(const afe_segmenter *s) { return s->seg.samples; }
int afe_segmenter_is_flushed(const afe_segmenter *s) { return s->seg.flushed; }
int afe_segmenter_was_flushed(const afe_segmenter *s) { return s->seg.last_calc_flushed; }

int afe_delta_create(int dim, int window_limit, int delta_size, int cuda_device, afe_delta **out)
{
    *out = nullptr;
    return guarded([&] {
        if (afe_device_count() <= cuda_device) throw Error("no usable CUDA device (the product has no CPU fallback)");
        std::unique_ptr<afe_delta> d(new afe_delta());
        d->device = cuda_device; d->dim = dim; d->window_limit = window_limit; d->L = delta_size;
        DeviceGuard g(cuda_device);
        AFE_CUDA(cudaStreamCreateWithFlags(&d->st, cudaStreamNonBlocking));
        d->out.alloc((size_t)dim * window_limit);
        *out = d.release();
    });
}
void afe_delta_destroy(afe_delta *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    cudaStreamSynchronize(d->st);
    cudaStreamDestroy(d->st);
    delete d;
}
int afe_delta_apply(afe_delta *d, const float *d_data, int window_count)
{
    return guarded([&] {
        if (window_count > d->window_limit) throw Error("Window count too high");
        DeviceGuard g(d->device);
        launch_delta(d_data, d->out.p, window_count, d->dim, d->L, d->st);
        AFE_CUDA(cudaStreamSynchronize(d->st));
    });
}
float *afe_delta_output(afe_delta *d) { return d->out.p; }

int afe_normalizer_create(int norm_type, int dim, int cuda_device, afe_normalizer **out)
{
    *out = nullptr;
    return guarded([&] {
        if (afe_device_count() <= cuda_device) throw Error("no usable CUDA device (the product has no CPU fallback)");
        std::unique_ptr<afe_normalizer> n(new afe_normalizer());
        n->device = cuda_device;
        DeviceGuard g(cuda_device);
        AFE_CUDA(cudaStreamCreateWithFlags(&n->st, cudaStreamNonBlocking));
        n->ns.init(norm_type, dim);
        *out = n.release();
    });
}
void afe_normalizer_destroy(afe_normalizer *n)
{
    if (!n) return;
    cudaSetDevice(n->device);
    cudaStreamSynchronize(n->st);
    cudaStreamDestroy(n->st);
    delete n;
}
int afe_normalizer_normalize(afe_normalizer *n, float *d_data, int offset, int window_count, int use_last_stats)
{
    return guarded([&] {
        DeviceGuard g(n->device);
        n->ns.normalize(d_data + offset, window_count, use_last_stats != 0, n->st);
        AFE_CUDA(cudaStreamSynchronize(n->st));
    });
}

int afe_device_malloc(int dev, size_t bytes, void **p)
{
    return guarded([&] { DeviceGuard g(dev); AFE_CUDA(cudaMalloc(p, bytes ? bytes : 1)); });
}
int afe_device_free(int dev, void *p)
{
    return guarded([&] { DeviceGuard g(dev); AFE_CUDA(cudaFree(p)); });
}
int afe_memcpy_h2d(int dev, void *d, const void *h, size_t bytes)
{
    return guarded([&] { DeviceGuard g(dev); AFE_CUDA(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice)); });
}
int afe_memcpy_d2h(int dev, void *h, const void *d, size_t bytes)
{
    return guarded([&] { DeviceGuard g(dev); AFE_CUDA(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost)); });
}

} // extern "C"
