// Accelerator variants of the stage objects, over the C ABI of libafe_cuda.so (include/afe_cuda.h).
//   MfccCuda       replaces MfccOpenCL       (mfccopencl.h:21-73)   — same 15 constructor arguments + CUDA device index
//   SegmenterCuda  replaces SegmenterOpenCL  (segmenteropencl.h:7-50)
//   DeltaCuda      replaces DeltaOpenCL      (deltaopencl.h:5-23)
//   NormalizerCuda replaces NormalizerOpenCL (normalizeropencl.h:5-29)
// `cl_mem` arguments become device pointers; (context, queue, device) become one `int cuda_device`.
// Errors surface as std::runtime_error carrying the library's message (the reference's strings where it has one).
#pragma once
#include <stdexcept>
#include <string>

#include "afe_cuda.h"
// Inside the reference's tree include its own "mfccbase.h" first and define AFE_USE_REFERENCE_HEADERS: the classes below
// then derive from the reference's ParamBase / MfccBase (same names, same protected members) instead of the mirror.
#ifndef AFE_USE_REFERENCE_HEADERS
#include "afe_stage_api.hpp"
#endif

namespace afe_detail {
inline void check(int rc)
{
    if (rc != 0) throw std::runtime_error(afe_last_error());
}
} // namespace afe_detail

class MfccCuda : public MfccBase {
public:
    MfccCuda(int input_buffer_size, int window_size, int shift, int num_banks, float sample_rate, float low_freq,
             float high_freq, int ceps_len, bool want_c0, float lift_coef,
             Normalizer::norm_t norm = Normalizer::NORM_NONE, dyn_t dyn = DYN_NONE, int delta_l1 = 1, int delta_l2 = 1,
             bool norm_after_dyn = true, int cuda_device = 0)
        : MfccBase(input_buffer_size, window_size, shift, num_banks, sample_rate, low_freq, high_freq, ceps_len, want_c0,
                   lift_coef, norm, dyn, delta_l1, delta_l2, norm_after_dyn),
          m_handle(nullptr)
    {
        afe_params p;
        p.input_buffer_size = input_buffer_size; p.window_size = window_size; p.shift = shift; p.num_banks = num_banks;
        p.sample_rate = sample_rate; p.low_freq = low_freq; p.high_freq = high_freq; p.ceps_len = ceps_len;
        p.want_c0 = want_c0 ? 1 : 0; p.lift_coef = lift_coef; p.norm = (int)norm; p.dyn = (int)dyn;
        p.delta_l1 = delta_l1; p.delta_l2 = delta_l2; p.norm_after_dyn = norm_after_dyn ? 1 : 0;
        afe_detail::check(afe_mfcc_create(&p, cuda_device, &m_handle));
    }
    ~MfccCuda() override { afe_mfcc_destroy(m_handle); }
    MfccCuda(const MfccCuda &) = delete;
    MfccCuda &operator=(const MfccCuda &) = delete;

    // ParamBase::set_alpha is not virtual and the driver calls it through a ParamBase* (parambase.h:25, ASR_OCL.cpp:241,276):
    // like MfccCpu::filter (mfcccpu.cpp:194) this object reads the protected m_alpha when apply() runs.
    void set_window(const float *window) override { afe_detail::check(afe_mfcc_set_window(m_handle, window)); }
    int set_input(const short *data, int samples) override
    {
        int frames = 0;
        afe_detail::check(afe_mfcc_set_input(m_handle, data, samples, &frames));
        return frames;
    }
    int flush() override
    {
        int frames = 0;
        afe_detail::check(afe_mfcc_flush(m_handle, &frames));
        m_last_block = true;
        return frames;
    }
    void apply() override
    {
        afe_detail::check(afe_mfcc_set_alpha(m_handle, m_alpha));
        afe_detail::check(afe_mfcc_apply(m_handle));
    }
    void get_output_data(float *data_out, int window_count) override
    {
        afe_detail::check(afe_mfcc_get_output(m_handle, data_out, window_count));
    }
    // extensions: the reference object cannot be reused after flush() (m_last_block is never cleared, Q3)
    void reset() { afe_detail::check(afe_mfcc_reset(m_handle)); m_last_block = false; }
    void fix_flush_statics(bool on) { afe_detail::check(afe_mfcc_set_option(m_handle, AFE_OPT_FIX_FLUSH_STATICS, on)); }
    // per-frame pre-emphasis before the window (the reference has none: 0 is its behaviour)
    void set_preemphasis(float coefficient) { afe_detail::check(afe_mfcc_set_preemphasis(m_handle, coefficient)); }
    bool uses_fused_kernel() const { return afe_mfcc_uses_fused_kernel(m_handle) != 0; }

private:
    afe_mfcc *m_handle;
};

class SegmenterCuda {
public:
    SegmenterCuda() : m_h(nullptr) {}
    void init(int window_size, int shift, int window_limit, int deltasize, int cuda_device = 0)
    {
        afe_detail::check(afe_segmenter_create(window_size, shift, window_limit, deltasize, cuda_device, &m_h));
        m_window_size = window_size; m_shift = shift;
    }
    void cleanup() { afe_segmenter_destroy(m_h); m_h = nullptr; }
    void set_window(const float *window) { afe_detail::check(afe_segmenter_set_window(m_h, window)); }
    void set_preemphasis(float coefficient) { afe_detail::check(afe_segmenter_set_preemphasis(m_h, coefficient)); }
    // d_data_out: DEVICE float[window_count_no_delta][ceil2(window_size)]
    void set_input(const short *data_in, float *d_data_out, int samples, int &window_count, int &window_count_no_delta)
    {
        afe_detail::check(afe_segmenter_set_input(m_h, data_in, d_data_out, samples, &window_count, &window_count_no_delta));
    }
    void flush(float *d_data_out, int &window_count, int &window_count_no_delta)
    {
        afe_detail::check(afe_segmenter_flush(m_h, d_data_out, &window_count, &window_count_no_delta));
    }
    int get_remaining_samples() const { return afe_segmenter_remaining_samples(m_h); }
    int get_samples() const { return afe_segmenter_samples(m_h); }
    bool is_flushed() const { return afe_segmenter_is_flushed(m_h) != 0; }
    bool was_flushed() const { return afe_segmenter_was_flushed(m_h) != 0; }
    int estimated_window_count(int samples) const { return afe_estimated_window_count(samples, m_window_size, m_shift); }

private:
    afe_segmenter *m_h;
    int m_window_size = 0, m_shift = 0;
};

class DeltaCuda {
public:
    DeltaCuda() : m_h(nullptr) {}
    void init(int dim, int window_limit, int delta_size, int cuda_device = 0)
    {
        afe_detail::check(afe_delta_create(dim, window_limit, delta_size, cuda_device, &m_h));
    }
    void cleanup() { afe_delta_destroy(m_h); m_h = nullptr; }
    void apply(const float *d_data, int window_count) { afe_detail::check(afe_delta_apply(m_h, d_data, window_count)); }
    float *get_output_buffer() { return afe_delta_output(m_h); } // DEVICE pointer

private:
    afe_delta *m_h;
};

class NormalizerCuda {
public:
    NormalizerCuda() : m_h(nullptr) {}
    void init(Normalizer::norm_t norm_type, int dim, int cuda_device = 0)
    {
        afe_detail::check(afe_normalizer_create((int)norm_type, dim, cuda_device, &m_h));
    }
    void cleanup() { afe_normalizer_destroy(m_h); m_h = nullptr; }
    void normalize(float *d_data, int offset, int window_count, bool use_last_stats = false)
    {
        afe_detail::check(afe_normalizer_normalize(m_h, d_data, offset, window_count, use_last_stats ? 1 : 0));
    }
    // Corpus-level CMVN (the exchange lives here): reset -> accumulate (every block / shard) -> allreduce (every rank) ->
    // finalize -> apply (every block). The three kernels of NormalizerOpenCL::normalize (normalizeropencl.cpp:123-158) as
    // separate verbs, the NCCL all-reduce of the record (sum | sumsq | count | min | max) between the first two.
    void reset() { afe_detail::check(afe_normalizer_reset(m_h)); }
    void accumulate(const float *d_data, int offset, int window_count)
    {
        afe_detail::check(afe_normalizer_accumulate(m_h, d_data, offset, window_count));
    }
    void allreduce(void *nccl_comm) { afe_detail::check(afe_normalizer_allreduce(m_h, nccl_comm)); }
    void finalize() { afe_detail::check(afe_normalizer_finalize(m_h)); }
    void apply(float *d_data, int offset, int window_count)
    {
        afe_detail::check(afe_normalizer_apply(m_h, d_data, offset, window_count));
    }
    int stats_len() const { return afe_normalizer_stats_len(m_h); }
    void get_stats(double *h_stats) { afe_detail::check(afe_normalizer_get_stats(m_h, h_stats)); }
    void set_stats(const double *h_stats) { afe_detail::check(afe_normalizer_set_stats(m_h, h_stats)); }

private:
    afe_normalizer *m_h;
};
