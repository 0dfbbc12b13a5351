// Host-side mirror of the reference's stage API for the MFCC path, so that code written against
// `ParamBase* p = new MfccOpenCL(...)` (ASR_OCL.cpp:141) compiles against `new MfccCuda(...)` unchanged.
//
// Same names, same argument order and meaning, same error behaviour (std::runtime_error with the reference's
// messages) as parambase.h:6-33, mfccbase.h:6-35 and normalizer.h:5 of mankeyboy/ASR-FeatExt-OpenCL. This file is a
// fresh declaration of that interface (nothing is copied); the arithmetic lives behind the C ABI in libafe_cuda.so.
#pragma once
#include <cmath>

namespace Normalizer {
// normalizer.h:5 — values are part of the interface (they travel through afe_params.norm)
enum norm_t { NORM_NONE = 0, NORM_CMN = 1, NORM_CVN = 2, NORM_MINMAX = 3 };
} // namespace Normalizer

// Abstract streaming feature extractor (parambase.h:6-33).
class ParamBase {
public:
    enum dyn_t { DYN_NONE = 0, DYN_DELTA = 1, DYN_ACC = 2 }; // parambase.h:9

    ParamBase(int input_buffer_size, int window_size, int shift, Normalizer::norm_t norm, dyn_t dyn)
        : m_window_size(window_size), m_shift(shift), m_alpha(1), m_norm(norm), m_dyn(dyn), m_last_block(false)
    {
        // parambase.cpp:12-13: capacity in whole frames, then the sample count that fills exactly those frames
        m_input_window_limit = estimated_window_count(input_buffer_size);
        m_input_buffer_size = m_input_window_limit * m_shift + m_window_size - m_shift;
    }
    virtual ~ParamBase() {}

    int get_input_buffer_size() const { return m_input_buffer_size; }
    // parambase.cpp:16-19 (evaluated in float like the reference)
    int estimated_window_count(int samples) const
    {
        return (int)std::floor(float(samples - (m_window_size - m_shift)) / m_shift);
    }
    void set_alpha(float alpha) { m_alpha = alpha; } // NOT virtual (parambase.h:25): objects read m_alpha in apply()

    virtual void set_window(const float *window) = 0;
    virtual int set_input(const short *data, int samples) = 0;
    virtual int flush() = 0;
    virtual void apply() = 0;
    virtual int get_output_data_width() const = 0;
    virtual void get_output_data(float *data_out, int window_count) = 0;

protected:
    int m_input_buffer_size, m_input_window_limit, m_window_size, m_shift;
    float m_alpha;
    Normalizer::norm_t m_norm;
    dyn_t m_dyn;
    bool m_last_block;
};

// MFCC parameter set + output-width rule (mfccbase.h:21-35, mfccbase.cpp:3-43).
class MfccBase : public ParamBase {
public:
    MfccBase(int input_buffer_size, int window_size, int shift, int num_banks, float sample_rate, float low_freq,
             float high_freq, int ceps_len, bool want_c0, float lift_coef,
             Normalizer::norm_t norm = Normalizer::NORM_NONE, dyn_t dyn = DYN_NONE, int delta_l1 = 1, int delta_l2 = 1,
             bool norm_after_dyn = true)
        : ParamBase(input_buffer_size, window_size, shift, norm, dyn), m_num_banks(num_banks), m_ceps_len(ceps_len),
          m_dct_len(want_c0 ? ceps_len + 1 : ceps_len), m_delta_l1(dyn != DYN_NONE ? delta_l1 : 0),
          m_delta_l2(dyn == DYN_ACC ? delta_l2 : 0), m_sample_rate(sample_rate), m_low_freq(low_freq),
          m_high_freq(high_freq), m_lift_coef(lift_coef), m_want_c0(want_c0), m_norm_after_dyn(norm_after_dyn)
    {
    }
    virtual ~MfccBase() {}

    // mfccbase.cpp:33-43: cols = ceps_len > 0 ? ceps_len + c0 : num_banks; x1 / x2 / x3 by dyn
    int get_output_data_width() const override
    {
        const int cols = m_ceps_len > 0 ? m_dct_len : m_num_banks;
        return cols * (m_dyn == DYN_ACC ? 3 : m_dyn == DYN_DELTA ? 2 : 1);
    }

protected:
    int m_num_banks, m_ceps_len, m_dct_len, m_delta_l1, m_delta_l2;
    float m_sample_rate, m_low_freq, m_high_freq, m_lift_coef;
    bool m_want_c0, m_norm_after_dyn;
};
