// Compile-only check that MfccCuda drops into the REFERENCE's own stage API: this translation unit includes the reference's
// headers (parambase.h / mfccbase.h, non-virtual ParamBase::set_alpha, parambase.h:25) instead of the mirror in
// asr-featext-opencl_b200/host/afe_stage_api.hpp and drives the object through a ParamBase* exactly as
// ASR_OCL.cpp:141,152,234-243,268-278 does. Built by tests/test_host_logic.py when /root/reference is present:
//   g++ -std=c++14 -fsyntax-only -DAFE_USE_REFERENCE_HEADERS -I/root/reference -Iinclude -Iasr-featext-opencl_b200/host
#include "mfccbase.h"   // the reference's
#include "mfcccuda.hpp" // ours, deriving from the reference's MfccBase

int drive(const short *pcm, int samples, const float *window, float *out)
{
    ParamBase *param = new MfccCuda(10000000, 400, 160, 23, 16000.f, 64.f, 8000.f, 12, true, 22.f, Normalizer::NORM_CMN,
                                    ParamBase::DYN_ACC, 3, 3, true, /*cuda_device=*/0);
    param->set_window(window);
    int rows = 0;
    const int wc = param->set_input(pcm, samples);
    for (float alpha = 0.9f; alpha <= 1.1f; alpha += 0.1f) { // VTLN sweep through the base pointer
        param->set_alpha(alpha);
        param->apply();
        param->get_output_data(out, wc);
    }
    rows += wc;
    const int tail = param->flush();
    if (tail > 0) {
        param->set_alpha(1.0f);
        param->apply();
        param->get_output_data(out, tail);
    }
    rows += tail;
    const int width = param->get_output_data_width();
    delete param;
    return rows * width;
}
