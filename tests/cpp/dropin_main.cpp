// Drop-in proof with the REFERENCE's own base classes: this driver is compiled against /root/reference/parambase.h and
// mfccbase.h and linked with the reference's parambase.cpp / mfccbase.cpp objects (oracle/Makefile target `dropin`,
// output oracle/_ref/dropin_ref); the only accelerator class is MfccCuda over libafe_cuda.so. It drives the object
// through a ParamBase* like process_files_worker does (ASR_OCL.cpp:141,152,227-301), sweeping VTLN alpha with the
// NON-virtual ParamBase::set_alpha, and dumps the rows per alpha as raw float32 for tests/test_gpu_driver.py.
//   dropin_ref in.s16 out.f32 sample_limit alpha_min alpha_max alpha_step
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <stdexcept>
#include <vector>

#include "mfccbase.h"   // the reference's header
#include "mfcccuda.hpp" // derives from the reference's MfccBase (AFE_USE_REFERENCE_HEADERS)

int main(int argc, char **argv)
{
    if (argc != 7) { fprintf(stderr, "usage: dropin_ref in.s16 out.f32 sample_limit alpha_min alpha_max alpha_step\n"); return 2; }
    try {
        FILE *f = fopen(argv[1], "rb");
        if (!f) throw std::runtime_error("can't open input");
        std::vector<short> pcm;
        short buf[4096];
        size_t n;
        while ((n = fread(buf, 2, 4096, f)) > 0) pcm.insert(pcm.end(), buf, buf + n);
        fclose(f);
        const int sample_limit = atoi(argv[3]);
        const float amin = (float)atof(argv[4]), amax = (float)atof(argv[5]), astep = (float)atof(argv[6]);
        std::vector<float> alphas;
        for (int i = 0; amin + i * astep <= amax + 1e-6f; i++) { alphas.push_back(amin + i * astep); if (astep <= 0) break; }
        std::vector<float> window(400);
        for (int i = 0; i < 400; i++) window[i] = (float)(0.56f - 0.46f * cos((2.0f * M_PI * i) / 400)) / 32768.f; // ASR_OCL.cpp:149-152
        ParamBase *param = new MfccCuda(sample_limit, 400, 160, 23, 16000.f, 64.f, 8000.f, 12, true, 22.f, Normalizer::NORM_CMN,
                                        ParamBase::DYN_ACC, 3, 3, true, 0);
        param->set_window(window.data());
        const int width = param->get_output_data_width(), limit = param->get_input_buffer_size();
        std::vector<std::vector<float>> rows(alphas.size());
        std::vector<float> tmp((size_t)width * (size_t)(param->estimated_window_count(limit) + 64));
        auto emit = [&](int wc) {
            for (size_t k = 0; k < alphas.size(); k++) {
                param->set_alpha(alphas[k]);
                param->apply();
                param->get_output_data(tmp.data(), wc);
                rows[k].insert(rows[k].end(), tmp.begin(), tmp.begin() + (size_t)wc * width);
            }
        };
        for (size_t pos = 0; pos < pcm.size();) {
            const int m = (int)std::min<size_t>(pcm.size() - pos, (size_t)limit);
            const int wc = param->set_input(pcm.data() + pos, m);
            if (wc > 0) emit(wc);
            pos += m;
        }
        const int wc = param->flush();
        if (wc > 0) emit(wc);
        delete param;
        FILE *o = fopen(argv[2], "wb");
        if (!o) throw std::runtime_error("can't create output");
        for (auto &r : rows) fwrite(r.data(), 4, r.size(), o);
        fclose(o);
        printf("dropin_ref: %zu alphas x %zu rows x %d\n", alphas.size(), rows[0].size() / width, width);
    } catch (const std::exception &e) {
        fprintf(stderr, "Exception caught %s\n", e.what());
        return 1;
    }
    return 0;
}
