// afe_extract — host driver around ParamBase*, the CUDA counterpart of the reference's process_files_worker
// (ASR_OCL.cpp:109-338): window synthesis (:149-152), block loop set_input -> set_alpha -> apply -> get_output_data
// (:227-267), flush (:268-301) and the text layout "| time | v | v | ... |" (:252-260) including its timestamp
// arithmetic (window and shift in ms divided by the sample rate in Hz, quirk Q6). The flag names follow the option
// block the reference documents but never enabled (ASR_OCL.cpp:570-612). libsndfile is replaced by a 44-byte RIFF /
// 1024-byte NIST header skip: 16-bit mono PCM only.
//
//   afe_extract [options] in.wav out.txt [in2.wav out2.txt ...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mfcccuda.hpp"

struct Config { // SConfig, ASR_OCL.cpp:81-100; defaults ASR_OCL.cpp:560
    float alpha = 1.f, window_ms = 25.f, shift_ms = 10.f;
    int num_banks = 15, ceps_len = 12, norm_type = 2, dyn_type = 0, delta_l1 = 3, delta_l2 = 3;
    float sample_rate = 16000.f, low_freq = 64.f, high_freq = 0.f, lift_coef = 22.f;
    bool want_c0 = true, norm_after_dyn = true, text_output = true, fix_flush = false;
    int sample_limit = 10000000, device = 0;
};

static std::vector<short> read_pcm(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Can't open \"" + path + "\"");
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> raw((size_t)size);
    if (fread(raw.data(), 1, raw.size(), f) != raw.size()) { fclose(f); throw std::runtime_error("Error while loading \"" + path + "\""); }
    fclose(f);
    const size_t skip = (size >= 4 && !memcmp(raw.data(), "NIST", 4)) ? 1024 : 44;
    if (raw.size() <= skip) throw std::runtime_error("Error while loading \"" + path + "\"");
    std::vector<short> pcm((raw.size() - skip) / 2);
    memcpy(pcm.data(), raw.data() + skip, pcm.size() * 2);
    return pcm;
}

static void write_rows(FILE *fout, const Config &cfg, const float *data, int rows, int width, long first_frame)
{
    if (!cfg.text_output) { fwrite(data, sizeof(float), (size_t)rows * width, fout); return; }
    const long double step = cfg.shift_ms / cfg.sample_rate, t0 = 0.5f * cfg.window_ms / cfg.sample_rate; // ASR_OCL.cpp:225-226
    for (int r = 0; r < rows; r++) {
        fprintf(fout, "| %f |", (double)(t0 + (first_frame + r) * step));
        for (int i = 0; i < width; i++) fprintf(fout, " %f |", data[(size_t)width * r + i]);
        fprintf(fout, "\n");
    }
}

static void process_file(ParamBase *param, const Config &cfg, const std::string &in, const std::string &out)
{
    std::vector<short> pcm = read_pcm(in);
    FILE *fout = fopen(out.c_str(), cfg.text_output ? "w" : "wb");
    if (!fout) throw std::runtime_error("Can't create output file: " + out);
    const int limit = param->get_input_buffer_size(), width = param->get_output_data_width();
    // a middle block can return more rows than estimated_window_count(limit) (carry-over): size for the object's frame
    // capacity, input_window_limit + 2 + 3*(l1+l2) (mfcccpu.cpp:95-103)
    std::vector<float> rows((size_t)width * (size_t)(std::max(1, param->estimated_window_count(limit)) + 2 + 3 * (cfg.delta_l1 + cfg.delta_l2)));
    long total = 0;
    size_t pos = 0;
    while (pos < pcm.size()) { // ASR_OCL.cpp:227-267
        const int n = (int)std::min<size_t>(pcm.size() - pos, (size_t)limit);
        const int wc = param->set_input(pcm.data() + pos, n);
        param->set_alpha(cfg.alpha);
        param->apply();
        if (wc > 0) {
            param->get_output_data(rows.data(), wc);
            write_rows(fout, cfg, rows.data(), wc, width, total);
            total += wc;
        }
        pos += n;
    }
    const int wc = param->flush(); // ASR_OCL.cpp:268-301
    if (wc > 0) {
        param->set_alpha(cfg.alpha);
        param->apply();
        param->get_output_data(rows.data(), wc);
        write_rows(fout, cfg, rows.data(), wc, width, total);
        total += wc;
    }
    fclose(fout);
    fprintf(stderr, "%s: %ld frames x %d -> %s\n", in.c_str(), total, width, out.c_str());
}

int main(int argc, char **argv)
{
    Config cfg;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--window-size") cfg.window_ms = (float)atof(val());
        else if (a == "--shift") cfg.shift_ms = (float)atof(val());
        else if (a == "--banks") cfg.num_banks = atoi(val());
        else if (a == "--ceps") cfg.ceps_len = atoi(val());
        else if (a == "--sample-rate") cfg.sample_rate = (float)atof(val());
        else if (a == "--low-freq") cfg.low_freq = (float)atof(val());
        else if (a == "--high-freq") cfg.high_freq = (float)atof(val());
        else if (a == "--lift-coef") cfg.lift_coef = (float)atof(val());
        else if (a == "--c0") cfg.want_c0 = atoi(val()) != 0;
        else if (a == "--alpha") cfg.alpha = (float)atof(val());
        else if (a == "--norm") cfg.norm_type = atoi(val());
        else if (a == "--dyn") cfg.dyn_type = atoi(val());
        else if (a == "--l1") cfg.delta_l1 = atoi(val());
        else if (a == "--l2") cfg.delta_l2 = atoi(val());
        else if (a == "--norm-after-dyn") cfg.norm_after_dyn = atoi(val()) != 0;
        else if (a == "--sample-limit") cfg.sample_limit = atoi(val());
        else if (a == "--text-output") cfg.text_output = atoi(val()) != 0;
        else if (a == "--fix-flush-statics") cfg.fix_flush = atoi(val()) != 0;
        else if (a == "--dev") cfg.device = atoi(val());
        else if (a.rfind("--", 0) == 0) { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
        else files.push_back(a);
    }
    if (files.empty() || files.size() % 2) { fprintf(stderr, "usage: afe_extract [options] in.wav out.txt [...]\n"); return 2; }
    if (cfg.high_freq <= 0) cfg.high_freq = cfg.sample_rate / 2; // ASR_OCL.cpp:359-360
    try {
        const long window_size = (long)cfg.sample_rate * cfg.window_ms * 1e-3, shift = (long)cfg.sample_rate * cfg.shift_ms * 1e-3; // :115-116
        std::vector<float> window((size_t)window_size);
        afe_make_window(window.data(), (int)window_size);
        for (size_t i = 0; i < files.size(); i += 2) {
            // one object per file: the reference never clears m_last_block (Q3)
            std::unique_ptr<MfccCuda> param(new MfccCuda(cfg.sample_limit, (int)window_size, (int)shift, cfg.num_banks, cfg.sample_rate,
                                                          cfg.low_freq, cfg.high_freq, cfg.ceps_len, cfg.want_c0, cfg.lift_coef,
                                                          (Normalizer::norm_t)cfg.norm_type, (ParamBase::dyn_t)cfg.dyn_type, cfg.delta_l1,
                                                          cfg.delta_l2, cfg.norm_after_dyn, cfg.device));
            param->set_window(window.data());
            if (cfg.fix_flush) param->fix_flush_statics(true);
            process_file(param.get(), cfg, files[i], files[i + 1]);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "Exception caught %s\n", e.what()); // ASR_OCL.cpp:326-331
        return 1;
    }
    return 0;
}
