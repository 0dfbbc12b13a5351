// afe_extract — host driver around ParamBase*, the CUDA counterpart of the reference's process_files_worker
// (ASR_OCL.cpp:109-338): window synthesis (:149-152), block loop set_input -> set_alpha -> apply -> get_output_data
// (:227-267), flush (:268-301) and the text layout "| time | v | v | ... |" (:252-260) including its timestamp
// arithmetic (window and shift in ms divided by the sample rate in Hz, quirk Q6). The flag names follow the option
// block the reference documents but never enabled (ASR_OCL.cpp:570-612). libsndfile is replaced by a 44-byte RIFF /
// 1024-byte NIST header skip: 16-bit mono PCM only.
//
// Beyond the reference: --batch packs ALL files into one shard and runs them through the fused batch
// kernel in a single call (afe_batch_*: pinned, chunked H2D / kernel / D2H pipeline), --scp reads "in out" pairs from a
// list, --htk writes HTK parameter files (the reference's binary branch writes nothing, ASR_OCL.cpp:212,315-318).
//
//   afe_extract [options] [--scp list] in.wav out.txt [in2.wav out2.txt ...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mfcccuda.hpp"

struct Config { // SConfig, ASR_OCL.cpp:81-100; defaults ASR_OCL.cpp:560
    float alpha = 1.f, alpha_max = 1.f, alpha_step = 0.f, preemphasis = 0.f, window_ms = 25.f, shift_ms = 10.f;
    int num_banks = 15, ceps_len = 12, norm_type = 2, dyn_type = 0, delta_l1 = 3, delta_l2 = 3;
    float sample_rate = 16000.f, low_freq = 64.f, high_freq = 0.f, lift_coef = 22.f;
    bool want_c0 = true, norm_after_dyn = true, text_output = true, fix_flush = false, batch = false, htk = false;
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

// HTK parameter file: big-endian header {nSamples int32, sampPeriod int32 [100 ns], sampSize int16 [bytes per row],
// parmKind int16} followed by big-endian float rows. Kind: MFCC (6) or FBANK (7) with the qualifiers _0 (c0 appended,
// last column like the reference), _D, _A, _Z (mean removed). The header is written first with nSamples = 0 and patched
// when the file is complete.
static void put_be32(unsigned char *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
static void write_htk_header(FILE *f, const Config &cfg, long rows, int width)
{
    unsigned kind = cfg.ceps_len > 0 ? 6 : 7;
    if (cfg.ceps_len > 0 && cfg.want_c0) kind |= 0x2000;
    if (cfg.dyn_type >= 1) kind |= 0x100;
    if (cfg.dyn_type >= 2) kind |= 0x200;
    if (cfg.norm_type != 0) kind |= 0x800;
    unsigned char h[12];
    put_be32(h, (uint32_t)rows);
    put_be32(h + 4, (uint32_t)(cfg.shift_ms * 1e4f + 0.5f));
    h[8] = (unsigned char)((4 * width) >> 8); h[9] = (unsigned char)(4 * width);
    h[10] = (unsigned char)(kind >> 8); h[11] = (unsigned char)kind;
    fseek(f, 0, SEEK_SET);
    fwrite(h, 1, 12, f);
    fseek(f, 0, SEEK_END);
}

static void write_rows(FILE *fout, const Config &cfg, const float *data, int rows, int width, long first_frame)
{
    if (cfg.htk) {
        std::vector<unsigned char> be((size_t)rows * width * 4);
        for (size_t i = 0; i < (size_t)rows * width; i++) {
            uint32_t v;
            memcpy(&v, data + i, 4);
            put_be32(be.data() + 4 * i, v);
        }
        fwrite(be.data(), 1, be.size(), fout);
        return;
    }
    if (!cfg.text_output) { fwrite(data, sizeof(float), (size_t)rows * width, fout); return; }
    const long double step = cfg.shift_ms / cfg.sample_rate, t0 = 0.5f * cfg.window_ms / cfg.sample_rate; // ASR_OCL.cpp:225-226
    for (int r = 0; r < rows; r++) {
        fprintf(fout, "| %f |", (double)(t0 + (first_frame + r) * step));
        for (int i = 0; i < width; i++) fprintf(fout, " %f |", data[(size_t)width * r + i]);
        fprintf(fout, "\n");
    }
}

// VTLN sweep of the reference driver (ASR_OCL.cpp:198-218): alpha = min, min + step, ... <= max; one output file per alpha,
// named stem + to_string(alpha) + extension when there is more than one.
static std::vector<float> alpha_list(const Config &cfg)
{
    std::vector<float> a;
    if (cfg.alpha_step <= 0.f || cfg.alpha_max - cfg.alpha < cfg.alpha_step) { a.push_back(cfg.alpha); return a; }
    int i = 0;
    for (float v = cfg.alpha; v <= cfg.alpha_max; v = cfg.alpha + i * cfg.alpha_step) { a.push_back(v); i++; }
    return a;
}

static void process_file(ParamBase *param, const Config &cfg, const std::string &in, const std::string &out)
{
    std::vector<short> pcm = read_pcm(in);
    const std::vector<float> alphas = alpha_list(cfg);
    std::vector<FILE *> fouts;
    for (float alpha : alphas) {
        std::string name = out;
        if (alphas.size() > 1 && out.size() > 4) name = out.substr(0, out.size() - 4) + std::to_string(alpha) + out.substr(out.size() - 4);
        FILE *f = fopen(name.c_str(), cfg.text_output && !cfg.htk ? "w" : "wb");
        if (!f) throw std::runtime_error("Can't create output file: " + name);
        fouts.push_back(f);
    }
    const int limit = param->get_input_buffer_size(), width = param->get_output_data_width();
    if (cfg.htk) for (FILE *f : fouts) write_htk_header(f, cfg, 0, width);
    // a middle block can return more rows than estimated_window_count(limit) (carry-over): size for the object's frame
    // capacity, input_window_limit + 2 + 3*(l1+l2) (mfcccpu.cpp:95-103)
    std::vector<float> rows((size_t)width * (size_t)(std::max(1, param->estimated_window_count(limit)) + 2 + 3 * (cfg.delta_l1 + cfg.delta_l2)));
    long total = 0;
    size_t pos = 0;
    // one set_input, then set_alpha -> apply -> get_output_data per alpha (ASR_OCL.cpp:234-243): set_alpha is the
    // NON-virtual ParamBase member, called through the base pointer exactly as the reference driver does
    auto emit = [&](int wc) {
        for (size_t k = 0; k < alphas.size(); k++) {
            param->set_alpha(alphas[k]);
            param->apply();
            param->get_output_data(rows.data(), wc);
            write_rows(fouts[k], cfg, rows.data(), wc, width, total);
        }
        total += wc;
    };
    while (pos < pcm.size()) { // ASR_OCL.cpp:227-267
        const int n = (int)std::min<size_t>(pcm.size() - pos, (size_t)limit);
        const int wc = param->set_input(pcm.data() + pos, n);
        if (wc > 0) emit(wc);
        else { param->set_alpha(alphas[0]); param->apply(); }
        pos += n;
    }
    const int wc = param->flush(); // ASR_OCL.cpp:268-301
    if (wc > 0) emit(wc);
    for (FILE *f : fouts) {
        if (cfg.htk) write_htk_header(f, cfg, total, width);
        fclose(f);
    }
    fprintf(stderr, "%s: %ld frames x %d x %zu alpha(s) -> %s\n", in.c_str(), total, width, alphas.size(), out.c_str());
}

// --batch: every file is one utterance of ONE shard (each processed like a file that fits a single block of the reference
// loop, ASR_OCL.cpp:227-301 with sample_limit >= N): one plan, one afe_batch_run_host call for all of them.
static void check(int rc) { if (rc) throw std::runtime_error(afe_last_error()); }
static void process_batch(const Config &cfg, const std::vector<std::string> &files, int window_size, int shift, const float *window)
{
    const size_t n = files.size() / 2;
    std::vector<int64_t> off(n), len(n), frame_off(n + 1);
    std::vector<short> pcm;
    for (size_t i = 0; i < n; i++) {
        const std::vector<short> x = read_pcm(files[2 * i]);
        off[i] = (int64_t)pcm.size();
        len[i] = (int64_t)x.size();
        pcm.insert(pcm.end(), x.begin(), x.end());
        pcm.resize((pcm.size() + 7) / 8 * 8, 0); // 16-byte aligned utterance starts: TMA staging
    }
    pcm.resize(pcm.size() + 16, 0);
    afe_params p{};
    p.input_buffer_size = cfg.sample_limit; p.window_size = window_size; p.shift = shift; p.num_banks = cfg.num_banks;
    p.sample_rate = cfg.sample_rate; p.low_freq = cfg.low_freq; p.high_freq = cfg.high_freq; p.ceps_len = cfg.ceps_len;
    p.want_c0 = cfg.want_c0; p.lift_coef = cfg.lift_coef; p.norm = cfg.norm_type; p.dyn = cfg.dyn_type;
    p.delta_l1 = cfg.delta_l1; p.delta_l2 = cfg.delta_l2; p.norm_after_dyn = cfg.norm_after_dyn;
    afe_batch *b = nullptr;
    check(afe_batch_create(&p, cfg.device, &b));
    try {
        check(afe_batch_set_window(b, window));
        check(afe_batch_set_alpha(b, cfg.alpha));
        if (cfg.preemphasis > 0.f) check(afe_batch_set_preemphasis(b, cfg.preemphasis));
        check(afe_batch_set_options(b, AFE_STATS_REFERENCE_BLOCK, cfg.fix_flush ? 0 : AFE_BATCH_Q1_EXACT));
        int64_t total = 0;
        check(afe_batch_plan(b, off.data(), len.data(), (int)n, &total));
        check(afe_batch_frame_offsets(b, frame_off.data()));
        const int width = afe_output_width(&p);
        std::vector<float> rows((size_t)total * width);
        check(afe_batch_run_host(b, pcm.data(), rows.data()));
        for (size_t i = 0; i < n; i++) {
            const std::string &out = files[2 * i + 1];
            FILE *fout = fopen(out.c_str(), cfg.text_output && !cfg.htk ? "w" : "wb");
            if (!fout) throw std::runtime_error("Can't create output file: " + out);
            const long T = (long)(frame_off[i + 1] - frame_off[i]);
            if (cfg.htk) write_htk_header(fout, cfg, T, width);
            write_rows(fout, cfg, rows.data() + (size_t)frame_off[i] * width, (int)T, width, 0);
            fclose(fout);
            fprintf(stderr, "%s: %ld frames x %d -> %s\n", files[2 * i].c_str(), T, width, out.c_str());
        }
        fprintf(stderr, "batch: %zu files, %lld frames, %d tile(s), %d kernel launch(es), %s\n", n, (long long)total,
                afe_batch_num_tiles(b), afe_batch_kernel_launches(b), afe_batch_kernel_name(b));
    } catch (...) { afe_batch_destroy(b); throw; }
    afe_batch_destroy(b);
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
        else if (a == "--alpha") { cfg.alpha = (float)atof(val()); if (cfg.alpha_max < cfg.alpha) cfg.alpha_max = cfg.alpha; }
        else if (a == "--alpha-max") cfg.alpha_max = (float)atof(val());
        else if (a == "--alpha-step") cfg.alpha_step = (float)atof(val());
        else if (a == "--preemphasis") cfg.preemphasis = (float)atof(val());
        else if (a == "--norm") cfg.norm_type = atoi(val());
        else if (a == "--dyn") cfg.dyn_type = atoi(val());
        else if (a == "--l1") cfg.delta_l1 = atoi(val());
        else if (a == "--l2") cfg.delta_l2 = atoi(val());
        else if (a == "--norm-after-dyn") cfg.norm_after_dyn = atoi(val()) != 0;
        else if (a == "--sample-limit") cfg.sample_limit = atoi(val());
        else if (a == "--text-output") cfg.text_output = atoi(val()) != 0;
        else if (a == "--fix-flush-statics") cfg.fix_flush = atoi(val()) != 0;
        else if (a == "--dev") cfg.device = atoi(val());
        else if (a == "--batch") cfg.batch = atoi(val()) != 0;
        else if (a == "--htk") cfg.htk = atoi(val()) != 0;
        else if (a == "--scp") {
            std::ifstream scp(val());
            if (!scp) { fprintf(stderr, "can't open list %s\n", argv[i]); return 2; }
            std::string line, in, out;
            while (std::getline(scp, line)) {
                std::istringstream ls(line);
                if (ls >> in >> out) { files.push_back(in); files.push_back(out); }
            }
        }
        else if (a.rfind("--", 0) == 0) { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
        else files.push_back(a);
    }
    if (files.empty() || files.size() % 2) { fprintf(stderr, "usage: afe_extract [options] [--scp list] in.wav out.txt [...]\n"); return 2; }
    if (cfg.high_freq <= 0) cfg.high_freq = cfg.sample_rate / 2; // ASR_OCL.cpp:359-360
    try {
        const long window_size = (long)cfg.sample_rate * cfg.window_ms * 1e-3, shift = (long)cfg.sample_rate * cfg.shift_ms * 1e-3; // :115-116
        std::vector<float> window((size_t)window_size);
        afe_make_window(window.data(), (int)window_size);
        if (cfg.batch) { process_batch(cfg, files, (int)window_size, (int)shift, window.data()); return 0; }
        for (size_t i = 0; i < files.size(); i += 2) {
            // one object per file: the reference never clears m_last_block (Q3)
            std::unique_ptr<MfccCuda> param(new MfccCuda(cfg.sample_limit, (int)window_size, (int)shift, cfg.num_banks, cfg.sample_rate,
                                                          cfg.low_freq, cfg.high_freq, cfg.ceps_len, cfg.want_c0, cfg.lift_coef,
                                                          (Normalizer::norm_t)cfg.norm_type, (ParamBase::dyn_t)cfg.dyn_type, cfg.delta_l1,
                                                          cfg.delta_l2, cfg.norm_after_dyn, cfg.device));
            param->set_window(window.data());
            if (cfg.fix_flush) param->fix_flush_statics(true);
            if (cfg.preemphasis > 0.f) param->set_preemphasis(cfg.preemphasis);
            process_file(param.get(), cfg, files[i], files[i + 1]);
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "Exception caught %s\n", e.what()); // ASR_OCL.cpp:326-331
        return 1;
    }
    return 0;
}
