// afe_stream_bench — throughput of the drop-in object through the reference's own stage API: T host threads, each driving
// ONE ParamBase* (a MfccCuda) over its share of F in-memory files with the reference driver's per-file sequence
// (ASR_OCL.cpp:227-301): set_input -> set_alpha -> apply -> get_output_data, flush -> set_alpha -> apply -> get_output_data.
// The object is made reusable between files with MfccCuda::reset() (the reference never clears m_last_block, Q3).
// Host buffers in, host buffers out: every H2D / D2H copy is inside the timed region. Prints one JSON line.
//   afe_stream_bench [--files 256] [--threads 8] [--seconds 10] [--banks 40] [--norm 1] [--dyn 2] [--dev 0] [--repeat 3]
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "mfcccuda.hpp"

int main(int argc, char **argv)
{
    int files = 256, threads = 8, seconds = 10, banks = 40, norm = 1, dyn = 2, dev = 0, repeat = 3, sample_limit = 10000000, profile = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string a = argv[i];
        const int v = atoi(argv[i + 1]);
        if (a == "--files") files = v; else if (a == "--threads") threads = v; else if (a == "--seconds") seconds = v;
        else if (a == "--banks") banks = v; else if (a == "--norm") norm = v; else if (a == "--dyn") dyn = v;
        else if (a == "--dev") dev = v; else if (a == "--repeat") repeat = v;
        else if (a == "--sample-limit") sample_limit = v; else if (a == "--profile") profile = v;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    try {
        const int sr = 16000, W = 400, S = 160, n = sr * seconds;
        // synthetic files: noise + one sinusoid each (SURVEY §8d shape; the exact generator does not matter for timing)
        std::vector<std::vector<short>> pcm((size_t)files, std::vector<short>((size_t)n));
        unsigned lcg = 12345u;
        for (int f = 0; f < files; f++) {
            const double w = 2.0 * M_PI * (100.0 + 3700.0 * ((f * 37) % 101) / 101.0) / sr;
            for (int i = 0; i < n; i++) {
                lcg = lcg * 1664525u + 1013904223u;
                const double noise = ((int)(lcg >> 16) - 32768) / 32768.0 * 5000.0;
                pcm[f][i] = (short)std::lrint(noise + 8000.0 * std::sin(w * i));
            }
        }
        std::vector<float> window((size_t)W);
        afe_make_window(window.data(), W);
        std::vector<std::unique_ptr<MfccCuda>> objs;
        for (int t = 0; t < threads; t++) {
            // input_buffer_size = the reference driver's default sample_limit (10 M, ASR_OCL.cpp:563): a file is one block
            objs.emplace_back(new MfccCuda(sample_limit, W, S, banks, (float)sr, 64.f, sr / 2.f, 12, true, 22.f, (Normalizer::norm_t)norm,
                                           (ParamBase::dyn_t)dyn, 3, 3, true, dev));
            objs.back()->set_window(window.data());
        }
        const int width = objs[0]->get_output_data_width();
        const int T = objs[0]->estimated_window_count(n);
        std::atomic<long long> rows_total{0};
        std::atomic<int> failed{0};
        double checksum = 0;
        double prof[6] = {0, 0, 0, 0, 0, 0}; // thread 0: set_input, apply, get_output, flush, apply+get (flush block), reset
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        auto work = [&](int t, bool count) {
            try {
                ParamBase *param = objs[t].get();           // the reference's interface; set_alpha is its non-virtual member
                const int limit = param->get_input_buffer_size();
                std::vector<float> out((size_t)(T + 64) * width);
                long long rows = 0;
                double sum = 0;
                for (int f = t; f < files; f += threads) {
                    for (int pos = 0; pos < n; pos += limit) { // ASR_OCL.cpp:227-267
                        const auto t0 = now();
                        const int wc = param->set_input(pcm[f].data() + pos, std::min(limit, n - pos));
                        const auto t1 = now();
                        param->set_alpha(1.0f);
                        param->apply();
                        const auto t2 = now();
                        if (wc > 0) {
                            param->get_output_data(out.data(), wc);
                            rows += wc;
                            sum += out[(size_t)(wc - 1) * width];
                        }
                        const auto t3 = now();
                        if (profile && t == 0 && count) { prof[0] += secs(t0, t1); prof[1] += secs(t1, t2); prof[2] += secs(t2, t3); }
                    }
                    const auto t3 = now();
                    const int wc = param->flush();
                    const auto t4 = now();
                    if (wc > 0) {
                        param->set_alpha(1.0f);
                        param->apply();
                        param->get_output_data(out.data(), wc);
                        rows += wc;
                    }
                    const auto t5 = now();
                    objs[t]->reset();
                    const auto t6 = now();
                    if (profile && t == 0 && count) { prof[3] += secs(t3, t4); prof[4] += secs(t4, t5); prof[5] += secs(t5, t6); }
                }
                if (count) { rows_total += rows; if (t == 0) checksum = sum; }
            } catch (const std::exception &e) {
                if (failed.fetch_add(1) == 0) fprintf(stderr, "Exception caught %s\n", e.what());
            }
        };
        auto run_all = [&](bool count) {
            std::vector<std::thread> th;
            for (int t = 0; t < threads; t++) th.emplace_back(work, t, count);
            for (auto &x : th) x.join();
        };
        run_all(false); // warm-up
        double best = 1e30;
        for (int r = 0; r < repeat; r++) {
            rows_total = 0;
            const auto t0 = std::chrono::steady_clock::now();
            run_all(true);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt < best) best = dt;
            if (failed.load()) return 1;
        }
        if (profile) {
            const double nf = (double)((files + threads - 1) / threads) * repeat;
            fprintf(stderr, "thread 0, us per file: set_input %.1f | apply %.1f | get_output_data %.1f | flush %.1f | flush-block apply+get %.1f | reset %.1f\n",
                    prof[0] / nf * 1e6, prof[1] / nf * 1e6, prof[2] / nf * 1e6, prof[3] / nf * 1e6, prof[4] / nf * 1e6, prof[5] / nf * 1e6);
        }
        printf("{\"frames_per_s\": %.1f, \"files\": %d, \"threads\": %d, \"seconds_per_file\": %d, \"frames_per_file\": %d, "
               "\"ms_per_file_per_thread\": %.4f, \"rows\": %lld, \"width\": %d, \"uses_fused_kernel\": %s, \"checksum\": %.6f}\n",
               (double)rows_total.load() / best, files, threads, seconds, T, best / files * threads * 1e3, rows_total.load(), width,
               objs[0]->uses_fused_kernel() ? "true" : "false", checksum);
    } catch (const std::exception &e) {
        fprintf(stderr, "Exception caught %s\n", e.what());
        return 1;
    }
    return 0;
}
