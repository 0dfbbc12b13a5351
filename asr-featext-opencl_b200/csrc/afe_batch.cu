// Batch extractor (afe_batch_*): planning, launch of the fused kernel K1, statistics finalize K2 and the normalise
// pass K3, plus the corpus-CMVN exchange hooks of the Normalizer subsystem.
#include <algorithm>
#include <atomic>
#include <memory>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "afe_internal.h"
#include "afe_fused.cuh"
#include "afe_fused_ws.cuh"
#include "afe_nccl.h"

namespace afe {

// ------------------------------------------------------------------------------------------------ FFT tables
void FftTables::build(int n2)
{
    release();
    N2 = n2;
    const int M = n2 / 2, R = M / 16;
    std::vector<float2> a((size_t)R * 16), p((size_t)M / 2);
    for (int l = 0; l < R; l++)
        for (int k1 = 0; k1 < 16; k1++) {
            const double ang = -2.0 * M_PI * (double)(l * k1) / (double)M;
            a[(size_t)l * 16 + k1] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int k = 0; k < M / 2; k++) {
        const double ang = -2.0 * M_PI * (double)k / (double)n2;
        p[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    AFE_CUDA(cudaMalloc(&d_tw_a, a.size() * sizeof(float2)));
    AFE_CUDA(cudaMalloc(&d_tw_p, p.size() * sizeof(float2)));
    AFE_CUDA(cudaMemcpy(d_tw_a, a.data(), a.size() * sizeof(float2), cudaMemcpyHostToDevice));
    AFE_CUDA(cudaMemcpy(d_tw_p, p.data(), p.size() * sizeof(float2), cudaMemcpyHostToDevice));
}
void FftTables::release()
{
    if (d_tw_a) cudaFree(d_tw_a);
    if (d_tw_p) cudaFree(d_tw_p);
    d_tw_a = d_tw_p = nullptr;
}

void MelTables::release()
{
    if (d_edges) cudaFree(d_edges);
    if (d_pairs) cudaFree(d_pairs);
    if (d_dct) cudaFree(d_dct);
    if (d_window) cudaFree(d_window);
    if (d_window2) cudaFree(d_window2);
    d_edges = nullptr; d_pairs = nullptr; d_dct = nullptr; d_window = nullptr; d_window2 = nullptr;
    alpha_built = -1.f;
}

void upload_mel_tables(const Derived &d, float alpha, MelTables &t, cudaStream_t st)
{
    std::vector<int> edges; std::vector<float> filters, pairs, dct;
    build_filters(d, alpha, edges, filters);
    for (int i = 0; i + 1 < (int)edges.size(); i++)
        if (edges[i + 1] < edges[i]) throw Error("mel filter edges are not monotonic");
    if (edges.front() < 0 || edges.back() > d.M) throw Error("mel filterbank exceeds the Nyquist bin (check low_freq/high_freq)");
    build_mel_pairs(d, edges, filters, pairs);
    if (!t.d_edges) AFE_CUDA(cudaMalloc(&t.d_edges, sizeof(int) * (d.nb + 2)));
    if (!t.d_pairs) AFE_CUDA(cudaMalloc(&t.d_pairs, sizeof(float) * 2 * d.bins));
    // pageable source + stream-ordered copy: cudaMemcpyAsync from pageable memory stages synchronously, safe with locals
    AFE_CUDA(cudaMemcpyAsync(t.d_edges, edges.data(), sizeof(int) * (d.nb + 2), cudaMemcpyHostToDevice, st));
    AFE_CUDA(cudaMemcpyAsync(t.d_pairs, pairs.data(), sizeof(float) * 2 * d.bins, cudaMemcpyHostToDevice, st));
    if (d.C > 0 && !t.d_dct) {
        build_dct(d, dct);
        AFE_CUDA(cudaMalloc(&t.d_dct, sizeof(float) * dct.size()));
        AFE_CUDA(cudaMemcpyAsync(t.d_dct, dct.data(), sizeof(float) * dct.size(), cudaMemcpyHostToDevice, st));
    }
    AFE_CUDA(cudaStreamSynchronize(st));
    t.alpha_built = alpha;
}

void upload_window(const Derived &d, const float *window, MelTables &t, cudaStream_t st)
{
    if (!t.d_window) AFE_CUDA(cudaMalloc(&t.d_window, sizeof(float) * d.W));
    if (!t.d_window2) AFE_CUDA(cudaMalloc(&t.d_window2, sizeof(float2) * d.M));
    std::vector<float> w2((size_t)d.N2, 0.f);
    memcpy(w2.data(), window, sizeof(float) * d.W);
    AFE_CUDA(cudaMemcpyAsync(t.d_window, window, sizeof(float) * d.W, cudaMemcpyHostToDevice, st));
    AFE_CUDA(cudaMemcpyAsync(t.d_window2, w2.data(), sizeof(float) * d.N2, cudaMemcpyHostToDevice, st));
    AFE_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------------------------------------ K2 / K3
// stats record per group: [sum[w], sumsq[w], count, min[w], max[w]]  (4w+1 doubles)
__global__ void k_reduce_partials(const double *__restrict__ partials, const int *__restrict__ tile_begin,
                                  const double *__restrict__ counts, int width, double *__restrict__ stats)
{
    const int g = blockIdx.x, c = threadIdx.x;
    if (c >= width) return;
    double s0 = 0.0, s1 = 0.0, lo = (double)FLT_MAX, hi = -(double)FLT_MAX;
#pragma unroll 8
    for (int t = tile_begin[g]; t < tile_begin[g + 1]; t++) { // fixed order: deterministic (loads are hoisted, adds stay ordered)
        const double *p = partials + ((long long)t * width + c) * 4;
        s0 += p[0]; s1 += p[1];
        lo = fmin(lo, p[2]); hi = fmax(hi, p[3]);
    }
    double *o = stats + (long long)g * (4 * width + 1);
    o[c] = s0; o[width + c] = s1; o[2 * width + 1 + c] = lo; o[3 * width + 1 + c] = hi;
    if (c == 0) o[2 * width] = counts[g];
}

// Corpus scope has ONE group over all tiles: two deterministic levels instead of one serial loop.
// level 1: block j sums tiles j, j+gridDim.x, ... -> scratch[j][width][4]; level 2 = k_reduce_partials over the scratch.
__global__ void k_reduce_partials_level1(const double *__restrict__ partials, int n_tiles, int width,
                                         double *__restrict__ scratch)
{
    const int c = threadIdx.x;
    if (c >= width) return;
    double s0 = 0.0, s1 = 0.0, lo = (double)FLT_MAX, hi = -(double)FLT_MAX;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const double *p = partials + ((long long)t * width + c) * 4;
        s0 += p[0]; s1 += p[1];
        lo = fmin(lo, p[2]); hi = fmax(hi, p[3]);
    }
    double *o = scratch + ((long long)blockIdx.x * width + c) * 4;
    o[0] = s0; o[1] = s1; o[2] = lo; o[3] = hi;
}

// mean / scale per group and column (normalizercpu.cpp:31-66). With norm_after_dyn == 0 the reference normalises the
// statics before the deltas are taken, which equals scaling delta columns by the static column's scale (mean 0).
__global__ void k_finalize_stats(const double *__restrict__ stats, int width, int cols, int norm_type, int norm_after_dyn,
                                 float *__restrict__ mean, float *__restrict__ scale)
{
    const int g = blockIdx.x, c = threadIdx.x;
    if (c >= width) return;
    const double *o = stats + (long long)g * (4 * width + 1);
    const double n = o[2 * width];
    const int src = norm_after_dyn ? c : (c % cols);
    const double s = o[src], s2 = o[width + src];
    const float mn = (float)o[2 * width + 1 + src], mx = (float)o[3 * width + 1 + src];
    float m = (float)(s / n), sc = 1.f;
    if (norm_type == AFE_NORM_CVN) sc = (float)sqrt((n - 1.0) / (s2 - s * (s / n)));
    else if (norm_type == AFE_NORM_MINMAX) sc = 1.f / fmaxf(fabsf(mn - m), fabsf(mx - m));
    if (!norm_after_dyn && c >= cols) m = 0.f;
    mean[(long long)g * width + c] = m;
    scale[(long long)g * width + c] = sc;
}

// K3: in-place (x - mean) [* scale] over one tile's rows: 128-bit accesses on the 16-byte aligned body of the tile's
// contiguous region, columns tracked incrementally (no division in the loop). HBM bound: 8 B per float.
__global__ void k_normalize_tiles(float *__restrict__ out, const Tile *__restrict__ tiles, int width, int norm_type,
                                  const float *__restrict__ mean, const float *__restrict__ scale)
{
    extern __shared__ float s_ms[]; // mean[width] | scale[width]
    const Tile tl = tiles[blockIdx.x];
    float *base = out + (tl.out_row0 + tl.t0) * (long long)width;
    const int n = tl.nout * width;
    for (int i = threadIdx.x; i < width; i += blockDim.x) {
        s_ms[i] = mean[(long long)tl.group * width + i];
        s_ms[width + i] = norm_type == AFE_NORM_CMN ? 1.f : scale[(long long)tl.group * width + i];
    }
    __syncthreads();
    const float *m = s_ms, *sc = s_ms + width;
    const bool cmn = norm_type == AFE_NORM_CMN;
    const int head = min(n, (int)(((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15) >> 2));
    const int n4 = (n - head) >> 2, tail0 = head + 4 * n4;
    if ((int)threadIdx.x < head) {
        const int i = threadIdx.x;
        const float v = base[i] - m[i % width];
        base[i] = cmn ? v : v * sc[i % width];
    }
    if ((int)threadIdx.x < n - tail0) {
        const int i = tail0 + threadIdx.x, c = i % width;
        const float v = base[i] - m[c];
        base[i] = cmn ? v : v * sc[c];
    }
    float4 *p4 = reinterpret_cast<float4 *>(base + head);
    int c = (head + 4 * (int)threadIdx.x) % width;
    const int cstep = (4 * (int)blockDim.x) % width;
    for (int j = threadIdx.x; j < n4; j += blockDim.x) {
        float4 v = p4[j];
        int c1 = c + 1; if (c1 >= width) c1 -= width;
        int c2 = c1 + 1; if (c2 >= width) c2 -= width;
        int c3 = c2 + 1; if (c3 >= width) c3 -= width;
        v.x -= m[c]; v.y -= m[c1]; v.z -= m[c2]; v.w -= m[c3];
        if (!cmn) { v.x *= sc[c]; v.y *= sc[c1]; v.z *= sc[c2]; v.w *= sc[c3]; }
        p4[j] = v;
        c += cstep; if (c >= width) c -= width;
    }
}

constexpr int kCorpusBlocks = 296; // 2 per SM
static std::atomic<int> g_launches{0};
int kernel_launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add(n); }

} // namespace afe

using namespace afe;

// ================================================================================================== afe_batch
struct afe_batch {
    Derived d;
    int device;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    FftTables fft;
    MelTables mel;
    float alpha = 1.f;
    bool window_set = false;
    int scope = AFE_STATS_REFERENCE_BLOCK, flags = 0;
    // plan
    int n_utts = 0, n_tiles = 0, n_groups = 0;
    bool aligned = false;
    std::vector<int64_t> sample_off, sample_len, frame_off;
    std::vector<int> h_tile_begin;          // [n_utts+1] first tile of every utterance
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k;
    int64_t pcm_extent = 0;
    Tile *d_tiles = nullptr;
    int *d_tile_begin = nullptr;
    double *d_counts = nullptr, *d_partials = nullptr, *d_stats = nullptr, *d_scratch = nullptr;
    int *d_counters = nullptr;
    int *d_scratch_begin = nullptr;
    float *d_mean = nullptr, *d_scale = nullptr;
    int tc_max = 0, nout_max = 0;
    const int warps = 8;      // warps per CTA of the fused kernel
    int max_tiles_per_utt = 0;
    // In-kernel normalisation lets ONE tile normalise its whole utterance: right for short utterances, a serial
    // bottleneck for a long stream (config 5: 720 tiles) -> those batches take the K2 + K3 kernels.
    bool fuse_norm() const { return !(flags & AFE_BATCH_UNFUSED_NORM) && max_tiles_per_utt <= 8; }
    FusedSmem L{};
    WsSmem Lws{};             // layout of the warp-specialised kernel (k_fused_ws)
    int sm_count = 0;
    // k_fused_ws (AFE_BATCH_WS_KERNEL, opt-in) covers the reference's default regression (static + delta + delta-delta,
    // l1 = l2 = 3); everything else takes k_fused_mfcc. Decided at plan time: the two kernels tile differently.
    bool ws_eligible() const { return (flags & AFE_BATCH_WS_KERNEL) && d.width == 3 * d.cols && d.l1 == 3 && d.l2 == 3; }
    bool ws_planned = false;
    bool use_ws() const { return ws_planned; }
    MelConst mc;
    float mc_alpha = -1.f;
    int last_launches = 0;
    // host staging for run_host
    int16_t *d_pcm_stage = nullptr; float *d_out_stage = nullptr;
    size_t pcm_stage_bytes = 0, out_stage_bytes = 0;

    explicit afe_batch(const afe_params &p, int dev) : d(p), device(dev) {}
    void free_plan()
    {
        if (d_tiles) cudaFree(d_tiles);
        if (d_tile_begin) cudaFree(d_tile_begin);
        if (d_counts) cudaFree(d_counts);
        if (d_partials) cudaFree(d_partials);
        if (d_stats) cudaFree(d_stats);
        if (d_scratch) cudaFree(d_scratch);
        if (d_counters) cudaFree(d_counters);
        d_counters = nullptr;
        if (d_scratch_begin) cudaFree(d_scratch_begin);
        d_scratch = nullptr; d_scratch_begin = nullptr;
        if (d_mean) cudaFree(d_mean);
        if (d_scale) cudaFree(d_scale);
        d_tiles = nullptr; d_tile_begin = nullptr; d_counts = d_partials = d_stats = nullptr; d_mean = d_scale = nullptr;
    }
};

static void check_fused_support(const Derived &d)
{
    if (d.N2 != 512 && d.N2 != 256)
        throw Error("fused batch path supports 256/512-point FFTs (window_size 129..512); use the streaming object");
    if (d.S % 2) throw Error("fused batch path needs an even shift");
    if (d.dct_len > 16) throw Error("fused batch path supports ceps_len + c0 <= 16");
    if (d.width > 128) throw Error("fused batch path supports output width <= 128");
    if (d.nb > kMaxBanks) throw Error("fused batch path supports num_banks <= 64");
}

template <int N2> static FusedSmem layout_for(const afe_batch *b)
{
    const Derived &d = b->d;
    return fused_smem_layout<N2>(b->warps, d.S, d.cols, b->tc_max, b->nout_max, d.l2, d.width / d.cols);
}

// Mel weights + DCT matrix as a by-value kernel parameter (constant bank). Per filter b: bins [edges[b], edges[b+2]) with
// weights filters[b%2][bin] — the per-filter form of the reference's two-running-sums sweep (mfcccpu.cpp:192-220).
static void build_mel_const(const Derived &d, float alpha, MelConst &mc)
{
    std::vector<int> edges; std::vector<float> filters, dct;
    build_filters(d, alpha, edges, filters);
    for (int i = 0; i + 1 < (int)edges.size(); i++)
        if (edges[i + 1] < edges[i]) throw Error("mel filter edges are not monotonic");
    if (edges.front() < 0 || edges.back() > d.M) throw Error("mel filterbank exceeds the Nyquist bin (check low_freq/high_freq)");
    memset(&mc, 0, sizeof mc);
    int off4 = 0;
    const float scale = 0.5f / (float)d.N2; // the kernel stores |2X|; scaling by a power of two commutes with rounding
    for (int w = 0; w < 8; w++) {           // warp class w owns the filters w, w + 8, ...: their lists are contiguous
        mc.wstart[w] = (short)off4;
        for (int b = w; b < d.nb; b += 8) {
            const int j0 = edges[b], j1 = edges[b + 2], s4 = j0 & ~3, n8 = std::max(1, (j1 - s4 + 7) / 8);
            if (off4 + 2 * n8 > kMaxWl4) throw Error("mel weight list exceeds the kernel-parameter budget");
            if (s4 + 8 * n8 > kMagStride + 16) throw Error("mel filter too wide for the fused kernel");
            mc.desc[b] = (s4 / 4) | (n8 << 16);
            float *wt = reinterpret_cast<float *>(mc.wl4 + off4);
            for (int j = j0; j < j1; j++) wt[j - s4] = filters[(size_t)(b % 2) * d.N2 + j] * scale;
            off4 += 2 * n8;
        }
    }
    if (d.C > 0) {
        build_dct(d, dct);
        for (int k = 0; k < d.nb; k++)
            for (int j = 0; j < d.dct_len; j++) reinterpret_cast<float *>(mc.dct4[k])[j] = dct[(size_t)k * d.dct_len + j];
    }
}

static FusedArgs make_fused_args(afe_batch *b, const int16_t *d_pcm, float *d_out, bool want_stats, int t0, bool fuse_norm)
{
    const Derived &d = b->d;
    FusedArgs a{};
    a.pcm = d_pcm; a.out = d_out; a.tiles = b->d_tiles; a.tile_base = t0;
    a.window2 = b->mel.d_window2; a.tw_a = b->fft.d_tw_a; a.tw_p = b->fft.d_tw_p;
    a.partials = want_stats ? b->d_partials : nullptr;
    a.counters = (want_stats && fuse_norm) ? b->d_counters : nullptr;
    a.norm_type = d.p.norm; a.norm_after_dyn = d.p.norm_after_dyn;
    a.W = d.W; a.S = d.S; a.nb = d.nb; a.dct_len = d.C > 0 ? d.dct_len : 0; a.cols = d.cols; a.width = d.width;
    a.l1 = d.l1; a.l2 = d.l2; a.nstreams = d.width / d.cols;
    a.q1 = (b->flags & AFE_BATCH_Q1_EXACT) && d.D > 0 ? 1 : 0;
    a.use_tma = (b->aligned && !(b->flags & AFE_BATCH_NO_TMA) && (reinterpret_cast<uintptr_t>(d_pcm) & 15) == 0) ? 1 : 0;
    if (!want_stats) a.stats_rows_mode = 0;
    else if (!d.p.norm_after_dyn) a.stats_rows_mode = 2;
    else a.stats_rows_mode = b->scope == AFE_STATS_REFERENCE_BLOCK ? 1 : 2;
    a.stats_kind = !want_stats ? 0 : (d.p.norm == AFE_NORM_CMN ? 1 : d.p.norm == AFE_NORM_CVN ? 2 : 3);
    a.tc_max = b->tc_max;
    { const char *dbg = getenv("AFE_DEBUG_SKIP"); a.debug_skip = dbg ? atoi(dbg) : 0; }
    float den1 = 0, den2 = 0;
    for (int l = 1; l <= d.l1; l++) den1 += l * l;   // float accumulation like deltacpu.cpp:26
    for (int l = 1; l <= d.l2; l++) den2 += l * l;
    a.rden1 = den1 > 0 ? 1.f / (2 * den1) : 0.f;
    a.rden2 = den2 > 0 ? 1.f / (2 * den2) : 0.f;
    return a;
}

template <int N2, int NZ, int KF>
static void launch_fused_ws(afe_batch *b, const FusedArgs &a, int t0, int t1)
{
    auto kern = k_fused_ws<N2, NZ, false, KF>; // AFE_BATCH_FAST_MATH applies to k_fused_mfcc only (halves the build time)
    AFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, b->Lws.total));
    const int ntiles = t1 - t0;
    const int grid = std::min(ntiles, b->sm_count); // persistent: one CTA per SM walks tiles blockIdx.x, +grid, ...
    kern<<<grid, kWsThreads, b->Lws.total, b->stream>>>(a, b->Lws, b->mc, ntiles);
    AFE_CUDA(cudaGetLastError());
    count_launch();
    b->last_launches++;
}

// cluster: 0 = plain launch (ticket scheme or no fused normalisation); 1, 2, ... = clustered launch, one cluster per
// utterance of `cluster` tiles, normalisation inside the cluster (FusedArgs::cluster_norm)
template <int N2, int NZ, int WARPS, int KF>
static void launch_fused(afe_batch *b, const int16_t *d_pcm, float *d_out, bool want_stats, int t0, int t1, bool fuse_norm,
                         int cluster = 0)
{
    FusedArgs a = make_fused_args(b, d_pcm, d_out, want_stats, t0, fuse_norm);
    if (b->use_ws()) { launch_fused_ws<N2, NZ, KF>(b, a, t0, t1); return; }
    const bool fast = (b->flags & AFE_BATCH_FAST_MATH) != 0;
    auto kern = fast ? k_fused_mfcc<N2, NZ, true, WARPS, KF> : k_fused_mfcc<N2, NZ, false, WARPS, KF>;
    AFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, b->L.total));
    if (cluster > 0) {
        a.cluster_norm = 1;
        a.counters = nullptr;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(t1 - t0); cfg.blockDim = dim3(32 * WARPS); cfg.dynamicSmemBytes = b->L.total; cfg.stream = b->stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        AFE_CUDA(cudaLaunchKernelEx(&cfg, kern, a, b->L, b->mc));
    } else
        kern<<<t1 - t0, 32 * WARPS, b->L.total, b->stream>>>(a, b->L, b->mc);
    AFE_CUDA(cudaGetLastError());
    count_launch();
    b->last_launches++;
}

static void dispatch_fused(afe_batch *b, const int16_t *d_pcm, float *d_out, bool want_stats, int t0, int t1, bool fuse_norm,
                           int cluster)
{
    const int R = b->d.M / 16;
    const bool pruned = b->d.W <= 26 * R; // window tail is zero from n1 = 13 on (400/512 and 200/256 both qualify)
    const int kf = (b->d.nb + 7) / 8; // filters per warp (8 warps): instantiated for 3, 5 and 8
    const int key = (b->d.N2 == 512 ? 0 : 6) + (pruned ? 0 : 3) + (kf <= 3 ? 0 : kf <= 5 ? 1 : 2);
    switch (key) {
    case 0: launch_fused<512, 13, 8, 3>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 1: launch_fused<512, 13, 8, 5>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 2: launch_fused<512, 13, 8, 8>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 3: launch_fused<512, 16, 8, 3>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 4: launch_fused<512, 16, 8, 5>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 5: launch_fused<512, 16, 8, 8>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 6: launch_fused<256, 13, 8, 3>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 7: launch_fused<256, 13, 8, 5>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 8: launch_fused<256, 13, 8, 8>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 9: launch_fused<256, 16, 8, 3>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    case 10: launch_fused<256, 16, 8, 5>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    default: launch_fused<256, 16, 8, 8>(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, cluster); break;
    }
}

// Tiles [t0, t1) (whole utterances). With fused normalisation and the default regression the utterances are taken in
// runs of equal tile count and every run of utterances with 1 to kMaxClusterTiles tiles becomes ONE clustered launch (cluster = utterance);
// a batch that would need more than kMaxClusterRuns launches (ragged lengths in random order) or other tile counts keeps
// the ticket scheme (last tile normalises through L2) in a single launch.
constexpr int kMaxClusterRuns = 8;
constexpr int kMaxClusterTiles = 4; // utterances of up to 4 tiles (~20 s) form a cluster; longer ones keep the ticket scheme
static void run_extract(afe_batch *b, const int16_t *d_pcm, float *d_out, int t0 = 0, int t1 = -1, bool fuse_norm = false)
{
    if (t1 < 0) t1 = b->n_tiles;
    if (t1 <= t0) return;
    if (!b->window_set) throw Error("set_window must be called before running");
    if (!b->d_tiles) throw Error("afe_batch_plan must be called before running");
    if (b->mc_alpha != b->alpha) { build_mel_const(b->d, b->alpha, b->mc); b->mc_alpha = b->alpha; }
    const Derived &d = b->d;
    const bool want_stats = d.p.norm != AFE_NORM_NONE;
    const bool cluster_ok = fuse_norm && want_stats && !b->use_ws() && !(b->flags & AFE_BATCH_NO_CLUSTER) &&
                            d.width == 3 * d.cols && d.l1 == 3 && d.l2 == 3;
    if (cluster_ok) {
        struct Run { int t0, t1, cluster; };
        std::vector<Run> runs;
        const int u0 = (int)(std::lower_bound(b->h_tile_begin.begin(), b->h_tile_begin.end(), t0) - b->h_tile_begin.begin());
        for (int u = u0; u < b->n_utts && b->h_tile_begin[u] < t1 && (int)runs.size() <= kMaxClusterRuns; u++) {
            const int nt = b->h_tile_begin[u + 1] - b->h_tile_begin[u], cl = nt <= kMaxClusterTiles ? nt : 0;
            if (!runs.empty() && runs.back().cluster == cl) runs.back().t1 = b->h_tile_begin[u + 1];
            else runs.push_back({b->h_tile_begin[u], b->h_tile_begin[u + 1], cl});
        }
        if ((int)runs.size() <= kMaxClusterRuns && !runs.empty() && runs.front().t0 == t0 && runs.back().t1 == t1) {
            for (const Run &r : runs) dispatch_fused(b, d_pcm, d_out, want_stats, r.t0, r.t1, fuse_norm, r.cluster);
            return;
        }
    }
    dispatch_fused(b, d_pcm, d_out, want_stats, t0, t1, fuse_norm, 0);
}

static void run_reduce(afe_batch *b, int g0 = 0, int g1 = -1)
{
    if (g1 < 0) g1 = b->n_groups;
    if (g1 <= g0) return;
    const int w = b->d.width;
    if (b->scope == AFE_STATS_CORPUS && b->n_tiles > 2 * kCorpusBlocks) {
        k_reduce_partials_level1<<<kCorpusBlocks, 128, 0, b->stream>>>(b->d_partials, b->n_tiles, w, b->d_scratch);
        AFE_CUDA(cudaGetLastError());
        k_reduce_partials<<<1, 128, 0, b->stream>>>(b->d_scratch, b->d_scratch_begin, b->d_counts, w, b->d_stats);
        AFE_CUDA(cudaGetLastError());
        count_launch(2); b->last_launches += 2;
        return;
    }
    k_reduce_partials<<<g1 - g0, 128, 0, b->stream>>>(b->d_partials, b->d_tile_begin + g0, b->d_counts + g0, w,
                                                       b->d_stats + (size_t)g0 * (4 * w + 1));
    AFE_CUDA(cudaGetLastError());
    count_launch(); b->last_launches++;
}

static void run_normalize(afe_batch *b, float *d_out, int t0 = 0, int t1 = -1, int g0 = 0, int g1 = -1)
{
    const Derived &d = b->d;
    if (t1 < 0) t1 = b->n_tiles;
    if (g1 < 0) g1 = b->n_groups;
    if (t1 <= t0 || g1 <= g0) return;
    const size_t w = d.width;
    k_finalize_stats<<<g1 - g0, 128, 0, b->stream>>>(b->d_stats + (size_t)g0 * (4 * w + 1), d.width, d.cols, d.p.norm,
                                                      d.p.norm_after_dyn, b->d_mean + g0 * w, b->d_scale + g0 * w);
    AFE_CUDA(cudaGetLastError());
    // Tile::group is absolute, so mean / scale keep their base
    k_normalize_tiles<<<t1 - t0, 256, 2 * d.width * sizeof(float), b->stream>>>(d_out, b->d_tiles + t0, d.width, d.p.norm,
                                                                              b->d_mean, b->d_scale);
    AFE_CUDA(cudaGetLastError());
    count_launch(2); b->last_launches += 2;
}

extern "C" {

int afe_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_error(std::string("CUDA error: ") + cudaGetErrorString(e)); return 0; }
    return n;
}

int afe_batch_create(const afe_params *p, int cuda_device, afe_batch **out)
{
    *out = nullptr;
    return guarded([&] {
        if (afe_device_count() <= cuda_device) throw Error("no usable CUDA device " + std::to_string(cuda_device) + " (the product has no CPU fallback)");
        std::unique_ptr<afe_batch> b(new afe_batch(*p, cuda_device));
        check_fused_support(b->d);
        DeviceGuard g(cuda_device);
        AFE_CUDA(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking));
        b->stream = b->own_stream;
        b->fft.build(b->d.N2);
        *out = b.release();
    });
}

void afe_batch_destroy(afe_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    b->free_plan();
    b->fft.release(); b->mel.release();
    if (b->d_pcm_stage) cudaFree(b->d_pcm_stage);
    if (b->d_out_stage) cudaFree(b->d_out_stage);
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    if (b->s_in) cudaStreamDestroy(b->s_in);
    if (b->s_out) cudaStreamDestroy(b->s_out);
    for (auto e : b->ev_in) cudaEventDestroy(e);
    for (auto e : b->ev_k) cudaEventDestroy(e);
    delete b;
}

int afe_batch_set_window(afe_batch *b, const float *window)
{
    return guarded([&] {
        DeviceGuard g(b->device);
        upload_window(b->d, window, b->mel, b->stream);
        b->window_set = true;
    });
}

int afe_batch_set_alpha(afe_batch *b, float alpha) { b->alpha = alpha; return 0; }

int afe_batch_set_options(afe_batch *b, int stats_scope, int flags)
{
    return guarded([&] {
        if (stats_scope < AFE_STATS_REFERENCE_BLOCK || stats_scope > AFE_STATS_CORPUS) throw Error("invalid stats scope");
        if (b->d_tiles && ((stats_scope == AFE_STATS_CORPUS) != (b->scope == AFE_STATS_CORPUS)))
            throw Error("set the statistics scope before afe_batch_plan");
        b->scope = stats_scope; b->flags = flags;
    });
}

int afe_batch_set_stream(afe_batch *b, void *cuda_stream)
{
    b->stream = cuda_stream ? (cudaStream_t)cuda_stream : b->own_stream;
    return 0;
}

int afe_batch_plan(afe_batch *b, const int64_t *off, const int64_t *len, int n_utts, int64_t *total_frames)
{
    return guarded([&] {
        const Derived &d = b->d;
        if (n_utts < 1) throw Error("plan: no utterances");
        DeviceGuard g(b->device);
        b->free_plan();
        // tile geometry. Generic kernel (2 CTAs per SM): the cepstra tile holds up to 512 frames (<= 27 KB of shared memory).
        // Warp-specialised kernel (1 CTA per SM): the tile takes what is left of the 227 KB, so that a whole utterance
        // (<= ~600 frames with 3 magnitude buffers, <= ~1250 with 2, at 13 columns) is ONE tile: no halo, and the
        // utterance is normalised before its rows are written.
        const char *env_tc = getenv("AFE_TILE_FRAMES");
        int tc = env_tc ? atoi(env_tc) : 512;
        b->ws_planned = false;
        if (b->ws_eligible()) {
            int t_max = 0;
            for (int u = 0; u < n_utts; u++)
                t_max = std::max<int64_t>(t_max, std::min<int64_t>(std::max<int64_t>(0, (len[u] - (d.W - d.S)) / d.S), 1 << 30));
            auto base = [&](int nbuf) { return d.N2 == 512 ? ws_smem_layout<512>(d.S, d.cols, 0, nbuf).total : ws_smem_layout<256>(d.S, d.cols, 0, nbuf).total; };
            auto cap = [&](int nbuf) { return (227 * 1024 - base(nbuf)) / (d.cols * 4) / kRoundFrames * kRoundFrames; };
            const int need = (t_max + kRoundFrames - 1) / kRoundFrames * kRoundFrames;
            const int min_tc = kRoundFrames * ((2 * d.D + 1 + kRoundFrames - 1) / kRoundFrames + 1);
            int nbuf = need <= cap(3) ? 3 : 2;
            int tcw = std::min(std::max(need, min_tc), cap(nbuf));
            if (env_tc) tcw = std::min(tcw, std::max(atoi(env_tc), min_tc));
            if (tcw >= min_tc) {
                b->ws_planned = true;
                tc = tcw;
                b->Lws = d.N2 == 512 ? ws_smem_layout<512>(d.S, d.cols, tc, nbuf) : ws_smem_layout<256>(d.S, d.cols, tc, nbuf);
            }
        }
        if (!b->ws_planned) {
            tc = std::min(tc, (6912 / d.cols) / kRoundFrames * kRoundFrames); // cepstra tile <= 27 KB: still 2 CTAs per SM
            tc = std::max(tc, kRoundFrames * ((2 * d.D + 1 + kRoundFrames - 1) / kRoundFrames + 1));
        }
        b->tc_max = tc; b->nout_max = tc - 2 * d.D;
        b->n_utts = n_utts;
        b->sample_off.assign(off, off + n_utts);
        b->sample_len.assign(len, len + n_utts);
        b->h_tile_begin.assign(n_utts + 1, 0);
        b->max_tiles_per_utt = 0;
        b->pcm_extent = 0;
        b->frame_off.assign(n_utts + 1, 0);
        std::vector<Tile> tiles;
        std::vector<int> tile_begin;
        std::vector<double> counts;
        const bool corpus = b->scope == AFE_STATS_CORPUS;
        bool aligned = d.S % 8 == 0;
        double corpus_count = 0;
        for (int u = 0; u < n_utts; u++) {
            const int64_t n = len[u];
            if (n < 0 || n > 0x7fffffff || off[u] < 0) throw Error("plan: invalid utterance offset/length");
            b->pcm_extent = std::max<int64_t>(b->pcm_extent, off[u] + n);
            if (off[u] % 2) throw Error("plan: utterance offsets must be even (32-bit PCM word loads)");
            if (off[u] % 8) aligned = false;
            // parambase.cpp:16-19 evaluates in float32, which is inexact above 2^24 samples: never let it exceed the
            // exact count (the extra frame would lie outside the utterance)
            const int T = std::min(afe_estimated_window_count((int)n, d.W, d.S), (int)std::max<int64_t>(0, (n - (d.W - d.S)) / d.S));
            if (T <= 2 * d.D || T < 1) throw Error("Can't process data, window count is too small"); // segmentercpu.cpp:65-66
            b->frame_off[u + 1] = b->frame_off[u] + T;
            // number of tiles: fewest 32-frame rounds (the halo of D frames per side is recomputed by every tile)
            int ntile = (T + b->nout_max - 1) / b->nout_max, best_cost = 1 << 30;
            const bool whole = b->ws_planned && T <= b->tc_max; // k_fused_ws: a whole utterance in one tile needs no halo
            if (whole) ntile = 1;
            for (int cand = ntile; cand <= ntile + 3 && !whole; cand++) {
                const int no = (T + cand - 1) / cand;
                int cost = 0;
                for (int t0 = 0; t0 < T; t0 += no) {
                    const int c0 = std::max(0, t0 - d.D), c1 = std::min(T, t0 + std::min(no, T - t0) + d.D);
                    cost += (c1 - c0 + kRoundFrames - 1) / kRoundFrames;
                }
                if (cost < best_cost) { best_cost = cost; ntile = cand; }
            }
            const int nout = (T + ntile - 1) / ntile;
            b->h_tile_begin[u] = (int)tiles.size();
            b->max_tiles_per_utt = std::max(b->max_tiles_per_utt, ntile);
            if (!corpus) tile_begin.push_back((int)tiles.size());
            for (int t0 = 0; t0 < T; t0 += nout) {
                Tile tl;
                tl.pcm_off = off[u]; tl.out_row0 = b->frame_off[u]; tl.T = T; tl.t0 = t0;
                tl.nout = std::min(nout, T - t0); tl.group = corpus ? 0 : u;
                tl.tile0 = b->h_tile_begin[u]; tl.ntiles = (T + nout - 1) / nout;
                tiles.push_back(tl);
            }
            const double cnt = !d.p.norm_after_dyn ? T : (b->scope == AFE_STATS_REFERENCE_BLOCK ? T - d.D : T);
            if (corpus) corpus_count += cnt; else counts.push_back(cnt);
        }
        if (corpus) { tile_begin.push_back(0); counts.push_back(corpus_count); }
        tile_begin.push_back((int)tiles.size());
        b->h_tile_begin[n_utts] = (int)tiles.size();
        b->aligned = aligned;
        b->n_tiles = (int)tiles.size();
        b->n_groups = corpus ? 1 : n_utts;
        if (b->ws_planned) {
            if (b->Lws.total > 227 * 1024) throw Error("fused kernel shared-memory budget exceeded");
        } else {
            b->L = d.N2 == 512 ? layout_for<512>(b) : layout_for<256>(b);
            if (b->L.total > 227 * 1024) throw Error("fused kernel shared-memory budget exceeded");
        }
        AFE_CUDA(cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, b->device));
        AFE_CUDA(cudaMalloc(&b->d_tiles, sizeof(Tile) * tiles.size()));
        AFE_CUDA(cudaMemcpy(b->d_tiles, tiles.data(), sizeof(Tile) * tiles.size(), cudaMemcpyHostToDevice));
        if (d.p.norm != AFE_NORM_NONE) {
            const size_t w = d.width;
            AFE_CUDA(cudaMalloc(&b->d_tile_begin, sizeof(int) * tile_begin.size()));
            AFE_CUDA(cudaMemcpy(b->d_tile_begin, tile_begin.data(), sizeof(int) * tile_begin.size(), cudaMemcpyHostToDevice));
            AFE_CUDA(cudaMalloc(&b->d_counts, sizeof(double) * counts.size()));
            AFE_CUDA(cudaMemcpy(b->d_counts, counts.data(), sizeof(double) * counts.size(), cudaMemcpyHostToDevice));
            AFE_CUDA(cudaMalloc(&b->d_partials, sizeof(double) * 4 * w * tiles.size()));
            AFE_CUDA(cudaMalloc(&b->d_stats, sizeof(double) * (4 * w + 1) * b->n_groups));
            if (corpus) {
                const int sb[2] = {0, kCorpusBlocks};
                AFE_CUDA(cudaMalloc(&b->d_scratch, sizeof(double) * 4 * w * kCorpusBlocks));
                AFE_CUDA(cudaMalloc(&b->d_scratch_begin, sizeof(sb)));
                AFE_CUDA(cudaMemcpy(b->d_scratch_begin, sb, sizeof(sb), cudaMemcpyHostToDevice));
            }
            AFE_CUDA(cudaMalloc(&b->d_counters, sizeof(int) * b->n_groups));
            AFE_CUDA(cudaMemset(b->d_counters, 0, sizeof(int) * b->n_groups));
            AFE_CUDA(cudaMalloc(&b->d_mean, sizeof(float) * w * b->n_groups));
            AFE_CUDA(cudaMalloc(&b->d_scale, sizeof(float) * w * b->n_groups));
        }
        if (total_frames) *total_frames = b->frame_off[n_utts];
    });
}

int afe_batch_frame_offsets(const afe_batch *b, int64_t *fo)
{
    memcpy(fo, b->frame_off.data(), sizeof(int64_t) * b->frame_off.size());
    return 0;
}
int afe_batch_num_tiles(const afe_batch *b) { return b->n_tiles; }
int afe_batch_kernel_launches(const afe_batch *b) { return b->last_launches; }
const char *afe_batch_kernel_name(const afe_batch *b) { return b->use_ws() ? "k_fused_ws" : "k_fused_mfcc"; }

int afe_batch_extract_device(afe_batch *b, const int16_t *d_pcm, float *d_out)
{
    return guarded([&] { DeviceGuard g(b->device); b->last_launches = 0; run_extract(b, d_pcm, d_out); });
}

int afe_batch_run_device(afe_batch *b, const int16_t *d_pcm, float *d_out)
{
    return guarded([&] {
        DeviceGuard g(b->device);
        if (b->scope == AFE_STATS_CORPUS && b->d.p.norm != AFE_NORM_NONE)
            throw Error("corpus statistics need the two-pass sequence: extract_device, corpus_stats, allreduce, normalize_device");
        b->last_launches = 0;
        const bool fuse = b->fuse_norm();
        run_extract(b, d_pcm, d_out, 0, -1, fuse);   // per-utterance scopes: the last tile of an utterance normalises it
        if (b->d.p.norm != AFE_NORM_NONE && !fuse) { run_reduce(b); run_normalize(b, d_out); }
    });
}

int afe_batch_corpus_stats(afe_batch *b, double **d_stats, int *stats_len)
{
    return guarded([&] {
        DeviceGuard g(b->device);
        if (b->d.p.norm == AFE_NORM_NONE) throw Error("corpus_stats: norm is NONE");
        run_reduce(b);
        if (d_stats) *d_stats = b->d_stats;
        if (stats_len) *stats_len = (4 * b->d.width + 1) * b->n_groups;
    });
}

int afe_normalizer_allreduce(afe_batch *b, void *comm)
{
    return guarded([&] {
        DeviceGuard g(b->device);
        if (b->scope != AFE_STATS_CORPUS) throw Error("allreduce: statistics scope is not CORPUS");
        const int w = b->d.width;
        // one fused group: sums + count (ncclSum), minima (ncclMin), maxima (ncclMax) — 4w+1 doubles, latency bound
        nccl_allreduce_stats(comm, b->d_stats, 2 * w + 1, b->d_stats + 2 * w + 1, w, b->d_stats + 3 * w + 1, w, b->stream);
    });
}

int afe_batch_set_corpus_stats(afe_batch *b, const double *h_stats, int stats_len)
{
    return guarded([&] {
        DeviceGuard g(b->device);
        if (stats_len != (4 * b->d.width + 1) * b->n_groups) throw Error("set_corpus_stats: wrong length");
        AFE_CUDA(cudaMemcpyAsync(b->d_stats, h_stats, sizeof(double) * stats_len, cudaMemcpyHostToDevice, b->stream));
        AFE_CUDA(cudaStreamSynchronize(b->stream));
    });
}

int afe_batch_normalize_device(afe_batch *b, float *d_out)
{
    return guarded([&] {
        DeviceGuard g(b->device);
        if (b->d.p.norm == AFE_NORM_NONE) return;
        run_normalize(b, d_out);
    });
}

int afe_batch_synchronize(afe_batch *b)
{
    return guarded([&] { DeviceGuard g(b->device); AFE_CUDA(cudaStreamSynchronize(b->stream)); });
}

// End-to-end with host buffers. h_pcm / h_out should be pinned for full PCIe speed (pageable memory works too).
// The shard is cut into up to 32 utterance chunks and pipelined over three streams: H2D of chunk c+1, the kernels of
// chunk c and D2H of chunk c-1 overlap (PCIe is full duplex), so the call is bound by max(H2D, D2H) instead of their sum.
int afe_batch_run_host(afe_batch *b, const int16_t *h_pcm, float *h_out)
{
    return guarded([&] {
        DeviceGuard g(b->device);
        const Derived &d = b->d;
        if (!b->d_tiles) throw Error("afe_batch_plan must be called before running");
        if (b->scope == AFE_STATS_CORPUS && d.p.norm != AFE_NORM_NONE)
            throw Error("run_host: corpus statistics need the two-pass device sequence");
        const size_t n_samples = (size_t)b->pcm_extent;
        const size_t pcm_bytes = n_samples * 2 + 32, out_bytes = (size_t)b->frame_off.back() * d.width * 4;
        if (pcm_bytes > b->pcm_stage_bytes) {
            if (b->d_pcm_stage) cudaFree(b->d_pcm_stage);
            AFE_CUDA(cudaMalloc(&b->d_pcm_stage, pcm_bytes));
            b->pcm_stage_bytes = pcm_bytes;
        }
        if (out_bytes > b->out_stage_bytes) {
            if (b->d_out_stage) cudaFree(b->d_out_stage);
            AFE_CUDA(cudaMalloc(&b->d_out_stage, out_bytes));
            b->out_stage_bytes = out_bytes;
        }
        if (!b->s_in) {
            AFE_CUDA(cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking));
            AFE_CUDA(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking));
        }
        // chunks are contiguous in samples only when the utterances are packed in increasing order
        bool ordered = true;
        for (int u = 0; u + 1 < b->n_utts; u++) ordered = ordered && b->sample_off[u] + b->sample_len[u] <= b->sample_off[u + 1];
        const int n_chunks = ordered ? std::max(1, std::min(32, b->n_utts / 8)) : 1;
        while ((int)b->ev_in.size() < n_chunks) {
            cudaEvent_t e1, e2;
            AFE_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            AFE_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            b->ev_in.push_back(e1); b->ev_k.push_back(e2);
        }
        b->last_launches = 0;
        if (b->mc_alpha != b->alpha) { build_mel_const(d, b->alpha, b->mc); b->mc_alpha = b->alpha; }
        const bool norm = d.p.norm != AFE_NORM_NONE;
        for (int c = 0; c < n_chunks; c++) {
            const int u0 = (int)((int64_t)b->n_utts * c / n_chunks), u1 = (int)((int64_t)b->n_utts * (c + 1) / n_chunks);
            if (u1 <= u0) continue;
            const int64_t s0 = n_chunks == 1 ? 0 : b->sample_off[u0];
            const int64_t s1 = n_chunks == 1 ? (int64_t)n_samples : b->sample_off[u1 - 1] + b->sample_len[u1 - 1];
            AFE_CUDA(cudaMemcpyAsync(b->d_pcm_stage + s0, h_pcm + s0, (size_t)(s1 - s0) * 2, cudaMemcpyHostToDevice, b->s_in));
            AFE_CUDA(cudaEventRecord(b->ev_in[c], b->s_in));
            AFE_CUDA(cudaStreamWaitEvent(b->stream, b->ev_in[c], 0));
            const int t0 = b->h_tile_begin[u0], t1 = b->h_tile_begin[u1];
            const bool fuse = b->fuse_norm();
            run_extract(b, b->d_pcm_stage, b->d_out_stage, t0, t1, fuse);
            if (norm && !fuse) { run_reduce(b, u0, u1); run_normalize(b, b->d_out_stage, t0, t1, u0, u1); }
            AFE_CUDA(cudaEventRecord(b->ev_k[c], b->stream));
            AFE_CUDA(cudaStreamWaitEvent(b->s_out, b->ev_k[c], 0));
            const size_t r0 = (size_t)b->frame_off[u0] * d.width, r1 = (size_t)b->frame_off[u1] * d.width;
            AFE_CUDA(cudaMemcpyAsync(h_out + r0, b->d_out_stage + r0, (r1 - r0) * 4, cudaMemcpyDeviceToHost, b->s_out));
        }
        AFE_CUDA(cudaStreamSynchronize(b->s_out));
        AFE_CUDA(cudaStreamSynchronize(b->stream));
    });
}

} // extern "C"
