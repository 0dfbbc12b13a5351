// K1: the fused hot-path kernel.  int16 PCM tile -> window -> real FFT -> |X|/N2 -> mel+log -> DCT -> delta/delta-delta
// -> feature rows (+ per-tile column statistics), one HBM read of PCM and one HBM write of features per frame.
//
// Work item = (utterance, tile of `nout` output frames). A CTA computes the cepstra of its tile plus a halo of
// D = l1+l2 frames on each side (clamped at the utterance edges, where the reference replicates the edge frame,
// mfcccpu.cpp:243-254) in sub-batches of 32 frames:
//   stage 0  PCM of the sub-batch lands in shared memory by one cp.async.bulk (TMA) per sub-batch, double buffered
//            (plain vector loads when the shard is not 16-byte aligned)
//   phase 1  4 warps x (32/R frames) : in-register FFT + magnitude      (afe_fft.cuh)   -> mags[32][bins] (smem)
//   phase 2  one thread per frame    : mel + log + DCT                   (afe_mel.cuh)   -> cep[tile][cols] (smem)
//   phase 3  delta on the extended axis -> smem, then rows [static | delta | delta-delta] are written coalesced;
//            column sums / sums of squares (double) / min / max of the tile go to a per-tile partial record.
// Replaces, for whole utterances: segmenter.cl, AppleFFT fft0, mfcc.cl kernelTranspose+kernelFilter, DCT.cl,
// delta.cl and norm.cl:kernelSum.
#pragma once
#include <cfloat>

#include "afe_fft.cuh"
#include "afe_mel.cuh"

namespace afe {

struct Tile {
    long long pcm_off;   // first sample of the utterance in the packed PCM buffer
    long long out_row0;  // output row of the utterance's frame 0
    int T;               // frames in the utterance
    int t0;              // first output frame of this tile
    int nout;            // output frames of this tile
    int group;           // statistics group (utterance index, or 0 for corpus scope)
};

struct FusedArgs {
    const int16_t *pcm;
    float *out;
    const Tile *tiles;
    const float2 *window2, *tw_a, *tw_p;
    const int *edges;
    const float2 *pairs;
    const float *dct;
    double *partials;    // [ntiles][width][4] or nullptr
    int W, S, nb, dct_len, cols, width, l1, l2, nstreams;
    int q1;              // reproduce the single-block flush quirk
    int use_tma;
    int stats_rows_mode; // 0: no stats, 1: rows < T-D, 2: all rows
    int tc_max;          // capacity (frames) of the cepstra tile
    int nz;              // non-zero n1 slots of the window
    float den1, den2;    // 2*sum(l^2)
};

constexpr int kFusedThreads = 128;
constexpr int kFusedWarps = 4;
constexpr int kSubBatch = 32;

struct FusedSmem {
    int off_mbar, off_edges, off_pairs, off_dct, off_pcm, pcm_bytes, off_scratch, off_mags, off_cep, off_red, total;
};

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

template <int N2> FusedSmem fused_smem_layout(int S, int nb, int dct_len, int cols, int tc_max, int nout_max, int l2)
{
    using C = dev::FftCfg<N2>;
    FusedSmem L;
    int o = 0;
    L.off_mbar = o; o += 16;
    L.off_edges = o; o += align_up((nb + 2) * 4, 16);
    L.off_pairs = o; o += align_up(C::BINS * 8, 16);
    L.off_dct = o; o += align_up((dct_len > 0 ? nb * dct_len : 1) * 4, 16);
    L.pcm_bytes = align_up(((kSubBatch - 1) * S + N2) * 2, 16) + 16;
    L.off_pcm = o; o += 2 * L.pcm_bytes;
    L.off_scratch = o; o += align_up(kFusedWarps * C::FPW * C::SCR * 8, 16);
    const int mags = kSubBatch * C::BINS * 4;
    const int dhat = (nout_max + 2 * l2) * cols * 4;
    L.off_mags = o; o += align_up(mags > dhat ? mags : dhat, 16);
    L.off_cep = o; o += align_up(tc_max * cols * 4, 16);
    L.off_red = o; o += kFusedThreads * 4 * 8;
    L.total = o;
    return L;
}

namespace dev {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // bounded: a TMA that never completes must trap (error to the host), never hang the GPU
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
// 1-D bulk async copy global -> shared through the TMA engine, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

} // namespace dev

template <int N2, bool FAST>
__global__ void __launch_bounds__(kFusedThreads, 2) k_fused_mfcc(const FusedArgs a, const FusedSmem L)
{
    using C = dev::FftCfg<N2>;
    constexpr int R = C::R, FPW = C::FPW, SCR = C::SCR, BINS = C::BINS, SB = kSubBatch;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(smem + L.off_mbar);
    int *s_edges = reinterpret_cast<int *>(smem + L.off_edges);
    float2 *s_pairs = reinterpret_cast<float2 *>(smem + L.off_pairs);
    float *s_dct = reinterpret_cast<float *>(smem + L.off_dct);
    unsigned char *s_pcm = smem + L.off_pcm;
    float2 *s_scratch = reinterpret_cast<float2 *>(smem + L.off_scratch);
    float *s_mags = reinterpret_cast<float *>(smem + L.off_mags);
    float *s_dhat = s_mags; // aliased after the last sub-batch
    float *s_cep = reinterpret_cast<float *>(smem + L.off_cep);
    double *s_red = reinterpret_cast<double *>(smem + L.off_red);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lf = lane % R, fw = lane / R;
    const Tile tl = a.tiles[blockIdx.x];
    const int D = a.l1 + a.l2, cols = a.cols;
    const int c0f = max(0, tl.t0 - D), c1f = min(tl.T, tl.t0 + tl.nout + D);
    const int ncomp = c1f - c0f;
    const int nsub = (ncomp + SB - 1) / SB;
    const int16_t *upcm = a.pcm + tl.pcm_off;

    for (int i = tid; i < a.nb + 2; i += kFusedThreads) s_edges[i] = a.edges[i];
    for (int i = tid; i < BINS; i += kFusedThreads) s_pairs[i] = a.pairs[i];
    if (a.dct_len > 0)
        for (int i = tid; i < a.nb * a.dct_len; i += kFusedThreads) s_dct[i] = a.dct[i];
    dev::LaneConsts<N2> lc;
    dev::load_lane_consts<N2>(lc, a.window2, a.tw_a, a.tw_p, lf);

    if (a.use_tma && tid == 0) {
        dev::mbar_init(&s_mbar[0], 1);
        dev::mbar_init(&s_mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto sub_samples = [&](int s) { return (min(SB, ncomp - s * SB) - 1) * a.S + a.W; };
    auto issue_tma = [&](int s) { // one thread
        const uint32_t bytes = (uint32_t)((sub_samples(s) * 2 + 15) & ~15);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        dev::mbar_expect_tx(&s_mbar[s & 1], bytes);
        dev::tma_bulk_g2s(s_pcm + (s & 1) * L.pcm_bytes, upcm + (long long)(c0f + s * SB) * a.S, bytes, &s_mbar[s & 1]);
    };
    if (a.use_tma && tid == 0) {
        issue_tma(0);
        if (nsub > 1) issue_tma(1);
    }

    for (int s = 0; s < nsub; s++) {
        const int nf = min(SB, ncomp - s * SB);
        const unsigned char *pcm_buf;
        if (a.use_tma) {
            pcm_buf = s_pcm + (s & 1) * L.pcm_bytes;
            dev::mbar_wait(&s_mbar[s & 1], (uint32_t)((s >> 1) & 1));
        } else {
            // plain staging: 16-bit elements (any alignment); the aligned fast path is the TMA branch
            pcm_buf = s_pcm;
            const int16_t *src = upcm + (long long)(c0f + s * SB) * a.S;
            int16_t *dst = reinterpret_cast<int16_t *>(s_pcm);
            const int n = sub_samples(s);
            if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
                const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
                uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
                for (int i = tid; i < n / 2; i += kFusedThreads) d32[i] = __ldg(s32 + i);
                if ((n & 1) && tid == 0) dst[n - 1] = src[n - 1];
            } else
                for (int i = tid; i < n; i += kFusedThreads) dst[i] = src[i];
            __syncthreads();
        }
        // ---- phase 1: FFT + magnitude, FPW frames per warp iteration
        for (int it = warp; it * FPW < nf; it += kFusedWarps) {
            const int fl = it * FPW + fw;
            const bool act = fl < nf;
            const int flc = act ? fl : nf - 1;
            const uint32_t *words = reinterpret_cast<const uint32_t *>(pcm_buf) + ((flc * a.S) >> 1);
            dev::fft_frame_mag<N2, FAST>(words, a.nz, lc, s_scratch + (warp * FPW + fw) * SCR,
                                         act ? s_mags + fl * BINS : nullptr, lf);
        }
        __syncthreads();
        if (a.use_tma && tid == 0 && s + 2 < nsub) issue_tma(s + 2);
        // ---- phase 2: mel + log + DCT, one thread per frame
        if (tid < nf)
            dev::mel_dct_frame<16, FAST>(s_mags + tid * BINS, s_edges, s_pairs, s_dct, a.nb, a.dct_len,
                                         s_cep + (s * SB + tid) * cols);
        __syncthreads();
    }

    // ---- phase 3a: delta on the extended axis u in [t0-l2, t0+nout+l2), edge frames replicated (clamped index)
    const int T = tl.T, t0 = tl.t0, nout = tl.nout, l1 = a.l1, l2 = a.l2;
    if (a.nstreams >= 2) {
        const int nd = nout + 2 * l2;
        for (int idx = tid; idx < nd * cols; idx += kFusedThreads) {
            const int i = idx / cols, c = idx - i * cols;
            const int u = t0 - l2 + i;
            float num = 0.f;
            for (int l = 1; l <= l1; l++) {
                const float hi = s_cep[(dev::clampi(u + l, 0, T - 1) - c0f) * cols + c];
                const float lo = s_cep[(dev::clampi(u - l, 0, T - 1) - c0f) * cols + c];
                num = __fadd_rn(num, __fmul_rn((float)l, __fsub_rn(hi, lo))); // deltacpu.cpp:25, unfused like the CPU
            }
            s_dhat[idx] = __fdiv_rn(num, a.den1);
        }
        __syncthreads();
    }

    // ---- phase 3b: rows out, column statistics
    const int width = a.width;
    const int rpp = kFusedThreads / width; // rows per pass (width <= 128 enforced by the host)
    const bool active = tid < rpp * width;
    const int r_off = tid / width, col = tid - r_off * width;
    const int strm = col / cols, c = col - strm * cols;
    const int n_stats = a.stats_rows_mode == 1 ? T - D : (a.stats_rows_mode == 2 ? T : 0);
    double sum = 0.0, sumsq = 0.0;
    float mn = FLT_MAX, mx = -FLT_MAX;
    if (active) {
        float *orow = a.out + (tl.out_row0 + t0) * (long long)width + col;
        for (int r = r_off; r < nout; r += rpp) {
            const int t = t0 + r;
            float val, sval; // sval: the value the statistics see
            if (strm == 0) {
                // Q1 shifts only what is WRITTEN for the flushed rows; the reference takes its statistics on the
                // first block's own statics (mfcccpu.cpp:274 / :383-384), i.e. always un-shifted
                sval = s_cep[(t - c0f) * cols + c];
                val = (a.q1 && t >= T - D) ? s_cep[(t - D - c0f) * cols + c] : sval;
            } else if (strm == 1) {
                val = sval = s_dhat[(r + l2) * cols + c];
            } else {
                float num = 0.f;
                for (int l = 1; l <= l2; l++)
                    num = __fadd_rn(num, __fmul_rn((float)l, __fsub_rn(s_dhat[(r + l2 + l) * cols + c],
                                                                      s_dhat[(r + l2 - l) * cols + c])));
                val = sval = __fdiv_rn(num, a.den2);
            }
            orow[(long long)r * width] = val;
            if (t < n_stats) { // normalizercpu.cpp:31-66: double sums of float values / float products
                sum += (double)sval;
                sumsq += (double)__fmul_rn(sval, sval);
                mn = fminf(mn, sval);
                mx = fmaxf(mx, sval);
            }
        }
    }
    if (a.partials) {
        s_red[tid * 4 + 0] = sum;
        s_red[tid * 4 + 1] = sumsq;
        s_red[tid * 4 + 2] = (double)mn;
        s_red[tid * 4 + 3] = (double)mx;
        __syncthreads();
        if (tid < width) {
            double s0 = 0.0, s1 = 0.0, lo = (double)FLT_MAX, hi = -(double)FLT_MAX;
            for (int g = 0; g < rpp; g++) {
                const double *p = s_red + (g * width + tid) * 4;
                s0 += p[0]; s1 += p[1];
                lo = fmin(lo, p[2]); hi = fmax(hi, p[3]);
            }
            double *dst = a.partials + ((long long)blockIdx.x * width + tid) * 4;
            dst[0] = s0; dst[1] = s1; dst[2] = lo; dst[3] = hi;
        }
    }
}

} // namespace afe
