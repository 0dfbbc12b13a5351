// K1: the fused hot-path kernel.  int16 PCM -> window -> real FFT -> |X|/N2 -> mel+log -> DCT -> delta/delta-delta
// -> feature rows (+ per-tile column statistics): one HBM read of PCM and one HBM write of features per frame.
//
// Work item = (utterance, tile of `nout` output frames). A CTA computes the cepstra of its tile plus a halo of
// D = l1+l2 frames on each side (clamped at the utterance edges, where the reference replicates the edge frame,
// mfcccpu.cpp:243-254). Inside the CTA every WARP IS AUTONOMOUS: it owns rounds of 8 consecutive frames and runs
//   stage 0  one cp.async.bulk (TMA, SASS UBLKCP) of the round's PCM into the warp's own staging buffer, completion on
//            the warp's own mbarrier; the next round's copy is issued as soon as the FFTs of this round are done
//   phase 1  8/FPW calls of the in-register FFT + magnitude (afe_fft.cuh)            -> warp-private mags[8][260]
//   phase 2  mel + log + DCT with 4 lanes per frame, each lane owning the filters b = q (mod 4) (summation order per
//            filter = the reference's ascending-bin order, mfcccpu.cpp:192-220)       -> cep[tile][cols] (CTA shared)
// with only __syncwarp between the steps, so warps of the resident CTAs interleave freely and no warp ever waits
// at a CTA barrier inside the loop (v1 lost 39 % of its issue slots there, profiles/r01_v1_k_fused_summary.txt).
// One __syncthreads later:
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
    const int *fidx;     // [3][nb]: per filter first bin (multiple of 4), float4 chunks, offset into wlist (float4 units)
    const float *wlist;  // concatenated triangular weights, filter by filter, zero padded to the float4 grid
    const float *dct;    // [nb][16] DCT rows zero padded to 16 columns
    double *partials;    // [ntiles][width][4] or nullptr
    int W, S, nb, dct_len, cols, width, l1, l2, nstreams, nwl;
    int q1;              // reproduce the single-block flush quirk
    int use_tma;
    int stats_rows_mode; // 0: no stats, 1: rows < T-D, 2: all rows
    int stats_kind;      // 0: none, 1: sums (CMN), 2: + sums of squares (CVN), 3: + min/max (MINMAX)
    int tc_max;          // capacity (frames) of the cepstra tile
    float rden1, rden2;  // 1 / (2*sum(l^2))
};

// Kernel shape: WARPS warps per CTA, each owning rounds of ROUND frames; phase 2 runs 32/ROUND lanes per frame.
//   <4, 8>: 4 lanes per frame in phase 2 (cheapest mel), 16 KB of shared memory per warp
//   <8, 4>: 8 lanes per frame, 10.6 KB per warp -> twice the resident warps per SM
// Magnitude row stride: 272 floats = 68 16-byte chunks. 272 = 16 (mod 32): the two rows written by one FFT call (adjacent
// frames) land 16 banks apart, and 68 = 4 (mod 8): in phase 2 the TPF lanes of a frame read TPF consecutive chunks and
// the next frame's lanes the chunks 4 further (mod 8), so a quarter warp always covers 8 distinct chunk groups.
// Columns M+1..271 are zero (weights there are zero padding; keeps 0*garbage from making NaNs).
__host__ __device__ constexpr int mag_stride(int) { return 272; }

struct FusedSmem {
    int off_mbar, off_win, off_twp, off_fidx, off_wlist, off_dct, off_warp, warp_bytes, w_pcm, w_scratch, w_mags,
        pcm_bytes, off_dd, off_red, off_cep, total;
};

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

template <int N2>
FusedSmem fused_smem_layout(int kFusedWarps, int kRound, int S, int nb, int nwl, int dct_len, int cols, int tc_max,
                            int nout_max, int l2, int nstreams)
{
    const int kFusedThreads = 32 * kFusedWarps, kMagStride = mag_stride(kRound);
    using C = dev::FftCfg<N2>;
    FusedSmem L;
    int o = 0;
    L.off_mbar = o; o += align_up(kFusedWarps * 8, 16); // one mbarrier per warp, never aliased
    L.off_win = o; o += align_up(C::M * 8, 16);
    L.off_twp = o; o += align_up(C::M / 2 * 8, 16);
    L.off_fidx = o; o += nb * 16;
    L.off_wlist = o; o += align_up(nwl * 4, 16);
    L.off_dct = o; o += align_up((dct_len > 0 ? nb * 16 : 1) * 4, 16);
    // per warp: [pcm | scratch | mags]
    int w = 0;
    L.pcm_bytes = align_up(((kRound - 1) * S + N2) * 2, 16) + 16;
    L.w_pcm = w; w += L.pcm_bytes;
    L.w_scratch = w; w += align_up(C::FPW * C::SCR * 8, 16);
    L.w_mags = w; w += align_up(kRound * kMagStride * 4, 16);
    L.warp_bytes = align_up(w, 128);
    // phase 3 reuses the per-warp area: [delta rows | delta-delta rows | reduction scratch]
    const int dhat = nstreams >= 2 ? align_up((nout_max + 2 * l2) * cols * 4, 16) : 0;
    const int dd = nstreams >= 3 ? align_up(nout_max * cols * 4, 16) : 0;
    const int phase3 = dhat + dd + kFusedThreads * 4 * 8;
    o = align_up(o, 128);
    L.off_warp = o;
    L.off_dd = o + dhat;
    L.off_red = o + dhat + dd;
    o += align_up(kFusedWarps * L.warp_bytes > phase3 ? kFusedWarps * L.warp_bytes : phase3, 128);
    L.off_cep = o; o += align_up(tc_max * cols * 4, 16);
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

template <int N2, int NZ, bool FAST, int kFusedWarps, int kRound>
__global__ void __launch_bounds__(32 * kFusedWarps, kFusedWarps == 4 ? 3 : 2) k_fused_mfcc(const FusedArgs a, const FusedSmem L)
{
    using C = dev::FftCfg<N2>;
    constexpr int kFusedThreads = 32 * kFusedWarps;
    constexpr int R = C::R, FPW = C::FPW, SCR = C::SCR, M = C::M, MS = mag_stride(kRound);
    constexpr int ITERS = kRound / FPW; // FFT calls per round
    constexpr int TPF = 32 / kRound;    // phase-2 lanes per frame
    static_assert(kRound % FPW == 0 && (TPF == 4 || TPF == 8), "unsupported kernel shape");
    extern __shared__ __align__(128) unsigned char smem[];
    float2 *s_win = reinterpret_cast<float2 *>(smem + L.off_win);
    float2 *s_twp = reinterpret_cast<float2 *>(smem + L.off_twp);
    int4 *s_fidx4 = reinterpret_cast<int4 *>(smem + L.off_fidx);
    float *s_wlist = reinterpret_cast<float *>(smem + L.off_wlist);
    float *s_dct = reinterpret_cast<float *>(smem + L.off_dct);
    float *s_cep = reinterpret_cast<float *>(smem + L.off_cep);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lf = lane % R, fw = lane / R;
    unsigned char *wbase = smem + L.off_warp + warp * L.warp_bytes;
    uint64_t *w_mbar = reinterpret_cast<uint64_t *>(smem + L.off_mbar) + warp;
    unsigned char *w_pcm = wbase + L.w_pcm;
    float2 *w_scratch = reinterpret_cast<float2 *>(wbase + L.w_scratch);
    float *w_mags = reinterpret_cast<float *>(wbase + L.w_mags);

    const Tile tl = a.tiles[blockIdx.x];
    const int D = a.l1 + a.l2, cols = a.cols;
    const int c0f = max(0, tl.t0 - D), c1f = min(tl.T, tl.t0 + tl.nout + D);
    const int ncomp = c1f - c0f;
    const int nrounds = (ncomp + kRound - 1) / kRound;
    const int16_t *upcm = a.pcm + tl.pcm_off + (long long)c0f * a.S;

    auto round_bytes = [&](int r) {
        return (uint32_t)((((min(kRound, ncomp - r * kRound) - 1) * a.S + a.W) * 2 + 15) & ~15);
    };
    // issue the first copy before anything else so that it overlaps the table loads
    if (a.use_tma && lane == 0) {
        dev::mbar_init(w_mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (warp < nrounds) {
            const uint32_t bytes = round_bytes(warp);
            dev::mbar_expect_tx(w_mbar, bytes);
            dev::tma_bulk_g2s(w_pcm, upcm + (long long)warp * kRound * a.S, bytes, w_mbar);
        }
    }
    for (int i = tid; i < M; i += kFusedThreads) s_win[i] = a.window2[i];
    for (int i = tid; i < M / 2; i += kFusedThreads) s_twp[i] = a.tw_p[i];
    for (int i = tid; i < a.nb; i += kFusedThreads) s_fidx4[i] = reinterpret_cast<const int4 *>(a.fidx)[i];
    for (int i = tid; i < a.nwl; i += kFusedThreads) s_wlist[i] = a.wlist[i];
    if (a.dct_len > 0)
        for (int i = tid; i < a.nb * 16; i += kFusedThreads) s_dct[i] = a.dct[i];
    float2 twa[16];
    dev::load_twa<N2>(twa, a.tw_a, lf);
    // the 128-bit mel loads may touch the 3 pad floats behind bin M of a magnitude row (with zero weights): keep them finite
    for (int i = lane; i < kRound * (MS - M - 1); i += 32) w_mags[(i / (MS - M - 1)) * MS + M + 1 + i % (MS - M - 1)] = 0.f;
    __syncthreads();

    const int f2 = lane / TPF, q = lane % TPF; // phase 2: frame within the round, lane within the frame
    uint32_t parity = 0;
    for (int r = warp; r < nrounds; r += kFusedWarps) {
        const int f0 = r * kRound;
        if (a.use_tma) {
            dev::mbar_wait(w_mbar, parity);
            parity ^= 1;
        } else {
            // plain staging (any alignment): the round's samples, 32-bit words when the source allows
            const int16_t *src = upcm + (long long)f0 * a.S;
            const int n = (min(kRound, ncomp - f0) - 1) * a.S + a.W;
            int16_t *dst = reinterpret_cast<int16_t *>(w_pcm);
            if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
                const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
                uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
                for (int i = lane; i < n / 2; i += 32) d32[i] = __ldg(s32 + i);
                if ((n & 1) && lane == 0) dst[n - 1] = src[n - 1];
            } else
                for (int i = lane; i < n; i += 32) dst[i] = src[i];
            __syncwarp();
        }
        // ---- phase 1: FFT + magnitude. Call `it` transforms the adjacent frames it*FPW + fw of the round: their PCM
        //      (S/2 words apart) and their magnitude rows (272 floats apart) start 16 / 8 banks apart
#pragma unroll 1
        for (int it = 0; it < ITERS; it++) {
            const int fl = it * FPW + fw;
            const uint32_t *words = reinterpret_cast<const uint32_t *>(w_pcm) + ((fl * a.S) >> 1);
            dev::fft_frame_mag<N2, NZ, true>(words, s_win, s_twp, twa, w_scratch + fw * SCR, w_mags + fl * MS, lf);
        }
        // the staging buffer is free again: prefetch this warp's next round while phase 2 runs
        if (a.use_tma && lane == 0 && r + kFusedWarps < nrounds) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bytes = round_bytes(r + kFusedWarps);
            dev::mbar_expect_tx(w_mbar, bytes);
            dev::tma_bulk_g2s(w_pcm, upcm + (long long)(r + kFusedWarps) * kRound * a.S, bytes, w_mbar);
        }
        // ---- phase 2: mel + log (+ DCT), TPF lanes per frame.
        //  (i)  the TPF lanes of a frame share every filter: lane q multiplies the 4-bin chunks q, q+TPF, ... of the
        //       filter's zero-padded weight list with the magnitudes (two 128-bit loads + 4 FMA per chunk) - no
        //       divergence, no bank conflicts;
        //  (ii) a transpose-reduce over the TPF lanes leaves the total of filter g*TPF + q in lane q, which takes the log
        //       and accumulates its share of the DCT; one final butterfly sums the DCT over the TPF lanes.
        {
            const float4 *mrow = reinterpret_cast<const float4 *>(w_mags + f2 * MS) + q;
            const float4 *wl = reinterpret_cast<const float4 *>(s_wlist) + q;
            float cep[16];
#pragma unroll
            for (int c = 0; c < 16; c++) cep[c] = 0.f;
            float *crow = s_cep + (f0 + f2) * cols;
            const bool live = f0 + f2 < ncomp;
            for (int g = 0; g < a.nb; g += TPF) {
                float p[TPF];
#pragma unroll
                for (int k = 0; k < TPF; k++) {
                    const int b = min(g + k, a.nb - 1); // (a short last group recomputes the last filter; unused)
                    const int4 fi = s_fidx4[b];         // {first chunk, iterations, weight offset (float4 units), -}
                    const float4 *mv = mrow + fi.x;
                    const float4 *wv = wl + fi.z;
                    float4 m = mv[0], w = wv[0];
                    float a0 = w.x * m.x, a1 = w.y * m.y, a2 = w.z * m.z, a3 = w.w * m.w;
#pragma unroll 1
                    for (int i = 1; i < fi.y; i++) { // most filters need one chunk per lane
                        m = mv[i * TPF]; w = wv[i * TPF];
                        a0 = fmaf(w.x, m.x, a0); a1 = fmaf(w.y, m.y, a1);
                        a2 = fmaf(w.z, m.z, a2); a3 = fmaf(w.w, m.w, a3);
                    }
                    p[k] = (a0 + a1) + (a2 + a3);
                }
                // transpose-reduce: afterwards lane q holds sum over the TPF lanes of p[q]
                float tot;
                if (TPF == 8) {
                    const bool h4 = q & 4;
                    float r[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float keep = h4 ? p[k + 4] : p[k], send = h4 ? p[k] : p[k + 4];
                        r[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) p[k] = r[k];
                }
                {
                    const bool h2 = q & 2, h1 = q & 1;
                    const float k0 = h2 ? p[2] : p[0], k1 = h2 ? p[3] : p[1];
                    const float s0 = h2 ? p[0] : p[2], s1 = h2 ? p[1] : p[3];
                    const float u0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
                    const float u1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
                    tot = (h1 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, h1 ? u0 : u1, 1);
                }
                const int b = g + q;
                const float e = dev::mel_log<FAST>(tot);
                if (b < a.nb) {
                    if (a.dct_len > 0) {
                        const float4 *row = reinterpret_cast<const float4 *>(s_dct + b * 16);
#pragma unroll
                        for (int c4 = 0; c4 < 4; c4++) {
                            const float4 d4 = row[c4];
                            cep[4 * c4 + 0] = fmaf(e, d4.x, cep[4 * c4 + 0]);
                            cep[4 * c4 + 1] = fmaf(e, d4.y, cep[4 * c4 + 1]);
                            cep[4 * c4 + 2] = fmaf(e, d4.z, cep[4 * c4 + 2]);
                            cep[4 * c4 + 3] = fmaf(e, d4.w, cep[4 * c4 + 3]);
                        }
                    } else if (live)
                        crow[b] = e;
                }
            }
            if (a.dct_len > 0) {
                // sum over the TPF lanes of a frame; lane q then writes the columns c = q (mod TPF)
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    cep[c] += __shfl_xor_sync(0xffffffffu, cep[c], 1);
                    cep[c] += __shfl_xor_sync(0xffffffffu, cep[c], 2);
                    if (TPF == 8) cep[c] += __shfl_xor_sync(0xffffffffu, cep[c], 4);
                }
#pragma unroll
                for (int c = 0; c < 16; c++)
                    if ((c % TPF) == q && live && c < a.dct_len) crow[c] = cep[c];
            }
        }
        __syncwarp(); // mags are rewritten by the next round's phase 1
    }
    __syncthreads();

    // ---- phase 3: every thread owns ONE column (c = tid % cols / col = tid % width) and strides over rows, so the loops
    //      are uniform (no integer division, no stream-dependent branch inside them).
    float *s_dhat = reinterpret_cast<float *>(smem + L.off_warp);
    float *s_dd = reinterpret_cast<float *>(smem + L.off_dd);
    double *s_red = reinterpret_cast<double *>(smem + L.off_red);
    const int T = tl.T, t0 = tl.t0, nout = tl.nout, l1 = a.l1, l2 = a.l2;
    if (a.nstreams >= 2) {
        // 3a: delta on the extended axis u in [t0-l2, t0+nout+l2), edge frames replicated (clamped index)
        const int rp = kFusedThreads / cols, rl = tid / cols, c = tid - rl * cols;
        const int nd = nout + 2 * l2;
        if (rl < rp) {
            for (int i = rl; i < nd; i += rp) {
                const int u = t0 - l2 + i;
                float num = 0.f;
                for (int l = 1; l <= l1; l++) {
                    const float hi = s_cep[(dev::clampi(u + l, 0, T - 1) - c0f) * cols + c];
                    const float lo = s_cep[(dev::clampi(u - l, 0, T - 1) - c0f) * cols + c];
                    num = fmaf((float)l, hi - lo, num); // deltacpu.cpp:25
                }
                s_dhat[i * cols + c] = num * a.rden1;
            }
        }
        __syncthreads();
        if (a.nstreams >= 3) { // 3b: delta-delta of the extended delta rows
            if (rl < rp) {
                for (int r = rl; r < nout; r += rp) {
                    const float *dc = s_dhat + (r + l2) * cols + c;
                    float num = 0.f;
                    for (int l = 1; l <= l2; l++) num = fmaf((float)l, dc[l * cols] - dc[-l * cols], num);
                    s_dd[r * cols + c] = num * a.rden2;
                }
            }
            __syncthreads();
        }
    }

    // 3c: rows out (coalesced: consecutive threads write consecutive floats), column statistics
    const int width = a.width;
    const int rpp = kFusedThreads / width; // rows per pass (width <= 128 enforced by the host)
    const bool active = tid < rpp * width;
    const int r_off = tid / width, col = tid - r_off * width;
    const int strm = col / cols, c = col - strm * cols;
    const int n_stats = a.stats_rows_mode == 1 ? T - D : (a.stats_rows_mode == 2 ? T : 0);
    // source of this thread's column: src[r * cols]; statics of the Q1 rows come from D rows earlier
    const float *src = strm == 0 ? s_cep + (t0 - c0f) * cols + c : strm == 1 ? s_dhat + l2 * cols + c : s_dd + c;
    const int rq = (a.q1 && strm == 0) ? max(0, T - D - t0) : nout; // first row written with the shifted static
    const int rs = min(nout, max(0, n_stats - t0));                 // rows [0, rs) enter the statistics
    double sum = 0.0, sumsq = 0.0;
    float mn = FLT_MAX, mx = -FLT_MAX;
    if (active) {
        float *orow = a.out + (tl.out_row0 + t0) * (long long)width + col;
        for (int r = r_off; r < nout; r += rpp) {
            const float sval = src[r * cols];
            // Q1 shifts only what is WRITTEN for the flushed rows; the reference takes its statistics on the first
            // block's own statics (mfcccpu.cpp:274 / :383-384), i.e. always un-shifted
            orow[(long long)r * width] = r >= rq ? src[(r - D) * cols] : sval;
            if (r < rs) { // normalizercpu.cpp:31-66: double sums of float values / float products
                if (a.stats_kind >= 1) sum += (double)sval;
                if (a.stats_kind >= 2) sumsq += (double)__fmul_rn(sval, sval);
                if (a.stats_kind >= 3) { mn = fminf(mn, sval); mx = fmaxf(mx, sval); }
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
