// K1-WS: the fused hot-path kernel, warp-specialised and persistent. OPT-IN (AFE_BATCH_WS_KERNEL): built to test whether
// mixing FFT and mel work on every scheduler by construction beats k_fused_mfcc, measured, and found 1.6 % slower at
// BASELINE config 3 and slower on ragged batches (profiles/r01_ws_vs_generic.txt). Kept as the measured alternative and
// for whole-utterance tiles; covers the reference's default regression (l1 = l2 = 3, static + delta + delta-delta).
//
// Idea: in k_fused_mfcc all 8 warps of a CTA run the same phase between CTA barriers. The FFT phase is bound by the FMA
// pipe (packed FADD2/FFMA2 occupy it for two cycles: 534 pipe cycles for the 530 instructions of one call), the mel / DCT
// phase by shared-memory and constant-bank latency with the FMA pipe nearly idle, so how well an SM is used depends on
// the two co-resident CTAs happening to be in different phases (spreading the partial-cepstra sum over all 8 warps, which
// removed a natural stagger, cost 6 %, tools/gpu_ab.sh). Here the mix is built in:
//   one CTA per SM, 16 warps:  warps 0-7  PRODUCERS  stage PCM (TMA bulk copies, 2 stages per warp), FFT + |X| -> mags[b]
//                              warps 8-15 CONSUMERS  mel + log + DCT of mags[b] -> cepstra tile; at the end of a tile:
//                                                    deltas, rows, statistics, normalisation
//   every scheduler (SMSP) holds 2 producer and 2 consumer warps. Producers and consumers meet only through a ring of
//   2-3 magnitude buffers guarded by full/empty mbarriers; the consumers synchronise among themselves with a named
//   barrier (bar.sync 1, 256), the producers not at all. The CTA walks tiles blockIdx.x, +gridDim.x, ...: lane constants,
//   barrier set-up and buffer zeroing are paid once per SM, and the producers run ahead into the next tile while the
//   consumers finish the previous one.
// With one CTA per SM the cepstra tile can hold a whole utterance of up to ~12 s (the planner sizes it from what is left of
// the 227 KB): no halo frames are recomputed, and per-utterance normalisation happens before the rows are written
// (statistics pass, then a normalised-write pass over the cepstra in shared memory) instead of in place through L2.
// Arithmetic and statistics records are those of k_fused_mfcc (afe_fused.cuh): bitwise identical rows without
// normalisation, and within one rounding of the mean with it (the record of one big tile is summed in another order than
// the records of two small ones); tests/test_gpu_parity.py::test_batch_ws_equals_generic_kernel.
//
// Outcome (B200, config 3): producers alone 4.78 ms, consumers alone 3.59 + 0.87 ms, together 6.45 ms against 6.35 ms for
// k_fused_mfcc; issue-slot utilisation is the same ~60 % in both. Packed FP32 instructions hold the issue port for two
// cycles (tools/ubench/issue.cu: 8 FADD2 + 8 LOP3 take 27.8 cycles, not 16), so both kernels already sit at ~80 % of the
// (instructions + packed instructions) bound and role imbalance (idle consumers) costs what the phase mixing gains.
#pragma once
#include "afe_fused.cuh"

namespace afe {

constexpr int kWsMaxBuffers = 3;     // magnitude buffers in flight between producers and consumers (2 or 3, WsSmem::nbuf)
constexpr int kWsPcmStages = 2;      // PCM staging buffers per producer warp: the copy of round i+1 flies during the FFTs of round i
constexpr int kWsProducers = 8, kWsConsumers = 8;
constexpr int kWsThreads = 32 * (kWsProducers + kWsConsumers);
constexpr int kWsConsumerThreads = 32 * kWsConsumers;

struct WsSmem {
    int nbuf, off_bar, off_mags, mags_bytes, off_pcm, pcm_bytes, off_scratch, w_scratch, off_exch, exch_bytes, off_cep, total;
};

// tc_max = 0: everything but the cepstra tile (the planner sizes the tile from what is left of the 227 KB)
template <int N2> WsSmem ws_smem_layout(int S, int cols, int tc_max, int nbuf)
{
    using C = dev::FftCfg<N2>;
    const int kWarpFrames = kRoundFrames / kWsProducers;
    WsSmem L;
    L.nbuf = nbuf;
    int o = 0;
    // mbarriers: PCM stages [producer warp][stage] | full[nbuf] | empty[nbuf]
    L.off_bar = o; o += align_up((kWsProducers * kWsPcmStages + 2 * kWsMaxBuffers) * 8, 128);
    L.mags_bytes = align_up((kRoundFrames * kMagStride + 16) * 4, 128);     // [32][260] + pad for whole-chunk reads
    L.off_mags = o; o += nbuf * L.mags_bytes;
    L.pcm_bytes = align_up(((kWarpFrames - 1) * S + N2) * 2, 16) + 16;      // per producer warp and stage
    o = align_up(o, 128);
    L.off_pcm = o; o += kWsProducers * kWsPcmStages * L.pcm_bytes;
    o = align_up(o, 128);
    L.w_scratch = align_up(C::FPW * C::SCR * 8, 16);                        // FFT exchange tiles, per producer warp
    L.off_scratch = o; o += kWsProducers * L.w_scratch;
    o = align_up(o, 128);
    // consumer-owned: partial cepstra float4 [column group 4][filter class 8][frame 32], double buffered by round parity;
    // phase 3 reuses both halves: reduction records [rp][width][4] doubles (<= 24.6 KB) | tile record [width][4] doubles
    // (<= 4 KB, at 25 KB) | mean, scale rows (<= 1 KB, at 30 KB)
    L.exch_bytes = 4 * kWsConsumers * kRoundFrames * 16;
    L.off_exch = o; o += 2 * L.exch_bytes;
    L.off_cep = o; o += align_up(tc_max * cols * 4, 16);
    L.total = o;
    return L;
}

namespace dev {

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the consumer warps only (barrier 0 is __syncthreads)
__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(kWsConsumerThreads) : "memory");
}

// geometry of a tile as both roles see it
struct TileGeom {
    int c0f, c1f, ncomp, nrounds;
};
__device__ __forceinline__ TileGeom tile_geom(const Tile &tl, int D)
{
    TileGeom g;
    g.c0f = max(0, tl.t0 - D);
    g.c1f = min(tl.T, tl.t0 + tl.nout + D);
    g.ncomp = g.c1f - g.c0f;
    g.nrounds = (g.ncomp + kRoundFrames - 1) / kRoundFrames;
    return g;
}

} // namespace dev

template <int N2, int NZ, bool FAST, int KF>
__global__ void __launch_bounds__(kWsThreads, 1)
k_fused_ws(const FusedArgs a, const WsSmem L, const __grid_constant__ MelConst mc, const int ntiles)
{
    using C = dev::FftCfg<N2>;
    constexpr int kWarpFrames = kRoundFrames / kWsProducers;
    constexpr int R = C::R, FPW = C::FPW, SCR = C::SCR;
    static_assert(kWarpFrames % FPW == 0, "a warp's frames must fill whole FFT calls");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    uint64_t *bar_full = bars + kWsProducers * kWsPcmStages, *bar_empty = bar_full + kWsMaxBuffers;
    const int nbuf = L.nbuf;
    float *s_cep = reinterpret_cast<float *>(smem + L.off_cep);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0); // warp-uniform for the compiler
    const int D = a.l1 + a.l2, cols = a.cols;

    if (tid == 0) {
        for (int w = 0; w < kWsProducers * kWsPcmStages; w++) dev::mbar_init(bars + w, 1);
        for (int b = 0; b < nbuf; b++) {
            dev::mbar_init(bar_full + b, kWsProducers);
            dev::mbar_init(bar_empty + b, kWsConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // phase 2 reads whole 8-bin chunks past a filter's end (into the next row or the pad) with a ZERO weight: every
    // word of the magnitude buffers must always hold a finite number. Zeroed once per CTA; the FFTs only store finite values.
    for (int i = tid; i < nbuf * L.mags_bytes / 16; i += kWsThreads)
        reinterpret_cast<float4 *>(smem + L.off_mags)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    if (warp < kWsProducers) {
        // ================================ PRODUCERS: PCM -> |X| ================================
        // A producer warp sees the CTA's work as a flat sequence of items (tile, round). Of a tile it needs only where its
        // samples start and how many frames are computed.
        struct PTile { long long src0; int ncomp; }; // first sample of computed frame 0 (packed buffer), computed frames
        const int lf = lane % R, fw = lane / R;
        uint64_t *w_mbar = bars + warp * kWsPcmStages;
        unsigned char *w_pcm0 = smem + L.off_pcm + warp * kWsPcmStages * L.pcm_bytes;
        float2 *w_scratch = reinterpret_cast<float2 *>(smem + L.off_scratch + warp * L.w_scratch);
        auto load_ptile = [&](int ti) {
            const Tile t = a.tiles[a.tile_base + ti];
            const dev::TileGeom g = dev::tile_geom(t, D);
            PTile p;
            p.src0 = t.pcm_off + (long long)g.c0f * a.S;
            p.ncomp = g.ncomp;
            return p;
        };
        auto warp_frames = [&](const PTile &t, int r) { return min(kWarpFrames, t.ncomp - r * kRoundFrames - warp * kWarpFrames); };
        auto warp_src = [&](const PTile &t, int r) { return a.pcm + t.src0 + (long long)(r * kRoundFrames + warp * kWarpFrames) * a.S; };
        auto issue_tma = [&](const PTile &t, int r, int stage) { // lane 0 only; nothing to copy without a live frame
            const int nf = warp_frames(t, r);
            if (nf <= 0) return;
            const uint32_t bytes = (uint32_t)((((nf - 1) * a.S + a.W) * 2 + 15) & ~15);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            dev::mbar_expect_tx(w_mbar + stage, bytes);
            dev::tma_bulk_g2s(w_pcm0 + stage * L.pcm_bytes, warp_src(t, r), bytes, w_mbar + stage);
        };
        PTile cur = load_ptile(blockIdx.x);
        int r = 0, stage = 0;
        if (a.use_tma && lane == 0) issue_tma(cur, 0, 0); // overlaps the constant loads below
        int ti_ahead = blockIdx.x + gridDim.x;            // descriptor of the following tile, loaded a whole tile early
        PTile ahead = cur;
        if (ti_ahead < ntiles) ahead = load_ptile(ti_ahead);
        dev::LaneConsts<N2, NZ> lc;
        dev::load_lane_consts<N2, NZ>(lc, a.window2, a.tw_a, a.tw_p, lf);
        uint32_t par = 0;     // bit s: parity of PCM stage s
        int b = 0;            // magnitude buffer of the current item
        uint32_t ph = 0;      // its use parity
        for (;;) {
            // successor item: its PCM copy is issued now and flies while this item's FFTs run
            PTile nxt = cur;
            int r_nxt = r + 1;
            bool has_nxt = true;
            if (r_nxt * kRoundFrames >= cur.ncomp) {
                r_nxt = 0;
                has_nxt = ti_ahead < ntiles;
                if (has_nxt) {
                    nxt = ahead;
                    ti_ahead += gridDim.x;
                    if (ti_ahead < ntiles) ahead = load_ptile(ti_ahead);
                }
            }
            if (has_nxt && a.use_tma && lane == 0) issue_tma(nxt, r_nxt, stage ^ 1);

            float *mags = reinterpret_cast<float *>(smem + L.off_mags + b * L.mags_bytes);
            unsigned char *w_pcm = w_pcm0 + stage * L.pcm_bytes;
            const int nfw = warp_frames(cur, r);
            // the consumers have finished reading this buffer's previous contents (first use: passes at once)
            dev::mbar_wait(bar_empty + b, ph ^ 1);
            if (nfw > 0) {
                if (a.use_tma) {
                    dev::mbar_wait(w_mbar + stage, (par >> stage) & 1);
                    par ^= 1u << stage;
                } else {
                    // plain staging (any alignment): the warp's samples, 32-bit words when the source allows
                    const int16_t *src = warp_src(cur, r);
                    const int n = (nfw - 1) * a.S + a.W;
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
#pragma unroll 1
                for (int it = 0; it * FPW < nfw && !(a.debug_skip & 1); it++) {
                    const int fl = it * FPW + fw;           // frame within the warp's (adjacent frames per call)
                    const int fr = warp * kWarpFrames + fl; // frame within the round
                    const uint32_t *words = reinterpret_cast<const uint32_t *>(w_pcm) + ((fl * a.S) >> 1);
                    dev::fft_frame_mag<N2, NZ, true, false>(words, lc, w_scratch + fw * SCR, mags + mag_row(fr) * kMagStride, lf);
                }
            }
            __syncwarp();
            if (lane == 0) dev::mbar_arrive(bar_full + b); // release: this warp's magnitude rows are complete
            if (++b == nbuf) { b = 0; ph ^= 1; }
            if (!has_nxt) break;
            cur = nxt; r = r_nxt; stage ^= 1;
        }
        return;
    }

    // ================================ CONSUMERS: |X| -> cepstra -> rows ================================
    const int wc = warp - kWsProducers;          // filter class: filters wc, wc + 8, ...
    const int ctid = tid - 32 * kWsProducers;    // thread index within the consumer group
    const uint32_t mrow_off = mag_row(lane) * kMagStride * 4; // lane = frame of the round
    const int width = a.width;
    int g = 0, b = 0;  // rounds consumed so far, magnitude buffer of the current round
    uint32_t ph = 0;   // its use parity
    for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const int tile_idx = a.tile_base + ti;
        const Tile tl = a.tiles[tile_idx];
        const dev::TileGeom tg = dev::tile_geom(tl, D);
        const int c0f = tg.c0f, c1f = tg.c1f;
        for (int r = 0; r < tg.nrounds; r++, g++) {
            const int f0 = r * kRoundFrames;
            const int nfr = min(kRoundFrames, tg.ncomp - f0);
            const bool live = lane < nfr && !(a.debug_skip & 2);
            unsigned char *exch = smem + L.off_exch + (g & 1) * L.exch_bytes;
            dev::mbar_wait(bar_full + b, ph); // acquire: all magnitude rows of the round are visible
            const uint32_t mrow_s = dev::opaque(dev::smem_u32(smem + L.off_mags + b * L.mags_bytes) + mrow_off);
            float es[KF];
            if (live) {
                int woff = mc.wstart[wc]; // running float4 offset into this warp class's weight lists (uniform)
                int dsc[KF];
                float4 mf0[KF], mf1[KF];
#pragma unroll
                for (int k = 0; k < KF; k++) { // every filter's first chunk up front: 2*KF independent loads in flight
                    const int fb = wc + k * kWsConsumers;
                    dsc[k] = fb < a.nb ? mc.desc[fb] : 0;
                    const uint32_t maddr = mrow_s + ((dsc[k] & 0xffff) << 4);
                    mf0[k] = dev::lds128(maddr);
                    mf1[k] = dev::lds128(maddr + 16);
                }
#pragma unroll
                for (int k = 0; k < KF; k++) {
                    // four chains (bins 0,1 | 2,3 of every 4-bin group) in two register pairs: FFMA2, ascending bins in each
                    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
                    int n8 = dsc[k] >> 16; // 0 for a slot past num_banks (warp uniform), else >= 1
                    if (n8 > 0) {
                        uint32_t maddr = mrow_s + ((dsc[k] & 0xffff) << 4);
                        float4 m0 = mf0[k], m1 = mf1[k];
                        for (;;) {
                            const float4 w0 = mc.wl4[woff], w1 = mc.wl4[woff + 1];
                            acc0 = __ffma2_rn(make_float2(m0.x, m0.y), make_float2(w0.x, w0.y), acc0);
                            acc1 = __ffma2_rn(make_float2(m0.z, m0.w), make_float2(w0.z, w0.w), acc1);
                            acc0 = __ffma2_rn(make_float2(m1.x, m1.y), make_float2(w1.x, w1.y), acc0);
                            acc1 = __ffma2_rn(make_float2(m1.z, m1.w), make_float2(w1.z, w1.w), acc1);
                            woff += 2;
                            if (--n8 == 0) break;
                            maddr += 32;
                            m0 = dev::lds128(maddr);
                            m1 = dev::lds128(maddr + 16);
                        }
                    }
                    const float2 t = __fadd2_rn(acc0, acc1);
                    es[k] = t.x + t.y;
                }
            }
            __syncwarp();
            if (lane == 0) dev::mbar_arrive(bar_empty + b); // this warp no longer reads the buffer: producers may refill it
            if (++b == nbuf) { b = 0; ph ^= 1; }
            if (live) {
                // logs two at a time (packed); a filter slot past num_banks holds 0 -> log(1e-30), never used
#pragma unroll
                for (int k = 0; k + 1 < KF; k += 2) {
                    const float2 e2 = dev::mel_log2<FAST>(make_float2(es[k], es[k + 1]));
                    es[k] = e2.x; es[k + 1] = e2.y;
                }
                if (KF & 1) es[KF - 1] = dev::mel_log<FAST>(es[KF - 1]);
                if (a.dct_len > 0) {
                    float2 cep[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) cep[c] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < KF; k++) { // branch free: DCT rows past num_banks are zero, their es is finite
                        const int fb = wc + k * kWsConsumers;
                        const float2 e2 = make_float2(es[k], es[k]);
#pragma unroll
                        for (int c4 = 0; c4 < 4; c4++) {
                            const float4 d4 = mc.dct4[fb][c4];
                            cep[2 * c4 + 0] = __ffma2_rn(e2, make_float2(d4.x, d4.y), cep[2 * c4 + 0]);
                            cep[2 * c4 + 1] = __ffma2_rn(e2, make_float2(d4.z, d4.w), cep[2 * c4 + 1]);
                        }
                    }
                    // partial cepstra of this filter class -> the region of the warp that will sum column group c4
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++)
                        reinterpret_cast<float4 *>(exch)[(c4 * kWsConsumers + wc) * kRoundFrames + lane] =
                            make_float4(cep[2 * c4].x, cep[2 * c4].y, cep[2 * c4 + 1].x, cep[2 * c4 + 1].y);
                } else {
#pragma unroll
                    for (int k = 0; k < KF; k++)
                        if (wc + k * kWsConsumers < a.nb) s_cep[(f0 + lane) * cols + wc + k * kWsConsumers] = es[k];
                }
            }
            dev::consumer_sync(); // partial cepstra of the round are complete. The region is double buffered by round
                                  // parity: its next writers (round g + 2) come after the barrier of round g + 1, which
                                  // the summing warps reach only after the sums below.
            if (a.dct_len > 0 && wc < 4 && lane < nfr) {
                // warp wc (< 4) sums columns 4wc..4wc+3 of every frame over the filter classes, in a fixed order
                const float4 *part = reinterpret_cast<const float4 *>(exch) + wc * kWsConsumers * kRoundFrames;
                float4 t = part[lane];
#pragma unroll
                for (int w2 = 1; w2 < kWsConsumers; w2++) {
                    const float4 u = part[w2 * kRoundFrames + lane];
                    t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
                float *crow = s_cep + (f0 + lane) * cols + 4 * wc;
                if (4 * wc + 0 < a.dct_len) crow[0] = t.x;
                if (4 * wc + 1 < a.dct_len) crow[1] = t.y;
                if (4 * wc + 2 < a.dct_len) crow[2] = t.z;
                if (4 * wc + 3 < a.dct_len) crow[3] = t.w;
            }
        }
        dev::consumer_sync(); // all cepstra of the tile are in shared memory

        // ---- phase 3 (default regression, dev::phase3_l3) + per-tile statistics record
        if (!(a.debug_skip & 4)) {
            const int T = tl.T, t0 = tl.t0, nout = tl.nout;
            const int n_stats = a.stats_rows_mode == 1 ? T - D : (a.stats_rows_mode == 2 ? T : 0);
            const int rs = min(nout, max(0, n_stats - t0)); // rows [0, rs) enter the statistics
            double *s_red3 = reinterpret_cast<double *>(smem + L.off_exch);
            const bool whole = a.counters && tl.ntiles == 1; // this tile holds a whole utterance that is to be normalised
            if (whole) {
                // statistics first (MODE 1), then normalised rows straight from the cepstra (MODE 2): the rows are written
                // once and never read back. Same records, same finalisation as the tile-by-tile path below.
                double *s_rec = reinterpret_cast<double *>(smem + L.off_exch + 25 * 1024);
                float *s_mean = reinterpret_cast<float *>(smem + L.off_exch + 30 * 1024), *s_scale = s_mean + width;
                dev::phase3_l3<1>(a.stats_kind, a, tl, s_cep, c0f, c1f, s_red3, ctid, kWsConsumerThreads, rs);
                dev::consumer_sync();
                if (ctid < width) {
                    const int rp = kWsConsumerThreads / cols;
                    double s0 = 0.0, s1 = 0.0, lo = (double)FLT_MAX, hi = -(double)FLT_MAX;
                    for (int gq = 0; gq < rp; gq++) {
                        const double *p = s_red3 + (gq * width + ctid) * 4;
                        s0 += p[0]; s1 += p[1];
                        lo = fmin(lo, p[2]); hi = fmax(hi, p[3]);
                    }
                    double *dst = a.partials + ((long long)tile_idx * width + ctid) * 4;
                    dst[0] = s0; dst[1] = s1; dst[2] = lo; dst[3] = hi;
                    double *rec = s_rec + ctid * 4;
                    rec[0] = s0; rec[1] = s1; rec[2] = lo; rec[3] = hi;
                }
                dev::consumer_sync();
                if (ctid < width) { // normalizercpu.cpp:31-66
                    const double *rec = s_rec + (a.norm_after_dyn ? ctid : ctid % cols) * 4;
                    const double s0 = 0.0 + rec[0], s1 = 0.0 + rec[1], n = (double)n_stats;
                    float m = (float)(s0 / n), sc = 1.f;
                    if (a.norm_type == AFE_NORM_CVN) sc = (float)sqrt((n - 1.0) / (s1 - s0 * (s0 / n)));
                    else if (a.norm_type == AFE_NORM_MINMAX) sc = 1.f / fmaxf(fabsf((float)rec[2] - m), fabsf((float)rec[3] - m));
                    if (!a.norm_after_dyn && ctid >= cols) m = 0.f;
                    s_mean[ctid] = m; s_scale[ctid] = sc;
                }
                dev::consumer_sync();
                const int c = ctid % cols;
                const float norm3[6] = {s_mean[c], s_mean[cols + c], s_mean[2 * cols + c],
                                        s_scale[c], s_scale[cols + c], s_scale[2 * cols + c]};
                dev::phase3_l3<2>(0, a, tl, s_cep, c0f, c1f, s_red3, ctid, kWsConsumerThreads, rs, norm3);
            } else {
            dev::phase3_l3<0>(a.stats_kind, a, tl, s_cep, c0f, c1f, s_red3, ctid, kWsConsumerThreads, rs);
            if (a.partials) {
                dev::consumer_sync();
                if (ctid < width) {
                    const int rp = kWsConsumerThreads / cols;
                    double s0 = 0.0, s1 = 0.0, lo = (double)FLT_MAX, hi = -(double)FLT_MAX;
                    for (int gq = 0; gq < rp; gq++) {
                        const double *p = s_red3 + (gq * width + ctid) * 4;
                        s0 += p[0]; s1 += p[1];
                        lo = fmin(lo, p[2]); hi = fmax(hi, p[3]);
                    }
                    double *dst = a.partials + ((long long)tile_idx * width + ctid) * 4;
                    dst[0] = s0; dst[1] = s1; dst[2] = lo; dst[3] = hi;
                }
            }
            // ---- fused normalisation (per-utterance statistics scopes), as in k_fused_mfcc: the LAST tile of an utterance
            //      to finish reduces the utterance's per-tile records in tile order, finalises mean / scale
            //      (normalizercpu.cpp:31-66) and normalises the utterance's rows in place while they are L2 resident.
            if (a.counters) { // tile-by-tile utterance (longer than one tile)
                __shared__ int s_last;
                float *s_mean = reinterpret_cast<float *>(smem + L.off_exch), *s_scale = s_mean + width;
                const int rpp = kWsConsumerThreads / width; // rows per pass
                const bool active = ctid < rpp * width;
                const int r_off = ctid / width, col = ctid - r_off * width;
                __threadfence(); // rows + record of this tile are visible device-wide before the ticket is taken
                dev::consumer_sync();
                if (ctid == 0) {
                    const int ticket = atomicAdd(a.counters + tl.group, 1);
                    s_last = ticket == tl.ntiles - 1;
                    if (s_last) a.counters[tl.group] = 0; // ready for the next launch
                }
                dev::consumer_sync();
                if (s_last) {
                    __threadfence();
                    if (ctid < width) {
                        const int src_col = a.norm_after_dyn ? ctid : ctid % cols;
                        double s0 = 0.0, s1 = 0.0, lo = (double)FLT_MAX, hi = -(double)FLT_MAX;
                        for (int t = 0; t < tl.ntiles; t++) {
                            const double *p = a.partials + ((long long)(tl.tile0 + t) * width + src_col) * 4;
                            s0 += __ldcg(p); s1 += __ldcg(p + 1);
                            lo = fmin(lo, __ldcg(p + 2)); hi = fmax(hi, __ldcg(p + 3));
                        }
                        const double n = (double)n_stats;
                        float m = (float)(s0 / n), sc = 1.f;
                        if (a.norm_type == AFE_NORM_CVN) sc = (float)sqrt((n - 1.0) / (s1 - s0 * (s0 / n)));
                        else if (a.norm_type == AFE_NORM_MINMAX) sc = 1.f / fmaxf(fabsf((float)lo - m), fabsf((float)hi - m));
                        if (!a.norm_after_dyn && ctid >= cols) m = 0.f;
                        s_mean[ctid] = m; s_scale[ctid] = sc;
                    }
                    dev::consumer_sync();
                    if (active) { // thread = (row group, output column): no division in the loop
                        float *o = a.out + tl.out_row0 * (long long)width + col;
                        const float m = s_mean[col], sc = a.norm_type == AFE_NORM_CMN ? 1.f : s_scale[col];
                        const bool cmn = a.norm_type == AFE_NORM_CMN;
                        const int step = rpp * width;
                        o += r_off * width;
                        int rr = r_off;
                        constexpr int U = 8; // loads in flight per thread: the rows come from L2, ~300 cycles away
                        for (; rr + (U - 1) * rpp < T; rr += U * rpp, o += U * step) {
                            float v[U];
#pragma unroll
                            for (int k = 0; k < U; k++) v[k] = __ldcg(o + k * step);
#pragma unroll
                            for (int k = 0; k < U; k++) o[k * step] = cmn ? v[k] - m : (v[k] - m) * sc;
                        }
                        for (; rr < T; rr += rpp, o += step) *o = cmn ? __ldcg(o) - m : (__ldcg(o) - m) * sc;
                    }
                }
            }
            } // tile-by-tile
        }
        dev::consumer_sync(); // the cepstra tile and the exchange region are free for the next tile
    }
}

} // namespace afe
