// tc_fft_proto.cu — PROTOTYPE (VERDICT r1 #3 i): stage A of the 512-point real FFT on the 5th-generation tensor cores.
//
//   int16 PCM --(CUDA cores: 5 ALU ops per sample pair)--> two fp16 planes in shared memory (x/2 = 128*h + l/2, exact)
//             --(tcgen05.mma kind::f16, fp32 accumulators in TMEM)--> Y[n2][k1] = sum_n1 w z[16 n1 + n2] e^{-2 pi i k1 n / 256}
//             --(tcgen05.ld, ONE THREAD PER FRAME)--> 16-point FFTs over n2 in registers, real split, |X| --> per-frame sum
//
// What the tensor cores take over from afe_fft.cuh: the int16 -> float conversions, the window multiplies, the first radix-16
// layer and the twiddle multiplies (all folded into the constant B operand), and the shared-memory exchange (TMEM holds
// frame x (n2, k1), so the thread that owns a frame reads all its n2 for one k1 directly).
//
// Operand A is never materialised per frame: frames overlap (hop 160 of 400 samples), so the converted samples of a round
// of 128 frames are stored ONCE, hop by hop, as Q[plane][row = hop][8 samples] and the M = 128 rows of the MMA are 128
// consecutive hops at the canonical 16-byte row pitch of the no-swizzle K-major layout; frame f's second and third hop are
// rows f+1 and f+2 of the same planes (descriptor start address / leading byte offset of 16 bytes). Inside a hop the 160
// samples are permuted so that the 8 samples of a K chunk belong to ONE residue n2 = n mod 16 (planes 0..15: n2, u = 0..3;
// planes 16..19: u = 4 for four residues), which makes most B tiles dense:
//   S1[n2]   rows f, f+1 of plane n2          -> n1 = 0..3, 5..8     N = 16 per k1-half
//   S2[j]    row f+2 of planes j and j+8      -> n1 = 10..12         N = 32 (two residues)
//   S3[g,e]  rows f, f+1 of plane 16+g        -> n1 = 4, 9           N = 32 (two of the chunk's four residues)
// Exactness: x/2 = 128*h + l/2 with h = x >> 8 (signed) and l = x & 255 are exact fp16 numbers; the constant matrix
// beta = w[.] * cos/sin * 2^22 is split beta = hi + lo (two fp16), products are exact in the fp32 accumulator.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../../asr-featext-opencl_b200/csrc \
//        -o tc_fft_proto tc_fft_proto.cu
// run:   ./tc_fft_proto [--frames-per-sm 4096] [--combos 4|3] [--validate 1] [--selftest 1]
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "afe_fft.cuh"

#define CK(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
    } while (0)

namespace {

constexpr int W = 400, S = 160, N2 = 512, M = 256, BINS = 257;
constexpr int kFrames = 128;                 // frames (MMA rows) per round
constexpr int kRows = kFrames + 2;           // hops a round touches
constexpr int kRowsAlloc = 137;              // plane pitch 137 * 16 B: consecutive planes land 4 banks apart (conflict-free stores)
constexpr int kPlanes = 20;
constexpr int kPlaneBytes = kRowsAlloc * 16;
constexpr int kQBytes = kPlanes * kPlaneBytes;        // one split
constexpr int kThreads = 512;
constexpr int kSteps = 32;                   // MMA steps per k1-half (16 S1 + 8 S2 + 8 S3)

__host__ __device__ constexpr int pi_of(int n2) { return 2 * (n2 % 8) + n2 / 8; } // TMEM column block of residue n2
__host__ __device__ constexpr int k1_of(int half, int slot)
{
    return half == 0 ? (slot == 0 ? 0 : slot == 1 ? 8 : (slot & 1) ? 16 - slot / 2 : slot / 2)
                     : ((slot & 1) ? 12 - slot / 2 : 4 + slot / 2);
}
// half 0 slots: k1 = 0, 8, 1, 15, 2, 14, 3, 13 ; half 1: 4, 12, 5, 11, 6, 10, 7, 9  (mirror pairs k1, 16 - k1 adjacent)

struct Step {          // one K = 16 MMA step (issued once per A-split x B-split combination)
    uint32_t a_off;    // byte offset of chunk 0 inside a Q split
    uint32_t a_lbo;    // byte distance chunk 0 -> chunk 1
    uint32_t b_off;    // byte offset of the tile inside a B split
    uint32_t n;        // N of the MMA (16 or 32)
    uint32_t d_col;    // first TMEM column (inside the half's 256)
};

struct Mma {            // one tcgen05.mma, descriptors relative to the dynamic shared-memory base (added at run time)
    uint64_t adesc, bdesc;
};
struct Program {
    Step step[2][kSteps];
    Mma mma[2][kSteps][4];    // [half][step][combo]: (hi,hi) (lo,hi) (hi,lo) (lo,lo)
    uint32_t b_split_bytes;   // bytes of one B split (hi or lo)
};

// ---------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spin = 0; spin < (1u << 26); spin++) { // bounded: a lost commit must trap, never hang the GPU
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (done) return;
    }
    __trap();
}

// shared-memory matrix descriptor, no swizzle, K-major: start address, leading (K) byte offset, stride (M/N) byte offset
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46; // descriptor version of sm_100
    return d;
}
// instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// four 32x32b.x4 loads (this thread's TMEM lane, 4 columns each) and ONE wait; the wait carries the 16 registers as
// read-write operands so that no use of them can be scheduled in front of it
__device__ __forceinline__ void tmem_ld4x4(const uint32_t (&taddr)[4], float (&v)[16])
{
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 4; i++)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[4 * i]), "=r"(r[4 * i + 1]), "=r"(r[4 * i + 2]), "=r"(r[4 * i + 3])
                     : "r"(taddr[i]));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------- conversion
// word = two int16 samples. x/2 = 128*h + l/2: fp16 pairs (hi, lo) for both samples, 5 ALU operations.
__device__ __forceinline__ void split_word(uint32_t w, uint32_t &hi2, uint32_t &lo2)
{
    // lo: bits 0x6000 | l  ==  512 + l/2  (ulp 0.5 in [512, 1024))
    const uint32_t lraw = (w & 0x00ff00ffu) | 0x60006000u;
    // hi: the byte b = (x >> 8) & 255 at mantissa bits [9:2] of 2^15 (ulp 32): 32768 + 128*b; b ^ 128 = h + 128
    const uint32_t hraw = ((w >> 6) & 0x03fc03fcu) ^ 0x7a007a00u; // 0x7800 | sign flip at bit 9
    const __half2 l = __hsub2(*reinterpret_cast<const __half2 *>(&lraw), __float2half2_rn(512.f));
    const __half2 h = __hsub2(*reinterpret_cast<const __half2 *>(&hraw), __float2half2_rn(49152.f)); // 32768 + 128*128
    lo2 = *reinterpret_cast<const uint32_t *>(&l);
    hi2 = *reinterpret_cast<const uint32_t *>(&h);
}

__constant__ float2 c_twp[M / 2 + 1]; // exp(-2 pi i k / 512), k = 0..128

struct Args {
    const int16_t *pcm;   // per CTA: rounds * 128 * 160 + 480 samples, CTA b starts at b * cta_stride
    long long cta_stride;
    int rounds;
    const __half *btab;   // [2 splits][b_split_bytes / 2]
    float *out_sum;       // [ctas * rounds * 128] sum of |2X| (scaled by 2^21)
    float *out_mag;       // validation: [frames][257] or nullptr
    int combos;           // 4: all of (hi,lo) x (hi,lo); 3: without lo x lo
    long long *cycles;    // [4] per CTA 0: convert, mma wait, stage B, total (clock64)
};

// One pair task of stage B: this thread's frame, mirror pair (a, b = 16 - a) (or the self-mirrored pair (0, 8)).
// ya / yb: the 16 residues n2 of Y[.][a] and Y[.][b]. Returns the sum of the |2X| it produces; optionally stores them.
__device__ __forceinline__ float stage_b_pair(float2 (&ya)[16], float2 (&yb)[16], int a, bool special, float *mag_row)
{
    using namespace afe::dev;
    fft16(ya);
    fft16(yb);
    float acc = 0.f;
    auto emit = [&](int k, float v) { acc += v; if (mag_row) mag_row[k] = v; };
    if (!special) {
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            // Z[k] with k = a + 16 k2 pairs with Z[M - k] = Z[b + 16 (15 - k2)], b = 16 - a
            const float2 za = ya[pos16(k2)], zb = yb[pos16(15 - k2)];
            const int k = a + 16 * k2;
            float2 w;
            if (k <= M / 2) w = c_twp[k];
            else { const float2 t = c_twp[M - k]; w = make_float2(-t.x, t.y); } // exp(-2 pi i k/512) = -conj(exp(-2 pi i (256-k)/512))
            const float2 c = __fadd2_rn(za, make_float2(zb.x, -zb.y));
            const float2 d = __fadd2_rn(za, make_float2(-zb.x, zb.y));
            const float2 p = cmul(d, w);
            const float2 x1 = add_mi(c, p), x2 = add_pi(c, p); // 2 X[k], 2 conj(X[M-k])
            emit(k, mag_sqrt<true>(x1.x * x1.x + x1.y * x1.y));
            emit(M - k, mag_sqrt<true>(x2.x * x2.x + x2.y * x2.y));
        }
    } else {
        // a = 0: Z[16 k2] pairs with Z[16 (16 - k2)] (Z[256] = Z[0]); ya holds k1 = 0, yb holds k1 = 8
#pragma unroll
        for (int k2 = 0; k2 <= 8; k2++) {
            const float2 za = ya[pos16(k2)], zb = ya[pos16((16 - k2) & 15)];
            const int k = 16 * k2;
            const float2 w = c_twp[k <= M / 2 ? k : 0];
            const float2 c = __fadd2_rn(za, make_float2(zb.x, -zb.y));
            const float2 d = __fadd2_rn(za, make_float2(-zb.x, zb.y));
            const float2 p = cmul(d, w);
            const float2 x1 = add_mi(c, p), x2 = add_pi(c, p);
            emit(k, mag_sqrt<true>(x1.x * x1.x + x1.y * x1.y));
            if (k2 < 8) emit(M - k, mag_sqrt<true>(x2.x * x2.x + x2.y * x2.y));
        }
        // k1 = 8: Z[8 + 16 k2] pairs with Z[8 + 16 (15 - k2)]
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            const float2 za = yb[pos16(k2)], zb = yb[pos16(15 - k2)];
            const int k = 8 + 16 * k2;
            const float2 w = c_twp[k];
            const float2 c = __fadd2_rn(za, make_float2(zb.x, -zb.y));
            const float2 d = __fadd2_rn(za, make_float2(-zb.x, zb.y));
            const float2 p = cmul(d, w);
            const float2 x1 = add_mi(c, p), x2 = add_pi(c, p);
            emit(k, mag_sqrt<true>(x1.x * x1.x + x1.y * x1.y));
            emit(M - k, mag_sqrt<true>(x2.x * x2.x + x2.y * x2.y));
        }
    }
    return acc;
}

__global__ void __launch_bounds__(kThreads, 1) k_tc_fft(const Args a, const __grid_constant__ Program prog)
{
    extern __shared__ __align__(128) unsigned char smem[];
    // [ Q hi | Q lo | B hi | B lo | partial sums | barriers ]
    unsigned char *q_hi = smem, *q_lo = smem + kQBytes;
    unsigned char *b_sm = smem + 2 * kQBytes;
    float *s_part = reinterpret_cast<float *>(b_sm + 2 * prog.b_split_bytes);           // [128][8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_part + kFrames * 8);                 // [2]
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- one-time: B operand to shared memory, TMEM allocation, barriers
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.btab);
        uint4 *dst = reinterpret_cast<uint4 *>(b_sm);
        for (uint32_t i = tid; i < 2 * prog.b_split_bytes / 16; i += kThreads) dst[i] = src[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const uint32_t q_base = smem_u32(q_hi);
    const int16_t *pcm = a.pcm + (long long)blockIdx.x * a.cta_stride;
    long long t_conv = 0, t_mma = 0, t_b = 0;
    const long long t_start = clock64();
    uint32_t parity = 0;

    for (int r = 0; r < a.rounds; r++) {
        const long long t0 = clock64();
        // ---- conversion: chunk (plane, row) = 8 samples of one hop, permuted (see the header)
        const uint32_t *words = reinterpret_cast<const uint32_t *>(pcm + (long long)r * kFrames * S);
        for (int c = tid; c < kRows * kPlanes; c += kThreads) {
            const int row = c / kPlanes, plane = c - row * kPlanes;
            const uint32_t *hop = words + row * (S / 2);
            uint32_t w[4];
            if (plane < 16) {
#pragma unroll
                for (int u = 0; u < 4; u++) w[u] = __ldg(hop + 16 * u + plane);         // n2 = plane, u = 0..3
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = __ldg(hop + 64 + (plane - 16) + 4 * j); // u = 4, n2 = g + 4 j
            }
            uint4 hi, lo;
            split_word(w[0], hi.x, lo.x); split_word(w[1], hi.y, lo.y); split_word(w[2], hi.z, lo.z); split_word(w[3], hi.w, lo.w);
            *reinterpret_cast<uint4 *>(q_hi + plane * kPlaneBytes + row * 16) = hi;
            *reinterpret_cast<uint4 *>(q_lo + plane * kPlaneBytes + row * 16) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        const long long t1 = clock64();
        // ---- MMA issue: one thread, both halves, a commit per half
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            // descriptors were built on the host relative to the shared-memory base: only the 14-bit start-address field moves
            const uint64_t qb = (uint64_t)(q_base >> 4);
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
#pragma unroll 4
                for (int s = 0; s < kSteps; s++) {
                    const uint32_t idesc = make_idesc((int)prog.step[half][s].n);
                    const uint32_t d = tmem + half * 256 + prog.step[half][s].d_col;
#pragma unroll
                    for (int combo = 0; combo < 4; combo++) {
                        if (combo < a.combos)
                            umma_f16(d, prog.mma[half][s][combo].adesc + qb, prog.mma[half][s][combo].bdesc + qb, idesc,
                                     (s < 16 && combo == 0) ? 0u : 1u); // S1 steps initialise their column block
                    }
                }
                umma_commit(&bars[half]);
            }
        }
        // ---- stage B per half: thread = frame (TMEM lane), warp / 4 = pair slot
        const int quad = warp & 3, task = warp >> 2;
        const int frame = quad * 32 + lane;
        float *mag_row = a.out_mag ? a.out_mag + ((long long)(blockIdx.x * a.rounds + r) * kFrames + frame) * BINS : nullptr;
        long long t2 = t1;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            mbar_wait(&bars[half], parity);
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (half == 0) t2 = clock64();
            float2 ya[16], yb[16];
            const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + half * 256 + task * 4;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const uint32_t ta[4] = {tbase + pi_of(4 * g) * 16, tbase + pi_of(4 * g + 1) * 16, tbase + pi_of(4 * g + 2) * 16,
                                        tbase + pi_of(4 * g + 3) * 16};
                float v[16];
                tmem_ld4x4(ta, v);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    ya[4 * g + i] = make_float2(v[4 * i], v[4 * i + 1]);
                    yb[4 * g + i] = make_float2(v[4 * i + 2], v[4 * i + 3]);
                }
            }
            const int ka = k1_of(half, 2 * task);
            const float sum = stage_b_pair(ya, yb, ka, half == 0 && task == 0, mag_row);
            s_part[frame * 8 + half * 4 + task] = sum;
        }
        parity ^= 1;
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads(); // TMEM and Q are free again; partial sums complete
        if (tid < kFrames) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) s += s_part[tid * 8 + i];
            a.out_sum[(long long)(blockIdx.x * a.rounds + r) * kFrames + tid] = s;
        }
        const long long t3 = clock64();
        t_conv += t1 - t0; t_mma += t2 - t1; t_b += t3 - t2;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    if (blockIdx.x == 0 && tid == 0 && a.cycles) {
        a.cycles[0] = t_conv; a.cycles[1] = t_mma; a.cycles[2] = t_b; a.cycles[3] = clock64() - t_start;
    }
}

// ---------------------------------------------------------------------------------------------- cost of ONE tcgen05.mma by N
// M = 128, K = 16 (one instruction), operands in shared memory (contents irrelevant), `count` instructions round-robin over
// independent accumulator blocks, one commit, one wait: cycles per instruction as the issuing thread sees them.
template <int n> __global__ void __launch_bounds__(128, 1) k_mma_cost(int count, long long *cycles)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem, base = smem_u32(smem);
    if (tid == 0) {
        constexpr uint32_t idesc = make_idesc(n);
        constexpr int blocks = 512 / n;                    // independent accumulator blocks
        const uint64_t ad = make_desc(base, 16, 128);      // A: 128 rows at 16 B pitch, second K chunk = next row (as in k_tc_fft)
        const uint64_t bd = make_desc(base + 32768, (n / 8) * 128, 128);
#pragma unroll
        for (int i = 0; i < blocks; i++) umma_f16(tmem + i * n, ad, bd, idesc, 0u); // initialise every block
        const long long t0 = clock64();
#pragma unroll 1
        for (int j = 0; j < count / 16; j++) {
#pragma unroll
            for (int i = 0; i < 16; i++) umma_f16(tmem + (i % blocks) * n, ad, bd, idesc, 1u); // straight-line: nothing but the MMAs
        }
        umma_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        cycles[0] = t1 - t0; cycles[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// ---------------------------------------------------------------------------------------------- baseline: afe_fft.cuh on CUDA cores
// the kernel's own lane-cooperative FFT (16 lanes per frame), PCM words read from global like k_fft_mag in afe_stages.cu
__global__ void __launch_bounds__(128, 4) k_base_fft(const int16_t *pcm, const float2 *window2, const float2 *tw_a, const float2 *tw_p,
                                                     float *out_sum, int frames)
{
    using C = afe::dev::FftCfg<512>;
    __shared__ float2 scratch[4 * C::FPW * C::SCR];
    __shared__ float mags[4 * C::FPW][260];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lf = lane % C::R, fw = lane / C::R;
    afe::dev::LaneConsts<512, 13> lc;
    afe::dev::load_lane_consts<512, 13>(lc, window2, tw_a, tw_p, lf);
    const int per_iter = 4 * C::FPW;
    for (int f0 = blockIdx.x * per_iter; f0 < frames; f0 += gridDim.x * per_iter) {
        const int f = f0 + warp * C::FPW + fw;
        const int fc = f < frames ? f : frames - 1;
        const uint32_t *words = reinterpret_cast<const uint32_t *>(pcm + (long long)fc * S);
        float *row = mags[warp * C::FPW + fw];
        afe::dev::fft_frame_mag<512, 13, true, false, false>(words, lc, scratch + (warp * C::FPW + fw) * C::SCR, row, lf);
        float s = 0.f;
        for (int k = lf; k < BINS; k += C::R) s += row[k];
#pragma unroll
        for (int o = C::R / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lf == 0 && f < frames) out_sum[f] = s;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- host: tables
void put_b(std::vector<__half> &hi, std::vector<__half> &lo, uint32_t tile_off_bytes, int n, int k, int col, double beta)
{
    // K-major, no swizzle: core matrix (k / 8, col / 8) of 8 rows (N) x 8 K-elements; LBO (next K chunk) = (n / 8) * 128 B
    const uint32_t off = tile_off_bytes + (k / 8) * (n / 8) * 128 + (col / 8) * 128 + (col % 8) * 16 + (k % 8) * 2;
    const __half h = __float2half_rn((float)beta);
    const __half l = __float2half_rn((float)(beta - (double)__half2float(h)));
    hi[off / 2] = h;
    lo[off / 2] = l;
}

// beta for sample (complex index n of the frame, part p) and output (k1, comp): window * twiddle * 2^22
double beta_of(const std::vector<float> &window, int n, int p, int k1, int comp)
{
    if (2 * n + p >= W) return 0.0;
    const double w = (double)window[2 * n + p] * 32768.0 * 128.0; // window carries 1/32768; 2^22 total with the x/2 of operand A
    const double th = 2.0 * M_PI * (double)k1 * (double)n / 256.0;
    // Re Y = sum w0 x0 cos + w1 x1 sin ; Im Y = sum w1 x1 cos - w0 x0 sin
    if (comp == 0) return p == 0 ? w * cos(th) : w * sin(th);
    return p == 0 ? -w * sin(th) : w * cos(th);
}

Program build_program(const std::vector<float> &window, std::vector<__half> &btab)
{
    Program pr{};
    // tile sizes per half: S1 16 x 512 B (N = 16), S2 8 x 1024 B, S3 8 x 1024 B  => 24 KB per half, 48 KB per split
    const uint32_t half_bytes = 16 * 512 + 8 * 1024 + 8 * 1024;
    pr.b_split_bytes = 2 * half_bytes;
    std::vector<__half> hi(pr.b_split_bytes / 2, __float2half(0.f)), lo(pr.b_split_bytes / 2, __float2half(0.f));
    for (int half = 0; half < 2; half++) {
        uint32_t off = half * half_bytes;
        int s = 0;
        for (int n2 = 0; n2 < 16; n2++, s++) { // S1: plane n2, rows f (hop 0) and f+1 (hop 1), u = 0..3
            Step &st = pr.step[half][s];
            st.a_off = n2 * kPlaneBytes; st.a_lbo = 16; st.b_off = off; st.n = 16; st.d_col = pi_of(n2) * 16;
            for (int c = 0; c < 2; c++)
                for (int u = 0; u < 4; u++)
                    for (int p = 0; p < 2; p++)
                        for (int slot = 0; slot < 8; slot++)
                            for (int comp = 0; comp < 2; comp++)
                                put_b(hi, lo, off, 16, 8 * c + 2 * u + p, slot * 2 + comp,
                                      beta_of(window, 80 * c + 16 * u + n2, p, k1_of(half, slot), comp));
            off += 512;
        }
        for (int j = 0; j < 8; j++, s++) { // S2: row f+2 (hop 2) of planes j and j+8
            Step &st = pr.step[half][s];
            st.a_off = j * kPlaneBytes + 2 * 16; st.a_lbo = 8 * kPlaneBytes; st.b_off = off; st.n = 32; st.d_col = pi_of(j) * 16;
            for (int c = 0; c < 2; c++) {
                const int n2 = j + 8 * c;
                for (int u = 0; u < 4; u++)
                    for (int p = 0; p < 2; p++)
                        for (int slot = 0; slot < 8; slot++)
                            for (int comp = 0; comp < 2; comp++)
                                put_b(hi, lo, off, 32, 8 * c + 2 * u + p, c * 16 + slot * 2 + comp,
                                      beta_of(window, 160 + 16 * u + n2, p, k1_of(half, slot), comp));
            }
            off += 1024;
        }
        for (int g = 0; g < 4; g++)
            for (int e = 0; e < 2; e++, s++) { // S3: rows f, f+1 of plane 16+g (u = 4), residues g + 4e and g + 4e + 8
                Step &st = pr.step[half][s];
                st.a_off = (16 + g) * kPlaneBytes; st.a_lbo = 16; st.b_off = off; st.n = 32; st.d_col = pi_of(g + 4 * e) * 16;
                for (int c = 0; c < 2; c++)
                    for (int jj = 0; jj < 4; jj++) {
                        const int n2 = g + 4 * jj;
                        if ((jj & 1) != e) continue;          // this MMA's two residues: jj = e and jj = e + 2
                        const int blk = jj >> 1;              // column block 0: n2 = g + 4e, 1: n2 = g + 4e + 8
                        for (int p = 0; p < 2; p++)
                            for (int slot = 0; slot < 8; slot++)
                                for (int comp = 0; comp < 2; comp++)
                                    put_b(hi, lo, off, 32, 8 * c + 2 * jj + p, blk * 16 + slot * 2 + comp,
                                          beta_of(window, 80 * c + 64 + n2, p, k1_of(half, slot), comp));
                    }
                off += 1024;
            }
    }
    btab = hi;
    btab.insert(btab.end(), lo.begin(), lo.end());
    for (int half = 0; half < 2; half++)
        for (int s = 0; s < kSteps; s++)
            for (int combo = 0; combo < 4; combo++) {
                const Step &st = pr.step[half][s];
                const int as = combo & 1, bs = combo >> 1;
                pr.mma[half][s][combo].adesc = make_desc(as * kQBytes + st.a_off, st.a_lbo, 128);
                pr.mma[half][s][combo].bdesc = make_desc(2 * kQBytes + bs * pr.b_split_bytes + st.b_off, (st.n / 8) * 128, 128);
            }
    return pr;
}

// float64 reference: |2 X[k]| * 2^21 of one frame (the kernel's scale: x/2 and beta * 2^22 -> 2^21 Y)
void ref_mags(const int16_t *x, const std::vector<float> &window, std::vector<double> &mag)
{
    mag.assign(BINS, 0.0);
    std::vector<double> fr(N2, 0.0);
    for (int j = 0; j < W; j++) fr[j] = (double)x[j] * (double)window[j];
    for (int k = 0; k < BINS; k++) {
        double re = 0, im = 0;
        for (int j = 0; j < W; j++) {
            const double th = -2.0 * M_PI * (double)k * j / N2;
            re += fr[j] * cos(th); im += fr[j] * sin(th);
        }
        mag[k] = 2.0 * sqrt(re * re + im * im) * 2097152.0;
    }
}

} // namespace

int main(int argc, char **argv)
{
    int frames_per_sm = 4096, combos = 4, validate = 1, sm_limit = 0, sweep = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        const int v = atoi(argv[i + 1]);
        if (k == "--frames-per-sm") frames_per_sm = v; else if (k == "--combos") combos = v; else if (k == "--validate") validate = v;
        else if (k == "--sms") sm_limit = v; else if (k == "--mma-sweep") sweep = v;
    }
    if (sweep) {
        // what one tcgen05.mma costs on this part as a function of N (M = 128, K = 16, kind::f16, SS): the figure that decides
        // whether small factorised transforms (FFT stage A: N = 16..32; mel bank: N = 48; DCT: N = 16) belong on tensor cores
        long long *d_c;
        CK(cudaMalloc(&d_c, 16));
        auto run = [&](int n, int count) {
            switch (n) {
            case 16: CK(cudaFuncSetAttribute(k_mma_cost<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); k_mma_cost<16><<<1, 128, 64 * 1024>>>(count, d_c); break;
            case 32: CK(cudaFuncSetAttribute(k_mma_cost<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); k_mma_cost<32><<<1, 128, 64 * 1024>>>(count, d_c); break;
            case 48: CK(cudaFuncSetAttribute(k_mma_cost<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); k_mma_cost<48><<<1, 128, 64 * 1024>>>(count, d_c); break;
            case 64: CK(cudaFuncSetAttribute(k_mma_cost<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); k_mma_cost<64><<<1, 128, 64 * 1024>>>(count, d_c); break;
            case 128: CK(cudaFuncSetAttribute(k_mma_cost<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); k_mma_cost<128><<<1, 128, 64 * 1024>>>(count, d_c); break;
            default: CK(cudaFuncSetAttribute(k_mma_cost<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); k_mma_cost<256><<<1, 128, 64 * 1024>>>(count, d_c); break;
            }
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
        };
        for (int n : {16, 32, 48, 64, 128, 256}) {
            const int count = 512;
            for (int rep = 0; rep < 2; rep++) run(n, count);
            long long c[2];
            CK(cudaMemcpy(c, d_c, 16, cudaMemcpyDeviceToHost));
            printf("tcgen05.mma M=128 N=%3d K=16 f16 SS: %6.1f cycles/instruction to issue, %6.1f cycles/instruction until complete "
                   "(%d instructions; ideal tensor floor 128*N/256 = %d)\n", n, (double)c[0] / count, (double)c[1] / count, count, n / 2);
        }
        return 0;
    }
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sm_limit > 0) sms = sm_limit;
    const int rounds = frames_per_sm / kFrames;
    const long long cta_samples = (long long)rounds * kFrames * S + 3 * S; // + the two extra hops and slack
    const long long total_frames = (long long)sms * rounds * kFrames;

    std::vector<float> window(W);
    for (int i = 0; i < W; i++) window[i] = (float)(0.56f - 0.46f * cos((2.0f * M_PI * i) / W)) / 32768.f;
    std::vector<__half> btab;
    const Program prog = build_program(window, btab);

    // PCM: noise + sinusoid per CTA, speech-like dynamic range
    std::vector<int16_t> pcm((size_t)sms * cta_samples);
    unsigned lcg = 1234567u;
    for (int b = 0; b < sms; b++) {
        const double w0 = 2.0 * M_PI * (100.0 + 3700.0 * ((b * 37) % 101) / 101.0) / 16000.0;
        for (long long i = 0; i < cta_samples; i++) {
            lcg = lcg * 1664525u + 1013904223u;
            const double noise = ((int)(lcg >> 16) - 32768) / 32768.0 * 5000.0;
            double v = noise + 8000.0 * sin(w0 * i);
            if (b == 1) v = (i % 1000 < 500) ? v * 0.001 : v;   // near-silent stretches
            if (b == 2) v = (i & 1) ? 32767 : -32768;            // full-scale extremes
            pcm[(size_t)b * cta_samples + i] = (int16_t)std::max(-32768.0, std::min(32767.0, std::round(v)));
        }
    }
    std::vector<float2> twp(M / 2 + 1);
    for (int k = 0; k <= M / 2; k++) twp[k] = make_float2((float)cos(-2.0 * M_PI * k / N2), (float)sin(-2.0 * M_PI * k / N2));
    CK(cudaMemcpyToSymbol(c_twp, twp.data(), sizeof(float2) * twp.size()));

    int16_t *d_pcm; __half *d_b; float *d_sum, *d_mag = nullptr; long long *d_cyc;
    CK(cudaMalloc(&d_pcm, pcm.size() * 2)); CK(cudaMemcpy(d_pcm, pcm.data(), pcm.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_b, btab.size() * 2)); CK(cudaMemcpy(d_b, btab.data(), btab.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_sum, total_frames * 4)); CK(cudaMalloc(&d_cyc, 4 * 8));
    const int val_ctas = std::min(sms, 3), val_rounds = 1;
    if (validate) CK(cudaMalloc(&d_mag, (size_t)val_ctas * val_rounds * kFrames * BINS * 4));

    const size_t smem = 2 * kQBytes + 2 * prog.b_split_bytes + kFrames * 8 * 4 + 64;
    CK(cudaFuncSetAttribute(k_tc_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    printf("tc_fft_proto: %d SMs x %d rounds x 128 frames = %lld frames, smem %zu B, B operand %u B, combos %d\n", sms, rounds,
           total_frames, smem, 2 * prog.b_split_bytes, combos);

    Args a{};
    a.pcm = d_pcm; a.cta_stride = cta_samples; a.btab = d_b; a.out_sum = d_sum; a.combos = combos; a.cycles = d_cyc;
    // ---- validation launch (small): full magnitudes against the float64 DFT
    if (validate) {
        a.rounds = val_rounds; a.out_mag = d_mag;
        k_tc_fft<<<val_ctas, kThreads, smem>>>(a, prog);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<float> mags((size_t)val_ctas * kFrames * BINS);
        CK(cudaMemcpy(mags.data(), d_mag, mags.size() * 4, cudaMemcpyDeviceToHost));
        double worst_rel_peak = 0, worst_rel_bin = 0;
        std::vector<double> ref;
        for (int b = 0; b < val_ctas; b++)
            for (int f = 0; f < kFrames; f += 7) {
                ref_mags(pcm.data() + (size_t)b * cta_samples + (size_t)f * S, window, ref);
                double peak = 1e-30;
                for (int k = 0; k < BINS; k++) peak = std::max(peak, ref[k]);
                for (int k = 0; k < BINS; k++) {
                    const double got = mags[((size_t)b * kFrames + f) * BINS + k];
                    const double err = std::fabs(got - ref[k]);
                    worst_rel_peak = std::max(worst_rel_peak, err / peak);
                    if (ref[k] > 1e-3 * peak) worst_rel_bin = std::max(worst_rel_bin, err / ref[k]);
                }
            }
        printf("validation vs float64 DFT: max |err| / frame peak = %.3g ; max rel err on bins within 60 dB of the peak = %.3g\n",
               worst_rel_peak, worst_rel_bin);
        a.out_mag = nullptr;
    }
    // ---- timing
    a.rounds = rounds;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; i++) k_tc_fft<<<sms, kThreads, smem>>>(a, prog);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int i = 0; i < reps; i++) k_tc_fft<<<sms, kThreads, smem>>>(a, prog);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    long long cyc[4];
    CK(cudaMemcpy(cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
    const double per_round = 1.0 / rounds;
    printf("tcgen05 path : %.3f ms for %lld frames = %.3f G frames/s ; per 128-frame round on CTA 0: convert %.0f, MMA wait %.0f, "
           "stage B + split + |X| %.0f cycles (total %.0f)\n",
           ms, total_frames, total_frames / (ms * 1e-3) / 1e9, cyc[0] * per_round, cyc[1] * per_round, cyc[2] * per_round,
           cyc[3] * per_round);

    // ---- baseline: the shipped CUDA-core FFT on the same frames (one long stream per launch, 4 x 128 threads per SM)
    {
        std::vector<float2> w2(M), twa(16 * 16), twpp(M / 2);
        for (int i = 0; i < M; i++) w2[i] = make_float2(2 * i < W ? window[2 * i] : 0.f, 2 * i + 1 < W ? window[2 * i + 1] : 0.f);
        for (int l = 0; l < 16; l++)
            for (int k1 = 0; k1 < 16; k1++) twa[l * 16 + k1] = make_float2((float)cos(-2.0 * M_PI * l * k1 / M), (float)sin(-2.0 * M_PI * l * k1 / M));
        for (int k = 0; k < M / 2; k++) twpp[k] = make_float2((float)cos(-2.0 * M_PI * k / N2), (float)sin(-2.0 * M_PI * k / N2));
        float2 *d_w2, *d_twa, *d_twp;
        CK(cudaMalloc(&d_w2, M * 8)); CK(cudaMalloc(&d_twa, 256 * 8)); CK(cudaMalloc(&d_twp, M / 2 * 8));
        CK(cudaMemcpy(d_w2, w2.data(), M * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_twa, twa.data(), 256 * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_twp, twpp.data(), M / 2 * 8, cudaMemcpyHostToDevice));
        const int frames = (int)std::min<long long>(total_frames, ((long long)pcm.size() - W) / S);
        float *d_sum2;
        CK(cudaMalloc(&d_sum2, (size_t)frames * 4));
        for (int i = 0; i < 2; i++) k_base_fft<<<4 * sms, 128>>>(d_pcm, d_w2, d_twa, d_twp, d_sum2, frames);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) k_base_fft<<<4 * sms, 128>>>(d_pcm, d_w2, d_twa, d_twp, d_sum2, frames);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        printf("CUDA-core path (afe_fft.cuh fft_frame_mag, 16 lanes per frame): %.3f ms for %d frames = %.3f G frames/s\n", ms, frames,
               frames / (ms * 1e-3) / 1e9);
        // per-frame sums of the two paths must agree (frames of CTA 0 are the stream's first frames)
        std::vector<float> s1(kFrames), s2(kFrames);
        CK(cudaMemcpy(s1.data(), d_sum, kFrames * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s2.data(), d_sum2, kFrames * 4, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int f = 0; f < kFrames; f++) worst = std::max(worst, std::fabs(s1[f] / 2097152.0 - s2[f]) / std::max(1e-30, (double)s2[f]));
        printf("per-frame sum of |2X|, tcgen05 vs CUDA cores, first 128 frames: max rel diff %.3g\n", worst);
    }
    return 0;
}
