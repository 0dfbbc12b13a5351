// Cost of feeding FFMA2 with warp-uniform weights from a kernel parameter (constant bank) in STRAIGHT-LINE code, the form a
// shape-specialised phase 2 of k_fused_mfcc has. One trip = 96 float4 of weights = 192 FFMA2 per warp.
//   MODE 0: compile-time addresses (ptxas: LDCU.128 / LDCU.64 with immediate offsets, uniform-register FFMA2 operands)
//   MODE 1: the same table indexed from a per-warp base held in a vector register (ptxas: LDC.64 R, c[0x0][R+imm])
//   MODE 2: the table in shared memory, broadcast LDS.128
//   MODE 3: no loads (weights in registers): the FFMA2 pipe alone
//   MODE 4: scalar FFMA with the weight as a direct constant-bank operand (FFMA R, R, c[0x0][imm], R), compile-time addresses
//   MODE 5: MODE 4 with one copy of the code per warp class (switch on the warp index), each copy reading its own 96 float4
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldcu2 ldcu2.cu ; cuobjdump -sass ldcu2 | grep -c LDCU.128
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N4 = 96, ITERS = 256;
struct Table { float4 w[448]; }; // 7 KB, like MelConst

template <int MODE>
__global__ void __launch_bounds__(512) k(float2 *out, const __grid_constant__ Table t, int start)
{
    __shared__ float4 s_w[448];
    for (int i = threadIdx.x; i < 448; i += blockDim.x) s_w[i] = t.w[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int base = MODE == 0 ? 0 : ((start + warp * 40) % (448 - N4));
    float2 m0 = make_float2(threadIdx.x, 1.f), m1 = make_float2(2.f, threadIdx.x);
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < N4; i++) {
            float4 w;
            if (MODE == 0) w = t.w[i + 7];
            else if (MODE == 1) w = t.w[base + i];
            else if (MODE == 2) w = s_w[base + i];
            else if (MODE >= 4) {
                w = t.w[i + 7];
                a0.x = fmaf(m0.x, w.x, a0.x); a0.y = fmaf(m0.y, w.y, a0.y); a1.x = fmaf(m1.x, w.z, a1.x); a1.y = fmaf(m1.y, w.w, a1.y);
                continue;
            } else w = make_float4(m0.y, m1.x, 0.5f, 0.25f);
            if (i & 1) { a2 = __ffma2_rn(m0, make_float2(w.x, w.y), a2); a3 = __ffma2_rn(m1, make_float2(w.z, w.w), a3); }
            else { a0 = __ffma2_rn(m0, make_float2(w.x, w.y), a0); a1 = __ffma2_rn(m1, make_float2(w.z, w.w), a1); }
        }
        m0.x += 1.f; m1.y += 1.f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(a0.x + a1.x + a2.x + a3.x, a0.y + a1.y + a2.y + a3.y);
}
template <int W> __device__ __forceinline__ void trip5(const Table &t, float2 m0, float2 m1, float2 &a0, float2 &a1)
{
#pragma unroll
    for (int i = 0; i < N4; i++) {
        const float4 w = t.w[(W * 44 + i) % 448];
        a0.x = fmaf(m0.x, w.x, a0.x); a0.y = fmaf(m0.y, w.y, a0.y); a1.x = fmaf(m1.x, w.z, a1.x); a1.y = fmaf(m1.y, w.w, a1.y);
    }
}
template <>
__global__ void __launch_bounds__(512) k<5>(float2 *out, const __grid_constant__ Table t, int start)
{
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) & 7;
    float2 m0 = make_float2(threadIdx.x, 1.f), m1 = make_float2(2.f, threadIdx.x);
    float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        switch (warp) {
        case 0: trip5<0>(t, m0, m1, a0, a1); break;
        case 1: trip5<1>(t, m0, m1, a0, a1); break;
        case 2: trip5<2>(t, m0, m1, a0, a1); break;
        case 3: trip5<3>(t, m0, m1, a0, a1); break;
        case 4: trip5<4>(t, m0, m1, a0, a1); break;
        case 5: trip5<5>(t, m0, m1, a0, a1); break;
        case 6: trip5<6>(t, m0, m1, a0, a1); break;
        default: trip5<7>(t, m0, m1, a0, a1); break;
        }
        m0.x += 1.f; m1.y += 1.f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(a0.x + a1.x, a0.y + a1.y);
}
template <int MODE> void run(const char *name, int warps_per_smsp, float2 *d, const Table &t)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 32 * 4 * warps_per_smsp;
    k<MODE><<<148, threads>>>(d, t, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148, threads>>>(d, t, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3 / ITERS; // per scheduler and trip: warps_per_smsp warps x (96 float4 + 192 FFMA2)
    printf("%-34s %d warp(s)/scheduler: %8.1f cycles per trip = %5.2f cycles per float4 of weights (2 FFMA2 = 4 pipe cycles)\n", name,
           warps_per_smsp, cyc, cyc / (N4 * warps_per_smsp));
}
int main()
{
    float2 *d; cudaMalloc(&d, 148 * 512 * 8);
    Table t; for (int i = 0; i < 448; i++) t.w[i] = make_float4(1e-3f * i, 0.5f, 0.25f, 0.125f);
    for (int w = 1; w <= 4; w *= 2) {
        run<0>("constant bank, static addresses", w, d, t); run<1>("constant bank, register base", w, d, t);
        run<2>("shared memory, broadcast LDS.128", w, d, t); run<3>("registers only", w, d, t);
        run<4>("scalar FFMA, c[] operand", w, d, t); run<5>("scalar FFMA, c[] operand, per-warp code", w, d, t);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
