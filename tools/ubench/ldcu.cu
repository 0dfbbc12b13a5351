// Throughput / latency of indexed constant-bank loads feeding FFMA2, the way phase 2 of k_fused_mfcc reads its mel weights
// (there ptxas proves the index warp-uniform and emits LDCU.64; here it emits LDC.64 with a register index): footprint 1 KB against 8 KB of a kernel parameter, 1 to 4 warps
// per scheduler, dependent FFMA2 on every load.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldcu ldcu.cu
#include <cstdio>
#include <cuda_runtime.h>
struct Table { float4 w[512]; }; // 8 KB
constexpr int ITERS = 2048;
template <int FOOT4> // footprint in float4
__global__ void __launch_bounds__(512) k(float2 *out, const __grid_constant__ Table t, int start)
{
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    float2 a0 = make_float2(threadIdx.x, 1.f), a1 = a0, a2 = a0, a3 = a0;
    int off = __shfl_sync(0xffffffffu, (start + warp * 37) & (FOOT4 - 1), 0); // warp uniform -> LDCU
    for (int it = 0; it < ITERS; it++) {
        const float4 w0 = t.w[off], w1 = t.w[(off + 1) & (FOOT4 - 1)]; // 4 LDCU.64
        a0 = __ffma2_rn(a0, make_float2(w0.x, w0.y), a1);
        a1 = __ffma2_rn(a1, make_float2(w0.z, w0.w), a2);
        a2 = __ffma2_rn(a2, make_float2(w1.x, w1.y), a3);
        a3 = __ffma2_rn(a3, make_float2(w1.z, w1.w), a0);
        off = (off + 2) & (FOOT4 - 1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(a0.x + a1.x + a2.x + a3.x, a0.y + a1.y + a2.y + a3.y);
}
template <int FOOT4> void run(int warps_per_smsp, float2 *d, const Table &t)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 32 * 4 * warps_per_smsp;
    k<FOOT4><<<148, threads>>>(d, t, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<FOOT4><<<148, threads>>>(d, t, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc_per_iter = ms * 1e-3 * clk * 1e3 / ITERS; // per scheduler: warps_per_smsp warps, 4 LDCU.64 + 4 FFMA2 each
    printf("footprint %4d B, %d warp(s)/scheduler: %7.1f cycles per trip of all warps = %5.1f cycles per LDCU.64\n", FOOT4 * 16,
           warps_per_smsp, cyc_per_iter, cyc_per_iter / (4.0 * warps_per_smsp));
}
int main()
{
    float2 *d; cudaMalloc(&d, 148 * 512 * 8);
    Table t; for (int i = 0; i < 512; i++) t.w[i] = make_float4(1e-3f * i, 0.5f, 0.25f, 0.125f);
    for (int w = 1; w <= 4; w *= 2) { run<64>(w, d, t); run<512>(w, d, t); }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
