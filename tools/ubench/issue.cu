// Do packed FP32 instructions (FADD2/FFMA2) block the issue port for two cycles? Mix them with ALU-pipe work (IADD3):
//   non-blocking: 8 FADD2 + 8 IADD3 take max(fma 16, alu 16, issue 16) = 16 cycles per SMSP
//   blocking    : 8*2 + 8 = 24 cycles.          Scalar reference: 16 FADD + 8 IADD3 = 24 issue cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue issue.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int MODE> __global__ void __launch_bounds__(256) k(float *out, float seed, int iseed)
{
    float2 a[8];
    int ia[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = make_float2(seed + i + threadIdx.x, seed - i); ia[i] = iseed + i * threadIdx.x; }
    const float2 c = make_float2(seed, 1.f - seed);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2) { a[i].x += c.x; a[i].y += c.y; }          // 2 FADD
            if (MODE == 1 || MODE == 3) a[i] = __fadd2_rn(a[i], c);                 // 1 FADD2
            if (MODE == 2 || MODE == 3 || MODE == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(iseed)); // 1 IADD3 (ALU pipe)
            if (MODE == 7 || MODE == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[i]) : "r"(iseed), "r"(it)); // 1 LOP3
            if (MODE == 8) a[i] = __fadd2_rn(a[i], c);
            if (MODE == 9) { a[i] = __fadd2_rn(a[i], c); asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(iseed)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[(i + 4) & 7]) : "r"(iseed), "r"(it)); }
            if (MODE == 5 || MODE == 6) { float t; asm volatile("ex2.approx.f32 %0, %1;" : "=f"(t) : "f"(a[i].x)); ia[i] ^= __float_as_int(t) & 1; }
            if (MODE == 6) a[i] = __fadd2_rn(a[i], c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y + (float)ia[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, float *d, int per_iter)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8;
    k<MODE><<<grid, 256>>>(d, 1.0001f, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<grid, 256>>>(d, 1.0001f, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // cycles per SMSP per unrolled iteration (8 elements): 16 warps per SMSP
    const double cyc = ms * 1e-3 * clk * 1e3 / ((double)ITERS * (grid * 8 / 592.0));
    printf("%-34s %7.3f ms  %6.1f cycles / (8 elements x warp)  [%d instructions]\n", name, ms, cyc, per_iter);
}
int main()
{
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("16 FADD", d, 16);
    run<1>("8 FADD2", d, 8);
    run<4>("8 IADD3", d, 8);
    run<2>("16 FADD + 8 IADD3", d, 24);
    run<3>("8 FADD2 + 8 IADD3", d, 16);
    run<7>("8 LOP3", d, 8);
    run<8>("8 FADD2 + 8 LOP3", d, 16);
    run<9>("8 FADD2 + 8 IADD3 + 8 LOP3", d, 24);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
