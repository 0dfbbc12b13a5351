// Microbenchmark: issue cost of packed FP32 (FFMA2/FADD2) against scalar FFMA/FADD on sm_100a, alone and mixed with
// integer work. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, float seed, int iseed)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
    const float2 m = make_float2(seed * 0.5f, seed * 0.25f), c = make_float2(seed, 1.f - seed);
    int ia[4] = {iseed, iseed + 1, iseed + 2, iseed + 3};
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2) { // scalar: 2 FFMA per element
                a[i].x = fmaf(a[i].x, m.x, c.x);
                a[i].y = fmaf(a[i].y, m.y, c.y);
            } else if (MODE == 1 || MODE == 3) { // packed: 1 FFMA2
                a[i] = __ffma2_rn(a[i], m, c);
            } else if (MODE == 4) { // scalar adds
                a[i].x = a[i].x + c.x;
                a[i].y = a[i].y + c.y;
            } else if (MODE == 5) {
                a[i] = __fadd2_rn(a[i], c);
            }
            if (MODE >= 6) { // one ALU-pipe instruction (LOP3) per element
                if (MODE == 6) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
                if (MODE == 7) a[i] = __ffma2_rn(a[i], m, c);
                if (MODE == 8) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
                if (MODE == 9) a[i] = __fadd2_rn(a[i], c);
                ia[i & 3] = ia[i & 3] ^ (ia[(i + 1) & 3] & it);
            }
            if (MODE == 2 || MODE == 3) { // + 2 integer ops per element (LOP3 / IADD3 mix, not FMA pipe)
                ia[i & 3] = (ia[i & 3] ^ iseed) + (ia[(i + 1) & 3] & 0x55);
                ia[(i + 2) & 3] = (ia[(i + 2) & 3] | it) - iseed;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(ia[0] + ia[1] + ia[2] + ia[3]);
}

template <int MODE> void run(const char *name, float *d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8;
    k<MODE><<<grid, 256>>>(d, 1.0001f, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, 1.0001f, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double elems = (double)grid * 256 * ITERS * 8; // float2 elements updated
    printf("%-28s %8.3f ms  %7.2f G float2-updates/s  (%.1f TFLOP/s if fma)\n", name, ms, elems / ms * 1e-6, elems * 4 / ms * 1e-9);
}

int main()
{
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA x2", d);
    run<1>("packed FFMA2", d);
    run<2>("scalar FFMA x2 + 4 int", d);
    run<3>("packed FFMA2 + 4 int", d);
    run<4>("scalar FADD x2", d);
    run<5>("packed FADD2", d);
    run<6>("scalar FFMA x2 + 1 LOP3", d);
    run<7>("packed FFMA2 + 1 LOP3", d);
    run<8>("scalar FADD x2 + 1 LOP3", d);
    run<9>("packed FADD2 + 1 LOP3", d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
