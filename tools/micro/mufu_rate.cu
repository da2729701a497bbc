// MUFU.EX2 throughput: f32 vs f16x2 (elements per clock per SM).  nvcc -arch=sm_100a -O3 mufu_rate.cu -o mufu_rate
#include <cstdio>
#include <cuda_fp16.h>
__global__ void k_f32(float* out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f16x2(float* out, int iters) {
    unsigned a[8];
    for (int i = 0; i < 8; ++i) a[i] = 0xB000B000u + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
    unsigned s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
__global__ void k_bf16x2(float* out, int iters) {
    unsigned a[8];
    for (int i = 0; i < 8; ++i) a[i] = 0xBE00BE00u + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
    unsigned s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int v = 0; v < 3; ++v) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) k_f32<<<148, 1024>>>(out, iters); else if (v == 1) k_f16x2<<<148, 1024>>>(out, iters); else k_bf16x2<<<148, 1024>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double instr = 148.0 * 1024 * 8.0 * iters;           // thread-level MUFU instructions
        printf("%s: %.3f ms  %.2f thread-instr/clk/SM (at %d MHz nominal) -> %.1f elements/clk/SM\n",
               v == 0 ? "ex2.f32  " : (v == 1 ? "ex2.f16x2" : "ex2.bf16x2"), ms, instr / 148 / (ms * 1e-3 * clk * 1e3),
               clk / 1000, instr * (v ? 2 : 1) / 148 / (ms * 1e-3 * clk * 1e3));
    }
    return 0;
}
