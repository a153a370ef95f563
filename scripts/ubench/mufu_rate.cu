// Micro-benchmark: MUFU.TANH (f32) vs MUFU.TANH.F16 / MUFU.EX2.F16 issue rate per SM, to decide whether
// 16-bit transcendental variants relieve the xu pipe that bounds the tcgen05 decoder.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, int iters) {
    unsigned a[8];
    for (int i = 0; i < 8; ++i) a[i] = 0x3800 + threadIdx.x + i * 77;
    float f[8];
    for (int i = 0; i < 8; ++i) f[i] = 0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
            if (MODE == 1) asm volatile("tanh.approx.f16 %0, %0;" : "+h"(*(unsigned short*)&a[i]));
            if (MODE == 2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
            if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            if (MODE == 4) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
            if (MODE == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        }
    }
    unsigned s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int per_op) {
    unsigned* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 20000;
    k<MODE><<<148, 1024>>>(out, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, 1024>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 1024 * 8.0 * iters * per_op;
    printf("%-22s %8.3f ms  %7.2f G results/s  = %.2f results/ns/SM\n", name, ms, ops / ms * 1e-6, ops / ms * 1e-6 / 148);
    cudaFree(out);
}
int main() {
    run<0>("tanh.approx.f32", 1);
    run<1>("tanh.approx.f16", 1);
    run<2>("tanh.approx.f16x2", 2);
    run<3>("ex2.approx.f32", 1);
    run<4>("ex2.approx.f16x2", 2);
    run<5>("rcp.approx.f32", 1);
    return 0;
}
