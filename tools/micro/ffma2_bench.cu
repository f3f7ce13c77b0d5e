// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_bench.cu -o ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 ffma2(float2 a, float b, float2 c) {
    float2 bb = make_float2(b, b);
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&bb), rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float2* y, float h0, float h1, int iters) {
    float2 acc[8];
    float2 x = make_float2(threadIdx.x * 1e-3f, blockIdx.x * 1e-3f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = make_float2(j, -j);
    for (int t = 0; t < iters; ++t) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) { acc[j].x = fmaf(x.x, h0, acc[j].x); acc[j].y = fmaf(x.y, h0, acc[j].y); }
            else acc[j] = ffma2(x, h0, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) { acc[j].x = fmaf(x.x, h1, acc[j].x); acc[j].y = fmaf(x.y, h1, acc[j].y); }
            else acc[j] = ffma2(x, h1, acc[j]);
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s.x += acc[j].x; s.y += acc[j].y; }
    y[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float2* y; cudaMalloc(&y, 148 * 8 * 1024 * sizeof(float2));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 2; ++mode) {
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(a);
                if (mode == 0) k<0><<<148, warps * 32>>>(y, 0.999f, 1.001f, iters);
                else k<1><<<148, warps * 32>>>(y, 0.999f, 1.001f, iters);
                cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            }
            const double cfma = 148.0 * warps * 32 * iters * 16;           // complex FMAs (2 real FMA each)
            printf("warps/SM %2d  %s  %.3f ms  %.2f T real-FMA/s  (%.1f real-FMA/clk/SM @1.965GHz)\n", warps, mode ? "FFMA2" : "FFMA ", ms,
                   2 * cfma / ms / 1e9, 2 * cfma / (ms * 1e-3) / 148 / 1.965e9);
        }
    }
    return 0;
}
