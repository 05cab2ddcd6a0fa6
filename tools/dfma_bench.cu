// Measures the sustained FP64 FMA rate of the GPU (no FP64 peak is recorded in MEASURED_PEAKS.json):
// every thread runs 8 independent DFMA chains; prints TFLOP/s.  nvcc -arch=sm_100a -O3 -o tools/dfma_bench tools/dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 256, blocks = sms * 8, iters = 4096;
    double *out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * 64.0 * iters * (double)threads * blocks;
    printf("{\"sms\": %d, \"fp64_tflops\": %.2f, \"ms\": %.3f, \"dfma_per_clk_per_sm_at_1.9GHz\": %.1f}\n", sms, flops / best / 1e9, best,
           flops / 2 / (best * 1e-3) / sms / 1.9e9);
    return 0;
}
