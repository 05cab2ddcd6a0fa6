// FP64 tensor-core (mma.sync m8n8k4 f64, SASS DMMA) check for sm_100a:
//  (1) correctness of the fragment layout used by the block kernels: C = X Y^T for 32x32 tiles in shared memory
//  (2) sustained DMMA throughput vs the vector DFMA rate (tools/dfma_bench.cu)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_bench tools/dmma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
constexpr int P = 36;   // tile pitch (doubles): conflict-free fragment loads
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void gemm_xyT(const double *X, const double *Y, double *C) {
    __shared__ double xs[32 * P], ys[32 * P];
    const int lane = threadIdx.x;
    for (int e = lane; e < 1024; e += 32) { xs[(e / 32) * P + e % 32] = X[e]; ys[(e / 32) * P + e % 32] = Y[e]; }
    __syncwarp();
    const int g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kk = 0; kk < 8; ++kk) {
        double a[4], b[4];
        for (int i = 0; i < 4; ++i) { a[i] = xs[(8 * i + g) * P + 4 * kk + q]; b[i] = ys[(8 * i + g) * P + 4 * kk + q]; }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        C[(8 * i + g) * 32 + 8 * j + 2 * q] = acc[i][j][0];
        C[(8 * i + g) * 32 + 8 * j + 2 * q + 1] = acc[i][j][1];
    }
}
__global__ void dmma_rate(double *out, int iters, double a, double b) {
    double c[16][2];
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0; for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double hx[1024], hy[1024], hc[1024];
    srand(1); for (int i = 0; i < 1024; ++i) { hx[i] = rand() / (double)RAND_MAX - 0.5; hy[i] = rand() / (double)RAND_MAX - 0.5; }
    double *dx, *dy, *dc; cudaMalloc(&dx, 8192); cudaMalloc(&dy, 8192); cudaMalloc(&dc, 8192);
    cudaMemcpy(dx, hx, 8192, cudaMemcpyHostToDevice); cudaMemcpy(dy, hy, 8192, cudaMemcpyHostToDevice);
    gemm_xyT<<<1, 32>>>(dx, dy, dc);
    cudaMemcpy(hc, dc, 8192, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int r = 0; r < 32; ++r) for (int c = 0; c < 32; ++c) { double s = 0; for (int k = 0; k < 32; ++k) s += hx[r * 32 + k] * hy[c * 32 + k]; err = fmax(err, fabs(s - hc[r * 32 + c])); }
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 256, blocks = sms * 8, iters = 2048;
    double *out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); dmma_rate<<<blocks, threads>>>(out, iters, 0.5, 0.25); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * 256.0 * 16.0 * iters * (double)(threads / 32) * blocks;
    printf("{\"dmma_layout_max_err\": %.3e, \"dmma_fp64_tflops\": %.2f, \"ms\": %.3f, \"cuda_err\": \"%s\"}\n", err, flops / best / 1e9, best,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
