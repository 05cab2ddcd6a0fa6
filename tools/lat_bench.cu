// Single-warp latency micro-benchmarks (cycles per dependent instruction) for the instructions the PDIP chain warp
// is made of: DFMA, DMMA (m8n8k4 f64), MUFU.RSQ64H-based rsqrt, SHFL of a double, LDS.64, and DMMA with ILP 2/4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_bench tools/lat_bench.cu && tools/lat_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

constexpr int N = 2048;

__global__ void bench(double *out, long long *cyc, double seed) {
    __shared__ double sm[64 * 36];
    const int lane = threadIdx.x;
    for (int i = lane; i < 64 * 36; i += 32) sm[i] = 1.0 + 1e-9 * i;
    __syncwarp();
    double x = seed + lane * 1e-3, y = 1.0000001;
    long long t0, t1;
    // 0: dependent DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
    t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    // 1: dependent DMMA
    double c[2] = {x, x};
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) dmma(c, 1e-3, y);
    t1 = clock64();
    if (lane == 0) cyc[1] = t1 - t0;
    x += c[0] + c[1];
    // 2: DMMA, 2 independent chains
    double c0[2] = {x, x}, c1[2] = {y, y};
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { dmma(c0, 1e-3, y); dmma(c1, 1e-3, x); }
    t1 = clock64();
    if (lane == 0) cyc[2] = t1 - t0;
    // 3: DMMA, 4 independent chains
    double d0[2] = {x, x}, d1[2] = {y, y}, d2[2] = {x, y}, d3[2] = {y, x};
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) { dmma(d0, 1e-3, y); dmma(d1, 1e-3, x); dmma(d2, 1e-3, y); dmma(d3, 1e-3, x); }
    t1 = clock64();
    if (lane == 0) cyc[3] = t1 - t0;
    x += c0[0] + c1[1] + d0[0] + d1[0] + d2[1] + d3[1];
    // 4: dependent rsqrt (approx + cubic correction, as in fast_rsqrt)
    double d = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        double r;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
        const double e = fma(-d * r, r, 1.0);
        r = fma(r * e, fma(0.375, e, 0.5), r);
        d = r + 2.0;
    }
    t1 = clock64();
    if (lane == 0) cyc[4] = t1 - t0;
    x += d;
    // 5: dependent SHFL of a double
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64();
    if (lane == 0) cyc[5] = t1 - t0;
    // 6: dependent LDS.64 (pointer chase through indices stored as doubles)
    for (int i = lane; i < 64 * 36; i += 32) sm[i] = (double)((i * 37 + 11) % (64 * 36));
    __syncwarp();
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int)sm[idx];
    t1 = clock64();
    if (lane == 0) cyc[6] = t1 - t0;
    // 7: STS -> __syncwarp -> LDS round trip
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
    t1 = clock64();
    if (lane == 0) cyc[7] = t1 - t0;
    // 8: dependent DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) y = y * 1.0000000001;
    t1 = clock64();
    if (lane == 0) cyc[8] = t1 - t0;
    out[lane] = x + idx + y;
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 32 * sizeof(double));
    cudaMallocManaged(&cyc, 16 * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) { bench<<<1, 32>>>(out, cyc, 1.5); cudaDeviceSynchronize(); }
    const char *names[] = {"DFMA dep", "DMMA dep", "DMMA ilp2 (per pair)", "DMMA ilp4 (per quad)", "rsqrt+corr dep", "SHFL.f64 dep", "LDS.64+cvt dep", "STS/sync/LDS/sync+DADD", "DMUL dep"};
    printf("{");
    for (int k = 0; k < 9; ++k) printf("\"%s\": %.1f%s", names[k], (double)cyc[k] / N, k < 8 ? ", " : "");
    printf("}\n");
    return cudaGetLastError() != cudaSuccess;
}
