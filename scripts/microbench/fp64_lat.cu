// Dependent-chain latencies on sm_100a that size the block-envelope Cholesky's pivot step (csrc/ba_band.cu): DFMA, DMUL, rsqrt(double),
// LDS, generic LD to shared, __syncthreads at 512 threads.  nvcc -arch=sm_100a -O3 -o build/fp64_lat scripts/microbench/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
    __shared__ double sm[512];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double x = seed, y = seed * 0.5;
    long long t0, t1;
    const int N = 256;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * y;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    x = fabs(x) + 2.0;
    // rsqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) x = rsqrt(x) + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // LDS dependent chain
    int idx = threadIdx.x & 255;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) idx = ((int)sm[idx]) & 255;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // generic load to shared
    double* g = sm;
    if (seed > 1e30) g = out;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) idx = ((int)g[idx]) & 255;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // syncthreads
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // float rsqrt + 2 Newton steps in double
    x = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        double r = (double)rsqrtf((float)x);
        double h = 0.5 * x;
        r = r * fma(-h * r, r, 1.5);
        r = r * fma(-h * r, r, 1.5);
        x = r + 1.5;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    // one-warp-only DFMA chain while others idle at barrier is the same; independent DFMA throughput per warp (8 chains)
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
        a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    out[threadIdx.x] = x + idx + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
    const char* names[8] = {"DFMA dependent", "DMUL dependent", "rsqrt(double)+DADD dependent", "LDS + cvt dependent", "generic LD (shared) + cvt dependent",
                            "__syncthreads", "rsqrtf + 2 Newton (double) + DADD", "8 independent DFMA chains (per 8 DFMA)"};
    for (int threads : {32, 512}) {
        for (int rep = 0; rep < 2; rep++) k<<<1, threads>>>(out, cyc, 1.0000001);
        long long h[8];
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads per block: %d\n", threads);
        for (int i = 0; i < 8; i++) printf("  %-45s %7.1f cycles per op\n", names[i], h[i] / 256.0);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
