// Cycle counts of the phases of one block column of the block-envelope Cholesky (csrc/ba_band.cu), in isolation: the 6x6 pivot
// factorisation by one thread, the forward substitution of one block row, one half-block trailing update.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o build/pivot_lat scripts/microbench/pivot_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void pivot(const double* A, double* Lkk, double* sdiag, int* fail) {
    double L[6][6], s[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[6 * j + j];
#pragma unroll
        for (int q = 0; q < 6; q++) if (q < j) d = fma(-L[j][q], L[j][q], d);
        ok = ok && d > 0 && isfinite(d);
        s[j] = rsqrt(d);
        L[j][j] = d * s[j];
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (i > j) {
                double v = A[6 * i + j];
#pragma unroll
                for (int q = 0; q < 6; q++) if (q < j) v = fma(-L[i][q], L[j][q], v);
                L[i][j] = v * s[j];
            }
    }
#pragma unroll
    for (int e = 0; e < 36; e++) Lkk[e] = e % 6 <= e / 6 ? L[e / 6][e % 6] : 0.0;
#pragma unroll
    for (int j = 0; j < 6; j++) sdiag[j] = s[j];
    if (!ok) *fail = 1;
}
// lane-parallel variant: lane i < 6 owns row i; column j: lane j takes the root, broadcasts; lanes below scale and update their rows
__device__ __forceinline__ void pivot_lanes(const double* A, double* Lkk, double* sdiag, int* fail, int lane) {
    double a[6];
    const int i = lane < 6 ? lane : 5;
#pragma unroll
    for (int c = 0; c < 6; c++) a[c] = A[6 * i + c];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        ok = ok && d > 0 && isfinite(d);
        const double s = rsqrt(d);
        a[j] *= s;   // l_ij (row j: l_jj = d s)
        if (lane == j) sdiag[j] = s;
#pragma unroll
        for (int c = 0; c < 6; c++)
            if (c > j) {
                const double lcj = __shfl_sync(0xffffffffu, a[j], c);
                a[c] = fma(-a[j], lcj, a[c]);
            }
    }
    if (lane < 6) {
#pragma unroll
        for (int c = 0; c < 6; c++) Lkk[6 * lane + c] = c <= lane ? a[c] : 0.0;
    }
    if (!ok) *fail = 1;
}
__global__ void k(double* out, long long* cyc, double seed, int busy) {
    __shared__ __align__(16) double A[36], Lkk[36], sdiag[8], lcol[36 * 32], dstb[36 * 32];
    __shared__ int fail;
    const int tid = threadIdx.x;
    if (tid < 36) A[tid] = (tid / 6 == tid % 6 ? 10.0 : 0.3) + seed * tid * 1e-3;
    for (int t = tid; t < 36 * 32; t += blockDim.x) { lcol[t] = seed + 1e-3 * t; dstb[t] = 1.0; }
    __syncthreads();
    long long t0 = 0, t1 = 0;
    double acc = seed;
    for (int rep = 0; rep < 4; rep++) {
        __syncthreads();
        if (tid == 0) {
            t0 = clock64();
            pivot(A, Lkk, sdiag, &fail);
            t1 = clock64();
            cyc[0] = t1 - t0;
        } else if (busy && tid >= 32) {
            for (int i = 0; i < 400; i++) acc = fma(acc, 1.0000001, 1e-9);
        }
        __syncthreads();
        if (tid < 32) {
            t0 = clock64();
            pivot_lanes(A, Lkk, sdiag, &fail, tid);
            t1 = clock64();
            if (tid == 0) cyc[1] = t1 - t0;
        } else if (busy) {
            for (int i = 0; i < 400; i++) acc = fma(acc, 1.0000001, 1e-9);
        }
        __syncthreads();
        // forward substitution of one block row (P2), one thread per row
        if (tid < 32) {
            t0 = clock64();
            const double2* A2 = (const double2*)(lcol + 36 * (tid / 6) + 6 * (tid % 6));
            const double2 A01 = A2[0], A23 = A2[1], A45 = A2[2];
            const double Ar[6] = {A01.x, A01.y, A23.x, A23.y, A45.x, A45.y};
            double l[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double v = Ar[c];
#pragma unroll
                for (int q = 0; q < 6; q++) if (q < c) v = fma(-l[q], Lkk[6 * c + q], v);
                l[c] = v * sdiag[c];
            }
            double2* lc = (double2*)(dstb + 36 * (tid / 6) + 6 * (tid % 6));
            lc[0] = make_double2(l[0], l[1]); lc[1] = make_double2(l[2], l[3]); lc[2] = make_double2(l[4], l[5]);
            t1 = clock64();
            if (tid == 0) cyc[2] = t1 - t0;
        }
        __syncthreads();
        // half-block update (P3), all threads of the block
        t0 = clock64();
        {
            const int pairi = (tid >> 1) & 15, half = tid & 1;
            double2* dst = (double2*)(dstb + 36 * pairi + 18 * half);
            const double2 *La = (const double2*)(lcol + 36 * (pairi / 4) + 18 * half), *Lb = (const double2*)(lcol + 36 * (pairi % 4 + 8));
            double2 Aa[9], O[9];
#pragma unroll
            for (int q = 0; q < 9; q++) { Aa[q] = La[q]; O[q] = dst[q]; }
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double2 b0 = Lb[3 * c], b1 = Lb[3 * c + 1], b2 = Lb[3 * c + 2];
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const double v = fma(Aa[3 * r + 2].y, b2.y, fma(Aa[3 * r + 2].x, b2.x, fma(Aa[3 * r + 1].y, b1.y, fma(Aa[3 * r + 1].x, b1.x, fma(Aa[3 * r].y, b0.y, Aa[3 * r].x * b0.x)))));
                    if (c & 1) O[3 * r + c / 2].y = fma(-1.0, v, O[3 * r + c / 2].y);
                    else O[3 * r + c / 2].x = fma(-1.0, v, O[3 * r + c / 2].x);
                }
            }
            if (tid < 32)
#pragma unroll
                for (int q = 0; q < 9; q++) dst[q] = O[q];
            else acc += O[0].x;
        }
        t1 = clock64();
        if (tid == 0) cyc[3] = t1 - t0;
        if (tid == 33) cyc[4] = t1 - t0;
    }
    out[tid] = acc + Lkk[tid % 36] + dstb[tid];
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
    const char* names[5] = {"pivot, one thread", "pivot, six lanes + shuffles", "forward substitution of a block row (one thread per row)", "half-block update (warp 0)", "half-block update (warp 1)"};
    for (int busy = 0; busy < 2; busy++)
        for (int threads : {32, 512}) {
            k<<<1, threads>>>(out, cyc, 1.0000001, busy);
            long long h[8];
            cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
            printf("threads per block: %d, other warps %s\n", threads, busy ? "issue dependent DFMAs meanwhile" : "idle");
            for (int i = 0; i < 5; i++) printf("  %-60s %7lld cycles\n", names[i], h[i]);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
