// ransac.cu — K14: RANSAC pose from 2D-3D matches (relocalisation / loop-closure candidate scoring).
//
// Replaces ucoslam::PnPSolver::solvePnPRansac (reference: src/optimization/pnpsolver.cpp:36-114), used by the relocaliser
// (src/utils/system.cpp:5026-5078) and the loop detector on every keyframe candidate:
//   per iteration: 4 random matches -> cv::solvePnP(SOLVEPNP_P3P) (:60-67) -> float 4x4 pose (Se3Transform(rv,tv)) -> reproject every
//   match in float (:72-77), inlier if squared pixel distance < 5.99 and the map point is seen under less than 60 degrees from its
//   normal (MapPoint::getViewCos(camCenter) >= 0.5, src/map_types/mappoint.h:99) (:84-97); the FIRST iteration with the most inliers
//   wins (:98-101); fewer than 4 inliers -> false; matches_io is reduced to the winner's inliers.
// The reference runs the iterations one after the other; they are independent, so here every iteration is one warp: all lanes form
// the hypothesis (p3p_math.h, double), then stride over the matches with the reference's float arithmetic, the winner is a 64-bit
// atomicMax on (inliers << 32 | ~iteration).  The 4-match samples are either given by the caller (tests replay the same samples
// through cv2.solvePnP) or drawn from a counter-based generator: the reference's std::random_shuffle stream (libstdc++ rand()) is
// not reproduced, only its distribution (4 distinct matches, uniformly).
#include "common.cuh"
#include "p3p_math.h"
#include <cstring>

namespace {

struct RansacArgs {
    const float* p3d; const float* p2d; const float* nrm; int n;
    float cam[4];   // fx fy cx cy
    int iters; const int32_t* samples; unsigned long long seed; float max_err;
    int32_t* counts; float* poses; unsigned long long* best;
};

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// the reference's inlier test for one match under the float pose M (3x4 row-major) with camera centre c (pnpsolver.cpp:72-97)
__device__ __forceinline__ bool ransac_inlier(const float* M, const float* c, const float* cam, float max_err, const float* P, const float* uv,
                                              const float* nr) {
    const float x = M[0] * P[0] + M[1] * P[1] + M[2] * P[2] + M[3];
    const float y = M[4] * P[0] + M[5] * P[1] + M[6] * P[2] + M[7];
    float z = M[8] * P[0] + M[9] * P[1] + M[10] * P[2] + M[11];
    z = (float)(1.0 / (double)z);
    const double rx = (double)(((cam[0] * x) * z) + cam[2]);
    const double ry = (double)(((cam[1] * y) * z) + cam[3]);
    const float dx = (float)((double)uv[0] - rx), dy = (float)((double)uv[1] - ry);
    if (!(dx * dx + dy * dy < max_err)) return false;
    float v[3] = {c[0] - P[0], c[1] - P[1], c[2] - P[2]};
    const double inv = 1.0 / sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
    v[0] = (float)(v[0] * inv); v[1] = (float)(v[1] * inv); v[2] = (float)(v[2] * inv);
    const float vc = v[0] * nr[0] + v[1] * nr[1] + v[2] * nr[2];
    return !(vc < 0.5f);
}

__device__ __forceinline__ void cam_centre(const float* M, float* c) {   // Se3Transform::inv() * (0,0,0), se3transform.h:89-108
    c[0] = -(M[3] * M[0] + M[7] * M[4] + M[11] * M[8]);
    c[1] = -(M[3] * M[1] + M[7] * M[5] + M[11] * M[9]);
    c[2] = -(M[3] * M[2] + M[7] * M[6] + M[11] * M[10]);
}

__global__ void __launch_bounds__(128) ransac_kernel(const __grid_constant__ RansacArgs A) {
    const int it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (it >= A.iters) return;
    int idx[4];
    if (A.samples) {
#pragma unroll
        for (int k = 0; k < 4; k++) idx[k] = A.samples[4 * it + k];
    } else {
        unsigned long long ctr = A.seed ^ ((unsigned long long)it << 20);
        for (int k = 0; k < 4; k++) {
            for (;;) {
                const int cand = (int)(splitmix64(ctr++) % (unsigned long long)A.n);
                bool dup = false;
                for (int j = 0; j < k; j++) dup = dup || idx[j] == cand;
                if (!dup) { idx[k] = cand; break; }
            }
        }
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++) ok = ok && idx[k] >= 0 && idx[k] < A.n;
    double X[4][3], px[4][2], Rb[9], tb[3];
    if (ok) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            X[k][0] = A.p3d[3 * idx[k]]; X[k][1] = A.p3d[3 * idx[k] + 1]; X[k][2] = A.p3d[3 * idx[k] + 2];
            px[k][0] = A.p2d[2 * idx[k]]; px[k][1] = A.p2d[2 * idx[k] + 1];
        }
        const double K[4] = {A.cam[0], A.cam[1], A.cam[2], A.cam[3]};
        ok = p3p_hypothesis(X, px, K, Rb, tb);
    }
    if (!ok) {   // cv::solvePnP returned false: the reference skips the iteration
        if (lane == 0) A.counts[it] = -1;
        return;
    }
    float M[12], c[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        M[4 * r] = (float)Rb[3 * r]; M[4 * r + 1] = (float)Rb[3 * r + 1]; M[4 * r + 2] = (float)Rb[3 * r + 2]; M[4 * r + 3] = (float)tb[r];
    }
    cam_centre(M, c);
    int cnt = 0;
    for (int j = lane; j < A.n; j += 32) cnt += ransac_inlier(M, c, A.cam, A.max_err, A.p3d + 3 * j, A.p2d + 2 * j, A.nrm + 3 * j) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
        A.counts[it] = cnt;
#pragma unroll
        for (int k = 0; k < 12; k++) A.poses[12 * (size_t)it + k] = M[k];
        atomicMax(A.best, ((unsigned long long)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)it));
    }
}

__global__ void __launch_bounds__(256) ransac_flags_kernel(const __grid_constant__ RansacArgs A, const float* M, uint8_t* flags) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.n) return;
    float Ml[12], c[3];
#pragma unroll
    for (int k = 0; k < 12; k++) Ml[k] = M[k];
    cam_centre(Ml, c);
    flags[j] = ransac_inlier(Ml, c, A.cam, A.max_err, A.p3d + 3 * j, A.p2d + 2 * j, A.nrm + 3 * j) ? 1 : 0;
}

}  // namespace

extern "C" {

int uco_b200_pnp_ransac(uco_b200_ctx* ctx, const float* p3d, const float* p2d, const float* normals, int n, const float* cam_fxfycxcy,
                        int max_iters, const int32_t* samples, uint64_t seed, float* pose44, int32_t* inliers, int* n_inliers,
                        int32_t* counts, int* best_iter) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!n_inliers || n < 0 || max_iters < 0 || !cam_fxfycxcy) return uco_fail(ctx, UCO_E_INVALID, "pnp_ransac: bad argument");
    *n_inliers = 0;
    if (best_iter) *best_iter = -1;
    if (n < 4 || max_iters == 0) return UCO_OK;   // pnpsolver.cpp:39: fewer than 4 matches -> false
    if (!p3d || !p2d || !normals || !pose44 || !inliers) return uco_fail(ctx, UCO_E_INVALID, "pnp_ransac: null pointer");
    const size_t b3 = (size_t)n * 12, b2 = (size_t)n * 8, bs = samples ? (size_t)max_iters * 16 : 0;
    const size_t o2 = b3, on = o2 + b2, os = on + b3, total = os + bs;
    uint8_t* h_in = (uint8_t*)uco_pinned(ctx, WS_RANSAC_IN, total);
    uint8_t* d_in = (uint8_t*)uco_ws(ctx, WS_RANSAC_IN, total);
    const size_t oc = 16, op = oc + (((size_t)max_iters * 4 + 15) & ~(size_t)15), of = op + (size_t)max_iters * 48;
    const size_t out_bytes = of + (size_t)n;
    uint8_t* d_out = (uint8_t*)uco_ws(ctx, WS_RANSAC_OUT, out_bytes);
    uint8_t* h_out = (uint8_t*)uco_pinned(ctx, WS_RANSAC_OUT, out_bytes);
    if (!h_in || !d_in || !d_out || !h_out) return UCO_E_NOMEM;
    memcpy(h_in, p3d, b3);
    memcpy(h_in + o2, p2d, b2);
    memcpy(h_in + on, normals, b3);
    if (samples) memcpy(h_in + os, samples, bs);
    UCO_CUDA(ctx, cudaMemcpyAsync(d_in, h_in, total, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemsetAsync(d_out, 0, 16, ctx->stream));
    RansacArgs A;
    A.p3d = (const float*)d_in; A.p2d = (const float*)(d_in + o2); A.nrm = (const float*)(d_in + on); A.n = n;
    memcpy(A.cam, cam_fxfycxcy, sizeof A.cam);
    A.iters = max_iters; A.samples = samples ? (const int32_t*)(d_in + os) : nullptr; A.seed = seed; A.max_err = 5.99f;   // :41
    A.counts = (int32_t*)(d_out + oc); A.poses = (float*)(d_out + op); A.best = (unsigned long long*)d_out;
    ransac_kernel<<<(max_iters + 3) / 4, 128, 0, ctx->stream>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    UCO_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, of, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned long long best = *(unsigned long long*)h_out;
    const int cnt = (int)(best >> 32), it = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFu));
    if (counts) memcpy(counts, h_out + oc, (size_t)max_iters * 4);
    if (best == 0 || cnt < 4) return UCO_OK;      // :103: fewer than 4 inliers -> false (pose and matches untouched)
    ransac_flags_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(A, A.poses + 12 * (size_t)it, d_out + of);
    UCO_LAUNCH_CHECK(ctx);
    UCO_CUDA(ctx, cudaMemcpyAsync(h_out + of, d_out + of, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int k = 0;
    for (int j = 0; j < n; j++)
        if (h_out[of + j]) inliers[k++] = j;
    *n_inliers = k;
    if (best_iter) *best_iter = it;
    const float* M = (const float*)(h_out + op) + 12 * (size_t)it;
    memcpy(pose44, M, 48);
    pose44[12] = pose44[13] = pose44[14] = 0.f;
    pose44[15] = 1.f;
    return UCO_OK;
}

// host-only: one hypothesis exactly as the kernel forms it (p3p_math.h compiled for the host), for CPU-side checks of the solver
int uco_b200_probe_p3p(const double* X4x3, const double* px4x2, const double* K_fxfycxcy, double* R9, double* t3) {
    double X[4][3], px[4][2];
    for (int i = 0; i < 4; i++) {
        for (int k = 0; k < 3; k++) X[i][k] = X4x3[3 * i + k];
        px[i][0] = px4x2[2 * i]; px[i][1] = px4x2[2 * i + 1];
    }
    return p3p_hypothesis(X, px, K_fxfycxcy, R9, t3) ? 1 : 0;
}

}  // extern "C"
