// undistort.cu — K15: keypoint undistortion when a Frame is built (the device-resident Frame's und_kpts).
//
// Replaces ucoslam::undistortPoints (reference: src/basictypes/misc.cpp:269-292) as FrameExtractor applies it to the extracted
// keypoints (Frame::und_kpts), the marker corners and the image bounds (src/utils/frameextractor.cpp, de-obfuscated: the statements
// after the extractor / marker-detector threads join).  One thread per point, undistort_math.h; with an empty distortion vector the
// reference's call reduces to (float)((u-cx)/fx)*fx+cx, which is what the same code computes with all coefficients zero.
#include "common.cuh"
#include "undistort_math.h"
#include <cstring>

namespace {

struct UndArgs {
    float K[4];
    double k[14];
};

__global__ void __launch_bounds__(256) undistort_points_kernel(const __grid_constant__ UndArgs A, const float2* __restrict__ in, size_t in_stride,
                                                               float2* __restrict__ out, size_t out_stride, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 p = *(const float2*)((const char*)in + (size_t)i * in_stride);
    float2 q;
    undistort_point(p.x, p.y, A.K, A.k, &q.x, &q.y);
    *(float2*)((char*)out + (size_t)i * out_stride) = q;
}

int fill_args(uco_b200_ctx* ctx, const float* K, const float* dist, int n_dist, UndArgs* A) {
    if (!K || n_dist < 0 || n_dist > 14 || (n_dist && !dist)) return uco_fail(ctx, UCO_E_INVALID, "undistort: bad camera / distortion argument");
    memcpy(A->K, K, sizeof A->K);
    for (int i = 0; i < 14; i++) A->k[i] = i < n_dist ? (double)dist[i] : 0.0;
    if (A->k[12] != 0.0 || A->k[13] != 0.0) return uco_fail(ctx, UCO_E_INVALID, "undistort: tilted sensor model (tauX, tauY) is not supported");
    return UCO_OK;
}

}  // namespace

extern "C" {

int uco_b200_undistort_points_dev(uco_b200_ctx* ctx, const float* pts_dev, size_t in_stride, int n, const float* K_fxfycxcy, const float* dist,
                                  int n_dist, float* out_dev, size_t out_stride) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    UndArgs A;
    int rc = fill_args(ctx, K_fxfycxcy, dist, n_dist, &A);
    if (rc != UCO_OK) return rc;
    if (n < 0 || in_stride < 8 || out_stride < 8 || (in_stride & 3) || (out_stride & 3)) return uco_fail(ctx, UCO_E_INVALID, "undistort: bad size / stride");
    if (n == 0) return UCO_OK;
    if (!pts_dev || !out_dev) return uco_fail(ctx, UCO_E_INVALID, "undistort: null pointer");
    undistort_points_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(A, (const float2*)pts_dev, in_stride, (float2*)out_dev, out_stride, n);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

int uco_b200_undistort_points(uco_b200_ctx* ctx, const float* pts, size_t in_stride, int n, const float* K_fxfycxcy, const float* dist,
                              int n_dist, float* out, size_t out_stride) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (n < 0 || in_stride < 8 || out_stride < 8) return uco_fail(ctx, UCO_E_INVALID, "undistort: bad size / stride");
    if (n == 0) return UCO_OK;
    if (!pts || !out) return uco_fail(ctx, UCO_E_INVALID, "undistort: null pointer");
    float* h = (float*)uco_pinned(ctx, WS_UNDISTORT, (size_t)n * 8);
    float* d = (float*)uco_ws(ctx, WS_UNDISTORT, (size_t)n * 8);
    if (!h || !d) return UCO_E_NOMEM;
    for (int i = 0; i < n; i++) memcpy(h + 2 * (size_t)i, (const char*)pts + (size_t)i * in_stride, 8);
    UCO_CUDA(ctx, cudaMemcpyAsync(d, h, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = uco_b200_undistort_points_dev(ctx, d, 8, n, K_fxfycxcy, dist, n_dist, d, 8);
    if (rc != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaMemcpyAsync(h, d, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) memcpy((char*)out + (size_t)i * out_stride, h + 2 * (size_t)i, 8);
    return UCO_OK;
}

// host-only: the same arithmetic compiled for the host, for CPU-side checks against cv2.undistortPoints
int uco_b200_probe_undistort(const float* pts, int n, const float* K_fxfycxcy, const float* dist, int n_dist, float* out) {
    if (n_dist < 0 || n_dist > 14) return UCO_E_INVALID;
    double k[14];
    for (int i = 0; i < 14; i++) k[i] = i < n_dist ? (double)dist[i] : 0.0;
    for (int i = 0; i < n; i++) undistort_point(pts[2 * i], pts[2 * i + 1], K_fxfycxcy, k, out + 2 * i, out + 2 * i + 1);
    return UCO_OK;
}

}  // extern "C"
