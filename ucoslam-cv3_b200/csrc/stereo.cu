// stereo.cu — K12: stereo depth association of a rectified pair (BASELINE config 3).
//
// Replaces the association loop of FrameExtractor::processStereo (reference: src/utils/frameextractor.cpp:1410-2634, macro-obfuscated;
// statements below are cited from the de-obfuscated text, SURVEY.md reading aid):
//   * right keypoints are bucketed by round(pt.y) (band of 0 rows), in keypoint order;
//   * for every left (undistorted) keypoint the bucket of round(pt.y) is scanned in order: a candidate is skipped if it lies to the
//     right of the left keypoint (pt.x > left pt.x) or more than one octave apart; MapPoint::getDescDistance
//     (src/map_types/mappoint.h:146-162,172-177: popcount over 4 x 64 bits, returned as float) must be < Params::maxDescDistance and
//     strictly below the best so far -> the FIRST candidate of least distance;
//   * both rounded keypoints must lie 3 px inside their images; the sum of absolute differences of the 6 x 6 patches
//     (cv::Range(c-3, c+3) is end-exclusive) is evaluated for horizontal offsets -7..7 clipped to the right image, first minimum wins;
//   * an interior minimum is refined by the parabola through its neighbours (double arithmetic) and
//     depth = (bl*fx) / (x_left - (x_right + offset)), narrowed to float; everything else keeps depth 0.
//
// Device form: one warp per left keypoint.  The right keypoints are flattened once into 16-byte (row, x, octave) records that the
// CTA's eight warps stream through shared memory in tiles, lanes test 32 candidates at a time, the first-minimum is a 64-bit
// (distance, index) minimum over the warp; the 15 SAD offsets run on 15 lanes and the parabola on lane 0.  The work per pair is a
// few hundred thousand integer operations on L2-resident data (4000 x 4000 row tests, ~5 Hamming distances and 15 x 36 byte
// differences per keypoint): the kernel is launch / latency bound by construction and exists so that a stereo frame never leaves
// the device between extraction and tracking.
#include "common.cuh"
#include <cstring>

namespace {

constexpr int ST_THREADS = 256;
constexpr int ST_TILE = 2048;

struct StereoArgs {
    const uint8_t* img_l; size_t pitch_l;
    const uint8_t* img_r; size_t pitch_r;
    int w, h;
    const uco_keypoint* kl; const uint32_t* dl; int nl;   // descriptors: dense 32-byte rows, 4-byte aligned
    const uco_keypoint* kr; const uint32_t* dr; int nr;
    float max_desc_dist, bl, fx;
    float* depth; int32_t* match; int32_t* counters;   // counters[0] = keypoints with depth, [1] = error flag
    int4* rsoa;
};

__global__ void __launch_bounds__(256) stereo_flatten_kernel(const uco_keypoint* __restrict__ kr, int nr, int h, int4* __restrict__ rsoa) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nr) return;
    const uco_keypoint k = kr[j];
    const double r = round((double)k.y);                       // int(std::round(y - 0.0)) on a double
    const int row = (r >= 0.0 && r <= (double)(h - 1)) ? (int)r : -1;   // outside [0, rows-1] the bucket range is empty
    rsoa[j] = make_int4(row, __float_as_int(k.x), k.octave, 0);
}

__global__ void __launch_bounds__(ST_THREADS) stereo_assoc_kernel(const __grid_constant__ StereoArgs A) {
    __shared__ int4 tile[ST_TILE];
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (ST_THREADS / 32) + (threadIdx.x >> 5);
    const bool active = i < A.nl;
    uco_keypoint kl;
    uint32_t dl[8];
    int yl = -2;
    if (active) {
        kl = A.kl[i];
        yl = (int)roundf(kl.y);                                // int _ = std::round(float)
#pragma unroll
        for (int k = 0; k < 8; k++) dl[k] = A.dl[(size_t)i * 8 + k];
    }
    unsigned long long best = ~0ull;
    for (int t0 = 0; t0 < A.nr; t0 += ST_TILE) {
        const int nt = min(ST_TILE, A.nr - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < nt; e += ST_THREADS) tile[e] = A.rsoa[t0 + e];
        __syncthreads();
        if (!active || yl < 0 || yl >= A.h) continue;
        for (int e = lane; e < nt; e += 32) {
            const int4 c = tile[e];
            if (c.x != yl || __int_as_float(c.y) > kl.x || abs(c.z - kl.octave) > 1) continue;
            const uint32_t* dr = A.dr + (size_t)(t0 + e) * 8;
            unsigned d = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) d += __popc(dl[k] ^ dr[k]);
            if ((float)d < A.max_desc_dist) best = min(best, ((unsigned long long)d << 32) | (unsigned)(t0 + e));
        }
    }
    if (!active) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = min(best, other);
    }
    float depth = 0.f;
    int matched = -1;
    if (best != ~0ull) {
        const int j = (int)(best & 0xffffffffu);
        matched = j;
        const uco_keypoint kr = A.kr[j];
        const int hw = 3, L = 7;
        const int xl = (int)roundf(kl.x), yl2 = (int)roundf(kl.y);
        const int xr = (int)roundf(kr.x), yr = (int)roundf(kr.y);
        const bool inside = !(xl < hw || xl + hw >= A.w) && !(yl2 < hw || yl2 + hw >= A.h) && !(xr < hw || xr + hw >= A.w) &&
                            !(yr < hw || yr + hw >= A.h);
        if (inside) {
            const int lo = max(-L, -xr), hi = min(L, A.w - 1 - xr);
            const int dx = lo + lane;
            unsigned key = 0xffffffffu;
            unsigned sad = 0;
            if (dx <= hi) {
                const int xc = xr + dx;
                if (xc - hw < 0 || xc + hw > A.w) {
                    A.counters[1] = 1;   // the reference's cv::Mat ROI constructor throws here (never with ORB keypoints: >= 16 px inside)
                } else {
                    for (int r = -hw; r < hw; r++) {
                        const uint8_t* pl = A.img_l + (size_t)(yl2 + r) * A.pitch_l + (xl - hw);
                        const uint8_t* pr = A.img_r + (size_t)(yr + r) * A.pitch_r + (xc - hw);
#pragma unroll
                        for (int c = 0; c < 2 * hw; c++) sad += (unsigned)abs((int)pl[c] - (int)pr[c]);
                    }
                    key = (sad << 8) | (unsigned)(dx + L);
                }
            }
            unsigned kmin = key;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            const int b = (int)(kmin & 0xffu);   // index of the first minimum in [0, 2L]
            const int bl_lane = b - L - lo;      // the lane that evaluated it
            const unsigned s1 = __shfl_sync(0xffffffffu, sad, (bl_lane - 1) & 31);
            const unsigned s2 = __shfl_sync(0xffffffffu, sad, bl_lane & 31);
            const unsigned s3 = __shfl_sync(0xffffffffu, sad, (bl_lane + 1) & 31);
            if (kmin != 0xffffffffu && b > lo + L && b < hi + L) {
                const double d1 = (double)s1, d2 = (double)s2, d3 = (double)s3;
                const double off = 0.5 * (d1 - d3) / (d1 + d3 - 2 * d2) + b - L;
                const double xs = (double)kr.x + off;
                depth = (float)((double)(A.bl * A.fx) / ((double)kl.x - xs));
                if (lane == 0) atomicAdd(A.counters, 1);
            }
        }
    }
    if (lane == 0) {
        A.depth[i] = depth;
        if (A.match) A.match[i] = matched;
    }
}

}  // namespace

extern "C" {

int uco_b200_stereo_depth_dev(uco_b200_ctx* ctx, const uint8_t* img_l_dev, size_t pitch_l, const uint8_t* img_r_dev, size_t pitch_r,
                              int w, int h, const uco_keypoint* kps_l_dev, const uint8_t* desc_l_dev, int n_l,
                              const uco_keypoint* kps_r_dev, const uint8_t* desc_r_dev, int n_r, float max_desc_dist, float bl, float fx,
                              float* depth_dev, int32_t* match_dev, int32_t* counters_dev) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (w <= 0 || h <= 0 || n_l < 0 || n_r < 0) return uco_fail(ctx, UCO_E_INVALID, "stereo_depth: bad size");
    if (!img_l_dev || !img_r_dev || !depth_dev || !counters_dev || (n_l && (!kps_l_dev || !desc_l_dev)) ||
        (n_r && (!kps_r_dev || !desc_r_dev)))
        return uco_fail(ctx, UCO_E_INVALID, "stereo_depth: null pointer");
    if (((uintptr_t)desc_l_dev | (uintptr_t)desc_r_dev) & 3) return uco_fail(ctx, UCO_E_INVALID, "stereo_depth: descriptors must be 4-byte aligned");
    if (pitch_l < (size_t)w || pitch_r < (size_t)w) return uco_fail(ctx, UCO_E_INVALID, "stereo_depth: pitch below width");
    UCO_CUDA(ctx, cudaMemsetAsync(counters_dev, 0, 2 * sizeof(int32_t), ctx->stream));
    if (n_l == 0) return UCO_OK;
    int4* rsoa = (int4*)uco_ws(ctx, WS_STEREO_SOA, (size_t)(n_r > 0 ? n_r : 1) * sizeof(int4));
    if (!rsoa) return UCO_E_NOMEM;
    if (n_r) {
        stereo_flatten_kernel<<<(n_r + 255) / 256, 256, 0, ctx->stream>>>(kps_r_dev, n_r, h, rsoa);
        UCO_LAUNCH_CHECK(ctx);
    }
    StereoArgs A{img_l_dev, pitch_l, img_r_dev, pitch_r, w, h, kps_l_dev, (const uint32_t*)desc_l_dev, n_l, kps_r_dev,
                 (const uint32_t*)desc_r_dev, n_r, max_desc_dist, bl, fx, depth_dev, match_dev, counters_dev, rsoa};
    const int per_block = ST_THREADS / 32;
    stereo_assoc_kernel<<<(n_l + per_block - 1) / per_block, ST_THREADS, 0, ctx->stream>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

int uco_b200_stereo_depth(uco_b200_ctx* ctx, const uint8_t* img_l, size_t stride_l, const uint8_t* img_r, size_t stride_r, int w, int h,
                          const uco_keypoint* kps_l, const uint8_t* desc_l, size_t desc_l_stride, int n_l, const uco_keypoint* kps_r,
                          const uint8_t* desc_r, size_t desc_r_stride, int n_r, float max_desc_dist, float bl, float fx, float* depth,
                          int32_t* match_r, int* n_with_depth) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (w <= 0 || h <= 0 || n_l < 0 || n_r < 0) return uco_fail(ctx, UCO_E_INVALID, "stereo_depth: bad size");
    if (!img_l || !img_r || (n_l && (!kps_l || !desc_l || !depth)) || (n_r && (!kps_r || !desc_r)))
        return uco_fail(ctx, UCO_E_INVALID, "stereo_depth: null pointer");
    if (stride_l < (size_t)w || stride_r < (size_t)w || (n_l && desc_l_stride < 32) || (n_r && desc_r_stride < 32))
        return uco_fail(ctx, UCO_E_INVALID, "stereo_depth: stride too small");
    if (n_with_depth) *n_with_depth = 0;
    if (n_l == 0) return UCO_OK;
    const size_t pitch = ((size_t)w + 63) & ~(size_t)63;
    const size_t img_bytes = pitch * h;
    const size_t kl_b = (size_t)n_l * sizeof(uco_keypoint), kr_b = (size_t)n_r * sizeof(uco_keypoint);
    const size_t off_kl = 2 * img_bytes, off_kr = off_kl + ((kl_b + 15) & ~(size_t)15), off_dl = off_kr + ((kr_b + 15) & ~(size_t)15);
    const size_t off_dr = off_dl + (size_t)n_l * 32, total = off_dr + (size_t)n_r * 32;
    uint8_t* d_in = (uint8_t*)uco_ws(ctx, WS_STEREO_IN, total);
    const size_t out_bytes = 16 + (size_t)n_l * 8;
    uint8_t* d_out = (uint8_t*)uco_ws(ctx, WS_STEREO_OUT, out_bytes);
    uint8_t* h_out = (uint8_t*)uco_pinned(ctx, WS_STEREO_OUT, out_bytes);
    if (!d_in || !d_out || !h_out) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemcpy2DAsync(d_in, pitch, img_l, stride_l, w, h, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpy2DAsync(d_in + img_bytes, pitch, img_r, stride_r, w, h, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(d_in + off_kl, kps_l, kl_b, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpy2DAsync(d_in + off_dl, 32, desc_l, desc_l_stride, 32, n_l, cudaMemcpyHostToDevice, ctx->stream));
    if (n_r) {
        UCO_CUDA(ctx, cudaMemcpyAsync(d_in + off_kr, kps_r, kr_b, cudaMemcpyHostToDevice, ctx->stream));
        UCO_CUDA(ctx, cudaMemcpy2DAsync(d_in + off_dr, 32, desc_r, desc_r_stride, 32, n_r, cudaMemcpyHostToDevice, ctx->stream));
    }
    int32_t* d_cnt = (int32_t*)d_out;
    float* d_depth = (float*)(d_out + 16);
    int32_t* d_match = (int32_t*)(d_out + 16 + (size_t)n_l * 4);
    int rc = uco_b200_stereo_depth_dev(ctx, d_in, pitch, d_in + img_bytes, pitch, w, h, (const uco_keypoint*)(d_in + off_kl), d_in + off_dl,
                                       n_l, (const uco_keypoint*)(d_in + off_kr), d_in + off_dr, n_r, max_desc_dist, bl, fx, d_depth,
                                       d_match, d_cnt);
    if (rc != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (((int32_t*)h_out)[1])
        return uco_fail(ctx, UCO_E_INVALID, "stereo_depth: a matched right keypoint lies within 10 px of the image border "
                                            "(the reference's cv::Mat ROI throws there)");
    memcpy(depth, h_out + 16, (size_t)n_l * 4);
    if (match_r) memcpy(match_r, h_out + 16 + (size_t)n_l * 4, (size_t)n_l * 4);
    if (n_with_depth) *n_with_depth = ((int32_t*)h_out)[0];
    return UCO_OK;
}

}  // extern "C"
