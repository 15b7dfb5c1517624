// triangulate.cu — K13: two-view triangulation of matched keypoints with the reference's acceptance gates (the
// arithmetic core of new-map-point creation).
//
// Replaces ucoslam::Triangulate(Train, Query, RT_Q2T, matches, maxChi2) (reference: src/basictypes/misc.cpp:921-1040; the twin
// triangulate_ :1042-1160 shares the body), called from the mapper's new-point creation (src/utils/mapmanager.cpp:10093) and the map
// initialiser (src/utils/mapinitializer.cpp:1574):
//   per match (trainIdx -> keypoint 1 in camera 1 = K1[I|0], queryIdx -> keypoint 2 in camera 2 = K2[R|t]):
//     rays through the two pixels, cosine of their angle; rejected if < 0 or > 0.9998 (:986-990);
//     homogeneous DLT: the 4x4 system of misc.cpp:923-929, null vector = last row of V^T of its SVD (:931-933), w == 0 rejects;
//     finite, in front of both cameras (:996-1009); reprojection chi2 in both images, scaled by 1/scaleFactor[octave]^2, <= maxChi2;
//   a rejected match yields (NaN, NaN, NaN).
// Optionally followed by the mapper's scale-consistency test on the survivors and the transform to global coordinates
// (src/utils/mapmanager.cpp:9772-10788, de-obfuscated: the statements after the Triangulate call of new-map-point creation).
//
// The reference runs OpenCV's float SVD per match (LAPACK sgesdd or OpenCV's Jacobi, depending on the build: not reproducible bit for
// bit across builds).  Here one thread per match forms the same float system, converts it to double and takes the eigenvector of the
// smallest eigenvalue of A^T A by cyclic Jacobi (4x4 symmetric, double: the squared condition number stays far below 1/eps for float
// data), then applies the gates in the reference's float arithmetic.  Parity is therefore a stated float tolerance (tests), not
// bit-exactness; the work is a few hundred flops per match on L2-resident inputs: latency bound.
#include "common.cuh"
#include <cmath>
#include <cstring>

namespace {

struct TriArgs {
    const uco_keypoint* kp1; int n1;
    const uco_keypoint* kp2; int n2;
    const uco_match* matches; int n;
    float K1[4], K2[4];       // fx fy cx cy
    float R[9], t[3];         // camera 1 -> camera 2
    const float* inv_sf1; int nl1;   // 1 / scaleFactor^2 per octave
    const float* inv_sf2; int nl2;
    const float* sf1; const float* sf2;   // the scale factors themselves
    float max_chi2, ratio_factor; int to_global; float G[12];
    float* xyz; int32_t* counters;   // [0] accepted, [1] bad index flag
};

__device__ void jacobi_smallest_eigvec4(double S[4][4], double v[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 12; sweep++) {
        double off = 0;
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) off += S[p][q] * S[p][q];
        const double diag = S[0][0] * S[0][0] + S[1][1] * S[1][1] + S[2][2] * S[2][2] + S[3][3] * S[3][3];
        if (off <= 1e-32 * diag) break;
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                if (S[p][q] == 0.0) continue;
                const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
#pragma unroll
                for (int k = 0; k < 4; k++) {   // S <- S J
                    const double a = S[k][p], b = S[k][q];
                    S[k][p] = c * a - s * b;
                    S[k][q] = s * a + c * b;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {   // S <- J^T S
                    const double a = S[p][k], b = S[q][k];
                    S[p][k] = c * a - s * b;
                    S[q][k] = s * a + c * b;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double a = V[k][p], b = V[k][q];
                    V[k][p] = c * a - s * b;
                    V[k][q] = s * a + c * b;
                }
            }
    }
    int m = 0;
#pragma unroll
    for (int k = 1; k < 4; k++)
        if (S[k][k] < S[m][m]) m = k;
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (m == 0) ? V[k][0] : (m == 1) ? V[k][1] : (m == 2) ? V[k][2] : V[k][3];
}

// one match: the gates of misc.cpp:981-1040 (+ the mapper's scale consistency / global frame); writes out[3] (NaN = rejected)
struct TriCam {
    float K1[4], K2[4];       // fx fy cx cy
    float R[9], t[3];         // camera 1 -> camera 2
};
struct TriGates {
    const float* inv_sf1; const float* inv_sf2; const float* sf1; const float* sf2;
    float max_chi2, ratio_factor; int to_global; float G[12];
};
__device__ __forceinline__ bool tri_one(const uco_keypoint& k1, const uco_keypoint& k2, const TriCam& A, const TriGates& T, float out[3]) {
    bool ok = true;
    {
        {
            const float fx1 = A.K1[0], fy1 = A.K1[1], cx1 = A.K1[2], cy1 = A.K1[3];
            const float fx2 = A.K2[0], fy2 = A.K2[1], cx2 = A.K2[2], cy2 = A.K2[3];
            const float* R = A.R;
            // parallax between the rays (misc.cpp:981-990)
            const float a1[3] = {(k1.x - cx1) * (1.f / fx1), (k1.y - cy1) * (1.f / fy1), 1.f};
            const float a2[3] = {(k2.x - cx2) * (1.f / fx2), (k2.y - cy2) * (1.f / fy2), 1.f};
            const float s1 = (float)(1.0 / sqrt((double)a1[0] * a1[0] + (double)a1[1] * a1[1] + 1.0));
            const float s2 = (float)(1.0 / sqrt((double)a2[0] * a2[0] + (double)a2[1] * a2[1] + 1.0));
            const float r1[3] = {a1[0] * s1, a1[1] * s1, a1[2] * s1};
            const float u2[3] = {a2[0] * s2, a2[1] * s2, a2[2] * s2};
            float r2[3];   // R^T * u2
#pragma unroll
            for (int c = 0; c < 3; c++) r2[c] = (float)((double)R[c] * u2[0] + (double)R[3 + c] * u2[1] + (double)R[6 + c] * u2[2]);
            const double cosp = (double)r1[0] * r2[0] + (double)r1[1] * r2[1] + (double)r1[2] * r2[2];
            ok = !(cosp < 0 || cosp > 0.9998);
            if (ok) {
                // P1 = K1 [I|0], P2 = K2 [R|t] in float; rows of A as misc.cpp:923-929
                float P1[3][4] = {{fx1, 0, cx1, 0}, {0, fy1, cy1, 0}, {0, 0, 1, 0}};
                float P2[3][4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float e0 = c < 3 ? R[c] : A.t[0], e1 = c < 3 ? R[3 + c] : A.t[1], e2 = c < 3 ? R[6 + c] : A.t[2];
                    P2[0][c] = (float)((double)fx2 * e0 + (double)cx2 * e2);
                    P2[1][c] = (float)((double)fy2 * e1 + (double)cy2 * e2);
                    P2[2][c] = e2;
                }
                double Am[4][4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    Am[0][c] = (double)(k1.x * P1[2][c] - P1[0][c]);
                    Am[1][c] = (double)(k1.y * P1[2][c] - P1[1][c]);
                    Am[2][c] = (double)(k2.x * P2[2][c] - P2[0][c]);
                    Am[3][c] = (double)(k2.y * P2[2][c] - P2[1][c]);
                }
                double S[4][4], v[4];
#pragma unroll
                for (int p = 0; p < 4; p++)
#pragma unroll
                    for (int q = 0; q < 4; q++) S[p][q] = Am[0][p] * Am[0][q] + Am[1][p] * Am[1][q] + Am[2][p] * Am[2][q] + Am[3][p] * Am[3][q];
                jacobi_smallest_eigvec4(S, v);
                const float w = (float)v[3];
                ok = w != 0.f;
                if (ok) {
                    const float X = (float)v[0] / w, Y = (float)v[1] / w, Z = (float)v[2] / w;   // x3D.rowRange(0,3) / x3D(3), float
                    ok = isfinite(X) && isfinite(Y) && isfinite(Z) && !(Z <= 0);
                    if (ok) {
                        const float X2 = (float)((double)R[0] * X + (double)R[1] * Y + (double)R[2] * Z) + A.t[0];
                        const float Y2 = (float)((double)R[3] * X + (double)R[4] * Y + (double)R[5] * Z) + A.t[1];
                        const float Z2 = (float)((double)R[6] * X + (double)R[7] * Y + (double)R[8] * Z) + A.t[2];
                        ok = !(Z2 <= 0);
                        if (ok) {
                            const float iz1 = 1.f / Z;
                            const float px = fx1 * X * iz1 + cx1, py = fy1 * Y * iz1 + cy1;
                            const float chi1 = T.inv_sf1[k1.octave] * ((px - k1.x) * (px - k1.x) + (py - k1.y) * (py - k1.y));
                            ok = !(chi1 > T.max_chi2);
                            if (ok) {
                                const float iz2 = 1.f / Z2;
                                const float qx = fx2 * X2 * iz2 + cx2, qy = fy2 * Y2 * iz2 + cy2;
                                const float chi2 = T.inv_sf2[k2.octave] * ((qx - k2.x) * (qx - k2.x) + (qy - k2.y) * (qy - k2.y));
                                ok = !(chi2 > T.max_chi2);
                                if (ok && T.ratio_factor != 0.f) {   // mapper's scale consistency (distances are pose invariant)
                                    const float d1 = (float)sqrt((double)X * X + (double)Y * Y + (double)Z * Z);
                                    const float d2 = (float)sqrt((double)X2 * X2 + (double)Y2 * Y2 + (double)Z2 * Z2);
                                    ok = !(d1 == 0.f || d2 == 0.f);
                                    if (ok) {
                                        const float rd = d1 / d2, ro = T.sf1[k1.octave] / T.sf2[k2.octave];
                                        ok = !(rd * T.ratio_factor < ro || rd > ro * T.ratio_factor);
                                    }
                                }
                                if (ok) {
                                    if (T.to_global) {               // Se3Transform::operator*(Point3f)
                                        out[0] = T.G[0] * X + T.G[1] * Y + T.G[2] * Z + T.G[3];
                                        out[1] = T.G[4] * X + T.G[5] * Y + T.G[6] * Z + T.G[7];
                                        out[2] = T.G[8] * X + T.G[9] * Y + T.G[10] * Z + T.G[11];
                                    } else {
                                        out[0] = X; out[1] = Y; out[2] = Z;
                                    }
                                    
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    return ok;
}

__global__ void __launch_bounds__(128) triangulate_kernel(const __grid_constant__ TriArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const float nanv = __int_as_float(0x7fc00000);
    float out[3] = {nanv, nanv, nanv};
    const uco_match m = A.matches[i];
    bool ok = m.trainIdx >= 0 && m.trainIdx < A.n1 && m.queryIdx >= 0 && m.queryIdx < A.n2;
    if (!ok) A.counters[1] = 1;
    if (ok) {
        const uco_keypoint k1 = A.kp1[m.trainIdx], k2 = A.kp2[m.queryIdx];
        ok = k1.octave >= 0 && k1.octave < A.nl1 && k2.octave >= 0 && k2.octave < A.nl2;
        if (!ok) A.counters[1] = 1;
        if (ok) {
            TriCam C;
            TriGates T;
#pragma unroll
            for (int q = 0; q < 4; q++) { C.K1[q] = A.K1[q]; C.K2[q] = A.K2[q]; }
#pragma unroll
            for (int q = 0; q < 9; q++) C.R[q] = A.R[q];
#pragma unroll
            for (int q = 0; q < 3; q++) C.t[q] = A.t[q];
            T.inv_sf1 = A.inv_sf1; T.inv_sf2 = A.inv_sf2; T.sf1 = A.sf1; T.sf2 = A.sf2;
            T.max_chi2 = A.max_chi2; T.ratio_factor = A.ratio_factor; T.to_global = A.to_global;
#pragma unroll
            for (int q = 0; q < 12; q++) T.G[q] = A.G[q];
            if (tri_one(k1, k2, C, T, out)) atomicAdd(A.counters, 1);
        }
    }
    A.xyz[3 * (size_t)i] = out[0];
    A.xyz[3 * (size_t)i + 1] = out[1];
    A.xyz[3 * (size_t)i + 2] = out[2];
}

// new-map-point creation: all (keyframe, neighbour f) pairs of one keyframe in one launch, reading the matcher's device output.
// blockIdx.y = neighbour; matches of neighbour f at matches[f * match_stride ..], n_matches[f] of them; the keyframe is camera 1.
struct TriPairsArgs {
    const uco_keypoint* kp1; int n1;                  // the keyframe's keypoints
    const uco_keypoint* kp2; int kp2_stride; const int32_t* n2;   // neighbour f: kp2 + f * kp2_stride, n2[f] keypoints
    const uco_match* matches; int match_stride; const int32_t* n_matches;
    float K1[4];
    const float* cam2;                                // per neighbour: K2 (4), R (9), t (3) = 16 floats
    TriGates T; int nl1, nl2;
    float* xyz; int32_t* counters;                    // xyz like matches; counters[0] = bad index flag, counters[1 + f] = accepted of f
};
__global__ void __launch_bounds__(128) triangulate_pairs_kernel(const __grid_constant__ TriPairsArgs A) {
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_matches[f]) return;
    const float nanv = __int_as_float(0x7fc00000);
    float out[3] = {nanv, nanv, nanv};
    const size_t at = (size_t)f * A.match_stride + i;
    const uco_match m = A.matches[at];
    bool ok = m.trainIdx >= 0 && m.trainIdx < A.n1 && m.queryIdx >= 0 && m.queryIdx < A.n2[f];
    if (!ok) A.counters[0] = 1;
    if (ok) {
        const uco_keypoint k1 = A.kp1[m.trainIdx], k2 = A.kp2[(size_t)f * A.kp2_stride + m.queryIdx];
        ok = k1.octave >= 0 && k1.octave < A.nl1 && k2.octave >= 0 && k2.octave < A.nl2;
        if (!ok) A.counters[0] = 1;
        if (ok) {
            TriCam C;
            const float* c2 = A.cam2 + 16 * (size_t)f;
#pragma unroll
            for (int q = 0; q < 4; q++) { C.K1[q] = A.K1[q]; C.K2[q] = c2[q]; }
#pragma unroll
            for (int q = 0; q < 9; q++) C.R[q] = c2[4 + q];
#pragma unroll
            for (int q = 0; q < 3; q++) C.t[q] = c2[13 + q];
            if (tri_one(k1, k2, C, A.T, out)) atomicAdd(A.counters + 1 + f, 1);
        }
    }
    A.xyz[3 * at] = out[0];
    A.xyz[3 * at + 1] = out[1];
    A.xyz[3 * at + 2] = out[2];
}

}  // namespace

// internal (match.cu: uco_b200_new_points): every pointer is a DEVICE pointer; sf_dev = 4 x UCO_MATCH_MAX_SCALES floats
// (1/sf1^2 | 1/sf2^2 | sf1 | sf2); cam2_dev = n_frames x 16; counters_dev = 1 + n_frames ints (zeroed here)
int uco_tri_pairs_launch(uco_b200_ctx* ctx, const uco_keypoint* kp1, int n1, const uco_keypoint* kp2, int kp2_stride, const int32_t* n2_dev, int n_frames,
                         const uco_match* matches, int match_stride, const int32_t* n_matches_dev, const float* K1, const float* cam2_dev,
                         const float* sf_dev, int nl1, int nl2, float max_chi2, float ratio_factor, const float* g2f_train, float* xyz_dev,
                         int32_t* counters_dev) {
    TriPairsArgs A;
    A.kp1 = kp1; A.n1 = n1; A.kp2 = kp2; A.kp2_stride = kp2_stride; A.n2 = n2_dev;
    A.matches = matches; A.match_stride = match_stride; A.n_matches = n_matches_dev;
    memcpy(A.K1, K1, sizeof A.K1);
    A.cam2 = cam2_dev;
    A.T.inv_sf1 = sf_dev; A.T.inv_sf2 = sf_dev + UCO_MATCH_MAX_SCALES; A.T.sf1 = sf_dev + 2 * UCO_MATCH_MAX_SCALES; A.T.sf2 = sf_dev + 3 * UCO_MATCH_MAX_SCALES;
    A.T.max_chi2 = max_chi2; A.T.ratio_factor = ratio_factor; A.T.to_global = g2f_train != nullptr;
    for (int q = 0; q < 12; q++) A.T.G[q] = g2f_train ? g2f_train[q] : 0.f;
    A.nl1 = nl1; A.nl2 = nl2;
    A.xyz = xyz_dev; A.counters = counters_dev;
    UCO_CUDA(ctx, cudaMemsetAsync(counters_dev, 0, 4 * (size_t)(1 + n_frames), ctx->stream));
    triangulate_pairs_kernel<<<dim3((match_stride + 127) / 128, n_frames), 128, 0, ctx->stream>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

extern "C" {

int uco_b200_triangulate(uco_b200_ctx* ctx, const uco_keypoint* kps_train, int n_train, const uco_keypoint* kps_query, int n_query,
                         const uco_match* matches, int n_matches, const uco_triangulate_params* prm, float* xyz, int* n_good) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (n_train < 0 || n_query < 0 || n_matches < 0 || !prm) return uco_fail(ctx, UCO_E_INVALID, "triangulate: bad argument");
    if (n_good) *n_good = 0;
    if (n_matches == 0) return UCO_OK;
    if (!kps_train || !kps_query || !matches || !xyz) return uco_fail(ctx, UCO_E_INVALID, "triangulate: null pointer");
    if (prm->n_levels_train <= 0 || prm->n_levels_query <= 0 || prm->n_levels_train > UCO_MATCH_MAX_SCALES ||
        prm->n_levels_query > UCO_MATCH_MAX_SCALES)
        return uco_fail(ctx, UCO_E_INVALID, "triangulate: scale factor tables of 1..%d levels expected", UCO_MATCH_MAX_SCALES);
    const size_t b1 = (size_t)n_train * sizeof(uco_keypoint), b2 = (size_t)n_query * sizeof(uco_keypoint);
    const size_t bm = (size_t)n_matches * sizeof(uco_match);
    const size_t o2 = (b1 + 15) & ~(size_t)15, om = o2 + ((b2 + 15) & ~(size_t)15), os = om + ((bm + 15) & ~(size_t)15);
    const size_t total = os + 4 * UCO_MATCH_MAX_SCALES * sizeof(float);
    uint8_t* h_in = (uint8_t*)uco_pinned(ctx, WS_TRI_IN, total);
    uint8_t* d_in = (uint8_t*)uco_ws(ctx, WS_TRI_IN, total);
    const size_t out_bytes = 16 + (size_t)n_matches * 12;
    uint8_t* d_out = (uint8_t*)uco_ws(ctx, WS_TRI_OUT, out_bytes);
    uint8_t* h_out = (uint8_t*)uco_pinned(ctx, WS_TRI_OUT, out_bytes);
    if (!h_in || !d_in || !d_out || !h_out) return UCO_E_NOMEM;
    memcpy(h_in, kps_train, b1);
    memcpy(h_in + o2, kps_query, b2);
    memcpy(h_in + om, matches, bm);
    float* sf = (float*)(h_in + os);
    // invScaleFactors: 1.f/(f*f) per octave (misc.cpp:943-945)
    for (int l = 0; l < prm->n_levels_train; l++) sf[l] = 1.f / (prm->scale_factors_train[l] * prm->scale_factors_train[l]);
    for (int l = 0; l < prm->n_levels_query; l++)
        sf[UCO_MATCH_MAX_SCALES + l] = 1.f / (prm->scale_factors_query[l] * prm->scale_factors_query[l]);
    for (int l = 0; l < prm->n_levels_train; l++) sf[2 * UCO_MATCH_MAX_SCALES + l] = prm->scale_factors_train[l];
    for (int l = 0; l < prm->n_levels_query; l++) sf[3 * UCO_MATCH_MAX_SCALES + l] = prm->scale_factors_query[l];
    UCO_CUDA(ctx, cudaMemcpyAsync(d_in, h_in, total, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemsetAsync(d_out, 0, 16, ctx->stream));
    TriArgs A;
    A.ratio_factor = prm->scale_ratio_factor;
    A.to_global = prm->to_global;
    memcpy(A.G, prm->g2f_train, sizeof A.G);
    A.kp1 = (const uco_keypoint*)d_in; A.n1 = n_train;
    A.kp2 = (const uco_keypoint*)(d_in + o2); A.n2 = n_query;
    A.matches = (const uco_match*)(d_in + om); A.n = n_matches;
    memcpy(A.K1, prm->K_train, sizeof A.K1);
    memcpy(A.K2, prm->K_query, sizeof A.K2);
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) A.R[3 * r + c] = prm->RT[4 * r + c];
        A.t[r] = prm->RT[4 * r + 3];
    }
    A.inv_sf1 = (const float*)(d_in + os); A.nl1 = prm->n_levels_train;
    A.inv_sf2 = A.inv_sf1 + UCO_MATCH_MAX_SCALES; A.nl2 = prm->n_levels_query;
    A.sf1 = A.inv_sf1 + 2 * UCO_MATCH_MAX_SCALES; A.sf2 = A.inv_sf1 + 3 * UCO_MATCH_MAX_SCALES;
    A.max_chi2 = prm->max_chi2;
    A.counters = (int32_t*)d_out;
    A.xyz = (float*)(d_out + 16);
    triangulate_kernel<<<(n_matches + 127) / 128, 128, 0, ctx->stream>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    UCO_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (((int32_t*)h_out)[1]) return uco_fail(ctx, UCO_E_INVALID, "triangulate: a match refers to a keypoint or an octave out of range");
    memcpy(xyz, h_out + 16, (size_t)n_matches * 12);
    if (n_good) *n_good = ((int32_t*)h_out)[0];
    return UCO_OK;
}

}  // extern "C"
