// match.cu — K8: frame-to-frame descriptor matcher = exact Hamming k-NN (knn.cu) + the post-filters of
// FrameMatcher_Flann::matchEpipolar (src/utils/framematcher.cpp:228-322), all on the device; and its BoW-guided twin
// FrameMatcher_BoW::matchEpipolar (:407-541), whose candidates are the keypoints sharing a level-3 vocabulary node.
//
// One thread block per frame pair.  The reference's sequential passes are restated as order-independent operations with the
// same result:
//   * per query (framematcher.cpp:246-283): the k candidates are scanned in the order the index returned them (the heap
//     order knn.cu reproduces exactly) — one thread per query;
//   * filter_ambiguous_train (misc.cpp:153-185) keeps, per train keypoint, the match of least distance and among equals the
//     one met first: that is the minimum of (distance, query position), taken with a 64-bit atomicMin — integer, so
//     independent of thread order;
//   * the rotation histogram (:288-316) needs only bin COUNTS (integer atomics) before the three maxima are chosen;
//   * remove_unused_matches is a stable compaction: block-wide prefix sums over query order.
#include "common.cuh"
#include <map>
#include <vector>
#include <algorithm>
#include <cmath>
#include <float.h>
#include <string.h>
#include <vector>

int uco_knn_launch_internal(uco_b200_ctx* ctx, const uint8_t* q_dev, int nq, const uint8_t* t_dev, int nt, int k, int order,
                            int32_t* idx_dev, int32_t* dist_dev, int n_pairs, const int* nq_dev, const int* nt_dev,
                            size_t q_stride, size_t t_stride);

namespace {

constexpr int MT = 1024;  // threads per block
constexpr int NN = 10;    // framematcher.cpp:122
constexpr int NBINS = 30;

struct MatchArgs {
    const int32_t* knn_idx;   // n_pairs x nq_max x NN
    const int32_t* knn_dist;
    const uco_keypoint* q_kps; size_t q_kps_stride;
    const uco_keypoint* t_kps; size_t t_kps_stride;
    const int32_t* q_map;     // optional: descriptor row -> keypoint index (pair p's map at q_map + p * q_map_stride)
    const int32_t* t_map;     // optional, shared by all pairs
    size_t q_map_stride;
    const int32_t *q_sel, *t_sel;   // optional frame selection of pair p (kps strides and nq_dev / nt_dev are then per FRAME)
    const float* f12_pair;    // optional: 9 floats per pair (matchEpipolar against several frames); null -> prm.f12
    int nq_max, nt_max, n_t_kps;   // n_t_kps: size of the `used` table of a pair
    const int32_t* nq_dev; const int32_t* nt_dev;
    unsigned long long* used; // n_pairs x n_t_kps
    int2* cand;               // n_pairs x nq_max : (train keypoint or -1, distance)
    uco_match* out;           // n_pairs x nq_max
    int32_t* n_out;
    uco_match_params prm;
    // BoW-guided candidates (FrameMatcher_BoW): entry i = i-th keypoint reference of the query's fBow2 in node order; q_map = the
    // keypoint of an entry.  Null bow_entry_node selects the k-NN candidates above.
    const int32_t* bow_entry_node;   // nq_max: query node of the entry
    const int32_t* bow_t_node;       // per query node: index of the train node with the same id, or -1
    const int32_t* bow_t_ptr;        // train node -> range in bow_t_kp
    const int32_t* bow_t_kp;
    const uint8_t *q_usable, *t_usable;   // optional isUsed(frame, keypoint, mode)
    const uint32_t *q_desc, *t_desc;      // descriptors by KEYPOINT index, 8 words each
};

__device__ __forceinline__ float epipolar_sq_dist(const uco_keypoint& kp1, const uco_keypoint& kp2, const float* F) {  // misc.h:72-81
    const float a = kp1.x * F[0] + kp1.y * F[3] + F[6];
    const float b = kp1.x * F[1] + kp1.y * F[4] + F[7];
    const float den = a * a + b * b;
    if (den == 0) return FLT_MAX;
    const float c = kp1.x * F[2] + kp1.y * F[5] + F[8];
    const float num = a * kp2.x + b * kp2.y + c;
    return num * num / den;
}

__global__ void __launch_bounds__(MT) match_filter_kernel(MatchArgs A) {
    __shared__ int hist[NBINS];
    __shared__ int keep[3];
    __shared__ int warp_tot[MT / 32];
    __shared__ int running;
    const int pair = blockIdx.x, tid = threadIdx.x;
    const int qf = A.q_sel ? A.q_sel[pair] : pair, tf = A.t_sel ? A.t_sel[pair] : pair;
    const int nq = A.nq_dev ? min(A.nq_max, A.nq_dev[qf]) : A.nq_max;
    const int nt = A.nt_dev ? min(A.nt_max, A.nt_dev[tf]) : A.nt_max;
    const int32_t* kidx = A.knn_idx + (size_t)pair * A.nq_max * NN;
    const int32_t* kdist = A.knn_dist + (size_t)pair * A.nq_max * NN;
    const uco_keypoint* qk = A.q_kps + (size_t)qf * A.q_kps_stride;
    const uco_keypoint* tk = A.t_kps + (size_t)tf * A.t_kps_stride;
    unsigned long long* used = A.used + (size_t)pair * A.n_t_kps;
    int2* cand = A.cand + (size_t)pair * A.nq_max;
    uco_match* out = A.out + (size_t)pair * A.nq_max;
    const uco_match_params& P = A.prm;
    const int32_t* q_map = A.q_map ? A.q_map + (size_t)pair * A.q_map_stride : nullptr;
    const float* F12 = A.f12_pair ? A.f12_pair + 9 * (size_t)pair : P.f12;

    for (int t = tid; t < A.n_t_kps; t += MT) used[t] = ~0ull;
    if (tid < NBINS) hist[tid] = 0;
    if (tid == 0) running = 0;
    __syncthreads();
    // per query: best / second best among the k candidates (framematcher.cpp:246-283)
    for (int i = tid; i < nq && !A.bow_entry_node; i += MT) {
        float bestDist = P.min_desc_dist, bestDist2 = FLT_MAX;
        int bestTrain = -1, octaveBest2 = -1;
        const int queryIndex = q_map ? q_map[i] : i;
        const uco_keypoint q = qk[queryIndex];
        for (int j = 0; j < NN; j++) {
            const int ti = kidx[i * NN + j];
            if (ti < 0 || ti >= nt) continue;
            const float d = (float)kdist[i * NN + j];
            if (d > P.min_desc_dist) continue;
            if (d < bestDist2) {
                const int trainIndex = A.t_map ? A.t_map[ti] : ti;
                const uco_keypoint t = tk[trainIndex];
                if (abs(t.octave - q.octave) > P.max_octave_diff) continue;
                if (P.use_f12) {
                    const float sf = P.scale_factors[min(max(q.octave, 0), UCO_MATCH_MAX_SCALES - 1)];
                    if ((double)epipolar_sq_dist(t, q, F12) >= 3.84 * (double)(sf * sf)) continue;
                }
                if (d < bestDist) { bestDist = d; bestTrain = trainIndex; }
                else { bestDist2 = d; octaveBest2 = t.octave; }
            }
        }
        if (bestTrain != -1 && (octaveBest2 == q.octave && bestDist > bestDist2 * P.nn_match_ratio)) bestTrain = -1;
        cand[i] = make_int2(bestTrain, (int)bestDist);
        if (bestTrain != -1) atomicMin(&used[bestTrain], ((unsigned long long)(unsigned)(int)bestDist << 32) | (unsigned)i);
    }
    // FrameMatcher_BoW::matchEpipolar (framematcher.cpp:433-480): the candidates of a query keypoint are the train keypoints of the
    // same level-3 vocabulary node, in the node's list order; any candidate that is not a new best overwrites the runner-up
    for (int i = tid; i < nq && A.bow_entry_node; i += MT) {
        const int qidx = q_map[i];
        int bestTrain = -1;
        float bestDist = P.min_desc_dist, bestDist2 = FLT_MAX;
        int octaveBest2 = -1;
        const int tn = A.bow_t_node[A.bow_entry_node[i]];
        const uco_keypoint q = qk[qidx];
        if (tn >= 0 && (!A.q_usable || A.q_usable[qidx])) {
            uint32_t qd[8];
#pragma unroll
            for (int w = 0; w < 8; w++) qd[w] = A.q_desc[8 * (size_t)qidx + w];
            for (int j = A.bow_t_ptr[tn]; j < A.bow_t_ptr[tn + 1]; j++) {
                const int tidx = A.bow_t_kp[j];
                if (A.t_usable && !A.t_usable[tidx]) continue;
                const uco_keypoint t = tk[tidx];
                if (abs(t.octave - q.octave) > P.max_octave_diff) continue;
                if (P.use_f12) {
                    const float sf = P.scale_factors[min(max(q.octave, 0), UCO_MATCH_MAX_SCALES - 1)];
                    if ((double)epipolar_sq_dist(t, q, F12) >= 3.84 * (double)(sf * sf)) continue;
                }
                int pc = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) pc += __popc(qd[w] ^ A.t_desc[8 * (size_t)tidx + w]);
                const float d = (float)pc;
                if (d < bestDist) { bestDist = d; bestTrain = tidx; }
                else { bestDist2 = d; octaveBest2 = t.octave; }
            }
        }
        if (bestTrain != -1 && (octaveBest2 == q.octave && bestDist > bestDist2 * P.nn_match_ratio)) bestTrain = -1;
        cand[i] = make_int2(bestTrain, (int)bestDist);
        if (bestTrain != -1) atomicMin(&used[bestTrain], ((unsigned long long)(unsigned)(int)bestDist << 32) | (unsigned)i);
    }
    __syncthreads();
    // survivors of filter_ambiguous_train and their rotation bin (:292-304); cand.x < 0 marks a dropped query
    for (int i = tid; i < nq; i += MT) {
        int2 c = cand[i];
        if (c.x < 0) continue;
        if (used[c.x] != (((unsigned long long)(unsigned)c.y << 32) | (unsigned)i)) { cand[i].x = -1; continue; }
        if (P.check_orientation) {
            const int queryIndex = q_map ? q_map[i] : i;
            float rot = tk[c.x].angle - qk[queryIndex].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * (1.0f / (float)NBINS));
            if (bin == NBINS) bin = 0;
            if (bin < 0 || bin >= NBINS) { cand[i].x = -1; continue; }   // angles outside [0, 360] (the reference asserts): the match is dropped
            atomicAdd(&hist[bin], 1);
            cand[i].y = c.y | (bin << 16);  // distance <= 256 fits the low half
        }
    }
    __syncthreads();
    if (tid == 0) {  // ucoslam_FM___computeThreeMaxima, framematcher.cpp:67-108
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < NBINS; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
    }
    __syncthreads();
    // stable compaction in query order
    const int lane = tid & 31, warp = tid >> 5;
    for (int base = 0; base < nq; base += MT) {
        const int i = base + tid;
        bool flag = false;
        int2 c = make_int2(-1, 0);
        if (i < nq) {
            c = cand[i];
            flag = c.x >= 0;
            if (flag && P.check_orientation) {
                const int bin = c.y >> 16;
                flag = bin == keep[0] || bin == keep[1] || bin == keep[2];
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; w++) before += warp_tot[w];
        if (flag) {
            uco_match m;
            m.queryIdx = q_map ? q_map[i] : i;
            m.trainIdx = c.x;
            m.imgIdx = -1;
            m.distance = (float)(c.y & 0xffff);
            out[before + __popc(bal & ((1u << lane) - 1))] = m;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < MT / 32; w++) tot += warp_tot[w];
            running += tot;
        }
        __syncthreads();
    }
    if (tid == 0) A.n_out[pair] = running;
}

int check_params(uco_b200_ctx* ctx, const uco_match_params* prm) {
    if (!prm) return uco_fail(ctx, UCO_E_INVALID, "frame_match: no parameters");
    if (prm->n_scales < 0 || prm->n_scales > UCO_MATCH_MAX_SCALES) return uco_fail(ctx, UCO_E_INVALID, "frame_match: n_scales out of range");
    if (!(prm->min_desc_dist <= 65535.f)) return uco_fail(ctx, UCO_E_INVALID, "frame_match: min_desc_dist out of range");
    return UCO_OK;
}

}  // namespace

extern "C" {

int uco_b200_frame_match_batch_dev(uco_b200_ctx* ctx, int n_pairs, const uint8_t* q_desc_dev, size_t q_pair_stride,
                                   const uco_keypoint* q_kps_dev, size_t q_kps_pair_stride, int nq_max, const int32_t* nq_dev,
                                   const uint8_t* t_desc_dev, size_t t_pair_stride, const uco_keypoint* t_kps_dev,
                                   size_t t_kps_pair_stride, int nt_max, const int32_t* nt_dev, const uco_match_params* prm,
                                   uco_match* out_dev, int32_t* n_out_dev) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    int rc = check_params(ctx, prm);
    if (rc) return rc;
    if (n_pairs <= 0 || nq_max <= 0 || nt_max <= 0) return uco_fail(ctx, UCO_E_INVALID, "frame_match: bad sizes");
    if (!q_desc_dev || !q_kps_dev || !t_desc_dev || !t_kps_dev || !out_dev || !n_out_dev) return uco_fail(ctx, UCO_E_INVALID, "frame_match: null pointer");
    const size_t knn_elems = (size_t)n_pairs * nq_max * NN;
    int32_t* knn = (int32_t*)uco_ws(ctx, WS_MATCH_KNN, knn_elems * 8);
    const size_t used_bytes = (size_t)n_pairs * nt_max * 8, cand_bytes = (size_t)n_pairs * nq_max * 8;
    uint8_t* scr = (uint8_t*)uco_ws(ctx, WS_MATCH_SCRATCH, used_bytes + cand_bytes);
    if (!knn || !scr) return UCO_E_NOMEM;
    rc = uco_knn_launch_internal(ctx, q_desc_dev, nq_max, t_desc_dev, nt_max, NN, UCO_KNN_HEAP, knn, knn + knn_elems, n_pairs, nq_dev,
                                 nt_dev, q_pair_stride, t_pair_stride);
    if (rc) return rc;
    MatchArgs A;
    A.knn_idx = knn; A.knn_dist = knn + knn_elems;
    A.q_kps = q_kps_dev; A.q_kps_stride = q_kps_pair_stride; A.t_kps = t_kps_dev; A.t_kps_stride = t_kps_pair_stride;
    A.q_map = nullptr; A.t_map = nullptr; A.q_map_stride = 0; A.f12_pair = nullptr; A.q_sel = A.t_sel = nullptr;
    A.nq_max = nq_max; A.nt_max = nt_max; A.n_t_kps = nt_max; A.nq_dev = nq_dev; A.nt_dev = nt_dev;
    A.used = (unsigned long long*)scr; A.cand = (int2*)(scr + used_bytes);
    A.out = out_dev; A.n_out = n_out_dev; A.prm = *prm;
    A.bow_entry_node = nullptr; A.bow_t_node = nullptr; A.bow_t_ptr = nullptr; A.bow_t_kp = nullptr; A.q_usable = A.t_usable = nullptr; A.q_desc = A.t_desc = nullptr;
    match_filter_kernel<<<n_pairs, MT, 0, ctx->stream>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

int uco_b200_frame_match(uco_b200_ctx* ctx, const uint8_t* q_desc, int nq, size_t q_stride, const uco_keypoint* q_kps,
                         int n_q_kps, const int32_t* q_map, const uint8_t* t_desc, int nt, size_t t_stride,
                         const uco_keypoint* t_kps, int n_t_kps, const int32_t* t_map, const uco_match_params* prm,
                         uco_match* out, int capacity, int* n_out) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    int rc = check_params(ctx, prm);
    if (rc) return rc;
    if (!n_out) return uco_fail(ctx, UCO_E_INVALID, "frame_match: null pointer");
    *n_out = 0;
    if (nq < 0 || nt < 0 || n_q_kps < 0 || n_t_kps < 0) return uco_fail(ctx, UCO_E_INVALID, "frame_match: bad sizes");
    if (nq == 0 || nt == 0) return UCO_OK;  // trainIndex.search fails on an empty index -> {} (framematcher.cpp:239)
    if (!q_desc || !q_kps || !t_desc || !t_kps || !out) return uco_fail(ctx, UCO_E_INVALID, "frame_match: null pointer");
    if (q_stride < 32 || t_stride < 32) return uco_fail(ctx, UCO_E_INVALID, "frame_match: row stride below 32 bytes");
    if (capacity < nq) return uco_fail(ctx, UCO_E_CAPACITY, "frame_match: output capacity %d below the %d query rows", capacity, nq);
    if ((!q_map && n_q_kps < nq) || (!t_map && n_t_kps < nt)) return uco_fail(ctx, UCO_E_INVALID, "frame_match: fewer keypoints than descriptor rows");
    for (int i = 0; q_map && i < nq; i++)
        if (q_map[i] < 0 || q_map[i] >= n_q_kps) return uco_fail(ctx, UCO_E_INVALID, "frame_match: q_map[%d] out of range", i);
    for (int i = 0; t_map && i < nt; i++)
        if (t_map[i] < 0 || t_map[i] >= n_t_kps) return uco_fail(ctx, UCO_E_INVALID, "frame_match: t_map[%d] out of range", i);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_qd = 0, o_td = al(o_qd + (size_t)nq * 32), o_qk = al(o_td + (size_t)nt * 32), o_tk = al(o_qk + sizeof(uco_keypoint) * n_q_kps),
                 o_qm = al(o_tk + sizeof(uco_keypoint) * n_t_kps), o_tm = al(o_qm + 4 * (size_t)nq), in_bytes = al(o_tm + 4 * (size_t)nt);
    uint8_t* din = (uint8_t*)uco_ws(ctx, WS_MATCH_IN, in_bytes);
    const size_t knn_elems = (size_t)nq * NN;
    int32_t* knn = (int32_t*)uco_ws(ctx, WS_MATCH_KNN, knn_elems * 8);
    const size_t used_bytes = (size_t)n_t_kps * 8, cand_bytes = (size_t)nq * 8;
    uint8_t* scr = (uint8_t*)uco_ws(ctx, WS_MATCH_SCRATCH, used_bytes + cand_bytes);
    uint8_t* dout = (uint8_t*)uco_ws(ctx, WS_MATCH_OUT, sizeof(uco_match) * (size_t)nq + 16);
    if (!din || !knn || !scr || !dout) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemcpy2DAsync(din + o_qd, 32, q_desc, q_stride, 32, nq, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpy2DAsync(din + o_td, 32, t_desc, t_stride, 32, nt, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_qk, q_kps, sizeof(uco_keypoint) * n_q_kps, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_tk, t_kps, sizeof(uco_keypoint) * n_t_kps, cudaMemcpyHostToDevice, ctx->stream));
    if (q_map) UCO_CUDA(ctx, cudaMemcpyAsync(din + o_qm, q_map, 4 * (size_t)nq, cudaMemcpyHostToDevice, ctx->stream));
    if (t_map) UCO_CUDA(ctx, cudaMemcpyAsync(din + o_tm, t_map, 4 * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
    rc = uco_knn_launch_internal(ctx, din + o_qd, nq, din + o_td, nt, NN, UCO_KNN_HEAP, knn, knn + knn_elems, 1, nullptr, nullptr, 0, 0);
    if (rc) return rc;
    MatchArgs A;
    A.knn_idx = knn; A.knn_dist = knn + knn_elems;
    A.q_kps = (const uco_keypoint*)(din + o_qk); A.q_kps_stride = 0; A.t_kps = (const uco_keypoint*)(din + o_tk); A.t_kps_stride = 0;
    A.q_map = q_map ? (const int32_t*)(din + o_qm) : nullptr; A.t_map = t_map ? (const int32_t*)(din + o_tm) : nullptr; A.q_map_stride = 0; A.f12_pair = nullptr; A.q_sel = A.t_sel = nullptr;
    A.nq_max = nq; A.nt_max = nt; A.n_t_kps = n_t_kps; A.nq_dev = nullptr; A.nt_dev = nullptr;
    A.used = (unsigned long long*)scr; A.cand = (int2*)(scr + used_bytes);
    A.out = (uco_match*)(dout + 16); A.n_out = (int32_t*)dout; A.prm = *prm;
    A.bow_entry_node = nullptr; A.bow_t_node = nullptr; A.bow_t_ptr = nullptr; A.bow_t_kp = nullptr; A.q_usable = A.t_usable = nullptr; A.q_desc = A.t_desc = nullptr;
    match_filter_kernel<<<1, MT, 0, ctx->stream>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    int n = 0;
    UCO_CUDA(ctx, cudaMemcpyAsync(&n, dout, 4, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n > 0) UCO_CUDA(ctx, cudaMemcpy(out, dout + 16, sizeof(uco_match) * (size_t)n, cudaMemcpyDeviceToHost));
    *n_out = n;
    return UCO_OK;
}

int uco_b200_frame_match_bow(uco_b200_ctx* ctx, const uint8_t* q_desc, size_t q_stride, const uco_keypoint* q_kps, int n_q_kps,
                             const uint8_t* q_usable, const uco_bow_index* q_bow, const uint8_t* t_desc, size_t t_stride,
                             const uco_keypoint* t_kps, int n_t_kps, const uint8_t* t_usable, const uco_bow_index* t_bow,
                             const uco_match_params* prm, uco_match* out, int capacity, int* n_out) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    int rc = check_params(ctx, prm);
    if (rc) return rc;
    if (!n_out || !q_bow || !t_bow) return uco_fail(ctx, UCO_E_INVALID, "frame_match_bow: null pointer");
    *n_out = 0;
    if (n_q_kps < 0 || n_t_kps < 0 || q_bow->n_nodes < 0 || t_bow->n_nodes < 0) return uco_fail(ctx, UCO_E_INVALID, "frame_match_bow: bad sizes");
    if (n_q_kps == 0 || n_t_kps == 0 || q_bow->n_nodes == 0 || t_bow->n_nodes == 0) return UCO_OK;
    if (!q_desc || !q_kps || !t_desc || !t_kps || !out || !q_bow->node_id || !q_bow->ptr || !q_bow->kp || !t_bow->node_id || !t_bow->ptr || !t_bow->kp)
        return uco_fail(ctx, UCO_E_INVALID, "frame_match_bow: null pointer");
    if (q_stride < 32 || t_stride < 32) return uco_fail(ctx, UCO_E_INVALID, "frame_match_bow: row stride below 32 bytes");
    const int ne = q_bow->ptr[q_bow->n_nodes], nte = t_bow->ptr[t_bow->n_nodes];
    if (capacity < ne) return uco_fail(ctx, UCO_E_CAPACITY, "frame_match_bow: output capacity %d below the %d query entries", capacity, ne);
    for (int i = 0; i < ne; i++)
        if ((unsigned)q_bow->kp[i] >= (unsigned)n_q_kps) return uco_fail(ctx, UCO_E_INVALID, "frame_match_bow: query keypoint %d out of range", q_bow->kp[i]);
    for (int i = 0; i < nte; i++)
        if ((unsigned)t_bow->kp[i] >= (unsigned)n_t_kps) return uco_fail(ctx, UCO_E_INVALID, "frame_match_bow: train keypoint %d out of range", t_bow->kp[i]);
    if (ne == 0 || nte == 0) return UCO_OK;
    // the in-step walk of the two std::maps (framematcher.cpp:426-431, 482-491): which train node carries the id of a query node
    std::vector<int32_t> t_of_q(q_bow->n_nodes, -1), entry_node(ne);
    for (int qi = 0, ti = 0; qi < q_bow->n_nodes && ti < t_bow->n_nodes;) {
        if (q_bow->node_id[qi] == t_bow->node_id[ti]) t_of_q[qi++] = ti++;
        else if (q_bow->node_id[qi] < t_bow->node_id[ti]) qi++;
        else ti++;
    }
    for (int qn = 0; qn < q_bow->n_nodes; qn++)
        for (int e = q_bow->ptr[qn]; e < q_bow->ptr[qn + 1]; e++) entry_node[e] = qn;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = al(off + b); return o; };
    const size_t o_qd = take((size_t)n_q_kps * 32), o_td = take((size_t)n_t_kps * 32), o_qk = take(sizeof(uco_keypoint) * (size_t)n_q_kps),
                 o_tk = take(sizeof(uco_keypoint) * (size_t)n_t_kps), o_qe = take(4 * (size_t)ne), o_en = take(4 * (size_t)ne),
                 o_tq = take(4 * (size_t)q_bow->n_nodes), o_tp = take(4 * (size_t)(t_bow->n_nodes + 1)), o_te = take(4 * (size_t)nte),
                 o_qu = take((size_t)n_q_kps), o_tu = take((size_t)n_t_kps);
    uint8_t* din = (uint8_t*)uco_ws(ctx, WS_MATCH_IN, off);
    const size_t used_bytes = (size_t)n_t_kps * 8, cand_bytes = (size_t)ne * 8;
    uint8_t* scr = (uint8_t*)uco_ws(ctx, WS_MATCH_SCRATCH, used_bytes + cand_bytes);
    uint8_t* dout = (uint8_t*)uco_ws(ctx, WS_MATCH_OUT, sizeof(uco_match) * (size_t)ne + 16);
    if (!din || !scr || !dout) return UCO_E_NOMEM;
    cudaStream_t st = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpy2DAsync(din + o_qd, 32, q_desc, q_stride, 32, n_q_kps, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaMemcpy2DAsync(din + o_td, 32, t_desc, t_stride, 32, n_t_kps, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_qk, q_kps, sizeof(uco_keypoint) * (size_t)n_q_kps, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_tk, t_kps, sizeof(uco_keypoint) * (size_t)n_t_kps, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_qe, q_bow->kp, 4 * (size_t)ne, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_en, entry_node.data(), 4 * (size_t)ne, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_tq, t_of_q.data(), 4 * (size_t)q_bow->n_nodes, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_tp, t_bow->ptr, 4 * (size_t)(t_bow->n_nodes + 1), cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaMemcpyAsync(din + o_te, t_bow->kp, 4 * (size_t)nte, cudaMemcpyHostToDevice, st));
    if (q_usable) UCO_CUDA(ctx, cudaMemcpyAsync(din + o_qu, q_usable, (size_t)n_q_kps, cudaMemcpyHostToDevice, st));
    if (t_usable) UCO_CUDA(ctx, cudaMemcpyAsync(din + o_tu, t_usable, (size_t)n_t_kps, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaStreamSynchronize(st));   // entry_node / t_of_q are stack-lifetime host vectors
    MatchArgs A;
    A.knn_idx = nullptr; A.knn_dist = nullptr;
    A.q_kps = (const uco_keypoint*)(din + o_qk); A.q_kps_stride = 0; A.t_kps = (const uco_keypoint*)(din + o_tk); A.t_kps_stride = 0;
    A.q_map = (const int32_t*)(din + o_qe); A.t_map = nullptr; A.q_map_stride = 0; A.f12_pair = nullptr; A.q_sel = A.t_sel = nullptr;
    A.nq_max = ne; A.nt_max = n_t_kps; A.n_t_kps = n_t_kps; A.nq_dev = nullptr; A.nt_dev = nullptr;
    A.used = (unsigned long long*)scr; A.cand = (int2*)(scr + used_bytes);
    A.out = (uco_match*)(dout + 16); A.n_out = (int32_t*)dout; A.prm = *prm;
    A.bow_entry_node = (const int32_t*)(din + o_en); A.bow_t_node = (const int32_t*)(din + o_tq); A.bow_t_ptr = (const int32_t*)(din + o_tp);
    A.bow_t_kp = (const int32_t*)(din + o_te);
    A.q_usable = q_usable ? din + o_qu : nullptr; A.t_usable = t_usable ? din + o_tu : nullptr;
    A.q_desc = (const uint32_t*)(din + o_qd); A.t_desc = (const uint32_t*)(din + o_td);
    match_filter_kernel<<<1, MT, 0, st>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    int n = 0;
    UCO_CUDA(ctx, cudaMemcpyAsync(&n, dout, 4, cudaMemcpyDeviceToHost, st));
    UCO_CUDA(ctx, cudaStreamSynchronize(st));
    if (n > 0) UCO_CUDA(ctx, cudaMemcpy(out, dout + 16, sizeof(uco_match) * (size_t)n, cudaMemcpyDeviceToHost));
    *n_out = n;
    return UCO_OK;
}

// The mapper's pattern (new-map-point creation, src/utils/mapmanager.cpp:9972-10065): FrameMatcher::setParams(train = the keyframe)
// once, then matchEpipolar(query = each neighbour keyframe, F12_i).  One call = one upload, two launches (k-NN of all the
// neighbours' rows against the keyframe's rows; the filters, one block per neighbour), one download.
}  // extern "C" (internal helpers follow)

// optional second stage of the multi-frame matcher: triangulation of every neighbour's matches on the device (triangulate.cu)
int uco_tri_pairs_launch(uco_b200_ctx* ctx, const uco_keypoint* kp1, int n1, const uco_keypoint* kp2, int kp2_stride, const int32_t* n2_dev, int n_frames,
                         const uco_match* matches, int match_stride, const int32_t* n_matches_dev, const float* K1, const float* cam2_dev,
                         const float* sf_dev, int nl1, int nl2, float max_chi2, float ratio_factor, const float* g2f_train, float* xyz_dev,
                         int32_t* counters_dev);
namespace {
struct TriStaged { const uco_match* matches; const float* xyz; int stride; };   // pinned staging of the last call (valid until the next one)
struct TriStage {
    const float* rt; const float* K_nb; const float* K_kf; const float* g2f;
    const float* sf1; int nl1; const float* sf2; int nl2;
    float max_chi2, ratio_factor;
    TriStaged* res;
};
}  // namespace

static int match_multi_impl(uco_b200_ctx* ctx, const uint8_t* t_desc, int nt, size_t t_stride, const uco_keypoint* t_kps, int n_t_kps,
                            const int32_t* t_map, int n_frames, const uint8_t* const* q_desc, const int32_t* nq, size_t q_stride,
                            const uco_keypoint* const* q_kps, const int32_t* n_q_kps, const int32_t* const* q_map, const float* f12,
                            const uco_match_params* prm, uco_match* const* out, int capacity, int32_t* n_out, const TriStage* tri) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    int rc = check_params(ctx, prm);
    if (rc) return rc;
    if (n_frames < 0 || !n_out) return uco_fail(ctx, UCO_E_INVALID, "frame_match_multi: bad arguments");
    for (int f = 0; f < n_frames; f++) n_out[f] = 0;
    if (n_frames == 0 || nt == 0) return UCO_OK;
    if (nt < 0 || n_t_kps < 0 || !t_desc || !t_kps || !q_desc || !nq || !q_kps || !n_q_kps || (!out && !tri) || q_stride < 32 || t_stride < 32)
        return uco_fail(ctx, UCO_E_INVALID, "frame_match_multi: bad arguments");
    if (!t_map && n_t_kps < nt) return uco_fail(ctx, UCO_E_INVALID, "frame_match_multi: fewer train keypoints than descriptor rows");
    if (prm->use_f12 && !f12) return uco_fail(ctx, UCO_E_INVALID, "frame_match_multi: use_f12 without the per-frame matrices");
    int nq_max = 0, nk_max = 0;
    for (int f = 0; f < n_frames; f++) {
        if (nq[f] < 0 || n_q_kps[f] < 0 || (nq[f] && (!q_desc[f] || !q_kps[f] || (!tri && !out[f])))) return uco_fail(ctx, UCO_E_INVALID, "frame_match_multi: frame %d malformed", f);
        if (!(q_map && q_map[f]) && n_q_kps[f] < nq[f]) return uco_fail(ctx, UCO_E_INVALID, "frame_match_multi: frame %d has fewer keypoints than rows", f);
        if (capacity < nq[f] && out && out[f]) return uco_fail(ctx, UCO_E_CAPACITY, "frame_match_multi: output capacity %d below the %d rows of frame %d", capacity, nq[f], f);
        for (int i = 0; q_map && q_map[f] && i < nq[f]; i++)
            if ((unsigned)q_map[f][i] >= (unsigned)n_q_kps[f]) return uco_fail(ctx, UCO_E_INVALID, "frame_match_multi: q_map[%d][%d] out of range", f, i);
        nq_max = nq[f] > nq_max ? nq[f] : nq_max;
        nk_max = n_q_kps[f] > nk_max ? n_q_kps[f] : nk_max;
    }
    for (int i = 0; t_map && i < nt; i++)
        if ((unsigned)t_map[i] >= (unsigned)n_t_kps) return uco_fail(ctx, UCO_E_INVALID, "frame_match_multi: t_map[%d] out of range", i);
    if (nq_max == 0) return UCO_OK;
    const bool any_qmap = q_map != nullptr;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = al(off + b); return o; };
    const size_t F = n_frames;
    const size_t o_td = take((size_t)nt * 32), o_tk = take(sizeof(uco_keypoint) * (size_t)n_t_kps), o_tm = take(t_map ? 4 * (size_t)nt : 0),
                 o_qd = take(F * nq_max * 32), o_qk = take(sizeof(uco_keypoint) * F * nk_max), o_qm = take(any_qmap ? 4 * F * nq_max : 0),
                 o_nq = take(4 * F), o_f = take(36 * F), o_nk = take(tri ? 4 * F : 0), o_cam = take(tri ? 64 * F : 0),
                 o_sf = take(tri ? 16 * UCO_MATCH_MAX_SCALES : 0);
    const size_t in_bytes = off;
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_MATCH_IN, in_bytes);
    uint8_t* din = (uint8_t*)uco_ws(ctx, WS_MATCH_IN, in_bytes);
    const size_t knn_elems = F * nq_max * NN;
    int32_t* knn = (int32_t*)uco_ws(ctx, WS_MATCH_KNN, knn_elems * 8);
    const size_t used_bytes = F * n_t_kps * 8, cand_bytes = F * nq_max * 8;
    uint8_t* scr = (uint8_t*)uco_ws(ctx, WS_MATCH_SCRATCH, used_bytes + cand_bytes);
    const size_t o_xyz = al(al(4 * F) + sizeof(uco_match) * F * nq_max), o_cnt = o_xyz + al(tri ? 12 * F * nq_max : 0);
    const size_t out_bytes = tri ? o_cnt + al(4 * (F + 1)) : al(4 * F) + sizeof(uco_match) * F * nq_max;
    uint8_t* dout = (uint8_t*)uco_ws(ctx, WS_MATCH_OUT, out_bytes);
    uint8_t* hout = (uint8_t*)uco_pinned(ctx, WS_MATCH_OUT, out_bytes);
    if (!h || !din || !knn || !scr || !dout || !hout) return UCO_E_NOMEM;
    for (int i = 0; i < nt; i++) memcpy(h + o_td + 32 * (size_t)i, t_desc + t_stride * (size_t)i, 32);
    memcpy(h + o_tk, t_kps, sizeof(uco_keypoint) * (size_t)n_t_kps);
    if (t_map) memcpy(h + o_tm, t_map, 4 * (size_t)nt);
    for (size_t f = 0; f < F; f++) {
        uint8_t* qd = h + o_qd + f * nq_max * 32;
        if (q_stride == 32) memcpy(qd, q_desc[f], 32 * (size_t)nq[f]);
        else for (int i = 0; i < nq[f]; i++) memcpy(qd + 32 * (size_t)i, q_desc[f] + q_stride * (size_t)i, 32);
        memcpy(h + o_qk + sizeof(uco_keypoint) * f * nk_max, q_kps[f], sizeof(uco_keypoint) * (size_t)n_q_kps[f]);
        if (any_qmap) {
            int32_t* m = (int32_t*)(h + o_qm) + f * nq_max;
            for (int i = 0; i < nq[f]; i++) m[i] = q_map[f] ? q_map[f][i] : i;
        }
        ((int32_t*)(h + o_nq))[f] = nq[f];
        if (f12) memcpy(h + o_f + 36 * f, f12 + 9 * f, 36);
        else memset(h + o_f + 36 * f, 0, 36);
        if (tri) {   // K2 | R | t of neighbour f (camera 1 = the keyframe)
            ((int32_t*)(h + o_nk))[f] = n_q_kps[f];
            float* c2 = (float*)(h + o_cam) + 16 * f;
            memcpy(c2, tri->K_nb + 4 * f, 16);
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++) c2[4 + 3 * r + c] = tri->rt[16 * f + 4 * r + c];
                c2[13 + r] = tri->rt[16 * f + 4 * r + 3];
            }
        }
    }
    if (tri) {   // invScaleFactors: 1.f/(f*f) per octave (misc.cpp:943-945), then the factors themselves
        float* sf = (float*)(h + o_sf);
        memset(sf, 0, 16 * UCO_MATCH_MAX_SCALES);
        for (int l = 0; l < tri->nl1; l++) { sf[l] = 1.f / (tri->sf1[l] * tri->sf1[l]); sf[2 * UCO_MATCH_MAX_SCALES + l] = tri->sf1[l]; }
        for (int l = 0; l < tri->nl2; l++) { sf[UCO_MATCH_MAX_SCALES + l] = 1.f / (tri->sf2[l] * tri->sf2[l]); sf[3 * UCO_MATCH_MAX_SCALES + l] = tri->sf2[l]; }
    }
    cudaStream_t st = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(din, h, in_bytes, cudaMemcpyHostToDevice, st));
    rc = uco_knn_launch_internal(ctx, din + o_qd, nq_max, din + o_td, nt, NN, UCO_KNN_HEAP, knn, knn + knn_elems, n_frames, (const int*)(din + o_nq),
                                 nullptr, (size_t)nq_max * 32, 0);
    if (rc) return rc;
    MatchArgs A;
    A.knn_idx = knn; A.knn_dist = knn + knn_elems;
    A.q_kps = (const uco_keypoint*)(din + o_qk); A.q_kps_stride = nk_max; A.t_kps = (const uco_keypoint*)(din + o_tk); A.t_kps_stride = 0;
    A.q_map = any_qmap ? (const int32_t*)(din + o_qm) : nullptr; A.q_map_stride = nq_max; A.q_sel = A.t_sel = nullptr; A.t_map = t_map ? (const int32_t*)(din + o_tm) : nullptr;
    A.f12_pair = f12 ? (const float*)(din + o_f) : nullptr;
    A.nq_max = nq_max; A.nt_max = nt; A.n_t_kps = n_t_kps; A.nq_dev = (const int32_t*)(din + o_nq); A.nt_dev = nullptr;
    A.used = (unsigned long long*)scr; A.cand = (int2*)(scr + used_bytes);
    A.out = (uco_match*)(dout + al(4 * F)); A.n_out = (int32_t*)dout; A.prm = *prm;
    A.bow_entry_node = nullptr; A.bow_t_node = nullptr; A.bow_t_ptr = nullptr; A.bow_t_kp = nullptr; A.q_usable = A.t_usable = nullptr; A.q_desc = A.t_desc = nullptr;
    match_filter_kernel<<<n_frames, MT, 0, st>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    if (tri) {   // Triangulate + the mapper's gates on the matcher's device output, all neighbours in one launch
        rc = uco_tri_pairs_launch(ctx, A.t_kps, n_t_kps, A.q_kps, nk_max, (const int32_t*)(din + o_nk), n_frames, A.out, nq_max, A.n_out, tri->K_kf,
                                  (const float*)(din + o_cam), (const float*)(din + o_sf), tri->nl1, tri->nl2, tri->max_chi2, tri->ratio_factor, tri->g2f,
                                  (float*)(dout + o_xyz), (int32_t*)(dout + o_cnt));
        if (rc) return rc;
    }
    UCO_CUDA(ctx, cudaMemcpyAsync(hout, dout, out_bytes, cudaMemcpyDeviceToHost, st));
    UCO_CUDA(ctx, cudaStreamSynchronize(st));
    if (tri && ((const int32_t*)(hout + o_cnt))[0]) return uco_fail(ctx, UCO_E_INVALID, "new_points: a match refers to a keypoint or an octave out of range");
    for (size_t f = 0; f < F; f++) {
        const int n = ((const int32_t*)hout)[f];
        n_out[f] = n;
        if (n > 0 && out && out[f]) memcpy(out[f], hout + al(4 * F) + sizeof(uco_match) * f * nq_max, sizeof(uco_match) * (size_t)n);
    }
    if (tri) {   // hand the staged results to the caller's merge
        tri->res->matches = (const uco_match*)(hout + al(4 * F));
        tri->res->xyz = (const float*)(hout + o_xyz);
        tri->res->stride = nq_max;
    }
    return UCO_OK;
}


extern "C" {

int uco_b200_frame_match_multi(uco_b200_ctx* ctx, const uint8_t* t_desc, int nt, size_t t_stride, const uco_keypoint* t_kps, int n_t_kps,
                               const int32_t* t_map, int n_frames, const uint8_t* const* q_desc, const int32_t* nq, size_t q_stride,
                               const uco_keypoint* const* q_kps, const int32_t* n_q_kps, const int32_t* const* q_map, const float* f12,
                               const uco_match_params* prm, uco_match* const* out, int capacity, int32_t* n_out) {
    return match_multi_impl(ctx, t_desc, nt, t_stride, t_kps, n_t_kps, t_map, n_frames, q_desc, nq, q_stride, q_kps, n_q_kps, q_map, f12, prm, out, capacity,
                            n_out, nullptr);
}

// New-map-point creation of one keyframe as a unit (MapManager's createNewPoints, src/utils/mapmanager.cpp:9772-10788, de-obfuscated):
//   FrameMatcher::setParams(frame, MODE_UNASSIGNED, ...) ; per neighbour f (the reference's OpenMP loop :9992): matchEpipolar(neighbour,
//   MODE_UNASSIGNED, T_f) -> Triangulate(frame, neighbour, T_f, matches) -> global coordinates -> scale-consistency gate -> (trainIdx,
//   neighbour, queryIdx, p, distance); then the merge: one point per keyframe keypoint (std::map order), position / distance of the LAST
//   neighbour that saw it (the reference's minimum-octave loop never updates its minimum, so it always ends on the last element),
//   observations in neighbour order; above max_points the points with the smallest distance survive (std::sort on dist + resize).
// One upload, three launches (k-NN, filters, triangulation), one download; the merge runs on the host over the downloaded lists.
int uco_b200_new_points(uco_b200_ctx* ctx, const uint8_t* t_desc, int nt, size_t t_stride, const uco_keypoint* t_kps, int n_t_kps, const int32_t* t_map,
                        int n_frames, const uint8_t* const* q_desc, const int32_t* nq, size_t q_stride, const uco_keypoint* const* q_kps,
                        const int32_t* n_q_kps, const int32_t* const* q_map, const float* f12, const float* rt, const float* K_nb,
                        const uco_new_points_params* prm, int32_t* n_points, int32_t* pt_kpt, float* pt_xyz, float* pt_dist, int32_t* obs_ptr,
                        int32_t* obs_frame, int32_t* obs_kpt, int capacity_points, int capacity_obs, uco_match* const* matches, int32_t* n_matches,
                        float* const* xyz) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    if (!prm || !n_points || n_frames < 0) return uco_fail(ctx, UCO_E_INVALID, "new_points: bad arguments");
    *n_points = 0;
    if (obs_ptr && capacity_points >= 0) obs_ptr[0] = 0;
    if (n_frames == 0 || nt == 0) return UCO_OK;
    if (!rt || !K_nb || !f12 || !prm->match.use_f12) return uco_fail(ctx, UCO_E_INVALID, "new_points: the epipolar matcher needs f12, rt and K_nb per neighbour (match.use_f12 = 1)");
    if (prm->n_levels_kf <= 0 || prm->n_levels_nb <= 0 || prm->n_levels_kf > UCO_MATCH_MAX_SCALES || prm->n_levels_nb > UCO_MATCH_MAX_SCALES ||
        !prm->scale_factors_kf || !prm->scale_factors_nb)
        return uco_fail(ctx, UCO_E_INVALID, "new_points: scale factor tables of 1..%d levels expected", UCO_MATCH_MAX_SCALES);
    if (!pt_kpt || !pt_xyz || !pt_dist || !obs_ptr || !obs_frame || !obs_kpt) return uco_fail(ctx, UCO_E_INVALID, "new_points: null output");
    TriStaged staged{nullptr, nullptr, 0};
    TriStage tri{rt, K_nb, prm->K_kf, prm->g2f_kf, prm->scale_factors_kf, prm->n_levels_kf, prm->scale_factors_nb, prm->n_levels_nb,
                 prm->max_chi2, prm->scale_ratio_factor, &staged};
    std::vector<int32_t> nm(n_frames, 0);
    int cap = 0;
    for (int f = 0; f < n_frames; f++) cap = nq && nq[f] > cap ? nq[f] : cap;
    int rc = match_multi_impl(ctx, t_desc, nt, t_stride, t_kps, n_t_kps, t_map, n_frames, q_desc, nq, q_stride, q_kps, n_q_kps, q_map, f12, &prm->match,
                              matches, cap, nm.data(), &tri);
    if (rc) return rc;
    if (n_matches) memcpy(n_matches, nm.data(), 4 * (size_t)n_frames);
    if (!staged.matches) return UCO_OK;   // nothing to match
    // accepted records per keyframe keypoint, in (neighbour, match) order
    struct Rec { int32_t f, q; float d; const float* p; };
    std::map<int32_t, std::vector<Rec>> groups;
    for (int f = 0; f < n_frames; f++) {
        const uco_match* m = staged.matches + (size_t)f * staged.stride;
        const float* p = staged.xyz + 3 * (size_t)f * staged.stride;
        if (xyz && xyz[f]) memcpy(xyz[f], p, 12 * (size_t)nm[f]);
        for (int i = 0; i < nm[f]; i++)
            if (!std::isnan(p[3 * i])) groups[m[i].trainIdx].push_back({f, m[i].queryIdx, m[i].distance, p + 3 * i});
    }
    std::vector<std::pair<int32_t, const std::vector<Rec>*>> pts;
    for (const auto& g : groups) pts.push_back({g.first, &g.second});
    std::vector<int> order(pts.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    size_t keep = pts.size();
    if (prm->max_points >= 0 && pts.size() > (size_t)prm->max_points) {   // std::sort on dist + resize (not stable: the same comparisons on the same sequence)
        std::sort(order.begin(), order.end(), [&](int a, int b) { return pts[a].second->back().d < pts[b].second->back().d; });
        keep = (size_t)prm->max_points;
    }
    size_t n_obs = 0;
    for (size_t j = 0; j < keep; j++) n_obs += pts[order[j]].second->size();
    if (keep > (size_t)capacity_points || n_obs > (size_t)capacity_obs)
        return uco_fail(ctx, UCO_E_CAPACITY, "new_points: %zu points / %zu observations exceed the output capacity (%d / %d)", keep, n_obs, capacity_points, capacity_obs);
    int at = 0;
    for (size_t j = 0; j < keep; j++) {
        const auto& g = pts[order[j]];
        const Rec& best = g.second->back();
        pt_kpt[j] = g.first;
        pt_xyz[3 * j] = best.p[0]; pt_xyz[3 * j + 1] = best.p[1]; pt_xyz[3 * j + 2] = best.p[2];
        pt_dist[j] = best.d;
        for (const Rec& r : *g.second) { obs_frame[at] = r.f; obs_kpt[at] = r.q; at++; }
        obs_ptr[j + 1] = at;
    }
    *n_points = (int)keep;
    return UCO_OK;
}

}  // extern "C"

// ---- per-keyframe work of the mapper on DEVICE-RESIDENT frames ------------------------------------------------------------------------
// What UcoSLAM's mapper does with a new keyframe before local BA (src/utils/mapmanager.cpp): its bag of words
// (KPFrameDataBase::computeBow, src/map_types/keyframedatabase.cpp:310-321 -> fbow transform at level 3) and, for new-map-point
// creation (:9972-10065), FrameMatcher::setParams(train = the keyframe) + matchEpipolar(query = each neighbour keyframe, F12).
// The frames are rows of a frame-strided device buffer (the extractor's batch output): n_kf keyframes, keyframe j has the
// neighbours nb_frame[nb_ptr[j] .. nb_ptr[j+1]).  Three launches for the whole batch: one fbow walk over all keyframes'
// descriptors, one k-NN over all (neighbour, keyframe) pairs, one filter pass.  Index arrays and f12 are HOST arrays.
int uco_knn_launch_selected(uco_b200_ctx* ctx, const uint8_t* q_dev, int nq, const uint8_t* t_dev, int nt, int k, int order,
                            int32_t* idx_dev, int32_t* dist_dev, int n_pairs, const int* nq_dev, const int* nt_dev,
                            size_t q_stride, size_t t_stride, const int* q_sel, const int* t_sel);
struct uco_b200_voc;
int uco_bow_transform_segments(uco_b200_ctx* ctx, const uco_b200_voc* voc, const uint8_t* desc_dev, size_t frame_stride_bytes, const int* n_dev,
                               const int* seg_sel_dev, int n_seg, int seg_rows, int level, uint32_t* word_dev, float* weight_dev,
                               uint32_t* node_dev, int* err_dev);
extern "C" int uco_orb_resident(uco_b200_ctx* ctx, const uco_keypoint** d_kps, const uint8_t** d_desc, const int** d_nout, int* max_features, int* n_frames);

extern "C" {

int uco_b200_keyframes_batch_dev(uco_b200_ctx* ctx, const uco_b200_voc* voc, int bow_level, const uco_keypoint* kps_dev, size_t kps_frame_stride,
                                 const uint8_t* desc_dev, size_t desc_frame_stride, const int32_t* n_kp_dev, int kp_cap, int n_frames,
                                 int n_kf, const int32_t* kf_frame, const int32_t* nb_ptr, const int32_t* nb_frame, const float* f12,
                                 const uco_match_params* prm, uint32_t* word_dev, float* weight_dev, uint32_t* node_dev,
                                 uco_match* match_dev, int32_t* n_match_dev) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    int rc = check_params(ctx, prm);
    if (rc) return rc;
    if (n_kf <= 0 || kp_cap <= 0 || n_frames <= 0 || !kps_dev || !desc_dev || !n_kp_dev || !kf_frame || !nb_ptr || (nb_ptr[n_kf] > 0 && !nb_frame))
        return uco_fail(ctx, UCO_E_INVALID, "keyframes_batch: bad arguments");
    const int n_pairs = nb_ptr[n_kf];
    if (n_pairs < 0 || (n_pairs > 0 && (!match_dev || !n_match_dev))) return uco_fail(ctx, UCO_E_INVALID, "keyframes_batch: bad arguments");
    if (voc && (!word_dev || !weight_dev || !node_dev)) return uco_fail(ctx, UCO_E_INVALID, "keyframes_batch: null bag-of-words output");
    if (prm->use_f12 && n_pairs > 0 && !f12) return uco_fail(ctx, UCO_E_INVALID, "keyframes_batch: use_f12 without the per-pair matrices");
    for (int j = 0; j < n_kf; j++) {
        if ((unsigned)kf_frame[j] >= (unsigned)n_frames || nb_ptr[j + 1] < nb_ptr[j]) return uco_fail(ctx, UCO_E_INVALID, "keyframes_batch: keyframe %d malformed", j);
        for (int e = nb_ptr[j]; e < nb_ptr[j + 1]; e++)
            if ((unsigned)nb_frame[e] >= (unsigned)n_frames) return uco_fail(ctx, UCO_E_INVALID, "keyframes_batch: neighbour frame %d out of range", nb_frame[e]);
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = al(off + b); return o; };
    const size_t o_kf = take(4 * (size_t)n_kf), o_q = take(4 * (size_t)n_pairs), o_t = take(4 * (size_t)n_pairs), o_f = take(36 * (size_t)n_pairs), o_err = take(16);
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_MATCH_IN, off);
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_MATCH_IN, off);
    if (!h || !d) return UCO_E_NOMEM;
    if (!ctx->stage_event) UCO_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_event, cudaEventDisableTiming));
    else UCO_CUDA(ctx, cudaEventSynchronize(ctx->stage_event));   // the previous call's upload has left the staging buffer
    memcpy(h + o_kf, kf_frame, 4 * (size_t)n_kf);
    for (int j = 0; j < n_kf; j++)
        for (int e = nb_ptr[j]; e < nb_ptr[j + 1]; e++) {
            ((int32_t*)(h + o_q))[e] = nb_frame[e];      // query = the neighbour
            ((int32_t*)(h + o_t))[e] = kf_frame[j];      // train = the keyframe
        }
    if (f12 && n_pairs) memcpy(h + o_f, f12, 36 * (size_t)n_pairs);
    memset(h + o_err, 0, 16);
    cudaStream_t st = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d, h, off, cudaMemcpyHostToDevice, st));
    UCO_CUDA(ctx, cudaEventRecord(ctx->stage_event, st));
    ctx->kf_err_dev = (int*)(d + o_err);
    if (voc) {
        rc = uco_bow_transform_segments(ctx, voc, desc_dev, desc_frame_stride, n_kp_dev, (const int*)(d + o_kf), n_kf, kp_cap, bow_level, word_dev,
                                        weight_dev, node_dev, (int*)(d + o_err));
        if (rc) return rc;
    }
    if (n_pairs == 0) return UCO_OK;
    const size_t knn_elems = (size_t)n_pairs * kp_cap * NN;
    int32_t* knn = (int32_t*)uco_ws(ctx, WS_MATCH_KNN, knn_elems * 8);
    const size_t used_bytes = (size_t)n_pairs * kp_cap * 8, cand_bytes = (size_t)n_pairs * kp_cap * 8;
    uint8_t* scr = (uint8_t*)uco_ws(ctx, WS_MATCH_SCRATCH, used_bytes + cand_bytes);
    if (!knn || !scr) return UCO_E_NOMEM;
    rc = uco_knn_launch_selected(ctx, desc_dev, kp_cap, desc_dev, kp_cap, NN, UCO_KNN_HEAP, knn, knn + knn_elems, n_pairs, n_kp_dev, n_kp_dev,
                                 desc_frame_stride, desc_frame_stride, (const int*)(d + o_q), (const int*)(d + o_t));
    if (rc) return rc;
    MatchArgs A;
    A.knn_idx = knn; A.knn_dist = knn + knn_elems;
    A.q_kps = kps_dev; A.q_kps_stride = kps_frame_stride; A.t_kps = kps_dev; A.t_kps_stride = kps_frame_stride;
    A.q_map = nullptr; A.t_map = nullptr; A.q_map_stride = 0; A.q_sel = (const int32_t*)(d + o_q); A.t_sel = (const int32_t*)(d + o_t);
    A.f12_pair = (f12 && prm->use_f12) ? (const float*)(d + o_f) : nullptr;
    A.nq_max = kp_cap; A.nt_max = kp_cap; A.n_t_kps = kp_cap; A.nq_dev = n_kp_dev; A.nt_dev = n_kp_dev;
    A.used = (unsigned long long*)scr; A.cand = (int2*)(scr + used_bytes);
    A.out = match_dev; A.n_out = n_match_dev; A.prm = *prm;
    A.bow_entry_node = nullptr; A.bow_t_node = nullptr; A.bow_t_ptr = nullptr; A.bow_t_kp = nullptr; A.q_usable = A.t_usable = nullptr; A.q_desc = A.t_desc = nullptr;
    match_filter_kernel<<<n_pairs, MT, 0, st>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

// the same on the frames of this context's LAST extraction call (uco_b200_track_frames / uco_b200_orb_extract_batch): host index arrays
// in, host results out — words / weights / level nodes of keyframe j at [j * max_features ..), matches of pair e at
// [e * max_features ..) with n_matches[e]; one small upload, three launches, one staged download
int uco_b200_keyframes_batch(uco_b200_ctx* ctx, const uco_b200_voc* voc, int bow_level, int n_kf, const int32_t* kf_frame, const int32_t* nb_ptr,
                             const int32_t* nb_frame, const float* f12, const uco_match_params* prm, uint32_t* word, float* weight, uint32_t* node,
                             uco_match* matches, int32_t* n_matches) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    const uco_keypoint* d_kps; const uint8_t* d_desc; const int* d_nout; int mf, nfr;
    int rc = uco_orb_resident(ctx, &d_kps, &d_desc, &d_nout, &mf, &nfr);
    if (rc) return rc;
    if (n_kf <= 0 || !nb_ptr) return uco_fail(ctx, UCO_E_INVALID, "keyframes_batch: bad arguments");
    const int n_pairs = nb_ptr[n_kf];
    if (n_pairs < 0 || (voc && (!word || !weight || !node)) || (n_pairs > 0 && (!matches || !n_matches))) return uco_fail(ctx, UCO_E_INVALID, "keyframes_batch: null output");
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t K = (size_t)n_kf * mf, P = (size_t)n_pairs * mf;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = al(off + b); return o; };
    const size_t o_w = take(voc ? 4 * K : 0), o_wt = take(voc ? 4 * K : 0), o_n = take(voc ? 4 * K : 0), o_nm = take(4 * (size_t)n_pairs + 16), o_m = take(sizeof(uco_match) * P);
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_MATCH_OUT, off);
    uint8_t* ho = (uint8_t*)uco_pinned(ctx, WS_MATCH_OUT, off);
    if (!d || !ho) return UCO_E_NOMEM;
    rc = uco_b200_keyframes_batch_dev(ctx, voc, bow_level, d_kps, (size_t)mf, d_desc, (size_t)32 * mf, d_nout, mf, nfr, n_kf, kf_frame, nb_ptr, nb_frame, f12,
                                      prm, (uint32_t*)(d + o_w), (float*)(d + o_wt), (uint32_t*)(d + o_n), (uco_match*)(d + o_m), (int32_t*)(d + o_nm));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(ho, d, off, cudaMemcpyDeviceToHost, st));
    int* herr = (int*)uco_pinned(ctx, WS_GENERIC0, 16);
    if (!herr) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemcpyAsync(herr, ctx->kf_err_dev, 4, cudaMemcpyDeviceToHost, st));
    if (n_pairs >= 16) UCO_CUDA(ctx, uco_sleep_sync(ctx));   // milliseconds of work: sleep instead of spinning (see uco_track_check_errors)
    else UCO_CUDA(ctx, cudaStreamSynchronize(st));
    if (*herr) return uco_fail(ctx, UCO_E_FORMAT, "keyframes_batch: malformed vocabulary (cycle or block index out of range)");
    if (voc) { memcpy(word, ho + o_w, 4 * K); memcpy(weight, ho + o_wt, 4 * K); memcpy(node, ho + o_n, 4 * K); }
    for (int e = 0; e < n_pairs; e++) {
        const int n = ((const int32_t*)(ho + o_nm))[e];
        n_matches[e] = n;
        if (n > 0) memcpy(matches + (size_t)e * mf, ho + o_m + sizeof(uco_match) * (size_t)e * mf, sizeof(uco_match) * (size_t)n);
    }
    return UCO_OK;
}

}  // extern "C"
