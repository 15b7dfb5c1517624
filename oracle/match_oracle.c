/*
 * match_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Plain-C restatement of the descriptor matcher's post-filters, i.e. everything FrameMatcher_Flann::matchEpipolar does
 * with the k-NN result:
 *   /root/reference/src/utils/framematcher.cpp:228-322   per query: scan the k=10 candidates in the order the index returns
 *                                                        them (xflann heap order, sorted=false), keep best / second best under
 *                                                        minDescDist, |octave difference| <= maxOctaveDiff and the optional
 *                                                        epipolar gate (chi2 3.84 * scaleFactor^2); second-best ratio test
 *                                                        only when the runner-up lies in the query keypoint's octave
 *   /root/reference/src/basictypes/misc.cpp:153-185      filter_ambiguous_train: per train keypoint keep the match of least
 *                                                        distance, the earlier one on ties; stable removal (:105-107)
 *   /root/reference/src/utils/framematcher.cpp:67-108, 288-316  30-bin rotation histogram (bin = round(rot/30), so only bins
 *                                                        0..12 are ever hit), keep the three fullest bins (10 % rule)
 *   /root/reference/src/basictypes/misc.h:72-81          epipolarLineSqDist in float
 * The k-NN itself is oracle/knn_oracle.c (exact "linear" search pinned to the reference's xflann).  The reference searches a
 * k-means tree with 16 checks (framematcher.cpp:214,239), an APPROXIMATION of these exact neighbours (SURVEY.md 7 hard part 6).
 * PARITY UNPINNED for the post-filters: framematcher.cpp / misc.cpp need OpenCV C++ and the Frame / Map classes and cannot
 * be compiled here, and the reference holds no golden vectors for them; this file follows the cited lines statement by
 * statement.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } kp_t;      /* cv::KeyPoint */
typedef struct { int32_t queryIdx, trainIdx, imgIdx; float distance; } dmatch_t;            /* cv::DMatch */

int oracle_hamming_knn(const uint8_t* q, int nq, size_t q_stride, const uint8_t* t, int nt, size_t t_stride, int k,
                       int order, int32_t* idx, int32_t* dist);

static float epipolarLineSqDist(const float* kp1, const float* kp2, const float* F) { /* misc.h:72-81, F row-major 3x3 */
    const float a = kp1[0] * F[0] + kp1[1] * F[3] + F[6];
    const float b = kp1[0] * F[1] + kp1[1] * F[4] + F[7];
    const float den = a * a + b * b;
    if (den == 0) return FLT_MAX;
    const float c = kp1[0] * F[2] + kp1[1] * F[5] + F[8];
    const float num = a * kp2[0] + b * kp2[1] + c;
    return num * num / den;
}

static int remove_unused(dmatch_t* m, int n) { /* misc.cpp:105-107 */
    int o = 0;
    for (int i = 0; i < n; i++)
        if (!(m[i].trainIdx == -1 || m[i].queryIdx == -1)) m[o++] = m[i];
    return o;
}

/* what both matchers do with their raw matches: filter_ambiguous_train (misc.cpp:153-185) and the rotation-consistency
 * histogram (framematcher.cpp:288-316 and, identically, :497-528) */
static int match_tail(dmatch_t* out, int n, const kp_t* q_kps, const kp_t* t_kps, int checkOrientation) {
    /* filter_ambiguous_train, misc.cpp:153-185 */
    if (n) {
        int maxT = -1;
        for (int i = 0; i < n; i++) if (out[i].trainIdx > maxT) maxT = out[i].trainIdx;
        int* used = malloc(sizeof(int) * (maxT + 1));
        for (int i = 0; i <= maxT; i++) used[i] = -1;
        int need = 0;
        for (int idx = 0; idx < n; idx++) {
            int t = out[idx].trainIdx;
            if (used[t] == -1) used[t] = idx;
            else {
                if (out[used[t]].distance > out[idx].distance) { out[used[t]].trainIdx = -1; used[t] = idx; need = 1; }
                else { out[idx].trainIdx = -1; need = 1; }
            }
        }
        free(used);
        if (need) n = remove_unused(out, n);
    }
    if (checkOrientation) { /* framematcher.cpp:288-316 */
        enum { NB = 30 };
        int cnt[NB];
        memset(cnt, 0, sizeof(cnt));
        int* bin_of = malloc(sizeof(int) * (n + 1));
        const float factor = 1.0f / (float)NB;
        for (int m = 0; m < n; m++) {
            float rot = t_kps[out[m].trainIdx].angle - q_kps[out[m].queryIdx].angle;
            if (rot < 0.0) rot += 360.0f;
            size_t bin = (size_t)roundf(rot * factor);
            if (bin == NB) bin = 0;
            bin_of[m] = (int)bin;
            cnt[bin]++;
        }
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1; /* :67-108 */
        for (int i = 0; i < NB; i++) {
            const int s = cnt[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int m = 0; m < n; m++) {
            int b = bin_of[m];
            if (b == ind1 || b == ind2 || b == ind3) continue;
            out[m].queryIdx = out[m].trainIdx = -1;
        }
        free(bin_of);
        n = remove_unused(out, n);
    }
    return n;
}

int oracle_frame_match_knn(const int32_t* indices, const int32_t* idist, int nq, const kp_t* q_kps, const int32_t* q_map, const kp_t* t_kps,
                           const int32_t* t_map, float minDescDist, float nn_match_ratio, int checkOrientation, int maxOctaveDiff,
                           const float* F12, const float* scaleFactors, int nScale, dmatch_t* out);

/* q_map / t_map: keypoint index of descriptor row i (FrameMatcher::manageMode, framematcher.cpp:160-198); NULL = identity.
 * F12: 9 floats or NULL.  Returns the number of matches written to out (capacity >= nq). */
int oracle_frame_match(const uint8_t* q_desc, int nq, size_t q_stride, const kp_t* q_kps, const int32_t* q_map,
                       const uint8_t* t_desc, int nt, size_t t_stride, const kp_t* t_kps, const int32_t* t_map,
                       float minDescDist, float nn_match_ratio, int checkOrientation, int maxOctaveDiff, const float* F12,
                       const float* scaleFactors, int nScale, dmatch_t* out) {
    const int nn = 10;
    if (nq <= 0 || nt <= 0) return 0;
    int32_t* indices = malloc(sizeof(int32_t) * nq * nn);
    int32_t* idist = malloc(sizeof(int32_t) * nq * nn);
    oracle_hamming_knn(q_desc, nq, q_stride, t_desc, nt, t_stride, nn, 0, indices, idist);
    int n = oracle_frame_match_knn(indices, idist, nq, q_kps, q_map, t_kps, t_map, minDescDist, nn_match_ratio, checkOrientation, maxOctaveDiff,
                                   F12, scaleFactors, nScale, out);
    free(indices); free(idist);
    return n;
}

/* the same filters on a GIVEN 10-NN table (nq x 10 indices / distances as xflann returns them): lets bench.py's CPU arm feed the
 * reference's own approximate index (HKMeans(32,0) + 16 checks, framematcher.cpp:213,239) instead of the exact search */
int oracle_frame_match_knn(const int32_t* indices, const int32_t* idist, int nq, const kp_t* q_kps, const int32_t* q_map, const kp_t* t_kps,
                           const int32_t* t_map, float minDescDist, float nn_match_ratio, int checkOrientation, int maxOctaveDiff,
                           const float* F12, const float* scaleFactors, int nScale, dmatch_t* out) {
    const int nn = 10;
    if (nq <= 0) return 0;
    float* sf2 = malloc(sizeof(float) * (nScale + 1));
    for (int i = 0; i < nScale; i++) sf2[i] = scaleFactors[i] * scaleFactors[i];
    int n = 0;
    for (int i = 0; i < nq; i++) {
        float bestDist = minDescDist, bestDist2 = FLT_MAX;
        int64_t bestQuery = -1, bestTrain = -1;
        int octaveBest2 = -1;
        int queryIndex = q_map ? q_map[i] : i;
        const kp_t* qk = &q_kps[queryIndex];
        for (int j = 0; j < nn; j++) {
            if (indices[i * nn + j] < 0) continue; /* fewer than k train rows (the reference would index out of range) */
            float d = (float)idist[i * nn + j];
            if (d > minDescDist) continue;
            if (d < bestDist2) {
                int trainIndex = t_map ? t_map[indices[i * nn + j]] : indices[i * nn + j];
                const kp_t* tk = &t_kps[trainIndex];
                if (abs(tk->octave - qk->octave) > maxOctaveDiff) continue;
                if (F12)
                    if (epipolarLineSqDist(&tk->x, &qk->x, F12) >= 3.84 * sf2[qk->octave]) continue;
                if (d < bestDist) { bestDist = d; bestQuery = queryIndex; bestTrain = trainIndex; }
                else { bestDist2 = d; octaveBest2 = tk->octave; }
            }
        }
        if (bestQuery != -1) {
            if (!(octaveBest2 == q_kps[bestQuery].octave && bestDist > bestDist2 * nn_match_ratio)) {
                out[n].queryIdx = (int32_t)bestQuery; out[n].trainIdx = (int32_t)bestTrain; out[n].imgIdx = -1; out[n].distance = bestDist;
                n++;
            }
        }
    }
    free(sf2);
    return match_tail(out, n, q_kps, t_kps, checkOrientation);
}


/* FrameMatcher_BoW::matchEpipolar, /root/reference/src/utils/framematcher.cpp:407-541: the two frames' fBow2 (level-3 node id ->
 * keypoint indices, std::map order = ascending node id) are walked in step; inside a common node every usable query keypoint
 * looks for its best train keypoint (octave window, optional epipolar gate, Hamming distance below minDescDist); ANY later
 * candidate that is not a new best overwrites the runner-up (:453-460), and the ratio test applies only when that runner-up lies
 * in the query keypoint's octave.  Then the same tail as the Flann matcher.  *_usable: isUsed(frame, idx, mode) per keypoint, or NULL.
 * PARITY UNPINNED (needs OpenCV C++ / Frame): restated statement by statement. */
int oracle_frame_match_bow(const uint8_t* q_desc, const kp_t* q_kps, const uint8_t* q_usable, int q_nodes, const uint32_t* q_node_id,
                           const int32_t* q_ptr, const int32_t* q_kp, const uint8_t* t_desc, const kp_t* t_kps, const uint8_t* t_usable,
                           int t_nodes, const uint32_t* t_node_id, const int32_t* t_ptr, const int32_t* t_kp, float minDescDist,
                           float nn_match_ratio, int checkOrientation, int maxOctaveDiff, const float* F12, const float* scaleFactors,
                           int nScale, dmatch_t* out) {
    float* sf2 = malloc(sizeof(float) * (nScale + 1));
    for (int i = 0; i < nScale; i++) sf2[i] = scaleFactors[i] * scaleFactors[i];
    int n = 0, qi = 0, ti = 0;
    while (qi < q_nodes && ti < t_nodes) {
        if (q_node_id[qi] == t_node_id[ti]) {
            for (int a = q_ptr[qi]; a < q_ptr[qi + 1]; a++) {
                const int qidx = q_kp[a];
                if (q_usable && !q_usable[qidx]) continue;
                const kp_t* qk = &q_kps[qidx];
                float bestDist = minDescDist, bestDist2 = FLT_MAX;
                int64_t bestQuery = -1, bestTrain = -1;
                int octaveBest2 = -1;
                for (int b = t_ptr[ti]; b < t_ptr[ti + 1]; b++) {
                    const int tidx = t_kp[b];
                    if (t_usable && !t_usable[tidx]) continue;
                    const kp_t* tk = &t_kps[tidx];
                    if (abs(tk->octave - qk->octave) > maxOctaveDiff) continue;
                    if (F12)
                        if (epipolarLineSqDist(&tk->x, &qk->x, F12) >= 3.84 * sf2[qk->octave]) continue;
                    int pc = 0;
                    const uint64_t* x = (const uint64_t*)(t_desc + 32 * (size_t)tidx);
                    const uint64_t* y = (const uint64_t*)(q_desc + 32 * (size_t)qidx);
                    for (int w = 0; w < 4; w++) pc += __builtin_popcountll(x[w] ^ y[w]);
                    const float dist = (float)pc;
                    if (dist < bestDist) { bestDist = dist; bestQuery = qidx; bestTrain = tidx; }
                    else { bestDist2 = dist; octaveBest2 = tk->octave; }
                }
                if (bestQuery != -1) {
                    if (!(octaveBest2 == q_kps[bestQuery].octave && bestDist > bestDist2 * nn_match_ratio)) {
                        out[n].queryIdx = (int32_t)bestQuery; out[n].trainIdx = (int32_t)bestTrain; out[n].imgIdx = -1; out[n].distance = bestDist;
                        n++;
                    }
                }
            }
            ++qi;
            ++ti;
        } else if (q_node_id[qi] < t_node_id[ti]) {
            while (qi < q_nodes && q_node_id[qi] < t_node_id[ti]) ++qi;
        } else {
            while (ti < t_nodes && t_node_id[ti] < q_node_id[qi]) ++ti;
        }
    }
    free(sf2);
    return match_tail(out, n, q_kps, t_kps, checkOrientation);
}
