// project_oracle.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
// CPU restatement of the reference's projection matcher, relative to /root/reference:
//   src/map.cpp:651-770                     Map::matchFrameToMapPoints (the per-map-point loop and filter_ambiguous_query)
//   src/map_types/frame.cpp:102-115         Frame::getKeyPointsInRegion (kd-tree radius search + octave window)
//   src/map_types/frame.h:129-136           Frame::predictScale
//   src/map_types/mappoint.h:99,146-162     MapPoint::getViewCos, getHammDescDistance_2
//   src/basictypes/se3transform.h:98-120    Se3Transform::inv, operator*(Point3f)
//   src/basictypes/picoflann.h:150-165,240-345,356-450,453-600   KdTreeIndex build (mean/variance split, planeSplit, std::sort
//                                           fallback) and the radius search (nearest child first; the VISIT ORDER decides the
//                                           best / second-best bookkeeping of map.cpp:722-737, which is order dependent)
//   src/basictypes/misc.cpp:117-150         filter_ambiguous_query
// C++ only because picoflann's build calls std::sort (unstable: the libstdc++ introsort decides the order of equal coordinates).
// Pinned against the reference's own picoflann.h compiled in oracle/_ref/libref_picoflann.so (tests/test_project_oracle.py).
// OpenCV arithmetic that is not under /root/reference (cv::norm(Point3f) in double, Point3f::operator*=(double), Point3f::dot in
// float; OpenCV core types.hpp) is restated from its published definition: "parity unpinned" for those three expressions.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Node {
    double div_val = 0;
    int col = 0;
    float divhigh = 0, divlow = 0;
    int left = -1, right = -1;
    std::vector<int> idx;
};
struct Tree {
    std::vector<Node> nodes;
    double bbox[2][2] = {{0, 0}, {0, 0}};  // [dim][first, second]
    const float* xy = nullptr;
    size_t stride = 2;  // floats between consecutive points
    std::vector<uint32_t> all;
    float at(uint32_t i, int d) const { return xy[i * stride + d]; }
};

void bounding_box(const Tree& T, double bb[2][2], int start, int end) {  // picoflann.h:356-370
    for (int i = 0; i < 2; i++) bb[i][0] = bb[i][1] = T.at(T.all[start], i);
    for (int k = start + 1; k < end; k++)
        for (int i = 0; i < 2; i++) {
            float v = T.at(T.all[k], i);
            if (v < bb[i][0]) bb[i][0] = v;
            if (v > bb[i][1]) bb[i][1] = v;
        }
}

void divide(Tree& T, size_t node, int start, int end, double bbox[2][2]) {  // picoflann.h:240-345
    const int count = end - start;
    if (count <= 10) {
        T.nodes[node].idx.resize(count);
        for (int i = 0; i < count; i++) T.nodes[node].idx[i] = (int)T.all[start + i];
        bounding_box(T, bbox, start, end);
        return;
    }
    const int left = (int)T.nodes.size(), right = left + 1;
    T.nodes[node].left = left;
    T.nodes[node].right = right;
    T.nodes.push_back(Node());
    T.nodes.push_back(Node());
    // mean_var_calculate, :372-400
    double mean[2] = {0, 0}, sum2[2] = {0, 0}, var[2];
    int cnt = 0, inc = 1;
    if (count >= 200) inc = count / 100;
    for (int i = start; i < end; i += inc) {
        for (int c = 0; c < 2; c++) {
            float val = T.at(T.all[i], c);
            mean[c] += val;
            sum2[c] += val * val;   // float product accumulated into a double, as `sum2[c] += val*val` with auto val = float
        }
        cnt++;
    }
    const double invcnt = 1. / double(cnt);
    for (int c = 0; c < 2; c++) {
        mean[c] *= invcnt;
        var[c] = sum2[c] * invcnt - mean[c] * mean[c];
    }
    int col = 0;
    if (var[1] > var[0]) col = 1;
    double div_val = mean[col];
    // planeSplit with float cutval, :410-432
    uint32_t* ind = &T.all[start];
    const float cutval = (float)div_val;
    int l = 0, r = count - 1;
    for (;;) {
        while (l <= r && T.at(ind[l], col) < cutval) ++l;
        while (l <= r && T.at(ind[r], col) >= cutval) --r;
        if (l > r) break;
        std::swap(ind[l], ind[r]);
        ++l;
        --r;
    }
    const int lim1 = l;
    r = count - 1;
    for (;;) {
        while (l <= r && T.at(ind[l], col) <= cutval) ++l;
        while (l <= r && T.at(ind[r], col) > cutval) --r;
        if (l > r) break;
        std::swap(ind[l], ind[r]);
        ++l;
        --r;
    }
    const int lim2 = l;
    int split;
    if (lim1 > count / 2) split = lim1;
    else if (lim2 < count / 2) split = lim2;
    else split = count / 2;
    if (lim1 == count || lim2 == 0) split = count / 2;
    if (split < 10 || count - split < 10) {
        std::sort(T.all.begin() + start, T.all.begin() + end, [&](const uint32_t& a, const uint32_t& b) { return T.at(a, col) < T.at(b, col); });
        split = count / 2;
        div_val = T.at(T.all[start + split], col);
    }
    T.nodes[node].col = col;
    T.nodes[node].div_val = div_val;
    double lb[2][2], rb[2][2];
    memcpy(lb, bbox, sizeof lb);
    lb[col][1] = div_val;
    divide(T, left, start, start + split, lb);
    lb[col][1] = div_val;          // :337: the bound the recursion tightened is overwritten again, so divlow == (float)div_val
    memcpy(rb, bbox, sizeof rb);
    rb[col][0] = div_val;
    divide(T, right, start + split, end, rb);
    T.nodes[node].divlow = (float)lb[col][1];
    T.nodes[node].divhigh = (float)rb[col][0];
    for (int i = 0; i < 2; i++) {
        bbox[i][0] = std::min(lb[i][0], rb[i][0]);
        bbox[i][1] = std::max(lb[i][1], rb[i][1]);
    }
}

void build(Tree& T, const float* xy, size_t stride, int n) {  // picoflann.h:150-165
    T.xy = xy;
    T.stride = stride;
    T.nodes.clear();
    T.all.resize(n);
    for (int i = 0; i < n; i++) T.all[i] = i;
    if (n == 0) return;
    T.nodes.reserve(2 * (size_t)n + 4);
    bounding_box(T, T.bbox, 0, n);
    T.nodes.push_back(Node());
    divide(T, 0, 0, n, T.bbox);
}

// picoflann.h:556-600, radius search: appends the accepted indices in visit order
void search_level(const Tree& T, int node, const float q[2], double r2, double mindistsq, double dists[2], std::vector<int>& out) {
    const Node& N = T.nodes[node];
    if (N.left == -1 && N.right == -1) {
        for (int id : N.idx) {
            double sqd = 0;
            for (int i = 0; i < 2; i++) {
                double d = q[i] - T.at(id, i);  // float subtraction, then widened
                sqd += d * d;
                if (sqd > r2) break;
            }
            if (sqd < r2) out.push_back(id);
        }
        return;
    }
    const double val = q[N.col];
    const double diff1 = val - N.divlow, diff2 = val - N.divhigh;
    int best, other;
    double cut;
    if (diff1 + diff2 < 0) { best = N.left; other = N.right; cut = diff2 * diff2; }
    else { best = N.right; other = N.left; cut = diff1 * diff1; }
    search_level(T, best, q, r2, mindistsq, dists, out);
    const float dst = (float)dists[N.col];
    mindistsq = mindistsq + cut - dst;
    dists[N.col] = cut;
    if (mindistsq * 1.0 <= r2) search_level(T, other, q, r2, mindistsq, dists, out);
    dists[N.col] = dst;
}

void radius_search(const Tree& T, const float q[2], double radius, std::vector<int>& out) {  // generalSearch, :453-463
    out.clear();
    if (T.nodes.empty()) return;
    double dists[2] = {0, 0};
    const double r2 = radius > 0 ? radius * radius : -1.f;
    float distsq = 0;  // computeInitialDistances, :435-451
    for (int i = 0; i < 2; i++) {
        const double e = q[i];
        if (e < T.bbox[i][0]) { double d = e - T.bbox[i][0]; dists[i] = d * d; distsq += dists[i]; }
        if (e > T.bbox[i][1]) { double d = e - T.bbox[i][1]; dists[i] = d * d; distsq += dists[i]; }
    }
    if (!(r2 > 0)) return;  // the tracker always passes a positive radius
    search_level(T, 0, q, r2, distsq, dists, out);
}

inline float hamming32(const uint8_t* a, const uint8_t* b) {  // mappoint.h:146-162
    const uint64_t* x = (const uint64_t*)a;
    const uint64_t* y = (const uint64_t*)b;
    int s = 0;
    for (int i = 0; i < 4; i++) s += __builtin_popcountll(x[i] ^ y[i]);
    return (float)s;
}
}  // namespace

extern "C" {

// flattened copy of the tree for comparisons: per node {col, left, right, leaf_begin, leaf_count} + divlow/divhigh/div_val, leaf index list
int oracle_kdtree_build(const float* xy, int stride_floats, int n, int32_t* node_i5, float* node_f2, double* node_div, int32_t* leaf_idx,
                        double* bbox4, int cap_nodes) {
    Tree T;
    build(T, xy, stride_floats, n);
    if ((int)T.nodes.size() > cap_nodes) return -1;
    int nl = 0;
    for (size_t i = 0; i < T.nodes.size(); i++) {
        const Node& N = T.nodes[i];
        node_i5[5 * i] = N.col; node_i5[5 * i + 1] = N.left; node_i5[5 * i + 2] = N.right; node_i5[5 * i + 3] = nl; node_i5[5 * i + 4] = (int)N.idx.size();
        node_f2[2 * i] = N.divlow; node_f2[2 * i + 1] = N.divhigh;
        node_div[i] = N.div_val;
        for (int v : N.idx) leaf_idx[nl++] = v;
    }
    bbox4[0] = T.bbox[0][0]; bbox4[1] = T.bbox[0][1]; bbox4[2] = T.bbox[1][0]; bbox4[3] = T.bbox[1][1];
    return (int)T.nodes.size();
}

// radius searches in visit order: out_ptr[nq+1], out_idx (capacity cap); returns the total or -1
int oracle_kdtree_radius(const float* xy, int stride_floats, int n, const float* queries, const float* radii, int nq, int32_t* out_ptr,
                         int32_t* out_idx, int cap) {
    Tree T;
    build(T, xy, stride_floats, n);
    std::vector<int> r;
    int tot = 0;
    for (int i = 0; i < nq; i++) {
        out_ptr[i] = tot;
        radius_search(T, queries + 2 * i, radii[i], r);
        for (int v : r) {
            if (tot >= cap) return -1;
            out_idx[tot++] = v;
        }
    }
    out_ptr[nq] = tot;
    return tot;
}

struct OracleMatch { int32_t queryIdx, trainIdx, imgIdx; float distance; };  // cv::DMatch

// Map::matchFrameToMapPoints after the map-point list has been gathered (map.cpp:651-672 is container walking): m map points in the
// order of smap_ids; returns the number of matches written (queryIdx = keypoint, trainIdx = ids[mpix]); visible[mpix] = setVisible()
int oracle_match_projected(int m, const uint32_t* ids, const float* pos, const float* normal, const float* min_dist, const float* max_dist,
                           const uint8_t* mp_desc, int n_kp, const float* kp_xy, const int32_t* kp_octave, const uint8_t* kp_desc,
                           const float* scale_factors, int n_levels, float fx, float fy, float cx, float cy, const float* min_xy,
                           const float* max_xy, const float* pose, float min_desc_dist, float max_reproj_dist, OracleMatch* out,
                           uint8_t* visible) {
    Tree T;
    build(T, kp_xy, 2, n_kp);
    // camCenter = pose_f2g.inv() * (0,0,0), se3transform.h:98-120
    float Mi[16];
    Mi[0] = pose[0]; Mi[1] = pose[4]; Mi[2] = pose[8]; Mi[4] = pose[1]; Mi[5] = pose[5]; Mi[6] = pose[9]; Mi[8] = pose[2]; Mi[9] = pose[6]; Mi[10] = pose[10];
    Mi[3] = -(pose[3] * Mi[0] + pose[7] * Mi[1] + pose[11] * Mi[2]);
    Mi[7] = -(pose[3] * Mi[4] + pose[7] * Mi[5] + pose[11] * Mi[6]);
    Mi[11] = -(pose[3] * Mi[8] + pose[7] * Mi[9] + pose[11] * Mi[10]);
    const float zero = 0.f;
    const float cc[3] = {Mi[0] * zero + Mi[1] * zero + Mi[2] * zero + Mi[3], Mi[4] * zero + Mi[5] * zero + Mi[6] * zero + Mi[7],
                         Mi[8] * zero + Mi[9] * zero + Mi[10] * zero + Mi[11]};
    std::vector<OracleMatch> matches;
    std::vector<int> region;
    for (int i = 0; i < m; i++) {
        if (visible) visible[i] = 0;
        const float* P = pos + 3 * i;
        const float* Nn = normal + 3 * i;
        // getViewCos: v = camCenter - pos3d; v *= 1./cv::norm(v); v.dot(normal)
        float v[3] = {cc[0] - P[0], cc[1] - P[1], cc[2] - P[2]};
        const double nv = std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
        const double inv = 1. / nv;
        for (int k = 0; k < 3; k++) v[k] = (float)(v[k] * inv);
        const float view_cos = v[0] * Nn[0] + v[1] * Nn[1] + v[2] * Nn[2];
        if (view_cos < 0.5) continue;
        float p3[3] = {pose[0] * P[0] + pose[1] * P[1] + pose[2] * P[2] + pose[3], pose[4] * P[0] + pose[5] * P[1] + pose[6] * P[2] + pose[7],
                       pose[8] * P[0] + pose[9] * P[1] + pose[10] * P[2] + pose[11]};
        if (p3[2] < 0) continue;
        const float dist = (float)std::sqrt((double)p3[0] * p3[0] + (double)p3[1] * p3[1] + (double)p3[2] * p3[2]);
        if (!(0.8f * min_dist[i] < dist && dist < 1.2f * max_dist[i])) continue;
        p3[2] = (float)(1. / p3[2]);
        const float p2[2] = {p3[0] * fx * p3[2] + cx, p3[1] * fy * p3[2] + cy};
        if (!(p2[0] > min_xy[0] && p2[1] > min_xy[1] && p2[0] < max_xy[0] && p2[1] < max_xy[1])) continue;
        if (visible) visible[i] = 1;
        // predictScale, frame.h:129-136 (float log: `using namespace std` is in force there)
        int octave;
        {
            const float lsf = std::log(scale_factors[1]);
            const int ns = (int)std::ceil(std::log(max_dist[i] / dist) / lsf);
            octave = ns < 0 ? 0 : (ns >= n_levels ? n_levels - 1 : ns);
        }
        float radius_scale = scale_factors[octave];
        if (view_cos < 0.98) radius_scale *= 1.6;
        radius_search(T, p2, radius_scale * max_reproj_dist, region);
        int best_kp = -1, best_level = 0, best_level2 = -1;
        float best = std::numeric_limits<float>::max(), best2 = std::numeric_limits<float>::max();
        for (int kp : region) {
            if (!(kp_octave[kp] >= octave - 1 && kp_octave[kp] <= octave)) continue;
            const float d = hamming32(mp_desc + 32 * (size_t)i, kp_desc + 32 * (size_t)kp);
            if (d < min_desc_dist) {
                if (d < best) {
                    best = d;
                    best_kp = kp;
                    best_level = kp_octave[kp];
                } else if (d < best2) {
                    best2 = d;
                    best_level2 = kp_octave[kp];
                }
            }
        }
        if (best_kp != -1) {
            bool valid = true;
            if (best_level2 == best_level && best > 0.8 * best2) valid = false;
            if (valid) matches.push_back({best_kp, (int32_t)ids[i], -1, best});   // cv::DMatch() leaves imgIdx = -1
        }
    }
    // filter_ambiguous_query, misc.cpp:117-150
    if (!matches.empty()) {
        int maxq = -1;
        for (auto& mm : matches) maxq = std::max(maxq, mm.queryIdx);
        std::vector<int> used(maxq + 1, -1);
        int idx = 0;
        for (auto& mm : matches) {
            if (used[mm.queryIdx] == -1) used[mm.queryIdx] = idx;
            else if (matches[used[mm.queryIdx]].distance > mm.distance) {
                matches[used[mm.queryIdx]].queryIdx = -1;
                used[mm.queryIdx] = idx;
            } else mm.queryIdx = -1;
            idx++;
        }
    }
    int n = 0;
    for (auto& mm : matches)
        if (mm.queryIdx != -1) out[n++] = mm;   // remove_unused_matches keeps the order
    return n;
}

// filter_ambiguous_query + remove_unused_matches (misc.cpp:105-150) on a match list, in place; returns the new count
int oracle_filter_ambiguous_query(OracleMatch* m, int n) {
    if (n == 0) return 0;
    int maxq = -1;
    for (int i = 0; i < n; i++) maxq = std::max(maxq, m[i].queryIdx);
    std::vector<int> used(maxq + 1, -1);
    for (int i = 0; i < n; i++) {
        OracleMatch& mm = m[i];
        if (used[mm.queryIdx] == -1) used[mm.queryIdx] = i;
        else if (m[used[mm.queryIdx]].distance > mm.distance) {
            m[used[mm.queryIdx]].queryIdx = -1;
            used[mm.queryIdx] = i;
        } else mm.queryIdx = -1;
    }
    int k = 0;
    for (int i = 0; i < n; i++)
        if (m[i].queryIdx != -1 && m[i].trainIdx != -1) m[k++] = m[i];
    return k;
}

// The tracker's search by projection from the previous frame, System::_11946837405316294395 (src/utils/system.cpp:5921-6456, macro-obfuscated;
// called first thing in tracking, :6559, with dist_thr = maxDescDistance*1.5 and proj_dist_thr = Params::projDistThr):
// every keypoint of the PREVIOUS frame that carries a valid, non-bad map point (prev_row[i] >= 0: row of that point in ids/pos) is
// projected with the current frame's pose guess (Frame::project(p, true, true), frame.h:140-161), the current frame's keypoints of
// the SAME octave within proj_dist_thr * scaleFactors[octave] are scanned in kd-tree visit order with the order-dependent best /
// second-best bookkeeping of the reference (a new best does not demote the old one), accepted when best < 0.7 * second, then
// filter_ambiguous_query.  queryIdx = current keypoint, trainIdx = map point id, distance = Hamming.
int oracle_track_projected(int n_prev, const int32_t* prev_octave, const uint8_t* prev_desc, const int32_t* prev_row, const uint32_t* ids,
                           const float* pos, int n_kp, const float* kp_xy, const int32_t* kp_octave, const uint8_t* kp_desc,
                           const float* scale_factors, int n_levels, float fx, float fy, float cx, float cy, const float* min_xy,
                           const float* max_xy, const float* pose, float dist_thr, float proj_dist_thr, OracleMatch* out) {
    (void)n_levels;
    Tree T;
    build(T, kp_xy, 2, n_kp);
    std::vector<OracleMatch> matches;
    std::vector<int> region;
    const float nanv = std::numeric_limits<float>::quiet_NaN();
    for (int i = 0; i < n_prev; i++) {
        const int row = prev_row[i];
        if (row < 0) continue;
        const float* P = pos + 3 * (size_t)row;
        // Frame::project(p3d, setNanIfDepthNegative = true, setNanIfDoNotProjectInImage = true)
        float p2[2];
        {
            float rz = P[0] * pose[8] + P[1] * pose[9] + P[2] * pose[10] + pose[11];
            if (rz < 0) p2[0] = p2[1] = nanv;
            else {
                const float rx = P[0] * pose[0] + P[1] * pose[1] + P[2] * pose[2] + pose[3];
                const float ry = P[0] * pose[4] + P[1] * pose[5] + P[2] * pose[6] + pose[7];
                rz = (float)(1. / rz);
                p2[0] = ((fx * rx) * rz) + cx;
                p2[1] = ((fy * ry) * rz) + cy;
                if (!(p2[0] >= min_xy[0] && p2[1] >= min_xy[1] && p2[0] < max_xy[0] && p2[1] < max_xy[1])) p2[0] = p2[1] = nanv;
            }
        }
        if (std::isnan(p2[0])) continue;
        const int oct = prev_octave[i];
        const float scale = scale_factors[oct];
        radius_search(T, p2, proj_dist_thr * scale, region);
        float best = (float)(dist_thr + 0.01), best2 = std::numeric_limits<float>::max();
        int best_kp = -1;
        for (int kp : region) {
            if (kp_octave[kp] != oct) continue;   // getKeyPointsInRegion(.., oct, oct) and the explicit test are the same condition
            const float d = hamming32(prev_desc + 32 * (size_t)i, kp_desc + 32 * (size_t)kp);
            if (d < best) {
                best = d;
                best_kp = kp;
            } else if (d < best2) best2 = d;
        }
        if (best_kp != -1 && best < 0.7 * best2) matches.push_back({best_kp, (int32_t)ids[row], -1, best});
    }
    int n = (int)matches.size();
    n = oracle_filter_ambiguous_query(matches.data(), n);
    for (int i = 0; i < n; i++) out[i] = matches[i];
    return n;
}

}  // extern "C"
