// project.cu — K9: projection matcher (local map -> frame), SURVEY.md 8a row a12.
//
// Replaces (reference, relative to /root/reference):
//   src/map.cpp:651-770              Map::matchFrameToMapPoints: per map point viewing-angle / depth / scale-invariance gates, pinhole
//                                    projection, predicted octave, radius search among the frame's keypoints, Hamming best / second
//                                    best with the 0.8 ratio test, filter_ambiguous_query
//   src/map_types/frame.cpp:102-115  Frame::getKeyPointsInRegion;  frame.h:129-136 predictScale;  mappoint.h:99,146-162
//   src/basictypes/picoflann.h       the frame's 2-d kd-tree: build (:150-165,240-345) on the host, radius search (:453-600) on the device
//   src/basictypes/misc.cpp:117-150  filter_ambiguous_query
// The best / second-best bookkeeping of the reference is ORDER dependent (map.cpp:722-737: a new best does not demote the old one),
// and the order is the kd-tree's visit order (nearest child first).  So the device walks the SAME tree in the same order: one
// thread per map point, explicit stack, the tree flattened to 28-byte nodes.  The tree itself is the caller's (Frame::keypoint_kdtree):
// uco_b200_kdtree_parse reads the byte stream KdTreeIndex::toStream writes, uco_b200_kdtree_build restates the build for callers
// that only hold keypoints.  Arithmetic follows the reference expression by expression (f32 without FMA, the three double-precision
// spots of cv::norm / 1./z / the kd-tree distances), see the comments.
#include "common.cuh"
#include "kdwalk.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

// ---- host: the frame's kd-tree -------------------------------------------------------------------------------------------------------
namespace {
struct KdBuilder {
    const uint8_t* base;
    size_t stride;
    std::vector<uint32_t> all;
    std::vector<uco_kdnode> nodes;
    std::vector<int32_t> leaf;
    float at(uint32_t i, int d) const { return ((const float*)(base + (size_t)i * stride))[d]; }
    void box(double bb[4], int s, int e) const {  // computeBoundingBox: float values widened
        for (int d = 0; d < 2; d++) bb[2 * d] = bb[2 * d + 1] = at(all[s], d);
        for (int k = s + 1; k < e; k++)
            for (int d = 0; d < 2; d++) {
                const float v = at(all[k], d);
                if (v < bb[2 * d]) bb[2 * d] = v;
                if (v > bb[2 * d + 1]) bb[2 * d + 1] = v;
            }
    }
    void divide(int node, int s, int e, double bb[4]) {
        const int count = e - s;
        if (count <= 10) {  // _maxLeafSize
            nodes[node].col = -1;
            nodes[node].left = nodes[node].right = -1;
            nodes[node].leaf_begin = (int32_t)leaf.size();
            nodes[node].leaf_count = count;
            for (int i = s; i < e; i++) leaf.push_back((int32_t)all[i]);
            box(bb, s, e);
            return;
        }
        const int L = (int)nodes.size();
        nodes.push_back(uco_kdnode{});
        nodes.push_back(uco_kdnode{});
        // split dimension = larger variance of a strided sample (>= 100 elements), split value = its mean
        double mean[2] = {0, 0}, sq[2] = {0, 0};
        const int inc = count >= 200 ? count / 100 : 1;
        int cnt = 0;
        for (int i = s; i < e; i += inc, cnt++)
            for (int d = 0; d < 2; d++) {
                const float v = at(all[i], d);
                mean[d] += v;
                sq[d] += v * v;  // the product is formed in float
            }
        const double ic = 1. / double(cnt);
        double var[2];
        for (int d = 0; d < 2; d++) {
            mean[d] *= ic;
            var[d] = sq[d] * ic - mean[d] * mean[d];
        }
        const int col = var[1] > var[0] ? 1 : 0;
        double div = mean[col];
        // three-way partition around the (float) cut value
        uint32_t* ind = all.data() + s;
        const float cut = (float)div;
        int a = 0, b = count - 1;
        for (;;) {
            while (a <= b && at(ind[a], col) < cut) a++;
            while (a <= b && at(ind[b], col) >= cut) b--;
            if (a > b) break;
            std::swap(ind[a++], ind[b--]);
        }
        const int lim1 = a;
        b = count - 1;
        for (;;) {
            while (a <= b && at(ind[a], col) <= cut) a++;
            while (a <= b && at(ind[b], col) > cut) b--;
            if (a > b) break;
            std::swap(ind[a++], ind[b--]);
        }
        const int lim2 = a, half = count / 2;
        int split = lim1 > half ? lim1 : (lim2 < half ? lim2 : half);
        if (lim1 == count || lim2 == 0) split = half;
        if (split < 10 || count - split < 10) {  // degenerate cut: sort the range along the dimension and cut in the middle
            std::sort(all.begin() + s, all.begin() + e, [&](const uint32_t& x, const uint32_t& y) { return at(x, col) < at(y, col); });
            split = half;
            div = at(all[s + split], col);
        }
        double lb[4], rb[4];
        memcpy(lb, bb, sizeof lb);
        lb[2 * col + 1] = div;
        divide(L, s, s + split, lb);
        lb[2 * col + 1] = div;  // picoflann.h:337 puts the split value back: divlow is the split value, divhigh the tight bound
        memcpy(rb, bb, sizeof rb);
        rb[2 * col] = div;
        divide(L + 1, s + split, e, rb);
        uco_kdnode& N = nodes[node];
        N.col = col;
        N.left = L;
        N.right = L + 1;
        N.leaf_begin = N.leaf_count = 0;
        N.divlow = (float)lb[2 * col + 1];
        N.divhigh = (float)rb[2 * col];
        for (int d = 0; d < 2; d++) {
            bb[2 * d] = std::min(lb[2 * d], rb[2 * d]);
            bb[2 * d + 1] = std::max(lb[2 * d + 1], rb[2 * d + 1]);
        }
    }
};
}  // namespace

extern "C" int uco_b200_kdtree_build(const float* xy, size_t stride_bytes, int n, uco_kdnode* nodes, int cap_nodes, int32_t* leaf_idx,
                                     double* bbox4, int* n_nodes) {
    if (n < 0 || !n_nodes || (n > 0 && (!xy || !nodes || !leaf_idx || !bbox4 || stride_bytes < 8))) return UCO_E_INVALID;
    *n_nodes = 0;
    if (n == 0) return UCO_OK;
    KdBuilder B;
    B.base = (const uint8_t*)xy;
    B.stride = stride_bytes;
    B.all.resize(n);
    for (int i = 0; i < n; i++) B.all[i] = i;
    B.nodes.reserve(2 * (size_t)n + 2);
    B.leaf.reserve(n);
    B.box(bbox4, 0, n);
    B.nodes.push_back(uco_kdnode{});
    B.divide(0, 0, n, bbox4);
    if ((int)B.nodes.size() > cap_nodes) return UCO_E_INVALID;
    memcpy(nodes, B.nodes.data(), sizeof(uco_kdnode) * B.nodes.size());
    memcpy(leaf_idx, B.leaf.data(), 4 * B.leaf.size());
    *n_nodes = (int)B.nodes.size();
    return UCO_OK;
}

// KdTreeIndex::toStream (picoflann.h:603-660): int dims, int nValues, u64 nb, nb x {double first, second}, u64 k,
// k x {double div_val, u16 col, float divhigh, float divlow, i64 left, i64 right, u64 s, s x int idx}
extern "C" int uco_b200_kdtree_parse(const void* bytes, size_t n_bytes, uco_kdnode* nodes, int cap_nodes, int32_t* leaf_idx, int cap_leaf,
                                     double* bbox4, int* n_nodes, int* n_leaf) {
    if (!bytes || !n_nodes || !n_leaf || !bbox4) return UCO_E_INVALID;
    const uint8_t* p = (const uint8_t*)bytes;
    const uint8_t* end = p + n_bytes;
    auto need = [&](size_t k) { return (size_t)(end - p) >= k; };
    auto rd = [&](void* dst, size_t k) { memcpy(dst, p, k); p += k; };
    int32_t dims, nvalues;
    uint64_t nb, k;
    if (!need(16)) return UCO_E_INVALID;
    rd(&dims, 4); rd(&nvalues, 4); rd(&nb, 8);
    if (nb > 2 || !need(16 * nb + 8)) return UCO_E_INVALID;
    bbox4[0] = bbox4[1] = bbox4[2] = bbox4[3] = 0;
    rd(bbox4, 16 * nb);
    rd(&k, 8);
    *n_nodes = *n_leaf = 0;
    if (k == 0) return UCO_OK;
    if (dims != 2 || k > (uint64_t)cap_nodes || !nodes || !leaf_idx) return UCO_E_INVALID;
    int nl = 0;
    for (uint64_t i = 0; i < k; i++) {
        if (!need(8 + 2 + 8 + 16 + 8)) return UCO_E_INVALID;
        double dv; uint16_t col; float dh, dl; int64_t l, r; uint64_t s;
        rd(&dv, 8); rd(&col, 2); rd(&dh, 4); rd(&dl, 4); rd(&l, 8); rd(&r, 8); rd(&s, 8);
        if (!need(4 * s) || nl + (int64_t)s > cap_leaf || l >= (int64_t)k || r >= (int64_t)k || col > 1) return UCO_E_INVALID;
        uco_kdnode& N = nodes[i];
        const bool is_leaf = l == -1 && r == -1;
        N.divlow = dl; N.divhigh = dh; N.col = is_leaf ? -1 : (int32_t)col; N.left = (int32_t)l; N.right = (int32_t)r;
        N.leaf_begin = is_leaf ? nl : 0; N.leaf_count = is_leaf ? (int32_t)s : 0;
        if (is_leaf) { rd(leaf_idx + nl, 4 * s); nl += (int)s; }
        else p += 4 * s;
    }
    *n_nodes = (int)k;
    *n_leaf = nl;
    return UCO_OK;
}

// ---- device --------------------------------------------------------------------------------------------------------------------------
namespace {
struct ProjDev {
    int m, n_kp, n_nodes, n_levels;
    const float *pos, *normal, *min_dist, *max_dist;
    const uint32_t* mp_desc;       // m x 8 words
    const uco_keypoint* kps;       // cv::KeyPoint layout: pt at 0, octave at 20
    const uint32_t* kp_desc;       // n_kp x 8 words
    const uco_kdnode* nodes;
    const int32_t* leaf_idx;
    double bbox[4];
    float scale_factors[UCO_ORB_MAX_LEVELS * 2];
    float fx, fy, cx, cy, min_x, min_y, max_x, max_y;
    float pose[16];
    float min_desc_dist, max_reproj_dist;
    int* best_kp;                  // m: keypoint chosen for the map point, or -1
    float* best_dist;              // m
    uint8_t* visible;              // m
    unsigned long long* kp_owner;  // n_kp: min over map points of (distance << 32 | map point index)
    int* err;                      // set when a walk overflowed its stack
};

__global__ void __launch_bounds__(128) project_match_kernel(const __grid_constant__ ProjDev D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.m) return;
    D.best_kp[i] = -1;
    D.best_dist[i] = 0.f;
    D.visible[i] = 0;
    const float* M = D.pose;
    // camCenter = pose_f2g.inv() * (0,0,0)   (se3transform.h:98-120; the three zero products are kept: they decide the sign of a zero)
    const float i0 = M[0], i1 = M[4], i2 = M[8], i4 = M[1], i5 = M[5], i6 = M[9], i8 = M[2], i9 = M[6], i10 = M[10];
    const float i3 = -(M[3] * i0 + M[7] * i1 + M[11] * i2), i7 = -(M[3] * i4 + M[7] * i5 + M[11] * i6), i11 = -(M[3] * i8 + M[7] * i9 + M[11] * i10);
    const float z0 = 0.f;
    const float ccx = i0 * z0 + i1 * z0 + i2 * z0 + i3, ccy = i4 * z0 + i5 * z0 + i6 * z0 + i7, ccz = i8 * z0 + i9 * z0 + i10 * z0 + i11;
    const float px = D.pos[3 * i], py = D.pos[3 * i + 1], pz = D.pos[3 * i + 2];
    // MapPoint::getViewCos: v = camCenter - pos3d; v *= 1./cv::norm(v) (double); v.dot(normal) (float)
    float vx = ccx - px, vy = ccy - py, vz = ccz - pz;
    const double inv = 1. / sqrt((double)vx * vx + (double)vy * vy + (double)vz * vz);
    vx = (float)(vx * inv); vy = (float)(vy * inv); vz = (float)(vz * inv);
    const float view_cos = vx * D.normal[3 * i] + vy * D.normal[3 * i + 1] + vz * D.normal[3 * i + 2];
    if (view_cos < 0.5) return;
    float cx3 = M[0] * px + M[1] * py + M[2] * pz + M[3], cy3 = M[4] * px + M[5] * py + M[6] * pz + M[7], cz3 = M[8] * px + M[9] * py + M[10] * pz + M[11];
    if (cz3 < 0) return;
    const float dist = (float)sqrt((double)cx3 * cx3 + (double)cy3 * cy3 + (double)cz3 * cz3);
    const float mxd = D.max_dist[i];
    if (!(0.8f * D.min_dist[i] < dist && dist < 1.2f * mxd)) return;
    cz3 = (float)(1. / cz3);
    const float q[2] = {cx3 * D.fx * cz3 + D.cx, cy3 * D.fy * cz3 + D.cy};
    if (!(q[0] > D.min_x && q[1] > D.min_y && q[0] < D.max_x && q[1] < D.max_y)) return;
    D.visible[i] = 1;
    // Frame::predictScale: float logs (std::log(float)); the device takes the double log rounded to float, which is the correctly
    // rounded value glibc's logf returns for all but ~1e-6 of its inputs
    int octave;
    {
        const float lsf = (float)log((double)D.scale_factors[1]);
        const float ns_f = ceilf((float)log((double)(mxd / dist)) / lsf);
        const int ns = (int)ns_f;
        octave = ns < 0 ? 0 : (ns >= D.n_levels ? D.n_levels - 1 : ns);
    }
    float radius_scale = D.scale_factors[octave];
    if (view_cos < 0.98) radius_scale = (float)(radius_scale * 1.6);
    const double radius = (double)(radius_scale * D.max_reproj_dist);
    if (D.n_nodes == 0 || !(radius > 0)) return;
    uint32_t md[8];
#pragma unroll
    for (int k = 0; k < 8; k++) md[k] = D.mp_desc[8 * (size_t)i + k];
    int best_kp = -1, best_level = 0, best_level2 = -1;
    float best = FLT_MAX, best2 = FLT_MAX;
    const bool walked = kd_radius_walk(D.nodes, D.leaf_idx, D.bbox, D.kps, q, radius, [&](int kp, const uco_keypoint& K) {
        if (!(K.octave >= octave - 1 && K.octave <= octave)) return;  // getKeyPointsInRegion's scale window
        const uint32_t* kd = D.kp_desc + 8 * (size_t)kp;
        int pc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) pc += __popc(md[w] ^ kd[w]);
        const float dsc = (float)pc;
        if (dsc < D.min_desc_dist) {   // map.cpp:722-737, order dependent on purpose
            if (dsc < best) {
                best = dsc;
                best_kp = kp;
                best_level = K.octave;
            } else if (dsc < best2) {
                best2 = dsc;
                best_level2 = K.octave;
            }
        }
    });
    if (!walked) atomicExch(D.err, 1);
    if (best_kp < 0) return;
    if (best_level2 == best_level && (double)best > 0.8 * (double)best2) return;
    D.best_kp[i] = best_kp;
    D.best_dist[i] = best;
    atomicMin(D.kp_owner + best_kp, ((unsigned long long)(unsigned)(int)best << 32) | (unsigned)i);
}

// filter_ambiguous_query (misc.cpp:117-150) + remove_unused_matches: a keypoint keeps the map point with the smallest distance, the
// earlier one on ties; the survivors come out in map-point order.  One CTA, ordered block scan.
__global__ void __launch_bounds__(1024) project_compact_kernel(const __grid_constant__ ProjDev D, const uint32_t* __restrict__ ids,
                                                               uco_match* __restrict__ out, int* __restrict__ n_out) {
    __shared__ int wsum[32];
    __shared__ int running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < D.m; base += 1024) {
        const int i = base + threadIdx.x;
        int kp = -1;
        if (i < D.m) {
            kp = D.best_kp[i];
            if (kp >= 0 && (unsigned)(D.kp_owner[kp] & 0xffffffffu) != (unsigned)i) kp = -1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, kp >= 0);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int before = running, tot = 0;
        for (int w = 0; w < 32; w++) {
            if (w < warp) before += wsum[w];
            tot += wsum[w];
        }
        if (kp >= 0) {
            uco_match mt;
            mt.queryIdx = kp;
            mt.trainIdx = (int32_t)ids[i];
            mt.imgIdx = -1;   // cv::DMatch() leaves imgIdx = -1 and the reference never sets it here
            mt.distance = D.best_dist[i];
            out[before + __popc(bal & ((1u << lane) - 1))] = mt;
        }
        __syncthreads();
        if (threadIdx.x == 0) running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = running;
}
}  // namespace

extern "C" int uco_b200_match_projected(uco_b200_ctx* ctx, const uco_mappoints* mp, const uco_frame_view* fr, const float* pose_f2g,
                                        float min_desc_dist, float max_reproj_dist, uco_match* out, int* n_out, uint8_t* visible) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!mp || !fr || !pose_f2g || !n_out) return uco_fail(ctx, UCO_E_INVALID, "match_projected: null argument");
    *n_out = 0;
    const int m = mp->n, nk = fr->n_kp;
    if (m < 0 || nk < 0 || fr->n_nodes < 0 || fr->n_levels < 2 || fr->n_levels > UCO_ORB_MAX_LEVELS * 2)
        return uco_fail(ctx, UCO_E_INVALID, "match_projected: bad sizes m=%d n_kp=%d n_nodes=%d n_levels=%d", m, nk, fr->n_nodes, fr->n_levels);
    if (m == 0) return UCO_OK;
    if (!out || !mp->ids || !mp->pos || !mp->normal || !mp->min_dist || !mp->max_dist || !mp->desc || !fr->scale_factors ||
        (nk > 0 && (!fr->kps || !fr->desc)) || (fr->n_nodes > 0 && (!fr->nodes || !fr->leaf_idx)))
        return uco_fail(ctx, UCO_E_INVALID, "match_projected: null array");
    if (nk == 0 || fr->n_nodes == 0) {  // no keypoints: nothing can match (the reference would dereference an empty tree)
        if (visible) memset(visible, 0, m);
        return UCO_OK;
    }
    // leaf indices must address keypoints (a stale tree would otherwise read out of bounds)
    int n_leaf = 0;
    for (int i = 0; i < fr->n_nodes; i++) {
        const uco_kdnode& N = fr->nodes[i];
        if (N.col < 0) n_leaf = std::max(n_leaf, N.leaf_begin + N.leaf_count);
        else if ((unsigned)N.left >= (unsigned)fr->n_nodes || (unsigned)N.right >= (unsigned)fr->n_nodes || N.col > 1 || N.left <= i || N.right <= i)  // children follow their parent in picoflann's numbering: rules out cycles
            return uco_fail(ctx, UCO_E_INVALID, "match_projected: kd-tree node %d is malformed", i);
    }
    for (int i = 0; i < n_leaf; i++)
        if ((unsigned)fr->leaf_idx[i] >= (unsigned)nk) return uco_fail(ctx, UCO_E_INVALID, "match_projected: kd-tree indexes keypoint %d of %d", fr->leaf_idx[i], nk);
    // one staging buffer -> one H2D copy
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += (b + 255) & ~(size_t)255; return o; };
    const size_t o_pos = take(12 * (size_t)m), o_nrm = take(12 * (size_t)m), o_min = take(4 * (size_t)m), o_max = take(4 * (size_t)m),
                 o_mdesc = take(32 * (size_t)m), o_ids = take(4 * (size_t)m), o_kps = take(sizeof(uco_keypoint) * (size_t)nk),
                 o_kdesc = take(32 * (size_t)nk), o_nodes = take(sizeof(uco_kdnode) * (size_t)fr->n_nodes), o_leaf = take(4 * (size_t)n_leaf + 4);
    const size_t in_bytes = off;
    const size_t o_bkp = take(4 * (size_t)m), o_bd = take(4 * (size_t)m), o_vis = take((size_t)m), o_own = take(8 * (size_t)nk),
                 o_out = take(sizeof(uco_match) * (size_t)m), o_n = take(16);
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_PROJ, off);
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_PROJ, in_bytes);
    uint8_t* ho = (uint8_t*)uco_pinned(ctx, WS_PROJ_OUT, sizeof(uco_match) * (size_t)m + (size_t)m + 64);
    if (!d || !h || !ho) return UCO_E_NOMEM;
    memcpy(h + o_pos, mp->pos, 12 * (size_t)m);
    memcpy(h + o_nrm, mp->normal, 12 * (size_t)m);
    memcpy(h + o_min, mp->min_dist, 4 * (size_t)m);
    memcpy(h + o_max, mp->max_dist, 4 * (size_t)m);
    memcpy(h + o_mdesc, mp->desc, 32 * (size_t)m);
    memcpy(h + o_ids, mp->ids, 4 * (size_t)m);
    memcpy(h + o_kps, fr->kps, sizeof(uco_keypoint) * (size_t)nk);
    {
        const size_t ds = fr->desc_stride ? fr->desc_stride : 32;
        if (ds == 32) memcpy(h + o_kdesc, fr->desc, 32 * (size_t)nk);
        else for (int i = 0; i < nk; i++) memcpy(h + o_kdesc + 32 * (size_t)i, fr->desc + ds * (size_t)i, 32);
    }
    memcpy(h + o_nodes, fr->nodes, sizeof(uco_kdnode) * (size_t)fr->n_nodes);
    memcpy(h + o_leaf, fr->leaf_idx, 4 * (size_t)n_leaf);
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_own, 0xff, 8 * (size_t)nk, s));
    ProjDev D;
    D.m = m; D.n_kp = nk; D.n_nodes = fr->n_nodes; D.n_levels = fr->n_levels;
    D.pos = (const float*)(d + o_pos); D.normal = (const float*)(d + o_nrm); D.min_dist = (const float*)(d + o_min); D.max_dist = (const float*)(d + o_max);
    D.mp_desc = (const uint32_t*)(d + o_mdesc); D.kps = (const uco_keypoint*)(d + o_kps); D.kp_desc = (const uint32_t*)(d + o_kdesc);
    D.nodes = (const uco_kdnode*)(d + o_nodes); D.leaf_idx = (const int32_t*)(d + o_leaf);
    for (int k = 0; k < 4; k++) D.bbox[k] = fr->bbox[k];
    for (int k = 0; k < UCO_ORB_MAX_LEVELS * 2; k++) D.scale_factors[k] = k < fr->n_levels ? fr->scale_factors[k] : 0.f;
    D.fx = fr->fx; D.fy = fr->fy; D.cx = fr->cx; D.cy = fr->cy;
    D.min_x = fr->min_xy[0]; D.min_y = fr->min_xy[1]; D.max_x = fr->max_xy[0]; D.max_y = fr->max_xy[1];
    memcpy(D.pose, pose_f2g, 64);
    D.min_desc_dist = min_desc_dist; D.max_reproj_dist = max_reproj_dist;
    D.best_kp = (int*)(d + o_bkp); D.best_dist = (float*)(d + o_bd); D.visible = d + o_vis; D.kp_owner = (unsigned long long*)(d + o_own);
    D.err = (int*)(d + o_n) + 1;
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_n, 0, 8, s));
    project_match_kernel<<<(m + 127) / 128, 128, 0, s>>>(D);
    UCO_LAUNCH_CHECK(ctx);
    project_compact_kernel<<<1, 1024, 0, s>>>(D, (const uint32_t*)(d + o_ids), (uco_match*)(d + o_out), (int*)(d + o_n));
    UCO_LAUNCH_CHECK(ctx);
    int* hn = (int*)(ho + sizeof(uco_match) * (size_t)m + (((size_t)m + 15) & ~(size_t)15));
    UCO_CUDA(ctx, cudaMemcpyAsync(ho, d + o_out, sizeof(uco_match) * (size_t)m, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + sizeof(uco_match) * (size_t)m, d + o_vis, (size_t)m, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(hn, d + o_n, 8, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    if (hn[1]) return uco_fail(ctx, UCO_E_CAPACITY, "match_projected: the kd-tree is deeper than the %d deferred branches the device walk keeps", KD_STACK);
    *n_out = *hn;
    memcpy(out, ho, sizeof(uco_match) * (size_t)*hn);
    if (visible) memcpy(visible, ho + sizeof(uco_match) * (size_t)m, (size_t)m);
    return UCO_OK;
}
