// track.cu — K17: the tracker's per-frame sequence for a batch of independent frames (cameras / streams), device resident.
//
// Replaces (reference, relative to /root/reference; src/utils/system.cpp is macro-obfuscated, lines are original file lines):
//   src/utils/system.cpp:5921-6456   System::_11946837405316294395 — search by projection from the PREVIOUS frame ("a12's twin"): every
//                                    keypoint of the previous frame that carries a valid map point is projected with the current
//                                    pose guess (Frame::project, src/map_types/frame.h:140-161), the current frame's keypoints of the
//                                    SAME octave within projDistThr * scaleFactors[octave] (Frame::getKeyPointsInRegion,
//                                    src/map_types/frame.cpp:102-115) are scanned in kd-tree visit order, best / second best with
//                                    the order-dependent bookkeeping, accepted when best < 0.7 * second, filter_ambiguous_query
//   src/utils/system.cpp:6460-6960   System::_11166622111371682966 (the tracker), main branch: that search (distance threshold
//                                    maxDescDistance*1.5) -> PnPSolver::solvePnp when > 30 matches -> on > 30 inliers take the pose,
//                                    mark the matched points as seen and search the local map with a 4 px radius, otherwise drop
//                                    the matches and search with projDistThr -> Map::matchFrameToMapPoints (src/map.cpp:651-770,
//                                    points already seen in this frame skipped :659-667, threshold maxDescDistance*2) -> append,
//                                    filter_ambiguous_query (src/basictypes/misc.cpp:117-150) -> PnPSolver::solvePnp
//   src/optimization/pnpsolver.cpp:196-259   the edges solvePnp builds from a match list (the kernel of pnp.cu does the solve)
// The branch the reference takes when the first search finds <= 30 matches (FrameMatcher against the reference keyframe,
// system.cpp:6600-6700) needs that keyframe's descriptors and is the caller's: such frames are flagged in `status` bit 0 and
// continue like the reference does when that fallback also fails (no matches, search radius projDistThr).
//
// Every frame of the batch is an independent tracking problem (its own previous frame, map block and pose prior), so the whole
// sequence is eight launches for the batch: kd-trees (kdtree.cu) -> tbp -> compact + edge assembly -> pose-only LM (pnp.cu) ->
// local-map search -> merge + filter + edge assembly -> pose-only LM -> finalize.  No host round trip in between.
#include "common.cuh"
#include "kdwalk.cuh"
#include "pnp_dev.cuh"
#include <cfloat>
#include <cmath>
#include <cstring>
#include <new>

int uco_kdtree_build_launch(uco_b200_ctx* ctx, int n_frames, const uco_keypoint* kps_dev, size_t kps_stride, const int32_t* n_kp_dev,
                            int n_fixed, int cap, uco_kdnode* nodes_dev, int node_cap, int32_t* leaf_dev, double* bbox_dev,
                            int32_t* n_nodes_dev, int32_t* err_dev);

namespace {

struct TrackDev {
    int F, kp_cap, prev_cap, map_cap, node_cap;
    // current frames
    const uco_keypoint* kps;
    const uint32_t* desc;
    const int32_t* n_kp;
    const float* depth;           // nullable
    const uco_kdnode* nodes;
    const int32_t* leaf;
    const double* bbox;
    const int32_t* n_nodes;
    // previous frames
    const uco_keypoint* prev_kps;
    const uint32_t* prev_desc;
    const int32_t* prev_n;
    const int32_t* prev_row;
    // map blocks
    const int32_t* map_n;
    const uint32_t* mp_id;
    const float *mp_pos, *mp_normal, *mp_min, *mp_max;
    const uint32_t* mp_desc;
    const uint8_t *mp_stable, *mp_local;   // nullable: all stable / all local
    const float* pose_prior;
    float fx, fy, cx, cy, bf, min_x, min_y, max_x, max_y;
    int n_levels;
    float sf[UCO_MATCH_MAX_SCALES];
    float thr1, thr2, proj_dist_thr;       // maxDescDistance*1.5, maxDescDistance*2, projDistThr
    int stage1_only;
    // scratch
    int* best_kp1; float* best_d1;         // F x prev_cap
    unsigned long long* owner1;            // F x kp_cap
    int* best_kp2; float* best_d2;         // F x map_cap
    unsigned long long* owner2;            // F x kp_cap
    int* kpsel;                            // F x kp_cap
    uint8_t* seen;                         // F x map_cap
    uco_match* m1; int* row1; int* n1;     // stage-1 list, F x kp_cap
    // pose-only problems
    PnpHead* heads;                        // 2 x F
    PnpOut* outs;                          // 2 x F
    float *p_pts, *p_uv, *p_ur, *p_isig;   // F x kp_cap
    uint8_t* p_flg;
    uint8_t *bad1, *bad2;
    // outputs
    uco_match* matches;                    // F x kp_cap
    int* n_matches;
    float* pose_out;
    int *n_good, *status, *n_tbp;
    uint8_t* visible;                      // F x map_cap, nullable
    int* err;
};

// ordered block-wide compaction helper: exclusive rank of `flag` among the flags of all threads with a lower index, plus the total
__device__ __forceinline__ int block_rank(bool flag, int* wsum, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    __syncthreads();  // wsum of the previous round has been consumed
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = 0, tot = 0;
    for (int w = 0; w < nw; w++) {
        const int v = wsum[w];
        if (w < warp) before += v;
        tot += v;
    }
    *total = tot;
    return before + __popc(bal & ((1u << lane) - 1));
}

__device__ __forceinline__ void write_edge(const TrackDev& D, int f, int slot, int kp, int row) {
    const size_t g = (size_t)f * D.kp_cap + slot;
    const size_t mr = (size_t)f * D.map_cap + row;
    const uco_keypoint K = D.kps[(size_t)f * D.kp_cap + kp];
    D.p_pts[3 * g] = D.mp_pos[3 * mr]; D.p_pts[3 * g + 1] = D.mp_pos[3 * mr + 1]; D.p_pts[3 * g + 2] = D.mp_pos[3 * mr + 2];
    D.p_uv[2 * g] = K.x; D.p_uv[2 * g + 1] = K.y;
    int oct = K.octave;
    oct = oct < 0 ? 0 : (oct >= D.n_levels ? D.n_levels - 1 : oct);
    D.p_isig[g] = (float)(1. / (double)D.sf[oct]);     // invScaleFactor, pnpsolver.cpp:192-193 / :246
    const float depth = D.depth ? D.depth[(size_t)f * D.kp_cap + kp] : 0.f;
    const bool stereo = depth > 0;
    D.p_ur[g] = stereo ? K.x - D.bf / depth : 0.f;     // :226
    const bool stable = !D.mp_stable || D.mp_stable[mr];
    D.p_flg[g] = (uint8_t)((stereo ? 1 : 0) | (stable ? 2 : 0));
}

__device__ __forceinline__ void write_head(const TrackDev& D, PnpHead* h, int f, int n, const float* pose) {
    h->n = n; h->nm = 0; h->off = f * D.kp_cap; h->moff = 0;
    for (int k = 0; k < 16; k++) h->pose44[k] = pose[k];
    h->fx = D.fx; h->fy = D.fy; h->cx = D.cx; h->cy = D.cy; h->bf = D.bf; h->wm = 0;
}

// ---- stage 1: search by projection from the previous frame ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tbp_kernel(const __grid_constant__ TrackDev D) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.prev_n[f] || i >= D.prev_cap) return;
    const size_t gi = (size_t)f * D.prev_cap + i;
    D.best_kp1[gi] = -1;
    const int row = D.prev_row[gi];
    if (row < 0 || row >= D.map_n[f] || D.n_nodes[f] == 0) return;
    const float* M = D.pose_prior + 16 * (size_t)f;
    const float* P = D.mp_pos + 3 * ((size_t)f * D.map_cap + row);
    // Frame::project(p3d, true, true), frame.h:140-161
    float rz = P[0] * M[8] + P[1] * M[9] + P[2] * M[10] + M[11];
    if (rz < 0) return;
    const float rx = P[0] * M[0] + P[1] * M[1] + P[2] * M[2] + M[3];
    const float ry = P[0] * M[4] + P[1] * M[5] + P[2] * M[6] + M[7];
    rz = (float)(1. / rz);
    const float q[2] = {((D.fx * rx) * rz) + D.cx, ((D.fy * ry) * rz) + D.cy};
    if (!(q[0] >= D.min_x && q[1] >= D.min_y && q[0] < D.max_x && q[1] < D.max_y)) return;
    const int oct = D.prev_kps[gi].octave;
    if (oct < 0 || oct >= D.n_levels) return;   // the reference would index scaleFactors out of range
    const double radius = (double)(D.proj_dist_thr * D.sf[oct]);
    if (!(radius > 0)) return;
    uint32_t pd[8];
#pragma unroll
    for (int k = 0; k < 8; k++) pd[k] = D.prev_desc[8 * gi + k];
    float best = (float)((double)D.thr1 + 0.01), best2 = FLT_MAX;
    int best_kp = -1;
    const uint32_t* cdesc = D.desc + 8 * (size_t)f * D.kp_cap;
    const bool walked = kd_radius_walk(D.nodes + (size_t)f * D.node_cap, D.leaf + (size_t)f * D.kp_cap, D.bbox + 4 * (size_t)f,
                                       D.kps + (size_t)f * D.kp_cap, q, radius, [&](int kp, const uco_keypoint& K) {
        if (K.octave != oct) return;
        const uint32_t* kd = cdesc + 8 * (size_t)kp;
        int pc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) pc += __popc(pd[w] ^ kd[w]);
        const float d = (float)pc;
        if (d < best) {
            best = d;
            best_kp = kp;
        } else if (d < best2) best2 = d;
    });
    if (!walked) atomicExch(D.err, 1);
    if (best_kp < 0 || !((double)best < 0.7 * (double)best2)) return;
    D.best_kp1[gi] = best_kp;
    D.best_d1[gi] = best;
    atomicMin(D.owner1 + (size_t)f * D.kp_cap + best_kp, ((unsigned long long)(unsigned)(int)best << 32) | (unsigned)i);
}

// filter_ambiguous_query over the stage-1 candidates (a keypoint keeps the closest previous keypoint, the earlier one on ties),
// survivors in previous-keypoint order, and the pose-only problem of the first solvePnp (only posed when > 30 matches)
__global__ void __launch_bounds__(1024) tbp_compact_kernel(const __grid_constant__ TrackDev D) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int f = blockIdx.x;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int np = min(D.prev_n[f], D.prev_cap);
    for (int base = 0; base < np; base += 1024) {
        const int i = base + threadIdx.x;
        int kp = -1;
        size_t gi = (size_t)f * D.prev_cap + i;
        if (i < np) {
            kp = D.best_kp1[gi];
            if (kp >= 0 && (unsigned)(D.owner1[(size_t)f * D.kp_cap + kp] & 0xffffffffu) != (unsigned)i) kp = -1;
        }
        int tot;
        const int r = block_rank(kp >= 0, wsum, &tot);
        if (kp >= 0) {
            const int slot = running + r;
            const int row = D.prev_row[gi];
            uco_match mt;
            mt.queryIdx = kp; mt.trainIdx = (int32_t)D.mp_id[(size_t)f * D.map_cap + row]; mt.imgIdx = -1; mt.distance = D.best_d1[gi];
            D.m1[(size_t)f * D.kp_cap + slot] = mt;
            D.row1[(size_t)f * D.kp_cap + slot] = row;
            D.seen[(size_t)f * D.map_cap + row] = 1;
            if (!D.stage1_only) write_edge(D, f, slot, kp, row);
            else D.matches[(size_t)f * D.kp_cap + slot] = mt;
        }
        __syncthreads();
        if (threadIdx.x == 0) running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        D.n1[f] = running;
        D.n_tbp[f] = running;
        if (D.stage1_only) D.n_matches[f] = running;
        else write_head(D, D.heads + f, f, running > 30 ? running : 0, D.pose_prior + 16 * (size_t)f);
    }
}

__device__ __forceinline__ bool stage1_ok(const TrackDev& D, int f) { return D.n1[f] > 30 && D.outs[f].n_good > 30; }

// ---- stage 2: Map::matchFrameToMapPoints over the frame's local map points, with the pose and radius stage 1 decided -----------------
__global__ void __launch_bounds__(128) local_map_kernel(const __grid_constant__ TrackDev D) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.map_n[f] || i >= D.map_cap) return;
    const size_t gi = (size_t)f * D.map_cap + i;
    D.best_kp2[gi] = -1;
    if (D.visible) D.visible[gi] = 0;
    if (D.mp_local && !D.mp_local[gi]) return;
    const bool ok1 = stage1_ok(D, f);
    if (ok1 && D.seen[gi]) return;   // lastFIdxSeen == curframe.fseq_idx, map.cpp:659-667
    const float* M = ok1 ? D.outs[f].pose44 : D.pose_prior + 16 * (size_t)f;
    const float max_reproj = ok1 ? 4.f : D.proj_dist_thr;
    // camCenter = pose_f2g.inv() * (0,0,0)   (se3transform.h:98-120; the zero products are kept: they decide the sign of a zero)
    const float i0 = M[0], i1 = M[4], i2 = M[8], i4 = M[1], i5 = M[5], i6 = M[9], i8 = M[2], i9 = M[6], i10 = M[10];
    const float i3 = -(M[3] * i0 + M[7] * i1 + M[11] * i2), i7 = -(M[3] * i4 + M[7] * i5 + M[11] * i6), i11 = -(M[3] * i8 + M[7] * i9 + M[11] * i10);
    const float z0 = 0.f;
    const float ccx = i0 * z0 + i1 * z0 + i2 * z0 + i3, ccy = i4 * z0 + i5 * z0 + i6 * z0 + i7, ccz = i8 * z0 + i9 * z0 + i10 * z0 + i11;
    const float px = D.mp_pos[3 * gi], py = D.mp_pos[3 * gi + 1], pz = D.mp_pos[3 * gi + 2];
    float vx = ccx - px, vy = ccy - py, vz = ccz - pz;
    const double inv = 1. / sqrt((double)vx * vx + (double)vy * vy + (double)vz * vz);
    vx = (float)(vx * inv); vy = (float)(vy * inv); vz = (float)(vz * inv);
    const float view_cos = vx * D.mp_normal[3 * gi] + vy * D.mp_normal[3 * gi + 1] + vz * D.mp_normal[3 * gi + 2];
    if (view_cos < 0.5) return;
    float cx3 = M[0] * px + M[1] * py + M[2] * pz + M[3], cy3 = M[4] * px + M[5] * py + M[6] * pz + M[7], cz3 = M[8] * px + M[9] * py + M[10] * pz + M[11];
    if (cz3 < 0) return;
    const float dist = (float)sqrt((double)cx3 * cx3 + (double)cy3 * cy3 + (double)cz3 * cz3);
    const float mxd = D.mp_max[gi];
    if (!(0.8f * D.mp_min[gi] < dist && dist < 1.2f * mxd)) return;
    cz3 = (float)(1. / cz3);
    const float q[2] = {cx3 * D.fx * cz3 + D.cx, cy3 * D.fy * cz3 + D.cy};
    if (!(q[0] > D.min_x && q[1] > D.min_y && q[0] < D.max_x && q[1] < D.max_y)) return;
    if (D.visible) D.visible[gi] = 1;
    int octave;
    {   // Frame::predictScale, frame.h:129-136 (see project.cu for the rounding note)
        const float lsf = (float)log((double)D.sf[1]);
        const float ns_f = ceilf((float)log((double)(mxd / dist)) / lsf);
        const int ns = (int)ns_f;
        octave = ns < 0 ? 0 : (ns >= D.n_levels ? D.n_levels - 1 : ns);
    }
    float radius_scale = D.sf[octave];
    if (view_cos < 0.98) radius_scale = (float)(radius_scale * 1.6);
    const double radius = (double)(radius_scale * max_reproj);
    if (D.n_nodes[f] == 0 || !(radius > 0)) return;
    uint32_t md[8];
#pragma unroll
    for (int k = 0; k < 8; k++) md[k] = D.mp_desc[8 * gi + k];
    int best_kp = -1, best_level = 0, best_level2 = -1;
    float best = FLT_MAX, best2 = FLT_MAX;
    const uint32_t* cdesc = D.desc + 8 * (size_t)f * D.kp_cap;
    const bool walked = kd_radius_walk(D.nodes + (size_t)f * D.node_cap, D.leaf + (size_t)f * D.kp_cap, D.bbox + 4 * (size_t)f,
                                       D.kps + (size_t)f * D.kp_cap, q, radius, [&](int kp, const uco_keypoint& K) {
        if (!(K.octave >= octave - 1 && K.octave <= octave)) return;
        const uint32_t* kd = cdesc + 8 * (size_t)kp;
        int pc = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) pc += __popc(md[w] ^ kd[w]);
        const float dsc = (float)pc;
        if (dsc < D.thr2) {
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
    D.best_kp2[gi] = best_kp;
    D.best_d2[gi] = best;
    atomicMin(D.owner2 + (size_t)f * D.kp_cap + best_kp, ((unsigned long long)(unsigned)(int)best << 32) | (unsigned)i);
}

// matchFrameToMapPoints' own filter_ambiguous_query, then the tracker's: stage-1 matches (when kept) followed by the local-map
// matches, a keypoint claimed twice keeps the smaller distance, the earlier entry on ties; survivors in list order; edge assembly
__global__ void __launch_bounds__(1024) merge_kernel(const __grid_constant__ TrackDev D) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int f = blockIdx.x;
    const bool ok1 = stage1_ok(D, f);
    const int n1 = ok1 ? D.n1[f] : 0;
    const int nm = min(D.map_n[f], D.map_cap);
    int* kpsel = D.kpsel + (size_t)f * D.kp_cap;
    const uco_match* m1 = D.m1 + (size_t)f * D.kp_cap;
    if (threadIdx.x == 0) running = 0;
    for (int j = threadIdx.x; j < n1; j += 1024) kpsel[m1[j].queryIdx] = j;
    __syncthreads();
    // a local-map match that survives its own filter meets the stage-1 match of the same keypoint (if any)
    for (int i = threadIdx.x; i < nm; i += 1024) {
        const size_t gi = (size_t)f * D.map_cap + i;
        int kp = D.best_kp2[gi];
        if (kp < 0) continue;
        if ((unsigned)(D.owner2[(size_t)f * D.kp_cap + kp] & 0xffffffffu) != (unsigned)i) { D.best_kp2[gi] = -1; continue; }
        const int j = kpsel[kp];
        if (j >= 0) {
            if (m1[j].distance > D.best_d2[gi]) kpsel[kp] = -2 - j;   // the stage-1 match is annulled
            else D.best_kp2[gi] = -1;
        }
    }
    __syncthreads();
    for (int base = 0; base < n1; base += 1024) {
        const int j = base + threadIdx.x;
        bool keep = false;
        if (j < n1) keep = kpsel[m1[j].queryIdx] == j;
        int tot;
        const int r = block_rank(keep, wsum, &tot);
        if (keep) {
            const int slot = running + r;
            uco_match mt = m1[j];
            mt.imgIdx = D.bad1[(size_t)f * D.kp_cap + j] ? -1 : 1;   // pnpsolver.cpp:398-404
            D.matches[(size_t)f * D.kp_cap + slot] = mt;
            write_edge(D, f, slot, mt.queryIdx, D.row1[(size_t)f * D.kp_cap + j]);
        }
        __syncthreads();
        if (threadIdx.x == 0) running += tot;
        __syncthreads();
    }
    for (int base = 0; base < nm; base += 1024) {
        const int i = base + threadIdx.x;
        int kp = -1;
        const size_t gi = (size_t)f * D.map_cap + i;
        if (i < nm) kp = D.best_kp2[gi];
        int tot;
        const int r = block_rank(kp >= 0, wsum, &tot);
        if (kp >= 0) {
            const int slot = running + r;
            uco_match mt;
            mt.queryIdx = kp; mt.trainIdx = (int32_t)D.mp_id[gi]; mt.imgIdx = -1; mt.distance = D.best_d2[gi];
            D.matches[(size_t)f * D.kp_cap + slot] = mt;
            write_edge(D, f, slot, kp, i);
        }
        __syncthreads();
        if (threadIdx.x == 0) running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        D.n_matches[f] = running;
        write_head(D, D.heads + D.F + f, f, running, ok1 ? D.outs[f].pose44 : D.pose_prior + 16 * (size_t)f);
    }
}

__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ TrackDev D) {
    const int f = blockIdx.x;
    const PnpOut& o = D.outs[D.F + f];
    const int n = D.n_matches[f];
    for (int j = threadIdx.x; j < n; j += blockDim.x)
        D.matches[(size_t)f * D.kp_cap + j].imgIdx = D.bad2[(size_t)f * D.kp_cap + j] ? -1 : 1;
    if (threadIdx.x < 16) D.pose_out[16 * (size_t)f + threadIdx.x] = o.pose44[threadIdx.x];
    if (threadIdx.x == 0) {
        D.n_good[f] = n ? o.n_good : 0;
        const int n1 = D.n1[f];
        D.status[f] = (n1 <= 30 ? 1 : 0) | ((n1 > 30 && D.outs[f].n_good <= 30) ? 2 : 0);
    }
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct Scratch {
    size_t o_bk1, o_bd1, o_own1, o_own2, o_kpsel, o_bk2, o_bd2, o_seen, o_m1, o_row1, o_n1, o_heads, o_outs, o_pts, o_uv, o_ur, o_isig,
        o_flg, o_bad1, o_bad2, o_chi2, o_act, o_err, total;
};
Scratch plan_scratch(int F, int kp_cap, int prev_cap, int map_cap) {
    Scratch S;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += al(b); return o; };
    const size_t K = (size_t)F * kp_cap, P = (size_t)F * prev_cap, M = (size_t)F * map_cap;
    S.o_own1 = take(8 * K); S.o_own2 = take(8 * K); S.o_kpsel = take(4 * K);   // contiguous: one 0xff fill
    S.o_seen = take(M); S.o_err = take(16);                                     // contiguous: one zero fill
    S.o_bk1 = take(4 * P); S.o_bd1 = take(4 * P); S.o_bk2 = take(4 * M); S.o_bd2 = take(4 * M);
    S.o_m1 = take(16 * K); S.o_row1 = take(4 * K); S.o_n1 = take(4 * (size_t)F);
    S.o_heads = take(sizeof(PnpHead) * 2 * (size_t)F); S.o_outs = take(sizeof(PnpOut) * 2 * (size_t)F);
    S.o_pts = take(12 * K); S.o_uv = take(8 * K); S.o_ur = take(4 * K); S.o_isig = take(4 * K); S.o_flg = take(K);
    S.o_bad1 = take(K); S.o_bad2 = take(K); S.o_chi2 = take(8 * K); S.o_act = take(K);
    S.total = off;
    return S;
}

int check_params(uco_b200_ctx* ctx, const uco_track_params* p) {
    if (!p) return uco_fail(ctx, UCO_E_INVALID, "track: null params");
    if (p->n_levels < 2 || p->n_levels > UCO_MATCH_MAX_SCALES) return uco_fail(ctx, UCO_E_INVALID, "track: n_levels %d outside [2, %d]", p->n_levels, UCO_MATCH_MAX_SCALES);
    return UCO_OK;
}

// scratch pointers + the two fills (before anything that reports into the error words)
int setup_scratch(uco_b200_ctx* ctx, TrackDev& D, uint8_t* scr, const Scratch& S) {
    cudaStream_t s = ctx->stream;
    D.owner1 = (unsigned long long*)(scr + S.o_own1); D.owner2 = (unsigned long long*)(scr + S.o_own2); D.kpsel = (int*)(scr + S.o_kpsel);
    D.seen = scr + S.o_seen; D.err = (int*)(scr + S.o_err);
    D.best_kp1 = (int*)(scr + S.o_bk1); D.best_d1 = (float*)(scr + S.o_bd1); D.best_kp2 = (int*)(scr + S.o_bk2); D.best_d2 = (float*)(scr + S.o_bd2);
    D.m1 = (uco_match*)(scr + S.o_m1); D.row1 = (int*)(scr + S.o_row1); D.n1 = (int*)(scr + S.o_n1);
    D.heads = (PnpHead*)(scr + S.o_heads); D.outs = (PnpOut*)(scr + S.o_outs);
    D.p_pts = (float*)(scr + S.o_pts); D.p_uv = (float*)(scr + S.o_uv); D.p_ur = (float*)(scr + S.o_ur); D.p_isig = (float*)(scr + S.o_isig);
    D.p_flg = scr + S.o_flg; D.bad1 = scr + S.o_bad1; D.bad2 = scr + S.o_bad2;
    UCO_CUDA(ctx, cudaMemsetAsync(scr + S.o_own1, 0xff, S.o_seen - S.o_own1, s));
    UCO_CUDA(ctx, cudaMemsetAsync(scr + S.o_seen, 0, S.o_bk1 - S.o_seen, s));
    return UCO_OK;
}

// the launches on device-resident inputs; the trees must already be in D (nodes / leaf / bbox / n_nodes)
int run_track(uco_b200_ctx* ctx, TrackDev& D, uint8_t* scr, const Scratch& S) {
    cudaStream_t s = ctx->stream;
    const dim3 g1((D.prev_cap + 127) / 128, D.F), g2((D.map_cap + 127) / 128, D.F);
    tbp_kernel<<<g1, 128, 0, s>>>(D);
    UCO_LAUNCH_CHECK(ctx);
    tbp_compact_kernel<<<D.F, 1024, 0, s>>>(D);
    UCO_LAUNCH_CHECK(ctx);
    if (D.stage1_only) return UCO_OK;
    PnpArrays A;
    A.pts = D.p_pts; A.uv = D.p_uv; A.ur = D.p_ur; A.isig = D.p_isig; A.flg = D.p_flg;
    A.mpose = nullptr; A.msize = nullptr; A.mobs = nullptr; A.mchi2 = nullptr; A.mrobust = nullptr;
    A.chi2 = (double*)(scr + S.o_chi2); A.active = scr + S.o_act;
    A.bad = D.bad1;
    int rc = uco_pnp_launch_dev(ctx, D.F, D.heads, A, D.outs);
    if (rc != UCO_OK) return rc;
    local_map_kernel<<<g2, 128, 0, s>>>(D);
    UCO_LAUNCH_CHECK(ctx);
    merge_kernel<<<D.F, 1024, 0, s>>>(D);
    UCO_LAUNCH_CHECK(ctx);
    A.bad = D.bad2;
    rc = uco_pnp_launch_dev(ctx, D.F, D.heads + D.F, A, D.outs + D.F);
    if (rc != UCO_OK) return rc;
    finalize_kernel<<<D.F, 256, 0, s>>>(D);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

void fill_params(TrackDev& D, const uco_track_params* p) {
    D.fx = p->fx; D.fy = p->fy; D.cx = p->cx; D.cy = p->cy; D.bf = p->bf;
    D.min_x = p->min_xy[0]; D.min_y = p->min_xy[1]; D.max_x = p->max_xy[0]; D.max_y = p->max_xy[1];
    D.n_levels = p->n_levels;
    for (int k = 0; k < UCO_MATCH_MAX_SCALES; k++) D.sf[k] = k < p->n_levels ? p->scale_factors[k] : 0.f;
    D.thr1 = (float)(p->max_desc_dist * 1.5);      // system.cpp:6559: maxDescDistance*1.5 (double product narrowed by the float parameter)
    D.thr2 = p->max_desc_dist * 2;                 // :6883: maxDescDistance*2 (float * int)
    D.proj_dist_thr = p->proj_dist_thr;
}

int node_cap_for(int kp_cap) { return 2 * (kp_cap / 5) + 2; }

// after the launches of a track_batch_dev(NO_SYNC) call: wait and turn the device error words into a status
// long_wait: the stream holds milliseconds of work (a batch of frames): the calling thread sleeps on an event instead of spinning, so
// that the tracker threads of several processes per host (one per GPU) do not burn the cores the mappers' planners need
int uco_track_check_errors(uco_b200_ctx* ctx, bool long_wait = false) {
    int32_t* herr = (int32_t*)uco_pinned(ctx, WS_TRACK_ERR, 16);
    if (!herr || !ctx->track_err_dev) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemcpyAsync(herr, ctx->track_err_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (long_wait) UCO_CUDA(ctx, uco_sleep_sync(ctx));
    else UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (herr[1]) return uco_fail(ctx, UCO_E_CAPACITY, "track_batch: kd-tree build failed (code %d)", herr[1]);
    if (herr[0]) return uco_fail(ctx, UCO_E_CAPACITY, "track_batch: a kd-tree is deeper than the %d deferred branches the device walk keeps", KD_STACK);
    return UCO_OK;
}

}  // namespace

extern "C" {

int uco_b200_track_batch_dev(uco_b200_ctx* ctx, const uco_track_batch* in, const uco_track_params* prm, const uco_track_out* out) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    int rc = check_params(ctx, prm);
    if (rc != UCO_OK) return rc;
    if (!in || !out || in->n_frames <= 0 || in->kp_cap <= 0 || in->prev_cap <= 0 || in->map_cap <= 0)
        return uco_fail(ctx, UCO_E_INVALID, "track_batch_dev: bad sizes");
    if (!in->kps || !in->desc || !in->n_kp || !in->prev_kps || !in->prev_desc || !in->prev_n_kp || !in->prev_mp_row || !in->map_n || !in->mp_id ||
        !in->mp_pos || !in->mp_normal || !in->mp_min_dist || !in->mp_max_dist || !in->mp_desc || !in->pose_prior || !out->matches ||
        !out->n_matches || !out->pose || !out->n_good || !out->status || !out->n_tbp)
        return uco_fail(ctx, UCO_E_INVALID, "track_batch_dev: null array");
    const int F = in->n_frames, node_cap = node_cap_for(in->kp_cap);
    const Scratch S = plan_scratch(F, in->kp_cap, in->prev_cap, in->map_cap);
    size_t toff = al(S.total);
    const size_t o_nodes = toff; toff += al(sizeof(uco_kdnode) * (size_t)F * node_cap);
    const size_t o_leaf = toff; toff += al(4 * (size_t)F * in->kp_cap);
    const size_t o_bbox = toff; toff += al(32 * (size_t)F);
    const size_t o_nn = toff; toff += al(4 * (size_t)F);
    uint8_t* scr = (uint8_t*)uco_ws(ctx, WS_TRACK, toff);
    int32_t* herr = (int32_t*)uco_pinned(ctx, WS_TRACK_ERR, 16);
    if (!scr || !herr) return UCO_E_NOMEM;
    TrackDev D;
    memset(&D, 0, sizeof D);
    D.F = F; D.kp_cap = in->kp_cap; D.prev_cap = in->prev_cap; D.map_cap = in->map_cap; D.node_cap = node_cap;
    D.kps = in->kps; D.desc = (const uint32_t*)in->desc; D.n_kp = in->n_kp; D.depth = in->depth;
    D.prev_kps = in->prev_kps; D.prev_desc = (const uint32_t*)in->prev_desc; D.prev_n = in->prev_n_kp; D.prev_row = in->prev_mp_row;
    D.map_n = in->map_n; D.mp_id = in->mp_id; D.mp_pos = in->mp_pos; D.mp_normal = in->mp_normal; D.mp_min = in->mp_min_dist; D.mp_max = in->mp_max_dist;
    D.mp_desc = (const uint32_t*)in->mp_desc; D.mp_stable = in->mp_stable; D.mp_local = in->mp_local; D.pose_prior = in->pose_prior;
    fill_params(D, prm);
    D.matches = out->matches; D.n_matches = out->n_matches; D.pose_out = out->pose; D.n_good = out->n_good; D.status = out->status;
    D.n_tbp = out->n_tbp; D.visible = out->visible;
    D.nodes = (const uco_kdnode*)(scr + o_nodes); D.leaf = (const int32_t*)(scr + o_leaf); D.bbox = (const double*)(scr + o_bbox);
    D.n_nodes = (const int32_t*)(scr + o_nn);
    cudaStream_t s = ctx->stream;
    rc = setup_scratch(ctx, D, scr, S);
    if (rc != UCO_OK) return rc;
    rc = uco_kdtree_build_launch(ctx, F, in->kps, (size_t)in->kp_cap, in->n_kp, 0, in->kp_cap, (uco_kdnode*)(scr + o_nodes), node_cap,
                                 (int32_t*)(scr + o_leaf), (double*)(scr + o_bbox), (int32_t*)(scr + o_nn), (int32_t*)(scr + S.o_err) + 1);
    if (rc != UCO_OK) return rc;
    rc = run_track(ctx, D, scr, S);
    if (rc != UCO_OK) return rc;
    ctx->track_err_dev = scr + S.o_err;
    if (!(in->flags & UCO_TRACK_NO_SYNC)) {
        UCO_CUDA(ctx, cudaMemcpyAsync(herr, scr + S.o_err, 8, cudaMemcpyDeviceToHost, s));
        UCO_CUDA(ctx, cudaStreamSynchronize(s));
        if (herr[1]) return uco_fail(ctx, UCO_E_CAPACITY, "track_batch_dev: kd-tree build failed (code %d)", herr[1]);
        if (herr[0]) return uco_fail(ctx, UCO_E_CAPACITY, "track_batch_dev: a kd-tree is deeper than the %d deferred branches the device walk keeps", KD_STACK);
    }
    return UCO_OK;
}

// host buffers in, host buffers out: one staged upload, the device sequence, one download
int uco_b200_track_batch(uco_b200_ctx* ctx, const uco_track_batch* in, const uco_track_params* prm, const uco_track_out* out) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    int rc = check_params(ctx, prm);
    if (rc != UCO_OK) return rc;
    if (!in || !out || in->n_frames <= 0 || in->kp_cap <= 0 || in->prev_cap <= 0 || in->map_cap <= 0)
        return uco_fail(ctx, UCO_E_INVALID, "track_batch: bad sizes");
    if (!in->kps || !in->desc || !in->n_kp || !in->prev_kps || !in->prev_desc || !in->prev_n_kp || !in->prev_mp_row || !in->map_n || !in->mp_id ||
        !in->mp_pos || !in->mp_normal || !in->mp_min_dist || !in->mp_max_dist || !in->mp_desc || !in->pose_prior || !out->matches ||
        !out->n_matches || !out->pose || !out->n_good || !out->status || !out->n_tbp)
        return uco_fail(ctx, UCO_E_INVALID, "track_batch: null array");
    const size_t F = in->n_frames, K = F * in->kp_cap, P = F * in->prev_cap, M = F * in->map_cap;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += al(b); return o; };
    const size_t i_kps = take(sizeof(uco_keypoint) * K), i_desc = take(32 * K), i_nkp = take(4 * F), i_depth = take(in->depth ? 4 * K : 0),
                 i_pkps = take(sizeof(uco_keypoint) * P), i_pdesc = take(32 * P), i_pn = take(4 * F), i_prow = take(4 * P), i_mn = take(4 * F),
                 i_id = take(4 * M), i_pos = take(12 * M), i_nrm = take(12 * M), i_min = take(4 * M), i_max = take(4 * M), i_mdesc = take(32 * M),
                 i_stab = take(in->mp_stable ? M : 0), i_loc = take(in->mp_local ? M : 0), i_pose = take(64 * F);
    const size_t in_bytes = off;
    const size_t o_match = take(sizeof(uco_match) * K), o_nm = take(4 * F), o_pose = take(64 * F), o_good = take(4 * F), o_stat = take(4 * F),
                 o_tbp = take(4 * F), o_vis = take(out->visible ? M : 0);
    const size_t out_bytes = off - o_match;
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_TRACK_IN, off);
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_TRACK_IN, in_bytes);
    uint8_t* ho = (uint8_t*)uco_pinned(ctx, WS_TRACK_OUT, out_bytes);
    if (!d || !h || !ho) return UCO_E_NOMEM;
    memcpy(h + i_kps, in->kps, sizeof(uco_keypoint) * K); memcpy(h + i_desc, in->desc, 32 * K); memcpy(h + i_nkp, in->n_kp, 4 * F);
    if (in->depth) memcpy(h + i_depth, in->depth, 4 * K);
    memcpy(h + i_pkps, in->prev_kps, sizeof(uco_keypoint) * P); memcpy(h + i_pdesc, in->prev_desc, 32 * P); memcpy(h + i_pn, in->prev_n_kp, 4 * F);
    memcpy(h + i_prow, in->prev_mp_row, 4 * P); memcpy(h + i_mn, in->map_n, 4 * F); memcpy(h + i_id, in->mp_id, 4 * M);
    memcpy(h + i_pos, in->mp_pos, 12 * M); memcpy(h + i_nrm, in->mp_normal, 12 * M); memcpy(h + i_min, in->mp_min_dist, 4 * M);
    memcpy(h + i_max, in->mp_max_dist, 4 * M); memcpy(h + i_mdesc, in->mp_desc, 32 * M);
    if (in->mp_stable) memcpy(h + i_stab, in->mp_stable, M);
    if (in->mp_local) memcpy(h + i_loc, in->mp_local, M);
    memcpy(h + i_pose, in->pose_prior, 64 * F);
    for (size_t f = 0; f < F; f++)
        if (in->n_kp[f] < 0 || in->n_kp[f] > in->kp_cap || in->prev_n_kp[f] < 0 || in->prev_n_kp[f] > in->prev_cap || in->map_n[f] < 0 || in->map_n[f] > in->map_cap)
            return uco_fail(ctx, UCO_E_INVALID, "track_batch: frame %zu has counts outside its capacities", f);
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, s));
    uco_track_batch din = *in;
    din.flags = 0;
    din.kps = (const uco_keypoint*)(d + i_kps); din.desc = d + i_desc; din.n_kp = (const int32_t*)(d + i_nkp);
    din.depth = in->depth ? (const float*)(d + i_depth) : nullptr;
    din.prev_kps = (const uco_keypoint*)(d + i_pkps); din.prev_desc = d + i_pdesc; din.prev_n_kp = (const int32_t*)(d + i_pn);
    din.prev_mp_row = (const int32_t*)(d + i_prow); din.map_n = (const int32_t*)(d + i_mn); din.mp_id = (const uint32_t*)(d + i_id);
    din.mp_pos = (const float*)(d + i_pos); din.mp_normal = (const float*)(d + i_nrm); din.mp_min_dist = (const float*)(d + i_min);
    din.mp_max_dist = (const float*)(d + i_max); din.mp_desc = d + i_mdesc; din.mp_stable = in->mp_stable ? d + i_stab : nullptr;
    din.mp_local = in->mp_local ? d + i_loc : nullptr; din.pose_prior = (const float*)(d + i_pose);
    uco_track_out dout;
    dout.matches = (uco_match*)(d + o_match); dout.n_matches = (int32_t*)(d + o_nm); dout.pose = (float*)(d + o_pose);
    dout.n_good = (int32_t*)(d + o_good); dout.status = (int32_t*)(d + o_stat); dout.n_tbp = (int32_t*)(d + o_tbp);
    dout.visible = out->visible ? d + o_vis : nullptr;
    din.flags = UCO_TRACK_NO_SYNC;
    rc = uco_b200_track_batch_dev(ctx, &din, prm, &dout);
    if (rc != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaMemcpyAsync(ho, d + o_match, out_bytes, cudaMemcpyDeviceToHost, s));
    rc = uco_track_check_errors(ctx);
    if (rc != UCO_OK) return rc;
    const uint8_t* b = ho - o_match;
    memcpy(out->n_matches, b + o_nm, 4 * F); memcpy(out->pose, b + o_pose, 64 * F); memcpy(out->n_good, b + o_good, 4 * F);
    memcpy(out->status, b + o_stat, 4 * F); memcpy(out->n_tbp, b + o_tbp, 4 * F);
    if (out->visible) memcpy(out->visible, b + o_vis, M);
    for (size_t f = 0; f < F; f++)
        memcpy(out->matches + f * in->kp_cap, b + o_match + sizeof(uco_match) * f * in->kp_cap, sizeof(uco_match) * (size_t)out->n_matches[f]);
    return UCO_OK;
}

// the search by projection alone, one frame, host buffers; the frame's kd-tree is the caller's (Frame::keypoint_kdtree)
int uco_b200_track_projected(uco_b200_ctx* ctx, int n_prev, const uco_keypoint* prev_kps, const uint8_t* prev_desc,
                             const int32_t* prev_mp_row, const uco_mappoints* mp, const uco_frame_view* fr, const float* pose_f2g,
                             float dist_thr, float proj_dist_thr, uco_match* out, int* n_out) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!mp || !fr || !pose_f2g || !n_out || n_prev < 0) return uco_fail(ctx, UCO_E_INVALID, "track_projected: bad arguments");
    *n_out = 0;
    const int m = mp->n, nk = fr->n_kp;
    if (m < 0 || nk < 0 || fr->n_nodes < 0 || fr->n_levels < 2 || fr->n_levels > UCO_MATCH_MAX_SCALES)
        return uco_fail(ctx, UCO_E_INVALID, "track_projected: bad sizes m=%d n_kp=%d n_nodes=%d n_levels=%d", m, nk, fr->n_nodes, fr->n_levels);
    if (n_prev == 0 || m == 0 || nk == 0 || fr->n_nodes == 0) return UCO_OK;
    if (!out || !prev_kps || !prev_desc || !prev_mp_row || !mp->ids || !mp->pos || !fr->kps || !fr->desc || !fr->nodes || !fr->leaf_idx || !fr->scale_factors)
        return uco_fail(ctx, UCO_E_INVALID, "track_projected: null array");
    int n_leaf = 0;
    for (int i = 0; i < fr->n_nodes; i++) {
        const uco_kdnode& N = fr->nodes[i];
        if (N.col < 0) n_leaf = n_leaf > N.leaf_begin + N.leaf_count ? n_leaf : N.leaf_begin + N.leaf_count;
        else if ((unsigned)N.left >= (unsigned)fr->n_nodes || (unsigned)N.right >= (unsigned)fr->n_nodes || N.col > 1 || N.left <= i || N.right <= i)
            return uco_fail(ctx, UCO_E_INVALID, "track_projected: kd-tree node %d is malformed", i);
    }
    if (n_leaf > nk) return uco_fail(ctx, UCO_E_INVALID, "track_projected: kd-tree lists %d keypoints, the frame has %d", n_leaf, nk);
    for (int i = 0; i < n_leaf; i++)
        if ((unsigned)fr->leaf_idx[i] >= (unsigned)nk) return uco_fail(ctx, UCO_E_INVALID, "track_projected: kd-tree indexes keypoint %d of %d", fr->leaf_idx[i], nk);
    for (int i = 0; i < n_prev; i++)
        if (prev_mp_row[i] >= m) return uco_fail(ctx, UCO_E_INVALID, "track_projected: previous keypoint %d refers to map row %d of %d", i, prev_mp_row[i], m);
    const Scratch S = plan_scratch(1, nk, n_prev, m);
    size_t off = al(S.total);
    auto take = [&](size_t b) { size_t o = off; off += al(b); return o; };
    const size_t base = off;
    const size_t i_kps = take(sizeof(uco_keypoint) * (size_t)nk), i_desc = take(32 * (size_t)nk), i_nkp = take(16), i_pkps = take(sizeof(uco_keypoint) * (size_t)n_prev),
                 i_pdesc = take(32 * (size_t)n_prev), i_prow = take(4 * (size_t)n_prev), i_id = take(4 * (size_t)m), i_pos = take(12 * (size_t)m),
                 i_pose = take(64), i_nodes = take(sizeof(uco_kdnode) * (size_t)fr->n_nodes), i_leaf = take(4 * (size_t)nk), i_bbox = take(32);
    const size_t in_end = off;
    const size_t o_match = take(sizeof(uco_match) * (size_t)nk), o_cnt = take(64);
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_TRACK, off);
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_TRACK_IN, in_end - base);
    uint8_t* ho = (uint8_t*)uco_pinned(ctx, WS_TRACK_OUT, off - o_match);
    if (!d || !h || !ho) return UCO_E_NOMEM;
    uint8_t* hb = h - base;
    memcpy(hb + i_kps, fr->kps, sizeof(uco_keypoint) * (size_t)nk);
    {
        const size_t ds = fr->desc_stride ? fr->desc_stride : 32;
        for (int i = 0; i < nk; i++) memcpy(hb + i_desc + 32 * (size_t)i, fr->desc + ds * (size_t)i, 32);
    }
    int32_t* cnt = (int32_t*)(hb + i_nkp);
    cnt[0] = nk; cnt[1] = n_prev; cnt[2] = m; cnt[3] = fr->n_nodes;
    memcpy(hb + i_pkps, prev_kps, sizeof(uco_keypoint) * (size_t)n_prev); memcpy(hb + i_pdesc, prev_desc, 32 * (size_t)n_prev);
    memcpy(hb + i_prow, prev_mp_row, 4 * (size_t)n_prev); memcpy(hb + i_id, mp->ids, 4 * (size_t)m); memcpy(hb + i_pos, mp->pos, 12 * (size_t)m);
    memcpy(hb + i_pose, pose_f2g, 64); memcpy(hb + i_nodes, fr->nodes, sizeof(uco_kdnode) * (size_t)fr->n_nodes);
    memcpy(hb + i_leaf, fr->leaf_idx, 4 * (size_t)n_leaf); memcpy(hb + i_bbox, fr->bbox, 32);
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d + base, h, in_end - base, cudaMemcpyHostToDevice, s));
    TrackDev D;
    memset(&D, 0, sizeof D);
    D.F = 1; D.kp_cap = nk; D.prev_cap = n_prev; D.map_cap = m; D.node_cap = fr->n_nodes; D.stage1_only = 1;
    D.kps = (const uco_keypoint*)(d + i_kps); D.desc = (const uint32_t*)(d + i_desc); D.n_kp = (const int32_t*)(d + i_nkp);
    D.prev_kps = (const uco_keypoint*)(d + i_pkps); D.prev_desc = (const uint32_t*)(d + i_pdesc); D.prev_n = (const int32_t*)(d + i_nkp) + 1;
    D.prev_row = (const int32_t*)(d + i_prow); D.map_n = (const int32_t*)(d + i_nkp) + 2; D.mp_id = (const uint32_t*)(d + i_id);
    D.mp_pos = (const float*)(d + i_pos); D.pose_prior = (const float*)(d + i_pose);
    D.nodes = (const uco_kdnode*)(d + i_nodes); D.leaf = (const int32_t*)(d + i_leaf); D.bbox = (const double*)(d + i_bbox);
    D.n_nodes = (const int32_t*)(d + i_nkp) + 3;
    D.fx = fr->fx; D.fy = fr->fy; D.cx = fr->cx; D.cy = fr->cy;
    D.min_x = fr->min_xy[0]; D.min_y = fr->min_xy[1]; D.max_x = fr->max_xy[0]; D.max_y = fr->max_xy[1];
    D.n_levels = fr->n_levels;
    for (int k = 0; k < UCO_MATCH_MAX_SCALES; k++) D.sf[k] = k < fr->n_levels ? fr->scale_factors[k] : 0.f;
    D.thr1 = dist_thr; D.proj_dist_thr = proj_dist_thr;
    D.matches = (uco_match*)(d + o_match); D.n_matches = (int*)(d + o_cnt); D.n_tbp = (int*)(d + o_cnt) + 1;
    int rc = setup_scratch(ctx, D, d, S);
    if (rc != UCO_OK) return rc;
    rc = run_track(ctx, D, d, S);
    if (rc != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaMemcpyAsync(ho, d + o_match, off - o_match, cudaMemcpyDeviceToHost, s));
    int32_t* herr = (int32_t*)uco_pinned(ctx, WS_TRACK_ERR, 16);
    if (!herr) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemcpyAsync(herr, d + S.o_err, 8, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    if (herr[0]) return uco_fail(ctx, UCO_E_CAPACITY, "track_projected: the kd-tree is deeper than the %d deferred branches the device walk keeps", KD_STACK);
    const int n = *(int*)(ho + (o_cnt - o_match));
    memcpy(out, ho, sizeof(uco_match) * (size_t)n);
    *n_out = n;
    return UCO_OK;
}

// ---- device-resident mirror of the tracking state of n independent streams (Frame / Map mirrors) ------------------
// What the tracker reads besides the new image: the previous frame's keypoints / descriptors / map-point assignment
// (Frame::und_kpts, ::desc, ::ids, src/map_types/frame.h:60-75) and the local map's points (MapPoint coordinates, normal, scale
// range, descriptor, stable flag; src/map_types/mappoint.h).  In the reference these live in host containers; here they live in
// HBM between calls, so a frame costs one image upload and one result download.
struct uco_b200_track_state {
    int n_streams, prev_cap, map_cap;
    uint8_t* d;      // one allocation
    size_t o_pkps, o_pdesc, o_pn, o_prow, o_mn, o_id, o_pos, o_nrm, o_min, o_max, o_mdesc, o_stab, o_loc, o_prior, total;
};

int uco_b200_track_state_create(uco_b200_ctx* ctx, int n_streams, int prev_cap, int map_cap, uco_b200_track_state** out) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!out || n_streams <= 0 || prev_cap <= 0 || map_cap <= 0) return uco_fail(ctx, UCO_E_INVALID, "track_state_create: bad arguments");
    uco_b200_track_state* st = new (std::nothrow) uco_b200_track_state();
    if (!st) return UCO_E_NOMEM;
    st->n_streams = n_streams; st->prev_cap = prev_cap; st->map_cap = map_cap;
    const size_t F = n_streams, P = F * prev_cap, M = F * map_cap;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += al(b); return o; };
    st->o_pkps = take(sizeof(uco_keypoint) * P); st->o_pdesc = take(32 * P); st->o_pn = take(4 * F); st->o_prow = take(4 * P);
    st->o_mn = take(4 * F); st->o_id = take(4 * M); st->o_pos = take(12 * M); st->o_nrm = take(12 * M); st->o_min = take(4 * M);
    st->o_max = take(4 * M); st->o_mdesc = take(32 * M); st->o_stab = take(M); st->o_loc = take(M); st->o_prior = take(64 * F);
    st->total = off;
    cudaError_t e = cudaMalloc(&st->d, off);
    if (e != cudaSuccess) { delete st; return uco_fail(ctx, UCO_E_NOMEM, "track_state_create: cudaMalloc(%zu) -> %s", off, cudaGetErrorString(e)); }
    cudaMemsetAsync(st->d, 0, off, ctx->stream);
    cudaMemsetAsync(st->d + st->o_stab, 1, M, ctx->stream);
    cudaMemsetAsync(st->d + st->o_loc, 1, M, ctx->stream);
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = st;
    return UCO_OK;
}

void uco_b200_track_state_free(uco_b200_ctx* ctx, uco_b200_track_state* st) {
    if (!st) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(st->d);
    delete st;
}

// previous frame of one stream: n keypoints (octave is read), descriptors, and per keypoint the ROW of its map point in the stream's
// map block (-1: none / bad point)
int uco_b200_track_state_set_prev(uco_b200_ctx* ctx, uco_b200_track_state* st, int stream, int n, const uco_keypoint* kps, const uint8_t* desc,
                                  const int32_t* mp_row) {
    if (!ctx || !st) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (stream < 0 || stream >= st->n_streams || n < 0 || n > st->prev_cap || (n && (!kps || !desc || !mp_row)))
        return uco_fail(ctx, UCO_E_INVALID, "track_state_set_prev: bad arguments");
    cudaStream_t s = ctx->stream;
    const size_t b = (size_t)stream * st->prev_cap;
    int32_t* hn = (int32_t*)uco_pinned(ctx, WS_TRACK_ERR, 64);
    if (!hn) return UCO_E_NOMEM;
    hn[2] = n;
    if (n) {
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_pkps + sizeof(uco_keypoint) * b, kps, sizeof(uco_keypoint) * (size_t)n, cudaMemcpyHostToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_pdesc + 32 * b, desc, 32 * (size_t)n, cudaMemcpyHostToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_prow + 4 * b, mp_row, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    }
    UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_pn + 4 * (size_t)stream, hn + 2, 4, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    return UCO_OK;
}

// the map block of one stream (the points the tracker may see: those of the previous frame + the local map of the reference keyframe)
int uco_b200_track_state_set_map(uco_b200_ctx* ctx, uco_b200_track_state* st, int stream, const uco_mappoints* mp, const uint8_t* stable,
                                 const uint8_t* local) {
    if (!ctx || !st) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (stream < 0 || stream >= st->n_streams || !mp || mp->n < 0 || mp->n > st->map_cap ||
        (mp->n && (!mp->ids || !mp->pos || !mp->normal || !mp->min_dist || !mp->max_dist || !mp->desc)))
        return uco_fail(ctx, UCO_E_INVALID, "track_state_set_map: bad arguments");
    cudaStream_t s = ctx->stream;
    const size_t b = (size_t)stream * st->map_cap, n = mp->n;
    int32_t* hn = (int32_t*)uco_pinned(ctx, WS_TRACK_ERR, 64);
    if (!hn) return UCO_E_NOMEM;
    hn[3] = mp->n;
    if (n) {
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_id + 4 * b, mp->ids, 4 * n, cudaMemcpyHostToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_pos + 12 * b, mp->pos, 12 * n, cudaMemcpyHostToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_nrm + 12 * b, mp->normal, 12 * n, cudaMemcpyHostToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_min + 4 * b, mp->min_dist, 4 * n, cudaMemcpyHostToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_max + 4 * b, mp->max_dist, 4 * n, cudaMemcpyHostToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_mdesc + 32 * b, mp->desc, 32 * n, cudaMemcpyHostToDevice, s));
        if (stable) UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_stab + b, stable, n, cudaMemcpyHostToDevice, s));
        else UCO_CUDA(ctx, cudaMemsetAsync(st->d + st->o_stab + b, 1, n, s));
        if (local) UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_loc + b, local, n, cudaMemcpyHostToDevice, s));
        else UCO_CUDA(ctx, cudaMemsetAsync(st->d + st->o_loc + b, 1, n, s));
    }
    UCO_CUDA(ctx, cudaMemcpyAsync(st->d + st->o_mn + 4 * (size_t)stream, hn + 3, 4, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    return UCO_OK;
}

namespace {
void state_view(const uco_b200_track_state* st, uco_track_batch* b) {
    b->n_frames = st->n_streams; b->prev_cap = st->prev_cap; b->map_cap = st->map_cap;
    b->prev_kps = (const uco_keypoint*)(st->d + st->o_pkps); b->prev_desc = st->d + st->o_pdesc; b->prev_n_kp = (const int32_t*)(st->d + st->o_pn);
    b->prev_mp_row = (const int32_t*)(st->d + st->o_prow); b->map_n = (const int32_t*)(st->d + st->o_mn); b->mp_id = (const uint32_t*)(st->d + st->o_id);
    b->mp_pos = (const float*)(st->d + st->o_pos); b->mp_normal = (const float*)(st->d + st->o_nrm); b->mp_min_dist = (const float*)(st->d + st->o_min);
    b->mp_max_dist = (const float*)(st->d + st->o_max); b->mp_desc = st->d + st->o_mdesc; b->mp_stable = st->d + st->o_stab; b->mp_local = st->d + st->o_loc;
    b->pose_prior = (const float*)(st->d + st->o_prior);
}
}  // namespace

// device-resident variant: the current frames' keypoints / descriptors (outputs of uco_b200_orb_extract_batch_dev) against the mirrored
// state; pose_prior_dev: n_streams x 16 (device).  Outputs device resident.
int uco_b200_track_state_step_dev(uco_b200_ctx* ctx, const uco_b200_track_state* st, const uco_keypoint* kps_dev, const uint8_t* desc_dev,
                                  const int32_t* n_kp_dev, int kp_cap, const float* pose_prior_dev, const uco_track_params* prm,
                                  const uco_track_out* out_dev, int flags) {
    if (!ctx || !st) return UCO_E_INVALID;
    uco_track_batch b;
    memset(&b, 0, sizeof b);
    state_view(st, &b);
    b.kp_cap = kp_cap; b.kps = kps_dev; b.desc = desc_dev; b.n_kp = n_kp_dev; b.flags = flags;
    if (pose_prior_dev) b.pose_prior = pose_prior_dev;
    return uco_b200_track_batch_dev(ctx, &b, prm, out_dev);
}

int uco_orb_extract_keep_dev(uco_b200_ctx* ctx, const uint8_t* const* imgs, int n_imgs, int w, int h, size_t stride, const uco_orb_params* prm,
                             uco_keypoint** d_kps, uint8_t** d_desc, int** d_nout, int** d_err);

// One tracking step of every stream through HOST buffers: the new frames in (n_streams images) + pose priors, on the device ORB
// extraction -> kd-trees -> the tracker's sequence against the mirrored state, out: the frames' keypoints / descriptors (what
// Frame::kpts / ::desc hold), poses, match lists with inlier flags.  One upload per image, one staged download.
int uco_b200_track_frames(uco_b200_ctx* ctx, const uco_b200_track_state* st, const uint8_t* const* imgs, int w, int h, size_t stride,
                          const uco_orb_params* orb, const uco_track_params* prm, const float* pose_prior, uco_keypoint* kps, uint8_t* desc,
                          int32_t* n_kp, const uco_track_out* out) {
    UCO_RANGE();
    if (!ctx || !st) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!imgs || !orb || !prm || !pose_prior || !out || !out->matches || !out->n_matches || !out->pose || !out->n_good || !out->status || !out->n_tbp)
        return uco_fail(ctx, UCO_E_INVALID, "track_frames: null argument");
    const int F = st->n_streams, mf = orb->max_features;
    uco_keypoint* d_kps; uint8_t* d_desc; int* d_nout; int* d_oerr;
    int rc = uco_orb_extract_keep_dev(ctx, imgs, F, w, h, stride, orb, &d_kps, &d_desc, &d_nout, &d_oerr);
    if (rc != UCO_OK) return rc;
    const size_t K = (size_t)F * mf;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += al(b); return o; };
    const size_t o_match = take(sizeof(uco_match) * K), o_nm = take(4 * (size_t)F), o_pose = take(64 * (size_t)F), o_good = take(4 * (size_t)F),
                 o_stat = take(4 * (size_t)F), o_tbp = take(4 * (size_t)F), o_nkp = take(4 * (size_t)F), o_oerr = take(16), o_prior = take(64 * (size_t)F);
    // page-locked caller buffers (cudaHostAlloc / cudaHostRegister'ed, e.g. a pinned cv::Mat allocator) receive the keypoints and
    // descriptors directly; pageable ones go through the context's pinned staging buffer + a host copy
    auto pinned = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool kps_direct = kps && pinned(kps), desc_direct = desc && pinned(desc);
    const size_t o_kps = take(kps && !kps_direct ? sizeof(uco_keypoint) * K : 0), o_desc = take(desc && !desc_direct ? 32 * K : 0);
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_TRACK_IN, off);
    uint8_t* ho = (uint8_t*)uco_pinned(ctx, WS_TRACK_OUT, off);
    if (!d || !ho) return UCO_E_NOMEM;
    cudaStream_t s = ctx->stream;
    memcpy(ho + o_prior, pose_prior, 64 * (size_t)F);
    UCO_CUDA(ctx, cudaMemcpyAsync(d + o_prior, ho + o_prior, 64 * (size_t)F, cudaMemcpyHostToDevice, s));
    uco_track_out dout;
    dout.matches = (uco_match*)(d + o_match); dout.n_matches = (int32_t*)(d + o_nm); dout.pose = (float*)(d + o_pose);
    dout.n_good = (int32_t*)(d + o_good); dout.status = (int32_t*)(d + o_stat); dout.n_tbp = (int32_t*)(d + o_tbp); dout.visible = nullptr;
    rc = uco_b200_track_state_step_dev(ctx, st, d_kps, d_desc, d_nout, mf, (const float*)(d + o_prior), prm, &dout, UCO_TRACK_NO_SYNC);
    if (rc != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaMemcpyAsync(ho, d, o_nkp, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_nkp, d_nout, 4 * (size_t)F, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_oerr, d_oerr, 4, cudaMemcpyDeviceToHost, s));
    if (kps) UCO_CUDA(ctx, cudaMemcpyAsync(kps_direct ? (void*)kps : (void*)(ho + o_kps), d_kps, sizeof(uco_keypoint) * K, cudaMemcpyDeviceToHost, s));
    if (desc) UCO_CUDA(ctx, cudaMemcpyAsync(desc_direct ? (void*)desc : (void*)(ho + o_desc), d_desc, 32 * K, cudaMemcpyDeviceToHost, s));
    rc = uco_track_check_errors(ctx, F >= 16);
    if (rc != UCO_OK) return rc;
    if (*(int*)(ho + o_oerr)) return uco_fail(ctx, UCO_E_CAPACITY, "track_frames: the extractor's internal selection list overflowed");
    memcpy(out->n_matches, ho + o_nm, 4 * (size_t)F); memcpy(out->pose, ho + o_pose, 64 * (size_t)F); memcpy(out->n_good, ho + o_good, 4 * (size_t)F);
    memcpy(out->status, ho + o_stat, 4 * (size_t)F); memcpy(out->n_tbp, ho + o_tbp, 4 * (size_t)F);
    if (n_kp) memcpy(n_kp, ho + o_nkp, 4 * (size_t)F);
    if (kps && !kps_direct) memcpy(kps, ho + o_kps, sizeof(uco_keypoint) * K);
    if (desc && !desc_direct) memcpy(desc, ho + o_desc, 32 * K);
    for (int f = 0; f < F; f++)
        memcpy(out->matches + (size_t)f * mf, ho + o_match + sizeof(uco_match) * (size_t)f * mf, sizeof(uco_match) * (size_t)out->n_matches[f]);
    return UCO_OK;
}

}  // extern "C"
