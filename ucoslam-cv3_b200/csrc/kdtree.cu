// kdtree.cu — K16: the frame's 2-d kd-tree (Frame::keypoint_kdtree) built ON THE DEVICE, node for node what the reference builds on
// the host (SURVEY.md 8f rank 4 "kd-tree replacement"; consumer: the projection matchers of project.cu / track.cu).
//
// Replaces (reference, relative to /root/reference):
//   src/map_types/frame.h:124-127          Frame::create_kdtree  (keypoint_kdtree.build(und_kpts), called by the frame extractor)
//   src/basictypes/picoflann.h:150-165     KdTreeIndex::build: index array 0..n-1, root bounding box, divideTree
//   src/basictypes/picoflann.h:240-345     divideTree: leaf at <= 10 points; split dimension = larger variance of a strided sample
//                                          (>= 100 points), cut = its mean; two Hoare passes (planeSplit :410-432) give lim1 / lim2;
//                                          the split index rule; std::sort + median cut when a side would hold < 10 points;
//                                          divlow = the cut, divhigh = the right child's tight lower bound; bounding boxes bottom-up
// Why identical and not "some spatial index": the best / second-best bookkeeping of Map::matchFrameToMapPoints (src/map.cpp:722-737)
// and of the tracker's search by projection (src/utils/system.cpp:5921-6456) depends on the ORDER in which the radius search reports
// keypoints, i.e. on this tree's shape and leaf contents.
//
// One CTA per frame builds its tree level by level in shared memory:
//   * per node of the level (one thread each): the sampled mean / variance in the reference's own sequential order (doubles of
//     float terms), split dimension and cut;
//   * the two Hoare passes for ALL nodes of the level at once: Hoare's scheme exchanges the k-th misplaced element from the left
//     with the k-th misplaced element from the right, so the resulting permutation follows from one block-wide prefix sum of the
//     predicate (rank of every misplaced element inside its node) and a parallel exchange — identical to the serial loop;
//   * per node: the split rule, and for degenerate cuts the serial replay of libstdc++'s std::sort (sort_exact.h) by one thread;
//   * bounding boxes / divhigh bottom-up over the levels, then the breadth-first working numbering is renamed to picoflann's
//     depth-first allocation order (children of the k-th divided node are 1+2k, 2+2k), so the node array equals the host's.
// Shared memory: 36 bytes per keypoint (n <= UCO_KDTREE_DEV_MAX_POINTS); everything else is registers.  No global traffic
// except reading the keypoints once and writing the nodes / leaf index list once.
#include "common.cuh"
#include "sort_exact.h"

namespace {

constexpr int KDB_THREADS = 1024;
constexpr int KDB_MAX_LEVELS = 1024;

struct KdBuildArgs {
    const uco_keypoint* kps;
    size_t kps_stride;       // records between frames
    const int32_t* n_kp;     // per frame (device); nullptr: n_fixed
    int n_fixed;
    int cap;                 // max points per frame
    int node_cap;            // nodes per frame in the output
    uco_kdnode* nodes;
    int32_t* leaf_idx;       // cap per frame
    double* bbox;            // 4 per frame
    int32_t* n_nodes;        // per frame
    int32_t* err;            // sticky error flag (device)
};

struct NodeW {               // working node (breadth-first numbering), 32 bytes
    double div;
    float cutf;
    float divhigh;
    uint16_t s, e;           // n <= UCO_KDTREE_DEV_MAX_POINTS < 65536
    uint16_t lim1, lim2;
    int16_t left;            // right = left + 1; -1 for a leaf
    uint16_t aux;            // internal-node count of the subtree
    uint16_t fin;            // final (depth-first) index
    int8_t col;
    int8_t pad;
};
static_assert(sizeof(NodeW) == 32, "NodeW layout");

// exclusive prefix sum of flag(p), p in [0, n), into pre[0..n] (pre[n] = total); all threads of the CTA call it
template <class F>
__device__ void block_scan(int n, int* pre, int* warp_tot, F flag) {
    const int items = (n + KDB_THREADS - 1) / KDB_THREADS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = threadIdx.x * items;
    int sum = 0;
    for (int k = 0; k < items; k++) {
        const int p = b + k;
        if (p < n) sum += flag(p) ? 1 : 0;
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        warp_tot[lane] = w;  // inclusive
    }
    __syncthreads();
    int run = inc - sum + (warp ? warp_tot[warp - 1] : 0);
    for (int k = 0; k < items; k++) {
        const int p = b + k;
        if (p < n) {
            pre[p] = run;
            run += flag(p) ? 1 : 0;
        }
    }
    if (threadIdx.x == KDB_THREADS - 1) pre[n] = warp_tot[31];
    __syncthreads();
}

__global__ void __launch_bounds__(KDB_THREADS, 1) kdtree_build_kernel(const KdBuildArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int f = blockIdx.x;
    const int n = A.n_kp ? A.n_kp[f] : A.n_fixed;
    uco_kdnode* out_nodes = A.nodes + (size_t)f * A.node_cap;
    int32_t* out_leaf = A.leaf_idx + (size_t)f * A.cap;
    double* out_bbox = A.bbox + 4 * (size_t)f;
    if (n <= 0 || n > A.cap) {
        if (threadIdx.x == 0) {
            A.n_nodes[f] = 0;
            out_bbox[0] = out_bbox[1] = out_bbox[2] = out_bbox[3] = 0;
            if (n > A.cap) atomicExch(A.err, 1);
        }
        return;
    }
    const int max_nodes = 2 * (n / 5) + 2;   // a node of 11..19 points is cut in the middle: leaves hold >= 5 points
    // shared layout
    NodeW* nd = (NodeW*)smem;                                   // max_nodes
    double* nbb = (double*)(nd + max_nodes);                    // 4 per node
    float2* xy = (float2*)(nbb + 4 * (size_t)max_nodes);        // n
    uint32_t* all = (uint32_t*)(xy + n);                        // n
    int* pre = (int*)(all + n);                                 // n + 1
    int* misL = pre + n + 1;                                    // n
    int* misR = misL + n;                                       // n
    int* seg = misR + n;                                        // n
    __shared__ int warp_tot[32];
    __shared__ int lvl_start[KDB_MAX_LEVELS + 1];
    __shared__ int n_nodes_s, n_internal_lvl, fail;

    const uco_keypoint* kps = A.kps + (size_t)f * A.kps_stride;
    for (int i = threadIdx.x; i < n; i += KDB_THREADS) {
        xy[i] = make_float2(kps[i].x, kps[i].y);
        all[i] = i;
        seg[i] = 0;
    }
    if (threadIdx.x == 0) {
        nd[0].s = 0; nd[0].e = n; nd[0].left = -1; nd[0].col = -1;
        n_nodes_s = 1;
        lvl_start[0] = 0;
        fail = 0;
    }
    __syncthreads();
    auto at = [&](uint32_t i, int d) -> float { return d ? xy[i].y : xy[i].x; };

    int level = 0, lb = 0, le = 1;  // nodes [lb, le) are the current level
    for (;;) {
        if (threadIdx.x == 0) n_internal_lvl = 0;
        __syncthreads();
        // A. per node: leaf or sampled mean / variance -> split dimension and cut (picoflann.h:372-400).  One node per WARP: the lanes fetch 32
        // samples side by side (two dependent shared-memory reads each: the serial form spent ~100 cycles per sample on them, one thread per
        // node while the rest of the CTA waited), then every lane accumulates them in the reference's own order (sample after sample, doubles of
        // float terms, the squares formed in float) - the same additions in the same order, without the load latency on the chain
        for (int k = lb + (threadIdx.x >> 5); k < le; k += KDB_THREADS / 32) {
            NodeW& N = nd[k];
            const int lane = threadIdx.x & 31;
            const int count = N.e - N.s, first = N.s;
            if (count <= 10) {
                __syncwarp();
                if (lane == 0) { N.col = -1; N.left = -1; }
                continue;
            }
            double mean[2] = {0, 0}, sq[2] = {0, 0};
            const int inc = count >= 200 ? count / 100 : 1;
            const int cnt = (count + inc - 1) / inc;
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                float2 v = make_float2(0.f, 0.f);
                if (c0 + lane < cnt) v = xy[all[first + (c0 + lane) * inc]];
                const int m = min(32, cnt - c0);
                for (int t = 0; t < m; t++) {
                    const float x = __shfl_sync(0xffffffffu, v.x, t), y = __shfl_sync(0xffffffffu, v.y, t);
                    mean[0] += x; sq[0] += x * x;   // the product is formed in float
                    mean[1] += y; sq[1] += y * y;
                }
            }
            const double ic = 1. / double(cnt);
            double var[2];
#pragma unroll
            for (int d = 0; d < 2; d++) {
                mean[d] *= ic;
                var[d] = sq[d] * ic - mean[d] * mean[d];
            }
            const int col = var[1] > var[0] ? 1 : 0;
            __syncwarp();
            if (lane == 0) {
                N.col = col;
                N.div = mean[col];
                N.cutf = (float)mean[col];
                atomicAdd(&n_internal_lvl, 1);
            }
        }
        __syncthreads();
        if (n_internal_lvl == 0) break;
        // B. first Hoare pass of every internal node of the level: (< cut | >= cut) over [s, e)
        {
            auto flag = [&](int p) -> bool {
                const NodeW& N = nd[seg[p]];
                return N.col >= 0 && at(all[p], N.col) < N.cutf;
            };
            block_scan(n, pre, warp_tot, flag);
            for (int p = threadIdx.x; p < n; p += KDB_THREADS) {
                const NodeW& N = nd[seg[p]];
                if (N.col < 0) continue;
                const int L = pre[N.e] - pre[N.s], before = pre[p] - pre[N.s];
                const bool lt = pre[p + 1] != pre[p];
                if (p < N.s + L) { if (!lt) misL[N.s + (p - N.s) - before] = p; }
                else if (lt) misR[N.s + (L - before - 1)] = p;
                if (p == N.s) nd[seg[p]].lim1 = L;
            }
            __syncthreads();
            for (int p = threadIdx.x; p < n; p += KDB_THREADS) {
                const NodeW& N = nd[seg[p]];
                if (N.col < 0) continue;
                const int L = N.lim1, M = L - (pre[N.s + L] - pre[N.s]);  // misplaced pairs
                if (p - N.s < M) {
                    const int a = misL[p], b = misR[p];
                    const uint32_t t = all[a]; all[a] = all[b]; all[b] = t;
                }
            }
            __syncthreads();
        }
        // C. second pass over [s + lim1, e): (<= cut | > cut)
        {
            auto flag = [&](int p) -> bool {
                const NodeW& N = nd[seg[p]];
                return N.col >= 0 && p >= N.s + N.lim1 && at(all[p], N.col) <= N.cutf;
            };
            block_scan(n, pre, warp_tot, flag);
            for (int p = threadIdx.x; p < n; p += KDB_THREADS) {
                const NodeW& N = nd[seg[p]];
                if (N.col < 0) continue;
                const int s2 = N.s + N.lim1;
                if (p == N.s) nd[seg[p]].lim2 = N.lim1 + (pre[N.e] - pre[s2]);
                if (p < s2) continue;
                const int L = pre[N.e] - pre[s2], before = pre[p] - pre[s2];
                const bool le_ = pre[p + 1] != pre[p];
                if (p < s2 + L) { if (!le_) misL[s2 + (p - s2) - before] = p; }
                else if (le_) misR[s2 + (L - before - 1)] = p;
            }
            __syncthreads();
            for (int p = threadIdx.x; p < n; p += KDB_THREADS) {
                const NodeW& N = nd[seg[p]];
                if (N.col < 0) continue;
                const int s2 = N.s + N.lim1;
                if (p < s2) continue;
                const int L = N.lim2 - N.lim1, M = L - (pre[s2 + L] - pre[s2]);
                if (p - s2 < M) {
                    const int a = misL[p], b = misR[p];
                    const uint32_t t = all[a]; all[a] = all[b]; all[b] = t;
                }
            }
            __syncthreads();
        }
        // D. split rule (+ the std::sort fallback), children.  One node per WARP, lane 0 works: the serial std::sort replays of neighbouring
        // nodes (every node of 11 .. 19 points takes one) used to sit in the lanes of ONE warp and ran one after the other (41 % of the kernel
        // was the barrier below)
        for (int k = lb + (threadIdx.x >> 5); k < le; k += KDB_THREADS / 32) {
            if (threadIdx.x & 31) continue;
            NodeW& N = nd[k];
            if (N.col < 0) continue;
            const int count = N.e - N.s, half = count / 2, lim1 = N.lim1, lim2 = N.lim2;
            int split = lim1 > half ? lim1 : (lim2 < half ? lim2 : half);
            if (lim1 == count || lim2 == 0) split = half;
            if (split < 10 || count - split < 10) {
                const int col = N.col;
                const float2* P = xy;
                auto less = [P, col](uint32_t a, uint32_t b) -> bool { return col ? P[a].y < P[b].y : P[a].x < P[b].x; };
                uco_sort::sort_(all + N.s, all + N.e, less);
                split = half;
                N.div = (double)at(all[N.s + split], col);
            }
            const int L = atomicAdd(&n_nodes_s, 2);
            if (L + 2 > max_nodes) { fail = 1; N.left = -1; N.col = -1; continue; }
            N.left = L;
            nd[L].s = N.s; nd[L].e = N.s + split; nd[L].left = -1; nd[L].col = -1;
            nd[L + 1].s = N.s + split; nd[L + 1].e = N.e; nd[L + 1].left = -1; nd[L + 1].col = -1;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < n; p += KDB_THREADS) {
            const NodeW& N = nd[seg[p]];
            if (N.col >= 0 && N.left >= 0) seg[p] = p < nd[N.left].e ? N.left : N.left + 1;
        }
        lb = le;
        le = n_nodes_s;
        level++;
        if (threadIdx.x == 0) lvl_start[level] = lb;
        __syncthreads();
        if (level >= KDB_MAX_LEVELS - 1 || fail) { fail = 1; break; }
        if (lb == le) break;
    }
    __syncthreads();
    if (fail) {
        if (threadIdx.x == 0) { atomicExch(A.err, 2); A.n_nodes[f] = 0; }
        return;
    }
    const int total = n_nodes_s;
    if (threadIdx.x == 0) lvl_start[level + 1] = total;
    __syncthreads();
    // bounding boxes, divhigh and subtree sizes bottom-up (picoflann.h:322-345); levels 0..level hold nodes
    for (int lv = level; lv >= 0; lv--) {
        const int b0 = lvl_start[lv], b1 = lvl_start[lv + 1];
        for (int k = b0 + threadIdx.x; k < b1; k += KDB_THREADS) {
            NodeW& N = nd[k];
            double* bb = nbb + 4 * (size_t)k;
            if (N.col < 0) {  // leaf: tight box of its points (computeBoundingBox :356-370)
                float2 v = xy[all[N.s]];
                float x0 = v.x, x1 = v.x, y0 = v.y, y1 = v.y;
                for (int i = N.s + 1; i < N.e; i++) {
                    v = xy[all[i]];
                    if (v.x < x0) x0 = v.x;
                    if (v.x > x1) x1 = v.x;
                    if (v.y < y0) y0 = v.y;
                    if (v.y > y1) y1 = v.y;
                }
                bb[0] = x0; bb[1] = x1; bb[2] = y0; bb[3] = y1;
                N.aux = 0;
            } else {
                const double* l = nbb + 4 * (size_t)N.left;
                const double* r = nbb + 4 * (size_t)(N.left + 1);
                double lbx[4] = {l[0], l[1], l[2], l[3]};
                lbx[2 * N.col + 1] = N.div;   // :337: the split value is put back over the bound the recursion tightened
                N.divhigh = (float)r[2 * N.col];
                bb[0] = fmin(lbx[0], r[0]); bb[1] = fmax(lbx[1], r[1]);
                bb[2] = fmin(lbx[2], r[2]); bb[3] = fmax(lbx[3], r[3]);
                N.aux = 1 + nd[N.left].aux + nd[N.left + 1].aux;
            }
        }
        __syncthreads();
    }
    // depth-first names top-down: fin := final index of the node
    if (threadIdx.x == 0) { nd[0].fin = 0; }
    __syncthreads();
    // rank pass needs the subtree counts of the left child before they are overwritten: keep counts in aux, ranks in `seg` (free now)
    int* rank = seg;  // per node (total <= n)
    if (threadIdx.x == 0) rank[0] = 0;
    __syncthreads();
    for (int lv = 0; lv <= level; lv++) {
        const int b0 = lvl_start[lv], b1 = lvl_start[lv + 1];
        for (int k = b0 + threadIdx.x; k < b1; k += KDB_THREADS) {
            const NodeW& N = nd[k];
            if (N.col < 0) continue;
            const int r = rank[k];
            nd[N.left].fin = (uint16_t)(1 + 2 * r);
            nd[N.left + 1].fin = (uint16_t)(2 + 2 * r);
            rank[N.left] = r + 1;
            rank[N.left + 1] = r + 1 + nd[N.left].aux;
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < total; k += KDB_THREADS) {
        const NodeW& N = nd[k];
        uco_kdnode o;
        if (N.col < 0) {
            o.divlow = 0.f; o.divhigh = 0.f; o.col = -1; o.left = -1; o.right = -1; o.leaf_begin = N.s; o.leaf_count = N.e - N.s;
        } else {
            o.divlow = (float)N.div; o.divhigh = N.divhigh; o.col = N.col;
            o.left = nd[N.left].fin; o.right = nd[N.left + 1].fin; o.leaf_begin = 0; o.leaf_count = 0;
        }
        out_nodes[N.fin] = o;
    }
    for (int i = threadIdx.x; i < n; i += KDB_THREADS) out_leaf[i] = (int32_t)all[i];
    if (threadIdx.x == 0) {
        A.n_nodes[f] = total;
        const double* bb = nbb;
        out_bbox[0] = bb[0]; out_bbox[1] = bb[1]; out_bbox[2] = bb[2]; out_bbox[3] = bb[3];
    }
}

size_t kdb_smem_bytes(int cap) {
    const size_t max_nodes = 2 * ((size_t)cap / 5) + 2;
    return max_nodes * (sizeof(NodeW) + 32) + (size_t)cap * (8 + 4 + 4 + 4 + 4 + 4) + 16;
}

}  // namespace

// internal launcher shared with track.cu
int uco_kdtree_build_launch(uco_b200_ctx* ctx, int n_frames, const uco_keypoint* kps_dev, size_t kps_stride, const int32_t* n_kp_dev,
                            int n_fixed, int cap, uco_kdnode* nodes_dev, int node_cap, int32_t* leaf_dev, double* bbox_dev,
                            int32_t* n_nodes_dev, int32_t* err_dev) {
    if (cap > UCO_KDTREE_DEV_MAX_POINTS) return uco_fail(ctx, UCO_E_CAPACITY, "kdtree_build_dev: %d points per frame exceed the shared-memory build's limit %d", cap, UCO_KDTREE_DEV_MAX_POINTS);
    if (node_cap < 2 * (cap / 5) + 2) return uco_fail(ctx, UCO_E_INVALID, "kdtree_build_dev: node capacity %d < %d", node_cap, 2 * (cap / 5) + 2);
    const size_t smem = kdb_smem_bytes(cap);
    static size_t configured = 0;   // grow-only opt-in; racing threads write the same attribute
    if (smem > configured) {
        UCO_CUDA(ctx, cudaFuncSetAttribute(kdtree_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    KdBuildArgs A;
    A.kps = kps_dev; A.kps_stride = kps_stride; A.n_kp = n_kp_dev; A.n_fixed = n_fixed; A.cap = cap; A.node_cap = node_cap;
    A.nodes = nodes_dev; A.leaf_idx = leaf_dev; A.bbox = bbox_dev; A.n_nodes = n_nodes_dev; A.err = err_dev;
    kdtree_build_kernel<<<n_frames, KDB_THREADS, smem, ctx->stream>>>(A);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

extern "C" {

int uco_b200_kdtree_build_batch_dev(uco_b200_ctx* ctx, int n_frames, const uco_keypoint* kps_dev, size_t kps_frame_stride,
                                    const int32_t* n_kp_dev, int cap, uco_kdnode* nodes_dev, int node_cap, int32_t* leaf_idx_dev,
                                    double* bbox_dev, int32_t* n_nodes_dev) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (n_frames <= 0 || cap <= 0 || !kps_dev || !n_kp_dev || !nodes_dev || !leaf_idx_dev || !bbox_dev || !n_nodes_dev)
        return uco_fail(ctx, UCO_E_INVALID, "kdtree_build_batch_dev: bad arguments");
    int32_t* err = (int32_t*)uco_ws(ctx, WS_KDTREE_ERR, 16);
    int32_t* herr = (int32_t*)uco_pinned(ctx, WS_KDTREE_ERR, 16);
    if (!err || !herr) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemsetAsync(err, 0, 4, ctx->stream));
    int rc = uco_kdtree_build_launch(ctx, n_frames, kps_dev, kps_frame_stride, n_kp_dev, 0, cap, nodes_dev, node_cap, leaf_idx_dev, bbox_dev,
                                     n_nodes_dev, err);
    if (rc != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaMemcpyAsync(herr, err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*herr) return uco_fail(ctx, UCO_E_CAPACITY, "kdtree_build_batch_dev: a frame %s", *herr == 1 ? "holds more keypoints than `cap`" : "needs more tree levels / nodes than the device build supports");
    return UCO_OK;
}

// host buffers in and out (tests, and callers that only want the tree): xy = n points, stride_bytes apart
int uco_b200_kdtree_build_dev(uco_b200_ctx* ctx, const float* xy, size_t stride_bytes, int n, uco_kdnode* nodes, int cap_nodes,
                              int32_t* leaf_idx, double* bbox4, int* n_nodes) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (n < 0 || !n_nodes || (n > 0 && (!xy || !nodes || !leaf_idx || !bbox4 || stride_bytes < 8))) return uco_fail(ctx, UCO_E_INVALID, "kdtree_build_dev: bad arguments");
    *n_nodes = 0;
    if (n == 0) return UCO_OK;
    const int node_cap = 2 * (n / 5) + 2;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += (b + 255) & ~(size_t)255; return o; };
    const size_t o_kps = take(sizeof(uco_keypoint) * (size_t)n), o_n = take(16), o_nodes = take(sizeof(uco_kdnode) * (size_t)node_cap),
                 o_leaf = take(4 * (size_t)n), o_bbox = take(32), o_nn = take(16), o_err = take(16);
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_KDTREE, off);
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_KDTREE, off);
    if (!d || !h) return UCO_E_NOMEM;
    uco_keypoint* hk = (uco_keypoint*)(h + o_kps);
    memset(hk, 0, sizeof(uco_keypoint) * (size_t)n);
    for (int i = 0; i < n; i++) {
        const float* p = (const float*)((const uint8_t*)xy + stride_bytes * (size_t)i);
        hk[i].x = p[0];
        hk[i].y = p[1];
    }
    *(int32_t*)(h + o_n) = n;
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d, h, o_nodes, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_err, 0, 4, s));
    int rc = uco_kdtree_build_launch(ctx, 1, (const uco_keypoint*)(d + o_kps), (size_t)n, (const int32_t*)(d + o_n), 0, n, (uco_kdnode*)(d + o_nodes),
                                     node_cap, (int32_t*)(d + o_leaf), (double*)(d + o_bbox), (int32_t*)(d + o_nn), (int32_t*)(d + o_err));
    if (rc != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaMemcpyAsync(h + o_nodes, d + o_nodes, off - o_nodes, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    if (*(int32_t*)(h + o_err)) return uco_fail(ctx, UCO_E_CAPACITY, "kdtree_build_dev: tree needs more levels / nodes than the device build supports");
    const int k = *(int32_t*)(h + o_nn);
    if (k > cap_nodes) return uco_fail(ctx, UCO_E_CAPACITY, "kdtree_build_dev: %d nodes, capacity %d", k, cap_nodes);
    memcpy(nodes, h + o_nodes, sizeof(uco_kdnode) * (size_t)k);
    memcpy(leaf_idx, h + o_leaf, 4 * (size_t)n);
    memcpy(bbox4, h + o_bbox, 32);
    *n_nodes = k;
    return UCO_OK;
}

// host: libstdc++ std::sort replay (sort_exact.h) on indices keyed by float values, for CPU unit tests of the restatement
int uco_b200_probe_sort_indices(uint32_t* idx, int n, const float* keys) {
    if (n < 0 || (n > 0 && (!idx || !keys))) return UCO_E_INVALID;
    auto less = [keys](uint32_t a, uint32_t b) -> bool { return keys[a] < keys[b]; };
    uco_sort::sort_(idx, idx + n, less);
    return UCO_OK;
}

}  // extern "C"
