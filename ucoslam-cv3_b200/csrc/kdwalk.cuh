// kdwalk.cuh — the radius search of the reference's kd-tree (picoflann) as a device-side walk, shared by the projection matchers
// (project.cu: Map::matchFrameToMapPoints; track.cu: the tracker's search by projection from the previous frame).
//   src/basictypes/picoflann.h:435-463   generalSearch / computeInitialDistances (float accumulator of double terms)
//   src/basictypes/picoflann.h:556-600   searchExactLevel: nearest child first, the other one deferred; leaf points tested in leaf order
// The VISIT ORDER is part of the result: the callers' best / second-best bookkeeping is order dependent (src/map.cpp:722-737).
// One thread walks one query with an explicit stack; `visit(kp, K)` is called for every keypoint inside the radius, in the
// reference's order.  Returns false when the deferred-branch stack overflowed (tree deeper than KD_STACK levels): the caller
// raises an error flag instead of silently dropping subtrees.
#pragma once
#include "common.cuh"

constexpr int KD_STACK = 48;

template <class Visit>
__device__ __forceinline__ bool kd_radius_walk(const uco_kdnode* __restrict__ nodes, const int32_t* __restrict__ leaf_idx, const double* bbox,
                                               const uco_keypoint* __restrict__ kps, const float q[2], double radius, Visit visit) {
    const double r2 = radius * radius;
    double d0 = 0, d1 = 0;
    float distsq = 0;
    {
        const double e0 = q[0], e1 = q[1];
        if (e0 < bbox[0]) { const double d = e0 - bbox[0]; d0 = d * d; distsq = (float)(distsq + d0); }
        if (e0 > bbox[1]) { const double d = e0 - bbox[1]; d0 = d * d; distsq = (float)(distsq + d0); }
        if (e1 < bbox[2]) { const double d = e1 - bbox[2]; d1 = d * d; distsq = (float)(distsq + d1); }
        if (e1 > bbox[3]) { const double d = e1 - bbox[3]; d1 = d * d; distsq = (float)(distsq + d1); }
    }
    int st_node[KD_STACK];
    double st_min[KD_STACK], st_d0[KD_STACK], st_d1[KD_STACK];
    int sp = 0;
    bool ok = true;
    st_node[0] = 0; st_min[0] = (double)distsq; st_d0[0] = d0; st_d1[0] = d1; sp = 1;
    while (sp > 0) {
        sp--;
        int node = st_node[sp];
        const double mind = st_min[sp];
        const double e0 = st_d0[sp], e1 = st_d1[sp];
        for (;;) {
            const uco_kdnode N = nodes[node];
            if (N.col < 0) {
                for (int t = 0; t < N.leaf_count; t++) {
                    const int kp = leaf_idx[N.leaf_begin + t];
                    const uco_keypoint K = kps[kp];
                    double dd = (double)(q[0] - K.x);   // L2::compute_distance: float difference, double square
                    double sqd = dd * dd;
                    if (!(sqd > r2)) {
                        dd = (double)(q[1] - K.y);
                        sqd += dd * dd;
                    }
                    if (!(sqd < r2)) continue;
                    visit(kp, K);
                }
                break;
            }
            const double val = (double)q[N.col];
            const double diff1 = val - (double)N.divlow, diff2 = val - (double)N.divhigh;
            int bestc, other;
            double cut;
            if (diff1 + diff2 < 0) { bestc = N.left; other = N.right; cut = diff2 * diff2; }
            else { bestc = N.right; other = N.left; cut = diff1 * diff1; }
            const float dst = (float)(N.col == 0 ? e0 : e1);
            const double mind2 = mind + cut - (double)dst;
            if (mind2 <= r2) {
                if (sp < KD_STACK) {
                    st_node[sp] = other; st_min[sp] = mind2;
                    st_d0[sp] = N.col == 0 ? cut : e0; st_d1[sp] = N.col == 0 ? e1 : cut;
                    sp++;
                } else ok = false;
            }
            node = bestc;   // mindistsq and dists are passed on unchanged to the nearer child
        }
    }
    return ok;
}
