// ref_picoflann_wrap.cpp — TEST INFRASTRUCTURE ONLY.  Thin C entry points over the reference's OWN kd-tree
// (/root/reference/src/basictypes/picoflann.h, header-only, compiled where it lies) so that the restatement in
// oracle/project_oracle.cpp and the product's host-side tree builder can be pinned against it: the serialised tree
// (KdTreeIndex::toStream, the format Frame::toStream embeds, frame.cpp:294) and radius searches in visit order
// (what Frame::getKeyPointsInRegion consumes, frame.cpp:102-115).
#include "basictypes/picoflann.h"
#include <sstream>
#include <cstring>

namespace {
struct Pt { float x, y; };
struct Adapter {
    inline float operator()(const Pt& p, int dim) const { return dim == 0 ? p.x : p.y; }  // Frame::KdTreeKeyPoints, frame.h:50-54
};
}

extern "C" {
// returns the number of bytes of the serialised tree (<= cap) or -1
long ref_picoflann_stream(const float* xy, int n, unsigned char* out, long cap) {
    std::vector<Pt> pts(n);
    for (int i = 0; i < n; i++) pts[i] = {xy[2 * i], xy[2 * i + 1]};
    picoflann::KdTreeIndex<2, Adapter> tree;
    tree.build(pts);
    std::stringstream ss;
    tree.toStream(ss);
    std::string s = ss.str();
    if ((long)s.size() > cap) return -1;
    memcpy(out, s.data(), s.size());
    return (long)s.size();
}
int ref_picoflann_radius(const float* xy, int n, const float* queries, const float* radii, int nq, int* out_ptr, int* out_idx, int cap) {
    std::vector<Pt> pts(n);
    for (int i = 0; i < n; i++) pts[i] = {xy[2 * i], xy[2 * i + 1]};
    picoflann::KdTreeIndex<2, Adapter> tree;
    tree.build(pts);
    int tot = 0;
    for (int i = 0; i < nq; i++) {
        out_ptr[i] = tot;
        Pt q{queries[2 * i], queries[2 * i + 1]};
        auto res = tree.radiusSearch(pts, q, radii[i], false);   // float radius widened to double, as getKeyPointsInRegion does
        for (auto& r : res) {
            if (tot >= cap) return -1;
            out_idx[tot++] = (int)r.first;
        }
    }
    out_ptr[nq] = tot;
    return tot;
}
}
