// ref_xflann_wrap.cpp — TEST INFRASTRUCTURE ONLY. C entry points around the REFERENCE's own xflann, compiled from the
// sources where they lie under /root/reference/3rdparty/xflann (see oracle/Makefile); output goes to oracle/_ref/.
// Used to validate oracle/knn_oracle.c, to generate tests/golden/knn_*.npz and as bench.py's CPU "reference" arm.
#include <xflann/xflann.h>
#include <cstdint>
#include <cstring>
#include <string>

extern "C" {
// type: 0 = LinearParams (exact), 1 = HKMeansParams(32,0) (what FrameMatcher_Flann builds, framematcher.cpp:213)
// max_checks / sorted: KnnSearchParams (framematcher.cpp:239 uses 16,false)
int ref_xflann_knn(const uint8_t* q, int nq, const uint8_t* t, int nt, int k, int type, int max_checks, int sorted,
                   int32_t* idx, int32_t* dist) {
    try {
        xflann::Matrix T(XFLANN_8U, nt, 32, (void*)t);
        xflann::Matrix Q(XFLANN_8U, nq, 32, (void*)q);
        xflann::Matrix I(XFLANN_32S, nq, k, idx);
        xflann::Matrix D(XFLANN_32S, nq, k, dist);
        xflann::Index index;
        if (type == 0) index.build(T, xflann::LinearParams());
        else index.build(T, xflann::HKMeansParams(32, 0));
        bool ok = index.search(Q, k, I, D, xflann::KnnSearchParams(max_checks, sorted != 0));
        return ok ? 0 : 1;
    } catch (std::exception& e) {
        return -1;
    }
}

// the reference's usage pattern: FrameMatcher_Flann::setParams builds the index over the train frame once (framematcher.cpp:200-215),
// every match / matchEpipolar call searches it (:239)
struct RefIndex { xflann::Index index; };
void* ref_xflann_build(const uint8_t* t, int nt, int type) {
    try {
        RefIndex* r = new RefIndex();
        xflann::Matrix T(XFLANN_8U, nt, 32, (void*)t);
        if (type == 0) r->index.build(T, xflann::LinearParams());
        else r->index.build(T, xflann::HKMeansParams(32, 0));
        return r;
    } catch (std::exception& e) {
        return nullptr;
    }
}
int ref_xflann_search(void* h, const uint8_t* q, int nq, int k, int max_checks, int sorted, int32_t* idx, int32_t* dist) {
    try {
        xflann::Matrix Q(XFLANN_8U, nq, 32, (void*)q);
        xflann::Matrix I(XFLANN_32S, nq, k, idx);
        xflann::Matrix D(XFLANN_32S, nq, k, dist);
        return ((RefIndex*)h)->index.search(Q, k, I, D, xflann::KnnSearchParams(max_checks, sorted != 0)) ? 0 : 1;
    } catch (std::exception& e) {
        return -1;
    }
}
void ref_xflann_free(void* h) { delete (RefIndex*)h; }
}
