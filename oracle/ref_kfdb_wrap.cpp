// ref_kfdb_wrap.cpp — TEST INFRASTRUCTURE ONLY. C entry points around the REFERENCE's own keyframe database:
// /root/reference/src/map_types/keyframedatabase.cpp and covisgraph.cpp compiled UNCHANGED where they lie (oracle/Makefile),
// together with the reference's fbow.  The only stand-ins are oracle/shim/opencv2/core/core.hpp (cv::Mat as a container)
// and oracle/shim/map_types/frame.h (a Frame with the four members keyframedatabase.cpp touches: idx, desc, bowvector,
// bowvector_level; FrameSet as the id -> Frame map it indexes).  Output goes to oracle/_ref/libref_kfdb.so.
#include <map_types/keyframedatabase.h>
#include <map_types/frame.h>
#include <map_types/covisgraph.h>
#include <cstdint>
#include <cstring>
#include <set>
#include <sstream>
#include <vector>

namespace {
struct RefDb {
    ucoslam::KeyFrameDataBase db;
    ucoslam::FrameSet fset;
    ucoslam::CovisGraph covis;
};
ucoslam::Frame make_frame(uint32_t idx, const uint8_t* desc, int n) {
    ucoslam::Frame f;
    f.idx = idx;
    f.desc = cv::Mat(n, 32, CV_8UC1, (void*)desc).clone();
    return f;
}
}  // namespace

extern "C" {
void* ref_kfdb_create(const char* voc_path) {
    try { auto* h = new RefDb(); h->db.loadFromFile(voc_path); return h; } catch (std::exception&) { return nullptr; }
}
void ref_kfdb_free(void* h) { delete (RefDb*)h; }
// the sections Map::toStream writes for these two members (map.cpp:316-325): KeyFrameDataBase::toStream (keyframedatabase.cpp:335-340 over
// KPFrameDataBase::toStream_, :278-287: vocabulary, inverted word index, frame ids) and CovisGraph::toStream (covisgraph.cpp:308-333)
long ref_kfdb_to_stream(void* h, unsigned char* out, long cap) {
    try {
        std::stringstream ss;
        ((RefDb*)h)->db.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -(long)b.size();
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return 0; }
}
long ref_covis_to_stream(void* h, unsigned char* out, long cap) {
    try {
        std::stringstream ss;
        ((RefDb*)h)->covis.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -(long)b.size();
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return 0; }
}
// KeyFrameDataBase::add (keyframedatabase.cpp:150-160); returns the frame's bag of words (map order) for the caller
int ref_kfdb_add(void* h, uint32_t idx, const uint8_t* desc, int n, uint32_t* bow_ids, float* bow_w, int* n_bow) {
    try {
        RefDb* R = (RefDb*)h;
        R->fset[idx] = make_frame(idx, desc, n);
        if (!R->db.add(R->fset[idx])) return -2;
        int i = 0;
        for (auto& e : *R->fset[idx].bowvector) { bow_ids[i] = e.first; bow_w[i] = (float)e.second; i++; }
        *n_bow = i;
        return 0;
    } catch (std::exception&) { return -1; }
}
int ref_kfdb_del(void* h, uint32_t idx) {
    try {
        RefDb* R = (RefDb*)h;
        if (!R->fset.count(idx)) return -2;
        R->db.del(R->fset[idx]);
        R->fset.erase(idx);
        return 0;
    } catch (std::exception&) { return -1; }
}
void ref_kfdb_covis_edge(void* h, uint32_t a, uint32_t b, float w) { ((RefDb*)h)->covis.createIncreaseEdge(a, b, w); }
// KeyFrameDataBase::relocalizationCandidates (keyframedatabase.cpp:195-276); also returns the query's bag of words
int ref_kfdb_query(void* h, const uint8_t* desc, int n, int sorted, float min_score, const uint32_t* excluded, int n_excluded,
                   uint32_t* out, int cap, uint32_t* bow_ids, float* bow_w, int* n_bow) {
    try {
        RefDb* R = (RefDb*)h;
        ucoslam::Frame q = make_frame(0xFFFFFFF0u, desc, n);
        std::set<uint32_t> exc(excluded, excluded + n_excluded);
        std::vector<uint32_t> c = R->db.relocalizationCandidates(q, R->fset, R->covis, sorted != 0, min_score, exc);
        int i = 0;
        for (auto& e : *q.bowvector) { bow_ids[i] = e.first; bow_w[i] = (float)e.second; i++; }
        *n_bow = i;
        if ((int)c.size() > cap) return -3;
        for (size_t k = 0; k < c.size(); k++) out[k] = c[k];
        return (int)c.size();
    } catch (std::exception&) { return -1; }
}
// the same two calls for callers that already hold bags of words (frames without descriptors: computeBow leaves the given
// bowvector alone, keyframedatabase.cpp:311) -- used to time the reference's query on large synthetic databases
int ref_kfdb_add_bow(void* h, uint32_t idx, const uint32_t* ids, const float* w, int n) {
    try {
        RefDb* R = (RefDb*)h;
        ucoslam::Frame f;
        f.idx = idx;
        for (int i = 0; i < n; i++) { float t = w[i]; (*f.bowvector)[ids[i]] = t; }
        R->fset[idx] = f;
        return R->db.add(R->fset[idx]) ? 0 : -2;
    } catch (std::exception&) { return -1; }
}
int ref_kfdb_query_bow(void* h, const uint32_t* ids, const float* w, int n, int sorted, float min_score, const uint32_t* excluded,
                       int n_excluded, uint32_t* out, int cap) {
    try {
        RefDb* R = (RefDb*)h;
        ucoslam::Frame q;
        q.idx = 0xFFFFFFF0u;
        for (int i = 0; i < n; i++) { float t = w[i]; (*q.bowvector)[ids[i]] = t; }
        std::set<uint32_t> exc(excluded, excluded + n_excluded);
        std::vector<uint32_t> c = R->db.relocalizationCandidates(q, R->fset, R->covis, sorted != 0, min_score, exc);
        if ((int)c.size() > cap) return -3;
        for (size_t k = 0; k < c.size(); k++) out[k] = c[k];
        return (int)c.size();
    } catch (std::exception&) { return -1; }
}
// KeyFrameDataBase::score (keyframedatabase.cpp:304-308): the float the loop detector thresholds
float ref_kfdb_score(void* h, uint32_t a, uint32_t b) {
    RefDb* R = (RefDb*)h;
    return R->db.score(R->fset[a], R->fset[b]);
}
}
