// ref_fbow_wrap.cpp — TEST INFRASTRUCTURE ONLY. C entry points around the REFERENCE's own fbow, compiled from
// /root/reference/3rdparty/fbow/fbow/fbow.cpp against oracle/shim (see oracle/Makefile); output goes to oracle/_ref/.
#include <fbow/fbow.h>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>

extern "C" {
void* ref_fbow_load_file(const char* path) {
    try { auto* v = new fbow::Vocabulary(); v->readFromFile(path); return v; } catch (std::exception&) { return nullptr; }
}
void* ref_fbow_load_bytes(const void* bytes, size_t n) {
    try {
        std::string s((const char*)bytes, n);
        std::istringstream is(s, std::ios::binary);
        auto* v = new fbow::Vocabulary(); v->fromStream(is); return v;
    } catch (std::exception&) { return nullptr; }
}
void ref_fbow_free(void* v) { delete (fbow::Vocabulary*)v; }
int ref_fbow_info(void* v, uint32_t* k, uint32_t* nblocks, uint32_t* desc_size) {
    auto* V = (fbow::Vocabulary*)v; *k = V->getK(); *nblocks = (uint32_t)V->size(); *desc_size = V->getDescSize(); return 0;
}
// Vocabulary::transform(features, level, fBow&, fBow2&) as keyframedatabase.cpp:319 calls it.
// bow: (word id, weight) pairs in map order; bow2: flattened (node id, feature idx) in map order / insertion order.
int ref_fbow_transform(void* v, const uint8_t* desc, int n, int level, uint32_t* bow_ids, float* bow_w, int* n_bow,
                       uint32_t* bow2_node, uint32_t* bow2_feat, int* n_bow2) {
    try {
        cv::Mat f(n, 32, CV_8UC1, (void*)desc);
        fbow::fBow b; fbow::fBow2 b2;
        ((fbow::Vocabulary*)v)->transform(f, level, b, b2);
        int i = 0; for (auto& e : b) { bow_ids[i] = e.first; bow_w[i] = (float)e.second; i++; } *n_bow = i;   // fBow value type
        i = 0; for (auto& e : b2) for (auto x : e.second) { bow2_node[i] = e.first; bow2_feat[i] = x; i++; } *n_bow2 = i;
        return 0;
    } catch (std::exception&) { return -1; }
}
double ref_fbow_score(const uint32_t* ids1, const float* w1, int n1, const uint32_t* ids2, const float* w2, int n2) {
    fbow::fBow a, b;
    for (int i = 0; i < n1; i++) { float t = w1[i]; a[ids1[i]] = t; }
    for (int i = 0; i < n2; i++) { float t = w2[i]; b[ids2[i]] = t; }
    return fbow::fBow::score(a, b);
}
}
