// bow_b200.h — the reference's bag-of-words seam: fbow::Vocabulary::transform(features, level, fBow&, fBow2&)
// (3rdparty/fbow/fbow/fbow.cpp:51-90) as KPFrameDataBase::computeBow calls it (src/map_types/keyframedatabase.cpp:310-321).
// The device walks the vocabulary tree for every descriptor; this adapter folds the per-descriptor (word, weight, node)
// triples into the reference's containers IN DESCRIPTOR ORDER, exactly as fbow.h:428-436 does, so the float weight sums
// (r1[word] += weight) and the fBow2 index lists are bit-identical to the CPU result.
// Needs only fbow.h and a cv::Mat (rows, cols, type(), ptr<T>(r), step) -- real OpenCV or oracle/shim.
#pragma once
#include <sstream>
#include <vector>
#include <fbow/fbow.h>
#include "uco_b200_cxx.h"

namespace uco_b200 {

class VocabularyB200 {
public:
    explicit VocabularyB200(int device = 0) : _ctx(device) {}
    ~VocabularyB200() { uco_b200_bow_free(_ctx.get(), _voc); }

    // the byte stream of a .fbow file / fbow::Vocabulary::toStream (fbow.cpp:169-190)
    void fromStream(std::istream& str) {
        std::string bytes((std::istreambuf_iterator<char>(str)), std::istreambuf_iterator<char>());
        uco_b200_bow_free(_ctx.get(), _voc);
        _voc = nullptr;
        _ctx.check(uco_b200_bow_load(_ctx.get(), bytes.data(), bytes.size(), &_voc));
    }
    void fromVocabulary(fbow::Vocabulary& v) {
        std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
        v.toStream(ss);
        fromStream(ss);
    }
    bool isValid() const { return _voc != nullptr; }

    void transform(const cv::Mat& features, int level, fbow::fBow& result, fbow::fBow2& r2) {
        if (features.rows == 0) throw std::runtime_error("Vocabulary::transform No input data");            // fbow.cpp:52
        if (features.type() != CV_8UC1 || features.cols != 32)
            throw std::runtime_error("Vocabulary::transform features are of different type than vocabulary");  // fbow.cpp:53
        const int n = features.rows;
        _word.resize(n); _weight.resize(n); _node.resize(n);
        const size_t stride = n > 1 ? (size_t)(features.ptr<unsigned char>(1) - features.ptr<unsigned char>(0)) : 32;
        _ctx.check(uco_b200_bow_transform(_ctx.get(), _voc, features.ptr<unsigned char>(0), n, stride, level, _word.data(),
                                          _weight.data(), _node.data()));
        result.clear();
        r2.clear();
        for (int i = 0; i < n; i++) {                      // fbow.h:428-436, descriptor order
            if (_node[i] != 0xFFFFFFFFu) r2[_node[i]].push_back((uint32_t)i);
            if (_word[i] != 0xFFFFFFFFu) result[_word[i]] += _weight[i];
        }
    }

private:
    Context _ctx;
    uco_b200_voc* _voc = nullptr;
    std::vector<uint32_t> _word, _node;
    std::vector<float> _weight;
};

}  // namespace uco_b200
