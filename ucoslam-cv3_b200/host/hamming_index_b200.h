// hamming_index_b200.h — the reference's k-NN seam: an xflann::impl::IndexImpl (3rdparty/xflann/xflann/types.h:415-430)
// whose search() is the sm_100a brute-force Hamming kernel.  It stands where xflann::impl::Linear stands
// (impl/linear.h:36-90): Index::build selects it by ParamSet::type (index.cpp:45-69, see INTEGRATION.md for the two-line
// patch adding type "linear_b200"), FrameMatcher_Flann then calls trainIndex.search(...) unchanged
// (src/utils/framematcher.cpp:239).  Results are bit-identical to Linear's, including the heap order of each row.
// Compiles against the reference's xflann alone (no OpenCV needed).
#pragma once
#include <cstring>
#include <sstream>
#include <vector>
#include <xflann/types.h>
#include "uco_b200_cxx.h"

namespace xflann {
namespace impl {

class LinearB200 : public IndexImpl {
public:
    explicit LinearB200(int device = 0) : _ctx(device) {}
    std::string getName() const override { return "linear_b200"; }

    // Linear::build (impl/linear.cpp): keeps a reference to, or a copy of, the train rows; here always a dense copy
    void build(Matrix features, const ParamSet& /*params*/) {
        if (features.type() != XFLANN_8U || features.cols != 32)
            throw std::runtime_error("LinearB200::build only 32-byte binary descriptors (CV_8UC1 x 32) are supported");
        _rows = features.rows;
        _train.resize((size_t)_rows * 32);
        for (int r = 0; r < _rows; r++) std::memcpy(&_train[(size_t)r * 32], features.ptr<char>(r), 32);
    }
    bool storeFeatures() const override { return true; }
    uint32_t size() const override { return (uint32_t)_rows; }

    void search(Matrix features, int nn, Matrix indices, Matrix distances, const ParamSet& search_params,
                std::shared_ptr<cpu> = std::shared_ptr<cpu>()) override {
        if (features.type() != XFLANN_8U || features.cols != 32) throw std::runtime_error("LinearB200::search descriptor type");
        if (indices.type() != XFLANN_32S || distances.type() != XFLANN_32S || indices.cols != nn || distances.cols != nn)
            throw std::runtime_error("LinearB200::search indices / distances must be int32 with nn columns");
        if (search_params.count("maxDist") && search_params.asDouble("maxDist") != -1)
            throw std::runtime_error("LinearB200::search radius search (maxDist) is not supported");
        // always the heap order Linear leaves: Index::_search sorts afterwards when "sorted" is set (index.cpp:91-102)
        // strided inputs are fine (types.h:282-290 copies cv::Mat::step); outputs are dense per the ABI -> stage if strided
        std::vector<int32_t> idx((size_t)features.rows * nn), dist((size_t)features.rows * nn);
        _ctx.check(uco_b200_hamming_knn(_ctx.get(), (const uint8_t*)features.ptr<char>(0), features.rows, features.stride,
                                        (const uint8_t*)_train.data(), _rows, 32, nn, UCO_KNN_HEAP,
                                        idx.data(), dist.data()));
        for (int r = 0; r < features.rows; r++) {
            std::memcpy(indices.ptr<int>(r), &idx[(size_t)r * nn], sizeof(int32_t) * nn);
            std::memcpy(distances.ptr<int>(r), &dist[(size_t)r * nn], sizeof(int32_t) * nn);
        }
    }

    void toStream(std::ostream& str) const override {
        str.write((const char*)&_rows, sizeof(_rows));
        str.write((const char*)_train.data(), _train.size());
    }
    void fromStream(std::istream& str) override {
        str.read((char*)&_rows, sizeof(_rows));
        _train.resize((size_t)_rows * 32);
        str.read((char*)_train.data(), _train.size());
    }
    uint64_t hash() const override {
        uint64_t h = 1469598103934665603ull;
        for (unsigned char c : _train) h = (h ^ c) * 1099511628211ull;
        return h ^ (uint64_t)_rows;
    }

private:
    uco_b200::Context _ctx;
    std::vector<char> _train;
    int _rows = 0;
};

}  // namespace impl
}  // namespace xflann
