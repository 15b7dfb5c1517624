// orb_extractor_b200.h — the reference's extractor seam: a ucoslam::Feature2DSerializable
// (src/featureextractors/feature2dserializable.h:30-95) backed by uco_b200_orb_extract.  It answers exactly what
// ucoslam::ORBextractor answers (src/featureextractors/ORBextractor.h:86-116): F2D_ORB, DESC_ORB, min descriptor distance
// 50, FeatParams streamed as the raw struct (ORBextractor.cpp:417-423), sensitivity -> FAST thresholds
// (ORBextractor.cpp:457-466).  Registered in Feature2DSerializable::create / fromStream.
// Compiled and driven through the reference's own Feature2DSerializable (its real header + member definitions) by
// tests/adapters/adapter_world_test.cpp against container stand-ins for OpenCV (oracle/shim2).
#pragma once
#include <cstring>
#include <vector>
#include "featureextractors/feature2dserializable.h"
#include "uco_b200_cxx.h"

namespace ucoslam {

class ORBextractorB200 : public Feature2DSerializable {
public:
    explicit ORBextractorB200(int device = 0) : _ctx(device) { uco_b200_orb_default_params(&_prm); }

    Feature2DSerializable::FeatParams getParams() const override { return _featParams; }
    float getMinDescDistance() const override { return 50; }
    DescriptorTypes::Type getDescriptorType() const override { return DescriptorTypes::DESC_ORB; }
    bool& doGaussianBlur() { return _blur; }

    void setSensitivity(float v) override {  // ORBextractor.cpp:457-466
        if (v > 1) v = 1;
        if (v <= 0) v = 0;
        _featParams.sensitivity = v;
        v = 1 - v;
        _prm.ini_th_fast = v * 10 + 10;
        _prm.min_th_fast = v * 4 + 3;
    }
    float getSensitivity() override { return _featParams.sensitivity; }

protected:
    F2S_Type getType() const override { return Feature2DSerializable::F2D_ORB; }

    void detectAndCompute_impl(cv::InputArray image, cv::InputArray /*mask: ignored by the reference too*/,
                               std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors,
                               Feature2DSerializable::FeatParams params) override {
        static_assert(sizeof(cv::KeyPoint) == sizeof(uco_keypoint), "uco_keypoint mirrors cv::KeyPoint");
        cv::Mat im = image.getMat();
        if (im.empty()) return;                                        // ORBextractor.cpp:1250
        if (im.type() != CV_8UC1) throw std::runtime_error("ORBextractorB200: image must be CV_8UC1");
        if (!(params == _featParams)) {                                // ORBextractor.cpp:1142-1145 (thresholds reset to 20 / 7)
            _featParams = params;
            _prm.ini_th_fast = 20;
            _prm.min_th_fast = 7;
        }
        _prm.max_features = params.maxFeatures;
        _prm.n_levels = params.nOctaveLevels;
        _prm.scale_factor = params.scaleFactor;
        _prm.blur_first = _blur;
        const int cap = params.maxFeatures;
        keypoints.resize(cap);
        _desc.resize((size_t)cap * 32);
        int n = 0;
        _ctx.check(uco_b200_orb_extract(_ctx.get(), im.data, im.cols, im.rows, im.step[0], &_prm,
                                        reinterpret_cast<uco_keypoint*>(keypoints.data()), _desc.data(), cap, &n));
        keypoints.resize(n);
        if (n == 0) { descriptors.release(); return; }                 // ORBextractor.cpp:1281-1283
        descriptors.create(n, 32, CV_8U);
        std::memcpy(descriptors.getMat().data, _desc.data(), (size_t)n * 32);
    }
    void toStream_impl(std::ostream& str) override { str.write((char*)&_featParams, sizeof(_featParams)); }
    void fromStream_impl(std::istream& str) override { str.read((char*)&_featParams, sizeof(_featParams)); }

private:
    uco_b200::Context _ctx;
    uco_orb_params _prm;
    Feature2DSerializable::FeatParams _featParams;
    std::vector<uint8_t> _desc;
    bool _blur = true;
};

}  // namespace ucoslam
