// TEST INFRASTRUCTURE ONLY — stand-in for src/basictypes/se3transform.h (Se3Transform : cv::Mat, 4x4 CV_32F).  The accessors mirror
// se3transform.h:24-88; inv() and operator*(Point3f) are the REFERENCE's own statements (:89-112), cut out by gen_ref_extract.py.
#pragma once
#include "se3.h"
namespace ucoslam {
class Se3Transform : public cv::Mat {
public:
    Se3Transform(bool makeInvalid = false) {
        create(4, 4, CV_32F);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) at<float>(i, j) = makeInvalid ? std::numeric_limits<float>::quiet_NaN() : (i == j ? 1.f : 0.f);
    }
    Se3Transform(const Se3Transform& R) { create(4, 4, CV_32F); memcpy(ptr<float>(0), R.ptr<float>(0), 64); }
    bool isValid() const { return !std::isnan(at<float>(0, 0)); }
    inline Se3Transform operator=(const cv::Mat& m) { memcpy(ptr<float>(0), m.ptr<float>(0), 64); return *this; }
    inline Se3Transform& operator=(const Se3Transform& R) { memcpy(ptr<float>(0), R.ptr<float>(0), 64); return *this; }
    inline float& operator[](uint32_t idx) { return ptr<float>()[idx]; }
    inline float operator[](uint32_t idx) const { return ptr<float>()[idx]; }
    inline float at_(uint32_t idx) const { return ptr<float>()[idx]; }
#include "gen/se3transform_ops.inc"
};
}
