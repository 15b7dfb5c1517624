// TEST INFRASTRUCTURE ONLY — stand-in for src/basictypes/se3.h: the pose handle the adapters pass around, kept as its 4x4 matrix
// (the reference stores a Rodrigues vector + translation and converts with cv::Rodrigues; the checkers never need that form).
#pragma once
#include <opencv2/core/core.hpp>
namespace ucoslam {
struct se3 {
    float m[16];
    se3() { for (int i = 0; i < 16; i++) m[i] = std::numeric_limits<float>::quiet_NaN(); }
    se3(const cv::Mat& rt) { *this = rt; }
    se3& operator=(const cv::Mat& rt) { memcpy(m, rt.ptr<float>(0), 64); return *this; }
    cv::Mat convert() const { cv::Mat o(4, 4, CV_32F); memcpy(o.ptr<float>(0), m, 64); return o; }
    operator cv::Mat() const { return convert(); }
    bool isValid() const { return !std::isnan(m[0]); }
};
}
