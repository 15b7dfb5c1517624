// TEST INFRASTRUCTURE ONLY — stand-in for src/map_types/marker.h: the members the adapters / typesg2o.h read (marker.h:33-56, 87-110)
#pragma once
#include <opencv2/core/core.hpp>
#include "basictypes/se3transform.h"
namespace ucoslam {
struct Marker {
    uint32_t id = 0;
    Se3Transform pose_g2m = Se3Transform(true);
    float size = 0;
    std::set<uint32_t> frames;
    static std::vector<cv::Point3f> get3DPointsLocalRefSystem(float size) {   // marker.cpp:58-62
        return {cv::Point3f(-size / 2., size / 2., 0), cv::Point3f(size / 2., size / 2., 0), cv::Point3f(size / 2., -size / 2., 0),
                cv::Point3f(-size / 2., -size / 2., 0)};
    }
};
struct MarkerPosesIPPE { cv::Mat sols[2]; double errs[2] = {0, 0}; double err_ratio = 0; };
struct MarkerObservation {
    uint32_t id = 0;
    std::vector<cv::Point2f> und_corners;
    MarkerPosesIPPE poses;
};
}
