// TEST INFRASTRUCTURE ONLY — stand-in for src/map_types/marker.h so that the reference's src/optimization/typesg2o.h
// compiles unchanged without OpenCV.  Only what typesg2o.h touches: Marker::get3DPointsLocalRefSystem
// (src/map_types/marker.cpp:58-62).
#pragma once
#include <opencv2/core/core.hpp>
#include <vector>
namespace ucoslam {
struct Marker {
    static std::vector<cv::Point3f> get3DPointsLocalRefSystem(float size) {
        return {cv::Point3f(-size / 2., size / 2., 0), cv::Point3f(size / 2., size / 2., 0), cv::Point3f(size / 2., -size / 2., 0),
                cv::Point3f(-size / 2., -size / 2., 0)};
    }
};
}  // namespace ucoslam
