#pragma once   // TEST INFRASTRUCTURE ONLY: stand-in for src/imageparams.h with the members the matcher / tracker code reads
#include <opencv2/core/core.hpp>
namespace ucoslam {
class ImageParams {
public:
    cv::Mat CameraMatrix, Distorsion;
    cv::Size CamSize;
    float bl = 0;
    float fx() const { return CameraMatrix.at<float>(0, 0); }
    float fy() const { return CameraMatrix.at<float>(1, 1); }
    float cx() const { return CameraMatrix.at<float>(0, 2); }
    float cy() const { return CameraMatrix.at<float>(1, 2); }
};
}
