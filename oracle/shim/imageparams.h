// TEST INFRASTRUCTURE ONLY — stand-in for the reference's src/imageparams.h with the members the adapters read
// (imageparams.h: CameraMatrix, Distorsion, CamSize, bl, fx()..cy()), for the adapters' compile check without OpenCV.
#pragma once
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
