// undistort_b200.h — the reference's undistortion helper ucoslam::undistortPoints(points_io, ImageParams, out)
// (src/basictypes/misc.h, misc.cpp:269-292), which FrameExtractor applies to the extracted keypoints (Frame::und_kpts), the marker
// corners and the image bounds when it builds a Frame.  Same signature plus the context.  In-place use (out == nullptr, the only
// form the reference calls) is identical.  With `out` given this adapter writes PIXEL coordinates to *out; the reference's out
// branch (misc.cpp:287-291) runs its rescale loop over an empty vector and so leaves *out in normalised camera coordinates — a
// deliberate difference: no caller relies on that branch and the in-place meaning is the documented one.
// Compile inside the reference tree (needs its headers and OpenCV C++; see the note in orb_extractor_b200.h).
#pragma once
#include <vector>
#include "imageparams.h"
#include "uco_b200_cxx.h"

namespace ucoslam {

inline void undistortPoints_b200(uco_b200::Context& ctx, std::vector<cv::Point2f>& points_io, const ImageParams& ip,
                                 std::vector<cv::Point2f>* out = nullptr) {
    const float K[4] = {ip.CameraMatrix.at<float>(0, 0), ip.CameraMatrix.at<float>(1, 1), ip.CameraMatrix.at<float>(0, 2),
                        ip.CameraMatrix.at<float>(1, 2)};
    cv::Mat d;
    ip.Distorsion.convertTo(d, CV_32F);
    d = d.reshape(1, 1);
    std::vector<cv::Point2f>& dst = out ? *out : points_io;
    if (out) out->resize(points_io.size());
    if (points_io.empty()) return;
    ctx.check(uco_b200_undistort_points(ctx.get(), &points_io[0].x, sizeof(cv::Point2f), (int)points_io.size(), K,
                                        d.total() ? d.ptr<float>(0) : nullptr, (int)d.total(), &dst[0].x, sizeof(cv::Point2f)));
}

// Frame::und_kpts = kpts with pt undistorted (frameextractor.cpp, after the extractor thread joins)
inline void undistortKeyPoints_b200(uco_b200::Context& ctx, const std::vector<cv::KeyPoint>& kpts, const ImageParams& ip,
                                    std::vector<cv::KeyPoint>& und_kpts) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(uco_keypoint), "layouts");
    und_kpts = kpts;
    if (kpts.empty()) return;
    const float K[4] = {ip.CameraMatrix.at<float>(0, 0), ip.CameraMatrix.at<float>(1, 1), ip.CameraMatrix.at<float>(0, 2),
                        ip.CameraMatrix.at<float>(1, 2)};
    cv::Mat d;
    ip.Distorsion.convertTo(d, CV_32F);
    d = d.reshape(1, 1);
    ctx.check(uco_b200_undistort_points(ctx.get(), &und_kpts[0].pt.x, sizeof(cv::KeyPoint), (int)und_kpts.size(), K,
                                        d.total() ? d.ptr<float>(0) : nullptr, (int)d.total(), &und_kpts[0].pt.x, sizeof(cv::KeyPoint)));
}

}  // namespace ucoslam
