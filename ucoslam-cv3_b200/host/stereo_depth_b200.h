// stereo_depth_b200.h — the reference's stereo association seam.  FrameExtractor::processStereo (src/utils/frameextractor.cpp:1410-2634,
// private and macro-obfuscated) has no virtual interface: after it has built the left Frame and extracted the right image's
// keypoints, the rest of its body -- row buckets, per-keypoint Hamming search, 6x6 SAD refinement, Frame::depth -- is replaced by one
// call of ucoslam::stereoDepth_b200 with the same inputs it reads: the two grey images the keypoints were extracted from, the
// frame (und_kpts, desc -> depth), the right keypoints / descriptors, ImageParams::bl and fx(), Params::maxDescDistance.
// Compile inside the reference tree (needs its headers and OpenCV C++; see the note in orb_extractor_b200.h).
#pragma once
#include <vector>
#include "map_types/frame.h"
#include "uco_b200_cxx.h"

namespace ucoslam {

// returns the number of keypoints that received a depth (the reference's nmatches counter)
inline int stereoDepth_b200(uco_b200::Context& ctx, const cv::Mat& leftGrey, const cv::Mat& rightGrey, Frame& frame,
                            const std::vector<cv::KeyPoint>& rightKpts, const cv::Mat& rightDesc, float bl, float fx,
                            float maxDescDistance) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(uco_keypoint), "layouts");
    if (leftGrey.type() != CV_8UC1 || rightGrey.type() != CV_8UC1 || leftGrey.size() != rightGrey.size())
        throw std::runtime_error("stereoDepth_b200: two grey images of one size expected");
    if ((frame.desc.rows && (frame.desc.type() != CV_8UC1 || frame.desc.cols != 32)) ||
        (rightDesc.rows && (rightDesc.type() != CV_8UC1 || rightDesc.cols != 32)))
        throw std::runtime_error("stereoDepth_b200: 256-bit binary descriptors only");
    const int nl = (int)frame.und_kpts.size(), nr = (int)rightKpts.size();
    frame.depth.assign(nl, 0.f);                                   // frameextractor.cpp: depth.resize + zero fill
    if (nl == 0) return 0;
    int n = 0;
    ctx.check(uco_b200_stereo_depth(ctx.get(), leftGrey.ptr<uchar>(0), leftGrey.step[0], rightGrey.ptr<uchar>(0), rightGrey.step[0],
                                    leftGrey.cols, leftGrey.rows, reinterpret_cast<const uco_keypoint*>(frame.und_kpts.data()),
                                    frame.desc.ptr<uchar>(0), frame.desc.step[0], nl,
                                    reinterpret_cast<const uco_keypoint*>(rightKpts.data()), nr ? rightDesc.ptr<uchar>(0) : nullptr,
                                    nr ? rightDesc.step[0] : 32, nr, maxDescDistance, bl, fx, frame.depth.data(), nullptr, &n));
    return n;
}

}  // namespace ucoslam
