// TEST INFRASTRUCTURE ONLY — stand-in for the reference's src/map_types/frame.h (which needs OpenCV's calib3d / features2d
// headers) with exactly the members src/map_types/keyframedatabase.cpp touches (frame.h:60-64,104-106: idx, desc, bowvector,
// bowvector_level) and FrameSet as the id -> Frame container it indexes (frameset: count / operator[]), so that the reference's
// keyframe database compiles unchanged into oracle/_ref/libref_kfdb.so; plus the members the stereo / triangulation / undistortion
// adapters read, so that those headers are at least compiled (tests/adapters/adapter_syntax.cpp).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <opencv2/core/core.hpp>
#include <fbow/fbow.h>
#include <vector>
#include "imageparams.h"
namespace ucoslam {
class Frame {
public:
    uint32_t idx = 0;
    cv::Mat desc;
    std::shared_ptr<fbow::fBow> bowvector = std::make_shared<fbow::fBow>();
    std::shared_ptr<fbow::fBow2> bowvector_level = std::make_shared<fbow::fBow2>();
    // read by the stereo / triangulation adapters' compile check (frame.h:66-90)
    std::vector<cv::KeyPoint> und_kpts;
    std::vector<float> depth, scaleFactors;
    ImageParams imageParams;
};
class FrameSet : public std::map<uint32_t, Frame> {};
}
