// TEST INFRASTRUCTURE ONLY — stand-in for src/map_types/frame.h: the data members of ucoslam::Frame with the reference's names and
// types (frame.h:56-88), its kd-tree as the reference's OWN picoflann index (basictypes/picoflann.h compiles unchanged), and the
// reference's OWN statements of predictScale / project (frame.h:129-161) and getKeyPointsInRegion (frame.cpp:102-115) cut out by
// oracle/gen_ref_extract.py.  Everything that needs OpenCV algorithms (extraction, undistortion, streams) is left out.
#pragma once
#include <opencv2/core/core.hpp>
#include "basictypes/picoflann.h"
#include "basictypes/se3transform.h"
#include "basictypes/flag.h"
#include "imageparams.h"
#include "ucoslamtypes.h"
#include "marker.h"
using namespace std;
namespace fbow { struct fBow; struct fBow2; }
namespace ucoslam {
class Frame {
    struct KdTreeKeyPoints {
        inline float operator()(const cv::KeyPoint& kp, int dim) const { return dim == 0 ? kp.pt.x : kp.pt.y; }
    };
public:
    enum FlagsTypes : uint8_t { FLAG_NONMAXIMA = 0x01, FLAG_OUTLIER = 0x02, FLAG_BAD = 0x04 };
    uint32_t idx = std::numeric_limits<uint32_t>::max();
    std::vector<ucoslam::MarkerObservation> markers;
    picoflann::KdTreeIndex<2, KdTreeKeyPoints> keypoint_kdtree;
    cv::Mat desc;
    std::vector<uint32_t> ids;
    std::vector<Flag> flags;
    Se3Transform pose_f2g;
    std::vector<cv::KeyPoint> und_kpts;
    std::vector<cv::Point2f> kpts;
    std::shared_ptr<fbow::fBow> bowvector;
    std::shared_ptr<fbow::fBow2> bowvector_level;
    uint32_t fseq_idx = std::numeric_limits<uint32_t>::max();
    vector<float> scaleFactors;
    ImageParams imageParams;
    bool isBad() const { return frame_flags.is(FLAG_BAD); }
    cv::Point minXY = cv::Point2f(0, 0), maxXY = cv::Point2f(std::numeric_limits<float>::max(), std::numeric_limits<float>::max());
    float getDepth(int i) const { return depth.empty() ? 0 : depth[i]; }   // frame.cpp: 0 without depth
    std::vector<float> depth;
    MarkerObservation getMarker(uint32_t id) const { for (auto& m : markers) if (m.id == id) return m; throw std::runtime_error("Frame::getMarker"); }
    void create_kdtree() { keypoint_kdtree.build(und_kpts); }              // frame.h:124-127
    std::vector<uint32_t> getKeyPointsInRegion(cv::Point2f p, float radius, int minScaleLevel = 0, int maxScaleLevel = std::numeric_limits<int>::max()) const;
#include "gen/frame_inline.inc"
private:
    Flag frame_flags;
};
}
