// TEST INFRASTRUCTURE ONLY — stand-in for src/map_types/mappoint.h: the members the projection matchers / solvePnp read
// (mappoint.h:40-100), with getViewCos and the descriptor distances as the reference's OWN statements (:99, :140-177).
#pragma once
#include <opencv2/core/core.hpp>
namespace ucoslam {
class MapPoint {
public:
    uint32_t id = std::numeric_limits<uint32_t>::max();
    std::map<uint32_t, uint32_t> frames;            // frame id -> keypoint index
    uint32_t lastFIdxSeen = std::numeric_limits<uint32_t>::max();
    bool isBad() const { return bad; }
    bool isStable() const { return stable; }
    bool isStereo() const { return stereo; }
    cv::Point3f getCoordinates() const { return pos3d; }
    void setCoordinates(const cv::Point3f& p) { pos3d = p; }
    cv::Point3f getNormal() const { return normal; }
    float getMinDistanceInvariance() const { return mfMinDistance; }
    float getMaxDistanceInvariance() const { return mfMaxDistance; }
    inline void getDescriptor(cv::Mat& copy) const { _desc.copyTo(copy); }
    inline float getDescDistance(const cv::Mat& dsc2, int row) const { return getDescDistance(_desc, 0, dsc2, row); }   // mappoint.h:84
    void setVisible() { nVisible++; }
    void setSeen() { nSeen++; }
    // state (public here: the checkers fill it directly)
    cv::Point3f pos3d, normal;
    float mfMinDistance = 0, mfMaxDistance = 0;
    cv::Mat _desc;
    bool bad = false, stable = true, stereo = false;
    int nVisible = 0, nSeen = 0;
#include "gen/mappoint_inline.inc"
};
}
