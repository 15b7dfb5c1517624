// ref_project_wrap.cpp — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own projection matchers, compiled from
// its statements where they lie under /root/reference (cut out whole, unedited, by oracle/gen_ref_extract.py into oracle/_ref/gen):
//   Map::matchFrameToMapPoints                       src/map.cpp:651-770
//   the tracker's search by projection               src/utils/system.cpp:5921-6456 (macro-obfuscated; de-obfuscated with cpp)
//   Frame::project / predictScale / getKeyPointsInRegion   src/map_types/frame.h:129-161, frame.cpp:102-115
//   MapPoint::getViewCos / getDescDistance           src/map_types/mappoint.h:99,140-177
//   Se3Transform::inv / operator*                    src/basictypes/se3transform.h:89-112
//   filter_ambiguous_query                           src/basictypes/misc.cpp:117-150
//   the frame's kd-tree                              src/basictypes/picoflann.h (compiled unchanged)
// against container stand-ins (oracle/shim2: cv::Mat / Point3f as OpenCV's types.hpp defines them, Frame / MapPoint / Map with the
// reference's member names).  Used to PIN oracle/project_oracle.cpp and the CUDA matchers (tests/golden/project_ref.npz).
#include <cstdio>
#include "map.h"
#include "basictypes/misc.h"
#include "basictypes/timers.h"

namespace ucoslam {
#include "gen/misc_filters.inc"
cv::Mat computeF12(const cv::Mat&, const cv::Mat&, const cv::Mat&, const cv::Mat&) { return cv::Mat(); }
#include "gen/frame_region.inc"
#include "gen/map_match.inc"

// the member names below are the obfuscated ones system.cpp uses (obfs.txt): _9098980761384425343 = the System's map
struct System {
    std::shared_ptr<Map> _9098980761384425343;
    std::vector<cv::DMatch> _11946837405316294395(Frame&, Frame&, float, float);
};
#include "gen/system_tbp.inc"
}  // namespace ucoslam

namespace {
struct KP { float x, y, size, angle, response; int octave, class_id; };
void fill_frame(ucoslam::Frame& f, int n, const float* kp_xy, const int32_t* kp_octave, const unsigned char* desc, const float* scale_factors,
                int n_levels, float fx, float fy, float cx, float cy, const float* min_xy, const float* max_xy, const float* pose) {
    f.und_kpts.resize(n);
    for (int i = 0; i < n; i++) { f.und_kpts[i].pt = cv::Point2f(kp_xy[2 * i], kp_xy[2 * i + 1]); f.und_kpts[i].octave = kp_octave[i]; }
    f.desc = cv::Mat(n, 32, CV_8UC1);
    if (n) memcpy(f.desc.ptr<uchar>(0), desc, 32 * (size_t)n);
    f.ids.assign(n, std::numeric_limits<uint32_t>::max());
    f.flags.assign(n, Flag());
    f.scaleFactors.assign(scale_factors, scale_factors + n_levels);
    f.imageParams.CameraMatrix = cv::Mat::eye(3, 3, CV_32F);
    f.imageParams.CameraMatrix.at<float>(0, 0) = fx; f.imageParams.CameraMatrix.at<float>(1, 1) = fy;
    f.imageParams.CameraMatrix.at<float>(0, 2) = cx; f.imageParams.CameraMatrix.at<float>(1, 2) = cy;
    f.imageParams.CamSize = cv::Size(1, 1);
    f.minXY = cv::Point2f(min_xy[0], min_xy[1]);
    f.maxXY = cv::Point2f(max_xy[0], max_xy[1]);
    cv::Mat P(4, 4, CV_32F);
    memcpy(P.ptr<float>(0), pose, 64);
    f.pose_f2g = P;
    f.fseq_idx = 7;
    if (n) f.create_kdtree();
}
}  // namespace

extern "C" {
// Map::matchFrameToMapPoints on m map points whose ids are their row numbers (the reference gathers its candidates in ascending id
// order, map.h:202-235); out / visible as oracle_match_projected.  Returns the number of matches, < 0 on an exception.
int ref_match_projected(int m, const float* pos, const float* normal, const float* min_dist, const float* max_dist, const unsigned char* mp_desc,
                        int n_kp, const float* kp_xy, const int32_t* kp_octave, const unsigned char* kp_desc, const float* scale_factors, int n_levels,
                        float fx, float fy, float cx, float cy, const float* min_xy, const float* max_xy, const float* pose, float min_desc_dist,
                        float max_reproj_dist, cv::DMatch* out, unsigned char* visible) {
    try {
        ucoslam::Map map;
        ucoslam::Frame& kf = map.keyframes.add(0);                  // one keyframe that observes every point
        kf.idx = 0;
        for (int i = 0; i < m; i++) {
            ucoslam::MapPoint& p = map.map_points.add(i);
            p.id = i;
            p.pos3d = cv::Point3f(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
            p.normal = cv::Point3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
            p.mfMinDistance = min_dist[i]; p.mfMaxDistance = max_dist[i];
            p._desc = cv::Mat(1, 32, CV_8UC1);
            memcpy(p._desc.ptr<uchar>(0), mp_desc + 32 * (size_t)i, 32);
            kf.ids.push_back(i);
        }
        ucoslam::Frame cur;
        fill_frame(cur, n_kp, kp_xy, kp_octave, kp_desc, scale_factors, n_levels, fx, fy, cx, cy, min_xy, max_xy, pose);
        cv::Mat P(4, 4, CV_32F);
        memcpy(P.ptr<float>(0), pose, 64);
        std::vector<cv::DMatch> r = map.matchFrameToMapPoints({0u}, cur, P, min_desc_dist, max_reproj_dist, true, true);
        for (int i = 0; i < m; i++) visible[i] = map.map_points[i].nVisible > 0;
        for (size_t i = 0; i < r.size(); i++) out[i] = r[i];
        return (int)r.size();
    } catch (std::exception& e) {
        fprintf(stderr, "ref_match_projected: %s\n", e.what());
        return -1;
    }
}

// the tracker's search by projection: previous frame (octaves, descriptors, map point row per keypoint or -1) against the current frame
int ref_track_projected(int n_prev, const int32_t* prev_octave, const unsigned char* prev_desc, const int32_t* prev_row, int m, const uint32_t* ids,
                        const float* pos, int n_kp, const float* kp_xy, const int32_t* kp_octave, const unsigned char* kp_desc,
                        const float* scale_factors, int n_levels, float fx, float fy, float cx, float cy, const float* min_xy, const float* max_xy,
                        const float* pose, float dist_thr, float proj_dist_thr, cv::DMatch* out) {
    try {
        ucoslam::System sys;
        sys._9098980761384425343 = std::make_shared<ucoslam::Map>();
        ucoslam::Map& map = *sys._9098980761384425343;
        for (int i = 0; i < m; i++) {
            ucoslam::MapPoint& p = map.map_points.add(ids[i]);
            p.id = ids[i];
            p.pos3d = cv::Point3f(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        }
        ucoslam::Frame cur, prev;
        fill_frame(cur, n_kp, kp_xy, kp_octave, kp_desc, scale_factors, n_levels, fx, fy, cx, cy, min_xy, max_xy, pose);
        std::vector<float> zero(2 * (size_t)n_prev, 0.f);
        fill_frame(prev, 0, nullptr, nullptr, nullptr, scale_factors, n_levels, fx, fy, cx, cy, min_xy, max_xy, pose);
        prev.und_kpts.resize(n_prev);
        prev.desc = cv::Mat(n_prev, 32, CV_8UC1);
        if (n_prev) memcpy(prev.desc.ptr<uchar>(0), prev_desc, 32 * (size_t)n_prev);
        prev.ids.resize(n_prev);
        for (int i = 0; i < n_prev; i++) {
            prev.und_kpts[i].octave = prev_octave[i];
            prev.ids[i] = prev_row[i] < 0 ? std::numeric_limits<uint32_t>::max() : ids[prev_row[i]];
        }
        std::vector<cv::DMatch> r = sys._11946837405316294395(cur, prev, dist_thr, proj_dist_thr);
        for (size_t i = 0; i < r.size(); i++) out[i] = r[i];
        return (int)r.size();
    } catch (std::exception& e) {
        fprintf(stderr, "ref_track_projected: %s\n", e.what());
        return -1;
    }
}
}
