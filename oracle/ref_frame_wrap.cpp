// ref_frame_wrap.cpp — TEST INFRASTRUCTURE ONLY.  The reference's OWN stream statements for a keyframe — Frame::toStream / fromStream
// (src/map_types/frame.cpp:260-341), MarkerObservation / MarkerPosesIPPE, cv::Mat / std::string (src/basictypes/io_utils.cpp),
// ImageParams (src/imageparams.cpp:68-84), Se3Transform (se3transform.h:178-188) — cut out where they lie by oracle/gen_ref_extract.py
// and compiled against container stand-ins (oracle/shim2 cv::Mat / Point / KeyPoint; the image has no OpenCV C++ headers), together
// with the reference's real io_utils.h templates, flag.h, picoflann.h and fbow.  The classes below only DECLARE the members those
// statements touch, with the reference's names and types (frame.h:56-88, marker.h:57-96, imageparams.h:33-40).
// Used by tests/test_frame_stream.py to pin csrc/frame_stream.cu: bytes written by the reference == bytes written by the codec.
#include <opencv2/core/core.hpp>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <basictypes/io_utils.h>      // the reference's (the include path puts its src/ first)
#include <basictypes/flag.h>
#include <basictypes/picoflann.h>
#include <fbow/fbow.h>
using namespace std;
namespace ucoslam {
struct DescriptorTypes { enum Type : std::int8_t { DESC_NONE = 0, DESC_ORB = 1, DESC_AKAZE = 2, DESC_BRISK = 3, DESC_FREAK = 4, DESC_SURF = 5 }; };   // ucoslamtypes.h:42
class Se3Transform : public cv::Mat {
public:
    Se3Transform() { create(4, 4, CV_32F); for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) at<float>(i, j) = i == j; }
#include "gen/se3transform_streams.inc"
};
struct MarkerPosesIPPE {
    cv::Mat sols[2];
    double errs[2];
    double err_ratio;
    void toStream(std::ostream& str) const;
    void fromStream(std::istream& str);
};
class MarkerObservation {
public:
    std::vector<cv::Point2f> corners, und_corners;
    float ssize;
    int id;
    std::string dict_info;
    MarkerPosesIPPE poses;
    void toStream(std::ostream& str) const;
    void fromStream(std::istream& str);
};
class ImageParams {
public:
    cv::Mat CameraMatrix, Distorsion;
    cv::Size CamSize;
    float bl = 0;
    float rgb_depthscale = 1;
    void toStream(std::ostream& str) const;
    void fromStream(std::istream& str);
};
class Frame {
    struct KdTreeKeyPoints {
        inline float operator()(const cv::KeyPoint& kp, int dim) const { return dim == 0 ? kp.pt.x : kp.pt.y; }
    };
public:
    uint32_t idx = std::numeric_limits<uint32_t>::max();
    std::vector<MarkerObservation> markers;
    picoflann::KdTreeIndex<2, KdTreeKeyPoints> keypoint_kdtree;
    cv::Mat desc;
    std::vector<uint32_t> ids;
    std::vector<Flag> flags;
    Se3Transform pose_f2g;
    std::vector<cv::KeyPoint> und_kpts;
    std::vector<cv::Point2f> kpts;
    std::vector<float> depth;
    cv::Mat image;
    std::shared_ptr<fbow::fBow> bowvector = std::make_shared<fbow::fBow>();
    std::shared_ptr<fbow::fBow2> bowvector_level = std::make_shared<fbow::fBow2>();
    uint32_t fseq_idx = std::numeric_limits<uint32_t>::max();
    vector<float> scaleFactors;
    ImageParams imageParams;
    DescriptorTypes::Type KpDescType = DescriptorTypes::DESC_NONE;
    cv::Point minXY = cv::Point2f(0, 0), maxXY = cv::Point2f(std::numeric_limits<float>::max(), std::numeric_limits<float>::max());
    void toStream(std::ostream& str) const;
    void fromStream(std::istream& str);
    Flag frame_flags;
};
class MapPoint {   // mappoint.h:111-131: the members MapPoint::toStream touches
public:
    uint32_t id = std::numeric_limits<uint32_t>::max();
    cv::Point3f pos3d, normal;
    std::map<uint32_t, uint32_t> frames;      // SafeMap<uint32_t,uint32_t> in the reference: a std::map behind a mutex
    uint64_t kfSinceAddition = 0;
    uint32_t lastFIdxSeen = std::numeric_limits<uint32_t>::max();
    Flag flags;
    cv::Mat _desc;
    uint16_t nTimesSeen = 0, nTimesVisible = 0;
    float mfMaxDistance = std::numeric_limits<float>::min(), mfMinDistance = std::numeric_limits<float>::max();
    void toStream(std::ostream& str) const;
    void fromStream(std::istream& str);
};
class Marker {     // marker.h:33-53: the members Marker::toStream touches
public:
    uint32_t id = 0;
    Se3Transform pose_g2m;
    float size = 0;
    std::set<uint32_t> frames;
    std::string dict_info;
    void toStream(std::ostream& str) const;
    void fromStream(std::istream& str);
};
#include "gen/frame_streams.inc"
#include "gen/mappoint_streams.inc"
#include "gen/marker_streams.inc"
}  // namespace ucoslam

#include "basictypes/reusablecontainer.h"   // the reference's own header (and expansiblecontainer.h), unchanged

extern "C" {
// builds a Frame from flat arrays, lets the reference write it.  markers: n_markers records of (id, ssize, 8 corner floats, 8
// undistorted corner floats, 16 doubles sols[0] as 4x4 CV_64F, errs[2], err_ratio) = 2 + 16 floats and 19 doubles each, dict "ARUCO_MIP_36h12".
// bow_level: n_nodes entries (node_id, ptr).  Returns the number of bytes (or -1: cap too small, -2: exception).
long ref_frame_to_stream(uint32_t idx, uint32_t fseq_idx, unsigned char frame_flags, int n_kp, const cv::KeyPoint* und_kpts, const unsigned char* desc,
                         const float* kpts, int n_depth, const float* depth, const uint32_t* ids, const unsigned char* flags, int n_markers,
                         const int32_t* marker_id, const float* marker_f, const double* marker_d, const float* pose16, int n_bow, const uint32_t* bow_word,
                         const float* bow_weight, int n_nodes, const uint32_t* node_id, const int32_t* node_ptr, const uint32_t* node_kp, int n_sf,
                         const float* sf, const float* K9, int n_dist, const float* dist, int cam_w, int cam_h, float bl, float depthscale, int img_rows,
                         int img_cols, const unsigned char* img, int build_tree, int min_x, int min_y, int max_x, int max_y, unsigned char* out, long cap) {
    try {
        ucoslam::Frame f;
        f.idx = idx; f.fseq_idx = fseq_idx; f.frame_flags.v = frame_flags; f.KpDescType = ucoslam::DescriptorTypes::DESC_ORB;
        f.und_kpts.assign(und_kpts, und_kpts + n_kp);
        if (n_kp) { f.desc.create(n_kp, 32, CV_8UC1); memcpy(f.desc.ptr<unsigned char>(0), desc, 32 * (size_t)n_kp); }
        f.kpts.resize(n_kp);
        for (int i = 0; i < n_kp; i++) f.kpts[i] = cv::Point2f(kpts[2 * i], kpts[2 * i + 1]);
        f.depth.assign(depth, depth + n_depth);
        f.ids.assign(ids, ids + n_kp);
        f.flags.resize(n_kp);
        for (int i = 0; i < n_kp; i++) f.flags[i].v = flags[i];
        for (int m = 0; m < n_markers; m++) {
            ucoslam::MarkerObservation mo;
            mo.id = marker_id[m]; mo.ssize = marker_f[17 * m]; mo.dict_info = "ARUCO_MIP_36h12";
            for (int c = 0; c < 4; c++) {
                mo.corners.push_back(cv::Point2f(marker_f[17 * m + 1 + 2 * c], marker_f[17 * m + 2 + 2 * c]));
                mo.und_corners.push_back(cv::Point2f(marker_f[17 * m + 9 + 2 * c], marker_f[17 * m + 10 + 2 * c]));
            }
            mo.poses.sols[0].create(4, 4, CV_64F);
            memcpy(mo.poses.sols[0].ptr<double>(0), marker_d + 19 * m, 128);
            mo.poses.errs[0] = marker_d[19 * m + 16]; mo.poses.errs[1] = marker_d[19 * m + 17]; mo.poses.err_ratio = marker_d[19 * m + 18];
            f.markers.push_back(mo);
        }
        memcpy(f.pose_f2g.ptr<float>(0), pose16, 64);
        for (int i = 0; i < n_bow; i++) { fbow::_float w; w.var = bow_weight[i]; (*f.bowvector)[bow_word[i]] = w; }
        for (int k = 0; k < n_nodes; k++) {
            std::vector<uint32_t>& v = (*f.bowvector_level)[node_id[k]];
            for (int e = node_ptr[k]; e < node_ptr[k + 1]; e++) v.push_back(node_kp[e]);
        }
        f.scaleFactors.assign(sf, sf + n_sf);
        f.imageParams.CameraMatrix.create(3, 3, CV_32F);
        memcpy(f.imageParams.CameraMatrix.ptr<float>(0), K9, 36);
        if (n_dist) { f.imageParams.Distorsion.create(1, n_dist, CV_32F); memcpy(f.imageParams.Distorsion.ptr<float>(0), dist, 4 * (size_t)n_dist); }
        f.imageParams.CamSize = cv::Size(cam_w, cam_h); f.imageParams.bl = bl; f.imageParams.rgb_depthscale = depthscale;
        if (img_rows * img_cols > 0) { f.image.create(img_rows, img_cols, CV_8UC1); memcpy(f.image.ptr<unsigned char>(0), img, (size_t)img_rows * img_cols); }
        if (build_tree && n_kp) f.keypoint_kdtree.build(f.und_kpts);     // Frame::create_kdtree, frame.h:124-127
        f.minXY = cv::Point(min_x, min_y); f.maxXY = cv::Point(max_x, max_y);
        std::stringstream ss;
        f.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -1;
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return -2; }
}
long ref_mappoint_to_stream(uint32_t id, const float* pos3, const unsigned char* desc32, int n_frames, const uint32_t* frame_kp, const float* normal3,
                            int seen, int visible, unsigned char flags, float maxd, float mind, unsigned long long kf_since, uint32_t last_seen,
                            unsigned char* out, long cap) {
    try {
        ucoslam::MapPoint p;
        p.id = id; p.pos3d = cv::Point3f(pos3[0], pos3[1], pos3[2]); p.normal = cv::Point3f(normal3[0], normal3[1], normal3[2]);
        if (desc32) { p._desc.create(1, 32, CV_8UC1); memcpy(p._desc.ptr<unsigned char>(0), desc32, 32); }
        for (int i = 0; i < n_frames; i++) p.frames[frame_kp[2 * i]] = frame_kp[2 * i + 1];
        p.nTimesSeen = (uint16_t)seen; p.nTimesVisible = (uint16_t)visible; p.flags.v = flags; p.mfMaxDistance = maxd; p.mfMinDistance = mind;
        p.kfSinceAddition = kf_since; p.lastFIdxSeen = last_seen;
        std::stringstream ss;
        p.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -1;
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return -2; }
}
long ref_mappoint_roundtrip(const unsigned char* in, long len, unsigned char* out, long cap) {
    try {
        std::stringstream is(std::string((const char*)in, (size_t)len));
        ucoslam::MapPoint p;
        p.fromStream(is);
        std::stringstream ss;
        p.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -1;
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return -2; }
}
// The map-point section of a map file (Map::toStream, map.cpp:316-325: map_points.toStream): the reference's OWN ReusableContainer /
// ExpansibleContainer headers, unchanged, over the MapPoint above.  The container is driven the way Map drives it: `n` points inserted in
// order (each read from its own stream by the reference's fromStream), the slots in `erase` erased in that order, then `n_again` more points
// inserted (they take the freed slots, last freed first); the reference writes the container.
long ref_mappoint_container(const unsigned char* streams, const long* lens, int n, const uint32_t* erase, int n_erase, int n_again, unsigned char* out, long cap) {
    try {
        ucoslam::ReusableContainer<ucoslam::MapPoint> c;
        const unsigned char* p = streams;
        auto next = [&](int i) {
            std::stringstream is(std::string((const char*)p, (size_t)lens[i]));
            p += lens[i];
            ucoslam::MapPoint mp;
            mp.fromStream(is);
            return mp;
        };
        for (int i = 0; i < n; i++) c.insert(next(i));
        for (int i = 0; i < n_erase; i++) c.erase(erase[i]);
        for (int i = 0; i < n_again; i++) c.insert(next(n + i));
        std::stringstream ss;
        c.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -1;
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return -2; }
}
// The keyframe section: FrameSet::toStream (frame.cpp:350-355) is `int magic = 88888` followed by ReusableContainer<Frame>::toStream; the Frame
// here carries the reference's own toStream / fromStream statements.  Same driving as above.
long ref_frame_container(const unsigned char* streams, const long* lens, int n, const uint32_t* erase, int n_erase, int n_again, unsigned char* out, long cap) {
    try {
        ucoslam::ReusableContainer<ucoslam::Frame> c;
        const unsigned char* p = streams;
        auto next = [&](int i) {
            std::stringstream is(std::string((const char*)p, (size_t)lens[i]));
            p += lens[i];
            ucoslam::Frame f;
            f.fromStream(is);
            return f;
        };
        for (int i = 0; i < n; i++) c.insert(next(i));
        for (int i = 0; i < n_erase; i++) c.erase(erase[i]);
        for (int i = 0; i < n_again; i++) c.insert(next(n + i));
        std::stringstream ss;
        int magic = 88888;
        ss.write((char*)&magic, sizeof(magic));
        c.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -1;
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return -2; }
}
// The marker section of a map file: Map::toStream writes toStream__kv_complex(map_markers, str) (map.cpp:321; io_utils.h:113-121, the
// reference's own template) over std::map<uint32_t, Marker> (SafeMap is a std::map) with the reference's Marker::toStream statements.
// markers: n records of (id, size, 16 pose floats), frames: n_frames[i] keyframe ids each, dict "ARUCO_MIP_36h12"
long ref_marker_map_to_stream(int n, const uint32_t* ids, const float* size, const float* pose16, const int* n_frames, const uint32_t* frames, unsigned char* out, long cap) {
    try {
        std::map<uint32_t, ucoslam::Marker> mm;
        for (int i = 0; i < n; i++) {
            ucoslam::Marker m;
            m.id = ids[i]; m.size = size[i]; m.dict_info = "ARUCO_MIP_36h12";
            memcpy(m.pose_g2m.ptr<float>(0), pose16 + 16 * i, 64);
            for (int k = 0; k < n_frames[i]; k++) m.frames.insert(*frames++);
            mm[ids[i]] = m;
        }
        std::stringstream ss;
        ucoslam::toStream__kv_complex(mm, ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -1;
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return -2; }
}
// the reference reads a container stream and writes it again; *n_valid = its size() (valid elements)
long ref_mappoint_container_roundtrip(const unsigned char* in, long len, unsigned char* out, long cap, long* n_valid) {
    try {
        std::stringstream is(std::string((const char*)in, (size_t)len));
        ucoslam::ReusableContainer<ucoslam::MapPoint> c;
        c.fromStream(is);
        if (n_valid) *n_valid = (long)c.size();
        std::stringstream ss;
        c.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -1;
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return -2; }
}
// the reference reads a stream (fromStream) and writes it again (toStream): -2 when its reader throws
long ref_frame_roundtrip(const unsigned char* in, long len, unsigned char* out, long cap) {
    try {
        std::stringstream is(std::string((const char*)in, (size_t)len));
        ucoslam::Frame f;
        f.fromStream(is);
        std::stringstream ss;
        f.toStream(ss);
        const std::string b = ss.str();
        if ((long)b.size() > cap) return -1;
        memcpy(out, b.data(), b.size());
        return (long)b.size();
    } catch (std::exception&) { return -2; }
}
}
