// projection_matcher_b200.h — the reference's projection matcher seam.  Map::matchFrameToMapPoints is a private member with no
// virtual interface (src/map.h:169; friends System / MapManager / LoopDetector), so the replacement is a drop-in BODY: the
// reference keeps gathering the candidate list (src/map.cpp:655-672: getMapPointsInFrames + the lastFIdxSeen filter, container
// walking) and hands it to ucoslam::matchFrameToMapPoints_b200, which flattens it, calls uco_b200_match_projected and applies
// MapPoint::setVisible() where the reference does (:711).  The frame's kd-tree travels as the bytes KdTreeIndex::toStream writes
// (picoflann.h:603-660), so the device walks exactly the tree the reference built in Frame (frame.h:125).
// Compiled and driven next to the reference's own statements by tests/adapters/adapter_world_test.cpp (oracle/shim2 stand-ins).
#pragma once
#include <sstream>
#include <vector>
#include "map.h"
#include "uco_b200_cxx.h"

namespace ucoslam {

// what follows `if (smap_ids.size()==0) return {};` in Map::matchFrameToMapPoints (src/map.cpp:672-770)
inline std::vector<cv::DMatch> matchFrameToMapPoints_b200(uco_b200::Context& ctx, Map& map, const std::vector<uint32_t>& smap_ids,
                                                         Frame& curframe, const cv::Mat& pose_f2g_, float minDescDist,
                                                         float maxRepjDist, bool markMapPointsAsVisible) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(uco_keypoint) && sizeof(cv::DMatch) == sizeof(uco_match), "layouts");
    const int m = (int)smap_ids.size();
    std::vector<float> pos(3 * m), nrm(3 * m), dmin(m), dmax(m);
    std::vector<uint8_t> desc(32 * (size_t)m);
    cv::Mat d;
    for (int i = 0; i < m; i++) {
        MapPoint& mp = map.map_points[smap_ids[i]];
        const cv::Point3f p = mp.getCoordinates(), n = mp.getNormal();
        pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
        nrm[3 * i] = n.x; nrm[3 * i + 1] = n.y; nrm[3 * i + 2] = n.z;
        dmin[i] = mp.getMinDistanceInvariance();
        dmax[i] = mp.getMaxDistanceInvariance();
        mp.getDescriptor(d);
        if (d.type() != CV_8UC1 || d.cols != 32) throw std::runtime_error("matchFrameToMapPoints_b200: 256-bit binary descriptors only");
        memcpy(&desc[32 * (size_t)i], d.ptr<uchar>(0), 32);
    }
    // Frame::keypoint_kdtree -> flattened nodes (cache this per frame if the matcher runs more than once on it)
    std::stringstream ss;
    curframe.keypoint_kdtree.toStream(ss);
    const std::string bytes = ss.str();
    const int nk = (int)curframe.und_kpts.size();
    std::vector<uco_kdnode> nodes(2 * (size_t)nk + 2);
    std::vector<int32_t> leaf(nk + 1);
    uco_frame_view fr{};
    int n_leaf = 0;
    if (uco_b200_kdtree_parse(bytes.data(), bytes.size(), nodes.data(), (int)nodes.size(), leaf.data(), (int)leaf.size(), fr.bbox,
                              &fr.n_nodes, &n_leaf) != UCO_OK)
        throw std::runtime_error("matchFrameToMapPoints_b200: cannot read the frame's kd-tree");
    fr.n_kp = nk;
    fr.kps = reinterpret_cast<const uco_keypoint*>(curframe.und_kpts.data());
    fr.desc = curframe.desc.ptr<uchar>(0);
    fr.desc_stride = curframe.desc.step[0];
    fr.nodes = nodes.data();
    fr.leaf_idx = leaf.data();
    fr.n_levels = (int)curframe.scaleFactors.size();
    fr.scale_factors = curframe.scaleFactors.data();
    fr.fx = curframe.imageParams.CameraMatrix.at<float>(0, 0); fr.fy = curframe.imageParams.CameraMatrix.at<float>(1, 1);
    fr.cx = curframe.imageParams.CameraMatrix.at<float>(0, 2); fr.cy = curframe.imageParams.CameraMatrix.at<float>(1, 2);
    fr.min_xy[0] = curframe.minXY.x; fr.min_xy[1] = curframe.minXY.y; fr.max_xy[0] = curframe.maxXY.x; fr.max_xy[1] = curframe.maxXY.y;
    uco_mappoints mp{m, smap_ids.data(), pos.data(), nrm.data(), dmin.data(), dmax.data(), desc.data()};
    Se3Transform pose; pose = pose_f2g_;
    std::vector<cv::DMatch> out(m);
    std::vector<uint8_t> visible(m);
    int n = 0;
    ctx.check(uco_b200_match_projected(ctx.get(), &mp, &fr, pose.ptr<float>(0), minDescDist, maxRepjDist,
                                       reinterpret_cast<uco_match*>(out.data()), &n, visible.data()));
    out.resize(n);
    if (markMapPointsAsVisible)
        for (int i = 0; i < m; i++)
            if (visible[i]) map.map_points[smap_ids[i]].setVisible();
    return out;
}

}  // namespace ucoslam
