// pnp_solver_b200.h — link-time replacement of the static PnPSolver::solvePnp (src/optimization/pnpsolver.h:33,
// src/optimization/pnpsolver.cpp:116-408): the same signature and effects (estimatedPose updated, map_matches[i].imgIdx = -1 / 1,
// return value = matches kept), the four-round Levenberg-Marquardt schedule runs in uco_b200_pose_only.  The matches are
// flattened exactly as solvePnp walks them (:200-259): und_kpts[queryIdx], map point coordinates, stability weight, stereo
// observations from Frame::getDepth, markers with a valid map pose seen from the neighbourhood of currentKeyFrame (:262-281).
// The tracker calls it 2-3 times per frame from one thread (src/utils/system.cpp); use one context per calling thread.
// Compiled and driven next to the reference's own code by tests/adapters/adapter_world_test.cpp (oracle/shim2 stand-ins).
#pragma once
#include <vector>
#include "map.h"
#include "basictypes/se3.h"
#include "uco_b200_cxx.h"

namespace ucoslam {

inline int solvePnp_b200(uco_b200::Context& ctx, const Frame& frame, std::shared_ptr<Map> TheMap, std::vector<cv::DMatch>& map_matches,
                         se3& estimatedPose, int64_t currentKeyFrame) {
    if (map_matches.empty() && frame.markers.empty()) return 0;                          // :144
    const size_t n = map_matches.size();
    std::vector<float> p3(3 * n), uv(2 * n), ur(n, 0.f), inv(n);
    std::vector<uint8_t> st(n, 0), stable(n, 1), bad(n, 0);
    const float mbf = frame.imageParams.bl * frame.imageParams.fx();
    for (size_t i = 0; i < n; i++) {
        const cv::KeyPoint& kpt = frame.und_kpts[map_matches[i].queryIdx];
        const MapPoint& mp = TheMap->map_points[map_matches[i].trainIdx];
        const cv::Point3f p = mp.getCoordinates();
        p3[3 * i] = p.x; p3[3 * i + 1] = p.y; p3[3 * i + 2] = p.z;
        uv[2 * i] = kpt.pt.x; uv[2 * i + 1] = kpt.pt.y;
        inv[i] = 1.f / frame.scaleFactors[kpt.octave];
        stable[i] = mp.isStable();
        const float depth = frame.getDepth(map_matches[i].queryIdx);
        if (depth > 0) { st[i] = 1; ur[i] = kpt.pt.x - mbf / depth; }                    // :226
    }
    std::vector<float> mpose, msize, mcorners;
    if (!frame.markers.empty() && currentKeyFrame != -1) {                               // :262-281
        auto neigh = TheMap->getNeighborKeyFrames(currentKeyFrame, true);
        for (const auto& mo : frame.markers) {
            auto it = TheMap->map_markers.find(mo.id);
            if (it == TheMap->map_markers.end() || !it->second.pose_g2m.isValid()) continue;
            bool seen = false;
            for (auto f : it->second.frames) if (neigh.count(f)) { seen = true; break; }
            if (!seen) continue;
            const float* m = it->second.pose_g2m.ptr<float>(0);
            mpose.insert(mpose.end(), m, m + 16);
            msize.push_back(it->second.size);
            for (const auto& c : mo.und_corners) { mcorners.push_back(c.x); mcorners.push_back(c.y); }
        }
    }
    cv::Mat pose_io = estimatedPose.convert();
    uco_pnp_problem pb{};
    pb.n_matches = (int)n; pb.pose44 = pose_io.ptr<float>(0); pb.points3 = p3.data(); pb.obs_uv = uv.data(); pb.obs_ur = ur.data();
    pb.obs_stereo = st.data(); pb.obs_inv_sigma2 = inv.data(); pb.stable = stable.data();
    pb.fx = frame.imageParams.fx(); pb.fy = frame.imageParams.fy(); pb.cx = frame.imageParams.cx(); pb.cy = frame.imageParams.cy(); pb.bf = mbf;
    pb.n_markers = (int)msize.size(); pb.marker_pose44 = mpose.data(); pb.marker_size = msize.data(); pb.marker_corners = mcorners.data();
    uco_pnp_result res{};
    res.bad = bad.data();
    ctx.check(uco_b200_pose_only(ctx.get(), &pb, &res));
    for (size_t i = 0; i < n; i++) map_matches[i].imgIdx = bad[i] ? -1 : 1;             // :398-404
    cv::Mat out(4, 4, CV_32F);
    memcpy(out.ptr<float>(0), res.pose44, 64);
    estimatedPose = out;
    return res.n_good;
}

// PnPSolver::solvePnPRansac (src/optimization/pnpsolver.h, pnpsolver.cpp:36-114): same signature and effects (false -> arguments
// untouched; true -> posef2g_io = winning hypothesis, matches_io = its inliers).  All maxIters hypotheses run side by side in
// uco_b200_pnp_ransac; the 4-match samples come from its counter-based generator seeded per call (the reference consumes the
// process-wide rand() stream through std::random_shuffle, which a parallel sampler cannot replay).
inline bool solvePnPRansac_b200(uco_b200::Context& ctx, const Frame& frame, std::shared_ptr<Map> map, std::vector<cv::DMatch>& matches_io,
                                se3& posef2g_io, int maxIters, uint64_t seed = 0) {
    if (matches_io.size() < 4) return false;                                             // :39
    const int n = (int)matches_io.size();
    std::vector<float> p3(3 * (size_t)n), p2(2 * (size_t)n), nr(3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        const MapPoint& mp = map->map_points[matches_io[i].trainIdx];
        const cv::Point3f p = mp.getCoordinates(), nn = mp.getNormal();
        const cv::Point2f k = frame.und_kpts[matches_io[i].queryIdx].pt;
        p3[3 * i] = p.x; p3[3 * i + 1] = p.y; p3[3 * i + 2] = p.z;
        nr[3 * i] = nn.x; nr[3 * i + 1] = nn.y; nr[3 * i + 2] = nn.z;
        p2[2 * i] = k.x; p2[2 * i + 1] = k.y;
    }
    const float* cam = frame.imageParams.CameraMatrix.ptr<float>(0);
    const float K[4] = {cam[0], cam[4], cam[2], cam[5]};
    cv::Mat pose(4, 4, CV_32F);
    std::vector<int32_t> inl(n);
    int ni = 0;
    ctx.check(uco_b200_pnp_ransac(ctx.get(), p3.data(), p2.data(), nr.data(), n, K, maxIters, nullptr, seed, pose.ptr<float>(0), inl.data(),
                                  &ni, nullptr, nullptr));
    if (ni < 4) return false;                                                            // :103
    std::vector<cv::DMatch> kept;
    kept.reserve(ni);
    for (int i = 0; i < ni; i++) kept.push_back(matches_io[inl[i]]);
    matches_io = kept;
    posef2g_io = pose;
    return true;
}

}  // namespace ucoslam
