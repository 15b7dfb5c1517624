// new_points_b200.h — the mapper's new-map-point creation on the device: a drop-in body for MapManager's createNewPoints(frame, nn,
// maxPoints) (src/utils/mapmanager.cpp:9772-10788; macro-obfuscated, `cpp -P` de-obfuscates it): the FrameMatcher::setParams + the
// OpenMP loop over the neighbour keyframes (matchEpipolar -> Triangulate -> global frame -> scale consistency) + the merge into
// NewPointInfo records become ONE call of uco_b200_new_points (one upload, three launches, one download).  The neighbour list
// (covisibility graph query, :9798) and the fundamental matrices (the reference's own computeF12, misc.cpp:893-920, through
// FrameMatcher_impl::getFund12's arguments) stay on the caller's side.
// MapManager::NewPointInfo is private to MapManager: paste this function into mapmanager.cpp (or befriend it) and call it from
// createNewPoints instead of the loop.  Needs the reference's headers and OpenCV C++ (see the note in orb_extractor_b200.h).
#pragma once
#include <limits>
#include <vector>
#include "uco_b200_cxx.h"

namespace ucoslam {

cv::Mat computeF12(const cv::Mat& RT1, const cv::Mat& CameraMatrix1, const cv::Mat& RT2, const cv::Mat& CameraMatrix2);   // basictypes/misc.h

// NewPointInfoT: MapManager::NewPointInfo { cv::Point3d pose; bool isStereo; std::vector<std::pair<uint32_t,uint32_t>> frame_kpt; float dist; }
template <class NewPointInfoT, class MapT>
std::vector<NewPointInfoT> createNewPoints_b200(uco_b200::Context& ctx, MapT& map, Frame& frame, const std::vector<uint32_t>& neighbours, uint32_t maxPoints,
                                               float maxDescDistance, float scaleFactor) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(uco_keypoint) && sizeof(cv::DMatch) == sizeof(uco_match), "layouts");
    if (frame.ids.size() == 0 || neighbours.empty()) return {};
    // FrameMatcher::manageMode(MODE_UNASSIGNED), framematcher.cpp:174-196: keypoints without a map point that are not FLAG_NONMAXIMA
    auto rows = [](const Frame& f, std::vector<int32_t>& map_, std::vector<uint8_t>& desc) {
        map_.clear();
        for (size_t i = 0; i < f.ids.size(); i++)
            if (f.ids[i] == std::numeric_limits<uint32_t>::max() && !f.flags[i].is(Frame::FLAG_NONMAXIMA)) map_.push_back((int32_t)i);
        desc.resize(32 * map_.size());
        for (size_t r = 0; r < map_.size(); r++) memcpy(&desc[32 * r], f.desc.template ptr<uchar>(map_[r]), 32);
    };
    const size_t F = neighbours.size();
    std::vector<int32_t> tRows;
    std::vector<uint8_t> tDesc;
    rows(frame, tRows, tDesc);
    std::vector<std::vector<int32_t>> qRows(F);
    std::vector<std::vector<uint8_t>> qDesc(F);
    std::vector<const uint8_t*> qd(F);
    std::vector<const uco_keypoint*> qk(F);
    std::vector<const int32_t*> qm(F);
    std::vector<int32_t> nq(F), nk(F);
    std::vector<float> f12(9 * F), rt(16 * F), Knb(4 * F);
    const Se3Transform g2f = frame.pose_f2g.inv();
    size_t obsCap = 0;
    for (size_t f = 0; f < F; f++) {
        Frame& nb = map.keyframes[neighbours[f]];
        rows(nb, qRows[f], qDesc[f]);
        qd[f] = qDesc[f].data(); qm[f] = qRows[f].data(); nq[f] = (int32_t)qRows[f].size();
        qk[f] = reinterpret_cast<const uco_keypoint*>(nb.und_kpts.data()); nk[f] = (int32_t)nb.und_kpts.size();
        obsCap += qRows[f].size();
        cv::Mat T = nb.pose_f2g * g2f;                               // :9994, train (the new keyframe) -> query (the neighbour)
        cv::Mat T32, F12;
        T.convertTo(T32, CV_32F);
        F12 = computeF12(cv::Mat::eye(4, 4, CV_32F), frame.imageParams.CameraMatrix, T32, nb.imageParams.CameraMatrix);   // getFund12, framematcher.cpp:58-64
        memcpy(&f12[9 * f], F12.ptr<float>(0), 36);
        memcpy(&rt[16 * f], T32.ptr<float>(0), 64);
        const cv::Mat& K = nb.imageParams.CameraMatrix;
        Knb[4 * f] = K.at<float>(0, 0); Knb[4 * f + 1] = K.at<float>(1, 1); Knb[4 * f + 2] = K.at<float>(0, 2); Knb[4 * f + 3] = K.at<float>(1, 2);
    }
    uco_new_points_params prm{};
    prm.match.min_desc_dist = maxDescDistance * 2; prm.match.nn_match_ratio = 0.6f; prm.match.check_orientation = 1;   // :9982
    prm.match.max_octave_diff = std::numeric_limits<int>::max(); prm.match.use_f12 = 1;
    const Frame& nb0 = map.keyframes[neighbours[0]];
    prm.match.n_scales = (int)nb0.scaleFactors.size();
    for (int i = 0; i < prm.match.n_scales && i < UCO_MATCH_MAX_SCALES; i++) prm.match.scale_factors[i] = nb0.scaleFactors[i];
    const cv::Mat& K = frame.imageParams.CameraMatrix;
    prm.K_kf[0] = K.at<float>(0, 0); prm.K_kf[1] = K.at<float>(1, 1); prm.K_kf[2] = K.at<float>(0, 2); prm.K_kf[3] = K.at<float>(1, 2);
    cv::Mat G = g2f;
    cv::Mat G32;
    G.convertTo(G32, CV_32F);
    memcpy(prm.g2f_kf, G32.ptr<float>(0), 64);
    prm.n_levels_kf = (int)frame.scaleFactors.size(); prm.scale_factors_kf = frame.scaleFactors.data();
    prm.n_levels_nb = (int)nb0.scaleFactors.size(); prm.scale_factors_nb = nb0.scaleFactors.data();
    prm.max_chi2 = 5.998f;                                             // Triangulate's default, misc.h:65
    prm.scale_ratio_factor = 1.5f * scaleFactor;                       // :10018
    prm.max_points = (int32_t)maxPoints;
    const int capP = (int)tRows.size();
    std::vector<int32_t> kpt(capP + 1), optr(capP + 2), ofr(obsCap + 1), okp(obsCap + 1);
    std::vector<float> xyz(3 * (size_t)capP + 3), dist(capP + 1);
    int32_t n = 0;
    ctx.check(uco_b200_new_points(ctx.get(), tDesc.data(), (int)tRows.size(), 32, reinterpret_cast<const uco_keypoint*>(frame.und_kpts.data()),
                                  (int)frame.und_kpts.size(), tRows.data(), (int)F, qd.data(), nq.data(), 32, qk.data(), nk.data(), qm.data(), f12.data(),
                                  rt.data(), Knb.data(), &prm, &n, kpt.data(), xyz.data(), dist.data(), optr.data(), ofr.data(), okp.data(), capP, (int)obsCap,
                                  nullptr, nullptr, nullptr));
    std::vector<NewPointInfoT> out(n);
    for (int j = 0; j < n; j++) {
        out[j].pose = cv::Point3d(xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]);   // the reference assigns its cv::Point3f to the cv::Point3d field
        out[j].dist = dist[j];
        out[j].frame_kpt.push_back({frame.idx, (uint32_t)kpt[j]});
        for (int o = optr[j]; o < optr[j + 1]; o++) out[j].frame_kpt.push_back({map.keyframes[neighbours[ofr[o]]].idx, (uint32_t)okp[o]});
    }
    return out;
}

}  // namespace ucoslam
