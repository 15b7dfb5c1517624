// triangulate_b200.h — the reference's triangulation seam: the free function ucoslam::Triangulate(Train, Query, RT_Q2T, matches,
// maxChi2) (src/basictypes/misc.h:65, misc.cpp:921-1040), called by the mapper's new-point creation (src/utils/mapmanager.cpp:10093)
// and the map initialiser (src/utils/mapinitializer.cpp:1574).  Same signature plus the context; same result convention (one
// cv::Point3f per match, NaN where rejected).
// Compile inside the reference tree (needs its headers and OpenCV C++; see the note in orb_extractor_b200.h).
#pragma once
#include <vector>
#include "map_types/frame.h"
#include "uco_b200_cxx.h"

namespace ucoslam {

inline std::vector<cv::Point3f> Triangulate_b200(uco_b200::Context& ctx, const Frame& Train, const Frame& Query, const cv::Mat& RT_Q2T,
                                                 const std::vector<cv::DMatch>& matches, float maxChi2 = 5.998f) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(uco_keypoint) && sizeof(cv::DMatch) == sizeof(uco_match) &&
                      sizeof(cv::Point3f) == 3 * sizeof(float), "layouts");
    std::vector<cv::Point3f> out(matches.size());
    if (matches.empty()) return out;
    uco_triangulate_params p{};
    const cv::Mat &K1 = Train.imageParams.CameraMatrix, &K2 = Query.imageParams.CameraMatrix;
    p.K_train[0] = K1.at<float>(0, 0); p.K_train[1] = K1.at<float>(1, 1); p.K_train[2] = K1.at<float>(0, 2); p.K_train[3] = K1.at<float>(1, 2);
    p.K_query[0] = K2.at<float>(0, 0); p.K_query[1] = K2.at<float>(1, 1); p.K_query[2] = K2.at<float>(0, 2); p.K_query[3] = K2.at<float>(1, 2);
    cv::Mat RT;
    RT_Q2T.convertTo(RT, CV_32F);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) p.RT[4 * r + c] = RT.at<float>(r, c);
    p.n_levels_train = (int)Train.scaleFactors.size(); p.scale_factors_train = Train.scaleFactors.data();
    p.n_levels_query = (int)Query.scaleFactors.size(); p.scale_factors_query = Query.scaleFactors.data();
    p.max_chi2 = maxChi2;   // scale_ratio_factor / to_global stay 0: plain ucoslam::Triangulate
    ctx.check(uco_b200_triangulate(ctx.get(), reinterpret_cast<const uco_keypoint*>(Train.und_kpts.data()), (int)Train.und_kpts.size(),
                                   reinterpret_cast<const uco_keypoint*>(Query.und_kpts.data()), (int)Query.und_kpts.size(),
                                   reinterpret_cast<const uco_match*>(matches.data()), (int)matches.size(), &p,
                                   reinterpret_cast<float*>(out.data()), nullptr));
    return out;
}

}  // namespace ucoslam
