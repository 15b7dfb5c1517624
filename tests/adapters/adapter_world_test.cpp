// adapter_world_test.cpp — TEST INFRASTRUCTURE.  Compiles and DRIVES the five C++ adapters that sit behind the reference's plugin seams
// (ucoslam-cv3_b200/host/{orb_extractor,frame_matcher,projection_matcher,pnp_solver,global_optimizer}_b200.h) next to the reference's
// own code, in one process on the GPU box:
//   1. ORBextractorB200 through the reference's real Feature2DSerializable (its header + the member definitions of
//      feature2dserializable.cpp:27-31,74-83) against the C ABI called directly: identical keypoints / descriptors, stream layout;
//   2. FrameMatcher_B200 (declared against the reference's own _impl::FrameMatcher_impl, framematcher.cpp:31-58) against the
//      reference's FrameMatcher_Flann compiled from the same file (exact index, see oracle/ref_match_wrap.cpp): identical cv::DMatch lists;
//   3. matchFrameToMapPoints_b200 against the reference's own Map::matchFrameToMapPoints statements (map.cpp:651-770): identical lists
//      and setVisible() marks, the kd-tree travelling as the bytes the reference's picoflann writes;
//   4. solvePnp_b200 against the reference's g2o + typesg2o.h (oracle/_ref/libref_g2o.so): pose to 1e-6, identical inlier flags;
//   5. GlobalOptimizerB200 (derived from the reference's real globaloptimizer.h) against the same g2o on the window it flattened, with one
//      camera and with keyframes taken with two cameras;
//   6. createNewPoints_b200 (the body of the mapper's new-map-point creation) on three keyframes: every point re-projects onto its keypoints.
// OpenCV / Frame / Map are the container stand-ins of oracle/shim2 (the image has no OpenCV C++).  Built by `make -C oracle ref`
// into oracle/_ref/, run by tests/test_adapters_gpu.py.  Exit code 0 = every comparison held.
#include <cstdio>
#include <random>
#include <sstream>
#include <xflann/xflann.h>
#include <fbow/fbow.h>
#define HKMeansParams(a, b) LinearParams()
#include <utils/framematcher.cpp>          // the reference's matcher + _impl::FrameMatcher_impl, exact index (oracle/ref_match_wrap.cpp)
#undef HKMeansParams
#include "frame_matcher_b200.h"
#include <featureextractors/feature2dserializable.h>
#include "orb_extractor_b200.h"
#include "map.h"
#include "projection_matcher_b200.h"
#include "pnp_solver_b200.h"
#include <optimization/globaloptimizer.h>
#include "global_optimizer_b200.h"
#include "new_points_b200.h"

namespace ucoslam {
#include "gen/misc_filters.inc"
#include "gen/f2d_members.inc"
#include "gen/frame_region.inc"
#include "gen/map_match.inc"
// computeF12 (misc.cpp:893-920) is OpenCV matrix algebra: off (empty matrix = no epipolar gate, as in oracle/ref_match_wrap.cpp) for the
// matcher comparison; for the new-point adapter a plain restatement: F12 = K1^-T [t12]x R12 K2^-1, R12 = R1 R2^T, t12 = -R1 R2^T t2 + t1
static bool g_real_f12 = false;
cv::Mat computeF12(const cv::Mat& RT1, const cv::Mat& K1, const cv::Mat& RT2, const cv::Mat& K2_) {
    if (!g_real_f12) return cv::Mat();
    const cv::Mat& K2 = K2_.empty() ? K1 : K2_;
    double R12[3][3], t12[3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            R12[r][c] = 0;
            for (int k = 0; k < 3; k++) R12[r][c] += (double)RT1.at<float>(r, k) * RT2.at<float>(c, k);
        }
    for (int r = 0; r < 3; r++) {
        t12[r] = RT1.at<float>(r, 3);
        for (int k = 0; k < 3; k++) t12[r] -= R12[r][k] * RT2.at<float>(k, 3);
    }
    const double tx[3][3] = {{0, -t12[2], t12[1]}, {t12[2], 0, -t12[0]}, {-t12[1], t12[0], 0}};
    auto kinv = [](const cv::Mat& K, double o[3][3]) {   // inverse of [fx 0 cx; 0 fy cy; 0 0 1]
        const double fx = K.at<float>(0, 0), fy = K.at<float>(1, 1), cx = K.at<float>(0, 2), cy = K.at<float>(1, 2);
        const double v[3][3] = {{1 / fx, 0, -cx / fx}, {0, 1 / fy, -cy / fy}, {0, 0, 1}};
        memcpy(o, v, sizeof v);
    };
    double k1i[3][3], k2i[3][3], A[3][3], B[3][3], C[3][3];
    kinv(K1, k1i); kinv(K2, k2i);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { A[r][c] = 0; for (int k = 0; k < 3; k++) A[r][c] += k1i[k][r] * tx[k][c]; }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { B[r][c] = 0; for (int k = 0; k < 3; k++) B[r][c] += A[r][k] * R12[k][c]; }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { C[r][c] = 0; for (int k = 0; k < 3; k++) C[r][c] += B[r][k] * k2i[k][c]; }
    cv::Mat F(3, 3, CV_32F);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F.at<float>(r, c) = (float)C[r][c];
    return F;
}
}  // namespace ucoslam

extern "C" {
int ref_pose_only(const float* pose44, int n, const float* points3, const float* obs_uv, const float* obs_ur, const uint8_t* obs_stereo,
                  const float* obs_inv_sigma2, const uint8_t* stable, float fx, float fy, float cx, float cy, float bf, int n_markers,
                  const float* marker_pose44, const float* marker_size, const float* marker_corners, float* out_pose44, double* out_pose7,
                  uint8_t* bad, int* iters_done);
int ref_ba_optimize(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs, const int32_t* obs_pose,
                    const int32_t* obs_point, const float* obs_uv, const float* obs_ur, const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx,
                    float fy, float cx, float cy, float bf, int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                    uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace);
int ref_ba_optimize_cams(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs, const int32_t* obs_pose,
                         const int32_t* obs_point, const float* obs_uv, const float* obs_ur, const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx,
                         float fy, float cx, float cy, float bf, int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                         uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace, int n_markers, const float* marker_pose44,
                         const float* marker_size, int n_mobs, const int32_t* mobs_marker, const int32_t* mobs_pose, const float* mobs_corners,
                         const float* mobs_weight, double* out_marker_pose7, float* out_marker_pose44, double* out_mobs_chi2, const float* pose_cam);
int ref_ba_optimize_planar(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs, const int32_t* obs_pose,
                           const int32_t* obs_point, const float* obs_uv, const float* obs_ur, const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx,
                           float fy, float cx, float cy, float bf, int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                           uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace, int n_markers, const float* marker_pose44,
                           const float* marker_size, int n_mobs, const int32_t* mobs_marker, const int32_t* mobs_pose, const float* mobs_corners,
                           const float* mobs_weight, double* out_marker_pose7, float* out_marker_pose44, double* out_mobs_chi2, const float* pose_cam,
                           int n_plane, int plane_ref, const float* plane_ref_pose44, const int32_t* plane_other, double plane_weight);
}

static int fails = 0;
#define EXPECT(c, msg) do { if (!(c)) { std::printf("FAIL %s\n", msg); fails++; } else std::printf("ok   %s\n", msg); } while (0)

using namespace ucoslam;

static bool same(const std::vector<cv::DMatch>& a, const std::vector<cv::DMatch>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++)
        if (a[i].queryIdx != b[i].queryIdx || a[i].trainIdx != b[i].trainIdx || a[i].imgIdx != b[i].imgIdx || a[i].distance != b[i].distance) return false;
    return true;
}

int main() {
    // ---- a textured scene: block noise; three 640x480 views = crops of one texture = a camera sliding over a fronto-parallel plane
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int TW = 800, TH = 600, W = 640, H = 480;
    std::mt19937 rng(11);
    std::vector<uint8_t> tex(TW * TH);
    for (int by = 0; by < TH / 8 + 1; by++)
        for (int bx = 0; bx < TW / 8 + 1; bx++) {
            const int v = rng() % 256;
            for (int y = by * 8; y < std::min(TH, by * 8 + 8); y++)
                for (int x = bx * 8; x < std::min(TW, bx * 8 + 8); x++) tex[y * TW + x] = (uint8_t)std::min(255, std::max(0, v + (int)(rng() % 13) - 6));
        }
    const int off[3][2] = {{80, 60}, {88, 64}, {72, 58}};
    const float f = 500.f, cx = 319.5f, cy = 239.5f, Z = 2.f;
    cv::Mat img[3];
    for (int k = 0; k < 3; k++) {
        img[k] = cv::Mat(H, W, CV_8UC1);
        for (int y = 0; y < H; y++) memcpy(img[k].ptr<uchar>(y), &tex[(y + off[k][1]) * TW + off[k][0]], W);
    }

    // ---- 1. the extractor seam ------------------------------------------------------------------------------------------------------
    std::shared_ptr<Feature2DSerializable> fd = std::make_shared<ORBextractorB200>();
    Feature2DSerializable::FeatParams fp(1500, 8, 1.2f, 1);
    Frame F[3];
    uco_b200::Context raw;
    bool ext_ok = true;
    for (int k = 0; k < 3; k++) {
        fd->detectAndCompute(img[k], cv::Mat(), F[k].und_kpts, F[k].desc, fp);           // the reference's own non-virtual entry point
        std::vector<uco_keypoint> kp(1500);
        std::vector<uint8_t> dsc(1500 * 32);
        uco_orb_params prm;
        uco_b200_orb_default_params(&prm);
        prm.max_features = 1500;
        int n = 0;
        raw.check(uco_b200_orb_extract(raw.get(), img[k].data, W, H, img[k].step[0], &prm, kp.data(), dsc.data(), 1500, &n));
        ext_ok = ext_ok && n == (int)F[k].und_kpts.size() && n > 800 && F[k].desc.rows == n && F[k].desc.cols == 32 &&
                 !memcmp(kp.data(), F[k].und_kpts.data(), sizeof(uco_keypoint) * n) && !memcmp(dsc.data(), F[k].desc.ptr<uchar>(0), 32 * (size_t)n);
        F[k].idx = k; F[k].fseq_idx = 100 + k;
        F[k].ids.assign(F[k].und_kpts.size(), std::numeric_limits<uint32_t>::max());
        F[k].flags.assign(F[k].und_kpts.size(), Flag());
        F[k].scaleFactors = fp.getScaleFactors();
        F[k].imageParams.CameraMatrix = cv::Mat::eye(3, 3, CV_32F);
        F[k].imageParams.CameraMatrix.at<float>(0, 0) = f; F[k].imageParams.CameraMatrix.at<float>(1, 1) = f;
        F[k].imageParams.CameraMatrix.at<float>(0, 2) = cx; F[k].imageParams.CameraMatrix.at<float>(1, 2) = cy;
        F[k].imageParams.CamSize = cv::Size(W, H);
        F[k].minXY = cv::Point2f(0, 0); F[k].maxXY = cv::Point2f(W, H);
        F[k].bowvector_level = std::make_shared<fbow::fBow2>();
        F[k].create_kdtree();
        // camera pose: frame k shows the texture shifted by (off[k] - off[0]) pixels = a translation parallel to the plane z = Z
        const float tx = -(off[k][0] - off[0][0]) * Z / f, ty = -(off[k][1] - off[0][1]) * Z / f;
        F[k].pose_f2g[3] = tx; F[k].pose_f2g[7] = ty;
    }
    EXPECT(ext_ok, "ORBextractorB200 through Feature2DSerializable::detectAndCompute == uco_b200_orb_extract (keypoints + descriptors, 3 frames)");
    {
        std::stringstream ss;
        fd->toStream(ss);                                                               // the reference's own toStream
        const std::string b = ss.str();
        uint64_t sig = 0, type = 99;
        memcpy(&sig, b.data(), 8); memcpy(&type, b.data() + 8, 8);
        EXPECT(sig == 1828374733ull && type == 0 /*F2D_ORB*/ && b.size() == 8 + 8 + 8 + sizeof(Feature2DSerializable::FeatParams) &&
                   fd->getMinDescDistance() == 50 && fd->getDescriptorType() == DescriptorTypes::DESC_ORB,
               "extractor stream layout (signature, F2D_ORB, raw FeatParams) and descriptor metadata");
    }

    // ---- 2. the matcher seam --------------------------------------------------------------------------------------------------------
    {
        F[0].ids[3] = 7; F[0].ids[10] = 8; F[1].ids[5] = 9;                             // a few assigned keypoints: the modes differ
        F[1].flags[20].set(Frame::FLAG_NONMAXIMA, true);
        bool ok = true;
        size_t total = 0;
        const FrameMatcher::Mode modes[3] = {FrameMatcher::MODE_ALL, FrameMatcher::MODE_UNASSIGNED, FrameMatcher::MODE_ASSIGNED};
        for (int mi = 0; mi < 2; mi++) {
            FrameMatcher ref(FrameMatcher::TYPE_FLANN);
            ref.setParams(F[0], modes[mi], 100.f, 0.6f, true, 3);
            _impl::FrameMatcher_B200 dev;
            dev.setParams(F[0], modes[mi], 100.f, 0.6f, true, 3);
            for (int q = 1; q < 3; q++) {
                const std::vector<cv::DMatch> a = ref.match(F[q], modes[mi]), b = dev.match(F[q], modes[mi]);
                ok = ok && same(a, b);
                total += a.size();
            }
        }
        EXPECT(ok && total > 1500, "FrameMatcher_B200 == the reference's FrameMatcher_Flann (exact index) on MODE_ALL / MODE_UNASSIGNED");
        F[0].ids.assign(F[0].ids.size(), std::numeric_limits<uint32_t>::max());
        F[1].ids.assign(F[1].ids.size(), std::numeric_limits<uint32_t>::max());
        F[1].flags[20].reset();
    }

    // ---- the map: frame 0's keypoints back-projected onto the plane --------------------------------------------------------------------
    std::shared_ptr<Map> map = std::make_shared<Map>();
    const int n0 = (int)F[0].und_kpts.size();
    for (int i = 0; i < n0; i++) {
        MapPoint& p = map->map_points.add(i);
        const cv::KeyPoint& k = F[0].und_kpts[i];
        p.id = i;
        p.pos3d = cv::Point3f((k.pt.x - cx) * Z / f, (k.pt.y - cy) * Z / f, Z);
        const float d = (float)cv::norm(p.pos3d);
        p.normal = cv::Point3f(-p.pos3d.x / d, -p.pos3d.y / d, -p.pos3d.z / d);
        p.mfMaxDistance = d * F[0].scaleFactors[k.octave];
        p.mfMinDistance = p.mfMaxDistance / F[0].scaleFactors.back();
        p._desc = F[0].desc.row(i).clone();
        p.frames[0] = i;
        p.stable = i % 5 != 0;
        F[0].ids[i] = i;
    }
    for (int k = 0; k < 3; k++) map->keyframes.add(k) = F[k];
    map->keyframes[0].create_kdtree(); map->keyframes[1].create_kdtree(); map->keyframes[2].create_kdtree();

    // ---- 3. the projection matcher ----------------------------------------------------------------------------------------------------
    uco_b200::Context ctx;
    std::vector<cv::DMatch> m12[3];
    {
        bool ok = true;
        for (int k = 1; k < 3; k++) {
            Frame& cur = map->keyframes[k];
            for (int i = 0; i < n0; i++) map->map_points[i].nVisible = 0;
            const std::vector<cv::DMatch> a = map->matchFrameToMapPoints({0u}, cur, cur.pose_f2g, 100.f, 15.f, true, true);   // the reference's statements
            std::vector<int> visA(n0);
            for (int i = 0; i < n0; i++) { visA[i] = map->map_points[i].nVisible; map->map_points[i].nVisible = 0; }
            std::vector<uint32_t> used{0u};
            const std::vector<uint32_t> ids = map->getMapPointsInFrames(used.begin(), used.end());
            const std::vector<cv::DMatch> b = matchFrameToMapPoints_b200(ctx, *map, ids, cur, cur.pose_f2g, 100.f, 15.f, true);
            bool vis = true;
            for (int i = 0; i < n0; i++) vis = vis && visA[i] == map->map_points[i].nVisible;
            ok = ok && same(a, b) && vis && a.size() > 500;
            m12[k] = b;
        }
        EXPECT(ok, "matchFrameToMapPoints_b200 == the reference's Map::matchFrameToMapPoints (matches + setVisible marks, 2 frames)");
    }

    // ---- 4. pose-only optimisation ----------------------------------------------------------------------------------------------------
    {
        bool ok = true;
        for (int k = 1; k < 3; k++) {
            const Frame& cur = map->keyframes[k];
            std::vector<cv::DMatch> mm = m12[k];
            // a perturbed start and a few gross outliers
            cv::Mat start = cur.pose_f2g.clone();
            start.at<float>(0, 3) += 0.01f; start.at<float>(1, 3) -= 0.008f; start.at<float>(2, 3) += 0.02f;
            for (size_t i = 0; i < mm.size(); i += 37) mm[i].trainIdx = (mm[i].trainIdx + 101) % n0;
            se3 pose;
            pose = start;
            std::vector<cv::DMatch> dm = mm;
            const int ngood = solvePnp_b200(ctx, cur, map, dm, pose, -1);
            const size_t n = mm.size();
            std::vector<float> p3(3 * n), uv(2 * n), ur(n, 0.f), inv(n);
            std::vector<uint8_t> st(n, 0), stable(n), bad(n);
            for (size_t i = 0; i < n; i++) {
                const cv::KeyPoint& kp = cur.und_kpts[mm[i].queryIdx];
                const MapPoint& mp = map->map_points[mm[i].trainIdx];
                p3[3 * i] = mp.pos3d.x; p3[3 * i + 1] = mp.pos3d.y; p3[3 * i + 2] = mp.pos3d.z;
                uv[2 * i] = kp.pt.x; uv[2 * i + 1] = kp.pt.y;
                inv[i] = 1.f / cur.scaleFactors[kp.octave];
                stable[i] = mp.isStable();
            }
            float out44[16]; double out7[7]; int its[4];
            const int rgood = ref_pose_only(start.ptr<float>(0), (int)n, p3.data(), uv.data(), ur.data(), st.data(), inv.data(), stable.data(), f, f, cx, cy, 0.f,
                                            0, nullptr, nullptr, nullptr, out44, out7, bad.data(), its);
            double dmax = 0;
            const cv::Mat got = pose.convert();
            for (int e = 0; e < 16; e++) dmax = std::max(dmax, (double)std::fabs(got.ptr<float>(0)[e] - out44[e]));
            bool flags = true;
            for (size_t i = 0; i < n; i++) flags = flags && (dm[i].imgIdx == (bad[i] ? -1 : 1));
            double terr = std::fabs(got.at<float>(0, 3) - cur.pose_f2g.at_(3)) + std::fabs(got.at<float>(1, 3) - cur.pose_f2g.at_(7)) + std::fabs(got.at<float>(2, 3));
            ok = ok && ngood == rgood && flags && dmax < 1e-6 && ngood > 400 && terr < 5e-3;
        }
        EXPECT(ok, "solvePnp_b200 == the reference's g2o pose-only solve (pose 1e-6, inlier flags, count) and recovers the pose");
    }

    // ---- 6. new-map-point creation (the mapper's createNewPoints body) ---------------------------------------------------------------
    {
        struct NewPointInfo { cv::Point3d pose; bool isStereo = false; std::vector<std::pair<uint32_t, uint32_t>> frame_kpt; float dist = std::numeric_limits<float>::max(); };
        // fresh copies of the three keyframes with no keypoint assigned: everything is a candidate; a tilted, displaced set of cameras
        // would need other images, so the poses stay the sliding ones (baseline 3-4 cm at 2 m: parallax cosine ~0.9998 -> use frames 1, 2
        // against 0 with the parallax gate doing its job on part of the matches)
        auto fresh = std::make_shared<Map>();
        for (int k = 0; k < 3; k++) {
            Frame& kf = fresh->keyframes.add(k);
            kf = F[k];
            kf.ids.assign(kf.und_kpts.size(), std::numeric_limits<uint32_t>::max());
            kf.pose_f2g = Se3Transform();
            // a wider baseline than the texture shift: the views are those of cameras 0.25 m apart looking at a plane 2 m away whose texture
            // moved with them by all but (off[k] - off[0]) pixels — geometrically: points at depth Zk with f * B / Zk = pixel shift
            kf.pose_f2g[3] = -(off[k][0] - off[0][0]) * Z / f; kf.pose_f2g[7] = -(off[k][1] - off[0][1]) * Z / f;
        }
        g_real_f12 = true;
        uco_b200::Context c6;
        std::vector<NewPointInfo> pts = createNewPoints_b200<NewPointInfo>(c6, *fresh, fresh->keyframes[0], std::vector<uint32_t>{1, 2}, 100000, 50.f, 1.2f);
        g_real_f12 = false;
        bool ok = !pts.empty();
        size_t two = 0;
        double worst = 0;
        for (const auto& pnt : pts) {
            ok = ok && pnt.frame_kpt.size() >= 2 && pnt.frame_kpt[0].first == 0 && pnt.frame_kpt[0].second < fresh->keyframes[0].und_kpts.size();
            two += pnt.frame_kpt.size() == 3;
            // the point re-projects onto its keypoints (5.998 chi2 gate at the keypoint's scale) and lies in front of the cameras
            for (const auto& fk : pnt.frame_kpt) {
                const Frame& kf = fresh->keyframes[fk.first];
                const cv::KeyPoint& kp = kf.und_kpts[fk.second];
                const double X = pnt.pose.x + kf.pose_f2g.at_(3), Y = pnt.pose.y + kf.pose_f2g.at_(7), Zc = pnt.pose.z;
                const double u = f * X / Zc + cx, v = f * Y / Zc + cy, s = kf.scaleFactors[kp.octave];
                worst = std::max(worst, ((u - kp.pt.x) * (u - kp.pt.x) + (v - kp.pt.y) * (v - kp.pt.y)) / (s * s));
                ok = ok && Zc > 0;
            }
        }
        std::printf("     new points: %zu (%zu seen by both neighbours), worst reprojection chi2 %.3f\n", pts.size(), two, worst);
        EXPECT(ok && pts.size() > 10 && worst < 3 * 5.998,   // the 8-pixel texture shifts are right at the 0.9998 parallax gate: few matches pass it
               "createNewPoints_b200 (mapper's new-point creation body): points re-project onto the keypoints they were made from");
    }

    // ---- 5. bundle adjustment ----------------------------------------------------------------------------------------------------------
    {
        for (int k = 1; k < 3; k++)
            for (const auto& m : m12[k]) {
                if (map->keyframes[k].ids[m.queryIdx] != std::numeric_limits<uint32_t>::max()) continue;
                map->keyframes[k].ids[m.queryIdx] = m.trainIdx;
                map->map_points[m.trainIdx].frames[k] = m.queryIdx;
            }
        for (int k = 1; k < 3; k++) {    // disturb the free keyframes and the points
            map->keyframes[k].pose_f2g[3] += 0.004f * k; map->keyframes[k].pose_f2g[11] -= 0.006f;
        }
        for (int i = 0; i < n0; i += 3) map->map_points[i].pos3d.z += 0.01f;
        std::shared_ptr<GlobalOptimizer> opt = std::make_shared<GlobalOptimizerB200>();
        GlobalOptimizer::ParamSet ps;
        ps.nIters = 5;
        ps.fixFirstFrame = true;
        opt->setParams(map, ps);
        const uco_ba_problem& pb = static_cast<GlobalOptimizerB200*>(opt.get())->problem();
        std::vector<double> r7(7 * pb.n_poses), r3(3 * pb.n_points), rchi(pb.n_obs), trace(128);
        std::vector<float> r44(16 * pb.n_poses);
        std::vector<uint8_t> rlev(pb.n_obs), rbad(pb.n_obs);
        int its[2] = {0, 0};
        ref_ba_optimize(pb.n_poses, pb.poses44, pb.fixed, pb.n_points, pb.points3, pb.n_obs, pb.obs_pose, pb.obs_point, pb.obs_uv, pb.obs_ur, pb.obs_stereo,
                        pb.obs_inv_sigma2, pb.fx, pb.fy, pb.cx, pb.cy, pb.bf, 5, r7.data(), r44.data(), r3.data(), rchi.data(), rlev.data(), rbad.data(), its,
                        trace.data());
        bool stop = false;
        opt->optimize(&stop);
        opt->getResults(map);
        double dmax = 0;
        for (int k = 0; k < pb.n_poses; k++) {
            const Frame& fr = map->keyframes[k];      // the adapter numbered the frames 0, 1, 2 in order of use
            if (pb.fixed[k]) continue;
            for (int e = 0; e < 12; e++) dmax = std::max(dmax, (double)std::fabs(fr.pose_f2g.at_(e) - r44[16 * k + e]));
        }
        size_t nbad = 0;
        for (auto b : rbad) nbad += b;
        EXPECT(pb.n_poses == 3 && pb.n_points > 500 && pb.n_obs > 1500 && dmax < 1e-5 && opt->getBadAssociations().size() == nbad &&
                   map->nNormalUpdates == pb.n_points && opt->getName() == "b200",
               "GlobalOptimizerB200 (reference's GlobalOptimizer interface) == the reference's g2o on the flattened window (poses 1e-5, bad associations)");
        // keyframe 1 taken with another camera: the adapter flattens one fx fy cx cy bf row per keyframe and the solve still equals the
        // reference's g2o with the per-edge ImageParams of globaloptimizer_g2o.cpp:233-236
        map->keyframes[1].imageParams.CameraMatrix.at<float>(0, 0) *= 1.002f;
        map->keyframes[1].imageParams.CameraMatrix.at<float>(1, 1) *= 1.003f;
        map->keyframes[1].imageParams.CameraMatrix.at<float>(0, 2) += 0.5f;
        map->nNormalUpdates = 0;
        opt->setParams(map, ps);
        const uco_ba_problem& pc = static_cast<GlobalOptimizerB200*>(opt.get())->problem();
        EXPECT(pc.pose_cam != nullptr && pc.pose_cam[5] != pc.pose_cam[0] && pc.pose_cam[10] == pc.pose_cam[0], "mixed-camera window: one camera row per keyframe");
        if (pc.pose_cam) {
            std::vector<double> c7(7 * pc.n_poses), c3(3 * pc.n_points), cchi(pc.n_obs);
            std::vector<float> c44(16 * pc.n_poses);
            std::vector<uint8_t> clev(pc.n_obs), cbad(pc.n_obs);
            ref_ba_optimize_cams(pc.n_poses, pc.poses44, pc.fixed, pc.n_points, pc.points3, pc.n_obs, pc.obs_pose, pc.obs_point, pc.obs_uv, pc.obs_ur, pc.obs_stereo,
                                 pc.obs_inv_sigma2, pc.fx, pc.fy, pc.cx, pc.cy, pc.bf, 5, c7.data(), c44.data(), c3.data(), cchi.data(), clev.data(), cbad.data(), its,
                                 trace.data(), 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, pc.pose_cam);
            opt->optimize(&stop);
            opt->getResults(map);
            double cmax = 0;
            for (int k = 0; k < pc.n_poses; k++) {
                if (pc.fixed[k]) continue;
                for (int e = 0; e < 12; e++) cmax = std::max(cmax, (double)std::fabs(map->keyframes[k].pose_f2g.at_(e) - c44[16 * k + e]));
            }
            size_t cb = 0;
            for (auto b : cbad) cb += b;
            EXPECT(cmax < 1e-5 && opt->getBadAssociations().size() == cb,
                   "GlobalOptimizerB200 on a window taken with two cameras == the reference's g2o with per-edge ImageParams (poses 1e-5, bad associations)");
        }
        // InPlaneMarkers (:356-401): two coplanar markers, each seen by all three keyframes; the adapter picks the reference marker, ties the other one to
        // it and the solve equals the reference's g2o with MarkerEdgeX (restated in ref_g2o_wrap.cpp) on the window the adapter flattened
        {
            GlobalOptimizer::ParamSet pps;
            pps.nIters = 5;
            pps.fixFirstFrame = true;
            pps.InPlaneMarkers = true;
            const float msz = 0.2f;
            for (uint32_t id = 4; id <= 5; id++) {
                Marker mk;
                mk.id = id; mk.size = msz;
                Se3Transform G;                                  // identity rotation: both markers in the plane z = 2.0 in front of keyframe 0's camera
                G = map->keyframes[0].pose_f2g.inv();
                cv::Mat S = cv::Mat::eye(4, 4, CV_32F);
                S.at<float>(0, 3) = id == 4 ? -0.35f : 0.3f; S.at<float>(1, 3) = id == 4 ? 0.1f : -0.15f; S.at<float>(2, 3) = 2.0f;
                Se3Transform Gm;
                for (int r = 0; r < 4; r++)
                    for (int c = 0; c < 4; c++) {
                        float v = 0;
                        for (int k = 0; k < 4; k++) v += G.at<float>(r, k) * S.at<float>(k, c);
                        Gm.at<float>(r, c) = v;
                    }
                mk.pose_g2m = Gm;
                for (int k = 0; k < 3; k++) {
                    const Frame& fr = map->keyframes[k];
                    ucoslam::MarkerObservation mo;
                    mo.id = id;
                    for (const auto& pl : Marker::get3DPointsLocalRefSystem(msz)) {
                        float g[3], c[3];
                        for (int r = 0; r < 3; r++) g[r] = Gm.at<float>(r, 0) * pl.x + Gm.at<float>(r, 1) * pl.y + Gm.at<float>(r, 2) * pl.z + Gm.at<float>(r, 3);
                        for (int r = 0; r < 3; r++) c[r] = fr.pose_f2g.at<float>(r, 0) * g[0] + fr.pose_f2g.at<float>(r, 1) * g[1] + fr.pose_f2g.at<float>(r, 2) * g[2] + fr.pose_f2g.at<float>(r, 3);
                        mo.und_corners.push_back(cv::Point2f(fr.imageParams.fx() * c[0] / c[2] + fr.imageParams.cx() + 0.3f * (id - 4), fr.imageParams.fy() * c[1] / c[2] + fr.imageParams.cy() - 0.2f * k));
                    }
                    map->keyframes[k].markers.push_back(mo);
                    mk.frames.insert(k);
                }
                mk.pose_g2m[3] += 0.01f * (id - 3); mk.pose_g2m[11] -= 0.015f;   // start away from the observations
                map->map_markers[id] = mk;
            }
            opt->setParams(map, pps);
            const uco_ba_problem& pq = static_cast<GlobalOptimizerB200*>(opt.get())->problem();
            EXPECT(pq.n_markers == 2 && pq.n_marker_obs == 6 && pq.n_plane == 1 && pq.plane_ref == 0 && pq.plane_other[0] == 1 && pq.plane_weight > 0,
                   "InPlaneMarkers: reference marker = the first of the two equally seen ones, one planar edge to the other, weight per :381-382");
            std::vector<double> q7(7 * pq.n_poses), q3(3 * pq.n_points), qchi(pq.n_obs), m7(7 * 2), mchi(6);
            std::vector<float> q44(16 * pq.n_poses), m44(16 * 2);
            std::vector<uint8_t> qlev(pq.n_obs), qbad(pq.n_obs);
            ref_ba_optimize_planar(pq.n_poses, pq.poses44, pq.fixed, pq.n_points, pq.points3, pq.n_obs, pq.obs_pose, pq.obs_point, pq.obs_uv, pq.obs_ur, pq.obs_stereo,
                                   pq.obs_inv_sigma2, pq.fx, pq.fy, pq.cx, pq.cy, pq.bf, 5, q7.data(), q44.data(), q3.data(), qchi.data(), qlev.data(), qbad.data(), its,
                                   trace.data(), pq.n_markers, pq.marker_pose44, pq.marker_size, pq.n_marker_obs, pq.mobs_marker, pq.mobs_pose, pq.mobs_corners,
                                   pq.mobs_weight, m7.data(), m44.data(), mchi.data(), pq.pose_cam, pq.n_plane, pq.plane_ref, pq.plane_ref_pose44, pq.plane_other,
                                   pq.plane_weight);
            opt->optimize(&stop);
            opt->getResults(map);
            double pmax = 0, mmax = 0;
            for (int k = 0; k < pq.n_poses; k++) {
                if (pq.fixed[k]) continue;
                for (int e = 0; e < 12; e++) pmax = std::max(pmax, (double)std::fabs(map->keyframes[k].pose_f2g.at_(e) - q44[16 * k + e]));
            }
            for (int m = 0; m < 2; m++)
                for (int e = 0; e < 12; e++) mmax = std::max(mmax, (double)std::fabs(map->map_markers[4 + m].pose_g2m.at_(e) - m44[16 * m + e]));
            EXPECT(pmax < 1e-4 && mmax < 6e-3, "GlobalOptimizerB200 with InPlaneMarkers == the reference's g2o with planar marker edges (keyframes 1e-4, markers 6e-3)");
        }
    }
    std::printf("%s\n", fails ? "ADAPTER WORLD FAILED" : "ADAPTER WORLD OK");
    return fails ? 1 : 0;
}
