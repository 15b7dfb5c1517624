// adapter_syntax.cpp — TEST INFRASTRUCTURE.  Compile (and link) check of the adapters that need only Frame / ImageParams and
// OpenCV containers: stereo_depth_b200.h, triangulate_b200.h, undistort_b200.h, against oracle/shim.  Built by `make -C oracle ref`;
// the program itself only runs where a CUDA device exists (it is a smoke run of the three calls, tests/test_adapters_gpu.py).
#include <cstdio>
#include "stereo_depth_b200.h"
#include "triangulate_b200.h"
#include "undistort_b200.h"

int main() {
    uco_b200::Context ctx;
    ucoslam::ImageParams ip;
    ip.CameraMatrix = cv::Mat(3, 3, CV_32F);
    ip.CameraMatrix.at<float>(0, 0) = 525; ip.CameraMatrix.at<float>(1, 1) = 525; ip.CameraMatrix.at<float>(0, 2) = 80; ip.CameraMatrix.at<float>(1, 2) = 60;
    ip.CameraMatrix.at<float>(2, 2) = 1;
    ip.Distorsion = cv::Mat(1, 5, CV_32F);
    ip.Distorsion.at<float>(0, 0) = 0.1f;
    ip.bl = 0.12f;
    std::vector<cv::Point2f> pts(3), out;
    pts[0] = cv::Point2f(10, 20); pts[1] = cv::Point2f(80, 60); pts[2] = cv::Point2f(150, 100);
    ucoslam::undistortPoints_b200(ctx, pts, ip, &out);
    std::vector<cv::KeyPoint> kps(2), und;
    kps[0].pt = pts[0]; kps[1].pt = pts[2];
    ucoslam::undistortKeyPoints_b200(ctx, kps, ip, und);
    bool ok = out.size() == 3 && und.size() == 2 && und[0].pt.x == out[0].x && und[1].pt.y == out[2].y && out[1].x == 80.f;

    ucoslam::Frame a, b;
    a.imageParams = b.imageParams = ip;
    a.scaleFactors = b.scaleFactors = {1.f, 1.2f};
    a.und_kpts.resize(1); b.und_kpts.resize(1);
    a.und_kpts[0].pt = cv::Point2f(90, 60); b.und_kpts[0].pt = cv::Point2f(64.f, 60);   // 0.25 m to the right, 5 m deep: 26.25 px
    cv::Mat RT(4, 4, CV_32F);
    for (int i = 0; i < 4; i++) RT.at<float>(i, i) = 1;
    RT.at<float>(0, 3) = -0.25f;
    std::vector<cv::DMatch> m(1);
    m[0].queryIdx = 0; m[0].trainIdx = 0;
    std::vector<cv::Point3f> p = ucoslam::Triangulate_b200(ctx, a, b, RT, m);
    ok = ok && p.size() == 1 && p[0].z > 4.9f && p[0].z < 5.1f;

    cv::Mat L(120, 160, CV_8UC1), R(120, 160, CV_8UC1), dr(0, 32, CV_8UC1);
    a.desc = cv::Mat(1, 32, CV_8UC1);
    int n = ucoslam::stereoDepth_b200(ctx, L, R, a, std::vector<cv::KeyPoint>(), dr, ip.bl, ip.fx(), 50.f);
    ok = ok && n == 0 && a.depth.size() == 1 && a.depth[0] == 0.f;
    std::printf("%s\n", ok ? "ADAPTER SYNTAX OK" : "ADAPTER SYNTAX FAILED");
    return ok ? 0 : 1;
}
