// undistort_math.h — one point through ucoslam::undistortPoints (src/basictypes/misc.cpp:269-292): cv::undistortPoints with its
// default termination (5 fixed-point iterations of the inverse Brown-Conrady model, OpenCV imgproc undistort.dispatch.cpp
// cvUndistortPointsInternal; OpenCV is not part of the reference tree, its published algorithm is restated here in its operation
// order, double arithmetic on float inputs, float result) followed by the reference's x*fx+cx in float.  Shared by the kernel
// (undistort.cu) and a host probe so that it can be pinned against cv2.undistortPoints without a device.
#pragma once
#ifdef __CUDACC__
#define UND_HD __host__ __device__ __forceinline__
#else
#define UND_HD inline
#endif

// K = fx fy cx cy (float, as ImageParams::CameraMatrix holds them); k[14] = distortion coefficients padded with zeros
// (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 taux tauy; the tilt terms must be zero)
UND_HD void undistort_point(float u, float v, const float* K, const double* k, float* ou, float* ov) {
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    double x = u, y = v;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) {   // OpenCV gives up on the point and returns the normalised input
            x = ((double)u - cx) * ifx;
            y = ((double)v - cy) * ify;
            break;
        }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const float xn = (float)x, yn = (float)y;       // cv::undistortPoints writes Point2f
    *ou = xn * K[0] + K[2];                         // misc.cpp:283-284 / :289-290, float
    *ov = yn * K[1] + K[3];
}
