// orb_math.h — bit-exact restatements of the scalar float routines the reference's extractor calls in third-party code.
// Compiled for the device with -fmad=false and for the host with -ffp-contract=off: every operation below is a single
// IEEE-754 operation in the order written, which is what the CPU libraries execute.
//   fast_atan2_deg : cv::fastAtan2(float y, float x)  (OpenCV core mathfuncs, called at ORBextractor.cpp:105)
//   sincos_glibc   : glibc >= 2.28 sinf / cosf (sysdeps/ieee754/flt-32/s_sincosf.h), the libm behind cos(float)/sin(float)
//                    at ORBextractor.cpp:119.  The double-precision evaluation order follows glibc; the float result is
//                    insensitive to FMA contraction (checked exhaustively for every float in [0, 6.5] on the build host).
//   cv_round_f     : cvRound(float) = round half to even.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
#ifdef __CUDACC__
#define UCO_MHD __host__ __device__ __forceinline__
#else
#define UCO_MHD inline
#endif

namespace uco_math {

UCO_MHD float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s,
                p7 = -0.04432655554792128f * s;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

UCO_MHD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
UCO_MHD uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ff; }

// polynomial of glibc's sinf_poly for quadrant parity n; tab = 0/1 selects the sign-flipped coefficient table
UCO_MHD float sincos_poly(double x, double x2, int tab, int n) {
    const double sg = tab ? -1.0 : 1.0;
    const double c0 = sg * 0x1p0, c1 = sg * -0x1.ffffffd0c621cp-2, c2_ = sg * 0x1.55553e1068f19p-5,
                 c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1_ = s2 + x2 * s3;
        double x7 = x3 * x2;
        double s = x + x3 * s1;
        return (float)(s + x7 * s1_);
    } else {
        double x4 = x2 * x2;
        double cc2 = c3 + x2 * c4;
        double cc1 = c0 + x2 * c1;
        double x6 = x4 * x2;
        double c = cc1 + x4 * c2_;
        return (float)(c + x6 * cc2);
    }
}
// valid for |y| < 120 (the extractor passes angles in [0, 2*pi])
UCO_MHD void sincos_glibc(float y, float* sn, float* cs) {
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    double x = (double)y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) {
            *sn = y;
            *cs = 1.0f;
            return;
        }
        *sn = sincos_poly(x, x2, 0, 0);
        *cs = sincos_poly(x, x2, 0, 1);
        return;
    }
    double r = x * hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = x - (double)n * hpi;
    const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // {1,-1,-1,1}
    const int tab = (n & 2) ? 1 : 0;
    *sn = sincos_poly(x * sign, x * x, tab, n);
    *cs = sincos_poly(x * sign, x * x, tab, n ^ 1);
}

UCO_MHD int cv_round_f(float v) {
#ifdef __CUDA_ARCH__
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}
}  // namespace uco_math
