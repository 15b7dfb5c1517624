// ba_math.cuh — per-observation device arithmetic of the bundle adjustment (f64): SE3 (unit quaternion + translation)
// mapping / exponential update, mono and stereo reprojection residuals with their Jacobians, Huber weighting.
// Each function names the reference code whose arithmetic it reproduces (operation order kept, no FMA contraction:
// the library is compiled with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ba {

struct Cam {
    double fx, fy, cx, cy, bf;
    float bf_f;  // EdgeStereoSE3ProjectXYZ::cam_project takes bf as const float& (typesg2o.h:398)
};

struct Pose {  // g2o::SE3Quat: q = (x, y, z, w), t
    double q[4], t[3];
};

__device__ __forceinline__ Pose load_pose(const double* p) {
    Pose T;
#pragma unroll
    for (int k = 0; k < 4; k++) T.q[k] = p[k];
#pragma unroll
    for (int k = 0; k < 3; k++) T.t[k] = p[4 + k];
    return T;
}
__device__ __forceinline__ void store_pose(double* p, const Pose& T) {
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = T.q[k];
#pragma unroll
    for (int k = 0; k < 3; k++) p[4 + k] = T.t[k];
}

// Eigen::Quaternion::_transformVector (what SE3Quat::map uses, se3quat.h:270-273)
__device__ __forceinline__ void quat_rot(const double* q, const double* v, double* o) {
    double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
    ux += ux; uy += uy; uz += uz;
    o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
__device__ __forceinline__ void se3_map(const Pose& T, const double* x, double* o) {
    quat_rot(T.q, x, o);
    o[0] += T.t[0]; o[1] += T.t[1]; o[2] += T.t[2];
}
// Eigen::Quaternion::toRotationMatrix, row-major
__device__ __forceinline__ void quat_to_R(const double* q, double* R) {
    double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen::Quaternion(Matrix3) (Shepperd), row-major m
__device__ __forceinline__ void quat_from_R(const double* m, double* q) {
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        double qq[4];
        qq[i] = 0.5 * t;
        t = 0.5 / t;
        qq[3] = (m[3 * k + j] - m[3 * j + k]) * t;
        qq[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        qq[k] = (m[3 * k + i] + m[3 * i + k]) * t;
        q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
    }
}
// SE3Quat::normalizeRotation, se3quat.h:345-350
__device__ __forceinline__ void quat_normalize(double* q) {
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// Frame::pose_f2g (row-major f32 4x4) -> SE3Quat, globaloptimizer_g2o.cpp:80-91
__device__ __forceinline__ Pose pose_from_m44f(const float* m) {
    double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
    Pose T;
    quat_from_R(R, T.q);
    quat_normalize(T.q);
    T.t[0] = m[3]; T.t[1] = m[7]; T.t[2] = m[11];
    return T;
}
// VertexSE3Expmap::oplusImpl: T <- SE3Quat::exp(u) * T  (typesg2o.h:76-79, se3quat.h:276-314 and 156-163); u = (omega, upsilon)
__device__ __forceinline__ void se3_oplus(Pose& T, const double* u) {
    double th = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double O[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0}, O2[9], R[9], V[9];
    mat3_mul(O, O, O2);
    double a, b, c, d;
    if (th < 0.00001) {
        a = 1; b = 0.5; c = 0.5; d = 1.0 / 6.0;
    } else {
        double s = sin(th), co = cos(th);
        a = s / th; b = (1 - co) / (th * th);
        c = b; d = (th - s) / pow(th, 3.0);
    }
#pragma unroll
    for (int i = 0; i < 9; i++) {
        double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + a * O[i] + b * O2[i];
        V[i] = I + c * O[i] + d * O2[i];
    }
    Pose E, Nw;
    quat_from_R(R, E.q);
    quat_normalize(E.q);
#pragma unroll
    for (int i = 0; i < 3; i++) E.t[i] = V[3 * i] * u[3] + V[3 * i + 1] * u[4] + V[3 * i + 2] * u[5];
    double rt[3];
    quat_rot(E.q, T.t, rt);
    Nw.t[0] = E.t[0] + rt[0]; Nw.t[1] = E.t[1] + rt[1]; Nw.t[2] = E.t[2] + rt[2];
    const double *x = E.q, *y = T.q;
    Nw.q[3] = x[3] * y[3] - x[0] * y[0] - x[1] * y[1] - x[2] * y[2];
    Nw.q[0] = x[3] * y[0] + x[0] * y[3] + x[1] * y[2] - x[2] * y[1];
    Nw.q[1] = x[3] * y[1] + x[1] * y[3] + x[2] * y[0] - x[0] * y[2];
    Nw.q[2] = x[3] * y[2] + x[2] * y[3] + x[0] * y[1] - x[1] * y[0];
    quat_normalize(Nw.q);
    T = Nw;
}

// residual z - project(p) for a point already in camera coordinates (typesg2o.h:267-272,318-320 mono; 349-354,398-405 stereo:
// the stereo projection keeps 1/z in a FLOAT)
__device__ __forceinline__ void residual(const double* p, const double* z, bool stereo, const Cam& c, double* e) {
    if (!stereo) {
        e[0] = z[0] - ((p[0] / p[2]) * c.fx + c.cx);
        e[1] = z[1] - ((p[1] / p[2]) * c.fy + c.cy);
        e[2] = 0;
    } else {
        const float invz = (float)(1.0 / p[2]);  // 1.0f / double -> double division, narrowed
        double r0 = p[0] * invz * c.fx + c.cx, r1 = p[1] * invz * c.fy + c.cy;
        double r2 = r0 - (double)(c.bf_f * invz);
        e[0] = z[0] - r0; e[1] = z[1] - r1; e[2] = z[2] - r2;
    }
}
// d residual / d pose (D x 6, rotation first), typesg2o.h:302-314 / 382-396.  The reference divides by z and z^2 entry by
// entry; here 1/z is formed once and multiplied in (a division is ~30 FP64 instructions): entries differ from the
// reference's by at most a couple of ulp, which moves an LM step by ~1e-16 relative and no fixed point at all.
__device__ __forceinline__ void jac_pose(const double* p, bool stereo, const Cam& c, double* JT) {
    const double x = p[0], y = p[1], iz = 1.0 / p[2], iz2 = iz * iz, fx = c.fx, fy = c.fy;
    const double xz = x * iz, yz = y * iz;
    JT[0] = xz * yz * fx; JT[1] = -(1 + xz * xz) * fx; JT[2] = yz * fx; JT[3] = -iz * fx; JT[4] = 0; JT[5] = x * iz2 * fx;
    JT[6] = (1 + yz * yz) * fy; JT[7] = -xz * yz * fy; JT[8] = -xz * fy; JT[9] = 0; JT[10] = -iz * fy; JT[11] = y * iz2 * fy;
    if (stereo) {
        const double bz = c.bf * iz2;
        JT[12] = JT[0] - bz * y; JT[13] = JT[1] + bz * x; JT[14] = JT[2]; JT[15] = JT[3]; JT[16] = 0; JT[17] = JT[5] - bz;
    } else {
#pragma unroll
        for (int k = 12; k < 18; k++) JT[k] = 0;
    }
}
// d residual / d point (D x 3), typesg2o.h:290-300 / 370-380 (same remark on 1/z)
__device__ __forceinline__ void jac_point(const double* p, const double* R, bool stereo, const Cam& c, double* JX) {
    const double x = p[0], y = p[1], iz = 1.0 / p[2], iz2 = iz * iz, fx = c.fx, fy = c.fy;
    const double a0 = -fx * iz, a2 = fx * x * iz2, b1 = -fy * iz, b2 = fy * y * iz2;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        JX[k] = a0 * R[k] + a2 * R[6 + k];
        JX[3 + k] = b1 * R[3 + k] + b2 * R[6 + k];
        JX[6 + k] = stereo ? JX[k] - c.bf * iz2 * R[6 + k] : 0.0;
    }
}
// g2o::RobustKernelHuber::robustify (robust_kernel_impl.cpp:65-79) scaled by WeightedHubberRobustKernel's weight
// (typesg2o.h:91-106): rho0 = robustified chi2, rho1 = weight of the information matrix
__device__ __forceinline__ void huber(double e2, double delta, double weight, double& rho0, double& rho1) {
    double dsqr = delta * delta;
    if (e2 <= dsqr) {
        rho0 = weight * e2; rho1 = 1.;
    } else {
        double s = sqrt(e2);
        rho0 = weight * (2 * s * delta - dsqr); rho1 = delta / s;
    }
}
// inverse of a symmetric 3x3 given as (00 01 02 11 12 22); result in the same packing
__device__ __forceinline__ void inv3_sym(const double* D, double* I) {
    double a = D[0], b = D[1], c = D[2], d = D[3], e = D[4], f = D[5];
    double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    double det = a * c00 + b * c01 + c * c02;
    double id = 1.0 / det;
    I[0] = c00 * id; I[1] = c01 * id; I[2] = c02 * id;
    I[3] = (a * f - c * c) * id; I[4] = (b * c - a * e) * id; I[5] = (a * d - b * b) * id;
}

}  // namespace ba
