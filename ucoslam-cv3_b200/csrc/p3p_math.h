// p3p_math.h — perspective-three-point solver and the hypothesis step of PnPSolver::solvePnPRansac, shared by the CUDA kernel
// (ransac.cu) and a host probe (uco_b200_probe_p3p) so that the arithmetic can be checked without a device.
//
// The reference obtains each RANSAC hypothesis from cv::solvePnP(4 points, SOLVEPNP_P3P) (src/optimization/pnpsolver.cpp:67): OpenCV
// solves the three-point problem on the first three correspondences and keeps, among the up to four solutions, the one that
// reprojects the FOURTH point best.  OpenCV is not part of the reference tree; the three-point problem is restated here from its
// textbook form (Grunert's distance equations, Haralick et al. 1994): with s_i the depths along the unit bearings f_i,
//   s_i^2 + s_j^2 - 2 s_i s_j cos(f_i, f_j) = |X_i - X_j|^2,   u = s2/s1, v = s3/s1
// eliminate u (it is a ratio of a quadratic and a linear polynomial in v) to get a quartic in v; the coefficients are formed by
// multiplying those small polynomials numerically instead of from a memorised closed form.  Every positive root gives the three
// camera-frame points s_i f_i; the pose follows from aligning the two congruent triangles (orthonormal triads).  All double.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define P3P_HD __host__ __device__ __forceinline__
#else
#define P3P_HD inline
#endif

P3P_HD double p3p_poly4(const double a[5], double x) { return (((a[4] * x + a[3]) * x + a[2]) * x + a[1]) * x + a[0]; }
P3P_HD double p3p_dpoly4(const double a[5], double x) { return ((4 * a[4] * x + 3 * a[3]) * x + 2 * a[2]) * x + a[1]; }

// real roots of a[4] x^4 + ... + a[0] (Ferrari through the resolvent cubic, three Newton steps of polish); returns their number
P3P_HD int p3p_solve_quartic(const double a[5], double roots[4]) {
    if (a[4] == 0.0) return 0;
    const double b = a[3] / a[4], c = a[2] / a[4], d = a[1] / a[4], e = a[0] / a[4];
    const double p = c - 3.0 * b * b / 8.0;
    const double q = d - b * c / 2.0 + b * b * b / 8.0;
    const double r = e - b * d / 4.0 + b * b * c / 16.0 - 3.0 * b * b * b * b / 256.0;
    int n = 0;
    double y[4];
    const double scale = fabs(p) + sqrt(fabs(r)) + 1e-300;
    if (fabs(q) <= 1e-12 * scale * sqrt(scale)) {   // biquadratic
        const double disc = p * p - 4.0 * r;
        if (disc >= 0) {
            const double sd = sqrt(disc);
            const double z1 = (-p + sd) / 2.0, z2 = (-p - sd) / 2.0;
            if (z1 >= 0) { y[n++] = sqrt(z1); y[n++] = -sqrt(z1); }
            if (z2 >= 0) { y[n++] = sqrt(z2); y[n++] = -sqrt(z2); }
        }
    } else {
        // resolvent: m^3 + A m^2 + B m + C = 0 has a positive real root (value -q^2/8 < 0 at m = 0)
        const double A = p, B = (p * p - 4.0 * r) / 4.0, C = -q * q / 8.0;
        const double P = B - A * A / 3.0, Q = 2.0 * A * A * A / 27.0 - A * B / 3.0 + C;
        const double D = Q * Q / 4.0 + P * P * P / 27.0;
        double z;
        if (D >= 0) {
            const double sD = sqrt(D);
            z = cbrt(-Q / 2.0 + sD) + cbrt(-Q / 2.0 - sD);
        } else {
            const double rho = sqrt(-P / 3.0);
            double arg = 3.0 * Q / (2.0 * P * rho);   // = cos(3 theta)
            arg = arg > 1.0 ? 1.0 : (arg < -1.0 ? -1.0 : arg);
            z = 2.0 * rho * cos(acos(arg) / 3.0);     // the largest of the three real roots
        }
        double m = z - A / 3.0;
        for (int it = 0; it < 3; it++) {              // polish on the cubic
            const double fm = ((m + A) * m + B) * m + C, dfm = (3.0 * m + 2.0 * A) * m + B;
            if (dfm != 0.0) m -= fm / dfm;
        }
        if (!(m > 0)) return 0;
        const double s = sqrt(2.0 * m), h = q / (2.0 * s);
        // y^2 - s y + (p/2 + m + h) = 0   and   y^2 + s y + (p/2 + m - h) = 0
        const double c1 = p / 2.0 + m + h, c2 = p / 2.0 + m - h;
        double d1 = s * s - 4.0 * c1, d2 = s * s - 4.0 * c2;
        const double tol = 1e-9 * (s * s + fabs(c1) + fabs(c2));
        if (d1 < 0 && d1 > -tol) d1 = 0;
        if (d2 < 0 && d2 > -tol) d2 = 0;
        if (d1 >= 0) { const double sd = sqrt(d1); y[n++] = (s + sd) / 2.0; y[n++] = (s - sd) / 2.0; }
        if (d2 >= 0) { const double sd = sqrt(d2); y[n++] = (-s + sd) / 2.0; y[n++] = (-s - sd) / 2.0; }
    }
    for (int i = 0; i < n; i++) {
        double x = y[i] - b / 4.0;
        for (int it = 0; it < 3; it++) {
            const double f = p3p_poly4(a, x), df = p3p_dpoly4(a, x);
            if (df != 0.0) x -= f / df;
        }
        roots[i] = x;
    }
    return n;
}

P3P_HD void p3p_cross(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
P3P_HD double p3p_dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
P3P_HD bool p3p_triad(const double p1[3], const double p2[3], const double p3[3], double E[3][3]) {   // columns e1 e2 e3 as E[k][col]
    double d12[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, d13[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
    const double n1 = sqrt(p3p_dot(d12, d12));
    if (!(n1 > 0)) return false;
    double e1[3] = {d12[0] / n1, d12[1] / n1, d12[2] / n1}, e3[3], e2[3];
    p3p_cross(e1, d13, e3);
    const double n3 = sqrt(p3p_dot(e3, e3));
    if (!(n3 > 1e-12 * sqrt(p3p_dot(d13, d13)))) return false;
    e3[0] /= n3; e3[1] /= n3; e3[2] /= n3;
    p3p_cross(e3, e1, e2);
    for (int k = 0; k < 3; k++) { E[k][0] = e1[k]; E[k][1] = e2[k]; E[k][2] = e3[k]; }
    return true;
}

// X: three world points, f: their unit bearing vectors in the camera.  Out: up to 4 poses (Xc = R Xw + t), R row-major.
P3P_HD int p3p_solve(const double X[3][3], const double f[3][3], double R[4][9], double t[4][3]) {
    double d[3];
    for (int k = 0; k < 3; k++) d[k] = X[1][k] - X[2][k];
    const double a2 = p3p_dot(d, d);
    for (int k = 0; k < 3; k++) d[k] = X[0][k] - X[2][k];
    const double b2 = p3p_dot(d, d);
    for (int k = 0; k < 3; k++) d[k] = X[0][k] - X[1][k];
    const double c2 = p3p_dot(d, d);
    if (!(a2 > 0 && b2 > 0 && c2 > 0)) return 0;
    double Ew[3][3];
    if (!p3p_triad(X[0], X[1], X[2], Ew)) return 0;   // collinear
    const double ca = p3p_dot(f[1], f[2]), cb = p3p_dot(f[0], f[2]), cg = p3p_dot(f[0], f[1]);
    // polynomials in v, ascending coefficients
    const double qv[3] = {1.0, -2.0 * cb, 1.0};
    const double k = c2 - a2;
    const double N[3] = {-b2 + k * qv[0], k * qv[1], b2 + k * qv[2]};
    const double D[2] = {-2.0 * b2 * cg, 2.0 * b2 * ca};
    const double G[3] = {b2 - c2 * qv[0], -c2 * qv[1], -c2 * qv[2]};   // b^2 - c^2 q(v)
    double NN[5] = {0, 0, 0, 0, 0}, ND[4] = {0, 0, 0, 0}, DD[3] = {D[0] * D[0], 2.0 * D[0] * D[1], D[1] * D[1]}, GDD[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) NN[i + j] += N[i] * N[j];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 2; j++) ND[i + j] += N[i] * D[j];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) GDD[i + j] += G[i] * DD[j];
    double poly[5];
    for (int i = 0; i < 5; i++) poly[i] = b2 * NN[i] - (i < 4 ? 2.0 * b2 * cg * ND[i] : 0.0) + GDD[i];
    // normalise (the coefficients scale with the 6th power of the scene size)
    double mx = 0;
    for (int i = 0; i < 5; i++) mx = fmax(mx, fabs(poly[i]));
    if (!(mx > 0)) return 0;
    for (int i = 0; i < 5; i++) poly[i] /= mx;
    double roots[4];
    const int nr = p3p_solve_quartic(poly, roots);
    int ns = 0;
    for (int i = 0; i < nr; i++) {
        const double v = roots[i];
        if (!(v > 0)) continue;
        const double Dv = D[0] + D[1] * v;
        if (fabs(Dv) < 1e-12 * b2) continue;
        const double u = (N[0] + (N[1] + N[2] * v) * v) / Dv;
        if (!(u > 0)) continue;
        const double den = 1.0 + u * u - 2.0 * u * cg;
        if (!(den > 0)) continue;
        const double s1 = sqrt(c2 / den), s2 = u * s1, s3 = v * s1;
        // the third distance equation must hold too (it does for true roots; spurious ones come from the polish of a double root)
        const double chk = s2 * s2 + s3 * s3 - 2.0 * s2 * s3 * ca;
        if (fabs(chk - a2) > 1e-6 * a2) continue;
        bool dup = false;
        for (int j = 0; j < i; j++) dup = dup || roots[j] == v;
        if (dup) continue;
        const double c1[3] = {s1 * f[0][0], s1 * f[0][1], s1 * f[0][2]}, c2p[3] = {s2 * f[1][0], s2 * f[1][1], s2 * f[1][2]},
                     c3[3] = {s3 * f[2][0], s3 * f[2][1], s3 * f[2][2]};
        double Ec[3][3];
        if (!p3p_triad(c1, c2p, c3, Ec)) continue;
        for (int r = 0; r < 3; r++)
            for (int cc = 0; cc < 3; cc++) R[ns][3 * r + cc] = Ec[r][0] * Ew[cc][0] + Ec[r][1] * Ew[cc][1] + Ec[r][2] * Ew[cc][2];
        for (int r = 0; r < 3; r++)
            t[ns][r] = c1[r] - (R[ns][3 * r] * X[0][0] + R[ns][3 * r + 1] * X[0][1] + R[ns][3 * r + 2] * X[0][2]);
        ns++;
    }
    return ns;
}

// One RANSAC hypothesis as cv::solvePnP(4 points, P3P) delivers it: pixels -> bearings with (fx, fy, cx, cy), P3P on points 0..2,
// the solution whose reprojection of point 3 is closest wins.  Returns false when there is no solution.
P3P_HD bool p3p_hypothesis(const double X[4][3], const double px[4][2], const double K[4], double Rb[9], double tb[3]) {
    double f[3][3];
    for (int i = 0; i < 3; i++) {
        const double x = (px[i][0] - K[2]) / K[0], y = (px[i][1] - K[3]) / K[1];
        const double n = sqrt(x * x + y * y + 1.0);
        f[i][0] = x / n; f[i][1] = y / n; f[i][2] = 1.0 / n;
    }
    double R[4][9], t[4][3];
    double X3[3][3];
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) X3[i][k] = X[i][k];
    const int ns = p3p_solve(X3, f, R, t);
    int best = -1;
    double best_err = 0;
    for (int s = 0; s < ns; s++) {
        const double xc = R[s][0] * X[3][0] + R[s][1] * X[3][1] + R[s][2] * X[3][2] + t[s][0];
        const double yc = R[s][3] * X[3][0] + R[s][4] * X[3][1] + R[s][5] * X[3][2] + t[s][1];
        const double zc = R[s][6] * X[3][0] + R[s][7] * X[3][1] + R[s][8] * X[3][2] + t[s][2];
        const double du = K[2] + K[0] * xc / zc - px[3][0], dv = K[3] + K[1] * yc / zc - px[3][1];
        const double err = du * du + dv * dv;
        if (best < 0 || err < best_err) { best = s; best_err = err; }
    }
    if (best < 0 || !(best_err == best_err)) return false;
    for (int k = 0; k < 9; k++) Rb[k] = R[best][k];
    for (int k = 0; k < 3; k++) tb[k] = t[best][k];
    return true;
}
