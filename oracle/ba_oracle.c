/*
 * ba_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle). Plain-C restatement of the reference's bundle adjustment and
 * pose-only optimisation, i.e. of what g2o does for the graph UcoSLAM assembles:
 *   /root/reference/src/optimization/globaloptimizer_g2o.cpp:77-401   graph: SE3 pose vertices (f32 4x4 -> SE3Quat), XYZ points
 *                                                                      (marginalised), one mono / stereo projection edge per
 *                                                                      observation, information = I * invScale[octave], Huber
 *   /root/reference/src/optimization/globaloptimizer_g2o.cpp:418-463  two stages: nIters LM robust; flag chi2 > 5.99 / 7.815 or
 *                                                                      depth <= 0 -> level 1, kernels off; 2*nIters LM
 *   /root/reference/src/optimization/globaloptimizer_g2o.cpp:466-538  results + bad associations
 *   /root/reference/src/optimization/typesg2o.h:249-325, 338-405      residuals and Jacobians of the mono / stereo edges
 *   /root/reference/src/optimization/typesg2o.h:76-79                 pose update: exp(dx) * T
 *   /root/reference/3rdparty/g2o/g2o/types/slam3d/se3quat.h:156-163,270-314,345-350   SE3Quat product / map / exp / normalise
 *   /root/reference/3rdparty/g2o/g2o/core/base_binary_edge.hpp:83-155 quadratic form (rho' weighting only, no rho'' term)
 *   /root/reference/3rdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-79 Huber
 *   /root/reference/3rdparty/g2o/g2o/core/block_solver.hpp:315-443    Schur complement, reduced solve, landmark back-substitution
 *   /root/reference/3rdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:58-175  LM trust region control
 *   /root/reference/3rdparty/g2o/g2o/core/sparse_optimizer.cpp:366-436 outer loop (float chi2 difference stop test)
 * The reduced system is solved by dense Cholesky here (g2o: sparse LDLT, an exact solver too), so results agree with g2o
 * to round-off; parity pin: tests/test_ba_oracle.py checks this file against the reference's own g2o + typesg2o.h compiled
 * from /root/reference (oracle/_ref/libref_g2o.so) and against the golden vectors tests/golden/ba_*.npz generated from it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ---- SE3Quat (q = x y z w, t) -------------------------------------------------------------------------------------- */
typedef struct { double q[4]; double t[3]; } se3;

static void quat_normalize(double* q) { /* se3quat.h:345-350 */
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat_from_R(const double m[9], double* q) { /* Eigen Quaternion(Matrix3), row-major m */
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
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
        q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}
static void quat_to_R(const double* q, double R[9]) { /* Eigen toRotationMatrix */
    double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static void quat_rot(const double* q, const double* v, double* o) { /* Eigen _transformVector */
    double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
    ux += ux; uy += uy; uz += uz;
    o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
static void quat_mul(const double* a, const double* b, double* o) {
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void se3_from_m44f(const float* m, se3* T) { /* globaloptimizer_g2o.cpp:80-91 */
    double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
    quat_from_R(R, T->q);
    quat_normalize(T->q);
    T->t[0] = m[3]; T->t[1] = m[7]; T->t[2] = m[11];
}
static void se3_map(const se3* T, const double* x, double* o) {
    quat_rot(T->q, x, o);
    o[0] += T->t[0]; o[1] += T->t[1]; o[2] += T->t[2];
}
static void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void se3_exp(const double* u, se3* E) { /* se3quat.h:276-314: u = (omega, upsilon) */
    double th = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double O[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0}, O2[9], R[9], V[9];
    mat3_mul(O, O, O2);
    double a, b, c, d;
    if (th < 0.00001) { a = 1; b = 0.5; c = 0.5; d = 1.0 / 6.0; }
    else {
        a = sin(th) / th; b = (1 - cos(th)) / (th * th);
        c = b; d = (th - sin(th)) / pow(th, 3);
    }
    for (int i = 0; i < 9; i++) {
        double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + a * O[i] + b * O2[i];
        V[i] = I + c * O[i] + d * O2[i];
    }
    quat_from_R(R, E->q);
    quat_normalize(E->q);
    for (int i = 0; i < 3; i++) E->t[i] = V[3 * i] * u[3] + V[3 * i + 1] * u[4] + V[3 * i + 2] * u[5];
}
static void se3_oplus(se3* T, const double* u) { /* typesg2o.h:76-79 + se3quat.h:156-163 */
    se3 E, N;
    se3_exp(u, &E);
    double rt[3];
    quat_rot(E.q, T->t, rt);
    N.t[0] = E.t[0] + rt[0]; N.t[1] = E.t[1] + rt[1]; N.t[2] = E.t[2] + rt[2];
    quat_mul(E.q, T->q, N.q);
    quat_normalize(N.q);
    *T = N;
}

/* ---- one observation ----------------------------------------------------------------------------------------------- */
typedef struct { double fx, fy, cx, cy, bf; } cam_t;

/* residual e (2 or 3), typesg2o.h:267-272 / 349-354 + cam_project :318-320 / :398-405 (stereo: float 1/z) */
static void obs_error(const se3* T, const double* X, const double* z, int stereo, const cam_t* c, double* e, double* depth) {
    double p[3];
    se3_map(T, X, p);
    *depth = p[2];
    if (!stereo) {
        e[0] = z[0] - ((p[0] / p[2]) * c->fx + c->cx);
        e[1] = z[1] - ((p[1] / p[2]) * c->fy + c->cy);
        e[2] = 0;
    } else {
        const float invz = 1.0f / p[2];
        double r0 = p[0] * invz * c->fx + c->cx, r1 = p[1] * invz * c->fy + c->cy;
        const float bf = (float)c->bf;
        double r2 = r0 - bf * invz;
        e[0] = z[0] - r0; e[1] = z[1] - r1; e[2] = z[2] - r2;
    }
}
/* Jacobians: JX (D x 3, wrt point), JT (D x 6, wrt pose: rotation first), typesg2o.h:282-315 / 364-396 */
static void obs_jac(const se3* T, const double* X, int stereo, const cam_t* c, double* JX, double* JT) {
    double p[3], R[9];
    se3_map(T, X, p);
    quat_to_R(T->q, R);
    double x = p[0], y = p[1], z = p[2], z_2 = z * z, fx = c->fx, fy = c->fy, bf = c->bf;
    if (!stereo) {
        double t02 = -x / z * fx, t12 = -y / z * fy, s = -1. / z;
        for (int k = 0; k < 3; k++) {
            JX[k] = (s * fx) * R[k] + (s * 0) * R[3 + k] + (s * t02) * R[6 + k];
            JX[3 + k] = (s * 0) * R[k] + (s * fy) * R[3 + k] + (s * t12) * R[6 + k];
        }
    } else {
        for (int k = 0; k < 3; k++) {
            JX[k] = -fx * R[k] / z + fx * x * R[6 + k] / z_2;
            JX[3 + k] = -fy * R[3 + k] / z + fy * y * R[6 + k] / z_2;
            JX[6 + k] = JX[k] - bf * R[6 + k] / z_2;
        }
    }
    JT[0] = x * y / z_2 * fx; JT[1] = -(1 + (x * x / z_2)) * fx; JT[2] = y / z * fx; JT[3] = -1. / z * fx; JT[4] = 0; JT[5] = x / z_2 * fx;
    JT[6] = (1 + y * y / z_2) * fy; JT[7] = -x * y / z_2 * fy; JT[8] = -x / z * fy; JT[9] = 0; JT[10] = -1. / z * fy; JT[11] = y / z_2 * fy;
    if (stereo) {
        JT[12] = JT[0] - bf * y / z_2; JT[13] = JT[1] + bf * x / z_2; JT[14] = JT[2]; JT[15] = JT[3]; JT[16] = 0; JT[17] = JT[5] - bf / z_2;
    }
}
/* Huber, robust_kernel_impl.cpp:65-79 (weight = WeightedHubberRobustKernel::Weight, typesg2o.h:91-106; 1 for plain Huber) */
static void huber(double e2, double delta, double weight, double* rho0, double* rho1) {
    double dsqr = delta * delta;
    if (e2 <= dsqr) { *rho0 = weight * e2; *rho1 = 1.; }
    else { double s = sqrt(e2); *rho0 = weight * (2 * s * delta - dsqr); *rho1 = delta / s; }
}

/* dense Cholesky solve A x = b (A n x n symmetric, lower part used, destroyed); returns 0 if not positive definite */
static int chol_solve(double* A, int n, const double* b, double* x) {
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
        if (!(d > 0)) return 0;
        d = sqrt(d);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= A[i * n + k] * x[k];
        x[i] = s / A[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= A[k * n + i] * x[k];
        x[i] = s / A[i * n + i];
    }
    return 1;
}
static void inv3_sym(const double* D, double* I) { /* D: 3x3 full */
    double a = D[0], b = D[1], c = D[2], d = D[4], e = D[5], f = D[8];
    double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    double det = a * c00 + b * c01 + c * c02;
    double id = 1.0 / det;
    I[0] = c00 * id; I[1] = c01 * id; I[2] = c02 * id;
    I[3] = I[1]; I[4] = (a * f - c * c) * id; I[5] = (b * c - a * e) * id;
    I[6] = I[2]; I[7] = I[5]; I[8] = (a * d - b * b) * id;
}

/* ---- problem ------------------------------------------------------------------------------------------------------- */
typedef struct {
    int P, N, M, Pf;
    se3 *pose, *pose_bak;
    double *pt, *pt_bak;
    const uint8_t* fixed;
    int* free_idx;
    const int32_t *op, *ox;
    double* z;       /* M x 3 */
    double* info;    /* M */
    const uint8_t* stereo;
    uint8_t* active; /* level == 0 */
    int robust;
    cam_t cam;
    double d2, d3;
    double *err, *chi2; /* M x 3, M: state of the LAST error evaluation (what g2o keeps in each edge's _error) */
    /* system */
    double *Hpp, *bp, *Hll, *bl, *Hpl; /* Pf x 36, Pf x 6, N x 9, N x 3, M x 18 */
    int *lm_ptr, *lm_obs;              /* observations grouped by landmark (input order kept inside a group) */
} ba_t;

static double ba_errors(ba_t* B) { /* computeActiveErrors + activeRobustChi2 (sparse_optimizer.cpp:102-116) */
    double chi = 0;
    for (int i = 0; i < B->M; i++) {
        if (!B->active[i]) continue;
        double depth;
        obs_error(&B->pose[B->op[i]], B->pt + 3 * B->ox[i], B->z + 3 * i, B->stereo[i], &B->cam, B->err + 3 * i, &depth);
        const double* e = B->err + 3 * i;
        B->chi2[i] = B->stereo[i] ? (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * B->info[i] : (e[0] * e[0] + e[1] * e[1]) * B->info[i];
        /* g2o: chi2 = e^T Omega e with Omega = I * info: e0*(info*e0) + e1*(info*e1); same value up to round-off */
        if (B->robust) {
            double r0, r1;
            huber(B->chi2[i], B->stereo[i] ? B->d3 : B->d2, 1.0, &r0, &r1);
            chi += r0;
        } else chi += B->chi2[i];
    }
    return chi;
}
static void ba_build(ba_t* B) { /* linearizeOplus + constructQuadraticForm for every active edge */
    memset(B->Hpp, 0, sizeof(double) * 36 * B->Pf);
    memset(B->bp, 0, sizeof(double) * 6 * B->Pf);
    memset(B->Hll, 0, sizeof(double) * 9 * B->N);
    memset(B->bl, 0, sizeof(double) * 3 * B->N);
    memset(B->Hpl, 0, sizeof(double) * 18 * B->M);
    for (int i = 0; i < B->M; i++) {
        if (!B->active[i]) continue;
        int D = B->stereo[i] ? 3 : 2, p = B->op[i], l = B->ox[i], fp = B->free_idx[p];
        double JX[9], JT[18];
        obs_jac(&B->pose[p], B->pt + 3 * l, B->stereo[i], &B->cam, JX, JT);
        double w = B->info[i];
        const double* e = B->err + 3 * i;
        double r1 = 1;
        if (B->robust) { double r0; huber(B->chi2[i], B->stereo[i] ? B->d3 : B->d2, 1.0, &r0, &r1); }
        double wo = r1 * w;                    /* weightedOmega = rho1 * information */
        double orr[3];                         /* omega_r = -(omega e) * rho1 */
        for (int d = 0; d < D; d++) orr[d] = -(w * e[d]) * r1;
        for (int a = 0; a < 3; a++) {
            double s = 0;
            for (int d = 0; d < D; d++) s += JX[3 * d + a] * orr[d];
            B->bl[3 * l + a] += s;
            for (int b = 0; b < 3; b++) {
                double h = 0;
                for (int d = 0; d < D; d++) h += JX[3 * d + a] * wo * JX[3 * d + b];
                B->Hll[9 * l + 3 * a + b] += h;
            }
        }
        if (fp >= 0) {
            for (int a = 0; a < 6; a++) {
                double s = 0;
                for (int d = 0; d < D; d++) s += JT[6 * d + a] * orr[d];
                B->bp[6 * fp + a] += s;
                for (int b = 0; b < 6; b++) {
                    double h = 0;
                    for (int d = 0; d < D; d++) h += JT[6 * d + a] * wo * JT[6 * d + b];
                    B->Hpp[36 * fp + 6 * a + b] += h;
                }
                for (int b = 0; b < 3; b++) { /* Hpl block (6 x 3) = JT^T wo JX */
                    double h = 0;
                    for (int d = 0; d < D; d++) h += JT[6 * d + a] * wo * JX[3 * d + b];
                    B->Hpl[18 * i + 3 * a + b] = h;
                }
            }
        }
    }
}

/* one LM stage = SparseOptimizer::optimize(iterations, minChi2BetweenIter = 1).  trace: {chi2, trials} per outer iteration. */
static int ba_stage(ba_t* B, int iterations, const volatile int* stop, double* trace, int* ntrace) {
    int n = 6 * B->Pf, N = B->N, M = B->M;
    double* S = malloc(sizeof(double) * (n ? n * n : 1));
    double* bs = malloc(sizeof(double) * (n + 1));
    double* xp = malloc(sizeof(double) * (n + 1));
    double* xl = malloc(sizeof(double) * 3 * N);
    double* Dinv = malloc(sizeof(double) * 9 * N);
    double lambda = 0, ni = 2;
    float prevChi2 = FLT_MAX, curChi2 = FLT_MAX, Chi2Diff = FLT_MAX;
    int ok = 1, its = 0;
    double lastChi = 0;
    for (int it = 0; it < iterations && !(stop && *stop) && ok && Chi2Diff > 1.0f; it++) {
        { float t = prevChi2; prevChi2 = curChi2; curChi2 = t; }
        double currentChi = ba_errors(B);
        double tempChi = currentChi;
        ba_build(B);
        if (it == 0) { /* computeLambdaInit :152-166 */
            double md = 0;
            for (int p = 0; p < B->Pf; p++)
                for (int j = 0; j < 6; j++) md = fmax(fabs(B->Hpp[36 * p + 7 * j]), md);
            for (int l = 0; l < N; l++)
                for (int j = 0; j < 3; j++) md = fmax(fabs(B->Hll[9 * l + 4 * j]), md);
            lambda = 1e-5 * md;
            ni = 2;
        }
        double rho = 0;
        int qmax = 0;
        do {
            memcpy(B->pose_bak, B->pose, sizeof(se3) * B->P);
            memcpy(B->pt_bak, B->pt, sizeof(double) * 3 * N);
            /* Schur complement, block_solver.hpp:329-400 */
            memset(S, 0, sizeof(double) * n * n);
            for (int p = 0; p < B->Pf; p++)
                for (int a = 0; a < 6; a++) {
                    for (int b = 0; b < 6; b++) S[(6 * p + a) * n + 6 * p + b] = B->Hpp[36 * p + 6 * a + b];
                    S[(6 * p + a) * n + 6 * p + a] += lambda;
                    bs[6 * p + a] = B->bp[6 * p + a];
                }
            /* observations are visited grouped by landmark */
            for (int l = 0; l < N; l++) {
                double D[9];
                memcpy(D, B->Hll + 9 * l, sizeof(D));
                D[0] += lambda; D[4] += lambda; D[8] += lambda;
                inv3_sym(D, Dinv + 9 * l);
            }
            for (int i = 0; i < M; i++) {
                int fi = B->free_idx[B->op[i]];
                if (!B->active[i] || fi < 0) continue;
                int l = B->ox[i];
                const double* Di = Dinv + 9 * l;
                double Y[18], db[3];
                for (int a = 0; a < 3; a++) db[a] = Di[3 * a] * B->bl[3 * l] + Di[3 * a + 1] * B->bl[3 * l + 1] + Di[3 * a + 2] * B->bl[3 * l + 2];
                for (int a = 0; a < 6; a++) {
                    for (int b = 0; b < 3; b++)
                        Y[3 * a + b] = B->Hpl[18 * i + 3 * a] * Di[b] + B->Hpl[18 * i + 3 * a + 1] * Di[3 + b] + B->Hpl[18 * i + 3 * a + 2] * Di[6 + b];
                    bs[6 * fi + a] -= B->Hpl[18 * i + 3 * a] * db[0] + B->Hpl[18 * i + 3 * a + 1] * db[1] + B->Hpl[18 * i + 3 * a + 2] * db[2];
                }
                for (int jj = B->lm_ptr[l]; jj < B->lm_ptr[l + 1]; jj++) { /* all observations of the same landmark */
                    int j = B->lm_obs[jj];
                    int fj = B->free_idx[B->op[j]];
                    if (!B->active[j] || fj < 0) continue;
                    for (int a = 0; a < 6; a++)
                        for (int b = 0; b < 6; b++)
                            S[(6 * fi + a) * n + 6 * fj + b] -= Y[3 * a] * B->Hpl[18 * j + 3 * b] + Y[3 * a + 1] * B->Hpl[18 * j + 3 * b + 1] + Y[3 * a + 2] * B->Hpl[18 * j + 3 * b + 2];
                }
            }
            int ok2 = n ? chol_solve(S, n, bs, xp) : 1;
            /* landmark back-substitution :413-443 : xl = Dinv (bl - Hpl^T xp) */
            for (int l = 0; l < N; l++) { xl[3 * l] = B->bl[3 * l]; xl[3 * l + 1] = B->bl[3 * l + 1]; xl[3 * l + 2] = B->bl[3 * l + 2]; }
            for (int i = 0; i < M; i++) {
                int fi = B->free_idx[B->op[i]];
                if (!B->active[i] || fi < 0) continue;
                for (int b = 0; b < 3; b++) {
                    double s = 0;
                    for (int a = 0; a < 6; a++) s += B->Hpl[18 * i + 3 * a + b] * xp[6 * fi + a];
                    xl[3 * B->ox[i] + b] -= s;
                }
            }
            for (int l = 0; l < N; l++) {
                double c[3] = {xl[3 * l], xl[3 * l + 1], xl[3 * l + 2]};
                const double* Di = Dinv + 9 * l;
                for (int a = 0; a < 3; a++) xl[3 * l + a] = Di[3 * a] * c[0] + Di[3 * a + 1] * c[1] + Di[3 * a + 2] * c[2];
            }
            /* update */
            if (ok2) {
                for (int p = 0; p < B->P; p++)
                    if (B->free_idx[p] >= 0) se3_oplus(&B->pose[p], xp + 6 * B->free_idx[p]);
                for (int k = 0; k < 3 * N; k++) B->pt[k] += xl[k];
            }
            tempChi = ba_errors(B);
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0; /* computeScale :168-175 */
            for (int k = 0; k < n; k++) scale += xp[k] * (lambda * xp[k] + B->bp[k]);
            for (int k = 0; k < 3 * N; k++) scale += xl[k] * (lambda * xl[k] + B->bl[k]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                double sf = fmax(1. / 3., alpha);
                lambda *= sf;
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                memcpy(B->pose, B->pose_bak, sizeof(se3) * B->P);
                memcpy(B->pt, B->pt_bak, sizeof(double) * 3 * N);
                if (!isfinite(lambda)) break;
            }
            qmax++;
        } while (rho < 0 && qmax < 10 && !(stop && *stop));
        if (qmax == 10 || rho == 0 || !isfinite(lambda)) ok = 0;
        /* curChi2 = activeRobustChi2() over the edges' stored errors (those of the last evaluation) */
        lastChi = 0;
        for (int i = 0; i < M; i++) {
            if (!B->active[i]) continue;
            if (B->robust) { double r0, r1; huber(B->chi2[i], B->stereo[i] ? B->d3 : B->d2, 1.0, &r0, &r1); lastChi += r0; }
            else lastChi += B->chi2[i];
        }
        curChi2 = (float)lastChi;
        Chi2Diff = prevChi2 - curChi2;
        if (trace && *ntrace < 64) { trace[2 * *ntrace] = lastChi; trace[2 * *ntrace + 1] = qmax; (*ntrace)++; }
        its++;
    }
    free(S); free(bs); free(xp); free(xl); free(Dinv);
    return its;
}

/* same signature as oracle/ref_g2o_wrap.cpp: ref_ba_optimize.  */
int oracle_ba_optimize(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs,
                       const int32_t* obs_pose, const int32_t* obs_point, const float* obs_uv, const float* obs_ur,
                       const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx, float fy, float cx, float cy, float bf,
                       int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                       uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace) {
    const float Chi2D = 5.99f, Chi3D = 7.815f; /* globaloptimizer_g2o.h:112-117 */
    ba_t B;
    memset(&B, 0, sizeof(B));
    B.P = n_poses; B.N = n_points; B.M = n_obs;
    B.pose = malloc(sizeof(se3) * n_poses); B.pose_bak = malloc(sizeof(se3) * n_poses);
    B.pt = malloc(sizeof(double) * 3 * n_points + 8); B.pt_bak = malloc(sizeof(double) * 3 * n_points + 8);
    B.free_idx = malloc(sizeof(int) * n_poses);
    B.fixed = fixed;
    for (int i = 0; i < n_poses; i++) {
        se3_from_m44f(poses44 + 16 * i, &B.pose[i]);
        B.free_idx[i] = fixed[i] ? -1 : B.Pf++;
    }
    for (int i = 0; i < 3 * n_points; i++) B.pt[i] = points3[i];
    B.op = obs_pose; B.ox = obs_point; B.stereo = obs_stereo;
    B.z = malloc(sizeof(double) * 3 * n_obs + 8); B.info = malloc(sizeof(double) * n_obs + 8);
    B.active = malloc(n_obs + 1);
    B.err = calloc(3 * n_obs + 1, sizeof(double)); B.chi2 = calloc(n_obs + 1, sizeof(double));
    for (int i = 0; i < n_obs; i++) {
        B.z[3 * i] = obs_uv[2 * i]; B.z[3 * i + 1] = obs_uv[2 * i + 1]; B.z[3 * i + 2] = obs_stereo[i] ? obs_ur[i] : 0;
        B.info[i] = obs_inv_sigma2[i];
        B.active[i] = 1;
    }
    B.cam.fx = fx; B.cam.fy = fy; B.cam.cx = cx; B.cam.cy = cy; B.cam.bf = bf;
    B.d2 = sqrtf(Chi2D); B.d3 = sqrtf(Chi3D);
    B.Hpp = malloc(sizeof(double) * 36 * (B.Pf + 1)); B.bp = malloc(sizeof(double) * 6 * (B.Pf + 1));
    B.Hll = malloc(sizeof(double) * 9 * (n_points + 1)); B.bl = malloc(sizeof(double) * 3 * (n_points + 1));
    B.Hpl = malloc(sizeof(double) * 18 * (n_obs + 1));
    B.lm_ptr = calloc(n_points + 2, sizeof(int)); B.lm_obs = malloc(sizeof(int) * (n_obs + 1));
    for (int i = 0; i < n_obs; i++) B.lm_ptr[obs_point[i] + 1]++;
    for (int l = 0; l < n_points; l++) B.lm_ptr[l + 1] += B.lm_ptr[l];
    {
        int* fill = calloc(n_points + 1, sizeof(int));
        for (int i = 0; i < n_obs; i++) B.lm_obs[B.lm_ptr[obs_point[i]] + fill[obs_point[i]]++] = i;
        free(fill);
    }
    int nt = 0;
    B.robust = 1;
    int it1 = ba_stage(&B, n_iters, 0, trace, &nt);
    for (int i = 0; i < n_obs; i++) { /* globaloptimizer_g2o.cpp:432-449 */
        double p[3];
        se3_map(&B.pose[obs_pose[i]], B.pt + 3 * obs_point[i], p);
        if (B.chi2[i] > (obs_stereo[i] ? Chi3D : Chi2D) || !(p[2] > 0.0)) B.active[i] = 0;
    }
    B.robust = 0;
    int it2 = ba_stage(&B, 2 * n_iters, 0, trace, &nt);
    if (iters_done) { iters_done[0] = it1; iters_done[1] = it2; }
    for (int i = 0; i < n_poses; i++) {
        const se3* T = &B.pose[i];
        for (int k = 0; k < 4; k++) out_pose7[7 * i + k] = T->q[k];
        for (int k = 0; k < 3; k++) out_pose7[7 * i + 4 + k] = T->t[k];
        if (fixed[i]) { memcpy(out_pose44 + 16 * i, poses44 + 16 * i, 64); continue; }
        double R[9];
        quat_to_R(T->q, R);
        float* m = out_pose44 + 16 * i;
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) m[4 * r + c] = (float)R[3 * r + c]; m[4 * r + 3] = (float)T->t[r]; }
        m[12] = m[13] = m[14] = 0; m[15] = 1;
    }
    for (int i = 0; i < 3 * n_points; i++) out_point3[i] = B.pt[i];
    for (int i = 0; i < n_obs; i++) { /* :494-521 */
        double p[3];
        se3_map(&B.pose[obs_pose[i]], B.pt + 3 * obs_point[i], p);
        int bad = 0;
        if (obs_stereo[i]) { if (B.chi2[i] > Chi3D || !(p[2] > 0.0)) bad = 1; }
        else if (B.chi2[i] > Chi2D) bad = 1;
        out_chi2[i] = B.chi2[i];
        out_level[i] = !B.active[i];
        if (!bad) {
            const float* m = out_pose44 + 16 * obs_pose[i];
            float px = (float)B.pt[3 * obs_point[i]], py = (float)B.pt[3 * obs_point[i] + 1], pz = (float)B.pt[3 * obs_point[i] + 2];
            float zc = m[8] * px + m[9] * py + m[10] * pz + m[11];
            if (zc < 0) bad = 1;
        }
        out_bad[i] = bad;
    }
    free(B.pose); free(B.pose_bak); free(B.pt); free(B.pt_bak); free(B.free_idx); free(B.z); free(B.info); free(B.active);
    free(B.err); free(B.chi2); free(B.Hpp); free(B.bp); free(B.Hll); free(B.bl); free(B.Hpl); free(B.lm_ptr); free(B.lm_obs);
    return 0;
}

/* ======================================================================================================================
 * Pose-only optimisation — PnPSolver::solvePnp, /root/reference/src/optimization/pnpsolver.cpp:116-408:
 *   one free SE3 vertex; per match a unary edge EdgeSE3ProjectXYZOnlyPose (typesg2o.h:590-663) or
 *   EdgeStereoSE3ProjectXYZOnlyPose (:521-588) with WeightedHubberRobustKernel (:82-105: the weight scales rho only);
 *   per visible map marker a MarkerEdgeOnlyProject (:414-470; float projections, NUMERIC Jacobian with delta = 1e-4f,
 *   base_binary_edge.hpp:167-232) against a fixed marker vertex; 4 rounds of optimize(10) (minChi2BetweenIter = 0), the
 *   estimate reset to the initial pose before every round (:358), inliers re-classified after each (:364-374), kernels
 *   dropped after round index 2, early exit when < 10 inliers and no markers (:383).
 *   Unary quadratic form: base_unary_edge.hpp:50-80.  Parity pin: oracle/_ref/libref_g2o.so ref_pose_only (the reference's
 *   own edge classes + g2o), tests/test_pnp_oracle.py and tests/golden/pnp_g2o.npz.
 * ==================================================================================================================== */
typedef struct {
    int n, nm;
    const float *pts, *uv, *ur, *isig;
    const uint8_t *stereo;
    double* w;       /* kernel weight per match */
    uint8_t* robust; /* per match: kernel still attached */
    uint8_t* active; /* level 0 */
    double *err, *chi2;
    cam_t cam;
    /* markers */
    se3* g2m; double* mpts; /* nm x 4 x 3 local corner coordinates */
    const float* mobs; double wm; uint8_t* mrobust; double *merr, *mchi2;
    se3 T, Tbak;
    double H[36], b[6];
} pnp_t;

static void pnp_edge_error(const pnp_t* P, const se3* T, int i, double* e) {
    double X[3] = {P->pts[3 * i], P->pts[3 * i + 1], P->pts[3 * i + 2]}, p[3];
    se3_map(T, X, p);
    if (!P->stereo[i]) { /* typesg2o.h:640-652 */
        e[0] = (double)P->uv[2 * i] - ((p[0] / p[2]) * P->cam.fx + P->cam.cx);
        e[1] = (double)P->uv[2 * i + 1] - ((p[1] / p[2]) * P->cam.fy + P->cam.cy);
        e[2] = 0;
    } else { /* :572-580: float 1/z, bf is a double member here */
        const float invz = 1.0f / p[2];
        double r0 = p[0] * invz * P->cam.fx + P->cam.cx, r1 = p[1] * invz * P->cam.fy + P->cam.cy;
        double r2 = r0 - P->cam.bf * invz;
        e[0] = (double)P->uv[2 * i] - r0; e[1] = (double)P->uv[2 * i + 1] - r1; e[2] = (double)P->ur[i] - r2;
    }
}
static void se3_mul(const se3* A, const se3* B, se3* C) { /* se3quat.h:156-163 */
    double rt[3];
    quat_rot(A->q, B->t, rt);
    se3 R;
    quat_mul(A->q, B->q, R.q);
    for (int k = 0; k < 3; k++) R.t[k] = A->t[k] + rt[k];
    quat_normalize(R.q);
    *C = R;
}
static void pnp_marker_error(const pnp_t* P, const se3* T, int m, double* e) { /* typesg2o.h:440-468 */
    se3 C2M;
    se3_mul(T, &P->g2m[m], &C2M);
    for (int i = 0; i < 4; i++) {
        double p[3];
        se3_map(&C2M, P->mpts + 12 * m + 3 * i, p);
        float projx = (p[0] / p[2]) * P->cam.fx + P->cam.cx;
        e[2 * i] = (double)P->mobs[8 * m + 2 * i] - projx;
        float projy = (p[1] / p[2]) * P->cam.fy + P->cam.cy;
        e[2 * i + 1] = (double)P->mobs[8 * m + 2 * i + 1] - projy;
    }
}
/* computeActiveErrors + activeRobustChi2 */
static double pnp_errors(pnp_t* P) {
    double s = 0;
    for (int i = 0; i < P->n; i++) {
        if (!P->active[i]) continue;
        double* e = P->err + 3 * i;
        pnp_edge_error(P, &P->T, i, e);
        double c = P->isig[i] * (e[0] * e[0]) + P->isig[i] * (e[1] * e[1]);
        if (P->stereo[i]) c += P->isig[i] * (e[2] * e[2]);
        P->chi2[i] = c;
    }
    for (int m = 0; m < P->nm; m++) {
        double* e = P->merr + 8 * m;
        pnp_marker_error(P, &P->T, m, e);
        double c = 0;
        for (int k = 0; k < 8; k++) c += e[k] * e[k];
        P->mchi2[m] = c;
    }
    for (int i = 0; i < P->n; i++) {
        if (!P->active[i]) continue;
        if (P->robust[i]) { double r0, r1; huber(P->chi2[i], P->stereo[i] ? sqrtf(7.815f) : sqrtf(5.99f), P->w[i], &r0, &r1); s += r0; }
        else s += P->chi2[i];
    }
    for (int m = 0; m < P->nm; m++) {
        if (P->mrobust[m]) { double r0, r1; huber(P->mchi2[m], sqrtf(15.507f), P->wm, &r0, &r1); s += r0; }
        else s += P->mchi2[m];
    }
    return s;
}
static void pnp_build(pnp_t* P) {
    memset(P->H, 0, sizeof(P->H)); memset(P->b, 0, sizeof(P->b));
    for (int i = 0; i < P->n; i++) {
        if (!P->active[i]) continue;
        double X[3] = {P->pts[3 * i], P->pts[3 * i + 1], P->pts[3 * i + 2]}, p[3], J[18];
        se3_map(&P->T, X, p);
        double x = p[0], y = p[1], invz = 1.0 / p[2], invz_2 = invz * invz, fx = P->cam.fx, fy = P->cam.fy, bf = P->cam.bf;
        J[0] = x * y * invz_2 * fx; J[1] = -(1 + (x * x * invz_2)) * fx; J[2] = y * invz * fx; J[3] = -invz * fx; J[4] = 0; J[5] = x * invz_2 * fx;
        J[6] = (1 + y * y * invz_2) * fy; J[7] = -x * y * invz_2 * fy; J[8] = -x * invz * fy; J[9] = 0; J[10] = -invz * fy; J[11] = y * invz_2 * fy;
        int D = 2;
        if (P->stereo[i]) {
            D = 3;
            J[12] = J[0] - bf * y * invz_2; J[13] = J[1] + bf * x * invz_2; J[14] = J[2]; J[15] = J[3]; J[16] = 0; J[17] = J[5] - bf * invz_2;
        }
        double rho1 = 1;
        if (P->robust[i]) { double r0; huber(P->chi2[i], P->stereo[i] ? sqrtf(7.815f) : sqrtf(5.99f), P->w[i], &r0, &rho1); }
        const double* e = P->err + 3 * i;
        double om = P->isig[i];
        for (int a = 0; a < 6; a++) {
            double s = 0;
            for (int d = 0; d < D; d++) s += J[6 * d + a] * (om * e[d]);
            P->b[a] -= rho1 * s;
            for (int c = 0; c < 6; c++) {
                double h = 0;
                for (int d = 0; d < D; d++) h += J[6 * d + a] * (rho1 * om) * J[6 * d + c];
                P->H[6 * a + c] += h;
            }
        }
    }
    for (int m = 0; m < P->nm; m++) { /* numeric Jacobian wrt the camera vertex, base_binary_edge.hpp:200-226 */
        const double delta = (double)1e-4f, scalar = 1 / (2 * delta);
        double J[48];
        for (int d = 0; d < 6; d++) {
            double u[6] = {0, 0, 0, 0, 0, 0}, e1[8], e2[8];
            se3 Tp = P->T;
            u[d] = delta; se3_oplus(&Tp, u); pnp_marker_error(P, &Tp, m, e1);
            Tp = P->T;
            u[d] = -delta; se3_oplus(&Tp, u); pnp_marker_error(P, &Tp, m, e2);
            for (int k = 0; k < 8; k++) J[6 * k + d] = scalar * (e1[k] - e2[k]);
        }
        double rho1 = 1;
        if (P->mrobust[m]) { double r0; huber(P->mchi2[m], sqrtf(15.507f), P->wm, &r0, &rho1); }
        const double* e = P->merr + 8 * m;
        for (int a = 0; a < 6; a++) {
            double s = 0;
            for (int d = 0; d < 8; d++) s += J[6 * d + a] * e[d];
            P->b[a] -= rho1 * s;
            for (int c = 0; c < 6; c++) {
                double h = 0;
                for (int d = 0; d < 8; d++) h += J[6 * d + a] * rho1 * J[6 * d + c];
                P->H[6 * a + c] += h;
            }
        }
    }
}
/* SparseOptimizer::optimize(iterations) with minChi2BetweenIter = 0 */
static int pnp_stage(pnp_t* P, int iterations) {
    double lambda = 0, ni = 2;
    float prevChi2 = FLT_MAX, curChi2 = FLT_MAX, Chi2Diff = FLT_MAX;
    int ok = 1, its = 0;
    for (int it = 0; it < iterations && ok && Chi2Diff > 0.0f; it++) {
        { float t = prevChi2; prevChi2 = curChi2; curChi2 = t; }
        double currentChi = pnp_errors(P), tempChi;
        pnp_build(P);
        if (it == 0) {
            double md = 0;
            for (int j = 0; j < 6; j++) md = fmax(fabs(P->H[7 * j]), md);
            lambda = 1e-5 * md;
            ni = 2;
        }
        double rho = 0;
        int qmax = 0;
        do {
            P->Tbak = P->T;
            double A[36], x[6] = {0, 0, 0, 0, 0, 0};
            memcpy(A, P->H, sizeof(A));
            for (int j = 0; j < 6; j++) A[7 * j] += lambda;
            int ok2 = chol_solve(A, 6, P->b, x);
            if (ok2) se3_oplus(&P->T, x);
            tempChi = pnp_errors(P);
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0;
            for (int k = 0; k < 6; k++) scale += x[k] * (lambda * x[k] + P->b[k]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                P->T = P->Tbak;
                if (!isfinite(lambda)) break;
            }
            qmax++;
        } while (rho < 0 && qmax < 10);
        if (qmax == 10 || rho == 0 || !isfinite(lambda)) ok = 0;
        double last = 0; /* activeRobustChi2 over the stored (possibly stale) edge errors */
        for (int i = 0; i < P->n; i++) {
            if (!P->active[i]) continue;
            if (P->robust[i]) { double r0, r1; huber(P->chi2[i], P->stereo[i] ? sqrtf(7.815f) : sqrtf(5.99f), P->w[i], &r0, &r1); last += r0; }
            else last += P->chi2[i];
        }
        for (int m = 0; m < P->nm; m++) {
            if (P->mrobust[m]) { double r0, r1; huber(P->mchi2[m], sqrtf(15.507f), P->wm, &r0, &r1); last += r0; }
            else last += P->mchi2[m];
        }
        curChi2 = (float)last;
        Chi2Diff = prevChi2 - curChi2;
        its++;
    }
    return its;
}

/* same signature as oracle/ref_g2o_wrap.cpp: ref_pose_only */
int oracle_pose_only(const float* pose44, int n, const float* points3, const float* obs_uv, const float* obs_ur,
                     const uint8_t* obs_stereo, const float* obs_inv_sigma2, const uint8_t* stable, float fx, float fy, float cx,
                     float cy, float bf, int n_markers, const float* marker_pose44, const float* marker_size,
                     const float* marker_corners, float* out_pose44, double* out_pose7, uint8_t* bad, int* iters_done) {
    if (n == 0 && n_markers == 0) return 0;
    const float Chi2D = 5.99f, Chi3D = 7.815f, Chi8D = 15.507f;
    pnp_t P;
    memset(&P, 0, sizeof(P));
    P.n = n; P.nm = n_markers; P.pts = points3; P.uv = obs_uv; P.ur = obs_ur; P.isig = obs_inv_sigma2; P.stereo = obs_stereo;
    P.w = malloc(sizeof(double) * (n + 1)); P.robust = malloc(n + 1); P.active = malloc(n + 1);
    P.err = calloc(3 * n + 1, sizeof(double)); P.chi2 = calloc(n + 1, sizeof(double));
    P.cam.fx = fx; P.cam.fy = fy; P.cam.cx = cx; P.cam.cy = cy; P.cam.bf = bf;
    double KpWeightSum = 0;
    for (int i = 0; i < n; i++) {
        float ew = 1;
        if (!stable[i]) ew = 0.5;
        if (obs_stereo[i]) ew *= 2;
        P.w[i] = ew; P.robust[i] = 1; P.active[i] = 1;
        KpWeightSum += ew;
    }
    P.g2m = malloc(sizeof(se3) * (n_markers + 1)); P.mpts = malloc(sizeof(double) * 12 * (n_markers + 1));
    P.mrobust = malloc(n_markers + 1); P.merr = calloc(8 * n_markers + 1, sizeof(double)); P.mchi2 = calloc(n_markers + 1, sizeof(double));
    P.mobs = marker_corners;
    {
        float w_markers = 0.3;
        int total = n + n_markers;
        P.wm = ((w_markers * total) / (1. - w_markers)) / (float)KpWeightSum; /* :298-300 */
    }
    for (int m = 0; m < n_markers; m++) {
        se3_from_m44f(marker_pose44 + 16 * m, &P.g2m[m]);
        float s = marker_size[m]; /* Marker::get3DPointsLocalRefSystem, marker.cpp:58-62: cv::Point3f of size/2. */
        float h = (float)(s / 2.), mh = (float)(-s / 2.);
        double c[12] = {mh, h, 0, h, h, 0, h, mh, 0, mh, mh, 0};
        memcpy(P.mpts + 12 * m, c, sizeof(c));
        P.mrobust[m] = 1;
    }
    se3 T0;
    se3_from_m44f(pose44, &T0);
    uint8_t* vbad = calloc(n + 1, 1);
    for (int it = 0; it < 4; it++) if (iters_done) iters_done[it] = 0;
    for (int it = 0; it < 4; it++) {
        P.T = T0;
        int r = pnp_stage(&P, 10);
        if (iters_done) iters_done[it] = r;
        int good = 0;
        for (int i = 0; i < n; i++) {
            if (vbad[i]) { /* e->computeError() on an inactive edge */
                double* e = P.err + 3 * i;
                pnp_edge_error(&P, &P.T, i, e);
                double c = P.isig[i] * (e[0] * e[0]) + P.isig[i] * (e[1] * e[1]);
                if (obs_stereo[i]) c += P.isig[i] * (e[2] * e[2]);
                P.chi2[i] = c;
            }
            vbad[i] = P.chi2[i] > (obs_stereo[i] ? Chi3D : Chi2D);
            P.active[i] = !vbad[i];
            if (it >= 2) P.robust[i] = 0;
            if (!vbad[i]) good++;
        }
        for (int m = 0; m < n_markers; m++) {
            double* e = P.merr + 8 * m;
            pnp_marker_error(&P, &P.T, m, e);
            double c = 0;
            for (int k = 0; k < 8; k++) c += e[k] * e[k];
            P.mchi2[m] = c;
            if (c > Chi8D || it >= 2) P.mrobust[m] = 0;
        }
        if (good < 10 && n_markers == 0) break;
    }
    {
        double R[9];
        quat_to_R(P.T.q, R);
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) out_pose44[4 * r + c] = (float)R[3 * r + c]; out_pose44[4 * r + 3] = (float)P.T.t[r]; }
        out_pose44[12] = out_pose44[13] = out_pose44[14] = 0; out_pose44[15] = 1;
        for (int k = 0; k < 4; k++) out_pose7[k] = P.T.q[k];
        for (int k = 0; k < 3; k++) out_pose7[4 + k] = P.T.t[k];
    }
    int nbad = 0;
    for (int i = 0; i < n; i++) { bad[i] = vbad[i]; nbad += vbad[i]; }
    free(P.w); free(P.robust); free(P.active); free(P.err); free(P.chi2); free(P.g2m); free(P.mpts); free(P.mrobust); free(P.merr);
    free(P.mchi2); free(vbad);
    return n - nbad;
}
