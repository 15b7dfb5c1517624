// ba.cu — K10-K13: bundle adjustment on the device (f64): residuals, per-observation Jacobians and quadratic forms,
// block-sparse Schur complement, dense reduced-camera solve, landmark back-substitution, Levenberg-Marquardt control.
//
// What it replaces (see include/ucoslam_b200.h for the entry point): GlobalOptimizerG2O's graph + g2o's
// OptimizationAlgorithmLevenberg / BlockSolver_6_3 / LinearSolverEigen for that graph.
//
// Layout in HBM (everything f64 unless noted; one arena per solve, carved from a grow-only context workspace):
//   pose[P x 7] (qx qy qz qw tx ty tz) + backup        pt[N x 3] + backup
//   observations SORTED BY LANDMARK: obs_pose[M] i32, obs_lm[M] i32, z[M x 3], info[M], stereo[M] u8, active[M] u8,
//                lm_ptr[N+1]; per free pose the list of its observations (pose_ptr[Pf+1], pose_obs[])
//   per observation: err[M x 3], chi2[M], rho0[M], Hpl[M x 18] (6x3 block W of the (pose, landmark) pair), Y[M x 18] = W D^-1
//   per landmark: Hll[N x 6] (symmetric packed), bl[N x 3], Dinv[N x 6], db[N x 3] = D^-1 bl, xl[N x 3]
//   per free pose: Hpp[Pf x 36], bp[Pf x 6];  reduced system S[n x n] (n = 6 Pf, dense), bs[n], xp[n]
//   Schur gather lists (built on the host once per solve): for every non-zero 6x6 block (i <= j) of S the list of
//   (obs_a, obs_b) pairs of landmarks seen by both poses, in landmark order -> every sum has a FIXED order: no atomics,
//   results are bitwise reproducible run to run.
// Kernels per LM trial: ba_prep (D^-1, Y, db) -> ba_schur_gather (S, bs) -> ba_chol_solve -> ba_update (back-substitution,
// backup, oplus) -> ba_errors -> ba_decide (ordered chi2 / scale reductions, gain ratio, lambda schedule, accept / restore).
// The LM state machine lives on the device (LmState); the host only reads two flags per trial to know what to enqueue next.
#include "common.cuh"
#include "ba_math.cuh"
#include "ba_plan.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>
#include <chrono>
#include <mutex>
#include "ba_band.h"

namespace {
using namespace ba;

struct LmState {
    double lambda, ni, currentChi, tempChi, rho, scale;
    float prevChi2, curChi2, chi2Diff;
    int it;          // outer iterations finished in this stage
    int qmax;        // LM trials of the current iteration
    int ok;          // SolverResult == OK
    int cont_trial;  // the do-while of OptimizationAlgorithmLevenberg::solve goes on
    int cont_iter;   // the for loop of SparseOptimizer::optimize goes on (before the iteration-count test)
    int chol_fail;
    int ntrace;
    double trace[128];
};

struct MkDev {  // ArUco markers: free SE3 vertices appended after the free keyframes (replicated on every rank), binary MarkerEdges
    int Nm = 0, Ne = 0;
    double *pose = nullptr, *pose_bak = nullptr;              // Nm x 7, global <- marker
    const float* size = nullptr;                              // Nm
    const int *e_marker = nullptr, *e_pose = nullptr;         // Ne
    const float *e_corners = nullptr, *e_weight = nullptr;    // Ne x 8, Ne
    double* e_chi2 = nullptr;                                 // Ne
    double* e_blk = nullptr;                                  // Ne x 120: Hcc(36) Hmm(36) Hcm(36) bc(6) bm(6)
    const int *cam_ptr = nullptr, *cam_edges = nullptr;       // free keyframe -> its marker edges (edge order)
    const int *mk_ptr = nullptr, *mk_edges = nullptr;         // marker -> its edges
    const int* blk_edge = nullptr;                            // Schur block -> marker edge whose Hcm fills it, or -1
    // InPlaneMarkers (globaloptimizer_g2o.cpp:356-401): Nx planar edges (reference marker, other marker) stored after the Ne marker edges in
    // e_chi2 / e_blk; x_ref < 0: the reference marker is a fixed vertex outside this window (its pose sits in slot Nm of `pose`).  In an edge's block record the "c"
    // slots belong to the marker with the smaller index, the "m" slots to the other one; a marker's edge list holds ~index for its "c" role.
    int Nx = 0, x_ref = -1;
    const int* x_other = nullptr;                             // Nx
    double x_w = 0;                                           // information = I4 * x_w
};

struct BaDev {
    MkDev mk;
    int P, N, M, Pf, n, nblk;
    double *pose, *pose_bak, *pt, *pt_bak;
    const int *free_idx, *free_list, *lm_ptr, *obs_pose, *obs_lm, *pose_ptr, *pose_obs;
    const double *z, *info;
    const uint8_t* stereo;
    uint8_t* active;
    double *err, *chi2, *rho0, *Hll, *bl, *Hpl, *Y, *Dinv, *db, *xl, *Hpp, *bp, *S, *bs, *xp, *scale_lm, *scale_pose;
    const int* blk_ptr;
    const int2 *blk_ij, *con;
    LmState* st;
    Cam cam;
    const Cam* cams = nullptr;   // per keyframe (uco_ba_problem::pose_cam: a window whose keyframes were taken with different cameras), else cam
    double d2, d3;      // Huber deltas sqrt(5.99f), sqrt(7.815f) (globaloptimizer_g2o.h:112-117)
    float chi2d, chi3d;
};

// every edge carries the ImageParams of ITS keyframe (globaloptimizer_g2o.cpp:233-236, :262-266, :335-338)
__device__ __forceinline__ const Cam& cam_of(const BaDev& B, int pi) { return B.cams ? B.cams[pi] : B.cam; }

// ---- K10 residuals: computeActiveErrors + the per-edge terms of activeRobustChi2 (sparse_optimizer.cpp:102-116) -------------
__global__ void __launch_bounds__(256) ba_errors_kernel(const __grid_constant__ BaDev B, int robust) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.M) return;
    if (!B.active[i]) {  // level-1 edges are not evaluated any more: they keep their last error (g2o) and add nothing
        B.rho0[i] = 0;
        return;
    }
    Pose T = load_pose(B.pose + 7 * B.obs_pose[i]);
    const double* X = B.pt + 3 * B.obs_lm[i];
    double x[3] = {X[0], X[1], X[2]}, p[3], e[3];
    se3_map(T, x, p);
    bool st = B.stereo[i];
    double z[3] = {B.z[3 * i], B.z[3 * i + 1], B.z[3 * i + 2]};
    residual(p, z, st, cam_of(B, B.obs_pose[i]), e);
    double c2 = st ? (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * B.info[i] : (e[0] * e[0] + e[1] * e[1]) * B.info[i];
    B.err[3 * i] = e[0]; B.err[3 * i + 1] = e[1]; B.err[3 * i + 2] = e[2];
    B.chi2[i] = c2;
    double r0 = c2, r1;
    if (robust) huber(c2, st ? B.d3 : B.d2, 1.0, r0, r1);
    B.rho0[i] = r0;
}

// weights of one observation's quadratic form (base_binary_edge.hpp:104-151): wo = rho1 * information, orr = -(information e) rho1
__device__ __forceinline__ void obs_weights(const BaDev& B, int i, int robust, double& wo, double* orr) {
    double w = B.info[i], r1 = 1, r0;
    if (robust) huber(B.chi2[i], B.stereo[i] ? B.d3 : B.d2, 1.0, r0, r1);
    wo = r1 * w;
#pragma unroll
    for (int d = 0; d < 3; d++) orr[d] = -(w * B.err[3 * i + d]) * r1;
}

// ---- K11a linearize, landmark side: one thread per landmark walks its observations: Hll, bl, and the W block of each ---------
__global__ void __launch_bounds__(128) ba_linearize_lm_kernel(const __grid_constant__ BaDev B, int robust) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= B.N) return;
    double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    double x[3] = {B.pt[3 * l], B.pt[3 * l + 1], B.pt[3 * l + 2]};
    for (int i = B.lm_ptr[l]; i < B.lm_ptr[l + 1]; i++) {
        double* W = B.Hpl + 18 * (size_t)i;
        if (!B.active[i]) {
#pragma unroll
            for (int k = 0; k < 18; k++) W[k] = 0;
            continue;
        }
        int pi = B.obs_pose[i];
        Pose T = load_pose(B.pose + 7 * pi);
        double p[3], R[9], JX[9], JT[18], wo, orr[3];
        se3_map(T, x, p);
        quat_to_R(T.q, R);
        bool st = B.stereo[i];
        jac_point(p, R, st, cam_of(B, pi), JX);
        obs_weights(B, i, robust, wo, orr);
        const int D = st ? 3 : 2;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double s = 0;
            for (int d = 0; d < D; d++) s += JX[3 * d + a] * orr[d];
            b[a] += s;
        }
        {
            int k = 0;
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int c = a; c < 3; c++, k++) {
                    double h = 0;
                    for (int d = 0; d < D; d++) h += JX[3 * d + a] * wo * JX[3 * d + c];
                    H[k] += h;
                }
        }
        if (B.free_idx[pi] >= 0) {
            jac_pose(p, st, cam_of(B, pi), JT);
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    double h = 0;
                    for (int d = 0; d < D; d++) h += JT[6 * d + a] * wo * JX[3 * d + c];
                    W[3 * a + c] = h;
                }
        } else {
#pragma unroll
            for (int k = 0; k < 18; k++) W[k] = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) B.Hll[6 * (size_t)l + k] = H[k];
#pragma unroll
    for (int k = 0; k < 3; k++) B.bl[3 * (size_t)l + k] = b[k];
}

// ---- K11b linearize, pose side: one CTA per free pose; ordered (strided + tree) reduction of J^T w J and J^T w r ----------------
constexpr int POSE_THREADS = 128;
__global__ void __launch_bounds__(POSE_THREADS) ba_linearize_pose_kernel(const __grid_constant__ BaDev B, int robust) {
    int f = blockIdx.x;
    int pi = B.free_list[f];
    Pose T = load_pose(B.pose + 7 * pi);
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
    for (int j = B.pose_ptr[f] + threadIdx.x; j < B.pose_ptr[f + 1]; j += POSE_THREADS) {
        int i = B.pose_obs[j];
        if (!B.active[i]) continue;
        const double* X = B.pt + 3 * B.obs_lm[i];
        double x[3] = {X[0], X[1], X[2]}, p[3], JT[18], wo, orr[3];
        se3_map(T, x, p);
        bool st = B.stereo[i];
        jac_pose(p, st, cam_of(B, pi), JT);
        obs_weights(B, i, robust, wo, orr);
        const int D = st ? 3 : 2;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = a; c < 6; c++, k++) {
                double h = 0;
                for (int d = 0; d < D; d++) h += JT[6 * d + a] * wo * JT[6 * d + c];
                acc[k] += h;
            }
#pragma unroll
        for (int a = 0; a < 6; a++) {
            double s = 0;
            for (int d = 0; d < D; d++) s += JT[6 * d + a] * orr[d];
            acc[21 + a] += s;
        }
    }
    __shared__ double red[POSE_THREADS / 32][27];
#pragma unroll
    for (int k = 0; k < 27; k++) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 27) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < POSE_THREADS / 32; w++) v += red[w][threadIdx.x];
        red[0][threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < 36) {
        int a = threadIdx.x / 6, c = threadIdx.x % 6;
        int lo = a < c ? a : c, hi = a < c ? c : a;
        int k = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
        B.Hpp[36 * (size_t)f + threadIdx.x] = red[0][k];
    }
    if (threadIdx.x < 6) B.bp[6 * (size_t)f + threadIdx.x] = red[0][21 + threadIdx.x];
}

// block-wide ordered sum / max over a strided sequence (blockDim.x == 1024)
template <bool MAX>
__device__ double block_reduce_1024(double v, double* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = MAX ? fmax(v, w) : v + w;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = sm[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double w = __shfl_xor_sync(0xffffffffu, v, o);
            v = MAX ? fmax(v, w) : v + w;
        }
        if (threadIdx.x == 0) sm[32] = v;
    }
    __syncthreads();
    return sm[32];
}

// ---- LM control, start of a stage / of an outer iteration (sparse_optimizer.cpp:366-381, levenberg.cpp:58-96,152-166) ----------
// ext (sharded form): the sums over ALL ranks' observations / landmarks, already all-reduced: ext[0] = robust chi2,
// ext[1] = landmark part of computeScale, ext[2] = max |diagonal|, ext[3] = number of ranks whose stop flag is raised
__global__ void __launch_bounds__(1024) ba_iter_begin_kernel(const __grid_constant__ BaDev B, int first_of_stage, const double* ext = nullptr) {
    __shared__ double sm[33];
    LmState* st = B.st;
    if (first_of_stage) {
        double s = 0, md = 0;
        if (ext) {
            s = ext[0];
            md = ext[2];
        } else {
            for (int i = threadIdx.x; i < B.M; i += 1024) s += B.rho0[i];
            s = block_reduce_1024<false>(s, sm);
            // computeLambdaInit: tau * max |diagonal| over all active vertices
            for (int k = threadIdx.x; k < 6 * B.Pf; k += 1024) md = fmax(md, fabs(B.Hpp[36 * (size_t)(k / 6) + 7 * (k % 6)]));
            for (int k = threadIdx.x; k < 3 * B.N; k += 1024) {
                int l = k / 3, j = k % 3;
                md = fmax(md, fabs(B.Hll[6 * (size_t)l + (j == 0 ? 0 : (j == 1 ? 3 : 5))]));
            }
            md = block_reduce_1024<true>(md, sm);
        }
        if (threadIdx.x == 0) {
            st->currentChi = s;
            st->lambda = 1e-5 * md;
            st->ni = 2;
            st->prevChi2 = st->curChi2 = st->chi2Diff = FLT_MAX;
            st->it = 0;
            st->ok = 1;
            st->cont_iter = 1;
        }
    }
    if (threadIdx.x == 0) {
        float t = st->prevChi2;
        st->prevChi2 = st->curChi2;
        st->curChi2 = t;
        st->qmax = 0;
        st->rho = 0;
        st->cont_trial = 1;
    }
}

// ---- K12a per landmark: D^-1 = (Hll + lambda I)^-1, db = D^-1 bl, Y = W D^-1 for each of its observations --------------------
__global__ void __launch_bounds__(128) ba_prep_kernel(const __grid_constant__ BaDev B) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= B.N) return;
    const double lambda = B.st->lambda;
    double D[6], I[6];
#pragma unroll
    for (int k = 0; k < 6; k++) D[k] = B.Hll[6 * (size_t)l + k];
    D[0] += lambda; D[3] += lambda; D[5] += lambda;
    inv3_sym(D, I);
#pragma unroll
    for (int k = 0; k < 6; k++) B.Dinv[6 * (size_t)l + k] = I[k];
    const double Im[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]};
    double b[3] = {B.bl[3 * (size_t)l], B.bl[3 * (size_t)l + 1], B.bl[3 * (size_t)l + 2]};
#pragma unroll
    for (int a = 0; a < 3; a++) B.db[3 * (size_t)l + a] = Im[3 * a] * b[0] + Im[3 * a + 1] * b[1] + Im[3 * a + 2] * b[2];
    for (int i = B.lm_ptr[l]; i < B.lm_ptr[l + 1]; i++) {
        const double* W = B.Hpl + 18 * (size_t)i;
        double* Y = B.Y + 18 * (size_t)i;
#pragma unroll
        for (int a = 0; a < 6; a++) {
            double w0 = W[3 * a], w1 = W[3 * a + 1], w2 = W[3 * a + 2];
#pragma unroll
            for (int c = 0; c < 3; c++) Y[3 * a + c] = w0 * Im[c] + w1 * Im[3 + c] + w2 * Im[6 + c];
        }
    }
}

// ---- K12b Schur complement by gather (block_solver.hpp:329-400): one CTA per non-zero 6x6 block (i <= j) of S ----------------
//   S(i,j) = [i == j] (Hpp(i) + lambda I) - sum over shared landmarks of Y_a W_b^T ;  bs(i) = bp(i) - sum over its observations W_a db
constexpr int GATHER_CHUNKS = 8;
__global__ void __launch_bounds__(36 * GATHER_CHUNKS) ba_schur_gather_kernel(const __grid_constant__ BaDev B) {
    const int blk = blockIdx.x;
    const int2 ij = B.blk_ij[blk];
    const int beg = B.blk_ptr[blk], end = B.blk_ptr[blk + 1];
    const int e = threadIdx.x % 36, chunk = threadIdx.x / 36, r = e / 6, c = e % 6;
    const int len = end - beg, per = (len + GATHER_CHUNKS - 1) / GATHER_CHUNKS;
    const int c0 = beg + chunk * per, c1 = min(end, c0 + per);
    const bool diag = ij.x == ij.y;
    double s = 0, sb = 0;
    for (int k = c0; k < c1; k++) {
        const int2 ab = B.con[k];
        const double* Y = B.Y + 18 * (size_t)ab.x + 3 * r;
        const double* W = B.Hpl + 18 * (size_t)ab.y + 3 * c;
        s += Y[0] * W[0] + Y[1] * W[1] + Y[2] * W[2];
        if (diag && c == 0) {
            const double* Wa = B.Hpl + 18 * (size_t)ab.x + 3 * r;
            const double* db = B.db + 3 * (size_t)B.obs_lm[ab.x];
            sb += Wa[0] * db[0] + Wa[1] * db[1] + Wa[2] * db[2];
        }
    }
    __shared__ double red[GATHER_CHUNKS][36], redb[GATHER_CHUNKS][6];
    red[chunk][e] = s;
    if (c == 0) redb[chunk][r] = sb;
    __syncthreads();
    if (chunk == 0) {
        double t = 0;
#pragma unroll
        for (int k = 0; k < GATHER_CHUNKS; k++) t += red[k][e];
        double h = 0;
        if (diag) {
            h = B.Hpp[36 * (size_t)ij.x + e];
            if (r == c) h += B.st->lambda;
        }
        h -= t;
        const int n = B.n;
        B.S[(size_t)(6 * ij.x + r) * n + 6 * ij.y + c] = h;
        if (!diag) B.S[(size_t)(6 * ij.y + c) * n + 6 * ij.x + r] = h;
        if (diag && c == 0) {
            double tb = 0;
#pragma unroll
            for (int k = 0; k < GATHER_CHUNKS; k++) tb += redb[k][r];
            B.bs[6 * ij.x + r] = B.bp[6 * ij.x + r] - tb;
        }
    }
}

// ---- K13 dense reduced-camera solve S xp = bs (replaces LinearSolverEigen / SimplicialLDLT, linear_solver_eigen.h:92-123) --------
// One CTA; the lower triangle of S plus bs as an extra row is held packed (row i at i(i+1)/2) in shared memory when it fits
// (n <= 238), else in a global workspace.  Right-looking L D L^T: after eliminating column j the extra row holds w = L^-1 bs,
// then one warp back-substitutes L^T x = D^-1 w.  A non-positive pivot raises chol_fail (the trial is then rejected).
constexpr int CHOL_THREADS = 1024;
__global__ void __launch_bounds__(CHOL_THREADS) ba_chol_solve_kernel(const __grid_constant__ BaDev B, double* gws, int use_smem) {
    extern __shared__ double sm_dyn[];
    __shared__ double col[1024];  // current column (chunks of up to 1024 rows at a time are not needed: n <= 1023 for this kernel)
    __shared__ int fail;
    const int n = B.n, tid = threadIdx.x;
    double* A = use_smem ? sm_dyn : gws;
    if (tid == 0) fail = 0;
    // load lower triangle (+ rhs row n)
    for (int i = tid / 32; i <= n; i += CHOL_THREADS / 32) {
        double* row = A + (size_t)i * (i + 1) / 2;
        if (i < n)
            for (int k = tid & 31; k <= i; k += 32) row[k] = B.S[(size_t)i * n + k];
        else
            for (int k = tid & 31; k < n; k += 32) row[k] = B.bs[k];
    }
    __syncthreads();
    for (int j = 0; j < n; j++) {
        const double p = A[(size_t)j * (j + 1) / 2 + j];
        if (!(p > 0) || !isfinite(p)) {
            if (tid == 0) fail = 1;
            break;  // uniform: every thread reads the same p
        }
        const int m = n - j;  // rows j+1 .. n
        for (int t = tid; t < m; t += CHOL_THREADS) col[t] = A[(size_t)(j + 1 + t) * (j + 2 + t) / 2 + j];
        __syncthreads();
        const double inv = 1.0 / p;
        for (int i = j + 1 + tid / 32; i <= n; i += CHOL_THREADS / 32) {
            double* row = A + (size_t)i * (i + 1) / 2;
            const double li = col[i - j - 1] * inv;
            const int kmax = i < n ? i : n - 1;
            for (int k = j + 1 + (tid & 31); k <= kmax; k += 32) row[k] -= li * col[k - j - 1];
        }
        __syncthreads();
    }
    __syncthreads();
    if (fail) {
        if (tid == 0) B.st->chol_fail = 1;
        for (int k = tid; k < n; k += CHOL_THREADS) B.xp[k] = 0;
        return;
    }
    if (tid == 0) B.st->chol_fail = 0;
    if (tid < 32) {  // back-substitution: x_j = (w_j - sum_{i>j} A[i][j] x_i) / d_j ; s (in col[]) carries w minus the known terms
        double* wrow = A + (size_t)n * (n + 1) / 2;
        for (int k = tid; k < n; k += 32) col[k] = wrow[k];
        __syncwarp();
        for (int j = n - 1; j >= 0; j--) {
            const double* row = A + (size_t)j * (j + 1) / 2;
            const double xj = col[j] / row[j];
            __syncwarp();
            if (tid == 0) col[j] = xj;
            for (int k = tid; k < j; k += 32) col[k] -= row[k] * xj;
            __syncwarp();
        }
        for (int k = tid; k < n; k += 32) B.xp[k] = col[k];
    }
}

// ---- K12c update: landmark back-substitution xl = D^-1 (bl - W^T xp) (block_solver.hpp:413-443), backup (push), oplus ----------
__global__ void __launch_bounds__(128) ba_update_kernel(const __grid_constant__ BaDev B) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    const LmState* st = B.st;
    const bool apply = !st->chol_fail;
    const double lambda = st->lambda;
    if (t < B.N) {
        const int l = t;
        double c[3] = {B.bl[3 * (size_t)l], B.bl[3 * (size_t)l + 1], B.bl[3 * (size_t)l + 2]};
        for (int i = B.lm_ptr[l]; i < B.lm_ptr[l + 1]; i++) {
            int f = B.free_idx[B.obs_pose[i]];
            if (f < 0 || !B.active[i]) continue;
            const double* W = B.Hpl + 18 * (size_t)i;
            const double* x = B.xp + 6 * f;
#pragma unroll
            for (int b = 0; b < 3; b++) {
                double s = 0;
#pragma unroll
                for (int a = 0; a < 6; a++) s += W[3 * a + b] * x[a];
                c[b] -= s;
            }
        }
        const double* I = B.Dinv + 6 * (size_t)l;
        double xl[3] = {I[0] * c[0] + I[1] * c[1] + I[2] * c[2], I[1] * c[0] + I[3] * c[1] + I[4] * c[2],
                        I[2] * c[0] + I[4] * c[1] + I[5] * c[2]};
        double sc = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double v = B.pt[3 * (size_t)l + k];
            B.pt_bak[3 * (size_t)l + k] = v;
            if (apply) B.pt[3 * (size_t)l + k] = v + xl[k];
            B.xl[3 * (size_t)l + k] = xl[k];
            sc += xl[k] * (lambda * xl[k] + B.bl[3 * (size_t)l + k]);
        }
        B.scale_lm[l] = sc;  // computeScale terms (levenberg.cpp:168-175)
    } else if (t < B.N + B.Pf) {
        const int f = t - B.N, pi = B.free_list[f];
        Pose T = load_pose(B.pose + 7 * pi);
        store_pose(B.pose_bak + 7 * pi, T);
        double u[6], sc = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            u[k] = B.xp[6 * f + k];
            sc += u[k] * (lambda * u[k] + B.bp[6 * f + k]);
        }
        B.scale_pose[f] = sc;
        if (apply) {
            se3_oplus(T, u);
            store_pose(B.pose + 7 * pi, T);
        }
    }
}

// ---- LM control, end of a trial (levenberg.cpp:96-150) and, when the trial loop ends, of the iteration (sparse_optimizer.cpp:403-436)
__global__ void __launch_bounds__(1024) ba_decide_kernel(const __grid_constant__ BaDev B, int stop, int max_iters, const double* ext = nullptr) {
    __shared__ double sm[33];
    __shared__ int reject;
    LmState* st = B.st;
    double s = 0;
    if (!ext)
        for (int i = threadIdx.x; i < B.M; i += 1024) s += B.rho0[i];
    const double chi_raw = ext ? ext[0] : block_reduce_1024<false>(s, sm);
    double sc = 0;
    for (int f = threadIdx.x; f < B.Pf + B.mk.Nm; f += 1024) sc += B.scale_pose[f];
    const double sc_p = block_reduce_1024<false>(sc, sm);
    sc = 0;
    if (!ext)
        for (int l = threadIdx.x; l < B.N; l += 1024) sc += B.scale_lm[l];
    const double sc_l = ext ? ext[1] : block_reduce_1024<false>(sc, sm);
    if (ext) stop = ext[3] != 0.0;
    if (threadIdx.x == 0) {
        double tempChi = st->chol_fail ? DBL_MAX : chi_raw;
        double rho = st->currentChi - tempChi;
        double scale = sc_p + sc_l + 1e-3;
        rho /= scale;
        int rej = 0, lam_bad = 0;
        if (rho > 0 && isfinite(tempChi) && !st->chol_fail) {
            double alpha = 1. - pow((2 * rho - 1), 3.0);
            alpha = fmin(alpha, 2. / 3.);
            double sf = fmax(1. / 3., alpha);
            st->lambda *= sf;
            st->ni = 2;
            st->currentChi = tempChi;
        } else {
            st->lambda *= st->ni;
            st->ni *= 2;
            rej = 1;
            if (!isfinite(st->lambda)) lam_bad = 1;
        }
        if (!lam_bad) st->qmax++;
        st->rho = rho;
        st->tempChi = tempChi;
        st->scale = scale;
        reject = rej;
        int cont = !lam_bad && rho < 0 && st->qmax < 10 && !stop;
        st->cont_trial = cont;
        if (!cont) {
            if (st->qmax == 10 || rho == 0 || !isfinite(st->lambda)) st->ok = 0;
            st->curChi2 = (float)chi_raw;
            st->chi2Diff = st->prevChi2 - st->curChi2;
            if (st->ntrace < 64) {
                st->trace[2 * st->ntrace] = chi_raw;
                st->trace[2 * st->ntrace + 1] = st->qmax;
                st->ntrace++;
            }
            st->it++;
            st->cont_iter = st->it < max_iters && !stop && st->ok && st->chi2Diff > 1.0f;
        }
    }
    __syncthreads();
    if (reject) {  // pop: restore the estimates saved before the update
        for (int k = threadIdx.x; k < 3 * B.N; k += 1024) B.pt[k] = B.pt_bak[k];
        for (int k = threadIdx.x; k < 7 * B.Pf; k += 1024) {
            int pi = B.free_list[k / 7];
            B.pose[7 * pi + k % 7] = B.pose_bak[7 * pi + k % 7];
        }
        for (int k = threadIdx.x; k < 7 * B.mk.Nm; k += 1024) B.mk.pose[k] = B.mk.pose_bak[k];
    }
}

// ---- between the stages (globaloptimizer_g2o.cpp:432-449): edges with chi2 above the gate or non-positive depth leave (level 1)
__global__ void __launch_bounds__(256) ba_flag_outliers_kernel(const __grid_constant__ BaDev B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.M) return;
    Pose T = load_pose(B.pose + 7 * B.obs_pose[i]);
    const double* X = B.pt + 3 * B.obs_lm[i];
    double x[3] = {X[0], X[1], X[2]}, p[3];
    se3_map(T, x, p);
    if (B.chi2[i] > (double)(B.stereo[i] ? B.chi3d : B.chi2d) || !(p[2] > 0.0)) B.active[i] = 0;
}

// ---- getResults (globaloptimizer_g2o.cpp:466-521): poses as f32 4x4, bad associations ---------------------------------------------
__global__ void __launch_bounds__(256) ba_results_kernel(const __grid_constant__ BaDev B, const float* poses44_in, float* poses44_out,
                                                         uint8_t* bad) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B.P) {
        float* m = poses44_out + 16 * t;
        if (B.free_idx[t] < 0) {
            for (int k = 0; k < 16; k++) m[k] = poses44_in[16 * t + k];
        } else {
            Pose T = load_pose(B.pose + 7 * t);
            double R[9];
            quat_to_R(T.q, R);
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++) m[4 * r + c] = (float)R[3 * r + c];
                m[4 * r + 3] = (float)T.t[r];
            }
            m[12] = m[13] = m[14] = 0;
            m[15] = 1;
        }
    }
}
__global__ void __launch_bounds__(256) ba_bad_kernel(const __grid_constant__ BaDev B, const float* poses44_out, uint8_t* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.M) return;
    Pose T = load_pose(B.pose + 7 * B.obs_pose[i]);
    const double* X = B.pt + 3 * B.obs_lm[i];
    double x[3] = {X[0], X[1], X[2]}, p[3];
    se3_map(T, x, p);
    int b = 0;
    if (B.stereo[i]) {
        if (B.chi2[i] > (double)B.chi3d || !(p[2] > 0.0)) b = 1;
    } else if (B.chi2[i] > (double)B.chi2d) b = 1;
    if (!b) {  // pincam = pose_f2g (f32) * point (f32), z < 0
        const float* m = poses44_out + 16 * B.obs_pose[i];
        float px = (float)x[0], py = (float)x[1], pz = (float)x[2];
        float zc = m[8] * px + m[9] * py + m[10] * pz + m[11];
        if (zc < 0) b = 1;
    }
    bad[i] = b;
}

__global__ void __launch_bounds__(128) ba_init_poses_kernel(const __grid_constant__ BaDev B, const float* poses44) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B.P) return;
    Pose T = pose_from_m44f(poses44 + 16 * t);
    store_pose(B.pose + 7 * t, T);
    store_pose(B.pose_bak + 7 * t, T);
}


// ---- sharded / large form (BASELINE config 5): the landmarks (with all their observations) are partitioned over the ranks, so
// every Hll / W / Y block lives on one GPU; each rank builds the partial Hpp / bp of its observations and the partial Schur
// complement of its landmarks as PACKED blocks (the block list is global: same layout on every rank), one all-reduce per LM
// trial sums them over NVLink, and every rank assembles and solves the reduced system redundantly (block_solver.hpp:329-400
// split by landmark; the sum over landmarks is associative up to rounding, see DESIGN.md for the tolerance).
__global__ void __launch_bounds__(36 * GATHER_CHUNKS) ba_schur_gather_packed_kernel(const __grid_constant__ BaDev B, double* __restrict__ Sp,
                                                                                    double* __restrict__ bsp) {
    const int blk = blockIdx.x;
    const int2 ij = B.blk_ij[blk];
    const int beg = B.blk_ptr[blk], end = B.blk_ptr[blk + 1];
    const int e = threadIdx.x % 36, chunk = threadIdx.x / 36, r = e / 6, c = e % 6;
    const int len = end - beg, per = (len + GATHER_CHUNKS - 1) / GATHER_CHUNKS;
    const int c0 = beg + chunk * per, c1 = min(end, c0 + per);
    const bool diag = ij.x == ij.y;
    double s = 0, sb = 0;
    for (int k = c0; k < c1; k++) {
        const int2 ab = B.con[k];
        const double* Y = B.Y + 18 * (size_t)ab.x + 3 * r;
        const double* W = B.Hpl + 18 * (size_t)ab.y + 3 * c;
        s += Y[0] * W[0] + Y[1] * W[1] + Y[2] * W[2];
        if (diag && c == 0) {
            const double* Wa = B.Hpl + 18 * (size_t)ab.x + 3 * r;
            const double* db = B.db + 3 * (size_t)B.obs_lm[ab.x];
            sb += Wa[0] * db[0] + Wa[1] * db[1] + Wa[2] * db[2];
        }
    }
    __shared__ double red[GATHER_CHUNKS][36], redb[GATHER_CHUNKS][6];
    red[chunk][e] = s;
    if (c == 0) redb[chunk][r] = sb;
    __syncthreads();
    if (chunk == 0) {
        double t = 0;
#pragma unroll
        for (int k = 0; k < GATHER_CHUNKS; k++) t += red[k][e];
        Sp[36 * (size_t)blk + e] = t;
        if (diag && c == 0) {
            double tb = 0;
#pragma unroll
            for (int k = 0; k < GATHER_CHUNKS; k++) tb += redb[k][r];
            bsp[6 * ij.x + r] = tb;
        }
    }
}

// S = [i == j](Hpp + lambda I) - (sum over all ranks of the packed Schur blocks), bs = bp - (summed right-hand side); S was zeroed
__global__ void __launch_bounds__(36) ba_assemble_kernel(const __grid_constant__ BaDev B, const double* __restrict__ Sp, const double* __restrict__ bsp) {
    const int blk = blockIdx.x, e = threadIdx.x, r = e / 6, c = e % 6;
    const int2 ij = B.blk_ij[blk];
    const bool diag = ij.x == ij.y;
    double h = 0;
    if (diag) {
        h = B.Hpp[36 * (size_t)ij.x + e];
        if (r == c) h += B.st->lambda;
    }
    h -= Sp[36 * (size_t)blk + e];
    if (B.mk.blk_edge) {  // (keyframe, marker) block: J_c^T Omega J_m of that marker edge
        const int ed = B.mk.blk_edge[blk];
        if (ed >= 0) h += B.mk.e_blk[120 * (size_t)ed + 72 + e];
    }
    const int n = B.n;
    B.S[(size_t)(6 * ij.x + r) * n + 6 * ij.y + c] = h;
    if (!diag) B.S[(size_t)(6 * ij.y + c) * n + 6 * ij.x + r] = h;
    if (diag && c == 0) B.bs[6 * ij.x + r] = B.bp[6 * ij.x + r] - bsp[6 * ij.x + r];
}


// ---- ArUco markers (globaloptimizer_g2o.cpp:304-350; typesg2o.h:108-167 MarkerEdge) ----------------------------------------------
__device__ __forceinline__ Pose pose_compose(const Pose& a, const Pose& b) {  // SE3Quat::operator*, se3quat.h:156-163
    Pose r;
    double rt[3];
    quat_rot(a.q, b.t, rt);
    const double *x = a.q, *y = b.q;
    r.q[3] = x[3] * y[3] - x[0] * y[0] - x[1] * y[1] - x[2] * y[2];
    r.q[0] = x[3] * y[0] + x[0] * y[3] + x[1] * y[2] - x[2] * y[1];
    r.q[1] = x[3] * y[1] + x[1] * y[3] + x[2] * y[0] - x[0] * y[2];
    r.q[2] = x[3] * y[2] + x[2] * y[3] + x[0] * y[1] - x[1] * y[0];
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    quat_normalize(r.q);
    return r;
}
// MarkerEdge::computeError: corners of the marker through camera * marker, projections narrowed to float (typesg2o.h:133-165)
__device__ void marker_edge_error(const Pose& c2g, const Pose& g2m, float size, const float* obs, const Cam& cam, double* e) {
    const Pose c2m = pose_compose(c2g, g2m);
    const float hp = (float)(size / 2.), hn = (float)(-size / 2.);  // Marker::get3DPointsLocalRefSystem, marker.cpp:58-62
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double c[3] = {(double)((i == 0 || i == 3) ? hn : hp), (double)((i < 2) ? hp : hn), 0.0}, p[3];
        se3_map(c2m, c, p);
        const float projx = (float)((p[0] / p[2]) * cam.fx + cam.cx);
        const float projy = (float)((p[1] / p[2]) * cam.fy + cam.cy);
        e[2 * i] = (double)obs[2 * i] - (double)projx;
        e[2 * i + 1] = (double)obs[2 * i + 1] - (double)projy;
    }
}
// one thread per marker edge: chi2 = w |e|^2; with `linearize` also the numeric Jacobians (central differences, delta 1e-4,
// base_binary_edge.hpp:165-230) of both vertices and the edge's quadratic-form blocks (base_binary_edge.hpp:83-155, no robust kernel)
// MarkerEdgeX::computeError (globaloptimizer_g2o.cpp:52-63): with M = inverse(ref) * other as 4x4 matrices, 10 * (M(0,2), M(1,2), 1 - M(2,2), M(2,3));
// in closed form M(i,2) = column i of R_ref . column 2 of R_other, M(2,3) = column 2 of R_ref . (t_other - t_ref)
__device__ void planar_edge_error(const Pose& A, const Pose& O, double* e) {
    double Ra[9], Ro[9];
    quat_to_R(A.q, Ra);
    quat_to_R(O.q, Ro);
    const double d0 = O.t[0] - A.t[0], d1 = O.t[1] - A.t[1], d2 = O.t[2] - A.t[2];
    e[0] = 10. * (Ra[0] * Ro[2] + Ra[3] * Ro[5] + Ra[6] * Ro[8]);
    e[1] = 10. * (Ra[1] * Ro[2] + Ra[4] * Ro[5] + Ra[7] * Ro[8]);
    e[2] = 10. * (1 - (Ra[2] * Ro[2] + Ra[5] * Ro[5] + Ra[8] * Ro[8]));
    e[3] = 10. * (Ra[2] * d0 + Ra[5] * d1 + Ra[8] * d2);
}
// one planar edge: chi2, and with `linearize` g2o's own numeric Jacobians (base_binary_edge.hpp:165-233: central differences with
// delta = 1e-9f on every free vertex) and the quadratic-form blocks, stored so that the "c" slots are the lower-numbered marker's
__device__ void planar_edge(const BaDev& B, int x, int linearize) {
    const int ro = B.mk.x_other[x], rr = B.mk.x_ref;
    const Pose O = load_pose(B.mk.pose + 7 * ro), A = load_pose(B.mk.pose + 7 * (rr >= 0 ? rr : B.mk.Nm));
    const double w = B.mk.x_w;
    double e0[4];
    planar_edge_error(A, O, e0);
    B.mk.e_chi2[B.mk.Ne + x] = e0[0] * w * e0[0] + e0[1] * w * e0[1] + e0[2] * w * e0[2] + e0[3] * w * e0[3];
    if (!linearize) return;
    const double delta = (double)1e-9f, scalar = 1 / (2 * delta);
    double Ja[24], Jo[24];  // [row * 6 + d]
    for (int d = 0; d < 6; d++) {
        double u[6] = {0, 0, 0, 0, 0, 0}, ea[4], eb[4];
        Pose Op = O, On = O;
        u[d] = delta; se3_oplus(Op, u);
        u[d] = -delta; se3_oplus(On, u);
        planar_edge_error(A, Op, ea);
        planar_edge_error(A, On, eb);
        for (int i = 0; i < 4; i++) Jo[i * 6 + d] = scalar * (ea[i] - eb[i]);
        if (rr >= 0) {
            Pose Ap = A, An = A;
            u[d] = delta; se3_oplus(Ap, u);
            u[d] = -delta; se3_oplus(An, u);
            planar_edge_error(Ap, O, ea);
            planar_edge_error(An, O, eb);
            for (int i = 0; i < 4; i++) Ja[i * 6 + d] = scalar * (ea[i] - eb[i]);
        } else {
            for (int i = 0; i < 4; i++) Ja[i * 6 + d] = 0;
        }
    }
    const bool ref_first = rr >= 0 && rr < ro;          // which marker owns the "c" slots
    const double* Jc = ref_first ? Ja : Jo;
    const double* Jm = ref_first ? Jo : Ja;
    double* o = B.mk.e_blk + 120 * (size_t)(B.mk.Ne + x);
    if (rr < 0) { Jc = Ja; Jm = Jo; }                   // fixed reference: only the "m" (other) slots are read
    for (int a = 0; a < 6; a++) {
        for (int c = 0; c < 6; c++) {
            double hcc = 0, hmm = 0, hcm = 0;
            for (int i = 0; i < 4; i++) {
                hcc += Jc[i * 6 + a] * w * Jc[i * 6 + c];
                hmm += Jm[i * 6 + a] * w * Jm[i * 6 + c];
                hcm += Jc[i * 6 + a] * w * Jm[i * 6 + c];
            }
            o[6 * a + c] = hcc;
            o[36 + 6 * a + c] = hmm;
            o[72 + 6 * a + c] = hcm;
        }
        double bc = 0, bm = 0;
        for (int i = 0; i < 4; i++) {
            bc += Jc[i * 6 + a] * (-(w * e0[i]));
            bm += Jm[i * 6 + a] * (-(w * e0[i]));
        }
        o[108 + a] = bc;
        o[114 + a] = bm;
    }
}
__global__ void __launch_bounds__(64) ba_marker_kernel(const __grid_constant__ BaDev B, int linearize) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= B.mk.Ne) {
        if (k < B.mk.Ne + B.mk.Nx) planar_edge(B, k - B.mk.Ne, linearize);
        return;
    }
    const int m = B.mk.e_marker[k], pi = B.mk.e_pose[k];
    const Pose T = load_pose(B.pose + 7 * pi), G = load_pose(B.mk.pose + 7 * m);
    const float size = B.mk.size[m];
    const float* obs = B.mk.e_corners + 8 * k;
    const double w = (double)B.mk.e_weight[k];
    const Cam& cam = cam_of(B, pi);
    double e0[8];
    marker_edge_error(T, G, size, obs, cam, e0);
    double c2 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) c2 += e0[i] * w * e0[i];
    B.mk.e_chi2[k] = c2;
    if (!linearize) return;
    const double delta = 1e-4, scalar = 1 / (2 * delta);
    const bool cam_free = B.free_idx[pi] >= 0;
    double Jm[48], Jc[48];  // [row * 6 + d]
    for (int d = 0; d < 6; d++) {
        double u[6] = {0, 0, 0, 0, 0, 0}, ea[8], eb[8];
        Pose Gp = G, Gn = G;
        u[d] = delta; se3_oplus(Gp, u);
        u[d] = -delta; se3_oplus(Gn, u);
        marker_edge_error(T, Gp, size, obs, cam, ea);
        marker_edge_error(T, Gn, size, obs, cam, eb);
        for (int i = 0; i < 8; i++) Jm[i * 6 + d] = scalar * (ea[i] - eb[i]);
        if (cam_free) {
            Pose Tp = T, Tn = T;
            u[d] = delta; se3_oplus(Tp, u);
            u[d] = -delta; se3_oplus(Tn, u);
            marker_edge_error(Tp, G, size, obs, cam, ea);
            marker_edge_error(Tn, G, size, obs, cam, eb);
            for (int i = 0; i < 8; i++) Jc[i * 6 + d] = scalar * (ea[i] - eb[i]);
        } else {
            for (int i = 0; i < 8; i++) Jc[i * 6 + d] = 0;
        }
    }
    double* o = B.mk.e_blk + 120 * (size_t)k;
    for (int a = 0; a < 6; a++) {
        for (int c = 0; c < 6; c++) {
            double hcc = 0, hmm = 0, hcm = 0;
            for (int i = 0; i < 8; i++) {
                hcc += Jc[i * 6 + a] * w * Jc[i * 6 + c];
                hmm += Jm[i * 6 + a] * w * Jm[i * 6 + c];
                hcm += Jc[i * 6 + a] * w * Jm[i * 6 + c];
            }
            o[6 * a + c] = hcc;
            o[36 + 6 * a + c] = hmm;
            o[72 + 6 * a + c] = hcm;
        }
        double bc = 0, bm = 0;
        for (int i = 0; i < 8; i++) {
            bc += Jc[i * 6 + a] * (-(w * e0[i]));
            bm += Jm[i * 6 + a] * (-(w * e0[i]));
        }
        o[108 + a] = bc;
        o[114 + a] = bm;
    }
}
// ordered accumulation of the marker edges' diagonal blocks: keyframe f gets += sum Hcc / bc of its edges (after the all-reduce of the
// keypoint part: the markers are replicated, not sharded), marker m gets its Hmm / bm
__global__ void __launch_bounds__(128) ba_marker_accumulate_kernel(const __grid_constant__ BaDev B) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B.Pf + B.mk.Nm) return;
    const bool is_cam = t < B.Pf;
    const int* ptr = is_cam ? B.mk.cam_ptr : B.mk.mk_ptr;
    const int* lst = is_cam ? B.mk.cam_edges : B.mk.mk_edges;
    const int i = is_cam ? t : t - B.Pf;
    double H[36], b[6];
    for (int k = 0; k < 36; k++) H[k] = is_cam ? B.Hpp[36 * (size_t)t + k] : 0.0;
    for (int k = 0; k < 6; k++) b[k] = is_cam ? B.bp[6 * (size_t)t + k] : 0.0;
    for (int j = ptr[i]; j < ptr[i + 1]; j++) {
        const int ed = lst[j];
        const bool c_role = is_cam || ed < 0;          // ~index: this marker owns the "c" slots of a planar edge
        const double* o = B.mk.e_blk + 120 * (size_t)(ed < 0 ? ~ed : ed);
        for (int k = 0; k < 36; k++) H[k] += o[(c_role ? 0 : 36) + k];
        for (int k = 0; k < 6; k++) b[k] += o[(c_role ? 108 : 114) + k];
    }
    for (int k = 0; k < 36; k++) B.Hpp[36 * (size_t)t + k] = H[k];
    for (int k = 0; k < 6; k++) B.bp[6 * (size_t)t + k] = b[k];
}
// marker part of the update (push + oplus + computeScale terms); xp / bp / scale_pose entries Pf .. Pf + Nm - 1 belong to the markers
__global__ void __launch_bounds__(128) ba_marker_update_kernel(const __grid_constant__ BaDev B) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= B.mk.Nm) return;
    const LmState* st = B.st;
    const bool apply = !st->chol_fail;
    const double lambda = st->lambda;
    Pose G = load_pose(B.mk.pose + 7 * m);
    store_pose(B.mk.pose_bak + 7 * m, G);
    double u[6], sc = 0;
    for (int k = 0; k < 6; k++) {
        u[k] = B.xp[6 * (B.Pf + m) + k];
        sc += u[k] * (lambda * u[k] + B.bp[6 * (size_t)(B.Pf + m) + k]);
    }
    B.scale_pose[B.Pf + m] = sc;
    if (apply) {
        se3_oplus(G, u);
        store_pose(B.mk.pose + 7 * m, G);
    }
}
__global__ void ba_marker_results_kernel(const __grid_constant__ BaDev B, float* __restrict__ m44) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= B.mk.Nm) return;
    const Pose G = load_pose(B.mk.pose + 7 * m);
    double R[9];
    quat_to_R(G.q, R);
    float* o = m44 + 16 * m;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) o[4 * r + c] = (float)R[3 * r + c];
        o[4 * r + 3] = (float)G.t[r];
    }
    o[12] = o[13] = o[14] = 0;
    o[15] = 1;
}

// this rank's ordered partial sums for the LM control: buf[0] = sum rho0, buf[1] = sum scale_lm, buf[3] = stop flag;
// mx[0] = max |diagonal| of (full) Hpp and of the local Hll
__global__ void __launch_bounds__(1024) ba_partial_sums_kernel(const __grid_constant__ BaDev B, double* __restrict__ buf, double* __restrict__ mx, int stop,
                                                               int add_markers) {
    __shared__ double sm[33];
    double s = 0;
    for (int i = threadIdx.x; i < B.M; i += 1024) s += B.rho0[i];
    if (add_markers)  // the marker edges are replicated: only one rank contributes their chi2 to the all-reduced sum
        for (int k = threadIdx.x; k < B.mk.Ne + B.mk.Nx; k += 1024) s += B.mk.e_chi2[k];
    s = block_reduce_1024<false>(s, sm);
    double sc = 0;
    for (int l = threadIdx.x; l < B.N; l += 1024) sc += B.scale_lm[l];
    sc = block_reduce_1024<false>(sc, sm);
    double md = 0;
    if (mx) {
        for (int k = threadIdx.x; k < 6 * (B.Pf + B.mk.Nm); k += 1024) md = fmax(md, fabs(B.Hpp[36 * (size_t)(k / 6) + 7 * (k % 6)]));
        for (int k = threadIdx.x; k < 3 * B.N; k += 1024) {
            int l = k / 3, j = k % 3;
            md = fmax(md, fabs(B.Hll[6 * (size_t)l + (j == 0 ? 0 : (j == 1 ? 3 : 5))]));
        }
        md = block_reduce_1024<true>(md, sm);
    }
    if (threadIdx.x == 0) {
        buf[0] = s;
        buf[1] = sc;
        buf[2] = 0;
        buf[3] = stop ? 1.0 : 0.0;
        if (mx) mx[0] = md;
    }
}

// after the library Cholesky of a large reduced system: publish the outcome where the update / decide kernels look for it
__global__ void ba_solve_finish_kernel(const __grid_constant__ BaDev B, const int* __restrict__ info, const double* __restrict__ x) {
    const bool fail = info[0] != 0 || info[1] != 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < B.n; k += gridDim.x * blockDim.x) B.xp[k] = fail ? 0.0 : x[k];
    if (blockIdx.x == 0 && threadIdx.x == 0) B.st->chol_fail = fail;
}

// ---- host side ---------------------------------------------------------------------------------------------------------------------
struct Arena {
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    }
};

}  // namespace

// ---- sharded / large form, host side -----------------------------------------------------------------------------------------------
// reduced systems beyond the single-CTA dense solver: two-level block-envelope Cholesky (ba_band.cu)
namespace {
// landmark range [L[r], L[r+1]) of every rank: contiguous, balanced by observation count (the Schur work follows it)
void ba_partition_landmarks(const std::vector<int>& lm_ptr, int N, int M, int R, std::vector<int>& L) {
    L.assign(R + 1, N);
    L[0] = 0;
    int l = 0;
    for (int k = 1; k < R; k++) {
        const long long target = (long long)M * k / R;
        while (l < N && lm_ptr[l] < target) l++;
        L[k] = l;
    }
    L[R] = N;
}
}  // namespace

struct uco_ba_state {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int smem_optin = 0;
};

void uco_ba_state_free(uco_b200_ctx* ctx) {
    if (!ctx->ba) return;
    if (ctx->ba->ev0) cudaEventDestroy(ctx->ba->ev0);
    if (ctx->ba->ev1) cudaEventDestroy(ctx->ba->ev1);
    delete ctx->ba;
    ctx->ba = nullptr;
}

cudaEvent_t* uco_ba_events(uco_b200_ctx* ctx) {
    if (!ctx->ba) {
        ctx->ba = new uco_ba_state();
        cudaEventCreate(&ctx->ba->ev0);
        cudaEventCreate(&ctx->ba->ev1);
        cudaDeviceGetAttribute(&ctx->ba->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    }
    return &ctx->ba->ev0;
}

// streamed form: one kernel per phase, the host enqueues the next trial after reading two flags (any problem size)
int ba_sharded_solve(uco_b200_ctx* ctx, uco_b200_comm* comm, const uco_ba_problem* pb, const volatile unsigned char* stop, uco_ba_result* res);
int ba_streamed_solve(uco_b200_ctx* ctx, const uco_ba_problem* pb, const volatile unsigned char* stop, uco_ba_result* res) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (!pb || !res) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: null problem / result");
    if (pb->pose_cam) return ba_sharded_solve(ctx, nullptr, pb, stop, res);   // one camera per keyframe: only the sharded solver's kernels read the per-pose table
    const int P = pb->n_poses, N = pb->n_points, M = pb->n_obs;
    if (P <= 0 || N < 0 || M < 0 || pb->n_iters < 0) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: bad sizes");
    if (!pb->poses44 || !pb->fixed || (N && !pb->points3) || (M && (!pb->obs_pose || !pb->obs_point || !pb->obs_uv || !pb->obs_inv_sigma2)))
        return uco_fail(ctx, UCO_E_INVALID, "ba_solve: null input array");
    for (int i = 0; i < M; i++) {
        if ((unsigned)pb->obs_pose[i] >= (unsigned)P || (unsigned)pb->obs_point[i] >= (unsigned)N)
            return uco_fail(ctx, UCO_E_INVALID, "ba_solve: observation %d references pose %d / point %d out of range", i,
                            pb->obs_pose[i], pb->obs_point[i]);
        if (pb->obs_stereo && pb->obs_stereo[i] && !pb->obs_ur) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: stereo observation without obs_ur");
    }
    uco_ba_events(ctx);
    // ---- structure (host): free-pose numbering, observations sorted by landmark, per-pose lists, Schur gather lists
    std::vector<int> free_idx(P), free_list;
    for (int i = 0; i < P; i++) {
        free_idx[i] = pb->fixed[i] ? -1 : (int)free_list.size();
        if (!pb->fixed[i]) free_list.push_back(i);
    }
    const int Pf = (int)free_list.size(), n = 6 * Pf;
    if (n > 1023) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: %d free poses exceed the single-CTA reduced solver (170)", Pf);
    std::vector<int> lm_ptr(N + 1, 0), order(M), fill(N, 0);
    for (int i = 0; i < M; i++) lm_ptr[pb->obs_point[i] + 1]++;
    for (int l = 0; l < N; l++) lm_ptr[l + 1] += lm_ptr[l];
    for (int i = 0; i < M; i++) order[lm_ptr[pb->obs_point[i]] + fill[pb->obs_point[i]]++] = i;  // sorted position -> caller index
    std::vector<int> s_pose(M), s_lm(M);
    for (int k = 0; k < M; k++) { s_pose[k] = pb->obs_pose[order[k]]; s_lm[k] = pb->obs_point[order[k]]; }
    std::vector<int> pose_ptr(Pf + 1, 0), pose_obs;
    for (int k = 0; k < M; k++) if (free_idx[s_pose[k]] >= 0) pose_ptr[free_idx[s_pose[k]] + 1]++;
    for (int f = 0; f < Pf; f++) pose_ptr[f + 1] += pose_ptr[f];
    pose_obs.resize(pose_ptr[Pf]);
    {
        std::vector<int> pf(Pf, 0);
        for (int k = 0; k < M; k++) { int f = free_idx[s_pose[k]]; if (f >= 0) pose_obs[pose_ptr[f] + pf[f]++] = k; }
    }
    // gather lists: key = i * Pf + j (i <= j), contributions in landmark order
    std::vector<int> blk_cnt((size_t)Pf * Pf, 0);
    for (int l = 0; l < N; l++)
        for (int a = lm_ptr[l]; a < lm_ptr[l + 1]; a++) {
            int fa = free_idx[s_pose[a]];
            if (fa < 0) continue;
            for (int b = lm_ptr[l]; b < lm_ptr[l + 1]; b++) {
                int fb = free_idx[s_pose[b]];
                if (fb < fa || (fb == fa && b != a)) continue;
                blk_cnt[(size_t)fa * Pf + fb]++;
            }
        }
    std::vector<int> blk_of((size_t)Pf * Pf, -1), blk_ptr(1, 0);
    std::vector<int2> blk_ij;
    for (int i = 0; i < Pf; i++)
        for (int j = i; j < Pf; j++)
            if (blk_cnt[(size_t)i * Pf + j] || i == j) {  // diagonal blocks always exist (Hpp + lambda I)
                blk_of[(size_t)i * Pf + j] = (int)blk_ij.size();
                blk_ij.push_back(make_int2(i, j));
                blk_ptr.push_back(blk_ptr.back() + blk_cnt[(size_t)i * Pf + j]);
            }
    const int nblk = (int)blk_ij.size();
    std::vector<int2> con(blk_ptr.back());
    {
        std::vector<int> bf(nblk, 0);
        for (int l = 0; l < N; l++)
            for (int a = lm_ptr[l]; a < lm_ptr[l + 1]; a++) {
                int fa = free_idx[s_pose[a]];
                if (fa < 0) continue;
                for (int b = lm_ptr[l]; b < lm_ptr[l + 1]; b++) {
                    int fb = free_idx[s_pose[b]];
                    if (fb < fa || (fb == fa && b != a)) continue;
                    int k = blk_of[(size_t)fa * Pf + fb];
                    con[blk_ptr[k] + bf[k]++] = make_int2(a, b);
                }
            }
    }
    // ---- arena: [inputs copied from the host in one transfer][device-only work arrays]
    Arena A;
    const size_t o_free_idx = A.take(4 * (size_t)P), o_free_list = A.take(4 * (size_t)(Pf + 1)), o_lm_ptr = A.take(4 * (size_t)(N + 1)),
                 o_obs_pose = A.take(4 * (size_t)(M + 1)), o_obs_lm = A.take(4 * (size_t)(M + 1)), o_pose_ptr = A.take(4 * (size_t)(Pf + 1)),
                 o_pose_obs = A.take(4 * (pose_obs.size() + 1)), o_blk_ptr = A.take(4 * (size_t)(nblk + 1)),
                 o_blk_ij = A.take(8 * (size_t)(nblk + 1)), o_con = A.take(8 * (con.size() + 1)), o_z = A.take(24 * (size_t)(M + 1)),
                 o_info = A.take(8 * (size_t)(M + 1)), o_stereo = A.take((size_t)M + 1), o_active = A.take((size_t)M + 1),
                 o_pt = A.take(24 * (size_t)(N + 1)), o_p44 = A.take(64 * (size_t)P);
    const size_t in_bytes = A.off;
    const size_t o_pose = A.take(56 * (size_t)P), o_pose_bak = A.take(56 * (size_t)P), o_pt_bak = A.take(24 * (size_t)(N + 1)),
                 o_err = A.take(24 * (size_t)(M + 1)), o_chi2 = A.take(8 * (size_t)(M + 1)), o_rho0 = A.take(8 * (size_t)(M + 1)),
                 o_Hll = A.take(48 * (size_t)(N + 1)), o_bl = A.take(24 * (size_t)(N + 1)), o_Hpl = A.take(144 * (size_t)(M + 1)),
                 o_Y = A.take(144 * (size_t)(M + 1)), o_Dinv = A.take(48 * (size_t)(N + 1)), o_db = A.take(24 * (size_t)(N + 1)),
                 o_xl = A.take(24 * (size_t)(N + 1)), o_Hpp = A.take(288 * (size_t)(Pf + 1)), o_bp = A.take(48 * (size_t)(Pf + 1)),
                 o_S = A.take(8 * ((size_t)n * n + 1)), o_bs = A.take(8 * (size_t)(n + 1)), o_xp = A.take(8 * (size_t)(n + 1)),
                 o_scl = A.take(8 * (size_t)(N + 1)), o_scp = A.take(8 * (size_t)(Pf + 1)), o_st = A.take(sizeof(LmState)),
                 o_p44o = A.take(64 * (size_t)P), o_bad = A.take((size_t)M + 1),
                 o_chol = A.take(8 * ((size_t)(n + 1) * (n + 2) / 2 + 1));
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_BA, A.off);
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_BA, in_bytes + sizeof(LmState));
    if (!d || !h) return UCO_E_NOMEM;
    memcpy(h + o_free_idx, free_idx.data(), 4 * (size_t)P);
    memcpy(h + o_free_list, free_list.data(), 4 * (size_t)Pf);
    memcpy(h + o_lm_ptr, lm_ptr.data(), 4 * (size_t)(N + 1));
    memcpy(h + o_obs_pose, s_pose.data(), 4 * (size_t)M);
    memcpy(h + o_obs_lm, s_lm.data(), 4 * (size_t)M);
    memcpy(h + o_pose_ptr, pose_ptr.data(), 4 * (size_t)(Pf + 1));
    memcpy(h + o_pose_obs, pose_obs.data(), 4 * pose_obs.size());
    memcpy(h + o_blk_ptr, blk_ptr.data(), 4 * (size_t)(nblk + 1));
    memcpy(h + o_blk_ij, blk_ij.data(), 8 * (size_t)nblk);
    memcpy(h + o_con, con.data(), 8 * con.size());
    {
        double* z = (double*)(h + o_z);
        double* info = (double*)(h + o_info);
        uint8_t* st = h + o_stereo;
        for (int k = 0; k < M; k++) {
            int i = order[k];
            bool s = pb->obs_stereo && pb->obs_stereo[i];
            z[3 * k] = pb->obs_uv[2 * i]; z[3 * k + 1] = pb->obs_uv[2 * i + 1]; z[3 * k + 2] = s ? pb->obs_ur[i] : 0.0;
            info[k] = pb->obs_inv_sigma2[i];
            st[k] = s;
        }
        memset(h + o_active, 1, (size_t)M);
        double* pt = (double*)(h + o_pt);
        for (int k = 0; k < 3 * N; k++) pt[k] = pb->points3[k];
        memcpy(h + o_p44, pb->poses44, 64 * (size_t)P);
    }
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_st, 0, sizeof(LmState), s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_err, 0, 24 * (size_t)(M + 1), s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_chi2, 0, 8 * (size_t)(M + 1), s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_S, 0, 8 * ((size_t)n * n + 1), s));  // blocks without shared landmarks are never written: they stay zero
    BaDev B;
    B.P = P; B.N = N; B.M = M; B.Pf = Pf; B.n = n; B.nblk = nblk;
    B.pose = (double*)(d + o_pose); B.pose_bak = (double*)(d + o_pose_bak); B.pt = (double*)(d + o_pt); B.pt_bak = (double*)(d + o_pt_bak);
    B.free_idx = (int*)(d + o_free_idx); B.free_list = (int*)(d + o_free_list); B.lm_ptr = (int*)(d + o_lm_ptr);
    B.obs_pose = (int*)(d + o_obs_pose); B.obs_lm = (int*)(d + o_obs_lm); B.pose_ptr = (int*)(d + o_pose_ptr); B.pose_obs = (int*)(d + o_pose_obs);
    B.z = (double*)(d + o_z); B.info = (double*)(d + o_info); B.stereo = d + o_stereo; B.active = d + o_active;
    B.err = (double*)(d + o_err); B.chi2 = (double*)(d + o_chi2); B.rho0 = (double*)(d + o_rho0); B.Hll = (double*)(d + o_Hll);
    B.bl = (double*)(d + o_bl); B.Hpl = (double*)(d + o_Hpl); B.Y = (double*)(d + o_Y); B.Dinv = (double*)(d + o_Dinv); B.db = (double*)(d + o_db);
    B.xl = (double*)(d + o_xl); B.Hpp = (double*)(d + o_Hpp); B.bp = (double*)(d + o_bp); B.S = (double*)(d + o_S); B.bs = (double*)(d + o_bs);
    B.xp = (double*)(d + o_xp); B.scale_lm = (double*)(d + o_scl); B.scale_pose = (double*)(d + o_scp);
    B.blk_ptr = (int*)(d + o_blk_ptr); B.blk_ij = (int2*)(d + o_blk_ij); B.con = (int2*)(d + o_con);
    B.st = (LmState*)(d + o_st);
    B.cam.fx = pb->fx; B.cam.fy = pb->fy; B.cam.cx = pb->cx; B.cam.cy = pb->cy; B.cam.bf = pb->bf; B.cam.bf_f = pb->bf;
    B.chi2d = 5.99f; B.chi3d = 7.815f;
    B.d2 = (double)sqrtf(B.chi2d); B.d3 = (double)sqrtf(B.chi3d);  // const float thHuber2D = sqrt(Chi2D)
    LmState* hst = (LmState*)(h + in_bytes);
    const size_t chol_bytes = 8 * ((size_t)(n + 1) * (n + 2) / 2);
    const int chol_smem = chol_bytes + 16 * 1024 <= (size_t)ctx->ba->smem_optin;
    if (chol_smem) UCO_CUDA(ctx, cudaFuncSetAttribute(ba_chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_bytes));
    if (n > 1023) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: reduced system too large");
    const int gM = (M + 255) / 256, gN = (N + 127) / 128, gU = (N + Pf + 127) / 128;

    UCO_CUDA(ctx, cudaEventRecord(ctx->ba->ev0, s));
    ba_init_poses_kernel<<<(P + 127) / 128, 128, 0, s>>>(B, (const float*)(d + o_p44));
    UCO_LAUNCH_CHECK(ctx);
    int iters[2] = {0, 0};
    bool stopped = false;
    for (int stage = 0; stage < 2 && !stopped; stage++) {
        const int robust = stage == 0, max_iters = stage == 0 ? pb->n_iters : 2 * pb->n_iters;
        if (stage == 1 && M) {
            ba_flag_outliers_kernel<<<gM, 256, 0, s>>>(B);
            UCO_LAUNCH_CHECK(ctx);
        }
        if (M) {
            ba_errors_kernel<<<gM, 256, 0, s>>>(B, robust);
            UCO_LAUNCH_CHECK(ctx);
        }
        bool cont_iter = max_iters > 0 && !(stop && *stop);
        for (int it = 0; cont_iter; it++) {
            if (N) {
                ba_linearize_lm_kernel<<<gN, 128, 0, s>>>(B, robust);
                UCO_LAUNCH_CHECK(ctx);
            }
            if (Pf) {
                ba_linearize_pose_kernel<<<Pf, POSE_THREADS, 0, s>>>(B, robust);
                UCO_LAUNCH_CHECK(ctx);
            }
            ba_iter_begin_kernel<<<1, 1024, 0, s>>>(B, it == 0);
            UCO_LAUNCH_CHECK(ctx);
            bool cont_trial = true;
            while (cont_trial) {
                if (N) {
                    ba_prep_kernel<<<gN, 128, 0, s>>>(B);
                    UCO_LAUNCH_CHECK(ctx);
                }
                if (Pf) {
                    ba_schur_gather_kernel<<<nblk, 36 * GATHER_CHUNKS, 0, s>>>(B);
                    UCO_LAUNCH_CHECK(ctx);
                    ba_chol_solve_kernel<<<1, CHOL_THREADS, chol_smem ? chol_bytes : 0, s>>>(B, (double*)(d + o_chol), chol_smem);
                    UCO_LAUNCH_CHECK(ctx);
                }
                ba_update_kernel<<<gU, 128, 0, s>>>(B);
                UCO_LAUNCH_CHECK(ctx);
                if (M) {
                    ba_errors_kernel<<<gM, 256, 0, s>>>(B, robust);
                    UCO_LAUNCH_CHECK(ctx);
                }
                ba_decide_kernel<<<1, 1024, 0, s>>>(B, stop && *stop ? 1 : 0, max_iters);
                UCO_LAUNCH_CHECK(ctx);
                UCO_CUDA(ctx, cudaMemcpyAsync(hst, B.st, offsetof(LmState, trace), cudaMemcpyDeviceToHost, s));
                UCO_CUDA(ctx, cudaStreamSynchronize(s));
                cont_trial = hst->cont_trial != 0;
            }
            cont_iter = hst->cont_iter != 0;
            iters[stage] = hst->it;
        }
        if (stop && *stop) stopped = true;  // GlobalOptimizerG2O::optimize: no second stage once stopASAP is raised
    }
    // ---- results
    float* p44o = (float*)(d + o_p44o);
    uint8_t* bad = d + o_bad;
    ba_results_kernel<<<(P + 255) / 256, 256, 0, s>>>(B, (const float*)(d + o_p44), p44o, bad);
    UCO_LAUNCH_CHECK(ctx);
    if (M) {
        ba_bad_kernel<<<gM, 256, 0, s>>>(B, p44o, bad);
        UCO_LAUNCH_CHECK(ctx);
    }
    UCO_CUDA(ctx, cudaEventRecord(ctx->ba->ev1, s));
    // D2H through a pinned staging area (reuses the input staging buffer), then scatter back to the caller's observation order
    const size_t o_h_pose = 0, o_h_p44 = o_h_pose + 56 * (size_t)P, o_h_pt = o_h_p44 + 64 * (size_t)P, o_h_chi = o_h_pt + 24 * (size_t)N,
                 o_h_act = o_h_chi + 8 * (size_t)M, o_h_bad = o_h_act + (size_t)M, o_h_st = ((o_h_bad + (size_t)M + 7) & ~(size_t)7),
                 out_bytes = o_h_st + sizeof(LmState);
    uint8_t* ho = (uint8_t*)uco_pinned(ctx, WS_BA_OUT, out_bytes);
    if (!ho) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_pose, B.pose, 56 * (size_t)P, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_p44, p44o, 64 * (size_t)P, cudaMemcpyDeviceToHost, s));
    if (N) UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_pt, B.pt, 24 * (size_t)N, cudaMemcpyDeviceToHost, s));
    if (M) {
        UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_chi, B.chi2, 8 * (size_t)M, cudaMemcpyDeviceToHost, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_act, B.active, (size_t)M, cudaMemcpyDeviceToHost, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_bad, bad, (size_t)M, cudaMemcpyDeviceToHost, s));
    }
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_st, B.st, sizeof(LmState), cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    if (res->pose7) memcpy(res->pose7, ho + o_h_pose, 56 * (size_t)P);
    if (res->poses44) memcpy(res->poses44, ho + o_h_p44, 64 * (size_t)P);
    if (res->points3) memcpy(res->points3, ho + o_h_pt, 24 * (size_t)N);
    const double* hchi = (const double*)(ho + o_h_chi);
    for (int k = 0; k < M; k++) {
        int i = order[k];
        if (res->obs_chi2) res->obs_chi2[i] = hchi[k];
        if (res->obs_level) res->obs_level[i] = !(ho + o_h_act)[k];
        if (res->obs_bad) res->obs_bad[i] = (ho + o_h_bad)[k];
    }
    const LmState* fst = (const LmState*)(ho + o_h_st);
    if (res->trace) {
        memset(res->trace, 0, sizeof(double) * 128);
        memcpy(res->trace, fst->trace, sizeof(double) * 2 * (size_t)std::min(fst->ntrace, 64));
    }
    res->iters[0] = iters[0];
    res->iters[1] = iters[1];
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ba->ev0, ctx->ba->ev1);
    res->device_ms = ms;
    if (res->profile) memset(res->profile, 0, sizeof(double) * 16);
    return UCO_OK;
}


int ba_sharded_solve(uco_b200_ctx* ctx, uco_b200_comm* comm, const uco_ba_problem* pb, const volatile unsigned char* stop, uco_ba_result* res) {
    if (!ctx) return UCO_E_INVALID;
    const auto t_enter = std::chrono::steady_clock::now();
    cudaSetDevice(ctx->device);
    if (!res) return uco_fail(ctx, UCO_E_INVALID, "ba_solve_sharded: null result");
    int rc = ba_validate(ctx, pb);
    if (rc != UCO_OK) return rc;
    const int R = uco_comm_world(comm), rank = uco_comm_rank(comm);
    const int P = pb->n_poses, N = pb->n_points, M = pb->n_obs;
    const int Nm = pb->n_markers, Ne = pb->n_marker_obs;
    if (Nm < 0 || Ne < 0 || (Nm && (!pb->marker_pose44 || !pb->marker_size)) || (Ne && (!Nm || !pb->mobs_marker || !pb->mobs_pose || !pb->mobs_corners || !pb->mobs_weight)))
        return uco_fail(ctx, UCO_E_INVALID, "ba_solve: malformed marker arrays");
    for (int k = 0; k < Ne; k++)
        if ((unsigned)pb->mobs_marker[k] >= (unsigned)Nm || (unsigned)pb->mobs_pose[k] >= (unsigned)P)
            return uco_fail(ctx, UCO_E_INVALID, "ba_solve: marker observation %d references marker %d / pose %d out of range", k, pb->mobs_marker[k], pb->mobs_pose[k]);
    const int Nx = pb->n_plane, xref = pb->plane_ref;
    if (Nx < 0 || (Nx && (!Nm || !pb->plane_other || xref >= Nm || (xref < 0 && !pb->plane_ref_pose44) || !(pb->plane_weight >= 0))))
        return uco_fail(ctx, UCO_E_INVALID, "ba_solve: malformed planar-marker arrays");
    for (int x = 0; x < Nx; x++)
        if ((unsigned)pb->plane_other[x] >= (unsigned)Nm || pb->plane_other[x] == xref)
            return uco_fail(ctx, UCO_E_INVALID, "ba_solve: planar edge %d references marker %d", x, pb->plane_other[x]);
    uco_ba_events(ctx);
    // ---- global structure (identical on every rank): free-pose numbering, observations sorted by landmark, Schur block list
    std::vector<int> free_idx(P), free_list;
    for (int i = 0; i < P; i++) {
        free_idx[i] = pb->fixed[i] ? -1 : (int)free_list.size();
        if (!pb->fixed[i]) free_list.push_back(i);
    }
    const int Pf = (int)free_list.size(), PT = Pf + Nm, n = 6 * PT;   // unknowns: free keyframes, then markers
    std::vector<int> lm_ptr(N + 1, 0), order(M), fill(N, 0);
    for (int i = 0; i < M; i++) lm_ptr[pb->obs_point[i] + 1]++;
    for (int l = 0; l < N; l++) lm_ptr[l + 1] += lm_ptr[l];
    for (int i = 0; i < M; i++) order[lm_ptr[pb->obs_point[i]] + fill[pb->obs_point[i]]++] = i;
    std::vector<int> g_free(M);  // free index of the pose of the k-th sorted observation
    for (int k = 0; k < M; k++) g_free[k] = free_idx[pb->obs_pose[order[k]]];
    std::vector<int> L;
    ba_partition_landmarks(lm_ptr, N, M, R, L);
    const int l0 = L[rank], l1 = L[rank + 1], NL = l1 - l0, o0 = lm_ptr[l0], o1 = lm_ptr[l1], ML = o1 - o0;
    // block list from ALL landmarks (so that the packed layout agrees across ranks), contributions from the local ones
    std::vector<int> blk_of((size_t)PT * PT, -1);
    {
        std::vector<uint8_t> present((size_t)PT * PT, 0);
        for (int k = 0; k < Ne; k++) {  // (keyframe, marker) blocks of the marker edges; marker diagonals exist like every diagonal
            const int fc = free_idx[pb->mobs_pose[k]];
            if (fc >= 0) present[(size_t)fc * PT + Pf + pb->mobs_marker[k]] = 1;
        }
        for (int x = 0; x < Nx && xref >= 0; x++)   // (reference marker, other marker) blocks of the planar edges
            present[(size_t)(Pf + std::min(xref, pb->plane_other[x])) * PT + Pf + std::max(xref, pb->plane_other[x])] = 1;
        for (int l = 0; l < N; l++)
            for (int a = lm_ptr[l]; a < lm_ptr[l + 1]; a++) {
                const int fa = g_free[a];
                if (fa < 0) continue;
                for (int b = lm_ptr[l]; b < lm_ptr[l + 1]; b++) {
                    const int fb = g_free[b];
                    if (fb < fa || (fb == fa && b != a)) continue;
                    present[(size_t)fa * PT + fb] = 1;
                }
            }
        int nb = 0;
        for (int i = 0; i < PT; i++)
            for (int j = i; j < PT; j++)
                if (present[(size_t)i * PT + j] || i == j) blk_of[(size_t)i * PT + j] = nb++;
    }
    std::vector<int2> blk_ij;
    for (int i = 0; i < PT; i++)
        for (int j = i; j < PT; j++)
            if (blk_of[(size_t)i * PT + j] >= 0) blk_ij.push_back(make_int2(i, j));
    const int nblk = (int)blk_ij.size();
    std::vector<int> blk_ptr(nblk + 1, 0);
    for (int l = l0; l < l1; l++)
        for (int a = lm_ptr[l]; a < lm_ptr[l + 1]; a++) {
            const int fa = g_free[a];
            if (fa < 0) continue;
            for (int b = lm_ptr[l]; b < lm_ptr[l + 1]; b++) {
                const int fb = g_free[b];
                if (fb < fa || (fb == fa && b != a)) continue;
                blk_ptr[blk_of[(size_t)fa * PT + fb] + 1]++;
            }
        }
    for (int k = 0; k < nblk; k++) blk_ptr[k + 1] += blk_ptr[k];
    std::vector<int2> con(blk_ptr[nblk]);
    {
        std::vector<int> bf(nblk, 0);
        for (int l = l0; l < l1; l++)
            for (int a = lm_ptr[l]; a < lm_ptr[l + 1]; a++) {
                const int fa = g_free[a];
                if (fa < 0) continue;
                for (int b = lm_ptr[l]; b < lm_ptr[l + 1]; b++) {
                    const int fb = g_free[b];
                    if (fb < fa || (fb == fa && b != a)) continue;
                    const int k = blk_of[(size_t)fa * PT + fb];
                    con[blk_ptr[k] + bf[k]++] = make_int2(a - o0, b - o0);  // local observation indices
                }
            }
    }
    // local observation lists
    std::vector<int> s_pose(ML), s_lm(ML), lm_ptr_loc(NL + 1);
    for (int k = 0; k < ML; k++) { s_pose[k] = pb->obs_pose[order[o0 + k]]; s_lm[k] = pb->obs_point[order[o0 + k]] - l0; }
    for (int l = 0; l <= NL; l++) lm_ptr_loc[l] = lm_ptr[l0 + l] - o0;
    std::vector<int> pose_ptr(Pf + 1, 0), pose_obs;
    for (int k = 0; k < ML; k++) if (g_free[o0 + k] >= 0) pose_ptr[g_free[o0 + k] + 1]++;
    for (int f = 0; f < Pf; f++) pose_ptr[f + 1] += pose_ptr[f];
    pose_obs.resize(pose_ptr[Pf]);
    {
        std::vector<int> pf(Pf, 0);
        for (int k = 0; k < ML; k++) { const int f = g_free[o0 + k]; if (f >= 0) pose_obs[pose_ptr[f] + pf[f]++] = k; }
    }
    // markers: per free keyframe / per marker the list of marker edges (edge order), per block the edge that fills it
    const int n_mk_entries = Ne + Nx + (xref >= 0 ? Nx : 0);   // a planar edge sits in the list of each of its free markers
    std::vector<int> cam_ptr(Pf + 1, 0), cam_edges, mk_ptr(Nm + 1, 0), mk_edges(n_mk_entries), blk_edge(nblk, -1);
    for (int k = 0; k < Ne; k++) {
        const int fc = free_idx[pb->mobs_pose[k]];
        if (fc >= 0) {
            cam_ptr[fc + 1]++;
            int& be = blk_edge[blk_of[(size_t)fc * PT + Pf + pb->mobs_marker[k]]];
            if (be >= 0) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: marker %d is observed twice by pose %d", pb->mobs_marker[k], pb->mobs_pose[k]);
            be = k;
        }
        mk_ptr[pb->mobs_marker[k] + 1]++;
    }
    for (int x = 0; x < Nx; x++) {
        const int o = pb->plane_other[x];
        mk_ptr[o + 1]++;
        if (xref >= 0) {
            mk_ptr[xref + 1]++;
            int& be = blk_edge[blk_of[(size_t)(Pf + std::min(xref, o)) * PT + Pf + std::max(xref, o)]];
            if (be >= 0) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: marker %d has two planar edges", o);
            be = Ne + x;
        }
    }
    for (int i = 0; i < Pf; i++) cam_ptr[i + 1] += cam_ptr[i];
    for (int i = 0; i < Nm; i++) mk_ptr[i + 1] += mk_ptr[i];
    cam_edges.resize(cam_ptr[Pf]);
    {
        std::vector<int> cf(cam_ptr.begin(), cam_ptr.end() - 1), mf(mk_ptr.begin(), mk_ptr.end() - 1);
        for (int k = 0; k < Ne; k++) {
            const int fc = free_idx[pb->mobs_pose[k]];
            if (fc >= 0) cam_edges[cf[fc]++] = k;
            mk_edges[mf[pb->mobs_marker[k]]++] = k;
        }
        for (int x = 0; x < Nx; x++) {    // after the marker edges, as the reference adds them; ~index = the marker owns the "c" slots
            const int o = pb->plane_other[x];
            if (xref < 0) mk_edges[mf[o]++] = Ne + x;
            else {
                mk_edges[mf[std::min(xref, o)]++] = ~(Ne + x);
                mk_edges[mf[std::max(xref, o)]++] = Ne + x;
            }
        }
    }
    // ---- arena: [inputs][work][full-size result arrays that are summed over the ranks]
    Arena A;
    const size_t o_free_idx = A.take(4 * (size_t)P), o_free_list = A.take(4 * (size_t)(Pf + 1)), o_lm_ptr = A.take(4 * (size_t)(NL + 1)),
                 o_obs_pose = A.take(4 * (size_t)(ML + 1)), o_obs_lm = A.take(4 * (size_t)(ML + 1)), o_pose_ptr = A.take(4 * (size_t)(Pf + 1)),
                 o_pose_obs = A.take(4 * (pose_obs.size() + 1)), o_blk_ptr = A.take(4 * (size_t)(nblk + 1)),
                 o_blk_ij = A.take(8 * (size_t)(nblk + 1)), o_con = A.take(8 * (con.size() + 1)), o_z = A.take(24 * (size_t)(ML + 1)),
                 o_info = A.take(8 * (size_t)(ML + 1)), o_stereo = A.take((size_t)ML + 1), o_active = A.take((size_t)ML + 1),
                 o_pt = A.take(24 * (size_t)(NL + 1)), o_p44 = A.take(64 * (size_t)P);
    const size_t o_mk44 = A.take(64 * (size_t)(Nm + 2)), o_mksz = A.take(4 * (size_t)(Nm + 1)), o_em = A.take(4 * (size_t)(Ne + 1)), o_ep = A.take(4 * (size_t)(Ne + 1)),
                 o_ec = A.take(32 * (size_t)(Ne + 1)), o_ew = A.take(4 * (size_t)(Ne + 1)), o_cptr = A.take(4 * (size_t)(Pf + 2)), o_cedg = A.take(4 * (cam_edges.size() + 1)),
                 o_mptr = A.take(4 * (size_t)(Nm + 2)), o_medg = A.take(4 * (size_t)(n_mk_entries + 1)), o_bedg = A.take(4 * (size_t)(nblk + 1));
    const size_t o_xoth = A.take(4 * (size_t)(Nx + 1));
    const size_t o_cams = A.take(pb->pose_cam ? sizeof(Cam) * (size_t)P : 0);   // one camera per keyframe (mixed-camera windows)
    const size_t in_bytes = A.off;
    const size_t o_mkpose = A.take(56 * (size_t)(Nm + 2)), o_mkbak = A.take(56 * (size_t)(Nm + 2)), o_echi = A.take(8 * (size_t)(Ne + Nx + 1)), o_eblk = A.take(960 * (size_t)(Ne + Nx + 1)),
                 o_mk44o = A.take(64 * (size_t)(Nm + 1));
    const size_t n_red = 36 * (size_t)nblk + (size_t)n;  // packed Schur blocks | right-hand side: the per-trial all-reduce payload
    const size_t o_pose = A.take(56 * (size_t)P), o_pose_bak = A.take(56 * (size_t)P), o_pt_bak = A.take(24 * (size_t)(NL + 1)),
                 o_err = A.take(24 * (size_t)(ML + 1)), o_chi2 = A.take(8 * (size_t)(ML + 1)), o_rho0 = A.take(8 * (size_t)(ML + 1)),
                 o_Hll = A.take(48 * (size_t)(NL + 1)), o_bl = A.take(24 * (size_t)(NL + 1)), o_Hpl = A.take(144 * (size_t)(ML + 1)),
                 o_Y = A.take(144 * (size_t)(ML + 1)), o_Dinv = A.take(48 * (size_t)(NL + 1)), o_db = A.take(24 * (size_t)(NL + 1)),
                 o_xl = A.take(24 * (size_t)(NL + 1)), o_HppBp = A.take(8 * (42 * (size_t)PT + 6)),
                 o_S = A.take(8 * ((n > 1023 ? 0 : (size_t)n * n) + 1)), o_bs = A.take(8 * (size_t)(n + 1)), o_xp = A.take(8 * (size_t)(n + 1)),
                 o_scl = A.take(8 * (size_t)(NL + 1)), o_scp = A.take(8 * (size_t)(PT + 1)), o_st = A.take(sizeof(LmState)),
                 o_p44o = A.take(64 * (size_t)P), o_bad = A.take((size_t)ML + 1), o_red = A.take(8 * (n_red + 1)), o_sums = A.take(64),
                 o_info2 = A.take(16);
    const size_t o_full_pt = A.take(24 * (size_t)(N + 1)), o_full_chi = A.take(8 * (size_t)(M + 1)), o_full_flags = A.take(2 * (size_t)M + 2);
    const bool big = n > 1023;   // beyond the single-CTA dense solver: block-envelope Cholesky over the RCM-ordered block graph (ba_band.cu)
    size_t o_chol = 0, o_bblob = 0, o_bZ = 0, o_bwin = 0;
    uco_band_plan band;
    if (!ctx->ba->smem_optin) cudaDeviceGetAttribute(&ctx->ba->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    uint8_t* d = nullptr;
    if (big) {
        uco_band_make_plan(n / 6, nblk, blk_ij.data(), ctx->ba->smem_optin, -1, band);
        o_bblob = A.take(band.blob.size() + 16);
        o_bZ = A.take(8 * (size_t)band.z_doubles + 16);
        o_bwin = A.take(8 * band.scratch_doubles + 16);
    } else {
        o_chol = A.take(8 * ((size_t)(n + 1) * (n + 2) / 2 + 1));
    }
    d = (uint8_t*)uco_ws(ctx, WS_BA, A.off);
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_BA, in_bytes + sizeof(LmState) + 64);
    if (!d || !h) return UCO_E_NOMEM;
    if (pb->pose_cam)
        for (int p = 0; p < P; p++) {
            const float* c = pb->pose_cam + 5 * (size_t)p;
            Cam k;
            k.fx = c[0]; k.fy = c[1]; k.cx = c[2]; k.cy = c[3]; k.bf = c[4]; k.bf_f = c[4];
            memcpy(h + o_cams + sizeof(Cam) * (size_t)p, &k, sizeof(Cam));
        }
    memcpy(h + o_free_idx, free_idx.data(), 4 * (size_t)P);
    memcpy(h + o_free_list, free_list.data(), 4 * (size_t)Pf);
    memcpy(h + o_lm_ptr, lm_ptr_loc.data(), 4 * (size_t)(NL + 1));
    memcpy(h + o_obs_pose, s_pose.data(), 4 * (size_t)ML);
    memcpy(h + o_obs_lm, s_lm.data(), 4 * (size_t)ML);
    memcpy(h + o_pose_ptr, pose_ptr.data(), 4 * (size_t)(Pf + 1));
    memcpy(h + o_pose_obs, pose_obs.data(), 4 * pose_obs.size());
    memcpy(h + o_blk_ptr, blk_ptr.data(), 4 * (size_t)(nblk + 1));
    memcpy(h + o_blk_ij, blk_ij.data(), 8 * (size_t)nblk);
    memcpy(h + o_con, con.data(), 8 * con.size());
    {
        double* z = (double*)(h + o_z);
        double* info = (double*)(h + o_info);
        uint8_t* st = h + o_stereo;
        for (int k = 0; k < ML; k++) {
            const int i = order[o0 + k];
            const bool s = pb->obs_stereo && pb->obs_stereo[i];
            z[3 * k] = pb->obs_uv[2 * i]; z[3 * k + 1] = pb->obs_uv[2 * i + 1]; z[3 * k + 2] = s ? pb->obs_ur[i] : 0.0;
            info[k] = pb->obs_inv_sigma2[i];
            st[k] = s;
        }
        memset(h + o_active, 1, (size_t)ML);
        double* pt = (double*)(h + o_pt);
        for (int k = 0; k < 3 * NL; k++) pt[k] = pb->points3[3 * (size_t)l0 + k];
        memcpy(h + o_p44, pb->poses44, 64 * (size_t)P);
    }
    if (Nm) {
        memcpy(h + o_mk44, pb->marker_pose44, 64 * (size_t)Nm);
        memcpy(h + o_mksz, pb->marker_size, 4 * (size_t)Nm);
    }
    if (Nx) {
        memcpy(h + o_xoth, pb->plane_other, 4 * (size_t)Nx);
        if (xref < 0) memcpy(h + o_mk44 + 64 * (size_t)Nm, pb->plane_ref_pose44, 64);   // the fixed reference marker: slot Nm, never updated
    }
    if (n_mk_entries) memcpy(h + o_medg, mk_edges.data(), 4 * (size_t)n_mk_entries);
    if (Ne) {
        memcpy(h + o_em, pb->mobs_marker, 4 * (size_t)Ne);
        memcpy(h + o_ep, pb->mobs_pose, 4 * (size_t)Ne);
        memcpy(h + o_ec, pb->mobs_corners, 32 * (size_t)Ne);
        memcpy(h + o_ew, pb->mobs_weight, 4 * (size_t)Ne);
        if (!cam_edges.empty()) memcpy(h + o_cedg, cam_edges.data(), 4 * cam_edges.size());
    }
    memcpy(h + o_cptr, cam_ptr.data(), 4 * (size_t)(Pf + 1));
    memcpy(h + o_mptr, mk_ptr.data(), 4 * (size_t)(Nm + 1));
    if (nblk) memcpy(h + o_bedg, blk_edge.data(), 4 * (size_t)nblk);
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_st, 0, sizeof(LmState), s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_err, 0, 24 * (size_t)(ML + 1), s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_chi2, 0, 8 * (size_t)(ML + 1), s));
    if (!big) UCO_CUDA(ctx, cudaMemsetAsync(d + o_S, 0, 8 * ((size_t)n * n + 1), s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_red, 0, 8 * (n_red + 1), s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_full_pt, 0, o_full_flags + 2 * (size_t)M + 2 - o_full_pt, s));
    BaDev B;
    B.P = P; B.N = NL; B.M = ML; B.Pf = Pf; B.n = n; B.nblk = nblk;
    B.pose = (double*)(d + o_pose); B.pose_bak = (double*)(d + o_pose_bak); B.pt = (double*)(d + o_pt); B.pt_bak = (double*)(d + o_pt_bak);
    B.free_idx = (int*)(d + o_free_idx); B.free_list = (int*)(d + o_free_list); B.lm_ptr = (int*)(d + o_lm_ptr);
    B.obs_pose = (int*)(d + o_obs_pose); B.obs_lm = (int*)(d + o_obs_lm); B.pose_ptr = (int*)(d + o_pose_ptr); B.pose_obs = (int*)(d + o_pose_obs);
    B.z = (double*)(d + o_z); B.info = (double*)(d + o_info); B.stereo = d + o_stereo; B.active = d + o_active;
    B.err = (double*)(d + o_err); B.chi2 = (double*)(d + o_chi2); B.rho0 = (double*)(d + o_rho0); B.Hll = (double*)(d + o_Hll);
    B.bl = (double*)(d + o_bl); B.Hpl = (double*)(d + o_Hpl); B.Y = (double*)(d + o_Y); B.Dinv = (double*)(d + o_Dinv); B.db = (double*)(d + o_db);
    B.xl = (double*)(d + o_xl);
    B.Hpp = (double*)(d + o_HppBp); B.bp = B.Hpp + 36 * (size_t)PT;   // contiguous: one all-reduce per outer iteration
    B.mk.Nm = Nm; B.mk.Ne = Ne;
    B.mk.Nx = Nx; B.mk.x_ref = xref; B.mk.x_other = (const int*)(d + o_xoth); B.mk.x_w = pb->plane_weight;
    B.mk.pose = (double*)(d + o_mkpose); B.mk.pose_bak = (double*)(d + o_mkbak); B.mk.size = (const float*)(d + o_mksz);
    B.mk.e_marker = (const int*)(d + o_em); B.mk.e_pose = (const int*)(d + o_ep); B.mk.e_corners = (const float*)(d + o_ec);
    B.mk.e_weight = (const float*)(d + o_ew); B.mk.e_chi2 = (double*)(d + o_echi); B.mk.e_blk = (double*)(d + o_eblk);
    B.mk.cam_ptr = (const int*)(d + o_cptr); B.mk.cam_edges = (const int*)(d + o_cedg); B.mk.mk_ptr = (const int*)(d + o_mptr);
    B.mk.mk_edges = (const int*)(d + o_medg); B.mk.blk_edge = Ne + Nx ? (const int*)(d + o_bedg) : nullptr;
    B.S = (double*)(d + o_S); B.bs = (double*)(d + o_bs);
    B.xp = (double*)(d + o_xp); B.scale_lm = (double*)(d + o_scl); B.scale_pose = (double*)(d + o_scp);
    B.blk_ptr = (int*)(d + o_blk_ptr); B.blk_ij = (int2*)(d + o_blk_ij); B.con = (int2*)(d + o_con);
    B.st = (LmState*)(d + o_st);
    B.cam.fx = pb->fx; B.cam.fy = pb->fy; B.cam.cx = pb->cx; B.cam.cy = pb->cy; B.cam.bf = pb->bf; B.cam.bf_f = pb->bf;
    if (pb->pose_cam) B.cams = (const Cam*)(d + o_cams);
    B.chi2d = 5.99f; B.chi3d = 7.815f;
    B.d2 = (double)sqrtf(B.chi2d); B.d3 = (double)sqrtf(B.chi3d);
    double* Sp = (double*)(d + o_red);
    double* bsp = Sp + 36 * (size_t)nblk;
    double* sums = (double*)(d + o_sums);      // [0..3] sums, [4] max
    int* info2 = (int*)(d + o_info2);
    (void)info2;
    if (big) {   // the ordering, the fronts and the storage map of the reduced system
        UCO_CUDA(ctx, cudaMemcpyAsync(d + o_bblob, band.blob.data(), band.blob.size(), cudaMemcpyHostToDevice, ctx->stream));
        UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    LmState* hst = (LmState*)(h + in_bytes);
    double* hsums = (double*)(h + in_bytes + sizeof(LmState));
    const size_t chol_bytes = 8 * ((size_t)(n + 1) * (n + 2) / 2);
    const int chol_smem = !big && chol_bytes + 16 * 1024 <= (size_t)ctx->ba->smem_optin;
    if (chol_smem) UCO_CUDA(ctx, cudaFuncSetAttribute(ba_chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_bytes));
    const int gM = (ML + 255) / 256, gN = (NL + 127) / 128, gU = (NL + Pf + 127) / 128;
    auto stop_now = [&] { return stop && *stop ? 1 : 0; };
    // all-reduce of the LM sums (+ the ranks' stop flags); with first = true also the max |diagonal| for lambda_0
    auto reduce_sums = [&](bool first) -> int {
        ba_partial_sums_kernel<<<1, 1024, 0, s>>>(B, sums, first ? sums + 4 : nullptr, stop_now(), rank == 0);
        UCO_LAUNCH_CHECK(ctx);
        int r2 = uco_comm_allreduce(comm, sums, sums, 4, 0, s);
        if (r2 != UCO_OK) return r2;
        if (first) {
            r2 = uco_comm_allreduce(comm, sums + 4, sums + 2, 1, 1, s);  // max lands in ext[2]
            if (r2 != UCO_OK) return r2;
        }
        return UCO_OK;
    };

    const auto t_prep = std::chrono::steady_clock::now();
    UCO_CUDA(ctx, cudaEventRecord(ctx->ba->ev0, s));
    ba_init_poses_kernel<<<(P + 127) / 128, 128, 0, s>>>(B, (const float*)(d + o_p44));
    UCO_LAUNCH_CHECK(ctx);
    if (Nm) {  // marker vertices: Marker::pose_g2m -> SE3Quat, like the keyframes
        BaDev Bm = B;
        Bm.P = Nm + (Nx && xref < 0 ? 1 : 0); Bm.pose = B.mk.pose; Bm.pose_bak = B.mk.pose_bak;
        ba_init_poses_kernel<<<(Bm.P + 127) / 128, 128, 0, s>>>(Bm, (const float*)(d + o_mk44));
        UCO_LAUNCH_CHECK(ctx);
    }
    const int gE = (Ne + Nx + 63) / 64, gPT = (PT + 127) / 128;
    int iters[2] = {0, 0};
    bool stopped = false;
    for (int stage = 0; stage < 2 && !stopped; stage++) {
        const int robust = stage == 0, max_iters = stage == 0 ? pb->n_iters : 2 * pb->n_iters;
        if (stage == 1 && ML) {
            ba_flag_outliers_kernel<<<gM, 256, 0, s>>>(B);
            UCO_LAUNCH_CHECK(ctx);
        }
        if (ML) {
            ba_errors_kernel<<<gM, 256, 0, s>>>(B, robust);
            UCO_LAUNCH_CHECK(ctx);
        }
        bool cont_iter = max_iters > 0;
        for (int it = 0; cont_iter; it++) {
            if (NL) {
                ba_linearize_lm_kernel<<<gN, 128, 0, s>>>(B, robust);
                UCO_LAUNCH_CHECK(ctx);
            }
            if (Pf) {
                ba_linearize_pose_kernel<<<Pf, POSE_THREADS, 0, s>>>(B, robust);
                UCO_LAUNCH_CHECK(ctx);
            }
            if (PT) {
                if (Nm) {  // marker rows of Hpp | bp hold no keypoint terms: cleared so that the all-reduce leaves them zero
                    UCO_CUDA(ctx, cudaMemsetAsync(B.Hpp + 36 * (size_t)Pf, 0, 8 * 36 * (size_t)Nm, s));
                    UCO_CUDA(ctx, cudaMemsetAsync(B.bp + 6 * (size_t)Pf, 0, 8 * 6 * (size_t)Nm, s));
                }
                if ((rc = uco_comm_allreduce(comm, B.Hpp, B.Hpp, 42 * (size_t)PT, 0, s)) != UCO_OK) return rc;
                if (Ne + Nx) {  // marker edges: replicated on every rank, added after the exchange
                    ba_marker_kernel<<<gE, 64, 0, s>>>(B, 1);
                    UCO_LAUNCH_CHECK(ctx);
                    ba_marker_accumulate_kernel<<<gPT, 128, 0, s>>>(B);
                    UCO_LAUNCH_CHECK(ctx);
                }
            }
            if (it == 0) {
                if ((rc = reduce_sums(true)) != UCO_OK) return rc;
                UCO_CUDA(ctx, cudaMemcpyAsync(hsums, sums, 32, cudaMemcpyDeviceToHost, s));
                UCO_CUDA(ctx, cudaStreamSynchronize(s));
                if (hsums[3] != 0.0) {  // some rank saw stopASAP before the stage started: every rank leaves together
                    stopped = true;
                    break;
                }
            }
            ba_iter_begin_kernel<<<1, 1024, 0, s>>>(B, it == 0, sums);
            UCO_LAUNCH_CHECK(ctx);
            bool cont_trial = true;
            while (cont_trial) {
                if (NL) {
                    ba_prep_kernel<<<gN, 128, 0, s>>>(B);
                    UCO_LAUNCH_CHECK(ctx);
                }
                if (PT) {
                    ba_schur_gather_packed_kernel<<<nblk, 36 * GATHER_CHUNKS, 0, s>>>(B, Sp, bsp);
                    UCO_LAUNCH_CHECK(ctx);
                    if ((rc = uco_comm_allreduce(comm, Sp, Sp, n_red, 0, s)) != UCO_OK) return rc;   // THE exchange step of the path
                    if (big) {   // two-level block-envelope Cholesky: assemble straight into the fronts, factor, solve (4 launches)
                        if ((rc = uco_band_solve_launch(ctx, band, d + o_bblob, (double*)(d + o_bZ), band.scratch_doubles ? (double*)(d + o_bwin) : nullptr,
                                                        B.blk_ij, B.Hpp, &B.st->lambda, Sp, B.bp, bsp, B.mk.blk_edge, B.mk.e_blk, B.xp, &B.st->chol_fail,
                                                        ctx->ba->smem_optin)) != UCO_OK) return rc;
                    } else {
                        ba_assemble_kernel<<<nblk, 36, 0, s>>>(B, Sp, bsp);
                        UCO_LAUNCH_CHECK(ctx);
                        ba_chol_solve_kernel<<<1, CHOL_THREADS, chol_smem ? chol_bytes : 0, s>>>(B, (double*)(d + o_chol), chol_smem);
                        UCO_LAUNCH_CHECK(ctx);
                    }
                }
                ba_update_kernel<<<gU, 128, 0, s>>>(B);
                UCO_LAUNCH_CHECK(ctx);
                if (Nm) {
                    ba_marker_update_kernel<<<(Nm + 127) / 128, 128, 0, s>>>(B);
                    UCO_LAUNCH_CHECK(ctx);
                }
                if (ML) {
                    ba_errors_kernel<<<gM, 256, 0, s>>>(B, robust);
                    UCO_LAUNCH_CHECK(ctx);
                }
                if (Ne + Nx) {
                    ba_marker_kernel<<<gE, 64, 0, s>>>(B, 0);
                    UCO_LAUNCH_CHECK(ctx);
                }
                if ((rc = reduce_sums(false)) != UCO_OK) return rc;
                ba_decide_kernel<<<1, 1024, 0, s>>>(B, 0, max_iters, sums);
                UCO_LAUNCH_CHECK(ctx);
                UCO_CUDA(ctx, cudaMemcpyAsync(hst, B.st, offsetof(LmState, trace), cudaMemcpyDeviceToHost, s));
                UCO_CUDA(ctx, cudaMemcpyAsync(hsums, sums, 32, cudaMemcpyDeviceToHost, s));
                UCO_CUDA(ctx, cudaStreamSynchronize(s));
                cont_trial = hst->cont_trial != 0;
            }
            cont_iter = hst->cont_iter != 0;
            iters[stage] = hst->it;
            if (hsums[3] != 0.0) stopped = true;
        }
    }
    // ---- results: poses are replicated; points / per-observation values are written into full-size arrays at this rank's
    // range and summed over the ranks (every rank returns the complete result)
    float* p44o = (float*)(d + o_p44o);
    uint8_t* bad = d + o_bad;
    ba_results_kernel<<<(P + 255) / 256, 256, 0, s>>>(B, (const float*)(d + o_p44), p44o, bad);
    UCO_LAUNCH_CHECK(ctx);
    if (ML) {
        ba_bad_kernel<<<gM, 256, 0, s>>>(B, p44o, bad);
        UCO_LAUNCH_CHECK(ctx);
    }
    if (Nm) {
        ba_marker_results_kernel<<<(Nm + 127) / 128, 128, 0, s>>>(B, (float*)(d + o_mk44o));
        UCO_LAUNCH_CHECK(ctx);
    }
    double* full_pt = (double*)(d + o_full_pt);
    double* full_chi = (double*)(d + o_full_chi);
    uint8_t* full_flags = d + o_full_flags;  // [0, M): active, [M, 2M): bad
    if (NL) UCO_CUDA(ctx, cudaMemcpyAsync(full_pt + 3 * (size_t)l0, B.pt, 24 * (size_t)NL, cudaMemcpyDeviceToDevice, s));
    if (ML) {
        UCO_CUDA(ctx, cudaMemcpyAsync(full_chi + o0, B.chi2, 8 * (size_t)ML, cudaMemcpyDeviceToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(full_flags + o0, B.active, (size_t)ML, cudaMemcpyDeviceToDevice, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(full_flags + (size_t)M + o0, bad, (size_t)ML, cudaMemcpyDeviceToDevice, s));
    }
    if (R > 1) {
        if ((rc = uco_comm_allreduce(comm, full_pt, full_pt, 3 * (size_t)N, 0, s)) != UCO_OK) return rc;
        if ((rc = uco_comm_allreduce(comm, full_chi, full_chi, (size_t)M, 0, s)) != UCO_OK) return rc;
        if ((rc = uco_comm_allreduce(comm, full_flags, full_flags, 2 * (size_t)M, 2, s)) != UCO_OK) return rc;
    }
    UCO_CUDA(ctx, cudaEventRecord(ctx->ba->ev1, s));
    const auto t_loop = std::chrono::steady_clock::now();
    const size_t o_h_pose = 0, o_h_p44 = o_h_pose + 56 * (size_t)P, o_h_pt = o_h_p44 + 64 * (size_t)P, o_h_chi = o_h_pt + 24 * (size_t)N,
                 o_h_fl = o_h_chi + 8 * (size_t)M, o_h_st = ((o_h_fl + 2 * (size_t)M + 7) & ~(size_t)7), o_h_mk7 = o_h_st + ((sizeof(LmState) + 7) & ~(size_t)7),
                 o_h_mk44 = o_h_mk7 + 56 * (size_t)Nm, o_h_echi = o_h_mk44 + 64 * (size_t)Nm, out_bytes = o_h_echi + 8 * (size_t)Ne + 8;
    uint8_t* ho = (uint8_t*)uco_pinned(ctx, WS_BA_OUT, out_bytes);
    if (!ho) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_pose, B.pose, 56 * (size_t)P, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_p44, p44o, 64 * (size_t)P, cudaMemcpyDeviceToHost, s));
    if (N) UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_pt, full_pt, 24 * (size_t)N, cudaMemcpyDeviceToHost, s));
    if (M) {
        UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_chi, full_chi, 8 * (size_t)M, cudaMemcpyDeviceToHost, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_fl, full_flags, 2 * (size_t)M, cudaMemcpyDeviceToHost, s));
    }
    UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_st, B.st, sizeof(LmState), cudaMemcpyDeviceToHost, s));
    if (Nm) {
        UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_mk7, B.mk.pose, 56 * (size_t)Nm, cudaMemcpyDeviceToHost, s));
        UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_mk44, d + o_mk44o, 64 * (size_t)Nm, cudaMemcpyDeviceToHost, s));
    }
    if (Ne) UCO_CUDA(ctx, cudaMemcpyAsync(ho + o_h_echi, B.mk.e_chi2, 8 * (size_t)Ne, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    if (res->marker_pose7 && Nm) memcpy(res->marker_pose7, ho + o_h_mk7, 56 * (size_t)Nm);
    if (res->marker_poses44 && Nm) memcpy(res->marker_poses44, ho + o_h_mk44, 64 * (size_t)Nm);
    if (res->mobs_chi2 && Ne) memcpy(res->mobs_chi2, ho + o_h_echi, 8 * (size_t)Ne);
    if (res->pose7) memcpy(res->pose7, ho + o_h_pose, 56 * (size_t)P);
    if (res->poses44) memcpy(res->poses44, ho + o_h_p44, 64 * (size_t)P);
    if (res->points3) memcpy(res->points3, ho + o_h_pt, 24 * (size_t)N);
    const double* hchi = (const double*)(ho + o_h_chi);
    for (int k = 0; k < M; k++) {
        const int i = order[k];
        if (res->obs_chi2) res->obs_chi2[i] = hchi[k];
        if (res->obs_level) res->obs_level[i] = !(ho + o_h_fl)[k];
        if (res->obs_bad) res->obs_bad[i] = (ho + o_h_fl)[(size_t)M + k];
    }
    const LmState* fst = (const LmState*)(ho + o_h_st);
    if (res->trace) {
        memset(res->trace, 0, sizeof(double) * 128);
        memcpy(res->trace, fst->trace, sizeof(double) * 2 * (size_t)std::min(fst->ntrace, 64));
    }
    res->iters[0] = iters[0];
    res->iters[1] = iters[1];
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ba->ev0, ctx->ba->ev1);
    res->device_ms = ms;
    if (res->profile) {
        memset(res->profile, 0, sizeof(double) * 16);
        res->profile[0] = NL; res->profile[1] = ML; res->profile[2] = nblk; res->profile[3] = (double)n_red * 8;  // shard sizes, all-reduce bytes per trial
        auto ms_between = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        res->profile[4] = ms_between(t_enter, t_prep);   // host: structure, solver plan, staging, upload
        res->profile[5] = ms_between(t_prep, t_loop);    // host wall time of the LM loop (kernel launches + one synchronisation per trial)
        res->profile[6] = ms_between(t_loop, std::chrono::steady_clock::now());   // download + scatter
    }
    return UCO_OK;
}

extern "C" {

// host-only: builds the window structure (no device needed) and reports its sizes; used by the CPU tests and to time the planner
int uco_b200_probe_ba_plan(const uco_ba_problem* pb, int cluster_size, int* out8) {
    uco_b200_ctx tmp;
    BaPlan p;
    const bool as_cluster = cluster_size >= 1000;   // 1000 + c: exactly the plan the cluster kernel gets (16 warps: fused prep + gather lists for small windows)
    if (as_cluster) cluster_size -= 1000;
    int rc = ba_plan_build(&tmp, *pb, BA_UNIT, p, cluster_size > 0 ? cluster_size : 8, 512, as_cluster ? 16 : 0);
    if (rc != UCO_OK) return rc;
    if (out8) {
        out8[0] = p.Pf; out8[1] = (int)p.blk_ij.size(); out8[2] = (int)p.unit.size(); out8[3] = p.ncon;
        out8[4] = (int)p.chunk_lm.size() - 1; out8[5] = (int)p.pose_obs.size();
        int mx = 0;
        for (size_t c = 0; c + 1 < p.chunk_lm.size(); c++) mx = std::max(mx, p.lm_ptr[p.chunk_lm[c + 1]] - p.lm_ptr[p.chunk_lm[c]]);
        out8[6] = mx;
        out8[7] = p.cta_lm.back();
    }
    return UCO_OK;
}

int uco_b200_ba_set_mode(uco_b200_ctx* ctx, int mode, int cluster_size) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (mode < 0 || mode > 2 || cluster_size < 0 || cluster_size > 16 || (cluster_size & (cluster_size - 1)))
        return uco_fail(ctx, UCO_E_INVALID, "ba_set_mode: mode 0..2, cluster size a power of two <= 16");
    ctx->ba_mode = mode;
    ctx->ba_cluster_size = cluster_size;
    return UCO_OK;
}

int uco_b200_ba_set_host_threads(uco_b200_ctx* ctx, int n_threads) {
    if (!ctx || n_threads < 0 || n_threads > 64) return UCO_E_INVALID;
    ctx->ba_host_threads = n_threads;
    return UCO_OK;
}

int uco_b200_ba_solve_batch(uco_b200_ctx* ctx, int n, const uco_ba_problem* pbs, const volatile unsigned char* stop, uco_ba_result* res) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (n < 0 || (n && (!pbs || !res))) return uco_fail(ctx, UCO_E_INVALID, "ba_solve_batch: bad arguments");
    std::vector<const uco_ba_problem*> cp;
    std::vector<uco_ba_result*> cr;
    for (int i = 0; i < n; i++) {
        int rc = ba_validate(ctx, pbs + i);
        if (rc != UCO_OK) return rc;
        if (pbs[i].n_markers > 0 || pbs[i].pose_cam) {  // ArUco markers / one camera per keyframe: the sharded solver (any size, one rank) handles them
            rc = ba_sharded_solve(ctx, nullptr, pbs + i, stop, res + i);
            if (rc != UCO_OK) return rc;
            continue;
        }
        const bool fits = 6 * ba_free_poses(pbs + i) <= BA_CLUSTER_MAX_N;
        if (ctx->ba_mode == 2 && !fits)
            return uco_fail(ctx, UCO_E_INVALID, "ba_solve: %d free poses exceed the cluster-resident solver (%d)", ba_free_poses(pbs + i),
                            BA_CLUSTER_MAX_N / 6);
        if (ctx->ba_mode != 1 && fits) {
            cp.push_back(pbs + i);
            cr.push_back(res + i);
        } else {
            rc = ba_streamed_solve(ctx, pbs + i, stop, res + i);
            if (rc != UCO_OK) return rc;
        }
    }
    const int per_launch = std::max(1, ctx->sm_count / (ctx->ba_cluster_size > 0 ? ctx->ba_cluster_size : 8));  // co-resident clusters
    for (size_t k = 0; k < cp.size(); k += per_launch) {
        int m = (int)std::min(cp.size() - k, (size_t)per_launch);
        int rc = ba_cluster_solve_batch(ctx, m, cp.data() + k, stop, cr.data() + k);
        if (rc != UCO_OK) return rc;
    }
    return UCO_OK;
}

int uco_b200_ba_solve_sharded(uco_b200_ctx* ctx, uco_b200_comm* comm, const uco_ba_problem* pb, const volatile unsigned char* stop,
                              uco_ba_result* res) {
    UCO_RANGE();
    return ba_sharded_solve(ctx, comm, pb, stop, res);
}

// host-only: the landmark ranges the sharded solver gives each of `world` ranks: out[0..world] boundaries, out[world+1 ..] observation counts
int uco_b200_probe_ba_partition(const uco_ba_problem* pb, int world, int* out) {
    uco_b200_ctx tmp;
    if (ba_validate(&tmp, pb) != UCO_OK || world < 1 || !out) return UCO_E_INVALID;
    const int N = pb->n_points, M = pb->n_obs;
    std::vector<int> lm_ptr(N + 1, 0), L;
    for (int i = 0; i < M; i++) lm_ptr[pb->obs_point[i] + 1]++;
    for (int l = 0; l < N; l++) lm_ptr[l + 1] += lm_ptr[l];
    ba_partition_landmarks(lm_ptr, N, M, world, L);
    for (int r = 0; r <= world; r++) out[r] = L[r];
    for (int r = 0; r < world; r++) out[world + 1 + r] = lm_ptr[L[r + 1]] - lm_ptr[L[r]];
    return UCO_OK;
}

int uco_b200_ba_solve(uco_b200_ctx* ctx, const uco_ba_problem* pb, const volatile unsigned char* stop, uco_ba_result* res) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (!pb || !res) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: null problem / result");
    return uco_b200_ba_solve_batch(ctx, 1, pb, stop, res);
}

}  // extern "C"
