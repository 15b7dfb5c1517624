// pnp.cu — K14: pose-only optimisation of the tracker (PnPSolver::solvePnp, src/optimization/pnpsolver.cpp:116-408).
//
// One thread block per problem (frame); a batch of frames is one launch.  The whole schedule of the reference — 4 rounds
// of optimize(10) with the estimate reset to the initial pose before every round (:358), inlier re-classification after
// each (:364-374), kernels dropped after round index 2, early exit below 10 inliers (:383) — and g2o's Levenberg-Marquardt
// controller (optimization_algorithm_levenberg.cpp:58-150, sparse_optimizer.cpp:366-436: float chi2-difference stop test
// with minChi2BetweenIter = 0) run inside the kernel: no host round trip between LM trials.
//
// Data flow per LM iteration: `linearize` = every thread walks its strided share of the matches (residual, chi2, Huber
// weight, analytic 2x6 / 3x6 Jacobian, 21 + 6 + 1 running sums in registers), marker edges (8-dim residual, NUMERIC Jacobian
// by central differences with delta = 1e-4f as base_binary_edge.hpp:167-232 does) are evaluated by 16-lane groups; partial
// sums are combined in a fixed order (xor-butterfly inside a warp, then warp 0..7, then marker group 0..15), so results
// are bitwise reproducible.  Thread 0 solves the damped 6x6 system and applies exp(dx)*T; `errors` re-evaluates chi2.
// The match arrays (36 B per match) are read from L2 on every pass; per-match chi2 lives in a global scratch array because
// the classification after each round needs the values of the LAST evaluation (stale after a rejected trial, as in g2o).
#include "common.cuh"
#include "ba_math.cuh"
#include "pnp_dev.cuh"
#include <float.h>
#include <math.h>
#include <string.h>

namespace {

constexpr int PNP_THREADS = 256;
constexpr int PNP_WARPS = PNP_THREADS / 32;
constexpr int PNP_GROUP = 16;  // lanes cooperating on one marker edge
constexpr int PNP_GROUPS = PNP_THREADS / PNP_GROUP;
constexpr int NACC = 28;       // 21 (upper triangle of H) + 6 (b) + 1 (robust chi2)

struct PnpShared {
    ba::Pose T, Tbak, T0;
    double H[21], b[6], x[6];
    double chi;                     // result of the last reduction
    double red[PNP_WARPS][NACC];
    double mH[PNP_GROUPS][NACC];
    double mE[PNP_GROUPS][13][8];
    double mJ[PNP_GROUPS][48];
    int flag, good;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// residual of match i at pose T (typesg2o.h:640-652 mono; :528-532,572-580 stereo: 1/z kept in a float, bf a double member)
__device__ __forceinline__ double edge_error(const PnpArrays& A, const PnpHead& h, const ba::Pose& T, int gi, bool stereo, double* e,
                                            double* p) {
    double X[3] = {(double)A.pts[3 * gi], (double)A.pts[3 * gi + 1], (double)A.pts[3 * gi + 2]};
    ba::se3_map(T, X, p);
    const double is = (double)A.isig[gi];
    if (!stereo) {
        e[0] = (double)A.uv[2 * gi] - ((p[0] / p[2]) * h.fx + h.cx);
        e[1] = (double)A.uv[2 * gi + 1] - ((p[1] / p[2]) * h.fy + h.cy);
        e[2] = 0;
        return is * (e[0] * e[0]) + is * (e[1] * e[1]);
    }
    const float invz = (float)(1.0 / p[2]);
    double r0 = p[0] * invz * h.fx + h.cx, r1 = p[1] * invz * h.fy + h.cy;
    double r2 = r0 - h.bf * invz;
    e[0] = (double)A.uv[2 * gi] - r0; e[1] = (double)A.uv[2 * gi + 1] - r1; e[2] = (double)A.ur[gi] - r2;
    return is * (e[0] * e[0]) + is * (e[1] * e[1]) + is * (e[2] * e[2]);
}

__device__ __forceinline__ ba::Pose pose_mul(const ba::Pose& a, const ba::Pose& b) {  // se3quat.h:156-163
    ba::Pose r;
    double rt[3];
    ba::quat_rot(a.q, b.t, rt);
    const double *x = a.q, *y = b.q;
    r.q[3] = x[3] * y[3] - x[0] * y[0] - x[1] * y[1] - x[2] * y[2];
    r.q[0] = x[3] * y[0] + x[0] * y[3] + x[1] * y[2] - x[2] * y[1];
    r.q[1] = x[3] * y[1] + x[1] * y[3] + x[2] * y[0] - x[0] * y[2];
    r.q[2] = x[3] * y[2] + x[2] * y[3] + x[0] * y[1] - x[1] * y[0];
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    ba::quat_normalize(r.q);
    return r;
}

// MarkerEdgeOnlyProject::computeError (typesg2o.h:440-468): corners through camera*marker, projections narrowed to float
__device__ void marker_error(const PnpArrays& A, const PnpHead& h, const ba::Pose& T, int gm, double* e) {
    ba::Pose g2m = ba::pose_from_m44f(A.mpose + 16 * gm);
    ba::Pose c2m = pose_mul(T, g2m);
    const float s = A.msize[gm];
    const float hp = (float)(s / 2.), hn = (float)(-s / 2.);  // Marker::get3DPointsLocalRefSystem, marker.cpp:58-62
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double c[3] = {(double)((i == 0 || i == 3) ? hn : hp), (double)((i < 2) ? hp : hn), 0.0}, p[3];
        ba::se3_map(c2m, c, p);
        float projx = (float)((p[0] / p[2]) * h.fx + h.cx);
        float projy = (float)((p[1] / p[2]) * h.fy + h.cy);
        e[2 * i] = (double)A.mobs[8 * gm + 2 * i] - (double)projx;
        e[2 * i + 1] = (double)A.mobs[8 * gm + 2 * i + 1] - (double)projy;
    }
}

// computeActiveErrors + activeRobustChi2 at S.T; every thread returns with S.chi valid
__device__ void pass_errors(const PnpArrays& A, const PnpHead& h, PnpShared& S, bool robust) {
    const ba::Pose T = S.T;
    const double d2 = (double)sqrtf(5.99f), d3 = (double)sqrtf(7.815f), d8 = (double)sqrtf(15.507f);
    double part = 0;
    for (int i = threadIdx.x; i < h.n; i += PNP_THREADS) {
        const int gi = h.off + i;
        if (!A.active[gi]) continue;
        const uint8_t f = A.flg[gi];
        double e[3], p[3];
        double c = edge_error(A, h, T, gi, f & 1, e, p);
        A.chi2[gi] = c;
        if (robust) {
            double r0, r1;
            ba::huber(c, (f & 1) ? d3 : d2, (f & 2 ? 1.0 : 0.5) * ((f & 1) ? 2.0 : 1.0), r0, r1);
            part += r0;
        } else part += c;
    }
    for (int m = threadIdx.x; m < h.nm; m += PNP_THREADS) {
        const int gm = h.moff + m;
        double e[8], c = 0;
        marker_error(A, h, T, gm, e);
#pragma unroll
        for (int k = 0; k < 8; k++) c += e[k] * e[k];
        A.mchi2[gm] = c;
        if (A.mrobust[gm]) {
            double r0, r1;
            ba::huber(c, d8, h.wm, r0, r1);
            part += r0;
        } else part += c;
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) S.red[threadIdx.x >> 5][0] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < PNP_WARPS; w++) s += S.red[w][0];
        S.chi = s;
    }
    __syncthreads();
}

// computeActiveErrors + activeRobustChi2 + linearizeOplus + constructQuadraticForm (base_unary_edge.hpp:50-80) at S.T:
// S.H (upper triangle, row-major a <= c), S.b, S.chi
__device__ void pass_linearize(const PnpArrays& A, const PnpHead& h, PnpShared& S, bool robust) {
    const ba::Pose T = S.T;
    const double d2 = (double)sqrtf(5.99f), d3 = (double)sqrtf(7.815f), d8 = (double)sqrtf(15.507f);
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = 0;
    if (h.nm) {
        for (int k = threadIdx.x; k < PNP_GROUPS * NACC; k += PNP_THREADS) (&S.mH[0][0])[k] = 0;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < h.n; i += PNP_THREADS) {
        const int gi = h.off + i;
        if (!A.active[gi]) continue;
        const uint8_t f = A.flg[gi];
        const bool stereo = f & 1;
        double e[3], p[3];
        double c = edge_error(A, h, T, gi, stereo, e, p);
        A.chi2[gi] = c;
        double rho1 = 1;
        if (robust) {
            double r0;
            ba::huber(c, stereo ? d3 : d2, (f & 2 ? 1.0 : 0.5) * (stereo ? 2.0 : 1.0), r0, rho1);
            acc[27] += r0;
        } else acc[27] += c;
        // typesg2o.h:618-638 / 539-569
        const double x = p[0], y = p[1], invz = 1.0 / p[2], invz_2 = invz * invz;
        double J[18];
        J[0] = x * y * invz_2 * h.fx; J[1] = -(1 + (x * x * invz_2)) * h.fx; J[2] = y * invz * h.fx; J[3] = -invz * h.fx; J[4] = 0; J[5] = x * invz_2 * h.fx;
        J[6] = (1 + y * y * invz_2) * h.fy; J[7] = -x * y * invz_2 * h.fy; J[8] = -x * invz * h.fy; J[9] = 0; J[10] = -invz * h.fy; J[11] = y * invz_2 * h.fy;
        if (stereo) {
            J[12] = J[0] - h.bf * y * invz_2; J[13] = J[1] + h.bf * x * invz_2; J[14] = J[2]; J[15] = J[3]; J[16] = 0; J[17] = J[5] - h.bf * invz_2;
        } else {
#pragma unroll
            for (int k = 12; k < 18; k++) J[k] = 0;
            e[2] = 0;
        }
        const double om = (double)A.isig[gi], wom = rho1 * om;
        const double oe0 = om * e[0], oe1 = om * e[1], oe2 = om * e[2];
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int cc = a; cc < 6; cc++) acc[k++] += J[a] * wom * J[cc] + J[6 + a] * wom * J[6 + cc] + J[12 + a] * wom * J[12 + cc];
        }
#pragma unroll
        for (int a = 0; a < 6; a++) acc[21 + a] -= rho1 * (J[a] * oe0 + J[6 + a] * oe1 + J[12 + a] * oe2);
    }
    // marker edges: one 16-lane group per marker
    {
        const int g = threadIdx.x / PNP_GROUP, l = threadIdx.x % PNP_GROUP;
        const double delta = (double)1e-4f, scalar = 1 / (2 * delta);
        const unsigned gmask = 0xFFFFu << (16 * (g & 1));  // the lanes of this group inside its warp
        for (int m = g; m < h.nm; m += PNP_GROUPS) {  // uniform trip count inside a group; the two groups of a warp may differ
            const int gm = h.moff + m;
            if (l < 13) {
                ba::Pose Tp = T;
                if (l > 0) {
                    double u[6] = {0, 0, 0, 0, 0, 0};
                    const int d = (l - 1) >> 1;
                    const double v = ((l - 1) & 1) ? -delta : delta;
#pragma unroll
                    for (int q = 0; q < 6; q++) u[q] = (q == d) ? v : 0.0;
                    ba::se3_oplus(Tp, u);
                }
                double e[8];
                marker_error(A, h, Tp, gm, e);
#pragma unroll
                for (int k = 0; k < 8; k++) S.mE[g][l][k] = e[k];
            }
            __syncwarp(gmask);
            for (int q = l; q < 48; q += PNP_GROUP) {  // J[k][d], k = residual row, d = pose dimension
                const int k = q / 6, d = q % 6;
                S.mJ[g][q] = scalar * (S.mE[g][1 + 2 * d][k] - S.mE[g][2 + 2 * d][k]);
            }
            __syncwarp(gmask);
            double c = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) c += S.mE[g][0][k] * S.mE[g][0][k];
            double rho1 = 1, r0 = c;
            if (A.mrobust[gm]) ba::huber(c, d8, h.wm, r0, rho1);
            if (l == 0) {
                A.mchi2[gm] = c;
                S.mH[g][27] += r0;
            }
            for (int q = l; q < 27; q += PNP_GROUP) {
                double s = 0;
                if (q < 21) {
                    int a = 0, r = q;
                    while (r >= 6 - a) { r -= 6 - a; a++; }
                    const int cc = a + r;
#pragma unroll
                    for (int k = 0; k < 8; k++) s += S.mJ[g][6 * k + a] * rho1 * S.mJ[g][6 * k + cc];
                    S.mH[g][q] += s;
                } else {
                    const int a = q - 21;
#pragma unroll
                    for (int k = 0; k < 8; k++) s += S.mJ[g][6 * k + a] * S.mE[g][0][k];
                    S.mH[g][q] -= rho1 * s;
                }
            }
            __syncwarp(gmask);
        }
    }
    {   // the 28 warp sums as ONE transposing butterfly: at offset o a lane keeps the half of its values whose index has that bit equal to
        // its own and hands the other half to its partner, so after 5 rounds lane L holds the warp total of accumulator L.  31 exchanges
        // instead of 28 x 5, and the same additions in the same tree as warp_sum (xor 16, 8, 4, 2, 1): bit-identical sums
        double a[32];
#pragma unroll
        for (int k = 0; k < 32; k++) a[k] = k < NACC ? acc[k] : 0.0;
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const bool up = lane & o;
#pragma unroll
            for (int j = 0; j < o; j++) {
                const double send = up ? a[j] : a[j + o], keep = up ? a[j + o] : a[j];
                a[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        if (lane < NACC) S.red[threadIdx.x >> 5][lane] = a[0];
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        const int k = threadIdx.x;
        double s = 0;
#pragma unroll
        for (int w = 0; w < PNP_WARPS; w++) s += S.red[w][k];
        if (h.nm)
            for (int g = 0; g < PNP_GROUPS; g++) s += S.mH[g][k];
        if (k < 21) S.H[k] = s;
        else if (k < 27) S.b[k - 21] = s;
        else S.chi = s;
    }
    __syncthreads();
}

// dense Cholesky of the damped 6x6 system (the reference: SimplicialLDLT on the single 6x6 block); false if not positive definite
__device__ bool solve6(const double* Hu, const double* b, double lambda, double* x) {
    double A[36];
    int k = 0;
    for (int a = 0; a < 6; a++)
        for (int c = a; c < 6; c++) { A[6 * c + a] = Hu[k]; A[6 * a + c] = Hu[k]; k++; }
    for (int j = 0; j < 6; j++) A[7 * j] += lambda;
    for (int j = 0; j < 6; j++) {
        double d = A[7 * j];
        for (int q = 0; q < j; q++) d -= A[6 * j + q] * A[6 * j + q];
        if (!(d > 0)) return false;
        d = sqrt(d);
        A[7 * j] = d;
        for (int i = j + 1; i < 6; i++) {
            double s = A[6 * i + j];
            for (int q = 0; q < j; q++) s -= A[6 * i + q] * A[6 * j + q];
            A[6 * i + j] = s / d;
        }
    }
    for (int i = 0; i < 6; i++) {
        double s = b[i];
        for (int q = 0; q < i; q++) s -= A[6 * i + q] * x[q];
        x[i] = s / A[7 * i];
    }
    for (int i = 5; i >= 0; i--) {
        double s = x[i];
        for (int q = i + 1; q < 6; q++) s -= A[6 * q + i] * x[q];
        x[i] = s / A[7 * i];
    }
    return true;
}

__global__ void __launch_bounds__(PNP_THREADS) pose_only_kernel(const PnpHead* heads, PnpArrays A, PnpOut* outs) {
    __shared__ PnpShared S;
    __shared__ PnpHead h;
    if (threadIdx.x == 0) h = heads[blockIdx.x];
    __syncthreads();
    PnpOut* out = outs + blockIdx.x;
    const float Chi2D = 5.99f, Chi3D = 7.815f, Chi8D = 15.507f;
    for (int i = threadIdx.x; i < h.n; i += PNP_THREADS) A.active[h.off + i] = 1;
    for (int m = threadIdx.x; m < h.nm; m += PNP_THREADS) A.mrobust[h.moff + m] = 1;
    if (threadIdx.x == 0) {
        S.T0 = ba::pose_from_m44f(h.pose44);
        S.T = S.T0;
        for (int k = 0; k < 4; k++) out->iters[k] = 0;
    }
    __syncthreads();
    if (h.n == 0 && h.nm == 0) {  // pnpsolver.cpp:144: nothing to do, pose untouched
        if (threadIdx.x < 16) out->pose44[threadIdx.x] = h.pose44[threadIdx.x];
        if (threadIdx.x < 7) out->pose7[threadIdx.x] = threadIdx.x < 4 ? S.T.q[threadIdx.x] : S.T.t[threadIdx.x - 4];
        if (threadIdx.x == 0) out->n_good = 0;
        return;
    }
    // LM controller state: identical in every thread (derived from shared values only)
    for (int round = 0; round < 4; round++) {
        const bool robust = round <= 2;  // kernels are removed after the classification of round index 2 (:371)
        if (threadIdx.x == 0) S.T = S.T0;  // :358
        __syncthreads();
        double lambda = 0, ni = 2;
        float prevChi2 = FLT_MAX, curChi2 = FLT_MAX, Chi2Diff = FLT_MAX;
        bool ok = true;
        int its = 0;
        for (int it = 0; it < 10 && ok && Chi2Diff > 0.0f; it++) {
            { float t = prevChi2; prevChi2 = curChi2; curChi2 = t; }
            pass_linearize(A, h, S, robust);
            double currentChi = S.chi, tempChi, lastSum = S.chi;
            if (it == 0) {  // computeLambdaInit, optimization_algorithm_levenberg.cpp:152-166
                double md = 0;
                const int dg[6] = {0, 6, 11, 15, 18, 20};
#pragma unroll
                for (int j = 0; j < 6; j++) md = fmax(fabs(S.H[dg[j]]), md);
                lambda = 1e-5 * md;
                ni = 2;
            }
            double rho = 0;
            int qmax = 0;
            do {
                if (threadIdx.x == 0) {
                    S.Tbak = S.T;
                    double x[6] = {0, 0, 0, 0, 0, 0};
                    bool ok2 = solve6(S.H, S.b, lambda, x);
                    if (ok2) ba::se3_oplus(S.T, x);
                    S.flag = ok2;
#pragma unroll
                    for (int k = 0; k < 6; k++) S.x[k] = x[k];
                }
                __syncthreads();
                pass_errors(A, h, S, robust);
                const bool ok2 = S.flag;
                lastSum = S.chi;
                tempChi = ok2 ? S.chi : DBL_MAX;
                rho = currentChi - tempChi;
                double scale = 0;  // computeScale :168-175
#pragma unroll
                for (int k = 0; k < 6; k++) scale += S.x[k] * (lambda * S.x[k] + S.b[k]);
                scale += 1e-3;
                rho /= scale;
                bool finite_l = true;
                if (rho > 0 && isfinite(tempChi)) {
                    double alpha = 1. - pow((2 * rho - 1), 3.0);
                    alpha = fmin(alpha, 2. / 3.);
                    lambda *= fmax(1. / 3., alpha);
                    ni = 2;
                    currentChi = tempChi;
                } else {
                    lambda *= ni;
                    ni *= 2;
                    __syncthreads();  // every thread has read S.T-dependent results
                    if (threadIdx.x == 0) S.T = S.Tbak;
                    __syncthreads();
                    if (!isfinite(lambda)) finite_l = false;
                }
                if (!finite_l) break;
                qmax++;
            } while (rho < 0 && qmax < 10);
            if (qmax == 10 || rho == 0 || !isfinite(lambda)) ok = false;
            curChi2 = (float)lastSum;  // activeRobustChi2() over the errors of the last evaluation
            Chi2Diff = prevChi2 - curChi2;
            its++;
        }
        __syncthreads();
        // classification, pnpsolver.cpp:364-381
        if (threadIdx.x == 0) { S.good = 0; out->iters[round] = its; }
        __syncthreads();
        {
            const ba::Pose T = S.T;
            int good = 0;
            for (int i = threadIdx.x; i < h.n; i += PNP_THREADS) {
                const int gi = h.off + i;
                const uint8_t f = A.flg[gi];
                if (!A.active[gi]) {
                    double e[3], p[3];
                    A.chi2[gi] = edge_error(A, h, T, gi, f & 1, e, p);
                }
                const bool bad = A.chi2[gi] > (double)((f & 1) ? Chi3D : Chi2D);
                A.active[gi] = !bad;
                good += !bad;
            }
            for (int m = threadIdx.x; m < h.nm; m += PNP_THREADS) {
                const int gm = h.moff + m;
                double e[8], c = 0;
                marker_error(A, h, T, gm, e);
#pragma unroll
                for (int k = 0; k < 8; k++) c += e[k] * e[k];
                A.mchi2[gm] = c;
                if (c > (double)Chi8D || round >= 2) A.mrobust[gm] = 0;
            }
            if (good) atomicAdd(&S.good, good);
        }
        __syncthreads();
        if (S.good < 10 && h.nm == 0) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double R[9];
        ba::quat_to_R(S.T.q, R);
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) out->pose44[4 * r + c] = (float)R[3 * r + c];
            out->pose44[4 * r + 3] = (float)S.T.t[r];
        }
        out->pose44[12] = out->pose44[13] = out->pose44[14] = 0;
        out->pose44[15] = 1;
        for (int k = 0; k < 4; k++) out->pose7[k] = S.T.q[k];
        for (int k = 0; k < 3; k++) out->pose7[4 + k] = S.T.t[k];
        out->n_good = S.good;
    }
    for (int i = threadIdx.x; i < h.n; i += PNP_THREADS) A.bad[h.off + i] = !A.active[h.off + i];
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// device-resident launch shared with track.cu: n problems described by heads / arrays already in device memory
int uco_pnp_launch_dev(uco_b200_ctx* ctx, int n, const PnpHead* heads_dev, const PnpArrays& A, PnpOut* outs_dev) {
    pose_only_kernel<<<n, PNP_THREADS, 0, ctx->stream>>>(heads_dev, A, outs_dev);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

extern "C" {

int uco_b200_pose_only_batch(uco_b200_ctx* ctx, int n, const uco_pnp_problem* pbs, uco_pnp_result* res) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (n <= 0 || !pbs || !res) return uco_fail(ctx, UCO_E_INVALID, "pose_only: bad arguments");
    size_t N = 0, M = 0;
    for (int p = 0; p < n; p++) {
        const uco_pnp_problem& pb = pbs[p];
        if (pb.n_matches < 0 || pb.n_markers < 0 || !pb.pose44) return uco_fail(ctx, UCO_E_INVALID, "pose_only: problem %d malformed", p);
        if (pb.n_matches && (!pb.points3 || !pb.obs_uv || !pb.obs_inv_sigma2))
            return uco_fail(ctx, UCO_E_INVALID, "pose_only: problem %d has null match arrays", p);
        if (pb.n_markers && (!pb.marker_pose44 || !pb.marker_size || !pb.marker_corners))
            return uco_fail(ctx, UCO_E_INVALID, "pose_only: problem %d has null marker arrays", p);
        N += pb.n_matches;
        M += pb.n_markers;
    }
    // packed input: heads | pts | uv | ur | isig | flags | mpose | msize | mobs
    const size_t o_head = 0, o_pts = al(o_head + sizeof(PnpHead) * n), o_uv = al(o_pts + 12 * N), o_ur = al(o_uv + 8 * N),
                 o_is = al(o_ur + 4 * N), o_fl = al(o_is + 4 * N), o_mp = al(o_fl + N), o_ms = al(o_mp + 64 * M),
                 o_mo = al(o_ms + 4 * M), in_bytes = al(o_mo + 32 * M);
    const size_t o_out = 0, o_bad = al(sizeof(PnpOut) * n), out_bytes = al(o_bad + N);
    const size_t s_chi = 0, s_act = al(8 * N), s_mchi = al(s_act + N), s_mrob = al(s_mchi + 8 * M), scr_bytes = al(s_mrob + M);
    uint8_t* hin = (uint8_t*)uco_pinned(ctx, WS_PNP_IN, in_bytes);
    uint8_t* hout = (uint8_t*)uco_pinned(ctx, WS_PNP_OUT, out_bytes);
    uint8_t* din = (uint8_t*)uco_ws(ctx, WS_PNP_IN, in_bytes);
    uint8_t* dout = (uint8_t*)uco_ws(ctx, WS_PNP_OUT, out_bytes);
    uint8_t* dscr = (uint8_t*)uco_ws(ctx, WS_PNP_SCRATCH, scr_bytes);
    if (!hin || !hout || !din || !dout || !dscr) return UCO_E_NOMEM;
    PnpHead* heads = (PnpHead*)(hin + o_head);
    size_t off = 0, moff = 0;
    for (int p = 0; p < n; p++) {
        const uco_pnp_problem& pb = pbs[p];
        PnpHead& h = heads[p];
        h.n = pb.n_matches; h.nm = pb.n_markers; h.off = (int)off; h.moff = (int)moff;
        memcpy(h.pose44, pb.pose44, 64);
        h.fx = pb.fx; h.fy = pb.fy; h.cx = pb.cx; h.cy = pb.cy; h.bf = pb.bf;
        const int nn = pb.n_matches;
        if (nn) {
            memcpy(hin + o_pts + 12 * off, pb.points3, 12 * (size_t)nn);
            memcpy(hin + o_uv + 8 * off, pb.obs_uv, 8 * (size_t)nn);
            if (pb.obs_ur) memcpy(hin + o_ur + 4 * off, pb.obs_ur, 4 * (size_t)nn);
            else memset(hin + o_ur + 4 * off, 0, 4 * (size_t)nn);
            memcpy(hin + o_is + 4 * off, pb.obs_inv_sigma2, 4 * (size_t)nn);
        }
        double kpw = 0;  // KpWeightSum, pnpsolver.cpp:201-258
        uint8_t* fl = hin + o_fl + off;
        for (int i = 0; i < nn; i++) {
            const bool st = pb.obs_stereo && pb.obs_stereo[i], stable = !pb.stable || pb.stable[i];
            fl[i] = (uint8_t)((st ? 1 : 0) | (stable ? 2 : 0));
            float ew = 1;
            if (!stable) ew = 0.5;
            if (st) ew *= 2;
            kpw += ew;
        }
        {
            float w_markers = 0.3;
            int total = pb.n_matches + pb.n_markers;
            h.wm = ((w_markers * total) / (1. - w_markers)) / float(kpw);  // :298-300
        }
        if (pb.n_markers) {
            memcpy(hin + o_mp + 64 * moff, pb.marker_pose44, 64 * (size_t)pb.n_markers);
            memcpy(hin + o_ms + 4 * moff, pb.marker_size, 4 * (size_t)pb.n_markers);
            memcpy(hin + o_mo + 32 * moff, pb.marker_corners, 32 * (size_t)pb.n_markers);
        }
        off += nn;
        moff += pb.n_markers;
    }
    UCO_CUDA(ctx, cudaMemcpyAsync(din, hin, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    PnpArrays A;
    A.pts = (const float*)(din + o_pts); A.uv = (const float*)(din + o_uv); A.ur = (const float*)(din + o_ur);
    A.isig = (const float*)(din + o_is); A.flg = din + o_fl; A.mpose = (const float*)(din + o_mp);
    A.msize = (const float*)(din + o_ms); A.mobs = (const float*)(din + o_mo);
    A.chi2 = (double*)(dscr + s_chi); A.active = dscr + s_act; A.mchi2 = (double*)(dscr + s_mchi); A.mrobust = dscr + s_mrob;
    A.bad = dout + o_bad;
    pose_only_kernel<<<n, PNP_THREADS, 0, ctx->stream>>>((const PnpHead*)(din + o_head), A, (PnpOut*)(dout + o_out));
    UCO_LAUNCH_CHECK(ctx);
    UCO_CUDA(ctx, cudaMemcpyAsync(hout, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const PnpOut* outs = (const PnpOut*)(hout + o_out);
    off = 0;
    for (int p = 0; p < n; p++) {
        memcpy(res[p].pose44, outs[p].pose44, 64);
        memcpy(res[p].pose7, outs[p].pose7, 56);
        res[p].n_good = outs[p].n_good;
        memcpy(res[p].iters, outs[p].iters, 16);
        if (res[p].bad && pbs[p].n_matches) memcpy(res[p].bad, hout + o_bad + off, pbs[p].n_matches);
        off += pbs[p].n_matches;
    }
    return UCO_OK;
}

int uco_b200_pose_only(uco_b200_ctx* ctx, const uco_pnp_problem* pb, uco_pnp_result* res) {
    return uco_b200_pose_only_batch(ctx, 1, pb, res);
}

}  // extern "C"
