// ba_cluster.cu — K10-K13, cluster-resident form: ONE thread-block cluster per bundle-adjustment window runs the whole
// two-stage Levenberg-Marquardt solve (every iteration, every trial, the outlier pass between the stages and the result
// extraction) inside a single kernel launch; a batch of independent windows is one launch with one cluster per window.
//
// Why: a local-BA window (10-30 keyframes, a few thousand points, ~15k observations) is far too small to fill a B200 and
// its LM loop is a chain of ~100 tiny dependent phases.  As separate kernels each phase pays a launch + drain (ba.cu, the
// streamed form, spends ~300 us per LM trial that way); inside one cluster the phases are separated by barrier.cluster
// (~0.2 us), the LM state machine is replicated in every CTA's shared memory (each CTA derives the same decision from the
// same ordered partial sums, so no broadcast is needed), and 18 windows run side by side on the 148 SMs.
//
// Phases of one LM trial (S = cluster barrier):  prep (D^-1, Y = W D^-1)  S  Schur gather into per-unit partials  S
//   CTA 0: assemble the reduced system in shared memory, blocked (6-wide) L D L^T, back-substitution  S
//   update (landmark back-substitution, backup, oplus)  S  residuals + ordered chi2 partials  S  decide (+ restore).
// All sums have a fixed order (strided partials, tree within a CTA, rank order across CTAs): results are bitwise
// reproducible for a given cluster size.  Arithmetic is the same as ba.cu's (ba_math.cuh); the reference code each phase
// restates is cited there and in include/ucoslam_b200.h.
#include "common.cuh"
#include "ba_math.cuh"
#include "ba_plan.h"
#include <cooperative_groups.h>
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {
using namespace ba;

constexpr int BS = 512;           // threads per CTA (f64 Jacobian code wants ~128 registers)
constexpr int TEAMS = BS / 36;    // Schur-gather teams of 36 threads (one 6x6 block element each)
constexpr int NW = BS / 32;

struct LmLocal {  // replicated per CTA
    double lambda, ni, currentChi, rho;
    float prevChi2, curChi2, chi2Diff;
    int it, qmax, ok, cont_trial, cont_iter, reject, ntrace, stopped;
};

__device__ __forceinline__ double block_sum(double v, double* red) {  // ordered: xor-tree in the warp, warps in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) s += red[w];
    return s;
}
__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) s = fmax(s, red[w]);
    return s;
}

__device__ __forceinline__ void obs_weights(const CbDev& B, int i, int robust, double& wo, double* orr) {
    double w = (double)B.info[i], r1 = 1, r0;
    if (robust) huber(B.chi2[i], B.stereo[i] ? B.d3 : B.d2, 1.0, r0, r1);
    wo = r1 * w;
#pragma unroll
    for (int d = 0; d < 3; d++) orr[d] = -(w * B.err[3 * i + d]) * r1;
}

// residuals of every active observation; returns this thread's ordered partial of the (robustified) chi2
__device__ __forceinline__ double phase_errors(const CbDev& B, int ct, int cn, int robust) {
    double s = 0;
    for (int i = ct; i < B.M; i += cn) {
        if (!B.active[i]) continue;
        Pose T = load_pose(B.pose + 7 * B.obs_pose[i]);
        const double* X = B.pt + 3 * B.obs_lm[i];
        double x[3] = {X[0], X[1], X[2]}, p[3], e[3];
        se3_map(T, x, p);
        bool st = B.stereo[i];
        double z[3] = {(double)B.z[3 * i], (double)B.z[3 * i + 1], (double)B.z[3 * i + 2]};
        residual(p, z, st, B.cam, e);
        double info = (double)B.info[i];
        double c2 = st ? (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * info : (e[0] * e[0] + e[1] * e[1]) * info;
        B.err[3 * i] = e[0]; B.err[3 * i + 1] = e[1]; B.err[3 * i + 2] = e[2];
        B.chi2[i] = c2;
        double r0 = c2, r1;
        if (robust) huber(c2, st ? B.d3 : B.d2, 1.0, r0, r1);
        s += r0;
    }
    return s;
}

// Work is cut along landmarks: CTA `rank` owns the landmark range [cta_lm[rank], cta_lm[rank+1]) and, observations being
// sorted by landmark, a contiguous range of observations; it walks that range in chunks of whole landmarks with <= BS
// observations (chunk_lm[]), so everything a landmark needs is inside one CTA and one __syncthreads away.  Per-observation
// 6x3 blocks live in HBM as 18 contiguous doubles (what the Schur gather wants); a thread never writes its own 144-byte
// row directly (32 lanes x 8 B scattered over 32 sectors) but stages it in shared memory and the warp copies it out
// contiguously.
constexpr int WPAD = 19;  // staged row pitch (doubles): odd-ish pitch keeps the per-lane row writes off the same banks

// linearize: per observation the W block and its Hll / bl terms; per landmark the ordered sums.  Returns max |diag(Hll)|.
__device__ __forceinline__ double phase_linearize_lm(const CbDev& B, int rank, int robust, double* stage) {
    const int tid = threadIdx.x;
    double* sW = stage;                 // [BS][WPAD]
    double* sC = stage + BS * WPAD;     // [BS][9]
    double md = 0;
    for (int ch = B.cta_chunk_ptr[rank]; ch < B.cta_chunk_ptr[rank + 1]; ch++) {
        const int l0 = B.chunk_lm[ch], l1 = B.chunk_lm[ch + 1], o0 = B.lm_ptr[l0], o1 = B.lm_ptr[l1], nobs = o1 - o0;
        if (tid < nobs) {
            const int i = o0 + tid;
            double* W = sW + tid * WPAD;
            double* C = sC + tid * 9;
            if (!B.active[i]) {
#pragma unroll
                for (int k = 0; k < 18; k++) W[k] = 0;
#pragma unroll
                for (int k = 0; k < 9; k++) C[k] = 0;
            } else {
                const int pi = B.obs_pose[i];
                Pose T = load_pose(B.pose + 7 * pi);
                const double* X = B.pt + 3 * B.obs_lm[i];
                double x[3] = {X[0], X[1], X[2]}, p[3], R[9], JX[9], JT[18], wo, orr[3];
                se3_map(T, x, p);
                quat_to_R(T.q, R);
                const bool st = B.stereo[i];
                jac_point(p, R, st, B.cam, JX);   // rows beyond the residual dimension are zero: adding their terms changes nothing
                obs_weights(B, i, robust, wo, orr);
                {
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < 3; a++)
#pragma unroll
                        for (int c = a; c < 3; c++, k++) {
                            double h = 0;
#pragma unroll
                            for (int d = 0; d < 3; d++) h += JX[3 * d + a] * wo * JX[3 * d + c];
                            C[k] = h;
                        }
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        double sacc = 0;
#pragma unroll
                        for (int d = 0; d < 3; d++) sacc += JX[3 * d + a] * orr[d];
                        C[6 + a] = sacc;
                    }
                }
                if (B.free_idx[pi] >= 0) {
                    jac_pose(p, st, B.cam, JT);
#pragma unroll
                    for (int a = 0; a < 6; a++)
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            double h = 0;
#pragma unroll
                            for (int d = 0; d < 3; d++) h += JT[6 * d + a] * wo * JX[3 * d + c];
                            W[3 * a + c] = h;
                        }
                } else {
#pragma unroll
                    for (int k = 0; k < 18; k++) W[k] = 0;
                }
            }
        }
        __syncthreads();
        {
            double* Wg = B.W + 18 * (size_t)o0;
            for (int e = tid; e < 18 * nobs; e += BS) Wg[e] = sW[(e / 18) * WPAD + e % 18];
        }
        if (tid < l1 - l0) {
            const int l = l0 + tid;
            double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = B.lm_ptr[l] - o0; i < B.lm_ptr[l + 1] - o0; i++) {
#pragma unroll
                for (int k = 0; k < 9; k++) acc[k] += sC[i * 9 + k];
            }
#pragma unroll
            for (int k = 0; k < 6; k++) B.Hll[6 * (size_t)l + k] = acc[k];
#pragma unroll
            for (int k = 0; k < 3; k++) B.bl[3 * (size_t)l + k] = acc[6 + k];
            md = fmax(md, fmax(fabs(acc[0]), fmax(fabs(acc[3]), fabs(acc[5]))));
        }
        __syncthreads();
    }
    return md;
}

// pose blocks Hpp / bp: one CTA per free pose, Jacobian recomputed from the pose's observation list, ordered reduction
__device__ __forceinline__ double phase_linearize_pose(const CbDev& B, int rank, int CL, int robust, double* red27) {
    double md = 0;
    const int tid = threadIdx.x;
    for (int f = rank; f < B.Pf; f += CL) {
        const int pi = B.free_list[f];
        Pose T = load_pose(B.pose + 7 * pi);
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] = 0;
        for (int j = B.pose_ptr[f] + tid; j < B.pose_ptr[f + 1]; j += BS) {
            const int i = B.pose_obs[j];
            if (!B.active[i]) continue;
            const double* X = B.pt + 3 * B.obs_lm[i];
            double x[3] = {X[0], X[1], X[2]}, p[3], JT[18], wo, orr[3];
            se3_map(T, x, p);
            const bool st = B.stereo[i];
            jac_pose(p, st, B.cam, JT);
            obs_weights(B, i, robust, wo, orr);
            int k = 0;
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int c = a; c < 6; c++, k++) {
                    double h = 0;
#pragma unroll
                    for (int d = 0; d < 3; d++) h += JT[6 * d + a] * wo * JT[6 * d + c];
                    acc[k] += h;
                }
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double sacc = 0;
#pragma unroll
                for (int d = 0; d < 3; d++) sacc += JT[6 * d + a] * orr[d];
                acc[21 + a] += sacc;
            }
        }
        __syncthreads();  // red27 reuse
#pragma unroll
        for (int k = 0; k < 27; k++) {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0) red27[(tid >> 5) * 27 + k] = v;
        }
        __syncthreads();
        if (tid < 27) {
            double v = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) v += red27[w * 27 + tid];
            red27[NW * 27 + tid] = v;
        }
        __syncthreads();
        if (tid < 36) {
            int a = tid / 6, c = tid % 6;
            int lo = a < c ? a : c, hi = a < c ? c : a;
            int k = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
            double v = red27[NW * 27 + k];
            B.Hpp[36 * (size_t)f + tid] = v;
            if (a == c) md = fmax(md, fabs(v));
        }
        if (tid < 6) B.bp[6 * (size_t)f + tid] = red27[NW * 27 + 21 + tid];
    }
    return md;
}

// per trial and chunk: D^-1 of the chunk's landmarks (kept in shared memory and in HBM for the back-substitution), then
// Y = W D^-1 one 3-wide row per thread (consecutive threads read and write consecutive 24 bytes), two rows in flight
__device__ __forceinline__ void phase_prep(const CbDev& B, int rank, double lambda, double* stage) {
    const int tid = threadIdx.x;
    double* sD = stage;  // [BS][6]
    for (int ch = B.cta_chunk_ptr[rank]; ch < B.cta_chunk_ptr[rank + 1]; ch++) {
        const int l0 = B.chunk_lm[ch], l1 = B.chunk_lm[ch + 1], o0 = B.lm_ptr[l0], o1 = B.lm_ptr[l1];
        if (tid < l1 - l0) {
            const int l = l0 + tid;
            double D[6], I[6];
#pragma unroll
            for (int k = 0; k < 6; k++) D[k] = B.Hll[6 * (size_t)l + k];
            D[0] += lambda; D[3] += lambda; D[5] += lambda;
            inv3_sym(D, I);
#pragma unroll
            for (int k = 0; k < 6; k++) {
                B.Dinv[6 * (size_t)l + k] = I[k];
                sD[6 * tid + k] = I[k];
            }
        }
        __syncthreads();
        const int r0 = 6 * o0, r1 = 6 * o1;
        for (int row = r0 + tid; row < r1; row += 2 * BS) {
            const int rowb = row + BS;
            const bool hb = rowb < r1;
            const int la = B.obs_lm[row / 6] - l0, lb = hb ? B.obs_lm[rowb / 6] - l0 : 0;
            const double* Wa = B.W + 3 * (size_t)row;
            const double* Wb = B.W + 3 * (size_t)(hb ? rowb : row);
            const double a0 = Wa[0], a1 = Wa[1], a2 = Wa[2], b0 = Wb[0], b1 = Wb[1], b2 = Wb[2];
            {
                const double* I = sD + 6 * la;
                double* Y = B.Y + 3 * (size_t)row;
                Y[0] = a0 * I[0] + a1 * I[1] + a2 * I[2];
                Y[1] = a0 * I[1] + a1 * I[3] + a2 * I[4];
                Y[2] = a0 * I[2] + a1 * I[4] + a2 * I[5];
            }
            if (hb) {
                const double* I = sD + 6 * lb;
                double* Y = B.Y + 3 * (size_t)rowb;
                Y[0] = b0 * I[0] + b1 * I[1] + b2 * I[2];
                Y[1] = b0 * I[1] + b1 * I[3] + b2 * I[4];
                Y[2] = b0 * I[2] + b1 * I[4] + b2 * I[5];
            }
        }
        __syncthreads();
    }
}

// Schur gather on the FP64 tensor pipe: one warp per unit (<= BA_UNIT contributions of one 6x6 block), one
// mma.m8n8k4.f64 per contribution: A = Y_a (6x3 in an 8x4 tile), B = W_b^T (3x6 in 4x8), accumulated in the warp's 8x8 C.
// Diagonal units (a == b; con holds (a, landmark)) put bl of the landmark in column 6 of B, so C[:,6] accumulates
// Y_a bl = W_a D^-1 bl, the Schur right-hand-side term.  Fragment layout (PTX ISA, mma.m8n8k4): A[lane/4][lane%4],
// B[lane%4][lane/4], C[lane/4][2(lane%4) + {0,1}]; element (row, k) of a 6x3 block sits at 3 row + k.
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// Instruction diet (the phase is issue-bound, ~16 warps per SM): units are padded on the host to a multiple of four
// contributions with the all-zero dummy observation M / landmark N, and every lane's operand address is
// base + stride * index with lane-constant base / stride (lanes outside the 6x3 tiles read a zero double with stride 0),
// so a contribution costs two shuffles, two address computations, two loads and one DMMA, with no branch.
__device__ __forceinline__ void phase_gather(const CbDev& B, int rank, int CL) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane >> 2, k4 = lane & 3, off = 3 * q + k4;
    const bool valid = k4 < 3 && q < 6, isbl = k4 < 3 && q == 6;
    const int stride = CL * NW;
    int u = rank * NW + warp;
    if (u >= B.nunits) return;
    const double* zero = B.W + 18 * (size_t)B.M;  // the dummy observation's (all-zero) block
    const double* baseA = valid ? B.Y + off : zero;
    const int strideA = valid ? 18 : 0;
    int4 un = B.unit[u];
    int2 i0 = make_int2(B.M, B.M), i1 = i0;
    if (lane < un.z - un.y) i0 = B.con[un.y + lane];
    if (lane + 32 < un.z - un.y) i1 = B.con[un.y + 32 + lane];
    while (true) {
        const int un_next = u + stride;
        int4 nn = make_int4(0, 0, 0, 0);
        int2 n0 = make_int2(B.M, B.M), n1 = n0;
        if (un_next < B.nunits) {
            nn = B.unit[un_next];
            if (lane < nn.z - nn.y) n0 = B.con[nn.y + lane];
            if (lane + 32 < nn.z - nn.y) n1 = B.con[nn.y + 32 + lane];
        }
        const bool diag = un.w != 0;
        const int cnt = un.z - un.y;  // multiple of 4
        // B operand: W_b (off-diagonal), W_a (diagonal, 6x3 tile lanes), bl of the landmark (diagonal, column 6 lanes)
        const double* baseB = valid ? B.W + off : (diag && isbl ? B.bl + k4 : zero);
        const int strideB = valid ? 18 : (diag && isbl ? 3 : 0);
        const bool b_from_a = diag && valid;
        double acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0}, av[4], bv[4], an[4], bn[4];  // four independent accumulator tiles: the
                                                                                           // DMMA dependent-issue latency is long
#define BA_FETCH(k, A_, B_)                                                                        \
    _Pragma("unroll") for (int j = 0; j < 4; j++) {                                               \
        const int kk = (k) + j;                                                                    \
        const int2 src = kk < 32 ? i0 : i1;                                                        \
        const int a = __shfl_sync(0xffffffffu, src.x, kk & 31), b = __shfl_sync(0xffffffffu, src.y, kk & 31); \
        A_[j] = baseA[(size_t)(strideA * a)];                                                     \
        B_[j] = baseB[(size_t)(strideB * (b_from_a ? a : b))];                                     \
    }
        BA_FETCH(0, av, bv)
        for (int k = 0; k < cnt; k += 4) {
            if (k + 4 < cnt) { BA_FETCH(k + 4, an, bn) }
#pragma unroll
            for (int j = 0; j < 4; j++) mma884(acc0[j], acc1[j], av[j], bv[j]);
#pragma unroll
            for (int j = 0; j < 4; j++) { av[j] = an[j]; bv[j] = bn[j]; }
        }
#undef BA_FETCH
        const double c0 = (acc0[0] + acc0[1]) + (acc0[2] + acc0[3]), c1 = (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]);
        if (q < 6) {
            const int col = 2 * k4;
            if (col < 6) {
                B.part[36 * (size_t)u + 6 * q + col] = c0;
                B.part[36 * (size_t)u + 6 * q + col + 1] = c1;
            } else if (diag) {
                B.partb[6 * (size_t)u + q] = c0;
            }
        }
        if (un_next >= B.nunits) break;
        u = un_next; un = nn; i0 = n0; i1 = n1;
    }
}

// Fused prep + Schur gather (windows of <= BA_GSLOTS blocks per warp).  The gather above streams every contribution's two 6x3 operands from
// L2 (144 B each, every block re-reading the landmarks it shares with the others) and leaves one partial block per 64 contributions in
// HBM.  Here a CTA walks its landmarks chunk by chunk: the chunk's W rows come in once (coalesced), Y = W D^-1 is formed in shared
// memory and never written out, and every warp runs the chunk's contributions of the blocks it OWNS from shared memory into register
// accumulators that live across the chunks (BA_GSLOTS slots x one m8n8k4 C tile).  One partial per (CTA, block) goes to HBM at the end; CTA 0
// adds them in rank order.  Fixed order throughout: bitwise reproducible for a given cluster size.
constexpr int PG_W = 0, PG_Y = (BS + 1) * 18, PG_D = 2 * (BS + 1) * 18, PG_BL = PG_D + BS * 6, PG_DOUBLES = PG_BL + (BS + 1) * 3 + 1;
__device__ __forceinline__ void phase_prep_gather(const CbDev& B, int rank, double lambda, double* stage) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* sW = stage + PG_W;
    double* sY = stage + PG_Y;
    double* sD = stage + PG_D;
    double* sbl = stage + PG_BL;
    const int q = lane >> 2, k4 = lane & 3, off = 3 * q + k4;
    const bool valid = k4 < 3 && q < 6, isbl = k4 < 3 && q == 6;
    double acc[BA_GSLOTS][2];
#pragma unroll
    for (int sl = 0; sl < BA_GSLOTS; sl++) acc[sl][0] = acc[sl][1] = 0;
    for (int ch = B.cta_chunk_ptr[rank]; ch < B.cta_chunk_ptr[rank + 1]; ch++) {
        const int l0 = B.chunk_lm[ch], l1 = B.chunk_lm[ch + 1], o0 = B.lm_ptr[l0], nobs = B.lm_ptr[l1] - o0, nl = l1 - l0;
        const long long tc0 = clock64();
        // the first segment's header and entries are requested now and arrive while the operands are staged
        int sg = B.gw_ptr[ch * NW + warp];
        const int sg_end = B.gw_ptr[ch * NW + warp + 1];
        int4 seg = make_int4(0, 0, 0, 0);
        uint32_t mine = 0;
        if (sg < sg_end) {
            seg = B.gseg[sg];
            mine = B.gcon[seg.y + lane];                                     // the entry array is padded by 32 at its end
        }
        if (tid < nl) {
            const int l = l0 + tid;
            double D[6], I[6];
#pragma unroll
            for (int k = 0; k < 6; k++) D[k] = B.Hll[6 * (size_t)l + k];
            D[0] += lambda; D[3] += lambda; D[5] += lambda;
            inv3_sym(D, I);
#pragma unroll
            for (int k = 0; k < 6; k++) {
                B.Dinv[6 * (size_t)l + k] = I[k];
                sD[6 * tid + k] = I[k];
            }
#pragma unroll
            for (int k = 0; k < 3; k++) sbl[3 * tid + k] = B.bl[3 * (size_t)l + k];
        }
        if (tid < 3) sbl[3 * nl + tid] = 0;                                  // the dummy landmark / observation of the padding
        if (tid >= 32 && tid < 50) sW[18 * nobs + tid - 32] = sY[18 * nobs + tid - 32] = 0;
        __syncthreads();
        for (int row = tid; row < 6 * nobs; row += BS) {
            const double* Wg = B.W + 3 * ((size_t)6 * o0 + row);
            const double a0 = Wg[0], a1 = Wg[1], a2 = Wg[2];
            const double* I = sD + 6 * (B.obs_lm[o0 + row / 6] - l0);
            sW[3 * row] = a0; sW[3 * row + 1] = a1; sW[3 * row + 2] = a2;
            sY[3 * row] = a0 * I[0] + a1 * I[1] + a2 * I[2];
            sY[3 * row + 1] = a0 * I[1] + a1 * I[3] + a2 * I[4];
            sY[3 * row + 2] = a0 * I[2] + a1 * I[4] + a2 * I[5];
        }
        __syncthreads();
        const long long tc1 = clock64();
        const double* zero = sW + 18 * nobs;
        const double* baseA = valid ? sY + off : zero;
        const int strideA = valid ? 18 : 0;
        while (sg < sg_end) {
            int4 nseg = make_int4(0, 0, 0, 0);
            uint32_t nmine = 0;
            if (sg + 1 < sg_end) {                                           // the next segment's header and first entries: in flight during this one
                nseg = B.gseg[sg + 1];
                nmine = B.gcon[nseg.y + lane];
            }
            const int slot = (seg.x >> 16) & 0xff;
            const bool diag = (seg.x >> 24) != 0;
            const double* baseB = valid ? sW + off : (diag && isbl ? sbl + k4 : zero);
            const int strideB = valid ? 18 : (diag && isbl ? 3 : 0);
            const bool b_from_a = diag && valid;
            // four independent accumulator tiles cover the DMMA dependent-issue latency; more do not help: at one m8n8k4 per
            // contribution (256 FMA issued for the 126 the 6x3 . 3x7 product needs) the phase is bound by the FP64 pipe of the SM
            double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
            for (int base = seg.y; base < seg.z; base += 32) {
                const int nb = min(32, seg.z - base);                      // a multiple of four
                if (base != seg.y) mine = B.gcon[base + lane];
                for (int k = 0; k < nb; k += 4) {
                    double av[4], bv[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t e = __shfl_sync(0xffffffffu, mine, k + j);
                        const int ia = (int)(e & 0xffffu), ib = (int)(e >> 16);
                        av[j] = baseA[strideA * ia];
                        bv[j] = baseB[strideB * (b_from_a ? ia : ib)];
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) mma884(a0[j], a1[j], av[j], bv[j]);
                }
            }
            const double c0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), c1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
#pragma unroll
            for (int sl = 0; sl < BA_GSLOTS; sl++)
                if (sl == slot) { acc[sl][0] += c0; acc[sl][1] += c1; }
            seg = nseg; mine = nmine; sg++;
        }
        const long long tc2 = clock64();
        __syncthreads();                                                     // the next chunk overwrites the staged operands
        if (rank == 0 && tid == 0) {                                         // profile slots 4 / 15: staging, this warp's own gather (the barrier wait is the rest of slot 5)
            B.res->phase_cycles[4] += (double)(tc1 - tc0);
            B.res->phase_cycles[15] += (double)(tc2 - tc1);
        }
    }
    // one partial per (CTA, block)
#pragma unroll
    for (int sl = 0; sl < BA_GSLOTS; sl++) {
        const int blk = B.wblk[warp * BA_GSLOTS + sl];
        if (blk < 0 || q >= 6) continue;
        const size_t at = (size_t)rank * B.nblk + blk;
        const int col = 2 * k4;
        if (col < 6) {
            B.part[36 * at + 6 * q + col] = acc[sl][0];
            B.part[36 * at + 6 * q + col + 1] = acc[sl][1];
        } else {
            B.partb[6 * at + q] = acc[sl][0];                                 // column 6 of a diagonal block's tile: Y_a bl (zero for the others)
        }
    }
}

// CTA 0: reduced system in shared memory (lower triangle packed by rows, row n = right-hand side), L D L^T in 6-wide block
// columns (same operation order as the column-by-column elimination), back-substitution by one warp.  Returns 0 on a
// non-positive pivot.
__device__ int phase_solve(const CbDev& B, double lambda, double* A, double* aux, int* sflag, int CL) {
    const int tid = threadIdx.x, n = B.n, lane = tid & 31, warp = tid >> 5;
    long long ts = clock64();
#define SOLVE_TICK(slot)                                          \
    do {                                                          \
        if (tid == 0) {                                           \
            const long long tn = clock64();                       \
            B.res->phase_cycles[slot] += (double)(tn - ts);       \
            ts = tn;                                              \
        }                                                         \
    } while (0)
    const int tot = (n + 1) * (n + 2) / 2;
    for (int k = tid; k < tot; k += BS) A[k] = 0;
    if (tid == 0) *sflag = 0;
    __syncthreads();
    for (int w = tid; w < B.nblk * 36; w += BS) {
        const int blk = w / 36, e = w % 36, r = e / 6, c = e % 6;
        const int2 ij = B.blk_ij[blk];
        const bool diag = ij.x == ij.y;
        if (diag && r > c) continue;  // the solver reads the upper triangle (SimplicialLDLT<Upper>)
        double t = 0;
        if (B.fused)
            for (int r = 0; r < CL; r++) t += B.part[36 * ((size_t)r * B.nblk + blk) + e];
        else
            for (int u = B.blk_unit_ptr[blk]; u < B.blk_unit_ptr[blk + 1]; u++) t += B.part[36 * (size_t)u + e];
        double h = 0;
        if (diag) {
            h = B.Hpp[36 * (size_t)ij.x + e];
            if (r == c) h += lambda;
        }
        h -= t;
        const int row = 6 * ij.y + c, col = 6 * ij.x + r;  // element (col, row) of the upper triangle -> (row, col) of the lower
        A[row * (row + 1) / 2 + col] = h;
    }
    for (int k = tid; k < n; k += BS) {
        const int f = k / 6, r = k % 6;
        const int blk = B.diag_blk[f];
        double t = 0;
        if (B.fused)
            for (int rr = 0; rr < CL; rr++) t += B.partb[6 * ((size_t)rr * B.nblk + blk) + r];
        else
            for (int u = B.blk_unit_ptr[blk]; u < B.blk_unit_ptr[blk + 1]; u++) t += B.partb[6 * (size_t)u + r];
        A[n * (n + 1) / 2 + k] = B.bp[k] - t;
    }
    __syncthreads();
    SOLVE_TICK(12);
    double* dinv = aux;        // 6 reciprocal pivots of the current block column
    double* Ps = aux + 8;      // scaled panel: Ps[(i - c0 - 6) * 6 + j] = A[i][c0 + j] / d_j for the rows below the block
    for (int kb = 0; kb < B.Pf; kb++) {
        const int c0 = 6 * kb;
        if (warp == 0) {  // diagonal 6x6 block, element (i, k) with k <= i on lane
            int li = 0, lk = 0;
            if (lane < 21) {
                int q = lane;
                while (q > li) { q -= li + 1; li++; }
                lk = q;
            }
            double* a = A + (c0 + li) * (c0 + li + 1) / 2 + c0 + lk;
            for (int j = 0; j < 6; j++) {
                const double p = A[(c0 + j) * (c0 + j + 1) / 2 + c0 + j];
                if (!(p > 0) || !isfinite(p)) {
                    if (lane == 0) *sflag = 1;
                    break;
                }
                double upd = 0;
                const bool mine = lane < 21 && lk > j;
                if (mine) upd = A[(c0 + li) * (c0 + li + 1) / 2 + c0 + j] * A[(c0 + lk) * (c0 + lk + 1) / 2 + c0 + j] * (1.0 / p);
                __syncwarp();
                if (mine) *a -= upd;
                if (lane == 0) dinv[j] = 1.0 / p;
                __syncwarp();
            }
        }
        __syncthreads();
        if (*sflag) break;
        // panel rows below the block (including the right-hand-side row n): forward substitution against the block
        for (int i = c0 + 6 + tid; i <= n; i += BS) {
            double* row = A + i * (i + 1) / 2 + c0;
            double v[6];
#pragma unroll
            for (int j = 0; j < 6; j++) v[j] = row[j];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const double lj = v[j] * dinv[j];
#pragma unroll
                for (int k = j + 1; k < 6; k++) v[k] -= lj * A[(c0 + k) * (c0 + k + 1) / 2 + c0 + j];
            }
#pragma unroll
            for (int j = 0; j < 6; j++) {
                row[j] = v[j];
                Ps[(i - c0 - 6) * 6 + j] = v[j] * dinv[j];
            }
        }
        __syncthreads();
        // trailing update, one term per eliminated column in column order
        for (int i = c0 + 6 + warp; i <= n; i += NW) {
            double* row = A + i * (i + 1) / 2;
            const double* pi = Ps + (i - c0 - 6) * 6;
            const double l0 = pi[0], l1 = pi[1], l2 = pi[2], l3 = pi[3], l4 = pi[4], l5 = pi[5];
            const int kmax = i < n ? i : n - 1;
            for (int k = c0 + 6 + lane; k <= kmax; k += 32) {
                const double* ak = A + k * (k + 1) / 2 + c0;
                double v = row[k];
                v -= l0 * ak[0]; v -= l1 * ak[1]; v -= l2 * ak[2]; v -= l3 * ak[3]; v -= l4 * ak[4]; v -= l5 * ak[5];
                row[k] = v;
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (*sflag) {
        for (int k = tid; k < n; k += BS) B.xp[k] = 0;
        return 0;
    }
    SOLVE_TICK(13);
    for (int j = tid; j < n; j += BS) aux[256 + j] = 1.0 / A[j * (j + 1) / 2 + j];  // pivots inverted side by side, off the serial chain
    __syncthreads();
    if (warp == 0) {  // x_j = (w_j - sum_{i>j} A[i][j] x_i) / d_j
        double* s = aux;  // n entries
        const double* rcp = aux + 256;
        const double* wrow = A + n * (n + 1) / 2;
        for (int k = lane; k < n; k += 32) s[k] = wrow[k];
        __syncwarp();
        for (int j = n - 1; j >= 0; j--) {
            const double* row = A + j * (j + 1) / 2;
            const double xj = s[j] * rcp[j];
            __syncwarp();
            if (lane == 0) s[j] = xj;
            for (int k = lane; k < j; k += 32) s[k] -= row[k] * xj;
            __syncwarp();
        }
        for (int k = lane; k < n; k += 32) B.xp[k] = s[k];
    }
    __syncthreads();
    SOLVE_TICK(14);
#undef SOLVE_TICK
    return 1;
}

// landmark back-substitution xl = D^-1 (bl - sum W^T xp) (block_solver.hpp:413-443), backup (push) and update of the points
// of this CTA's landmark range and of the free poses (spread over the cluster); returns the thread's partial of computeScale.
// Per chunk: one thread per 3-wide row of W multiplies it by its xp entry (coalesced), one thread per landmark adds the
// rows up in (observation, row) order.
__device__ __forceinline__ double phase_update(const CbDev& B, int rank, int ct, int cn, double lambda, bool apply, double* stage) {
    const int tid = threadIdx.x;
    double* sx = stage;              // xp, n doubles
    double* sP = stage + 256;        // [BS * 6][3] products
    for (int k = tid; k < B.n; k += BS) sx[k] = B.xp[k];
    __syncthreads();
    double sc = 0;
    for (int ch = B.cta_chunk_ptr[rank]; ch < B.cta_chunk_ptr[rank + 1]; ch++) {
        const int l0 = B.chunk_lm[ch], l1 = B.chunk_lm[ch + 1], o0 = B.lm_ptr[l0], o1 = B.lm_ptr[l1], nrows = 6 * (o1 - o0);
        for (int rr = tid; rr < nrows; rr += 2 * BS) {  // two rows in flight; W of an inactive / fixed-pose observation is zero
            const int rb = rr + BS;
            const bool hb = rb < nrows;
            const int fa = B.obs_free[o0 + rr / 6], fb = hb ? B.obs_free[o0 + rb / 6] : -1;
            const double* Wa = B.W + 3 * ((size_t)6 * o0 + rr);
            const double* Wb = B.W + 3 * ((size_t)6 * o0 + (hb ? rb : rr));
            const double a0 = Wa[0], a1 = Wa[1], a2 = Wa[2], b0 = Wb[0], b1 = Wb[1], b2 = Wb[2];
            const double xa = fa >= 0 ? sx[6 * fa + rr % 6] : 0.0;
            sP[3 * rr] = a0 * xa; sP[3 * rr + 1] = a1 * xa; sP[3 * rr + 2] = a2 * xa;
            if (hb) {
                const double xb = fb >= 0 ? sx[6 * fb + rb % 6] : 0.0;
                sP[3 * rb] = b0 * xb; sP[3 * rb + 1] = b1 * xb; sP[3 * rb + 2] = b2 * xb;
            }
        }
        __syncthreads();
        if (tid < l1 - l0) {
            const int l = l0 + tid;
            if (B.lm_ptr[l] != B.lm_ptr[l + 1]) {  // a point without edges is not an active vertex
                double c[3] = {B.bl[3 * (size_t)l], B.bl[3 * (size_t)l + 1], B.bl[3 * (size_t)l + 2]};
                for (int i = B.lm_ptr[l] - o0; i < B.lm_ptr[l + 1] - o0; i++) {
                    double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
                        s0 += sP[3 * (6 * i + a)]; s1 += sP[3 * (6 * i + a) + 1]; s2 += sP[3 * (6 * i + a) + 2];
                    }
                    c[0] -= s0; c[1] -= s1; c[2] -= s2;
                }
                const double* I = B.Dinv + 6 * (size_t)l;
                const double xl[3] = {I[0] * c[0] + I[1] * c[1] + I[2] * c[2], I[1] * c[0] + I[3] * c[1] + I[4] * c[2],
                                      I[2] * c[0] + I[4] * c[1] + I[5] * c[2]};
                double sacc = 0;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double v = B.pt[3 * (size_t)l + k];
                    B.pt_bak[3 * (size_t)l + k] = v;
                    if (apply) B.pt[3 * (size_t)l + k] = v + xl[k];
                    sacc += xl[k] * (lambda * xl[k] + B.bl[3 * (size_t)l + k]);
                }
                sc += sacc;
            }
        }
        __syncthreads();
    }
    for (int f = ct; f < B.Pf; f += cn) {
        const int pi = B.free_list[f];
        Pose T = load_pose(B.pose + 7 * pi);
        store_pose(B.pose_bak + 7 * pi, T);
        double u[6], sacc = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            u[k] = sx[6 * f + k];
            sacc += u[k] * (lambda * u[k] + B.bp[6 * f + k]);
        }
        sc += sacc;
        if (apply) {
            se3_oplus(T, u);
            store_pose(B.pose + 7 * pi, T);
        }
    }
    return sc;
}
__device__ __forceinline__ void phase_restore(const CbDev& B, int rank, int ct, int cn) {  // pop: same owners as phase_update
    const int L0 = B.cta_lm[rank], L1 = B.cta_lm[rank + 1];
    for (int k = 3 * L0 + threadIdx.x; k < 3 * L1; k += BS) B.pt[k] = B.pt_bak[k];
    for (int f = ct; f < B.Pf; f += cn) {
        const int pi = B.free_list[f];
#pragma unroll
        for (int k = 0; k < 7; k++) B.pose[7 * pi + k] = B.pose_bak[7 * pi + k];
    }
}

__global__ void __launch_bounds__(BS, 1) ba_cluster_kernel(const CbDev* __restrict__ probs) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int prob = blockIdx.x / CL, tid = threadIdx.x, ct = rank * BS + tid, cn = CL * BS;
    extern __shared__ double sm_dyn[];   // staging area of the chunked phases; CTA 0 also holds the packed reduced system here
    __shared__ CbDev B;
    __shared__ LmLocal L;
    __shared__ double red[NW + 1];
    __shared__ double red27[(NW + 1) * 27];
    __shared__ double aux[BA_CLUSTER_MAX_N * 6 + 16];
    __shared__ int sflag;
    for (int k = tid; k < (int)(sizeof(CbDev) / 4); k += BS) ((int*)&B)[k] = ((const int*)(probs + prob))[k];
    if (tid == 0) { L.ntrace = 0; L.stopped = 0; }
    __syncthreads();
    // per-phase cycle counters of CTA 0 (read back through uco_ba_result::profile)
    long long tprev = clock64();
    const bool timing = rank == 0 && tid == 0;
#define BA_TICK(slot)                                        \
    do {                                                     \
        if (timing) {                                        \
            long long tnow = clock64();                      \
            B.res->phase_cycles[slot] += (double)(tnow - tprev); \
            tprev = tnow;                                    \
        }                                                    \
    } while (0)
    // vertices: Frame::pose_f2g -> SE3Quat, cv::Point3f -> Vector3d
    for (int t = ct; t < B.P; t += cn) {
        Pose T = pose_from_m44f(B.p44_in + 16 * t);
        store_pose(B.pose + 7 * t, T);
        store_pose(B.pose_bak + 7 * t, T);
    }
    for (int k = ct; k < 3 * B.N; k += cn) B.pt[k] = B.pt_bak[k] = (double)B.pt_in[k];
    if (ct < 18) B.W[18 * (size_t)B.M + ct] = B.Y[18 * (size_t)B.M + ct] = 0;  // the dummy observation / landmark the padded
    if (ct < 3) B.bl[3 * (size_t)B.N + ct] = 0;                                  // Schur-gather units point at
    for (int i = ct; i < B.M; i += cn) {
        B.active[i] = 1;
        B.chi2[i] = 0;
        B.err[3 * i] = B.err[3 * i + 1] = B.err[3 * i + 2] = 0;
    }
    cluster.sync();
    BA_TICK(0);
    for (int stage = 0; stage < 2; stage++) {
        const int robust = stage == 0, max_iters = stage == 0 ? B.n_iters : 2 * B.n_iters;
        if (stage == 1) {  // globaloptimizer_g2o.cpp:432-449
            for (int i = ct; i < B.M; i += cn) {
                Pose T = load_pose(B.pose + 7 * B.obs_pose[i]);
                const double* X = B.pt + 3 * B.obs_lm[i];
                double x[3] = {X[0], X[1], X[2]}, p[3];
                se3_map(T, x, p);
                if (B.chi2[i] > (double)(B.stereo[i] ? B.chi3d : B.chi2d) || !(p[2] > 0.0)) B.active[i] = 0;
            }
            __syncthreads();  // active[] is re-read by the same strided owner below; other CTAs see it after the barrier
            cluster.sync();
            BA_TICK(11);
        }
        {
            double s = block_sum(phase_errors(B, ct, cn, robust), red);
            if (tid == 0) B.parts[rank] = s;
        }
        if (rank == 0 && tid == 0) B.chol_fail[1] = B.stop ? *(volatile const int*)B.stop : 0;  // one reader: every CTA must see the same value
        cluster.sync();
        if (tid == 0) {
            L.prevChi2 = L.curChi2 = L.chi2Diff = FLT_MAX;
            L.it = 0;
            L.ok = 1;
            const int stop = ((volatile int*)B.chol_fail)[1];
            L.cont_iter = max_iters > 0 && !stop;
            if (stop) L.stopped = 1;
        }
        BA_TICK(1);
        while (true) {
            __syncthreads();
            if (!L.cont_iter) break;
            // ---- linearize at the current estimate
            {
                double md = phase_linearize_lm(B, rank, robust, sm_dyn);
                BA_TICK(2);
                md = block_max(fmax(md, phase_linearize_pose(B, rank, CL, robust, red27)), red);
                if (tid == 0) B.parts[32 + rank] = md;
            }
            cluster.sync();
            if (tid == 0) {
                if (L.it == 0) {  // computeLambdaInit, levenberg.cpp:152-166
                    double s = 0, md = 0;
                    for (int r = 0; r < CL; r++) { s += B.parts[r]; md = fmax(md, B.parts[32 + r]); }
                    L.currentChi = s;
                    L.lambda = 1e-5 * md;
                    L.ni = 2;
                }
                float t = L.prevChi2; L.prevChi2 = L.curChi2; L.curChi2 = t;
                L.qmax = 0;
                L.rho = 0;
                L.cont_trial = 1;
            }
            __syncthreads();
            BA_TICK(3);
            while (L.cont_trial) {
                const double lambda = L.lambda;
                if (B.fused) {
                    phase_prep_gather(B, rank, lambda, sm_dyn);
                    cluster.sync();
                    BA_TICK(5);
                } else {
                    phase_prep(B, rank, lambda, sm_dyn);
                    cluster.sync();
                    BA_TICK(4);
                    phase_gather(B, rank, CL);
                    cluster.sync();
                    BA_TICK(5);
                }
                if (rank == 0) {
                    int ok = 1;
                    if (B.Pf) ok = phase_solve(B, lambda, sm_dyn, aux, &sflag, CL);
                    if (tid == 0) {
                        B.chol_fail[0] = !ok;
                        B.chol_fail[1] = B.stop ? *(volatile const int*)B.stop : 0;
                    }
                }
                cluster.sync();
                BA_TICK(6);
                const bool fail = *(volatile int*)B.chol_fail != 0;
                {
                    double s = block_sum(phase_update(B, rank, ct, cn, lambda, !fail, sm_dyn), red);
                    if (tid == 0) B.parts[64 + rank] = s;
                }
                cluster.sync();
                BA_TICK(7);
                {
                    double s = block_sum(phase_errors(B, ct, cn, robust), red);
                    if (tid == 0) B.parts[rank] = s;
                }
                cluster.sync();
                BA_TICK(8);
                if (tid == 0) {  // levenberg.cpp:96-150, identical in every CTA
                    double chi_raw = 0, scale = 0;
                    for (int r = 0; r < CL; r++) { chi_raw += B.parts[r]; scale += B.parts[64 + r]; }
                    const int stop = ((volatile int*)B.chol_fail)[1];
                    double tempChi = fail ? DBL_MAX : chi_raw;
                    double rho = L.currentChi - tempChi;
                    scale += 1e-3;
                    rho /= scale;
                    int rej = 0, lam_bad = 0;
                    if (rho > 0 && isfinite(tempChi) && !fail) {
                        double alpha = 1. - pow((2 * rho - 1), 3.0);
                        alpha = fmin(alpha, 2. / 3.);
                        L.lambda *= fmax(1. / 3., alpha);
                        L.ni = 2;
                        L.currentChi = tempChi;
                    } else {
                        L.lambda *= L.ni;
                        L.ni *= 2;
                        rej = 1;
                        if (!isfinite(L.lambda)) lam_bad = 1;
                    }
                    if (!lam_bad) L.qmax++;
                    L.rho = rho;
                    L.reject = rej;
                    const int cont = !lam_bad && rho < 0 && L.qmax < 10 && !stop;
                    L.cont_trial = cont;
                    if (stop) L.stopped = 1;
                    if (!cont) {  // sparse_optimizer.cpp:403-436
                        if (L.qmax == 10 || rho == 0 || !isfinite(L.lambda)) L.ok = 0;
                        L.curChi2 = (float)chi_raw;
                        L.chi2Diff = L.prevChi2 - L.curChi2;
                        if (rank == 0 && L.ntrace < 64) {
                            B.res->trace[2 * L.ntrace] = chi_raw;
                            B.res->trace[2 * L.ntrace + 1] = L.qmax;
                        }
                        if (L.ntrace < 64) L.ntrace++;
                        L.it++;
                        L.cont_iter = L.it < max_iters && !stop && L.ok && L.chi2Diff > 1.0f;
                    }
                }
                __syncthreads();
                if (L.reject) phase_restore(B, rank, ct, cn);
                BA_TICK(9);
            }
            cluster.sync();  // estimates settled (accepted or restored) before anyone linearizes or flags outliers
            BA_TICK(9);
        }
        if (tid == 0 && rank == 0) B.res->iters[stage] = L.it;
        __syncthreads();
        if (L.stopped) {  // GlobalOptimizerG2O::optimize: no second stage after stopASAP
            if (stage == 0 && tid == 0 && rank == 0) B.res->iters[1] = 0;
            break;
        }
    }
    cluster.sync();
    // ---- getResults
    if (rank == 0 && tid == 0) B.res->ntrace = L.ntrace;
    for (int t = ct; t < B.P; t += cn) {
        float* m = B.p44_out + 16 * t;
        if (B.free_idx[t] < 0) {
            for (int k = 0; k < 16; k++) m[k] = B.p44_in[16 * t + k];
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
    for (int i = ct; i < B.M; i += cn) {
        const int pi = B.obs_pose[i];
        Pose T = load_pose(B.pose + 7 * pi);
        const double* X = B.pt + 3 * B.obs_lm[i];
        double x[3] = {X[0], X[1], X[2]}, p[3];
        se3_map(T, x, p);
        int b = 0;
        if (B.stereo[i]) {
            if (B.chi2[i] > (double)B.chi3d || !(p[2] > 0.0)) b = 1;
        } else if (B.chi2[i] > (double)B.chi2d) b = 1;
        if (!b) {  // pincam = pose_f2g (f32) * point (f32), z < 0  (:515-518)
            float m8, m9, m10, m11;
            if (B.free_idx[pi] < 0) {
                const float* m = B.p44_in + 16 * pi;
                m8 = m[8]; m9 = m[9]; m10 = m[10]; m11 = m[11];
            } else {
                double R[9];
                quat_to_R(T.q, R);
                m8 = (float)R[6]; m9 = (float)R[7]; m10 = (float)R[8]; m11 = (float)T.t[2];
            }
            const float zc = m8 * (float)x[0] + m9 * (float)x[1] + m10 * (float)x[2] + m11;
            if (zc < 0) b = 1;
        }
        B.bad[i] = (uint8_t)(b | (B.active[i] ? 0 : 2));  // bit 1: the edge left the problem after stage 1 (level 1)
    }
    BA_TICK(10);
#undef BA_TICK
}

}  // namespace

// ---- host side: one launch for a batch of windows ------------------------------------------------------------------------------
struct BaArena {
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    }
};

// host-side planning / staging of a batch: a few worker threads, each taking windows in turn.  The count is bounded by this
// process's share of the host cores (one process per GPU, several mapper contexts per process): oversubscribing them was measured
// to stretch the 8-GPU step by 50 %.  uco_b200_ba_set_host_threads / UCO_BA_HOST_THREADS override.
template <class F>
static void ba_parallel_for(int n, int want, F&& f) {
    static const int auto_cap = [] {
        if (const char* e = getenv("UCO_BA_HOST_THREADS")) return std::max(1, atoi(e));
        int ndev = 1;
        cudaGetDeviceCount(&ndev);
        const int hc = (int)std::thread::hardware_concurrency();
        return std::max(1, std::min(8, hc / std::max(1, ndev) / 4));
    }();
    const int nt = std::min(n, want > 0 ? want : auto_cap);
    if (nt <= 1) {
        for (int i = 0; i < n; i++) f(i);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] { for (int i = t; i < n; i += nt) f(i); });
    for (auto& t : th) t.join();
}

int ba_cluster_solve_batch(uco_b200_ctx* ctx, int n, const uco_ba_problem* const* pbs, const volatile unsigned char* stop, uco_ba_result* const* res) {
    if (n <= 0) return UCO_OK;
    for (int i = 0; i < n; i++)
        if (pbs[i]->pose_cam || pbs[i]->n_markers > 0)   // the callers route these to the sharded solver
            return uco_fail(ctx, UCO_E_INVALID, "ba_solve: markers / per-keyframe cameras are not handled by the cluster-resident solver");
    static const bool trace = getenv("UCO_BA_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t0 = now();
    std::vector<BaPlan> plans(n);
    int max_n = 0;
    const int CL = ctx->ba_cluster_size > 0 ? ctx->ba_cluster_size : 8;
    {  // the planner is pure host work (~0.5 ms per window): one thread per window
        std::vector<int> rcs(n, UCO_OK);
        std::vector<std::string> errs(n);
        auto job = [&](int i) {
            uco_b200_ctx tmp;
            rcs[i] = ba_plan_build(&tmp, *pbs[i], BA_UNIT, plans[i], CL, BS, NW);
            errs[i] = tmp.err;
        };
        ba_parallel_for(n, ctx->ba_host_threads, job);
        for (int i = 0; i < n; i++) {
            if (rcs[i] != UCO_OK) return uco_fail(ctx, rcs[i], "%s", errs[i].c_str());
            max_n = std::max(max_n, 6 * plans[i].Pf);
        }
    }
    if (max_n > BA_CLUSTER_MAX_N) return uco_fail(ctx, UCO_E_INVALID, "ba cluster path: reduced system of %d unknowns", max_n);
    // layout: [inputs of all windows | CbDev array | stop flag][work][outputs of all windows]
    struct Off {
        size_t p44, free_idx, free_list, lm_ptr, obs_pose, obs_lm, obs_free, pose_ptr, pose_obs, blk_unit_ptr, blk_ij, diag_blk, unit, con, z, info,
            stereo, pt_in, cta_lm, cta_chunk_ptr, chunk_lm, gw_ptr, gseg, gcon, wblk;
        size_t pose_bak, pt_bak, err, lmc, Hll, bl, W, Y, Dinv, db, Hpp, bp, part, partb, xp, parts, chol_fail, active;
        size_t pose, p44o, pt, chi2, level_dummy, bad, resd;
    };
    std::vector<Off> off(n);
    BaArena A;
    for (int i = 0; i < n; i++) {
        const BaPlan& p = plans[i];
        Off& o = off[i];
        o.p44 = A.take(64 * (size_t)p.P); o.free_idx = A.take(4 * (size_t)p.P); o.free_list = A.take(4 * (size_t)(p.Pf + 1));
        o.lm_ptr = A.take(4 * (size_t)(p.N + 1)); o.obs_pose = A.take(4 * (size_t)(p.M + 1)); o.obs_lm = A.take(4 * (size_t)(p.M + 1));
        o.obs_free = A.take(4 * (size_t)(p.M + 1));
        o.pose_ptr = A.take(4 * (size_t)(p.Pf + 1)); o.pose_obs = A.take(4 * (p.pose_obs.size() + 1));
        o.blk_unit_ptr = A.take(4 * (p.blk_unit_ptr.size() + 1)); o.blk_ij = A.take(8 * (p.blk_ij.size() + 1));
        o.diag_blk = A.take(4 * (size_t)(p.Pf + 1)); o.unit = A.take(16 * (p.unit.size() + 1)); o.con = A.take(8 * (p.con.size() + 1));
        o.z = A.take(12 * (size_t)(p.M + 1)); o.info = A.take(4 * (size_t)(p.M + 1)); o.stereo = A.take((size_t)p.M + 1);
        o.pt_in = A.take(12 * (size_t)(p.N + 1));
        o.cta_lm = A.take(4 * p.cta_lm.size()); o.cta_chunk_ptr = A.take(4 * p.cta_chunk_ptr.size()); o.chunk_lm = A.take(4 * p.chunk_lm.size());
        o.gw_ptr = A.take(4 * (p.gw_ptr.size() + 1)); o.gseg = A.take(16 * (p.gseg.size() + 1)); o.gcon = A.take(4 * (p.gcon.size() + 1));
        o.wblk = A.take(4 * (p.wblk.size() + 1));
    }
    const size_t o_probs = A.take(sizeof(CbDev) * (size_t)n);
    const size_t in_bytes = A.off;
    for (int i = 0; i < n; i++) {
        const BaPlan& p = plans[i];
        Off& o = off[i];
        const size_t M1 = p.M + 1, N1 = p.N + 1, F1 = p.Pf + 1, U1 = std::max(p.unit.size(), (size_t)CL * p.blk_ij.size()) + 1;
        o.pose_bak = A.take(56 * (size_t)p.P); o.pt_bak = A.take(24 * N1); o.err = A.take(24 * M1); o.lmc = A.take(72 * M1);
        o.Hll = A.take(48 * N1); o.bl = A.take(24 * N1); o.W = A.take(144 * M1); o.Y = A.take(144 * M1); o.Dinv = A.take(48 * N1);
        o.db = A.take(24 * N1); o.Hpp = A.take(288 * F1); o.bp = A.take(48 * F1); o.part = A.take(288 * U1); o.partb = A.take(48 * U1);
        o.xp = A.take(48 * F1); o.parts = A.take(8 * 96); o.chol_fail = A.take(16); o.active = A.take(M1);
    }
    const size_t out_begin = A.off;
    for (int i = 0; i < n; i++) {
        const BaPlan& p = plans[i];
        Off& o = off[i];
        o.pose = A.take(56 * (size_t)p.P); o.p44o = A.take(64 * (size_t)p.P); o.pt = A.take(24 * (size_t)(p.N + 1));
        o.chi2 = A.take(8 * (size_t)(p.M + 1)); o.bad = A.take((size_t)p.M + 1); o.resd = A.take(sizeof(CbResult));
    }
    const size_t out_bytes = A.off - out_begin;
    auto t1 = now();
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_BA, A.off);
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_BA, in_bytes);
    uint8_t* ho = (uint8_t*)uco_pinned(ctx, WS_BA_OUT, out_bytes + 64);
    if (!d || !h || !ho) return UCO_E_NOMEM;
    // the stop flag the kernel polls: pinned, mapped (zero-copy) memory owned by the context
    volatile int* hstop = (volatile int*)uco_pinned(ctx, WS_BA_STOP, 64);
    if (!hstop) return UCO_E_NOMEM;
    *hstop = (stop && *stop) ? 1 : 0;
    int* dstop = nullptr;
    UCO_CUDA(ctx, cudaHostGetDevicePointer((void**)&dstop, (void*)hstop, 0));
    CbDev* hp = (CbDev*)(h + o_probs);
    auto fill = [&](int i) {
        const BaPlan& p = plans[i];
        const uco_ba_problem& pb = *pbs[i];
        const Off& o = off[i];
        memcpy(h + o.p44, pb.poses44, 64 * (size_t)p.P);
        memcpy(h + o.free_idx, p.free_idx.data(), 4 * (size_t)p.P);
        memcpy(h + o.free_list, p.free_list.data(), 4 * (size_t)p.Pf);
        memcpy(h + o.lm_ptr, p.lm_ptr.data(), 4 * (size_t)(p.N + 1));
        memcpy(h + o.obs_pose, p.s_pose.data(), 4 * (size_t)p.M);
        memcpy(h + o.obs_lm, p.s_lm.data(), 4 * (size_t)p.M);
        memcpy(h + o.obs_free, p.s_free.data(), 4 * (size_t)p.M);
        memcpy(h + o.pose_ptr, p.pose_ptr.data(), 4 * (size_t)(p.Pf + 1));
        memcpy(h + o.pose_obs, p.pose_obs.data(), 4 * p.pose_obs.size());
        memcpy(h + o.blk_unit_ptr, p.blk_unit_ptr.data(), 4 * p.blk_unit_ptr.size());
        memcpy(h + o.blk_ij, p.blk_ij.data(), 8 * p.blk_ij.size());
        memcpy(h + o.diag_blk, p.diag_blk.data(), 4 * (size_t)p.Pf);
        memcpy(h + o.unit, p.unit.data(), 16 * p.unit.size());
        memcpy(h + o.con, p.con.data(), 8 * p.con.size());
        float* z = (float*)(h + o.z);
        float* info = (float*)(h + o.info);
        uint8_t* st = h + o.stereo;
        for (int k = 0; k < p.M; k++) {
            const int j = p.order[k];
            const bool s = pb.obs_stereo && pb.obs_stereo[j];
            z[3 * k] = pb.obs_uv[2 * j]; z[3 * k + 1] = pb.obs_uv[2 * j + 1]; z[3 * k + 2] = s ? pb.obs_ur[j] : 0.f;
            info[k] = pb.obs_inv_sigma2[j];
            st[k] = s;
        }
        if (p.N) memcpy(h + o.pt_in, pb.points3, 12 * (size_t)p.N);
        memcpy(h + o.cta_lm, p.cta_lm.data(), 4 * p.cta_lm.size());
        memcpy(h + o.cta_chunk_ptr, p.cta_chunk_ptr.data(), 4 * p.cta_chunk_ptr.size());
        memcpy(h + o.chunk_lm, p.chunk_lm.data(), 4 * p.chunk_lm.size());
        if (p.fused) {
            memcpy(h + o.gw_ptr, p.gw_ptr.data(), 4 * p.gw_ptr.size());
            memcpy(h + o.gseg, p.gseg.data(), 16 * p.gseg.size());
            memcpy(h + o.gcon, p.gcon.data(), 4 * p.gcon.size());
            memcpy(h + o.wblk, p.wblk.data(), 4 * p.wblk.size());
        }
        CbDev& B = hp[i];
        memset(&B, 0, sizeof(B));
        B.P = p.P; B.N = p.N; B.M = p.M; B.Pf = p.Pf; B.n = 6 * p.Pf; B.nblk = (int)p.blk_ij.size(); B.nunits = (int)p.unit.size();
        B.n_iters = pb.n_iters;
        B.p44_in = (const float*)(d + o.p44); B.free_idx = (const int*)(d + o.free_idx); B.free_list = (const int*)(d + o.free_list);
        B.lm_ptr = (const int*)(d + o.lm_ptr); B.obs_pose = (const int*)(d + o.obs_pose); B.obs_lm = (const int*)(d + o.obs_lm); B.obs_free = (const int*)(d + o.obs_free);
        B.pose_ptr = (const int*)(d + o.pose_ptr); B.pose_obs = (const int*)(d + o.pose_obs);
        B.blk_unit_ptr = (const int*)(d + o.blk_unit_ptr); B.blk_ij = (const int2*)(d + o.blk_ij); B.diag_blk = (const int*)(d + o.diag_blk);
        B.unit = (const int4*)(d + o.unit); B.con = (const int2*)(d + o.con); B.z = (const float*)(d + o.z); B.info = (const float*)(d + o.info);
        B.stereo = d + o.stereo; B.pt_in = (const float*)(d + o.pt_in);
        B.cta_lm = (const int*)(d + o.cta_lm); B.cta_chunk_ptr = (const int*)(d + o.cta_chunk_ptr); B.chunk_lm = (const int*)(d + o.chunk_lm);
        B.fused = p.fused; B.n_chunks = (int)p.chunk_lm.size() - 1;
        B.gw_ptr = (const int*)(d + o.gw_ptr); B.gseg = (const int4*)(d + o.gseg); B.gcon = (const uint32_t*)(d + o.gcon); B.wblk = (const int*)(d + o.wblk);
        B.pose_bak = (double*)(d + o.pose_bak); B.pt_bak = (double*)(d + o.pt_bak); B.err = (double*)(d + o.err); B.lmc = (double*)(d + o.lmc);
        B.Hll = (double*)(d + o.Hll); B.bl = (double*)(d + o.bl); B.W = (double*)(d + o.W); B.Y = (double*)(d + o.Y);
        B.Dinv = (double*)(d + o.Dinv); B.db = (double*)(d + o.db); B.Hpp = (double*)(d + o.Hpp); B.bp = (double*)(d + o.bp);
        B.part = (double*)(d + o.part); B.partb = (double*)(d + o.partb); B.xp = (double*)(d + o.xp); B.parts = (double*)(d + o.parts);
        B.chol_fail = (int*)(d + o.chol_fail); B.active = d + o.active;
        B.pose = (double*)(d + o.pose); B.p44_out = (float*)(d + o.p44o); B.pt = (double*)(d + o.pt); B.chi2 = (double*)(d + o.chi2);
        B.bad = d + o.bad; B.res = (CbResult*)(d + o.resd);
        B.cam.fx = pb.fx; B.cam.fy = pb.fy; B.cam.cx = pb.cx; B.cam.cy = pb.cy; B.cam.bf = pb.bf; B.cam.bf_f = pb.bf;
        B.chi2d = 5.99f; B.chi3d = 7.815f;
        B.d2 = (double)sqrtf(B.chi2d); B.d3 = (double)sqrtf(B.chi3d);
        B.stop = dstop;
    };
    ba_parallel_for(n, ctx->ba_host_threads, fill);
    auto t2 = now();
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + out_begin, 0, out_bytes, s));
    bool any_fused = false;
    for (int i = 0; i < n; i++) any_fused = any_fused || plans[i].fused;
    const size_t smem = std::max<size_t>(std::max<size_t>(8 * ((size_t)(max_n + 1) * (max_n + 2) / 2), 8 * (size_t)BS * (WPAD + 9)), any_fused ? 8 * (size_t)PG_DOUBLES : 0);
    UCO_CUDA(ctx, cudaFuncSetAttribute(ba_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL > 8) UCO_CUDA(ctx, cudaFuncSetAttribute(ba_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n * CL));
    cfg.blockDim = dim3(BS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    UCO_CUDA(ctx, cudaEventRecord(uco_ba_events(ctx)[0], s));
    UCO_CUDA(ctx, cudaLaunchKernelEx(&cfg, ba_cluster_kernel, (const CbDev*)(d + o_probs)));
    UCO_LAUNCH_CHECK(ctx);
    UCO_CUDA(ctx, cudaEventRecord(uco_ba_events(ctx)[1], s));
    UCO_CUDA(ctx, cudaMemcpyAsync(ho, d + out_begin, out_bytes, cudaMemcpyDeviceToHost, s));
    if (stop) {  // forward an asynchronous stopASAP to the flag the kernel polls
        cudaError_t q;
        while ((q = cudaStreamQuery(s)) == cudaErrorNotReady)
            if (*stop) *hstop = 1;
        if (q != cudaSuccess) return uco_fail(ctx, UCO_E_CUDA, "ba cluster kernel -> %s", cudaGetErrorString(q));
    }
    UCO_CUDA(ctx, uco_sleep_sync(ctx));   // milliseconds of device work: sleep, do not spin
    auto t3 = now();
    float ms = 0;
    cudaEventElapsedTime(&ms, uco_ba_events(ctx)[0], uco_ba_events(ctx)[1]);
    for (int i = 0; i < n; i++) {
        const BaPlan& p = plans[i];
        const Off& o = off[i];
        uco_ba_result& r = *res[i];
        const uint8_t* base = ho - out_begin;
        if (r.pose7) memcpy(r.pose7, base + o.pose, 56 * (size_t)p.P);
        if (r.poses44) memcpy(r.poses44, base + o.p44o, 64 * (size_t)p.P);
        if (r.points3 && p.N) memcpy(r.points3, base + o.pt, 24 * (size_t)p.N);
        const double* chi = (const double*)(base + o.chi2);
        const uint8_t* bad = base + o.bad;
        for (int k = 0; k < p.M; k++) {
            const int j = p.order[k];
            if (r.obs_chi2) r.obs_chi2[j] = chi[k];
            if (r.obs_bad) r.obs_bad[j] = bad[k] & 1;
            if (r.obs_level) r.obs_level[j] = (bad[k] >> 1) & 1;
        }
        const CbResult* cr = (const CbResult*)(base + o.resd);
        if (r.trace) {
            memset(r.trace, 0, sizeof(double) * 128);
            memcpy(r.trace, cr->trace, sizeof(double) * 2 * (size_t)std::min(cr->ntrace, 64));
        }
        r.iters[0] = cr->iters[0];
        r.iters[1] = cr->iters[1];
        r.device_ms = ms;
        if (r.profile) memcpy(r.profile, cr->phase_cycles, sizeof(double) * 16);
    }
    if (trace) {
        auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::micro>(b - a).count();
        };
        fprintf(stderr, "[ba] n=%d plan %.0f us  fill %.0f us  h2d+kernel+d2h %.0f us (kernel %.0f us, in %zu B, out %zu B)  unpack %.0f us\n", n,
                us(t0, t1), us(t1, t2), us(t2, t3), ms * 1e3, in_bytes, out_bytes, us(t3, now()));
    }
    return UCO_OK;
}
