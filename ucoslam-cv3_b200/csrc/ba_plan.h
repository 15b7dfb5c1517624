// ba_plan.h — host-side structure of one bundle-adjustment window (built once per solve, no arithmetic):
// free-pose numbering, observations sorted by landmark, per-pose observation lists and the Schur gather lists
// (for every 6x6 block (i <= j) of the reduced system the (obs_a, obs_b) pairs of the landmarks both poses see, in
// landmark order, cut into units of <= `unit` contributions).  Mirrors what BlockSolver::buildStructure prepares
// (3rdparty/g2o/g2o/core/block_solver.hpp:103-312) for the graph GlobalOptimizerG2O::setParams assembles.
#pragma once
#include <algorithm>
#include "common.cuh"
#include "ba_math.cuh"
#include <vector>

constexpr int BA_UNIT = 64;            // contributions per Schur-gather unit (two index registers per lane)
constexpr int BA_CLUSTER_MAX_N = 228;  // reduced-system size the cluster kernel holds in shared memory (38 free keyframes)

struct CbResult {
    int iters[2];
    int ntrace, pad;
    double trace[128];
    double phase_cycles[16];
};

constexpr int BA_GSLOTS = 8;   // blocks a warp owns in the fused prep + gather phase (register accumulators)
struct CbDev {  // one window as the cluster kernel sees it (device pointers)
    int P, N, M, Pf, n, nblk, nunits, n_iters;
    const float *p44_in, *pt_in, *z, *info;
    const uint8_t* stereo;
    const int *free_idx, *free_list, *lm_ptr, *obs_pose, *obs_lm, *obs_free, *pose_ptr, *pose_obs, *blk_unit_ptr, *diag_blk;
    const int *cta_lm, *cta_chunk_ptr, *chunk_lm;  // landmark range per CTA, its chunks (whole landmarks, <= CTA-size observations)
    const int2 *blk_ij, *con;
    const int4* unit;
    // fused prep + Schur gather (windows of <= BA_GSLOTS blocks per warp): per (chunk, warp) the segments (block | slot << 16 | diag << 24,
    // first entry, one past the last (padded to 4)) and their entries (chunk-local observation a | b << 16; diagonal: a | landmark << 16)
    int fused, n_chunks;
    const int* gw_ptr;
    const int4* gseg;
    const uint32_t* gcon;
    const int* wblk;       // [warp * BA_GSLOTS + slot] -> block (or -1)
    double *pose, *pose_bak, *pt, *pt_bak, *err, *chi2, *lmc, *Hll, *bl, *W, *Y, *Dinv, *db, *Hpp, *bp, *part, *partb, *xp, *parts;
    int* chol_fail;  // [0] reduced solve failed in this trial, [1] stop flag as latched by CTA 0
    uint8_t *active, *bad;
    float* p44_out;
    CbResult* res;
    ba::Cam cam;
    double d2, d3;
    float chi2d, chi3d;
    const int* stop;
};

struct BaPlan {
    int P = 0, N = 0, M = 0, Pf = 0, ncon = 0;  // ncon: contributions before padding
    std::vector<int> free_idx, free_list, lm_ptr, order, s_pose, s_lm, s_free, pose_ptr, pose_obs, blk_unit_ptr, diag_blk;
    std::vector<int2> blk_ij, con;
    std::vector<int4> unit;  // (block, first contribution, one past the last, block is diagonal)
    std::vector<int> cta_lm, cta_chunk_ptr, chunk_lm;
    bool fused = false;                       // fused prep + gather lists instead of con / unit
    std::vector<int> gw_ptr, wblk;
    std::vector<int4> gseg;
    std::vector<uint32_t> gcon;
};

inline int ba_validate(uco_b200_ctx* ctx, const uco_ba_problem* pb) {
    if (!pb) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: null problem");
    const int P = pb->n_poses, N = pb->n_points, M = pb->n_obs;
    if (P <= 0 || N < 0 || M < 0 || pb->n_iters < 0) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: bad sizes");
    if (!pb->poses44 || !pb->fixed || (N && !pb->points3) || (M && (!pb->obs_pose || !pb->obs_point || !pb->obs_uv || !pb->obs_inv_sigma2)))
        return uco_fail(ctx, UCO_E_INVALID, "ba_solve: null input array");
    for (int i = 0; i < M; i++) {
        if ((unsigned)pb->obs_pose[i] >= (unsigned)P || (unsigned)pb->obs_point[i] >= (unsigned)N)
            return uco_fail(ctx, UCO_E_INVALID, "ba_solve: observation %d references pose %d / point %d out of range", i,
                            pb->obs_pose[i], pb->obs_point[i]);
        if (pb->obs_stereo && pb->obs_stereo[i] && !pb->obs_ur)
            return uco_fail(ctx, UCO_E_INVALID, "ba_solve: stereo observation without obs_ur");
    }
    return UCO_OK;
}

inline int ba_free_poses(const uco_ba_problem* pb) {
    int f = 0;
    for (int i = 0; i < pb->n_poses; i++) f += pb->fixed[i] ? 0 : 1;
    return f;
}

// nwarps > 0: the cluster kernel's warp count; windows with <= BA_GSLOTS blocks per warp get the lists of the fused prep + gather phase
// (per-chunk, per-owner-warp segments over chunk-local indices) INSTEAD of the per-block contribution list `con`
inline int ba_plan_build(uco_b200_ctx* ctx, const uco_ba_problem& pb, int unit, BaPlan& p, int n_cta, int cta_threads, int nwarps = 0) {
    int rc = ba_validate(ctx, &pb);
    if (rc != UCO_OK) return rc;
    const int P = pb.n_poses, N = pb.n_points, M = pb.n_obs;
    p.P = P; p.N = N; p.M = M;
    p.free_idx.assign(P, -1);
    p.free_list.clear();
    for (int i = 0; i < P; i++)
        if (!pb.fixed[i]) {
            p.free_idx[i] = (int)p.free_list.size();
            p.free_list.push_back(i);
        }
    const int Pf = p.Pf = (int)p.free_list.size();
    p.lm_ptr.assign(N + 1, 0);
    for (int i = 0; i < M; i++) p.lm_ptr[pb.obs_point[i] + 1]++;
    for (int l = 0; l < N; l++) p.lm_ptr[l + 1] += p.lm_ptr[l];
    p.order.resize(M);
    {
        std::vector<int> fill(p.lm_ptr.begin(), p.lm_ptr.end() - 1);
        for (int i = 0; i < M; i++) p.order[fill[pb.obs_point[i]]++] = i;  // sorted position -> caller index (stable)
    }
    p.s_pose.resize(M);
    p.s_lm.resize(M);
    for (int k = 0; k < M; k++) {
        p.s_pose[k] = pb.obs_pose[p.order[k]];
        p.s_lm[k] = pb.obs_point[p.order[k]];
    }
    p.pose_ptr.assign(Pf + 1, 0);
    for (int k = 0; k < M; k++) {
        int f = p.free_idx[p.s_pose[k]];
        if (f >= 0) p.pose_ptr[f + 1]++;
    }
    for (int f = 0; f < Pf; f++) p.pose_ptr[f + 1] += p.pose_ptr[f];
    p.pose_obs.resize(p.pose_ptr[Pf]);
    {
        std::vector<int> fill(p.pose_ptr.begin(), p.pose_ptr.end() - 1);
        for (int k = 0; k < M; k++) {
            int f = p.free_idx[p.s_pose[k]];
            if (f >= 0) p.pose_obs[fill[f]++] = k;
        }
    }
    // contributions per block (i <= j), in landmark order.  Within a landmark every pose pair appears once (a pose observes
    // a landmark once), so visiting position pairs a <= b and ordering each by free-pose index gives every block its
    // contributions sorted by landmark: Y comes from the lower-indexed pose (the block row), W from the higher one.
    p.s_free.resize(M);
    for (int k = 0; k < M; k++) p.s_free[k] = p.free_idx[p.s_pose[k]];
    // landmark ranges per CTA (balanced by observation count), cut into chunks of whole landmarks
    p.cta_lm.assign(n_cta + 1, N);
    p.cta_lm[0] = 0;
    {
        int l = 0;
        for (int c = 1; c < n_cta; c++) {
            const long long target = (long long)M * c / n_cta;
            while (l < N && p.lm_ptr[l] < target) l++;
            p.cta_lm[c] = l;
        }
    }
    p.chunk_lm.assign(1, 0);
    p.cta_chunk_ptr.assign(1, 0);
    for (int c = 0; c < n_cta; c++) {
        int l = p.cta_lm[c];
        while (l < p.cta_lm[c + 1]) {
            int e = l;
            while (e < p.cta_lm[c + 1] && e - l < cta_threads && p.lm_ptr[e + 1] - p.lm_ptr[l] <= cta_threads) e++;
            if (e == l) return uco_fail(ctx, UCO_E_INVALID, "ba_solve: landmark %d has %d observations (> %d per chunk)", l, p.lm_ptr[l + 1] - p.lm_ptr[l], cta_threads);
            p.chunk_lm.push_back(e);
            l = e;
        }
        p.cta_chunk_ptr.push_back((int)p.chunk_lm.size() - 1);
    }
    std::vector<int> cnt((size_t)Pf * Pf, 0);
    const int* sf = p.s_free.data();
    // When the fused lists may be built (at most BA_GSLOTS blocks per warp even if every pose pair is covisible) the contributions are
    // counted per (chunk, pose pair) in ONE pass: the sums give the blocks, the per-chunk counts the segment sizes of the fused lists.
    const int n_chunks_all = (int)p.chunk_lm.size() - 1;
    const bool may_fuse = nwarps > 0 && Pf * (Pf + 1) / 2 <= BA_GSLOTS * nwarps && cta_threads < 65535;
    std::vector<int> ccnt(may_fuse ? (size_t)n_chunks_all * Pf * Pf : 0, 0);
    // free observations in landmark order, compacted (fixed-pose observations contribute nothing): position -> sorted observation, free pose
    std::vector<int> fo_ptr(N + 1, 0), fo_obs, fo_free;
    fo_obs.reserve(M); fo_free.reserve(M);
    for (int l = 0; l < N; l++) {
        for (int a = p.lm_ptr[l]; a < p.lm_ptr[l + 1]; a++)
            if (sf[a] >= 0) { fo_obs.push_back(a); fo_free.push_back(sf[a]); }
        fo_ptr[l + 1] = (int)fo_obs.size();
    }
    const int* fof = fo_free.data();
    for (int ch = 0; ch < (may_fuse ? n_chunks_all : 1); ch++) {
        int* cc = may_fuse ? ccnt.data() + (size_t)ch * Pf * Pf : cnt.data();
        const int lb = may_fuse ? p.chunk_lm[ch] : 0, le = may_fuse ? p.chunk_lm[ch + 1] : N;
        for (int l = lb; l < le; l++) {
            const int e0 = fo_ptr[l], e1 = fo_ptr[l + 1];
            for (int a = e0; a < e1; a++) {
                const int fa = fof[a];
                cc[fa * Pf + fa]++;
                for (int b = a + 1; b < e1; b++) {
                    const int fb = fof[b];
                    if (fb == fa) continue;                       // a pose listed twice for a landmark: never in a valid map
                    const int lo = fa < fb ? fa : fb, hi = fa < fb ? fb : fa;
                    cc[lo * Pf + hi]++;
                }
            }
        }
        if (may_fuse)
            for (int k = 0; k < Pf * Pf; k++) cnt[k] += cc[k];
    }
    // Blocks in (i, j >= i) order.  A block's contributions are written straight into the PADDED layout the gather kernel reads:
    // units of <= `unit` contributions (unit is a multiple of four), the last unit of a block padded to a multiple of four with the
    // all-zero dummy observation M (diagonal blocks: dummy landmark N), so the gather loop needs no bounds checks.
    std::vector<int> pos((size_t)Pf * Pf, -1), blk_ptr(1, 0);   // pos: next write position of block (i, j) in the padded array
    p.blk_ij.clear();
    p.diag_blk.assign(Pf, -1);
    p.unit.clear();
    p.blk_unit_ptr.assign(1, 0);
    int padded_total = 0, ncon = 0;
    for (int i = 0; i < Pf; i++)
        for (int j = i; j < Pf; j++) {
            const int c = cnt[(size_t)i * Pf + j];
            if (!c && i != j) continue;   // diagonal blocks always exist (Hpp + lambda I)
            const int k = (int)p.blk_ij.size();
            if (i == j) p.diag_blk[i] = k;
            p.blk_ij.push_back(make_int2(i, j));
            pos[(size_t)i * Pf + j] = padded_total;
            for (int c0 = 0; c0 < c; c0 += unit) {
                const int c1 = std::min(c0 + unit, c);
                p.unit.push_back(make_int4(k, padded_total + c0, padded_total + ((c1 + 3) & ~3), i == j));
            }
            p.blk_unit_ptr.push_back((int)p.unit.size());
            ncon += c;
            padded_total += (c + 3) & ~3;
            blk_ptr.push_back(padded_total);
        }
    const int nblk = (int)p.blk_ij.size();
    p.fused = may_fuse;   // nblk <= Pf (Pf + 1) / 2 <= BA_GSLOTS * nwarps
    p.con.resize(p.fused ? 0 : padded_total);
    if (!p.fused) {
        int2* con = p.con.data();
        int* ps = pos.data();
        for (int l = 0; l < N; l++) {
            const int e0 = p.lm_ptr[l], e1 = p.lm_ptr[l + 1];
            for (int a = e0; a < e1; a++) {
                const int fa = sf[a];
                if (fa < 0) continue;
                con[ps[(size_t)fa * Pf + fa]++] = make_int2(a, l);  // diagonal: (observation, landmark)
                for (int b = a + 1; b < e1; b++) {
                    const int fb = sf[b];
                    if (fb < 0 || fb == fa) continue;
                    if (fa < fb) con[ps[(size_t)fa * Pf + fb]++] = make_int2(a, b);
                    else con[ps[(size_t)fb * Pf + fa]++] = make_int2(b, a);
                }
            }
        }
        for (size_t k = 0; k < p.blk_ij.size(); k++) {   // the padding of each block's last unit
            const int2 ij = p.blk_ij[k];
            const bool dg = ij.x == ij.y;
            for (int q = ps[(size_t)ij.x * Pf + ij.y]; q < blk_ptr[k + 1]; q++) con[q] = make_int2(M, dg ? N : M);
        }
    }
    p.ncon = ncon;
    p.gw_ptr.clear(); p.gseg.clear(); p.gcon.clear(); p.wblk.clear();
    if (p.fused) {
        // block ownership: heaviest block first to the least loaded warp that still has a free slot (same for every CTA)
        std::vector<int> ord(nblk), load(nwarps, 0), nslot(nwarps, 0), bcnt(nblk), blk_at((size_t)Pf * Pf, -1);
        for (int k = 0; k < nblk; k++) {
            ord[k] = k;
            bcnt[k] = cnt[(size_t)p.blk_ij[k].x * Pf + p.blk_ij[k].y];
            blk_at[(size_t)p.blk_ij[k].x * Pf + p.blk_ij[k].y] = k;
        }
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return bcnt[a] > bcnt[b]; });
        p.wblk.assign((size_t)nwarps * BA_GSLOTS, -1);
        std::vector<int> slot_of(nblk, 0), owner(nblk, 0);
        for (int k : ord) {
            int w = -1;
            for (int c = 0; c < nwarps; c++)
                if (nslot[c] < BA_GSLOTS && (w < 0 || load[c] < load[w])) w = c;
            owner[k] = w; slot_of[k] = nslot[w];
            p.wblk[(size_t)w * BA_GSLOTS + nslot[w]++] = k;
            load[w] += bcnt[k] + 8;
        }
        const int n_chunks = (int)p.chunk_lm.size() - 1;
        p.gw_ptr.assign((size_t)n_chunks * nwarps + 1, 0);
        p.gseg.reserve((size_t)n_chunks * nblk);
        p.gcon.reserve((size_t)ncon + 4 * (size_t)n_chunks * nblk + 32);
        std::vector<int> cb(nblk), at(nblk);
        const int* ba = blk_at.data();
        for (int ch = 0; ch < n_chunks; ch++) {
            const int l0 = p.chunk_lm[ch], l1 = p.chunk_lm[ch + 1], o0 = p.lm_ptr[l0], nobs = p.lm_ptr[l1] - o0, nl = l1 - l0;
            for (int k = 0; k < nblk; k++) cb[k] = ccnt[(size_t)ch * Pf * Pf + (size_t)p.blk_ij[k].x * Pf + p.blk_ij[k].y];
            for (int w = 0; w < nwarps; w++) {
                p.gw_ptr[(size_t)ch * nwarps + w] = (int)p.gseg.size();
                for (int sl = 0; sl < BA_GSLOTS; sl++) {
                    const int k = p.wblk[(size_t)w * BA_GSLOTS + sl];
                    if (k < 0 || !cb[k]) continue;
                    const bool dg = p.blk_ij[k].x == p.blk_ij[k].y;
                    const int st = (int)p.gcon.size(), padded = (cb[k] + 3) & ~3;
                    at[k] = st;
                    p.gseg.push_back(make_int4(k | sl << 16 | (dg ? 1 << 24 : 0), st, st + padded, 0));
                    p.gcon.resize((size_t)st + padded, (uint32_t)nobs | (uint32_t)(dg ? nl : nobs) << 16);   // the padding: all-zero dummy observation / landmark
                }
            }
            uint32_t* gc = p.gcon.data();
            for (int l = l0; l < l1; l++) {
                const int e0 = fo_ptr[l], e1 = fo_ptr[l + 1];
                const uint32_t ll = (uint32_t)(l - l0) << 16;
                for (int a = e0; a < e1; a++) {
                    const int fa = fof[a];
                    const uint32_t oa = (uint32_t)(fo_obs[a] - o0);
                    gc[at[ba[fa * Pf + fa]]++] = oa | ll;
                    for (int b = a + 1; b < e1; b++) {
                        const int fb = fof[b];
                        if (fb == fa) continue;
                        const uint32_t ob = (uint32_t)(fo_obs[b] - o0);
                        if (fa < fb) gc[at[ba[fa * Pf + fb]]++] = oa | ob << 16;
                        else gc[at[ba[fb * Pf + fa]]++] = ob | oa << 16;
                    }
                }
            }
        }
        p.gw_ptr[(size_t)n_chunks * nwarps] = (int)p.gseg.size();
        p.gcon.resize(p.gcon.size() + 32, 0u);   // a warp reads 32 entries at a segment's start whatever its length
    }
    return UCO_OK;
}

// ba.cu
cudaEvent_t* uco_ba_events(uco_b200_ctx* ctx);
int ba_streamed_solve(uco_b200_ctx* ctx, const uco_ba_problem* pb, const volatile unsigned char* stop, uco_ba_result* res);
// ba_cluster.cu
int ba_cluster_solve_batch(uco_b200_ctx* ctx, int n, const uco_ba_problem* const* pbs, const volatile unsigned char* stop, uco_ba_result* const* res);
