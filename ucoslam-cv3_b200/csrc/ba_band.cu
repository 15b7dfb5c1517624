// ba_band.cu — K13b: the reduced-camera solve of LARGE windows (global BA, BASELINE config 5) as a block-ENVELOPE Cholesky.
//
// Replaces (reference, relative to /root/reference):
//   3rdparty/g2o/g2o/solvers/eigen/linear_solver_eigen.h:92-123,145-199   LinearSolverEigen::solve: SimplicialLDLT of the reduced
//                                        camera system with a fill-reducing (AMD) block ordering computed once per structure
//   3rdparty/g2o/g2o/core/block_solver.hpp:315-448                        BlockSolver::solve around it
// The reduced system of a keyframe graph is block sparse (config 5: 498 free keyframes, 2.4 % of the 6x6 blocks non-zero, a band of
// 5 blocks); the dense potrf this file replaces spent 9 GFLOP per LM trial on zeros and was 82 % of the solve on every rank.
//
// Host (once per solve): reverse Cuthill-McKee ordering of the block graph, then the ROW ENVELOPE of the permuted pattern — row i
// keeps the blocks of columns f(i)..i; Cholesky fill stays inside it.  Device, per LM trial:
//   band_assemble : S = [i == j](Hpp + lambda I) - sum Schur (+ marker blocks) scattered into the envelope, right-hand side permuted
//   band_solve    : ONE thread block, right-looking block Cholesky over a sliding window of (bmax+1)^2 blocks held in shared memory
//                   (circular in rows and columns; in a global scratch when the envelope is too wide for shared memory), the
//                   right-hand side riding along (forward substitution for free), then the back substitution with the factor
//                   streamed back from L2 one step ahead.  Per block column: the 6x6 pivot is factored and inverted by one thread
//                   in registers (rsqrt, no shuffles: the dependent chain is the cost, not the flops), the column is scaled by the
//                   inverse (a product), the trailing window is updated by all threads.
// The factor is exact Cholesky in f64; g2o's SimplicialLDLT differs only in operation order (parity: poses 1e-7 as for the other
// BA forms, identical LM decisions on the test windows).
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <queue>
#include <vector>

struct uco_band_plan {
    int nb = 0, bmax = 0;
    size_t n_env = 0;                 // blocks in the envelope
    std::vector<int> perm, fcol, rowptr;   // old block -> new; new row -> first column; new row -> first envelope block
};

// reverse Cuthill-McKee over the block graph of the reduced system, then the row envelope
void uco_band_make_plan(int nb, int nblk, const int2* blk_ij, uco_band_plan& P) {
    P.nb = nb;
    std::vector<std::vector<int>> adj(nb);
    for (int b = 0; b < nblk; b++)
        if (blk_ij[b].x != blk_ij[b].y) { adj[blk_ij[b].x].push_back(blk_ij[b].y); adj[blk_ij[b].y].push_back(blk_ij[b].x); }
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::vector<int> order;
    order.reserve(nb);
    std::vector<char> seen(nb, 0);
    auto bfs_far = [&](int start, std::vector<int>* out) {   // BFS from start (neighbours by increasing degree); returns the last node
        std::vector<char> vis(nb, 0);
        std::queue<int> q;
        q.push(start); vis[start] = 1;
        int last = start;
        while (!q.empty()) {
            const int u = q.front(); q.pop();
            last = u;
            if (out) out->push_back(u);
            std::vector<int> nb_ = adj[u];
            std::sort(nb_.begin(), nb_.end(), [&](int a, int b) { return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b; });
            for (int v : nb_) if (!vis[v] && !seen[v]) { vis[v] = 1; q.push(v); }
        }
        return last;
    };
    for (int s0 = 0; s0 < nb; s0++) {          // one breadth-first numbering per connected component
        if (seen[s0]) continue;
        const int start = bfs_far(bfs_far(s0, nullptr), nullptr);   // pseudo-peripheral node: the far end of the far end
        std::vector<int> comp;
        bfs_far(start, &comp);
        for (int u : comp) seen[u] = 1;
        order.insert(order.end(), comp.begin(), comp.end());
    }
    std::reverse(order.begin(), order.end());
    P.perm.assign(nb, 0);
    for (int k = 0; k < nb; k++) P.perm[order[k]] = k;
    P.fcol.assign(nb, 0);
    for (int i = 0; i < nb; i++) P.fcol[i] = i;
    for (int b = 0; b < nblk; b++) {
        int i = P.perm[blk_ij[b].x], j = P.perm[blk_ij[b].y];
        if (i < j) std::swap(i, j);
        P.fcol[i] = std::min(P.fcol[i], j);
    }
    P.rowptr.assign(nb + 1, 0);
    P.bmax = 0;
    for (int i = 0; i < nb; i++) {
        P.rowptr[i + 1] = P.rowptr[i] + (i - P.fcol[i] + 1);
        P.bmax = std::max(P.bmax, i - P.fcol[i]);
    }
    P.n_env = (size_t)P.rowptr[nb];
}

namespace {

struct BandDev {
    int nb, W;                 // block unknowns; window size = bmax + 1
    const int *perm, *fcol, *rowptr;
    double* E;                 // envelope blocks (36 doubles each, row-major 6x6), becomes the factor (diagonal slots: inverse of L_kk)
    double* rhs;               // permuted right-hand side
    double* xp;                // solution in the CALLER's (unpermuted) order
    int* fail;
};

// S = [i == j](Hpp + lambda I) - (summed Schur blocks) [+ marker block], scattered into the (zeroed) envelope; rhs permuted
__global__ void __launch_bounds__(36) band_assemble_kernel(BandDev D, const int2* __restrict__ blk_ij, const double* __restrict__ Hpp, const double* lambda_p,
                                                           const double* __restrict__ Sp, const double* __restrict__ bp, const double* __restrict__ bsp,
                                                           const int* __restrict__ mk_blk_edge, const double* __restrict__ mk_e_blk) {
    const int blk = blockIdx.x, e = threadIdx.x, r = e / 6, c = e % 6;
    const int2 ij = blk_ij[blk];
    const bool diag = ij.x == ij.y;
    double h = 0;
    if (diag) {
        h = Hpp[36 * (size_t)ij.x + e];
        if (r == c) h += *lambda_p;
    }
    h -= Sp[36 * (size_t)blk + e];
    if (mk_blk_edge) {
        const int ed = mk_blk_edge[blk];
        if (ed >= 0) h += mk_e_blk[120 * (size_t)ed + 72 + e];
    }
    int pi = D.perm[ij.x], pj = D.perm[ij.y];
    int rr = r, cc = c;
    if (pi < pj) { const int t = pi; pi = pj; pj = t; rr = c; cc = r; }   // the block of the lower triangle is the transpose
    D.E[36 * (size_t)(D.rowptr[pi] + pj - D.fcol[pi]) + 6 * rr + cc] = h;
    if (diag && c == 0) D.rhs[6 * pi + r] = bp[6 * ij.x + r] - bsp[6 * ij.x + r];
}

constexpr int BAND_THREADS = 1024;

__global__ void __launch_bounds__(BAND_THREADS) band_solve_kernel(BandDev D, double* wglobal, int win_in_smem) {
    extern __shared__ double sm[];
    const int nb = D.nb, W = D.W, B = W - 1, tid = threadIdx.x;
    double* y = sm;                                   // 6 nb: right-hand side -> forward solution -> solution
    double* Linv = y + 6 * (size_t)nb;                // 36: inverse of the current pivot's Cholesky factor
    double* ynew = Linv + 36;                         // 6 (+2 pad)
    double* part = ynew + 8;                          // 6 W: partial sums of the back substitution
    double* win = win_in_smem ? part + 6 * (size_t)W : wglobal;   // W x W blocks, circular: block (i, j) at ((i % W) * W + (j % W)) * 36
    double* lcol = win + 36 * (size_t)W * W;          // W blocks: the scaled column of the current step
    __shared__ int fail;
    auto blk = [&](int i, int j) -> double* { return win + 36 * (size_t)((i % W) * W + (j % W)); };
    auto load_row = [&](int i) {                      // row i of the envelope into its window slot (zero left of f(i))
        const int f = D.fcol[i];
        const double* src = D.E + 36 * (size_t)D.rowptr[i];
        for (int t = tid; t < 36 * W; t += BAND_THREADS) {
            const int j = i - B + t / 36;
            if (j >= 0) blk(i, j)[t % 36] = j >= f ? src[36 * (size_t)(j - f) + t % 36] : 0.0;
        }
    };
    if (tid == 0) fail = 0;
    for (int t = tid; t < 6 * nb; t += BAND_THREADS) y[t] = D.rhs[t];
    for (int i = 0; i < nb && i < W; i++) load_row(i);
    __syncthreads();

    for (int k = 0; k < nb; k++) {
        const int nrow = min(B, nb - 1 - k);          // window rows below the pivot
        // P1: Cholesky of the pivot block and the inverse of its factor: one thread, registers, no shuffles (the dependent chain of
        // six rsqrt's is the cost, not the 100 flops)
        if (tid == 0) {
            const double* A = blk(k, k);
            double L[6][6], I[6][6], s[6];
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 6; j++) {
                double d = A[6 * j + j];
#pragma unroll
                for (int q = 0; q < 6; q++) if (q < j) d -= L[j][q] * L[j][q];
                ok = ok && d > 0 && isfinite(d);
                s[j] = rsqrt(d);
                L[j][j] = d * s[j];
#pragma unroll
                for (int i = 0; i < 6; i++)
                    if (i > j) {
                        double v = A[6 * i + j];
#pragma unroll
                        for (int q = 0; q < 6; q++) if (q < j) v -= L[i][q] * L[j][q];
                        L[i][j] = v * s[j];
                    }
            }
#pragma unroll
            for (int c = 0; c < 6; c++)                // columns of the inverse by forward substitution
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    if (r < c) I[r][c] = 0;
                    else {
                        double v = r == c ? 1.0 : 0.0;
#pragma unroll
                        for (int q = 0; q < 6; q++) if (q >= c && q < r) v -= L[r][q] * I[q][c];
                        I[r][c] = v * s[r];
                    }
                }
#pragma unroll
            for (int e = 0; e < 36; e++) Linv[e] = I[e / 6][e % 6];
            if (!ok) fail = 1;
        }
        __syncthreads();
        if (fail) break;
        // P2: L_ik = A_ik Linv^T into `lcol` and the factor; y_k = Linv b_k; the pivot's inverse goes to the factor's diagonal slot
        for (int t = tid; t < 36 * nrow; t += BAND_THREADS) {
            const int i = k + 1 + t / 36, r = (t % 36) / 6, c = t % 6;
            const double* A = blk(i, k);
            double v = 0;
#pragma unroll
            for (int q = 0; q < 6; q++) if (q <= c) v += A[6 * r + q] * Linv[6 * c + q];
            lcol[t] = v;
            if (k >= D.fcol[i]) D.E[36 * (size_t)(D.rowptr[i] + k - D.fcol[i]) + t % 36] = v;
        }
        if (tid >= BAND_THREADS - 6) {
            const int r = tid - (BAND_THREADS - 6);
            double v = 0;
#pragma unroll
            for (int q = 0; q < 6; q++) if (q <= r) v += Linv[6 * r + q] * y[6 * k + q];
            ynew[r] = v;
        } else if (tid >= BAND_THREADS - 64 && tid < BAND_THREADS - 64 + 36) {
            const int e = tid - (BAND_THREADS - 64);
            D.E[36 * (size_t)(D.rowptr[k] + k - D.fcol[k]) + e] = Linv[e];
        }
        __syncthreads();
        // P3: trailing update of the window, right-hand side, and the row that enters the window
        for (int t = tid; t < 36 * nrow * nrow; t += BAND_THREADS) {
            const int a = t / (36 * nrow), b = (t / 36) % nrow;
            if (b > a) continue;
            const int r = (t % 36) / 6, c = t % 6;
            const double *La = lcol + 36 * a + 6 * r, *Lb = lcol + 36 * b + 6 * c;
            double v = 0;
#pragma unroll
            for (int q = 0; q < 6; q++) v += La[q] * Lb[q];
            blk(k + 1 + a, k + 1 + b)[6 * r + c] -= v;
        }
        for (int t = tid; t < 6 * nrow; t += BAND_THREADS) {
            const double* La = lcol + 6 * t;
            double v = 0;
#pragma unroll
            for (int q = 0; q < 6; q++) v += La[q] * ynew[q];
            y[6 * (k + 1) + t] -= v;
        }
        if (tid < 6) y[6 * k + tid] = ynew[tid];
        if (k + W < nb) load_row(k + W);
        __syncthreads();
    }
    __syncthreads();
    if (fail) {
        if (tid == 0) *D.fail = 1;
        for (int t = tid; t < 6 * nb; t += BAND_THREADS) D.xp[t] = 0;
        return;
    }
    if (tid == 0) *D.fail = 0;
    // back substitution L^T x = y: x_k = Linv_k^T (y_k - sum_{i > k} L_ik^T x_i); thread (ii, c) owns column c of block (k+1+ii, k); the
    // factor is streamed from L2 one step ahead (registers), so its latency is off the dependent chain
    const int ii = tid / 6, c = tid % 6;
    if (6 * B > BAND_THREADS) {   // very wide envelope (close to dense): plain multi-pass form, no prefetch
        for (int k = nb - 1; k >= 0; k--) {
            for (int t = tid; t < 6 * B; t += BAND_THREADS) {
                const int i = k + 1 + t / 6, cc = t % 6;
                double v = 0;
                if (i < nb && k >= D.fcol[i]) {
                    const double* src = D.E + 36 * (size_t)(D.rowptr[i] + k - D.fcol[i]);
                    for (int r = 0; r < 6; r++) v += src[6 * r + cc] * y[6 * i + r];
                }
                part[t] = v;
            }
            __syncthreads();
            if (tid < 6) {
                double sv = y[6 * k + tid];
                for (int a = 0; a < B; a++) sv -= part[6 * a + tid];
                ynew[tid] = sv;
            }
            __syncthreads();
            if (tid < 6) {
                const double* dg = D.E + 36 * (size_t)(D.rowptr[k] + k - D.fcol[k]);
                double v = 0;
                for (int r = tid; r < 6; r++) v += dg[6 * r + tid] * ynew[r];
                y[6 * k + tid] = v;
            }
            __syncthreads();
        }
        for (int t = tid; t < nb; t += BAND_THREADS) {
            const int p = D.perm[t];
            for (int r = 0; r < 6; r++) D.xp[6 * t + r] = y[6 * p + r];
        }
        return;
    }
    const bool worker = tid < 6 * B;
    double pre[6], dinv[6];
    auto fetch = [&](int k) {
        const int i = k + 1 + ii;
        const bool live = worker && i < nb && k >= D.fcol[i];
        const double* src = live ? D.E + 36 * (size_t)(D.rowptr[i] + k - D.fcol[i]) : nullptr;
#pragma unroll
        for (int r = 0; r < 6; r++) pre[r] = live ? src[6 * r + c] : 0.0;
        if (tid < 6) {                               // column c of Linv_k = row c of Linv_k^T
            const double* dg = D.E + 36 * (size_t)(D.rowptr[k] + k - D.fcol[k]);
#pragma unroll
            for (int r = 0; r < 6; r++) dinv[r] = dg[6 * r + c];
        }
    };
    if (nb > 0) fetch(nb - 1);
    for (int k = nb - 1; k >= 0; k--) {
        double cur[6], dcur[6];
#pragma unroll
        for (int r = 0; r < 6; r++) { cur[r] = pre[r]; dcur[r] = dinv[r]; }
        if (k > 0) fetch(k - 1);
        if (worker) {
            const int i = k + 1 + ii;
            double v = 0;
            if (i < nb) {
#pragma unroll
                for (int r = 0; r < 6; r++) v += cur[r] * y[6 * i + r];
            }
            part[tid] = v;
        }
        __syncthreads();
        if (tid < 6) {
            double sv = y[6 * k + tid];
            for (int a = 0; a < B; a++) sv -= part[6 * a + tid];
            ynew[tid] = sv;
        }
        __syncthreads();
        if (tid < 6) {
            double v = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) if (r >= c) v += dcur[r] * ynew[r];
            y[6 * k + tid] = v;
        }
        __syncthreads();
    }
    for (int t = tid; t < nb; t += BAND_THREADS) {
        const int p = D.perm[t];
#pragma unroll
        for (int r = 0; r < 6; r++) D.xp[6 * t + r] = y[6 * p + r];
    }
}

size_t band_smem_fixed(int nb, int W) { return 8 * (6 * (size_t)nb + 36 + 8 + 6 * (size_t)W); }
size_t band_smem_window(int W) { return 8 * 36 * ((size_t)W * W + W); }

}  // namespace

// ---- launchers used by ba.cu -------------------------------------------------------------------------------------------------------
// device-side plan arrays + buffers live in the caller's arena: perm | fcol | rowptr (ints), E (36 n_env doubles), rhs (6 nb), wglobal
int uco_band_assemble_launch(uco_b200_ctx* ctx, int nb, int W, const int* perm_dev, const int* fcol_dev, const int* rowptr_dev, double* E, size_t n_env,
                             double* rhs, int nblk, const int2* blk_ij_dev, const double* Hpp, const double* lambda_dev, const double* Sp,
                             const double* bp, const double* bsp, const int* mk_blk_edge, const double* mk_e_blk) {
    BandDev D{nb, W, perm_dev, fcol_dev, rowptr_dev, E, rhs, nullptr, nullptr};
    UCO_CUDA(ctx, cudaMemsetAsync(E, 0, 8 * 36 * n_env, ctx->stream));
    band_assemble_kernel<<<nblk, 36, 0, ctx->stream>>>(D, blk_ij_dev, Hpp, lambda_dev, Sp, bp, bsp, mk_blk_edge, mk_e_blk);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

// bytes of global scratch the solve needs when the window does not fit shared memory (0 otherwise)
size_t uco_band_scratch_bytes(int nb, int W, int smem_optin) {
    return band_smem_fixed(nb, W) + band_smem_window(W) <= (size_t)smem_optin ? 0 : band_smem_window(W);
}

int uco_band_solve_launch(uco_b200_ctx* ctx, int nb, int W, const int* perm_dev, const int* fcol_dev, const int* rowptr_dev, double* E, double* rhs,
                          double* xp, int* fail_dev, double* wglobal, int smem_optin) {
    BandDev D{nb, W, perm_dev, fcol_dev, rowptr_dev, E, rhs, xp, fail_dev};
    const size_t fixed = band_smem_fixed(nb, W), winb = band_smem_window(W);
    if (fixed > (size_t)smem_optin) return uco_fail(ctx, UCO_E_CAPACITY, "band solve: %d block unknowns exceed the shared-memory right-hand side", nb);
    const bool in_smem = fixed + winb <= (size_t)smem_optin;
    if (!in_smem && !wglobal) return uco_fail(ctx, UCO_E_INVALID, "band solve: no scratch for the window");
    const size_t smem = in_smem ? fixed + winb : fixed;
    static size_t configured = 0;
    if (smem > configured) {
        UCO_CUDA(ctx, cudaFuncSetAttribute(band_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    band_solve_kernel<<<1, BAND_THREADS, smem, ctx->stream>>>(D, wglobal, in_smem ? 1 : 0);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}
