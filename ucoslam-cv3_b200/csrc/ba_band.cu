// ba_band.cu — K13b: the reduced-camera solve of LARGE windows (global BA, BASELINE config 5) as a two-level block-envelope
// Cholesky: a one-level nested dissection of the keyframe graph whose pieces are factored by independent thread blocks.
//
// Replaces (reference, relative to /root/reference):
//   3rdparty/g2o/g2o/solvers/eigen/linear_solver_eigen.h:92-123,145-199   LinearSolverEigen::solve: SimplicialLDLT of the reduced
//                                        camera system with a fill-reducing (AMD) block ordering computed once per structure
//   3rdparty/g2o/g2o/core/linear_solver.h:44-90                           the LinearSolver<MatrixType>::solve seam (uco_b200_block_solve)
//   3rdparty/g2o/g2o/core/block_solver.hpp:315-448                        BlockSolver::solve around it
// The reduced system of a keyframe graph is block sparse (config 5: 498 free keyframes, 2.4 % of the 6x6 blocks non-zero).  A band
// Cholesky of it is a chain of nb dependent pivots; what costs is the length of that chain, not the flops.  So, once per structure
// (host):
//   1. breadth-first level structure of the block graph from a pseudo-peripheral keyframe;
//   2. K of its levels become SEPARATORS (K chosen by a cost model: longest interior chain + separator chain); the connected
//      components of what is left are the INTERIOR FRONTS (a trajectory with a loop closure: 2 arcs per gap between separator levels);
//   3. every front is ordered by reverse Cuthill-McKee and stored as a row envelope (fill stays inside it) plus its BORDER: the dense
//      block rows of the separator keyframes it touches; the separators themselves form the ROOT front, ordered and stored the same
//      way over the graph "separator edges + one clique per front border".
// Device, per LM trial (4 launches, no library call, no atomics, fixed summation order):
//   band_assemble : S = [i == j](Hpp + lambda I) - sum Schur (+ marker blocks) scattered to its place (front envelope / border / root)
//   band_front    : one thread block per interior front: right-looking block Cholesky over a sliding window of (bmax+1)^2 blocks in
//                   shared memory (circular in rows and columns), the border rows and the right-hand side riding along; leaves the
//                   factor, the border's Schur complement and reduced right-hand side
//   band_front (root): gathers the fronts' Schur complements in a fixed order, factors the separator system, forward + back substitution
//   band_back     : one thread block per interior front: back substitution given the separator solution
// Per block column: the 6x6 pivot is factored by one thread in registers (the dependent chain of six rsqrt's is the cost), the column
// (window rows + border rows) is scaled by forward substitution, the trailing window / border / Schur blocks are updated by all
// threads, the row entering the window arrives by cp.async.  K = 0 (graph without a level structure worth cutting) degenerates to
// the plain envelope Cholesky in one thread block.
// The factor is exact Cholesky in f64; g2o's SimplicialLDLT differs only in operation order (parity: poses 1e-7 as for the other
// BA forms, identical LM decisions on the test windows).
#include "ba_band.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>
#include <vector>

namespace {

constexpr int BAND_THREADS = 512;

size_t front_smem_fixed(int n, int W, int nbr) {   // doubles: y | bv | Lkk + sdiag + ynew (pad to 56) | lcol ; ints: fcol, rowptr, pair table
    const int B = W - 1;
    const size_t npairs = (size_t)(B + nbr) * (B + nbr + 1) / 2;
    return 8 * (6 * (size_t)n + 6 * (size_t)nbr + 56 + 36 * (size_t)(B + nbr) + 36 * (size_t)W * nbr + 36 * (size_t)nbr * nbr) +
           4 * (2 * (size_t)n + npairs + 4);
}
size_t front_smem_window(int W) { return 8 * 36 * (size_t)std::max(W * W, 5 * W); }   // the back substitution's row ring reuses lcol + window: (4 + 1) W blocks

using Adj = std::vector<std::vector<int>>;

// breadth-first search inside `mask` (mask[v] == tag) from `start`, neighbours by increasing degree; fills order / level
int bfs(const Adj& adj, const std::vector<int>& mask, int tag, int start, std::vector<int>* order, std::vector<int>* level) {
    std::vector<int> lev(adj.size(), -1);
    std::queue<int> q;
    q.push(start);
    lev[start] = 0;
    int last = start;
    std::vector<int> nbs;
    while (!q.empty()) {
        const int u = q.front();
        q.pop();
        last = u;
        if (order) order->push_back(u);
        nbs.clear();
        for (int v : adj[u])
            if (mask[v] == tag && lev[v] < 0) nbs.push_back(v);
        std::sort(nbs.begin(), nbs.end(), [&](int a, int b) { return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b; });
        for (int v : nbs) {
            lev[v] = lev[u] + 1;
            q.push(v);
        }
    }
    if (level) *level = lev;
    // the deepest level's node of the smallest degree: the end of the graph a Cuthill-McKee numbering should start from
    int best = last;
    for (size_t v = 0; v < adj.size(); v++)
        if (lev[v] == lev[last] && (adj[v].size() < adj[best].size() || (adj[v].size() == adj[best].size() && (int)v < best))) best = (int)v;
    return best;
}

// reverse Cuthill-McKee of the connected piece of `mask == tag` that holds `seed` (appended to `out`)
void rcm_component(const Adj& adj, const std::vector<int>& mask, int tag, int seed, std::vector<int>& out) {
    const int start = bfs(adj, mask, tag, bfs(adj, mask, tag, seed, nullptr, nullptr), nullptr, nullptr);   // pseudo-peripheral: far end of the far end
    std::vector<int> comp;
    bfs(adj, mask, tag, start, &comp, nullptr);
    std::reverse(comp.begin(), comp.end());
    out.insert(out.end(), comp.begin(), comp.end());
}

struct Piece {   // a set of block unknowns factored by one thread block
    std::vector<int> nodes;   // in elimination order
    std::vector<int> border;  // separator nodes (caller's indices), later sorted by root-local row
};

}  // namespace

// ---- host: ordering, fronts, storage map ------------------------------------------------------------------------------------------
void uco_band_make_plan(int nb, int nblk, const int2* blk_ij, int smem_optin, int force_k, uco_band_plan& P) {
    P = uco_band_plan();
    P.nb = nb;
    P.nblk = nblk;
    Adj adj(nb);
    for (int b = 0; b < nblk; b++)
        if (blk_ij[b].x != blk_ij[b].y) {
            adj[blk_ij[b].x].push_back(blk_ij[b].y);
            adj[blk_ij[b].y].push_back(blk_ij[b].x);
        }
    for (auto& a : adj) {
        std::sort(a.begin(), a.end());
        a.erase(std::unique(a.begin(), a.end()), a.end());
    }
    // connected components of the whole graph; the largest one is the candidate for cutting
    std::vector<int> comp_of(nb, -1), all(nb, 0);
    std::vector<std::vector<int>> comps;
    for (int s = 0; s < nb; s++) {
        if (comp_of[s] >= 0) continue;
        std::vector<int> c;
        bfs(adj, all, 0, s, &c, nullptr);
        // bfs does not know about comp_of: mark afterwards (components are disjoint by construction)
        for (int v : c) comp_of[v] = (int)comps.size();
        comps.push_back(c);
    }
    int big = 0;
    for (size_t c = 1; c < comps.size(); c++)
        if (comps[c].size() > comps[big].size()) big = (int)c;
    // level structure of the largest component
    std::vector<int> level;
    int L = 0;
    std::vector<int> mask(nb, 0);   // 0 = interior, 1 = separator
    if (nb > 0) {
        std::vector<int> cm(nb, 1);
        for (int v : comps[big]) cm[v] = 0;
        const int start = bfs(adj, cm, 0, bfs(adj, cm, 0, comps[big][0], nullptr, nullptr), nullptr, nullptr);
        bfs(adj, cm, 0, start, nullptr, &level);
        for (int v : comps[big]) L = std::max(L, level[v] + 1);
    }
    auto cut = [&](int K, int style, std::vector<int>& m) {   // K separator levels; style 0: evenly spaced, 1: end pieces half as long (the two ends of a
        m.assign(nb, 0);                                        // closed loop's level structure hold both sides of the loop)
        if (K <= 0) return;
        std::vector<char> is_sep(L, 0);
        for (int q = 1; q <= K; q++) {
            const double at = style == 0 ? (double)q * L / (K + 1) : ((double)q - 0.5) * L / K;
            is_sep[std::min(L - 2, std::max(1, (int)std::lround(at)))] = 1;
        }
        for (int v : comps[big])
            if (is_sep[level[v]]) m[v] = 1;
    };
    auto pieces_of = [&](const std::vector<int>& m, std::vector<Piece>& out) {
        out.clear();
        std::vector<char> seen(nb, 0);
        for (int s = 0; s < nb; s++) {
            if (m[s] != 0 || seen[s]) continue;
            Piece pc;
            rcm_component(adj, m, 0, s, pc.nodes);
            for (int v : pc.nodes) {
                seen[v] = 1;
                for (int u : adj[v])
                    if (m[u] == 1) pc.border.push_back(u);
            }
            std::sort(pc.border.begin(), pc.border.end());
            pc.border.erase(std::unique(pc.border.begin(), pc.border.end()), pc.border.end());
            out.push_back(std::move(pc));
        }
    };
    auto local_W = [&](const std::vector<int>& nodes) {
        std::vector<int> pos(nb, -1);
        for (size_t k = 0; k < nodes.size(); k++) pos[nodes[k]] = (int)k;
        int bmax = 0;
        for (size_t k = 0; k < nodes.size(); k++)
            for (int u : adj[nodes[k]])
                if (pos[u] >= 0) bmax = std::max(bmax, std::abs(pos[u] - (int)k));
        return std::max(2, bmax + 1);
    };
    // cost of a cut: the chain of dependent pivots (longest interior front, then the separators) — and everything has to fit shared memory
    int bestK = 0, best_style = 0;
    {
        struct Cand { double cost; int K, style; };
        std::vector<Cand> cands;
        std::vector<int> m;
        std::vector<Piece> pcs;
        const int Kmax = std::min(48, (L - 1) / 3);
        for (int K = 0; K <= Kmax; K++)
            for (int style = 0; style < (K == 0 ? 1 : 2); style++) {
                if (force_k < 0 && K > 8 && (K & (K > 16 ? 3 : 1))) continue;   // beyond 8 levels the cost curve is flat: every 2nd, then every 4th
                if (force_k >= 0 && (K != std::min(force_k % 100, Kmax) || (force_k >= 100 && style != 1))) continue;
                cut(K, style, m);
                // sizes of the pieces only (one sweep); the orderings are built for the winner
                size_t longest = 0, nsep = 0, npieces = 0;
                std::vector<char> seen(nb, 0);
                std::vector<int> stack;
                for (int s0 = 0; s0 < nb; s0++) {
                    if (m[s0] == 1) { nsep++; continue; }
                    if (seen[s0]) continue;
                    size_t sz = 0;
                    stack.assign(1, s0);
                    seen[s0] = 1;
                    while (!stack.empty()) {
                        const int u = stack.back();
                        stack.pop_back();
                        sz++;
                        for (int v : adj[u])
                            if (m[v] == 0 && !seen[v]) { seen[v] = 1; stack.push_back(v); }
                    }
                    longest = std::max(longest, sz);
                    npieces++;
                }
                if (npieces > 256) continue;
                // forward + back steps; the root's band is about twice as wide
                cands.push_back({K == 0 ? 1.5 * (double)longest : 1.5 * (double)longest + 2.5 * (double)nsep + 8.0, K, style});
            }
        std::stable_sort(cands.begin(), cands.end(), [](const Cand& x, const Cand& y) { return x.cost < y.cost; });
        for (const Cand& cd : cands) {   // the cheapest cut whose fronts fit shared memory
            bool fits = true;
            if (cd.K > 0) {
                cut(cd.K, cd.style, m);
                pieces_of(m, pcs);
                for (const Piece& pc : pcs) {
                    const int w = local_W(pc.nodes);
                    if (front_smem_fixed((int)pc.nodes.size(), w, (int)pc.border.size()) + front_smem_window(w) > (size_t)smem_optin) fits = false;
                }
            }
            if (fits) {
                bestK = cd.K;
                best_style = cd.style;
                break;
            }
        }
    }
    if (force_k >= 0 && bestK != std::min(force_k % 100, std::min(48, (L - 1) / 3))) {   // the forced cut does not fit shared memory: choose freely
        uco_band_make_plan(nb, nblk, blk_ij, smem_optin, -1, P);
        return;
    }
    P.K = bestK;
    cut(bestK, best_style, mask);
    std::vector<Piece> pcs;
    pieces_of(mask, pcs);
    // the root: the separators, ordered by RCM over "separator edges + one clique per front border"; without separators the largest piece is the root
    Piece root;
    std::vector<int> root_pos(nb, -1);
    if (bestK > 0) {
        Adj radj(nb);
        for (int v = 0; v < nb; v++)
            if (mask[v] == 1)
                for (int u : adj[v])
                    if (mask[u] == 1) radj[v].push_back(u);
        for (const Piece& pc : pcs)
            for (int a : pc.border)
                for (int b : pc.border)
                    if (a != b) radj[a].push_back(b);
        for (auto& a : radj) {
            std::sort(a.begin(), a.end());
            a.erase(std::unique(a.begin(), a.end()), a.end());
        }
        std::vector<char> seen(nb, 0);
        for (int s = 0; s < nb; s++) {
            if (mask[s] != 1 || seen[s]) continue;
            const size_t at = root.nodes.size();
            rcm_component(radj, mask, 1, s, root.nodes);
            for (size_t k = at; k < root.nodes.size(); k++) seen[root.nodes[k]] = 1;
        }
    } else if (!pcs.empty()) {
        size_t bi = 0;
        for (size_t c = 1; c < pcs.size(); c++)
            if (pcs[c].nodes.size() > pcs[bi].nodes.size()) bi = c;
        root = std::move(pcs[bi]);
        pcs.erase(pcs.begin() + bi);
    }
    for (size_t k = 0; k < root.nodes.size(); k++) root_pos[root.nodes[k]] = (int)k;
    for (Piece& pc : pcs) std::sort(pc.border.begin(), pc.border.end(), [&](int a, int b) { return root_pos[a] < root_pos[b]; });
    pcs.push_back(std::move(root));
    const int nf = (int)pcs.size();
    // per-front storage
    std::vector<int> front_of(nb, -1), local_of(nb, -1);
    P.fronts.resize(nf);
    long long z = 0;
    for (int f = 0; f < nf; f++) {
        const Piece& pc = pcs[f];
        uco_band_front& F = P.fronts[f];
        F.n = (int)pc.nodes.size();
        F.nbr = (int)pc.border.size();
        F.row0 = (int)P.row_node.size();
        F.bord0 = (int)P.bord.size();
        for (int k = 0; k < F.n; k++) {
            front_of[pc.nodes[k]] = f;
            local_of[pc.nodes[k]] = k;
        }
        for (int b : pc.border) P.bord.push_back(root_pos[b]);
        for (int k = 0; k < F.n; k++) {
            P.row_node.push_back(pc.nodes[k]);
            P.row_fcol.push_back(k);
            P.row_ptr.push_back(0);
        }
    }
    // envelopes: original edges inside a front; for the root also the border cliques
    auto touch = [&](int f, int li, int lj) {
        if (li < lj) std::swap(li, lj);
        int& fc = P.row_fcol[P.fronts[f].row0 + li];
        fc = std::min(fc, lj);
    };
    for (int b = 0; b < nblk; b++) {
        const int i = blk_ij[b].x, j = blk_ij[b].y;
        if (front_of[i] == front_of[j]) touch(front_of[i], local_of[i], local_of[j]);
    }
    for (int f = 0; f + 1 < nf; f++) {
        const uco_band_front& F = P.fronts[f];
        for (int a = 0; a < F.nbr; a++)
            for (int b = 0; b <= a; b++) touch(nf - 1, P.bord[F.bord0 + a], P.bord[F.bord0 + b]);
    }
    for (int f = 0; f < nf; f++) {
        uco_band_front& F = P.fronts[f];
        int bmax = 0, ptr = 0;
        for (int k = 0; k < F.n; k++) {
            P.row_ptr[F.row0 + k] = ptr;
            ptr += k - P.row_fcol[F.row0 + k] + 1;
            bmax = std::max(bmax, k - P.row_fcol[F.row0 + k]);
        }
        F.W = std::max(2, bmax + 1);   // >= 2: the look-ahead factors pivot k + 1 one column early, its row must be in the window by then
        F.npairs = (F.W - 1 + F.nbr) * (F.W + F.nbr) / 2;
        F.oE = z; z += 36LL * ptr;
        F.oB = z; z += 36LL * F.n * F.nbr;
        F.oS = z; z += 36LL * F.nbr * F.nbr;
        F.oY = z; z += 6LL * F.n;
        F.oR = z; z += 6LL * F.nbr;
        if (f == nf - 1 && front_smem_fixed(F.n, F.W, 0) + front_smem_window(F.W) > (size_t)smem_optin) P.scratch_doubles = 36 * (size_t)F.W * F.W;
    }
    P.z_doubles = z;
    // where every caller block lands
    P.blk_dst.assign(nblk, -1);
    for (int b = 0; b < nblk; b++) {
        const int i = blk_ij[b].x, j = blk_ij[b].y;   // the block holds S(i, j): rows = unknowns of i, columns = unknowns of j
        const int fi = front_of[i], fj = front_of[j];
        if (fi == fj) {
            const uco_band_front& F = P.fronts[fi];
            const int li = local_of[i], lj = local_of[j];
            const int hi = std::max(li, lj), lo = std::min(li, lj);
            const long long off = F.oE + 36LL * (P.row_ptr[F.row0 + hi] + lo - P.row_fcol[F.row0 + hi]);
            P.blk_dst[b] = off << 1 | (li < lj ? 1 : 0);   // stored block = (row hi, column lo)
        } else {
            // one interior, one separator: border block (row = border slot, column = interior row k) of the interior's front
            const bool i_int = fi != nf - 1;
            const int fint = i_int ? fi : fj, inode = i_int ? i : j, snode = i_int ? j : i;
            const uco_band_front& F = P.fronts[fint];
            int slot = -1;
            for (int a = 0; a < F.nbr; a++)
                if (P.bord[F.bord0 + a] == root_pos[snode]) slot = a;
            const long long off = F.oB + 36LL * ((long long)local_of[inode] * F.nbr + slot);
            P.blk_dst[b] = slot < 0 || (fi != nf - 1 && fj != nf - 1) ? -1 : (off << 1 | (i_int ? 1 : 0));   // S(sep, int) as is, S(int, sep) transposed
        }
    }
    P.rhs_dst.assign(nb, 0);
    for (int v = 0; v < nb; v++) P.rhs_dst[v] = P.fronts[front_of[v]].oY + 6LL * local_of[v];
    // root gathers, destination-major, sources in front order
    {
        const uco_band_front& R = P.fronts[nf - 1];
        struct Src { long long dst, src; };
        std::vector<Src> s;
        std::vector<std::vector<long long>> rs(R.n);
        for (int f = 0; f + 1 < nf; f++) {
            const uco_band_front& F = P.fronts[f];
            for (int a = 0; a < F.nbr; a++) {
                const int ra = P.bord[F.bord0 + a];   // borders are sorted by root row: ra > rb for a > b
                rs[ra].push_back(F.oR + 6LL * a);
                for (int b = 0; b <= a; b++) {
                    const int rb = P.bord[F.bord0 + b];
                    const long long dst = R.oE + 36LL * (P.row_ptr[R.row0 + ra] + rb - P.row_fcol[R.row0 + ra]);
                    s.push_back({dst, (F.oS + 36LL * ((long long)a * F.nbr + b)) << 1});
                }
            }
        }
        std::stable_sort(s.begin(), s.end(), [](const Src& x, const Src& y) { return x.dst < y.dst; });
        for (size_t k = 0; k < s.size(); k++) {
            if (k == 0 || s[k].dst != s[k - 1].dst) {
                P.g_dst.push_back(s[k].dst);
                P.g_ptr.push_back((int)k);
            }
            P.g_src.push_back(s[k].src);
        }
        P.g_ptr.push_back((int)s.size());
        P.r_ptr.assign(R.n + 1, 0);
        for (int k = 0; k < R.n; k++) {
            P.r_ptr[k + 1] = P.r_ptr[k] + (int)rs[k].size();
            P.r_src.insert(P.r_src.end(), rs[k].begin(), rs[k].end());
        }
    }
    // device blob
    auto put = [&](const void* p, size_t bytes) {
        const size_t at = (P.blob.size() + 15) / 16 * 16;
        P.blob.resize(at + bytes);
        if (bytes) memcpy(P.blob.data() + at, p, bytes);
        return at;
    };
    P.o_fronts = put(P.fronts.data(), sizeof(uco_band_front) * P.fronts.size());
    P.o_fcol = put(P.row_fcol.data(), 4 * P.row_fcol.size());
    P.o_rowptr = put(P.row_ptr.data(), 4 * P.row_ptr.size());
    P.o_node = put(P.row_node.data(), 4 * P.row_node.size());
    P.o_bord = put(P.bord.data(), 4 * P.bord.size());
    P.o_blk_dst = put(P.blk_dst.data(), 8 * P.blk_dst.size());
    P.o_rhs_dst = put(P.rhs_dst.data(), 8 * P.rhs_dst.size());
    P.o_g_dst = put(P.g_dst.data(), 8 * P.g_dst.size());
    P.o_g_src = put(P.g_src.data(), 8 * P.g_src.size());
    P.o_g_ptr = put(P.g_ptr.data(), 4 * P.g_ptr.size());
    P.o_r_ptr = put(P.r_ptr.data(), 4 * P.r_ptr.size());
    P.o_r_src = put(P.r_src.data(), 8 * P.r_src.size());
    P.blob.resize((P.blob.size() + 15) / 16 * 16);
}

bool uco_band_plan_valid(const uco_band_plan& P) {
    for (long long d : P.blk_dst)
        if (d < 0) return false;
    return true;
}

// ---- host: the same algebra on the same storage, without windows (inspection hook of the planner; never on the product path) -------
namespace {
// in-place Cholesky of the 6x6 block A (row-major, lower part used); returns false on a non-positive pivot; L in the lower part
bool host_chol6(double* A) {
    for (int j = 0; j < 6; j++) {
        double d = A[6 * j + j];
        for (int q = 0; q < j; q++) d -= A[6 * j + q] * A[6 * j + q];
        if (!(d > 0) || !std::isfinite(d)) return false;
        const double s = 1.0 / std::sqrt(d);
        A[6 * j + j] = d * s;
        for (int i = j + 1; i < 6; i++) {
            double v = A[6 * i + j];
            for (int q = 0; q < j; q++) v -= A[6 * i + q] * A[6 * j + q];
            A[6 * i + j] = v * s;
        }
    }
    return true;
}
void host_trsm_right(double* X, const double* L) {   // X <- X L^-T (rows of X solved against L)
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
            double v = X[6 * r + c];
            for (int q = 0; q < c; q++) v -= X[6 * r + q] * L[6 * c + q];
            X[6 * r + c] = v / L[6 * c + c];
        }
}
void host_sub_abt(double* C, const double* A, const double* B, double sign) {   // C += sign * A B^T
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
            double v = 0;
            for (int q = 0; q < 6; q++) v += A[6 * r + q] * B[6 * c + q];
            C[6 * r + c] += sign * v;
        }
}
bool host_front(const uco_band_plan& P, int f, double* Z) {
    const uco_band_front& F = P.fronts[f];
    const int* fcol = P.row_fcol.data() + F.row0;
    const int* rowptr = P.row_ptr.data() + F.row0;
    auto E = [&](int i, int j) { return Z + F.oE + 36LL * (rowptr[i] + j - fcol[i]); };
    auto Bd = [&](int k, int b) { return Z + F.oB + 36LL * ((long long)k * F.nbr + b); };
    double* y = Z + F.oY;
    for (int k = 0; k < F.n; k++) {
        double* Lkk = E(k, k);
        if (!host_chol6(Lkk)) return false;
        for (int c = 1; c < 6; c++)
            for (int r = 0; r < c; r++) Lkk[6 * r + c] = 0;
        for (int r = 0; r < 6; r++) {   // y_k <- Lkk^-1 y_k
            double v = y[6 * k + r];
            for (int q = 0; q < r; q++) v -= Lkk[6 * r + q] * y[6 * k + q];
            y[6 * k + r] = v / Lkk[6 * r + r];
        }
        std::vector<int> rows;
        for (int i = k + 1; i < F.n && i < k + F.W; i++)
            if (fcol[i] <= k) rows.push_back(i);
        for (int i : rows) host_trsm_right(E(i, k), Lkk);
        for (int b = 0; b < F.nbr; b++) host_trsm_right(Bd(k, b), Lkk);
        for (size_t a = 0; a < rows.size(); a++) {
            for (size_t b = 0; b <= a; b++) host_sub_abt(E(rows[a], rows[b]), E(rows[a], k), E(rows[b], k), -1.0);
            for (int r = 0; r < 6; r++)
                for (int q = 0; q < 6; q++) y[6 * rows[a] + r] -= E(rows[a], k)[6 * r + q] * y[6 * k + q];
        }
        for (int a = 0; a < F.nbr; a++) {
            for (int i : rows) host_sub_abt(Bd(i, a), Bd(k, a), E(i, k), -1.0);
            for (int b = 0; b <= a; b++) host_sub_abt(Z + F.oS + 36LL * ((long long)a * F.nbr + b), Bd(k, a), Bd(k, b), 1.0);
            for (int r = 0; r < 6; r++)
                for (int q = 0; q < 6; q++) Z[F.oR + 6 * a + r] += Bd(k, a)[6 * r + q] * y[6 * k + q];
        }
    }
    return true;
}
void host_back(const uco_band_plan& P, int f, double* Z, double* x) {
    const uco_band_front& F = P.fronts[f];
    const uco_band_front& R = P.fronts.back();
    const int* fcol = P.row_fcol.data() + F.row0;
    const int* rowptr = P.row_ptr.data() + F.row0;
    auto E = [&](int i, int j) { return Z + F.oE + 36LL * (rowptr[i] + j - fcol[i]); };
    double* y = Z + F.oY;
    for (int k = F.n - 1; k >= 0; k--) {
        double v[6];
        for (int c = 0; c < 6; c++) v[c] = y[6 * k + c];
        for (int i = k + 1; i < F.n && i < k + F.W; i++)
            if (fcol[i] <= k)
                for (int c = 0; c < 6; c++)
                    for (int r = 0; r < 6; r++) v[c] -= E(i, k)[6 * r + c] * y[6 * i + r];
        for (int b = 0; b < F.nbr; b++) {
            const double* xb = Z + R.oY + 6LL * P.bord[F.bord0 + b];
            const double* Lb = Z + F.oB + 36LL * ((long long)k * F.nbr + b);
            for (int c = 0; c < 6; c++)
                for (int r = 0; r < 6; r++) v[c] -= Lb[6 * r + c] * xb[r];
        }
        const double* Lkk = E(k, k);
        for (int c = 5; c >= 0; c--) {   // Lkk^T x = v
            double s = v[c];
            for (int r = c + 1; r < 6; r++) s -= Lkk[6 * r + c] * y[6 * k + r];
            y[6 * k + c] = s / Lkk[6 * c + c];
        }
        for (int c = 0; c < 6; c++) x[6 * P.row_node[F.row0 + k] + c] = y[6 * k + c];
    }
}
}  // namespace

extern "C" int uco_b200_probe_block_solve(int nb, int nblk, const int* blk_ij, const double* blocks, const double* rhs, int smem_optin, int force_k,
                                          double* x, int* info8) {
    if (nb < 0 || nblk < 0 || (nblk && (!blk_ij || !blocks)) || (nb && !rhs)) return -1;
    for (int b = 0; b < nblk; b++)
        if (blk_ij[2 * b] < 0 || blk_ij[2 * b] > blk_ij[2 * b + 1] || blk_ij[2 * b + 1] >= nb) return -1;
    uco_band_plan P;
    uco_band_make_plan(nb, nblk, (const int2*)blk_ij, smem_optin > 0 ? smem_optin : 232448, force_k, P);
    if (info8) {
        int maxn = 0, maxW = 0, maxbr = 0;
        for (size_t f = 0; f + 1 < P.fronts.size(); f++) {
            maxn = std::max(maxn, P.fronts[f].n);
            maxW = std::max(maxW, P.fronts[f].W);
            maxbr = std::max(maxbr, P.fronts[f].nbr);
        }
        info8[0] = (int)P.fronts.size(); info8[1] = P.K; info8[2] = P.fronts.empty() ? 0 : P.fronts.back().n; info8[3] = P.fronts.empty() ? 0 : P.fronts.back().W;
        info8[4] = maxn; info8[5] = maxW; info8[6] = maxbr; info8[7] = (int)std::min<long long>(P.z_doubles, 0x7fffffff);
    }
    if (!uco_band_plan_valid(P)) return -2;
    if (!x) return 0;
    std::vector<double> Z((size_t)P.z_doubles + 1, 0.0);
    for (int b = 0; b < nblk; b++) {
        const long long off = P.blk_dst[b] >> 1;
        const bool tr = P.blk_dst[b] & 1;
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) Z[off + (tr ? 6 * c + r : 6 * r + c)] = blocks[36 * (size_t)b + 6 * r + c];
    }
    for (int v = 0; v < nb; v++)
        for (int r = 0; r < 6; r++) Z[P.rhs_dst[v] + r] = rhs[6 * (size_t)v + r];
    const int nf = (int)P.fronts.size();
    for (int f = 0; f + 1 < nf; f++)
        if (!host_front(P, f, Z.data())) return 1;
    if (nf) {
        const uco_band_front& R = P.fronts.back();
        for (size_t u = 0; u < P.g_dst.size(); u++)
            for (int s = P.g_ptr[u]; s < P.g_ptr[u + 1]; s++)
                for (int e = 0; e < 36; e++) Z[P.g_dst[u] + e] -= Z[(P.g_src[s] >> 1) + e];
        for (int k = 0; k < R.n; k++)
            for (int s = P.r_ptr[k]; s < P.r_ptr[k + 1]; s++)
                for (int r = 0; r < 6; r++) Z[R.oY + 6 * k + r] -= Z[P.r_src[s] + r];
        if (!host_front(P, nf - 1, Z.data())) return 1;
        host_back(P, nf - 1, Z.data(), x);
        for (int f = 0; f + 1 < nf; f++) host_back(P, f, Z.data(), x);
    }
    return 0;
}

// ---- device ------------------------------------------------------------------------------------------------------------------------
namespace {

struct BandDev {
    const uco_band_front* fr;
    const int *fcol, *rowptr, *node, *bord;
    const long long *blk_dst, *rhs_dst, *g_dst, *g_src, *r_src;
    const int *g_ptr, *r_ptr;
    int nfronts, g_n;
    double *Z, *xp, *wglobal;
    int* fail;
};

// S = [i == j](Hpp + lambda I) - (summed Schur blocks) [+ marker block], written to its place in the fronts; right-hand sides too
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
    const long long code = D.blk_dst[blk];
    D.Z[(code >> 1) + ((code & 1) ? 6 * c + r : e)] = h;
    if (diag && c == 0) D.Z[D.rhs_dst[ij.x] + r] = bp[6 * ij.x + r] - bsp[6 * ij.x + r];
    if (blk == 0 && e == 0) *D.fail = 0;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

// back substitution L^T x = y over a front's envelope factor (diagonal slots hold the INVERSE of L_kk), column-oriented: once x_k is
// known, y_j -= L_kj^T x_k for the blocks of ROW k — which the row envelope stores contiguously.  No reduction: every participating
// thread computes x_k itself (21 FMA) and owns one entry of one y_j; one named barrier per step.  Rows arrive through a cp.async ring
// BACK_DEPTH steps ahead, so no L2 latency is on the chain.  Needs 18 W <= BAND_THREADS and (BACK_DEPTH + 1) 36 W doubles of ring.
constexpr int BACK_DEPTH = 4;
__device__ void band_back_substitute_ring(const double* __restrict__ E, const int* fcol, const int* rowptr, int n, int W, double* y, double* ring) {
    const int tid = threadIdx.x, B = W - 1;
    const int nthr = (18 * W + 31) / 32 * 32;         // whole warps: they copy the rows; the first 6 B threads also compute
    if (tid >= nthr || n <= 0) return;                // the others wait at the caller's next __syncthreads
    const int jj = tid / 6, c = tid % 6;
    auto issue = [&](int i) {                         // row i -> slot i % (BACK_DEPTH + 1); one commit group per row (empty below row 0)
        if (i >= 0 && tid < 18 * (i - fcol[i] + 1)) cp_async16(ring + 36 * W * (i % (BACK_DEPTH + 1)) + 2 * tid, E + 36 * (size_t)rowptr[i] + 2 * tid);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    for (int i = n - 1; i > n - 1 - BACK_DEPTH; i--) issue(i);
    double xprev = 0;
    for (int k = n - 1; k >= 0; k--) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(BACK_DEPTH - 1) : "memory");
        asm volatile("bar.sync 1, %0;\n" ::"r"(nthr) : "memory");   // row k has landed for everyone; the y updates of step k + 1 are visible
        if (tid < 6 && k + 1 < n) y[6 * (k + 1) + tid] = xprev;       // y_(k+1) is not read any more: it becomes x_(k+1)
        issue(k - BACK_DEPTH);                         // into the slot row k + 1 has just left
        const double* row = ring + 36 * W * (k % (BACK_DEPTH + 1));
        const int f = fcol[k];
        const double* Linv = row + 36 * (k - f);
        double yk[6], x[6];
#pragma unroll
        for (int r = 0; r < 6; r++) yk[r] = y[6 * k + r];
#pragma unroll
        for (int cc = 0; cc < 6; cc++) {
            double v = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) if (r >= cc) v = fma(Linv[6 * r + cc], yk[r], v);
            x[cc] = v;
        }
        const int j = k - 1 - jj;
        if (jj < B && j >= f) {
            const double* Lb = row + 36 * (j - f);
            double v = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) v = fma(Lb[6 * r + c], x[r], v);
            y[6 * j + c] -= v;
        }
        if (tid < 6) {
#pragma unroll
            for (int r = 0; r < 6; r++) if (r == tid) xprev = x[r];
        }
    }
    if (tid < 6) y[tid] = xprev;
}

// back substitution L^T x = y over a front's envelope factor (diagonal slots hold the INVERSE of L_kk): x_k = Linv_k^T (y_k - sum_{i > k}
// L_ik^T x_i).  Thread (ii, c) owns column c of block (k + 1 + ii, k), streamed from L2 one step ahead; warp 0 folds the partial sums.
__device__ void band_back_substitute(const double* __restrict__ E, const int* fcol, const int* rowptr, int n, int B, double* y, double* part) {
    const int tid = threadIdx.x;
    if (n <= 0) return;
    if (6 * B > BAND_THREADS) {   // very wide envelope (close to dense): plain multi-pass form, no prefetch
        __shared__ double sv6[6];
        for (int k = n - 1; k >= 0; k--) {
            for (int t = tid; t < 6 * B; t += BAND_THREADS) {
                const int i = k + 1 + t / 6, cc = t % 6;
                double v = 0;
                if (i < n && k >= fcol[i]) {
                    const double* src = E + 36 * (size_t)(rowptr[i] + k - fcol[i]);
                    for (int r = 0; r < 6; r++) v += src[6 * r + cc] * y[6 * i + r];
                }
                part[t] = v;
            }
            __syncthreads();
            if (tid < 6) {
                double sv = y[6 * k + tid];
                for (int a = 0; a < B; a++) sv -= part[6 * a + tid];
                sv6[tid] = sv;
            }
            __syncthreads();
            if (tid < 6) {
                const double* dg = E + 36 * (size_t)(rowptr[k] + k - fcol[k]);
                double v = 0;
                for (int r = tid; r < 6; r++) v += dg[6 * r + tid] * sv6[r];
                y[6 * k + tid] = v;
            }
            __syncthreads();
        }
        return;
    }
    const int ii = tid / 6, c = tid % 6;
    const bool worker = tid < 6 * B;
    double pre[6], dinv[6];
    auto fetch = [&](int k) {
        const int i = k + 1 + ii;
        const bool live = worker && i < n && k >= fcol[i];
        const double* src = live ? E + 36 * (size_t)(rowptr[i] + k - fcol[i]) : nullptr;
#pragma unroll
        for (int r = 0; r < 6; r++) pre[r] = live ? src[6 * r + c] : 0.0;
        if (tid < 6) {   // column c of Linv_k = row c of Linv_k^T
            const double* dg = E + 36 * (size_t)(rowptr[k] + k - fcol[k]);
#pragma unroll
            for (int r = 0; r < 6; r++) dinv[r] = dg[6 * r + c];
        }
    };
    fetch(n - 1);
    for (int k = n - 1; k >= 0; k--) {
        double cur[6], dcur[6];
#pragma unroll
        for (int r = 0; r < 6; r++) { cur[r] = pre[r]; dcur[r] = dinv[r]; }
        if (k > 0) fetch(k - 1);
        if (worker) {
            const int i = k + 1 + ii;
            double v = 0;
            if (i < n) {
#pragma unroll
                for (int r = 0; r < 6; r++) v += cur[r] * y[6 * i + r];
            }
            part[tid] = v;
        }
        __syncthreads();
        if (tid < 32) {
            double sv = 0;
            if (tid < 6) {
                sv = y[6 * k + tid];
                for (int a = 0; a < B; a++) sv -= part[6 * a + tid];
            }
            double v = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                const double svr = __shfl_sync(0xffffffffu, sv, r);
                if (tid < 6 && r >= c) v += dcur[r] * svr;
            }
            if (tid < 6) y[6 * k + tid] = v;
        }
        __syncthreads();
    }
}

// one front: partial block Cholesky of [interior envelope ; border rows], see the file header.  blockIdx.x + f0 = front.
template <bool win_in_smem>
__global__ void __launch_bounds__(BAND_THREADS) band_front_kernel(BandDev D, int f0, int is_root) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x;
    const uco_band_front F = D.fr[f0 + blockIdx.x];
    const int n = F.n, W = F.W, B = W - 1, nbr = F.nbr, nslot = B + nbr;
    if (n == 0) return;
    double* y = sm;                                   // 6 n: right-hand side -> forward solution -> solution
    double* bv = y + 6 * (size_t)n;                   // 6 nbr: reduced right-hand side of the border
    double* Lkk = bv + 6 * (size_t)nbr;               // 36: Cholesky factor of the pivot
    double* sdiag = Lkk + 36;                         // 6: reciprocals of its diagonal
    double* ynew = sdiag + 6;                         // 6 (+8 pad)
    double* lcol = ynew + 14;                         // (B + nbr) blocks: the scaled column of the current step (window slots, then border)
    double* bwin = lcol + 36 * (size_t)nslot;         // W x nbr blocks: border rows over the window's columns (circular)
    double* Sbb = bwin + 36 * (size_t)W * nbr;        // nbr x nbr blocks: Schur complement of the border (lower block triangle)
    double* after = Sbb + 36 * (size_t)nbr * nbr;
    double* win = win_in_smem ? after : D.wglobal;    // W x W blocks, circular: block (i, j) at ((i % W) * W + (j % W)) * 36 (template: shared loads when in smem)
    int* fcol = (int*)(win_in_smem ? after + 36 * (size_t)max(W * W, 5 * W) : after);
    int* rowptr = fcol + n;
    int* ptab = rowptr + n;                           // slot pairs a >= b, a << 16 | b
    __shared__ int s_fail;
    double* const E = D.Z + F.oE;
    double* const Bd = D.Z + F.oB;
    // window slots are circular in rows and columns; inside the loop they are tracked incrementally (s0 = k % W), no integer division
    auto wrap = [&](int x) { return x >= W ? x - W : x; };   // x < 2 W
    auto wblk = [&](int rs, int cs) -> double* { return win + 36 * (rs * W + cs); };   // by slot
    auto load_row = [&](int i, int rs) {              // row i of the envelope into row slot rs = i % W (zero left of f(i)), 16 bytes per copy
        const int f = fcol[i];
        const double* src = E + 36 * (size_t)rowptr[i];
        for (int t = tid; t < 18 * W; t += BAND_THREADS) {
            const int jj = t / 18, ch = t % 18, j = i - B + jj;   // column j sits in column slot (rs + 1 + jj) % W
            if (j < 0) continue;
            double* dst = wblk(rs, wrap(rs + 1 + jj)) + 2 * ch;
            if (j >= f) {
                if (win_in_smem) cp_async16(dst, src + 36 * (size_t)(j - f) + 2 * ch);
                else { dst[0] = src[36 * (size_t)(j - f) + 2 * ch]; dst[1] = src[36 * (size_t)(j - f) + 2 * ch + 1]; }
            } else { dst[0] = 0.0; dst[1] = 0.0; }
        }
    };
    auto load_bcol = [&](int k, int cs) {             // border blocks of column k into column slot cs = k % W
        const double* src = Bd + 36 * (size_t)k * nbr;
        double* dst = bwin + 36 * cs * nbr;
        for (int t = tid; t < 18 * nbr; t += BAND_THREADS) cp_async16(dst + 2 * t, src + 2 * t);
    };
    if (tid == 0) s_fail = 0;
    for (int t = tid; t < n; t += BAND_THREADS) { fcol[t] = D.fcol[F.row0 + t]; rowptr[t] = D.rowptr[F.row0 + t]; }
    for (int a = tid; a < nslot; a += BAND_THREADS)
        for (int b = 0; b <= a; b++) ptab[a * (a + 1) / 2 + b] = a << 16 | b;
    if (is_root) {   // the fronts' Schur complements and reduced right-hand sides, destination-major, sources in front order
        for (int t = tid; t < 36 * D.g_n; t += BAND_THREADS) {
            const int u = t / 36, e = t % 36;
            double v = D.Z[D.g_dst[u] + e];
            for (int s = D.g_ptr[u]; s < D.g_ptr[u + 1]; s++) v -= D.Z[(D.g_src[s] >> 1) + e];
            D.Z[D.g_dst[u] + e] = v;
        }
        for (int t = tid; t < 6 * n; t += BAND_THREADS) {
            const int k = t / 6, r = t % 6;
            double v = D.Z[F.oY + t];
            for (int s = D.r_ptr[k]; s < D.r_ptr[k + 1]; s++) v -= D.Z[D.r_src[s] + r];
            y[t] = v;
        }
    } else {
        for (int t = tid; t < 6 * n; t += BAND_THREADS) y[t] = D.Z[F.oY + t];
    }
    for (int t = tid; t < 6 * nbr; t += BAND_THREADS) bv[t] = 0.0;
    for (int t = tid; t < 36 * nbr * nbr; t += BAND_THREADS) Sbb[t] = 0.0;
    __syncthreads();
    if (*(volatile int*)D.fail) s_fail = 2;   // a front of this trial already failed: nothing downstream is used
    for (int i = 0; i < n && i < W; i++) {
        load_row(i, i);
        if (nbr) load_bcol(i, i);
    }
    cp_async_wait_all();
    __syncthreads();

    // Cholesky of a pivot block: one thread, registers, no shuffles (the dependent chain of six rsqrt's is the cost)
    auto pivot = [&](const double* A) {
        double L[6][6], s[6];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 6; j++) {
            double d = A[6 * j + j];
#pragma unroll
            for (int q = 0; q < 6; q++) if (q < j) d = fma(-L[j][q], L[j][q], d);
            ok = ok && d > 0 && isfinite(d);
            s[j] = rsqrt(d);
            L[j][j] = d * s[j];
#pragma unroll
            for (int i = 0; i < 6; i++)
                if (i > j) {
                    double v = A[6 * i + j];
#pragma unroll
                    for (int q = 0; q < 6; q++) if (q < j) v = fma(-L[i][q], L[j][q], v);
                    L[i][j] = v * s[j];
                }
        }
#pragma unroll
        for (int e = 0; e < 36; e++) Lkk[e] = e % 6 <= e / 6 ? L[e / 6][e % 6] : 0.0;
#pragma unroll
        for (int j = 0; j < 6; j++) sdiag[j] = s[j];
        if (!ok) s_fail = 1;
    };
    // the trailing update: a thread owns HALF a 6x6 block (three rows) of one block pair (a, b) of the scaled column: 9 + 18 16-byte
    // operand loads, 9 + 9 for the destination, 126 DFMA — the step is bound by instruction issue and shared-memory latency, not by
    // flops, so the tile is as large as the registers allow.  Pair 0 (the next pivot block) belongs to warp 0, 18 lanes x 2 entries.
    constexpr int NPASS = (BAND_THREADS - 32) / 2;    // pairs per pass of the other 15 warps
    const int gt = tid - 32, half = gt & 1;
    auto ptab_pair = [&](int p) {                     // pair p of the enumeration a >= b, a-major (= ptab[p], which is not filled yet)
        int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
        while (a * (a + 1) / 2 > p) a--;
        while ((a + 1) * (a + 2) / 2 <= p) a++;
        return a << 16 | (p - a * (a + 1) / 2);
    };
    const int mypr = gt >= 0 && 1 + (gt >> 1) < F.npairs ? ptab_pair(1 + (gt >> 1)) : -1;   // a thread's pair of the first pass never changes
    auto dst_of = [&](int s1, int a, int b, double& sign) -> double* {
        sign = -1.0;
        if (a < B) return wblk(wrap(s1 + a), wrap(s1 + b));
        if (b < B) return bwin + 36 * (wrap(s1 + b) * nbr + (a - B));
        sign = 1.0;
        return Sbb + 36 * ((a - B) * nbr + (b - B));
    };
    auto update = [&](int s1, int nrow, int pr, int ur, int uc) {     // two entries (ur, uc), (ur, uc + 1); s1 = (k + 1) % W
        const int a = pr >> 16, b = pr & 0xffff;
        if ((a < B && a >= nrow) || (b < B && b >= nrow)) return;
        const double2 *La = (const double2*)(lcol + 36 * a + 6 * ur), *Lb = (const double2*)(lcol + 36 * b + 6 * uc);
        const double2 a0 = La[0], a1 = La[1], a2 = La[2], b0 = Lb[0], b1 = Lb[1], b2 = Lb[2], c0 = Lb[3], c1 = Lb[4], c2 = Lb[5];
        const double v0 = fma(a2.y, b2.y, fma(a2.x, b2.x, fma(a1.y, b1.y, fma(a1.x, b1.x, fma(a0.y, b0.y, a0.x * b0.x)))));
        const double v1 = fma(a2.y, c2.y, fma(a2.x, c2.x, fma(a1.y, c1.y, fma(a1.x, c1.x, fma(a0.y, c0.y, a0.x * c0.x)))));
        double sign;
        double2* dst = (double2*)(dst_of(s1, a, b, sign) + 6 * ur + uc);
        double2 o = *dst;
        o.x = fma(sign, v0, o.x);
        o.y = fma(sign, v1, o.y);
        *dst = o;
    };
    auto update_half = [&](int s1, int nrow, int pr) {   // rows 3 half .. 3 half + 2 of the pair's block
        const int a = pr >> 16, b = pr & 0xffff;
        if (pr < 0 || (a < B && a >= nrow) || (b < B && b >= nrow)) return;
        double sign;
        double2* dst = (double2*)(dst_of(s1, a, b, sign) + 18 * half);
        const double2 *La = (const double2*)(lcol + 36 * a + 18 * half), *Lb = (const double2*)(lcol + 36 * b);
        double2 A[9], O[9];
#pragma unroll
        for (int q = 0; q < 9; q++) { A[q] = La[q]; O[q] = dst[q]; }
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const double2 b0 = Lb[3 * c], b1 = Lb[3 * c + 1], b2 = Lb[3 * c + 2];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double v = fma(A[3 * r + 2].y, b2.y, fma(A[3 * r + 2].x, b2.x, fma(A[3 * r + 1].y, b1.y, fma(A[3 * r + 1].x, b1.x, fma(A[3 * r].y, b0.y, A[3 * r].x * b0.x)))));
                if (c & 1) O[3 * r + c / 2].y = fma(sign, v, O[3 * r + c / 2].y);
                else O[3 * r + c / 2].x = fma(sign, v, O[3 * r + c / 2].x);
            }
        }
#pragma unroll
        for (int q = 0; q < 9; q++) dst[q] = O[q];
    };
    if (tid == 0 && !s_fail) pivot(wblk(0, 0));
    __syncthreads();
    // per block column k (its pivot is factored already: look-ahead): P2 scale the column, P3 trailing update — with warp 0 updating
    // the NEXT pivot block first and factoring it while the other warps update the rest
    int s0 = 0;                                       // k % W
    for (int k = 0; k < n && !s_fail; k++) {
        const int nrow = min(B, n - 1 - k);           // window rows below the pivot
        const int s1 = wrap(s0 + 1);
        if (k + W < n) load_row(k + W, s0);           // the slot of row k is free: its only live block was the pivot
        // P2: L_ik = A_ik Lkk^-T by forward substitution, one thread per block row, for the window rows and the border rows; y_k; Linv
        for (int t = tid; t < 6 * (nrow + nbr); t += BAND_THREADS) {
            const int which = t / 6, r = t % 6;
            const bool wrow = which < nrow;
            const int i = k + 1 + which, b = which - nrow;
            const double2* A2 = (const double2*)((wrow ? wblk(wrap(s1 + which), s0) : bwin + 36 * (s0 * nbr + b)) + 6 * r);
            const double2 A01 = A2[0], A23 = A2[1], A45 = A2[2];
            const double A[6] = {A01.x, A01.y, A23.x, A23.y, A45.x, A45.y};
            double l[6];
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double v = A[c];
#pragma unroll
                for (int q = 0; q < 6; q++) if (q < c) v = fma(-l[q], Lkk[6 * c + q], v);
                l[c] = v * sdiag[c];
            }
            double2* lc = (double2*)(lcol + 36 * (wrow ? which : B + b) + 6 * r);
            lc[0] = make_double2(l[0], l[1]); lc[1] = make_double2(l[2], l[3]); lc[2] = make_double2(l[4], l[5]);
            double2* g = (double2*)(wrow ? (k >= fcol[i] ? E + 36 * (size_t)(rowptr[i] + k - fcol[i]) + 6 * r : nullptr) : Bd + 36 * ((size_t)k * nbr + b) + 6 * r);
            if (g) { g[0] = make_double2(l[0], l[1]); g[1] = make_double2(l[2], l[3]); g[2] = make_double2(l[4], l[5]); }
        }
        if (tid == BAND_THREADS - 1) {                // y_k <- Lkk^-1 y_k
            double v[6];
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double a = y[6 * k + r];
#pragma unroll
                for (int q = 0; q < 6; q++) if (q < r) a = fma(-Lkk[6 * r + q], v[q], a);
                v[r] = a * sdiag[r];
                ynew[r] = v[r];
            }
        } else if (tid >= BAND_THREADS - 64 && tid < BAND_THREADS - 58) {   // column c of Lkk^-1 -> the factor's diagonal slot (back substitution)
            const int c = tid - (BAND_THREADS - 64);
            double inv[6];
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double a = r == c ? 1.0 : 0.0;
#pragma unroll
                for (int q = 0; q < 6; q++) if (q < r && q >= c) a = fma(-Lkk[6 * r + q], inv[q], a);
                inv[r] = r < c ? 0.0 : a * sdiag[r];
            }
            double* dg = E + 36 * (size_t)(rowptr[k] + k - fcol[k]);
#pragma unroll
            for (int r = 0; r < 6; r++) dg[6 * r + c] = inv[r];
        }
        __syncthreads();
        if (nbr && k + W < n) load_bcol(k + W, s0);   // the border slot of column k is free from here on
        // P3: trailing update of the window, the border rows, the border's Schur complement and the right-hand sides
        if (tid < 32) {
            if (nrow > 0) {                           // look-ahead: block (k+1, k+1) first, then its Cholesky while the other warps work
                if (tid < 18) update(s1, nrow, 0, tid / 3, (tid % 3) * 2);
                __syncwarp();
                if (tid == 0) pivot(wblk(s1, s1));
            }
        } else {
            update_half(s1, nrow, mypr);
            for (int p = 1 + NPASS + (gt >> 1); p < F.npairs; p += NPASS) update_half(s1, nrow, ptab[p]);
            for (int t = gt; t < 6 * nslot; t += BAND_THREADS - 32) {
                const int a = t / 6, r = t % 6;
                if (a < B && a >= nrow) continue;
                const double* La = lcol + 36 * a + 6 * r;
                double v = 0;
#pragma unroll
                for (int q = 0; q < 6; q++) v = fma(La[q], ynew[q], v);
                if (a < B) y[6 * (k + 1 + a) + r] -= v;
                else bv[6 * (a - B) + r] += v;
            }
            if (gt < 6) y[6 * k + gt] = ynew[gt];
        }
        cp_async_wait_all();
        __syncthreads();
        s0 = s1;
    }
    __syncthreads();
    if (s_fail) {
        if (tid == 0) *D.fail = 1;
        if (nbr == 0)
            for (int t = tid; t < 6 * n; t += BAND_THREADS) {
                D.xp[6 * D.node[F.row0 + t / 6] + t % 6] = 0;
                D.Z[F.oY + t] = 0;
            }
        return;
    }
    if (nbr) {   // an interior front: hand the Schur complement, the reduced right-hand side and the forward solution on
        for (int t = tid; t < 36 * nbr * nbr; t += BAND_THREADS) D.Z[F.oS + t] = Sbb[t];
        for (int t = tid; t < 6 * nbr; t += BAND_THREADS) D.Z[F.oR + t] = bv[t];
        for (int t = tid; t < 6 * n; t += BAND_THREADS) D.Z[F.oY + t] = y[t];
        return;
    }
    if (win_in_smem && 18 * W <= BAND_THREADS) band_back_substitute_ring(E, fcol, rowptr, n, W, y, lcol);   // the ring takes over lcol + window
    else band_back_substitute(E, fcol, rowptr, n, B, y, lcol);
    __syncthreads();
    for (int t = tid; t < 6 * n; t += BAND_THREADS) {
        D.xp[6 * D.node[F.row0 + t / 6] + t % 6] = y[t];
        D.Z[F.oY + t] = y[t];                         // the root's solution in root order: the interior fronts' borders read it
    }
}

// interior fronts with a border: x = L^-T (y - L_border^T x_border)
__global__ void __launch_bounds__(BAND_THREADS) band_back_kernel(BandDev D) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x;
    const uco_band_front F = D.fr[blockIdx.x];
    const int n = F.n, B = F.W - 1, nbr = F.nbr;
    if (n == 0 || nbr == 0) return;                   // fronts without a border finished in band_front_kernel
    if (*(volatile int*)D.fail) {
        for (int t = tid; t < 6 * n; t += BAND_THREADS) D.xp[6 * D.node[F.row0 + t / 6] + t % 6] = 0;
        return;
    }
    double* y = sm;
    double* xb = y + 6 * (size_t)n;
    double* part = xb + 6 * (size_t)nbr;
    double* ring = part + 6 * (size_t)max(B, 1);
    int* fcol = (int*)(ring + 36 * (size_t)(BACK_DEPTH + 1) * F.W);
    int* rowptr = fcol + n;
    const uco_band_front R = D.fr[D.nfronts - 1];
    for (int t = tid; t < n; t += BAND_THREADS) { fcol[t] = D.fcol[F.row0 + t]; rowptr[t] = D.rowptr[F.row0 + t]; }
    for (int t = tid; t < 6 * nbr; t += BAND_THREADS) xb[t] = D.Z[R.oY + 6 * (size_t)D.bord[F.bord0 + t / 6] + t % 6];
    __syncthreads();
    for (int t = tid; t < 6 * n; t += BAND_THREADS) {
        const int k = t / 6, c = t % 6;
        const double* Lb = D.Z + F.oB + 36 * (size_t)k * nbr;
        double v = D.Z[F.oY + t];
        for (int b = 0; b < nbr; b++)
#pragma unroll
            for (int r = 0; r < 6; r++) v -= Lb[36 * b + 6 * r + c] * xb[6 * b + r];
        y[t] = v;
    }
    __syncthreads();
    if (18 * F.W <= BAND_THREADS) band_back_substitute_ring(D.Z + F.oE, fcol, rowptr, n, F.W, y, ring);
    else band_back_substitute(D.Z + F.oE, fcol, rowptr, n, B, y, part);
    __syncthreads();
    for (int t = tid; t < 6 * n; t += BAND_THREADS) D.xp[6 * D.node[F.row0 + t / 6] + t % 6] = y[t];
}

size_t back_smem(int n, int W, int nbr) {
    return 8 * (6 * (size_t)n + 6 * (size_t)nbr + 6 * (size_t)std::max(W - 1, 1) + 36 * (size_t)(BACK_DEPTH + 1) * W) + 8 * (size_t)n + 16;
}

BandDev band_dev(const uco_band_plan& P, const unsigned char* blob_dev, double* Z, double* xp, int* fail_dev, double* wglobal) {
    BandDev D;
    D.fr = (const uco_band_front*)(blob_dev + P.o_fronts);
    D.fcol = (const int*)(blob_dev + P.o_fcol); D.rowptr = (const int*)(blob_dev + P.o_rowptr); D.node = (const int*)(blob_dev + P.o_node);
    D.bord = (const int*)(blob_dev + P.o_bord);
    D.blk_dst = (const long long*)(blob_dev + P.o_blk_dst); D.rhs_dst = (const long long*)(blob_dev + P.o_rhs_dst);
    D.g_dst = (const long long*)(blob_dev + P.o_g_dst); D.g_src = (const long long*)(blob_dev + P.o_g_src); D.r_src = (const long long*)(blob_dev + P.o_r_src);
    D.g_ptr = (const int*)(blob_dev + P.o_g_ptr); D.r_ptr = (const int*)(blob_dev + P.o_r_ptr);
    D.nfronts = (int)P.fronts.size(); D.g_n = (int)P.g_dst.size();
    D.Z = Z; D.xp = xp; D.wglobal = wglobal; D.fail = fail_dev;
    return D;
}

}  // namespace

// ---- launcher used by ba.cu and uco_b200_block_solve ---------------------------------------------------------------------------------
// blob_dev: P.blob on the device; Z: P.z_doubles doubles; wglobal: P.scratch_doubles doubles (or null when 0).  Asynchronous on the stream.
int uco_band_solve_launch(uco_b200_ctx* ctx, const uco_band_plan& P, const unsigned char* blob_dev, double* Z, double* wglobal, const int2* blk_ij_dev,
                          const double* Hpp, const double* lambda_dev, const double* Sp, const double* bp, const double* bsp, const int* mk_blk_edge,
                          const double* mk_e_blk, double* xp, int* fail_dev, int smem_optin) {
    const int nf = (int)P.fronts.size();
    if (nf == 0 || P.nb == 0) return UCO_OK;
    if (!uco_band_plan_valid(P)) return uco_fail(ctx, UCO_E_INVALID, "band solve: a block couples two interior fronts (planner bug)");
    BandDev D = band_dev(P, blob_dev, Z, xp, fail_dev, wglobal);
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemsetAsync(Z, 0, 8 * (size_t)P.z_doubles, s));
    band_assemble_kernel<<<P.nblk, 36, 0, s>>>(D, blk_ij_dev, Hpp, lambda_dev, Sp, bp, bsp, mk_blk_edge, mk_e_blk);
    UCO_LAUNCH_CHECK(ctx);
    size_t smem_fr = 0, smem_bk = 0;
    for (int f = 0; f + 1 < nf; f++) {
        const uco_band_front& F = P.fronts[f];
        smem_fr = std::max(smem_fr, front_smem_fixed(F.n, F.W, F.nbr) + front_smem_window(F.W));
        smem_bk = std::max(smem_bk, back_smem(F.n, F.W, F.nbr));
    }
    const uco_band_front& R = P.fronts[nf - 1];
    const size_t root_fixed = front_smem_fixed(R.n, R.W, 0), root_win = front_smem_window(R.W);
    const bool root_in_smem = root_fixed + root_win <= (size_t)smem_optin;
    if (root_fixed > (size_t)smem_optin) return uco_fail(ctx, UCO_E_CAPACITY, "band solve: %d block unknowns exceed the shared-memory right-hand side", R.n);
    if (smem_fr > (size_t)smem_optin) return uco_fail(ctx, UCO_E_CAPACITY, "band solve: an interior front exceeds shared memory (planner bug)");
    if (!root_in_smem && !wglobal) return uco_fail(ctx, UCO_E_INVALID, "band solve: no scratch for the root window");
    const size_t smem_root = root_in_smem ? root_fixed + root_win : root_fixed;
    static size_t configured_front = 0, configured_back = 0;   // grow-only attributes (same values from every thread)
    const size_t need = std::max(smem_fr, smem_root);
    if (need > configured_front) {
        UCO_CUDA(ctx, cudaFuncSetAttribute(band_front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        UCO_CUDA(ctx, cudaFuncSetAttribute(band_front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        configured_front = need;
    }
    if (smem_bk > configured_back) {
        UCO_CUDA(ctx, cudaFuncSetAttribute(band_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bk));
        configured_back = smem_bk;
    }
    if (nf > 1) {
        band_front_kernel<true><<<nf - 1, BAND_THREADS, smem_fr, s>>>(D, 0, 0);
        UCO_LAUNCH_CHECK(ctx);
    }
    if (root_in_smem) band_front_kernel<true><<<1, BAND_THREADS, smem_root, s>>>(D, nf - 1, 1);
    else band_front_kernel<false><<<1, BAND_THREADS, smem_root, s>>>(D, nf - 1, 1);
    UCO_LAUNCH_CHECK(ctx);
    bool any_border = false;
    for (int f = 0; f + 1 < nf; f++) any_border = any_border || P.fronts[f].nbr > 0;
    if (any_border) {
        band_back_kernel<<<nf - 1, BAND_THREADS, smem_bk, s>>>(D);
        UCO_LAUNCH_CHECK(ctx);
    }
    return UCO_OK;
}

static thread_local float g_last_solve_ms = 0.f;
extern "C" int uco_b200_block_solve_profile(uco_b200_ctx* ctx, float* ms) {
    if (!ctx || !ms) return UCO_E_INVALID;
    *ms = g_last_solve_ms;
    return UCO_OK;
}

// ---- the LinearSolver seam on its own: S x = b for a symmetric positive definite block-sparse S given by its upper block triangle ----
extern "C" int uco_b200_block_solve(uco_b200_ctx* ctx, int nb, int nblk, const int* blk_ij, const double* blocks, const double* rhs, int force_k, double* x,
                                    int* info8) {
    if (!ctx) return UCO_E_INVALID;
    if (nb < 0 || nblk < 0 || (nblk && (!blk_ij || !blocks)) || (nb && (!rhs || !x))) return uco_fail(ctx, UCO_E_INVALID, "block_solve: null argument");
    std::vector<char> has_diag(nb, 0);
    for (int b = 0; b < nblk; b++) {
        if (blk_ij[2 * b] < 0 || blk_ij[2 * b] > blk_ij[2 * b + 1] || blk_ij[2 * b + 1] >= nb)
            return uco_fail(ctx, UCO_E_INVALID, "block_solve: block %d = (%d, %d) is not in the upper triangle of %d block rows", b, blk_ij[2 * b], blk_ij[2 * b + 1], nb);
        if (blk_ij[2 * b] == blk_ij[2 * b + 1]) has_diag[blk_ij[2 * b]] = 1;
    }
    for (int v = 0; v < nb; v++)
        if (!has_diag[v]) return uco_fail(ctx, UCO_E_INVALID, "block_solve: block row %d has no diagonal block", v);
    if (info8) memset(info8, 0, 32);
    if (nb == 0) return UCO_OK;
    int smem_optin = 0;
    UCO_CUDA(ctx, cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    uco_band_plan P;
    uco_band_make_plan(nb, nblk, (const int2*)blk_ij, smem_optin, force_k, P);
    if (info8) {
        int maxn = 0, maxW = 0, maxbr = 0;
        for (size_t f = 0; f + 1 < P.fronts.size(); f++) {
            maxn = std::max(maxn, P.fronts[f].n); maxW = std::max(maxW, P.fronts[f].W); maxbr = std::max(maxbr, P.fronts[f].nbr);
        }
        info8[0] = (int)P.fronts.size(); info8[1] = P.K; info8[2] = P.fronts.back().n; info8[3] = P.fronts.back().W;
        info8[4] = maxn; info8[5] = maxW; info8[6] = maxbr;
    }
    // device layout: blob | blk_ij | negated blocks (the assemble kernel subtracts its Schur input) | rhs | zeros (Hpp, bsp) | lambda | x | fail | Z | scratch
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_blob = 0, o_ij = al(P.blob.size()), o_S = o_ij + al(8 * (size_t)nblk), o_b = o_S + al(8 * 36 * (size_t)nblk), o_zero = o_b + al(48 * (size_t)nb),
                 o_lam = o_zero + al(8 * 36 * (size_t)nb), o_x = o_lam + 256, o_fail = o_x + al(48 * (size_t)nb), o_Z = o_fail + 256,
                 o_scr = o_Z + al(8 * (size_t)P.z_doubles), total = o_scr + al(8 * P.scratch_doubles);
    uint8_t* d = (uint8_t*)uco_ws(ctx, WS_GENERIC0, total);
    if (!d) return UCO_E_NOMEM;
    std::vector<double> neg(36 * (size_t)nblk);
    for (size_t t = 0; t < neg.size(); t++) neg[t] = -blocks[t];
    cudaStream_t s = ctx->stream;
    UCO_CUDA(ctx, cudaMemcpyAsync(d + o_blob, P.blob.data(), P.blob.size(), cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(d + o_ij, blk_ij, 8 * (size_t)nblk, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(d + o_S, neg.data(), 8 * neg.size(), cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(d + o_b, rhs, 48 * (size_t)nb, cudaMemcpyHostToDevice, s));
    UCO_CUDA(ctx, cudaMemsetAsync(d + o_zero, 0, o_x - o_zero, s));
    int rc = uco_band_solve_launch(ctx, P, d + o_blob, (double*)(d + o_Z), P.scratch_doubles ? (double*)(d + o_scr) : nullptr, (const int2*)(d + o_ij),
                                   (const double*)(d + o_zero), (const double*)(d + o_lam), (const double*)(d + o_S), (const double*)(d + o_b),
                                   (const double*)(d + o_zero), nullptr, nullptr, (double*)(d + o_x), (int*)(d + o_fail), smem_optin);
    if (rc != UCO_OK) return rc;
    if (ctx->profiling) {   // the same launches once more between events (the first pass warmed the caches the way an LM loop's previous trial does)
        cudaEvent_t e0, e1;
        UCO_CUDA(ctx, cudaEventCreate(&e0));
        UCO_CUDA(ctx, cudaEventCreate(&e1));
        UCO_CUDA(ctx, cudaEventRecord(e0, s));
        rc = uco_band_solve_launch(ctx, P, d + o_blob, (double*)(d + o_Z), P.scratch_doubles ? (double*)(d + o_scr) : nullptr, (const int2*)(d + o_ij),
                                   (const double*)(d + o_zero), (const double*)(d + o_lam), (const double*)(d + o_S), (const double*)(d + o_b),
                                   (const double*)(d + o_zero), nullptr, nullptr, (double*)(d + o_x), (int*)(d + o_fail), smem_optin);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&g_last_solve_ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (rc != UCO_OK) return rc;
    }
    int fail = 0;
    UCO_CUDA(ctx, cudaMemcpyAsync(x, d + o_x, 48 * (size_t)nb, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaMemcpyAsync(&fail, d + o_fail, 4, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    if (info8) info8[7] = fail;
    return UCO_OK;
}
