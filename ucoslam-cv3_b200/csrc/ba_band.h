// ba_band.h — plan of the two-level block-envelope Cholesky (ba_band.cu) and its launcher, shared with ba.cu.
#pragma once
#include "common.cuh"
#include <vector>

struct uco_band_front {
    int n, W, nbr, row0, bord0, npairs;   // interior block rows; window = local bmax + 1; border rows; offsets into the per-row / border arrays; slot pairs
    long long oE, oB, oS, oY, oR;         // offsets (doubles) into Z: envelope, border blocks [k][b], Schur out [a][b], rhs / y / x, reduced rhs out
};

struct uco_band_plan {
    int nb = 0, nblk = 0, K = 0;
    std::vector<uco_band_front> fronts;          // interior fronts, the root LAST
    std::vector<int> row_fcol, row_ptr, row_node;   // per front row (concatenated): first column (local), first envelope block (local), caller's block index
    std::vector<int> bord;                          // per front border slot (concatenated): root-local row
    std::vector<long long> blk_dst;                 // per caller block: destination offset in Z, bit 0 = store the transpose
    std::vector<long long> rhs_dst;                 // per caller block row
    std::vector<long long> g_dst, g_src;            // root gather: unique destination blocks; sources (offset << 1 | transpose)
    std::vector<int> g_ptr;
    std::vector<int> r_ptr;                         // root rows: reduced right-hand side sources
    std::vector<long long> r_src;
    long long z_doubles = 0;
    size_t scratch_doubles = 0;                     // global window of a root too wide for shared memory
    std::vector<unsigned char> blob;                // everything above, laid out for the device
    size_t o_fronts = 0, o_fcol = 0, o_rowptr = 0, o_node = 0, o_bord = 0, o_blk_dst = 0, o_rhs_dst = 0, o_g_dst = 0, o_g_src = 0, o_g_ptr = 0,
           o_r_ptr = 0, o_r_src = 0;
};

// smem_optin: opt-in shared memory per block of the device; force_k < 0: choose the number of separator levels by the cost model
void uco_band_make_plan(int nb, int nblk, const int2* blk_ij, int smem_optin, int force_k, uco_band_plan& P);
bool uco_band_plan_valid(const uco_band_plan& P);
// blob_dev: P.blob on the device; Z: P.z_doubles doubles; wglobal: P.scratch_doubles doubles (null when 0).  Asynchronous on the stream.
int uco_band_solve_launch(uco_b200_ctx* ctx, const uco_band_plan& P, const unsigned char* blob_dev, double* Z, double* wglobal, const int2* blk_ij_dev,
                          const double* Hpp, const double* lambda_dev, const double* Sp, const double* bp, const double* bsp, const int* mk_blk_edge,
                          const double* mk_e_blk, double* xp, int* fail_dev, int smem_optin);
