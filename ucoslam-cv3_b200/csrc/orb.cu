// orb.cu — K1..K6: the ORB pyramid extractor of the reference, device resident end to end.
//
// Replaces ORBextractor::compute and everything below it (reference, relative to /root/reference):
//   src/featureextractors/ORBextractor.cpp:1247-1351  compute(): blur, pyramid, per-level keypoints + descriptors, concatenate
//   :1355-1392  ComputePyramid   (cv::resize INTER_CUBIC chain + copyMakeBorder REFLECT_101, 19 px)
//   :899-1076   ComputeKeyPoints_thread (grid cells, cv::FAST th 20 / fallback 7, quota redistribution, retainBest)
//   :79-106     IC_Angle,  :113-153 computeOrbDescriptor,  :1120-1137 computeDescriptors,  :1229 coordinate rescale
// and the OpenCV primitives those call (GaussianBlur 8U fixed point, resize INTER_CUBIC 8U, FAST-9/16 score + 3x3 NMS,
// KeyPointsFilter::retainBest, fastAtan2, cvRound), restated bit-exactly (see DESIGN.md "ORB: arithmetic contracts").
//
// Data layout in HBM (per context, sized for a batch of frames):
//   pyramid : per frame one block holding all levels; each level is a (h+38) x pitch u8 image with the 19-px reflected
//             border already filled (pitch = w+38 rounded up to 64 B), so FAST / orientation / rBRIEF never branch on edges
//   cand    : per frame, per grid cell a slot of packed candidates  score<<24 | y<<12 | x  in row-major scan order
//   sel     : per frame, per level the selected keypoints in the reference's final order
//   out     : per frame max_features x 28-B cv::KeyPoint + max_features x 32-B descriptors, level-major
// Kernels (one launch each per batch, grid.z / grid.y = frame): blur7 -> 7 x resize_cubic -> fast_cells -> select ->
// orient_describe.  No host round trip: the selection that the reference does with std::nth_element runs on the
// device as an exact replay (select_exact.h).
#include "common.cuh"
#include <cuda.h>          // CUtensorMap + the cuTensorMapEncodeTiled prototype (resolved at run time through cudaGetDriverEntryPoint: no libcuda link)
#include "orb_math.h"
#include "orb_pattern.h"
#include "select_exact.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#define ORB_E 19            // EDGE_THRESHOLD
#define ORB_MAXL UCO_ORB_MAX_LEVELS
#define ORB_MAX_CELLS_PER_LEVEL 1024

namespace {

struct LevelDev {
    int w, h, pitch;        // level image size, bordered-buffer pitch
    unsigned off;           // byte offset of the bordered buffer inside a frame's pyramid block
    int n_desired;
    int rows, cols, nf_cell, n_cells, cell_begin;
    int sel_off, sel_cap;   // offset (entries) / capacity of this level's selected list inside a frame's sel block
    float scale;            // mvScaleFactor[level]
    int patch;              // (int)(PATCH_SIZE * scale)
    int vec_limit;          // x < vec_limit takes OpenCV's float SIMD path in the vertical resize pass
    int tab_x, tab_y;       // offsets into the resize tables (level >= 1)
};
struct CellDev {
    short level, valid;
    short x0, y0, w, h;     // cv::FAST ROI in level coordinates (iniX, iniY, hX, hY)
    int cand_off, cand_cap; // slot inside a frame's candidate block (entries)
};
struct PlanDev {
    int n_levels, n_cells_total, max_features, ini_th, min_th;
    unsigned frame_bytes;   // pyramid block per frame
    int cand_per_frame, sel_per_frame;
    LevelDev lv[ORB_MAXL];
};

__constant__ int8_t c_pattern[UCO_ORB_NPTS * 2];
__constant__ int c_umax[16];

// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {  // cv::borderInterpolate(BORDER_REFLECT_101)
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// store level pixel (x,y) and its reflections into the 19-px border (copyMakeBorder REFLECT_101 of the finished level)
__device__ __forceinline__ void store_with_border(uint8_t* buf, int pitch, int w, int h, int x, int y, uint8_t v) {
    int xs[3], ys[3], nx = 1, ny = 1;
    xs[0] = x;
    ys[0] = y;
    if (x >= 1 && x <= ORB_E) xs[nx++] = -x;
    if (x <= w - 2 && x >= w - 1 - ORB_E) xs[nx++] = 2 * (w - 1) - x;
    if (y >= 1 && y <= ORB_E) ys[ny++] = -y;
    if (y <= h - 2 && y >= h - 1 - ORB_E) ys[ny++] = 2 * (h - 1) - y;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) buf[(size_t)(ys[j] + ORB_E) * pitch + xs[i] + ORB_E] = v;
}

// ---- K1: 7x7 sigma=2 Gaussian blur, OpenCV 8U fixed-point path (kernel {18,34,48,56,48,34,18}/256 per axis, 8.8 after the
// horizontal pass, 16.16 after the vertical pass, + 0.5 and truncate) -> level 0 with border ------------------------------
#define BL_TW 64
#define BL_TH 16
__global__ void __launch_bounds__(256) blur7_kernel(const __grid_constant__ PlanDev c_plan, const uint8_t* __restrict__ in, size_t in_pitch, size_t in_frame,
                                                    uint8_t* __restrict__ pyr, int do_blur) {
    const LevelDev& L = c_plan.lv[0];
    const int w = L.w, h = L.h;
    const uint8_t* src = in + (size_t)blockIdx.z * in_frame;
    uint8_t* dst = pyr + (size_t)blockIdx.z * c_plan.frame_bytes + L.off;
    const int x0 = blockIdx.x * BL_TW, y0 = blockIdx.y * BL_TH;
    __shared__ uint8_t tin[BL_TH + 6][BL_TW + 8];
    __shared__ uint16_t hb[BL_TH + 6][BL_TW];
    for (int i = threadIdx.x; i < (BL_TH + 6) * (BL_TW + 6); i += 256) {
        int r = i / (BL_TW + 6), c = i % (BL_TW + 6);
        int yy = reflect101(y0 + r - 3, h), xx = reflect101(x0 + c - 3, w);
        tin[r][c] = src[(size_t)yy * in_pitch + xx];
    }
    __syncthreads();
    if (!do_blur) {
        for (int i = threadIdx.x; i < BL_TH * BL_TW; i += 256) {
            int r = i / BL_TW, c = i % BL_TW;
            if (x0 + c < w && y0 + r < h) store_with_border(dst, L.pitch, w, h, x0 + c, y0 + r, tin[r + 3][c + 3]);
        }
        return;
    }
    for (int i = threadIdx.x; i < (BL_TH + 6) * BL_TW; i += 256) {
        int r = i / BL_TW, c = i % BL_TW;
        const uint8_t* p = &tin[r][c];
        hb[r][c] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BL_TH * BL_TW; i += 256) {
        int r = i / BL_TW, c = i % BL_TW;
        if (x0 + c < w && y0 + r < h) {
            unsigned v = 18u * (hb[r][c] + hb[r + 6][c]) + 34u * (hb[r + 1][c] + hb[r + 5][c]) +
                         48u * (hb[r + 2][c] + hb[r + 4][c]) + 56u * hb[r + 3][c];
            store_with_border(dst, L.pitch, w, h, x0 + c, y0 + r, (uint8_t)((v + 32768u) >> 16));
        }
    }
}

// ---- K2: cv::resize INTER_CUBIC, 8U: integer horizontal pass (11-bit coefficients), vertical pass in float for
// x < vec_limit (OpenCV's SSE VResizeCubicVec_32s8u) and in 22-bit fixed point for the row tail ---------------------------
// Separable form: a CTA owns a 32 x 32 tile of the destination level.  Pass 1 runs the horizontal filter once per (source row the
// tile needs, destination column) into shared memory (the 4 source rows of vertically adjacent outputs overlap: ~1.3 filtered
// rows per output row instead of 4), pass 2 combines 4 of those rows per output pixel.  Same integers as the per-pixel form.
#define RS_T 32
#define RS_ROWS 72   // source rows a tile can need: 31 * scale + 4 with scale <= 2 (checked in orb_prepare)
__global__ void __launch_bounds__(256) resize_cubic_kernel(const __grid_constant__ PlanDev c_plan, uint8_t* __restrict__ pyr, int level,
                                                           const int* __restrict__ tab_ofs,
                                                           const short4* __restrict__ tab_coef) {
    const LevelDev& D = c_plan.lv[level];
    const LevelDev& S = c_plan.lv[level - 1];
    __shared__ int sr[RS_ROWS][RS_T + 1];
    __shared__ int s_sx[RS_T];
    __shared__ short4 s_a[RS_T];
    const int x0 = blockIdx.x * RS_T, y0 = blockIdx.y * RS_T, tid = threadIdx.x;
    uint8_t* frame = pyr + (size_t)blockIdx.z * c_plan.frame_bytes;
    const uint8_t* src = frame + S.off + (size_t)ORB_E * S.pitch + ORB_E;
    const int ny = min(RS_T, D.h - y0), nx = min(RS_T, D.w - x0);
    // clamped source row range of the tile
    const int r_lo = min(max(tab_ofs[D.tab_y + y0] - 1, 0), S.h - 1);
    const int r_hi = min(max(tab_ofs[D.tab_y + y0 + ny - 1] + 2, 0), S.h - 1);
    const int nr = r_hi - r_lo + 1;
    if (tid < RS_T) {
        const int dx = min(x0 + tid, D.w - 1);
        s_sx[tid] = tab_ofs[D.tab_x + dx];
        s_a[tid] = tab_coef[D.tab_x + dx];
    }
    __syncthreads();
    for (int i = tid; i < nr * RS_T; i += 256) {   // pass 1: horizontal, 11-bit coefficients
        const int r = i >> 5, x = i & 31;
        const int sx = s_sx[x];
        const short4 a = s_a[x];
        const uint8_t* p = src + (size_t)(r_lo + r) * S.pitch;
        const int c0 = min(max(sx - 1, 0), S.w - 1), c1 = min(max(sx, 0), S.w - 1), c2 = min(max(sx + 1, 0), S.w - 1),
                  c3 = min(max(sx + 2, 0), S.w - 1);
        sr[r][x] = p[c0] * a.x + p[c1] * a.y + p[c2] * a.z + p[c3] * a.w;
    }
    __syncthreads();
    for (int i = tid; i < RS_T * RS_T; i += 256) {  // pass 2: vertical (float for x < vec_limit as OpenCV's SIMD path, else 22-bit fixed point)
        const int y = i >> 5, x = i & 31;
        const int dx = x0 + x, dy = y0 + y;
        if (x >= nx || y >= ny) continue;
        const int sy = tab_ofs[D.tab_y + dy];
        const short4 b = tab_coef[D.tab_y + dy];
        int Sr[4];
#pragma unroll
        for (int k = 0; k < 4; k++) Sr[k] = sr[min(max(sy - 1 + k, 0), S.h - 1) - r_lo][x];
        int v;
        if (dx < D.vec_limit) {
            const float scale = 1.f / (2048.f * 2048.f);
            float b0 = __fmul_rn((float)b.x, scale), b1 = __fmul_rn((float)b.y, scale), b2 = __fmul_rn((float)b.z, scale),
                  b3 = __fmul_rn((float)b.w, scale);
            float t = __fmul_rn((float)Sr[3], b3);
            t = __fadd_rn(__fmul_rn((float)Sr[2], b2), t);
            t = __fadd_rn(__fmul_rn((float)Sr[1], b1), t);
            t = __fadd_rn(__fmul_rn((float)Sr[0], b0), t);
            v = __float2int_rn(t);
        } else {
            int acc = Sr[0] * b.x + Sr[1] * b.y + Sr[2] * b.z + Sr[3] * b.w;
            v = (acc + (1 << 21)) >> 22;
        }
        v = min(max(v, 0), 255);
        store_with_border(frame + D.off, D.pitch, D.w, D.h, dx, dy, (uint8_t)v);
    }
}

// ---- K1 / K2, strip forms (the ones launched; the per-pixel forms above stay as the A/B reference under UCO_ORB_PYRAMID_V1=1) ---------
// Both kernels treat the BORDERED level (w + 38 columns) as their output domain: destination word kd holds columns 4kd..4kd+3 of the
// bordered buffer = level pixels reflect101(4kd - 19 + j).  For the blur that is exact because the filter is symmetric and its own
// border rule is the same reflection (blur of the reflected extension at -x = blur at x, the same products in another order, integer
// adds); for the resize the coefficient tables are simply read at the reflected pixel.  So the left / right borders cost no special
// stores, every store is one aligned 32-bit word, and the top / bottom borders are the same word stored a second time.
__device__ __forceinline__ void store_word_rows(uint8_t* dst, int pitch, int h, int y, uint32_t word) {  // dst: column already applied, row 0 of the bordered buffer
    *(uint32_t*)(dst + (size_t)(y + ORB_E) * pitch) = word;
    if (y >= 1 && y <= ORB_E) *(uint32_t*)(dst + (size_t)(ORB_E - y) * pitch) = word;
    if (y <= h - 2 && y >= h - 1 - ORB_E) *(uint32_t*)(dst + (size_t)(2 * (h - 1) - y + ORB_E) * pitch) = word;
}

// Blur: a thread owns one destination word (4 pixels) and walks down BS_R rows.  Per input row it reads three aligned words of the
// staged tile (10 of their 12 bytes are the taps of its 4 pixels), runs the horizontal filter on packed u16x2 pairs (a lane never
// exceeds 256 * 255, so one IMAD serves two pixels) and keeps the last 7 rows of horizontal results in registers (the window rotates
// by unrolling 7 steps); the vertical filter reads that window.  The tile is staged once per CTA: interior words are copied as words,
// the reflected edge bytes (and every byte of an unaligned source) one by one.
#define BS_R 29                      // output rows per CTA
#define BS_STEPS (BS_R + 6)          // input rows per CTA = 5 x 7
#define BS_MAXW 256                  // destination words per CTA, at most
template <bool DO_BLUR>
__global__ void __launch_bounds__(256) blur7_strip_kernel(const __grid_constant__ PlanDev c_plan, const uint8_t* __restrict__ in, size_t in_pitch, size_t in_frame,
                                                          uint8_t* __restrict__ pyr, int wpt, int aligned) {
    const LevelDev& L = c_plan.lv[0];
    const int w = L.w, h = L.h;
    const int nwords = (w + 2 * ORB_E + 3) >> 2;
    const int k0 = blockIdx.x * wpt, nk = min(wpt, nwords - k0);
    const int y0 = blockIdx.y * BS_R;
    const int spw = wpt + 2;
    __shared__ uint32_t tile[BS_STEPS * (BS_MAXW + 2)];
    const uint8_t* src = in + (size_t)blockIdx.z * in_frame;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int vb0 = 4 * (k0 - 6);                                        // level pixel (virtual: before reflection) of byte 0 of a tile row
    int s_lo = nk + 2, s_hi = nk + 2;                                    // slots [s_lo, s_hi) are whole words inside the image
    if (aligned) {
        s_lo = min(max(0, 6 - k0), nk + 2);
        s_hi = max(s_lo, min(nk + 2, ((w - 4 - vb0) >> 2) + 1));
    }
    const int n_left = 4 * s_lo, n_edge = n_left + 4 * (nk + 2 - s_hi);
    for (int r = warp; r < BS_STEPS; r += nwarps) {
        const uint8_t* row = src + (size_t)reflect101(y0 - 3 + r, h) * in_pitch;
        uint32_t* trow = tile + r * spw;
        for (int s = s_lo + lane; s < s_hi; s += 32) trow[s] = *(const uint32_t*)(row + vb0 + 4 * s);
        for (int e = lane; e < n_edge; e += 32) {
            const int bi = e < n_left ? e : 4 * s_hi + (e - n_left);
            ((uint8_t*)trow)[bi] = row[reflect101(vb0 + bi, w)];
        }
    }
    __syncthreads();
    if (tid >= nk) return;
    uint8_t* dst = pyr + (size_t)blockIdx.z * c_plan.frame_bytes + L.off + 4 * (k0 + tid);
    const uint32_t* tp = tile + tid;
    if (!DO_BLUR) {
        for (int r = 0; r < BS_R && y0 + r < h; r++)
            store_word_rows(dst, L.pitch, h, y0 + r, __funnelshift_r(tp[(r + 3) * spw + 1], tp[(r + 3) * spw + 2], 8));
        return;
    }
    uint32_t hw[7][4];
#pragma unroll
    for (int u = 0; u < 7; u++)
#pragma unroll
        for (int j = 0; j < 4; j++) hw[u][j] = 0;
#pragma unroll 1
    for (int g = 0; g < BS_STEPS / 7; g++) {
#pragma unroll
        for (int u = 0; u < 7; u++) {
            const int step = g * 7 + u;
            const uint32_t W0 = tp[step * spw], W1 = tp[step * spw + 1], W2 = tp[step * spw + 2];
            // pair s = (byte s, byte s + 1) of the 12 staged bytes as u16x2; pixel j of the word is centred on byte 5 + j
            const uint32_t p2 = __byte_perm(W0, 0, 0x4342), p3 = __byte_perm(W0, W1, 0x0403) & 0x00ff00ffu, p4 = __byte_perm(W1, 0, 0x4140),
                           p5 = __byte_perm(W1, 0, 0x4241), p6 = __byte_perm(W1, 0, 0x4342), p7 = __byte_perm(W1, W2, 0x0403) & 0x00ff00ffu,
                           p8 = __byte_perm(W2, 0, 0x4140), p9 = __byte_perm(W2, 0, 0x4241), p10 = __byte_perm(W2, 0, 0x4342);
            const uint32_t A = 18u * (p2 + p8) + 34u * (p3 + p7) + 48u * (p4 + p6) + 56u * p5;
            const uint32_t B = 18u * (p4 + p10) + 34u * (p5 + p9) + 48u * (p6 + p8) + 56u * p7;
            hw[u][0] = A & 0xffffu; hw[u][1] = A >> 16; hw[u][2] = B & 0xffffu; hw[u][3] = B >> 16;
            const int yo = y0 + step - 6;                                // window = tile rows step-6 .. step; row step-6+k sits in slot (u+1+k) % 7
            if ((g > 0 || u == 6) && yo < h) {
                uint32_t v[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    v[j] = 18u * (hw[(u + 1) % 7][j] + hw[u][j]) + 34u * (hw[(u + 2) % 7][j] + hw[(u + 6) % 7][j]) +
                           48u * (hw[(u + 3) % 7][j] + hw[(u + 5) % 7][j]) + 56u * hw[(u + 4) % 7][j] + 32768u;   // < 2^24: the result is byte 2
                store_word_rows(dst, L.pitch, h, yo, __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410));
            }
        }
    }
}

__device__ __forceinline__ int reflect_once(int p, int len) {  // reflect101 for -len < p < 2 * len - 1 (one fold), branch free
    p = p < 0 ? -p : p;
    return p >= len ? 2 * (len - 1) - p : p;
}

// Resize: a CTA owns 64 bordered destination columns x 32 level rows.  The source pixels the tile needs are staged in shared memory as
// aligned words (the gathers of pass 1 then cost one LDS.U8 with an immediate row offset each, no 64-bit address arithmetic), the
// per-row source indices and float weights are made once per tile row.  Pass 1: a thread keeps ONE destination column (its 4 source
// offsets and 11-bit coefficients in registers) and filters the staged rows into shared memory; pass 2: a thread makes 4 adjacent
// pixels of one row from four 128-bit shared-memory reads.  Same integers / floats per pixel as resize_cubic_kernel.
#define RS2_TW 64
#define RS2_TH 32
#define RS2_SP 144                   // staged bytes per source row: 63 * 2 + 5 source columns at scale 2, + 3 of alignment, rounded to 16
__global__ void __launch_bounds__(256) resize_cubic_strip_kernel(const __grid_constant__ PlanDev c_plan, uint8_t* __restrict__ pyr, int level,
                                                                 const int* __restrict__ tab_ofs, const short4* __restrict__ tab_coef) {
    const LevelDev& D = c_plan.lv[level];
    const LevelDev& S = c_plan.lv[level - 1];
    __shared__ __align__(16) int sr[RS_ROWS][RS2_TW + 4];
    __shared__ __align__(16) uint8_t s_src[RS_ROWS][RS2_SP];
    __shared__ __align__(16) float4 s_bw[RS2_TH];
    __shared__ short4 s_b[RS2_TH];
    __shared__ uint32_t s_ri[RS2_TH];
    __shared__ __align__(4) uint8_t s_fix[RS2_TW];
    const int c0 = blockIdx.x * RS2_TW, y0 = blockIdx.y * RS2_TH, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wext = (D.w + 2 * ORB_E + 3) & ~3;                         // bordered columns, whole words
    uint8_t* frame = pyr + (size_t)blockIdx.z * c_plan.frame_bytes;
    const int ny = min(RS2_TH, D.h - y0);
    const int r_lo = min(max(tab_ofs[D.tab_y + y0] - 1, 0), S.h - 1);
    const int r_hi = min(max(tab_ofs[D.tab_y + y0 + ny - 1] + 2, 0), S.h - 1);
    const int nr = r_hi - r_lo + 1;
    // source columns of the tile: the level pixels of its columns are reflect101(c - 19), piecewise linear in c, and the source offset grows with the pixel
    const int cl = min(c0 + RS2_TW, wext) - 1;
    int pa = reflect_once(c0 - ORB_E, D.w), pb = reflect_once(cl - ORB_E, D.w);   // D.w >= 46: the 19 + 3 columns beyond an edge fold once
    int pmin = min(pa, pb), pmax = max(pa, pb);
    if (c0 <= ORB_E && ORB_E <= cl) pmin = 0;
    if (c0 <= D.w - 1 + ORB_E && D.w - 1 + ORB_E <= cl) pmax = D.w - 1;
    const int xlo = min(max(tab_ofs[D.tab_x + pmin] - 1, 0), S.w - 1), xhi = min(max(tab_ofs[D.tab_x + pmax] + 2, 0), S.w - 1);
    const int a0 = (xlo + ORB_E) & ~3, nw = min(((xhi + ORB_E - a0) >> 2) + 1, RS2_SP / 4);   // bordered source column of staged byte 0, staged words per row (<= RS2_SP / 4)
    if (tid < RS2_TH) {                                                  // per destination row: its 4 source rows (relative to r_lo) and weights
        const int dy = min(y0 + tid, D.h - 1);
        const int sy = tab_ofs[D.tab_y + dy];
        const short4 b = tab_coef[D.tab_y + dy];
        uint32_t ri = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) ri |= (uint32_t)(min(max(sy - 1 + k, 0), S.h - 1) - r_lo) << (8 * k);
        const float scale = 1.f / (2048.f * 2048.f);
        s_ri[tid] = ri;
        s_b[tid] = b;
        s_bw[tid] = make_float4(__fmul_rn((float)b.x, scale), __fmul_rn((float)b.y, scale), __fmul_rn((float)b.z, scale), __fmul_rn((float)b.w, scale));
    }
    {   // stage: a warp per source row, a lane per word (nw <= 36: lanes 0..3 take a second word at scale factors near 2); 4 rows in flight
        const uint8_t* sb = frame + S.off + (size_t)(r_lo + ORB_E) * S.pitch + a0 + 4 * lane;
        uint8_t* ss = &s_src[0][0] + 4 * lane;
        const unsigned sp = (unsigned)S.pitch;
        for (int r0 = warp; r0 < nr; r0 += 32) {
            uint32_t v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = (r0 + 8 * j < nr && lane < nw) ? *(const uint32_t*)(sb + (unsigned)(r0 + 8 * j) * sp) : 0u;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (r0 + 8 * j < nr && lane < nw) *(uint32_t*)(ss + (r0 + 8 * j) * RS2_SP) = v[j];
        }
        if (nw > 32 && lane + 32 < nw)
            for (int r = warp; r < nr; r += 8) *(uint32_t*)(ss + r * RS2_SP + 128) = *(const uint32_t*)(sb + (unsigned)r * sp + 128);
    }
    const int x = tid & (RS2_TW - 1), c = c0 + x;
    int q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    short4 a = make_short4(0, 0, 0, 0);
    if (c < wext) {
        const int px = reflect_once(c - ORB_E, D.w);
        const int sx = tab_ofs[D.tab_x + px];
        a = tab_coef[D.tab_x + px];
        const int sh = ORB_E - a0;
        q0 = min(max(sx - 1, 0), S.w - 1) + sh; q1 = min(max(sx, 0), S.w - 1) + sh; q2 = min(max(sx + 1, 0), S.w - 1) + sh; q3 = min(max(sx + 2, 0), S.w - 1) + sh;
        if (tid < RS2_TW) s_fix[x] = px >= D.vec_limit;
    } else if (tid < RS2_TW)
        s_fix[x] = 0;
    __syncthreads();
    if (c < wext) {                                                      // pass 1: horizontal, 11-bit coefficients
        const uint8_t* p = &s_src[tid >> 6][0];
        int* o = &sr[tid >> 6][x];
#pragma unroll 4
        for (int r = tid >> 6; r < nr; r += 4, p += 4 * RS2_SP, o += 4 * (RS2_TW + 4)) *o = p[q0] * a.x + p[q1] * a.y + p[q2] * a.z + p[q3] * a.w;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < (RS2_TW / 4) * RS2_TH / 256; it++) {           // pass 2: vertical, one destination word per thread and round
        const int i = tid + it * 256, y = i >> 4, gx = i & 15;
        if (y >= ny || c0 + 4 * gx >= wext) continue;
        const uint32_t ri = s_ri[y];
        const float4 bw = s_bw[y];
        int4 R[4];
#pragma unroll
        for (int k = 0; k < 4; k++) R[k] = *(const int4*)&sr[(ri >> (8 * k)) & 0xffu][4 * gx];
        const uint32_t fix4 = *(const uint32_t*)&s_fix[4 * gx];
        int v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int S0 = (&R[0].x)[j], S1 = (&R[1].x)[j], S2 = (&R[2].x)[j], S3 = (&R[3].x)[j];
            float t = __fmul_rn((float)S3, bw.w);
            t = __fadd_rn(__fmul_rn((float)S2, bw.z), t);
            t = __fadd_rn(__fmul_rn((float)S1, bw.y), t);
            t = __fadd_rn(__fmul_rn((float)S0, bw.x), t);
            v[j] = __float2int_rn(t);
        }
        if (fix4) {                                                       // columns past OpenCV's SIMD width: 22-bit fixed point
            const short4 b = s_b[y];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((fix4 >> (8 * j)) & 0xffu) {
                    const int acc = (&R[0].x)[j] * b.x + (&R[1].x)[j] * b.y + (&R[2].x)[j] * b.z + (&R[3].x)[j] * b.w;
                    v[j] = (acc + (1 << 21)) >> 22;
                }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = min(max(v[j], 0), 255);
        const uint32_t word = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
        store_word_rows(frame + D.off + c0 + 4 * gx, D.pitch, D.h, y0 + y, word);
    }
}

// ---- K3 + K4a: FAST-9/16 corner score on each grid cell's interior + 3x3 non-maximum suppression clipped to the cell
// (cv::FAST is called per cell ROI, ORBextractor.cpp:976-986, so neighbours in the adjacent cell never suppress) + ordered
// compaction (row-major, the order cv::FAST emits).
// The kernel is integer-ALU bound, so the work is arranged to keep every issued instruction useful:
//   1. decision, four horizontally adjacent pixels per thread in packed bytes: the 16 ring pixels arrive as 21 aligned 32-bit
//      shared-memory words + funnel shifts, "brighter than v+th" / "darker than v-th" are byte-wise unsigned compares whose
//      result sits in bit 7 of every byte, and the "9 contiguous of 16" test is 2 x 40 three-input LOPs on those words (a
//      rotation of the ring is a renaming of registers).  ~65 instructions per pixel, no divergence.
//   2. only the pixels that passed (a few percent) are appended to a list and scored afterwards by dense warps, on packed
//      (d, -d) 16-bit pairs so that one VIMNMX3.S16x2 advances the "min over the arc" of both polarities.
//   3. non-maximum suppression and the ordered output walk the corner bit mask (one thread per 32 pixels).
__device__ __forceinline__ unsigned gtu4(unsigned a, unsigned b) {  // bit 7 of every byte: a > b (unsigned); other bits undefined
    unsigned t = (a & 0x7f7f7f7fu) + (~b & 0x7f7f7f7fu);
    return (a & ~b) | (~(a ^ b) & t);
}
__device__ __forceinline__ unsigned arc9(const unsigned (&m)[16]) {  // bit 7 of byte j set iff 9 circularly contiguous m[k] have it
    unsigned r3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) r3[k] = m[k] & m[(k + 1) & 15] & m[(k + 2) & 15];
    unsigned any = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) any |= r3[k] & r3[(k + 3) & 15] & r3[(k + 6) & 15];
    return any;
}
// score = max(th, A, B) - 1 with A = max over the 16 arcs of 9 of min(d), B = max over arcs of min(-d) (cv cornerScore<16>),
// d = centre - ring.  x[k] packs (d, -d) as s16x2; arcs k..k+8 and k+1..k+9 share the min over k+1..k+8.
__device__ __forceinline__ int fast_score(const uint8_t* p, int pitch, int th) {
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * pitch];
    d[1] = v - p[3 * pitch + 1];
    d[2] = v - p[2 * pitch + 2];
    d[3] = v - p[pitch + 3];
    d[4] = v - p[3];
    d[5] = v - p[-pitch + 3];
    d[6] = v - p[-2 * pitch + 2];
    d[7] = v - p[-3 * pitch + 1];
    d[8] = v - p[-3 * pitch];
    d[9] = v - p[-3 * pitch - 1];
    d[10] = v - p[-2 * pitch - 2];
    d[11] = v - p[-pitch - 3];
    d[12] = v - p[-3];
    d[13] = v - p[pitch - 3];
    d[14] = v - p[2 * pitch - 2];
    d[15] = v - p[3 * pitch - 1];
    unsigned x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = __byte_perm((unsigned)d[k], (unsigned)(-d[k]), 0x5410);
    unsigned best = 0xff00ff00u;  // (-256, -256)
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        unsigned a = __vimin3_s16x2(x[(k + 1) & 15], x[(k + 2) & 15], x[(k + 3) & 15]);
        a = __vimin3_s16x2(a, x[(k + 4) & 15], x[(k + 5) & 15]);
        a = __vimin3_s16x2(a, x[(k + 6) & 15], x[(k + 7) & 15]);
        a = __vmins2(a, x[(k + 8) & 15]);
        best = __vimax3_s16x2(best, __vmins2(a, x[k]), __vmins2(a, x[(k + 9) & 15]));
    }
    const int A = (int)(short)(best & 0xffffu), B = (int)(short)(best >> 16);
    return max(max(th, A), B) - 1;
}

// shared-memory geometry of a cell (host and device agree through these)
__host__ __device__ inline int fc_roi_pitch(int cw) { return (cw + 8) & ~3; }            // ROI column c at byte c + 1; slack for the last group
__host__ __device__ inline int fc_sc_pitch(int iw) { return ((iw + 3) & ~3) + 8; }       // interior column c at byte c + 4, zero frame around
__host__ __device__ inline int fc_list_cap(int n) { return (max(512, n / 4) + 1) & ~1; }
__host__ __device__ inline int fc_smem_bytes(int cw, int ch) {
    const int iw = cw - 6, ih = ch - 6, n = iw * ih;
    return ch * fc_roi_pitch(cw) + (ih + 2) * fc_sc_pitch(iw) + 4 * ((n + 31) / 32) + 2 * fc_list_cap(n) + 16;
}

__global__ void __launch_bounds__(256, 5) fast_cells_kernel(const __grid_constant__ PlanDev c_plan, const uint8_t* __restrict__ pyr, const CellDev* __restrict__ cells,
                                                         uint32_t* __restrict__ cand, int* __restrict__ cand_cnt) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int ci = blockIdx.x, f = blockIdx.y;
    const CellDev C = cells[ci];
    int* cnt = cand_cnt + ((size_t)f * c_plan.n_cells_total + ci) * 2;
    const int iw = C.w - 6, ih = C.h - 6;
    if (!C.valid || iw <= 0 || ih <= 0) {
        if (threadIdx.x == 0) cnt[0] = cnt[1] = 0;
        return;
    }
    const LevelDev& L = c_plan.lv[C.level];
    const uint8_t* src = pyr + (size_t)f * c_plan.frame_bytes + L.off + (size_t)(C.y0 + ORB_E) * L.pitch + C.x0 + ORB_E;
    const int n = iw * ih, nw = (n + 31) >> 5;
    const int rp = fc_roi_pitch(C.w), sp = fc_sc_pitch(iw), lcap = fc_list_cap(n);
    uint8_t* roi = smem;                                       // C.h x rp
    uint8_t* sc = roi + (size_t)rp * C.h;                      // (ih + 2) x sp scores, zero where there is no corner
    unsigned* cmask = (unsigned*)(sc + (size_t)sp * (ih + 2)); // nw words: pixel i = r * iw + c passed the min_th segment test
    unsigned short* list = (unsigned short*)(cmask + nw);      // the same pixels, unordered, for the dense scoring pass
    __shared__ int nlist;
    __shared__ int wsum[8];
    __shared__ int running, running_ini;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {   // ROI -> shared memory as 32-bit words: the byte phase of a row start, (x0 + 19) & 3, is the same for every row (level offsets
        // and pitches are multiples of 64), so destination word k of a row (ROI columns 4k-1 .. 4k+2) is one funnel shift of two aligned
        // source words; 4 words in flight per thread.  Bytes past the ROI's last column are whatever the image holds there (inside the
        // bordered row: x0 + w + 8 <= level width + 38); the segment test masks those columns.
        const int nwr = rp >> 2, total = nwr * C.h;
        const unsigned wm = (unsigned)((0x100000000ull + nwr - 1) / (unsigned)nwr);   // i / nwr == umulhi(i, wm) for i < 2^16
        const int phi = (int)((size_t)src & 3);
        const uint8_t* base = src - phi;
        for (int i0 = tid; i0 < total; i0 += 4 * 256) {
            unsigned lo[4], hi[4];
            int so[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = i0 + 256 * j;
                const int r = total < 65536 ? (int)__umulhi((unsigned)i, wm) : i / nwr, k = i - r * nwr;
                const int o = 4 * k - 1 + phi;                              // byte offset from the aligned row start, >= -1
                const unsigned* q = (const unsigned*)(base + (size_t)r * L.pitch) + (o >> 2);
                so[j] = r * rp + 4 * k;
                lo[j] = hi[j] = 0;
                if (i < total) { lo[j] = q[0]; hi[j] = q[1]; }
                lo[j] = __funnelshift_r(lo[j], hi[j], 8 * (o & 3));
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (i0 + 256 * j < total) *(unsigned*)(roi + so[j]) = lo[j];
        }
    }
    for (int i = tid; i < ((sp * (ih + 2)) >> 2) + nw; i += 256) ((unsigned*)sc)[i] = 0;  // sc and cmask are contiguous
    if (tid == 0) nlist = running = running_ini = 0;
    __syncthreads();
    const int min_th = c_plan.min_th, ini_th = c_plan.ini_th;
    const unsigned iwm = (unsigned)((0x100000000ull + iw - 1) / (unsigned)iw);   // i / iw == umulhi(i, iwm) for i < 2^16 (iw = 1 would wrap the constant)
    const bool fastdiv = n < 65536 && iw > 1;
    // 1. segment test, 4 pixels per thread
    {
        const int gpr = (iw + 3) >> 2, G = gpr * ih, rpw = rp >> 2;
        const unsigned gm = (unsigned)((0x100000000ull + gpr - 1) / (unsigned)gpr);  // g / gpr == umulhi(g, gm) for g < 2^16 ...
        const bool small = G < 65536 && gpr > 1;
        const unsigned th4 = (unsigned)min_th * 0x01010101u;
        for (int g = tid; g < G; g += 256) {
            const int r = small ? (int)__umulhi((unsigned)g, gm) : g / gpr;
            const int c0 = (g - r * gpr) << 2;
            const unsigned* row = (const unsigned*)(roi + (size_t)(r + 3) * rp) + (c0 >> 2) + 1;
            unsigned ring[16];
            const unsigned v4 = row[0];
            {
                const unsigned* q = row + 3 * rpw;
                const unsigned a = q[-1], b = q[0], c = q[1];
                ring[15] = __funnelshift_r(a, b, 24); ring[0] = b; ring[1] = __funnelshift_r(b, c, 8);
            }
            {
                const unsigned* q = row + 2 * rpw;
                const unsigned a = q[-1], b = q[0], c = q[1];
                ring[14] = __funnelshift_r(a, b, 16); ring[2] = __funnelshift_r(b, c, 16);
            }
            {
                const unsigned* q = row + rpw;
                const unsigned a = q[-1], b = q[0], c = q[1];
                ring[13] = __funnelshift_r(a, b, 8); ring[3] = __funnelshift_r(b, c, 24);
            }
            {
                const unsigned a = row[-1], c = row[1];
                ring[12] = __funnelshift_r(a, v4, 8); ring[4] = __funnelshift_r(v4, c, 24);
            }
            {
                const unsigned* q = row - rpw;
                const unsigned a = q[-1], b = q[0], c = q[1];
                ring[11] = __funnelshift_r(a, b, 8); ring[5] = __funnelshift_r(b, c, 24);
            }
            {
                const unsigned* q = row - 2 * rpw;
                const unsigned a = q[-1], b = q[0], c = q[1];
                ring[10] = __funnelshift_r(a, b, 16); ring[6] = __funnelshift_r(b, c, 16);
            }
            {
                const unsigned* q = row - 3 * rpw;
                const unsigned a = q[-1], b = q[0], c = q[1];
                ring[9] = __funnelshift_r(a, b, 24); ring[8] = b; ring[7] = __funnelshift_r(b, c, 8);
            }
            const unsigned hi = __vaddus4(v4, th4), lo = __vsubus4(v4, th4);  // saturation = "no pixel can be beyond"
            unsigned m[16];
#pragma unroll
            for (int k = 0; k < 16; k++) m[k] = gtu4(ring[k], hi);
            unsigned res = arc9(m);
#pragma unroll
            for (int k = 0; k < 16; k++) m[k] = gtu4(lo, ring[k]);
            res = (res | arc9(m)) & 0x80808080u;
            while (res) {
                const int j = (__ffs(res) - 1) >> 3;
                res &= res - 1;
                if (c0 + j < iw) {
                    const int i = r * iw + c0 + j;
                    atomicOr(&cmask[i >> 5], 1u << (i & 31));
                    const int pos = atomicAdd(&nlist, 1);
                    if (pos < lcap && i < 65536) list[pos] = (unsigned short)i;
                    else sc[(r + 1) * sp + c0 + j + 4] = (uint8_t)fast_score(roi + (r + 3) * rp + c0 + j + 4, rp, min_th);
                }
            }
        }
    }
    __syncthreads();
    // 2. scores of the listed pixels, dense
    {
        const int nl = min(nlist, lcap);
        for (int e = tid; e < nl; e += 256) {
            const int i = list[e];
            const int r = fastdiv ? (int)__umulhi((unsigned)i, iwm) : i / iw, c = i - r * iw;
            sc[(r + 1) * sp + c + 4] = (uint8_t)fast_score(roi + (r + 3) * rp + c + 4, rp, min_th);
        }
    }
    __syncthreads();
    // 3. 3x3 non-maximum suppression (strictly greater than all 8 neighbours; the zero frame stands for the ROI edge) and
    //    ordered emission: a thread owns the 32 pixels of one mask word, a block scan orders the words
    uint32_t* out = cand + (size_t)f * c_plan.cand_per_frame + C.cand_off;
    for (int wbase = 0; wbase < nw; wbase += 256) {
        const int w = wbase + tid;
        unsigned bits = w < nw ? cmask[w] : 0u, keep = 0, ini = 0;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int i = (w << 5) + b;
            const int r = fastdiv ? (int)__umulhi((unsigned)i, iwm) : i / iw, c = i - r * iw;
            const uint8_t* q = sc + (r + 1) * sp + c + 4;
            const int s = q[0];
            const int nb = max(max(max(q[-sp - 1], q[-sp]), max(q[-sp + 1], q[-1])), max(max(q[1], q[sp - 1]), max(q[sp], q[sp + 1])));
            if (s > nb) {
                keep |= 1u << b;
                if (s >= ini_th) ini |= 1u << b;
            }
        }
        const int mine = __popc(keep) | (__popc(ini) << 16);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int before = running, tot = 0;
        for (int k = 0; k < 8; k++) {
            const int ws = wsum[k];
            if (k < warp) before += ws & 0xffff;
            tot += ws;
        }
        int pos = before + ((incl - mine) & 0xffff);
        while (keep) {
            const int b = __ffs(keep) - 1;
            keep &= keep - 1;
            const int i = (w << 5) + b;
            const int r = fastdiv ? (int)__umulhi((unsigned)i, iwm) : i / iw, c = i - r * iw;
            const uint32_t s = sc[(r + 1) * sp + c + 4];
            if (pos < C.cand_cap) out[pos] = (s << 24) | ((uint32_t)(C.y0 + 3 + r) << 12) | (uint32_t)(C.x0 + 3 + c);
            pos++;
        }
        __syncthreads();
        if (tid == 0) {
            running += tot & 0xffff;
            running_ini += tot >> 16;
        }
        __syncthreads();
    }
    if (tid == 0) {
        cnt[0] = min(running, C.cand_cap);
        cnt[1] = running_ini;
    }
}

// ---- K4b: per (frame, level) keypoint selection: threshold fallback, quota redistribution, per-cell retainBest, level-wide
// retainBest — ComputeKeyPoints_thread, ORBextractor.cpp:980-1073, replayed exactly ------------------------------------------
// Block-cooperative EXACT replay of libstdc++'s introselect (select_exact.h) on a list in shared memory.  The costly part of the serial
// form is __unguarded_partition: two pointers walk towards each other, each step a dependent load.  Its result is a pure function of the
// input, though: the left pointer stops at the elements whose score is <= the pivot's, in ascending position; the right pointer at the
// elements whose score is >= the pivot's, in descending position; stop k of the one is swapped with stop k of the other as long as the
// left one is still left of the right one, and no position takes part in two swaps.  So: positions of both kinds by a block scan, the
// number of swaps by a count of the (monotone) predicate, the swaps in parallel, the returned cut = the first stop of the left pointer
// after the last swap.  median-of-3, the depth limit (heap-select fallback, serial) and the final insertion sort stay as they are.
__device__ int par_unguarded_partition(uint32_t* v, int lo, int hi, uint32_t ps, int* Lpos, int* Rpos, int* sh) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int per = (hi - lo + nt - 1) / nt, b = min(hi, lo + tid * per), e = min(hi, b + per);
    int cl = 0, cr = 0;
    for (int i = b; i < e; i++) {
        const uint32_t sc = v[i] >> 24;
        cl += sc <= ps;
        cr += sc >= ps;
    }
    int il = cl, ir = cr;                                // inclusive scans: warp, then across warps
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, il, o), c = __shfl_up_sync(0xffffffffu, ir, o);
        if (lane >= o) { il += a; ir += c; }
    }
    if (lane == 31) { sh[warp] = il; sh[32 + warp] = ir; }
    __syncthreads();
    int bl = 0, br = 0, nL = 0, nR = 0;
    for (int w = 0; w < nw; w++) {
        if (w < warp) { bl += sh[w]; br += sh[32 + w]; }
        nL += sh[w]; nR += sh[32 + w];
    }
    int pl = bl + il - cl, pr = br + ir - cr;            // stops before this thread's chunk
    for (int i = b; i < e; i++) {
        const uint32_t sc = v[i] >> 24;
        if (sc <= ps) Lpos[pl++] = i;
        if (sc >= ps) Rpos[nR - 1 - pr++] = i;            // descending position
    }
    __syncthreads();
    const int m = min(nL, nR);
    int cnt = 0;
    for (int k = tid; k < m; k += nt) cnt += Lpos[k] < Rpos[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) sh[64 + warp] = cnt;
    __syncthreads();
    int ks = 0;
    for (int w = 0; w < nw; w++) ks += sh[64 + w];
    for (int k = tid; k < ks; k += nt) {
        const int a = Lpos[k], c = Rpos[k];
        const uint32_t t = v[a]; v[a] = v[c]; v[c] = t;
    }
    const int cut = min(ks < nL ? Lpos[ks] : hi, ks > 0 ? Rpos[ks - 1] : hi);
    __syncthreads();
    return cut;
}
// retainBest(v, n) + resize(n) by the whole block; v, Lpos, Rpos in shared memory (Lpos / Rpos: count entries each); returns the new count
__device__ int par_retain_best_truncate(uint32_t* v, int count, int n, int* Lpos, int* Rpos, int* sh) {
    if (!(n >= 0 && count > n)) return count;
    if (n == 0) return 0;
    const int nth = n - 1;
    int first = 0, last = count, depth = uco_sel::lg2(count) * 2;
    while (last - first > 3) {
        if (depth == 0) {                                // never seen on FAST scores; kept for exactness
            if (threadIdx.x == 0) {
                uco_sel::heap_select(v + first, v + nth + 1, v + last);
                uco_sel::iswap(v + first, v + nth);
            }
            __syncthreads();
            return n;
        }
        --depth;
        if (threadIdx.x == 0) uco_sel::move_median_to_first(v + first, v + first + 1, v + first + (last - first) / 2, v + last - 1);
        __syncthreads();
        const int cut = par_unguarded_partition(v, first + 1, last, v[first] >> 24, Lpos, Rpos, sh);
        if (cut <= nth) first = cut;
        else last = cut;
    }
    if (threadIdx.x == 0) uco_sel::insertion_sort(v + first, v + last);
    __syncthreads();
    return n;
}

// All of a level's lists live in SHARED memory while they are permuted: the per-cell retainBest (one thread per cell) and the level's
// final retainBest (one thread) are serial introselect replays whose every dependent access would otherwise be an L2 round trip
// (~0.3 us each: the kernel used to spend 146 us per clip that way).  Dynamic shared memory: cand_budget entries for the cells'
// candidate lists (packed back to back by a prefix sum of the counts) + list_entries for the concatenated level list; a level that
// exceeds either works in global memory as before.
__global__ void __launch_bounds__(256) select_kernel(const __grid_constant__ PlanDev c_plan, const CellDev* __restrict__ cells, uint32_t* __restrict__ cand,
                                                     const int* __restrict__ cand_cnt, uint32_t* __restrict__ sel,
                                                     int* __restrict__ sel_cnt, int* __restrict__ err_flag, int cand_budget, int list_entries) {
    const int level = blockIdx.x, f = blockIdx.y;
    const LevelDev& L = c_plan.lv[level];
    __shared__ int n_total[ORB_MAX_CELLS_PER_LEVEL];
    __shared__ short n_retain[ORB_MAX_CELLS_PER_LEVEL];
    __shared__ short th_used[ORB_MAX_CELLS_PER_LEVEL];
    __shared__ int offs[ORB_MAX_CELLS_PER_LEVEL];
    __shared__ int coff[ORB_MAX_CELLS_PER_LEVEL];        // start of the cell's staged candidate list
    __shared__ unsigned char valid_s[ORB_MAX_CELLS_PER_LEVEL];
    __shared__ int total_s, cand_total_s;
    extern __shared__ uint32_t s_dyn[];
    uint32_t* s_cand = s_dyn;
    uint32_t* s_list = s_dyn + cand_budget;
    const int nc = L.n_cells;
    const int* cnt = cand_cnt + ((size_t)f * c_plan.n_cells_total + L.cell_begin) * 2;
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        int n_all = cnt[c * 2], n_ini = cnt[c * 2 + 1];
        bool use_min = n_ini <= 3;                       // :982-987 (second FAST call with minThFAST)
        n_total[c] = use_min ? n_all : n_ini;
        th_used[c] = (short)(use_min ? c_plan.min_th : c_plan.ini_th);
        const bool v = cells[L.cell_begin + c].valid != 0;
        valid_s[c] = v;
        offs[c] = v ? n_all : 0;                         // entries to stage (scratch use of offs until the quotas are dealt)
    }
    __syncthreads();
    if (threadIdx.x == 32) {                             // staging offsets (a second thread: thread 0 deals the quotas meanwhile)
        int run = 0;
        for (int c = 0; c < nc; c++) { coff[c] = run; run += offs[c]; }
        cand_total_s = run;
    }
    __syncthreads();
    const bool staged_c = cand_total_s <= cand_budget;
    if (staged_c) {                                      // a warp copies a cell's list: coalesced
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int c = warp; c < nc; c += nw) {
            const uint32_t* src = cand + (size_t)f * c_plan.cand_per_frame + cells[L.cell_begin + c].cand_off;
            const int n = offs[c];
            for (int i = lane; i < n; i += 32) s_cand[coff[c] + i] = src[i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                              // :988-1039, serial exactly as the reference
        const int nf = L.nf_cell;
        int n_no_more = 0, n_dist = 0;
        // 0 = may take more, 1 = bNoMore.  Cells the reference skips with `continue` keep bNoMore=false, nTotal=0.
        for (int c = 0; c < nc; c++) {
            if (!valid_s[c]) {
                n_retain[c] = 0;
                offs[c] = 0;
                continue;
            }
            int nk = n_total[c];
            if (nk > nf) {
                n_retain[c] = (short)nf;
                offs[c] = 0;
            } else {
                n_retain[c] = (short)nk;
                n_dist += nf - nk;
                offs[c] = 1;
                n_no_more++;
            }
        }
        while (n_dist > 0 && n_no_more < nc) {
            int n_new = nf + (int)ceilf((float)n_dist / (float)(nc - n_no_more));
            n_dist = 0;
            for (int c = 0; c < nc; c++) {
                if (!offs[c]) {
                    if (n_total[c] > n_new) {
                        n_retain[c] = (short)n_new;
                    } else {
                        n_retain[c] = (short)n_total[c];
                        n_dist += n_new - n_total[c];
                        offs[c] = 1;
                        n_no_more++;
                    }
                }
            }
        }
    }
    __syncthreads();
    // per-cell: keep candidates with score >= the threshold in force (stable), then retainBest + truncate, in place
    auto cell_select = [&](uint32_t* lst, int c) {
        int n_all = valid_s[c] ? cnt[c * 2] : 0;
        int n = 0;
        const uint32_t th = (uint32_t)th_used[c];
        if (th == (uint32_t)c_plan.min_th) n = n_all;
        else
            for (int i = 0; i < n_all; i++) {
                uint32_t e = lst[i];
                if ((e >> 24) >= th) lst[n++] = e;
            }
        n_total[c] = uco_sel::retain_best_truncate(lst, n, n_retain[c]);
    };
    if (staged_c)
        for (int c = threadIdx.x; c < nc; c += blockDim.x) cell_select(s_cand + coff[c], c);   // shared-memory loads
    else
        for (int c = threadIdx.x; c < nc; c += blockDim.x) cell_select(cand + (size_t)f * c_plan.cand_per_frame + cells[L.cell_begin + c].cand_off, c);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c < nc; c++) {
            offs[c] = run;
            run += n_total[c];
        }
        if (run > L.sel_cap) {
            atomicExch(err_flag, 1);
            run = L.sel_cap;
        }
        total_s = run;
    }
    __syncthreads();
    uint32_t* out = sel + (size_t)f * c_plan.sel_per_frame + L.sel_off;
    const bool staged = staged_c && L.sel_cap <= list_entries && 2 * L.sel_cap <= cand_budget;
    __shared__ int scan_s[96];
    if (staged) {
        for (int c = 0; c < nc; c++) {                   // concatenation in cell (row-major) order, :1046-1066
            const uint32_t* lst = s_cand + coff[c];
            for (int i = threadIdx.x; i < n_total[c]; i += blockDim.x)
                if (offs[c] + i < L.sel_cap) s_list[offs[c] + i] = lst[i];
        }
        __syncthreads();
        // :1069-1073 by the whole block; the candidate area is dead and holds the stop positions
        const int n = par_retain_best_truncate(s_list, total_s, L.n_desired, (int*)s_cand, (int*)s_cand + L.sel_cap, scan_s);
        if (threadIdx.x == 0) sel_cnt[f * ORB_MAXL + level] = n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s_list[i];
        return;
    }
    for (int c = 0; c < nc; c++) {
        const uint32_t* lst = staged_c ? s_cand + coff[c] : cand + (size_t)f * c_plan.cand_per_frame + cells[L.cell_begin + c].cand_off;
        for (int i = threadIdx.x; i < n_total[c]; i += blockDim.x)
            if (offs[c] + i < L.sel_cap) out[offs[c] + i] = lst[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {                              // :1069-1073
        int n = total_s;
        if (n > L.n_desired) n = uco_sel::retain_best_truncate(out, n, L.n_desired);
        sel_cnt[f * ORB_MAXL + level] = n;
    }
}

// ---- TMA: one tiled tensor map per pyramid level over (bordered column, bordered row, frame) ---------------------------------
struct OrbTmaps {
    CUtensorMap m[ORB_MAXL];
};
#define ORB_WIN_COLS 64   // the box of a keypoint's window: the inner START has to be a multiple of 16 bytes (measured: scripts/microbench/tma_window.cu;
                          // any other x raises an illegal-instruction fault), so the box starts at x & ~15 and is 15 + 39 <= 64 columns wide
#define ORB_WIN_ROWS 39
__device__ __forceinline__ uint32_t orb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void orb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(orb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- K5 + K6: intensity-centroid orientation + steered rBRIEF, one warp per keypoint ------------------------------------
template <bool TMA>
__global__ void __launch_bounds__(256) orient_describe_kernel(const __grid_constant__ PlanDev c_plan, const CUtensorMap* __restrict__ tmaps,
                                                              const uint8_t* __restrict__ pyr, const uint32_t* __restrict__ sel,
                                                              const int* __restrict__ sel_cnt, uco_keypoint* __restrict__ kps,
                                                              uint8_t* __restrict__ desc, int* __restrict__ n_out,
                                                              int capacity) {
    __shared__ float pat[UCO_ORB_NPTS * 2];
    // the 39 x 39 window around the keypoint (orientation disc radius 15, rotated pattern radius <= 18.4 -> +-19 = EDGE_THRESHOLD),
    // staged per warp — by ONE 2-D TMA tile load (cp.async.bulk.tensor: the box is addressed by element coordinates, so the
    // unaligned window origin costs nothing and no lane computes an address), or by row-coalesced lane loads when the driver cannot
    // encode tensor maps: the 709 disc reads and 512 pattern gathers then hit shared memory, not L1 sectors
    __shared__ __align__(128) uint8_t patch_all[8][ORB_WIN_ROWS + 1][ORB_WIN_COLS];   // 2560 B per warp: a multiple of 128
    __shared__ __align__(8) uint64_t bars[8];
    if (TMA) {
        if ((threadIdx.x & 31) == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(orb_smem_u32(&bars[threadIdx.x >> 5])), "r"(1) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < UCO_ORB_NPTS * 2; i += blockDim.x) pat[(i & 31) * 32 + (i >> 5)] = (float)c_pattern[i];
    __syncthreads();
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int* cnt = sel_cnt + f * ORB_MAXL;
    int level = -1, first = 0, total = 0;
    for (int l = 0; l < c_plan.n_levels; l++) {
        int n = cnt[l];
        if (level < 0 && slot < total + n) {
            level = l;
            first = total;
        }
        total += n;
    }
    if (slot == 0 && lane == 0) n_out[f] = min(total, capacity);
    if (level < 0 || slot >= capacity) return;
    const LevelDev& L = c_plan.lv[level];
    const uint32_t e = sel[(size_t)f * c_plan.sel_per_frame + L.sel_off + (slot - first)];
    const int x = e & 0xfff, y = (e >> 12) & 0xfff, score = e >> 24;
    const uint8_t* center = pyr + (size_t)f * c_plan.frame_bytes + L.off + (size_t)(y + ORB_E) * L.pitch + x + ORB_E;
    uint8_t (*patch)[ORB_WIN_COLS] = patch_all[threadIdx.x >> 5];
    const int dx = TMA ? (x & 15) : 0;           // column of the window's first pixel inside the staged rows
    if (TMA) {
        uint64_t* bar = &bars[threadIdx.x >> 5];
        if (lane == 0) {   // box origin in bordered coordinates: (x + ORB_E - 19, y + ORB_E - 19) = (x, y), x rounded down to 16
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(orb_smem_u32(bar)), "r"(ORB_WIN_COLS * ORB_WIN_ROWS) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                             orb_smem_u32(&patch[0][0])),
                         "l"(tmaps + level), "r"(x & ~15), "r"(y), "r"(f), "r"(orb_smem_u32(bar))
                         : "memory");
        }
        orb_mbar_wait(bar, 0);
    } else {
        const uint8_t* p = center - 19 * L.pitch - 19;
#pragma unroll 3
        for (int r = 0; r < 39; r++, p += L.pitch) {
            patch[r][lane] = p[lane];
            if (lane < 7) patch[r][32 + lane] = p[32 + lane];
        }
        __syncwarp();
    }
    // IC_Angle: lane u+15 sums image COLUMN u of the radius-15 disc.  The disc is symmetric under transposition (the umax table is built
    // that way, ORBextractor.cpp:436-451: |u| <= umax[|v|] <=> |v| <= umax[|u|]) and the moments are integer sums, so the order is free;
    // with a lane per column the 32 lanes of every step read one row (8 consecutive words), where a lane per row had them 64 bytes
    // apart - 16 lanes on each of two banks (the kernel's top stall was the shared-memory queue)
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - 15;
        const int d = c_umax[u < 0 ? -u : u];
        const uint8_t* col = &patch[19][19 + dx + u];
        int sv = 0, s1 = 0;
#pragma unroll
        for (int v = -15; v <= 15; v++) {
            const int val = (v >= -d && v <= d) ? col[v * ORB_WIN_COLS] : 0;
            sv += v * val;
            s1 += val;
        }
        m10 = u * s1;
        m01 = sv;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = uco_math::fast_atan2_deg((float)m01, (float)m10);
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float a, b;
    uco_math::sincos_glibc(__fmul_rn(angle, factorPI), &b, &a);   // a = cos, b = sin
    // lane computes descriptor byte `lane` (tests 8*lane .. 8*lane+7)
    const float* pp = pat + lane;  // pat[j * 32 + lane] = coordinate j of this lane's 16 sample points
    unsigned byte = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        float x0 = pp[(t * 4) * 32], y0 = pp[(t * 4 + 1) * 32], x1 = pp[(t * 4 + 2) * 32], y1 = pp[(t * 4 + 3) * 32];
        int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        int t0 = patch[19 + r0][19 + dx + c0], t1 = patch[19 + r1][19 + dx + c1];
        byte |= (unsigned)(t0 < t1) << t;
    }
    const size_t o = (size_t)f * capacity + slot;
    desc[o * 32 + lane] = (uint8_t)byte;
    if (lane == 0) {
        uco_keypoint k;
        float fx = (float)x, fy = (float)y;
        if (level != 0) {  // :1229  kp.pt = (kp.pt + (0.5,0.5)) * scale
            fx = __fmul_rn(__fadd_rn(fx, 0.5f), L.scale);
            fy = __fmul_rn(__fadd_rn(fy, 0.5f), L.scale);
        }
        k.x = fx;
        k.y = fy;
        k.size = (float)L.patch;
        k.angle = angle;
        k.response = (float)score;
        k.octave = level;
        k.class_id = -1;
        kps[o] = k;
    }
}

}  // namespace

// =====================================================================================================================
// host side: plan (precalculateParams + grid geometry + resize tables), buffers, launches
// =====================================================================================================================
struct uco_orb_state {
    uco_orb_params prm{};
    int w = 0, h = 0, batch_cap = 0;
    PlanDev plan{};
    std::vector<CellDev> cells;
    int max_cell_smem = 0;
    uint8_t* d_pyr = nullptr;
    OrbTmaps tmaps{};              // per-level tensor maps over d_pyr (orient_describe stages keypoint windows with them)
    CUtensorMap* d_tmaps = nullptr; // ... in global memory: the kernel indexes them by level
    bool use_tma = false;
    bool pyramid_v1 = false;       // UCO_ORB_PYRAMID_V1=1: launch the per-pixel blur / resize kernels instead of the strip forms (A/B checks)
    uint8_t* d_in = nullptr;       // staging for host images
    size_t in_pitch = 0;
    CellDev* d_cells = nullptr;
    int* d_tab_ofs = nullptr;
    short4* d_tab_coef = nullptr;
    uint32_t* d_cand = nullptr;
    int* d_cand_cnt = nullptr;
    uint32_t* d_sel = nullptr;
    int* d_sel_cnt = nullptr;
    int* d_err = nullptr;
    uco_keypoint* d_kps = nullptr;
    uint8_t* d_desc = nullptr;
    int* d_nout = nullptr;
    int last_n = 0, last_max_features = 0;   // frames / row capacity of the last extraction (its results stay resident)
    int* h_nout = nullptr;         // pinned
    int* h_err = nullptr;          // pinned
    cudaEvent_t ev[6] = {};        // stage boundaries when ctx->profiling is set
    bool ev_valid = false;
};

static void orb_free_buffers(uco_orb_state* s) {
    cudaFree(s->d_pyr); cudaFree(s->d_tmaps); cudaFree(s->d_in); cudaFree(s->d_cells); cudaFree(s->d_tab_ofs); cudaFree(s->d_tab_coef);
    cudaFree(s->d_cand); cudaFree(s->d_cand_cnt); cudaFree(s->d_sel); cudaFree(s->d_sel_cnt); cudaFree(s->d_err);
    cudaFree(s->d_kps); cudaFree(s->d_desc); cudaFree(s->d_nout);
    if (s->h_nout) cudaFreeHost(s->h_nout);
    if (s->h_err) cudaFreeHost(s->h_err);
    for (auto& e : s->ev)
        if (e) { cudaEventDestroy(e); e = nullptr; }
    s->ev_valid = false;
    s->d_tmaps = nullptr; s->d_pyr = s->d_in = nullptr; s->d_cells = nullptr; s->d_tab_ofs = nullptr; s->d_tab_coef = nullptr;
    s->d_cand = nullptr; s->d_cand_cnt = nullptr; s->d_sel = nullptr; s->d_sel_cnt = nullptr; s->d_err = nullptr;
    s->d_kps = nullptr; s->d_desc = nullptr; s->d_nout = nullptr; s->h_nout = nullptr; s->h_err = nullptr;
}

void uco_orb_state_free(uco_b200_ctx* ctx) {
    if (!ctx->orb) return;
    orb_free_buffers(ctx->orb);
    delete ctx->orb;
    ctx->orb = nullptr;
}

static inline int cv_round_host(float v) { return (int)lrintf(v); }

// cv::resize INTER_CUBIC coefficient table for one axis (OpenCV imgproc resize.cpp: interpolateCubic, A = -0.75,
// INTER_RESIZE_COEF_SCALE = 2048, saturate_cast<short> = round half even)
static void cubic_table(int dst, int src, std::vector<int>& ofs, std::vector<short4>& coef) {
    const double inv_scale = (double)dst / src;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dst; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        const float A = -0.75f;
        float c[4];
        c[0] = ((A * (fx + 1) - 5 * A) * (fx + 1) + 8 * A) * (fx + 1) - 4 * A;
        c[1] = ((A + 2) * fx - (A + 3)) * fx * fx + 1;
        c[2] = ((A + 2) * (1 - fx) - (A + 3)) * (1 - fx) * (1 - fx) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
        short4 s;
        short* sp = &s.x;
        for (int k = 0; k < 4; k++) {
            long r = lrintf(c[k] * 2048.f);
            sp[k] = (short)std::min(32767L, std::max(-32768L, r));
        }
        ofs.push_back(sx);
        coef.push_back(s);
    }
}

static int orb_prepare(uco_b200_ctx* ctx, int w, int h, const uco_orb_params* prm, int batch) {
    uco_orb_state* s = ctx->orb;
    if (!s) s = ctx->orb = new uco_orb_state();
    const bool same_plan = s->w == w && s->h == h && memcmp(&s->prm, prm, sizeof *prm) == 0;
    if (same_plan && batch <= s->batch_cap) return UCO_OK;
    if (w > 4096 || h > 4096 || w < 64 || h < 64) return uco_fail(ctx, UCO_E_INVALID, "orb: image size %dx%d unsupported", w, h);
    if (prm->n_levels < 1 || prm->n_levels > ORB_MAXL || prm->max_features < 1 || !(prm->scale_factor > 1.f))
        return uco_fail(ctx, UCO_E_INVALID, "orb: bad parameters");
    if (prm->scale_factor > 2.f) return uco_fail(ctx, UCO_E_INVALID, "orb: scale factor %g above 2 is not supported", prm->scale_factor);
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    orb_free_buffers(s);
    s->batch_cap = 0;
    s->prm = *prm;
    s->w = w;
    s->h = h;
    s->cells.clear();
    s->max_cell_smem = 0;
    s->pyramid_v1 = getenv("UCO_ORB_PYRAMID_V1") != nullptr;
    PlanDev& P = s->plan;
    memset(&P, 0, sizeof P);
    const int NL = prm->n_levels;
    P.n_levels = NL;
    P.max_features = prm->max_features;
    P.ini_th = prm->ini_th_fast;
    P.min_th = prm->min_th_fast;
    // precalculateParams, ORBextractor.cpp:466-514
    float sf[ORB_MAXL], inv[ORB_MAXL];
    sf[0] = 1.0f;
    for (int i = 1; i < NL; i++) sf[i] = sf[i - 1] * prm->scale_factor;
    for (int i = 0; i < NL; i++) inv[i] = 1.0f / sf[i];
    {
        float factor = 1.0f / prm->scale_factor;
        float nd = prm->max_features * (1 - factor) / (1 - (float)pow((double)factor, (double)NL));
        int sum = 0;
        for (int l = 0; l < NL - 1; l++) {
            P.lv[l].n_desired = cv_round_host(nd);
            sum += P.lv[l].n_desired;
            nd *= factor;
        }
        P.lv[NL - 1].n_desired = std::max(prm->max_features - sum, 0);
    }
    std::vector<int> tab_ofs;
    std::vector<short4> tab_coef;
    unsigned off = 0;
    int cand_off = 0, sel_off = 0;
    const float image_ratio = (float)w / h;
    for (int l = 0; l < NL; l++) {
        LevelDev& L = P.lv[l];
        L.w = cv_round_host((float)w * inv[l]);   // ComputePyramid :1370
        L.h = cv_round_host((float)h * inv[l]);
        if (L.w < 2 * ORB_E + 8 || L.h < 2 * ORB_E + 8)
            return uco_fail(ctx, UCO_E_INVALID, "orb: level %d is %dx%d, too small for the 19-px border", l, L.w, L.h);
        L.pitch = (L.w + 2 * ORB_E + 63) & ~63;
        L.off = off;
        off += (unsigned)L.pitch * (L.h + 2 * ORB_E);
        off = (off + 255u) & ~255u;
        L.scale = sf[l];
        L.patch = (int)(31 * sf[l]);              // const int scaledPatchSize = PATCH_SIZE*mvScaleFactor[level]
        L.vec_limit = (L.w / 8) * 8;
        if (l > 0) {
            L.tab_x = (int)tab_ofs.size();
            cubic_table(L.w, P.lv[l - 1].w, tab_ofs, tab_coef);
            L.tab_y = (int)tab_ofs.size();
            cubic_table(L.h, P.lv[l - 1].h, tab_ofs, tab_coef);
        }
        // grid geometry, ComputeKeyPoints_thread :900-976
        const int nd = L.n_desired;
        const int level_cols = (int)sqrtf((float)nd / (5 * image_ratio));
        const int level_rows = (int)(image_ratio * level_cols);
        if (level_cols <= 0 || level_rows <= 0)
            return uco_fail(ctx, UCO_E_INVALID, "orb: level %d has an empty cell grid (max_features too small)", l);
        if (level_cols * level_rows > ORB_MAX_CELLS_PER_LEVEL)
            return uco_fail(ctx, UCO_E_INVALID, "orb: level %d has too many cells", l);
        const int min_b = ORB_E, max_bx = L.w - ORB_E, max_by = L.h - ORB_E;
        const int W = max_bx - min_b, H = max_by - min_b;
        const int cell_w = (int)ceilf((float)W / level_cols), cell_h = (int)ceilf((float)H / level_rows);
        L.rows = level_rows;
        L.cols = level_cols;
        L.n_cells = level_rows * level_cols;
        L.nf_cell = (int)ceilf((float)nd / L.n_cells);
        L.cell_begin = (int)s->cells.size();
        float hY = cell_h + 6;
        std::vector<int> ini_x(level_cols);
        for (int i = 0; i < level_rows; i++) {
            const float iniY = min_b + i * cell_h - 3;
            bool row_valid = true;
            if (i == level_rows - 1) {
                hY = max_by + 3 - iniY;
                if (hY <= 0) row_valid = false;
            }
            float hX = cell_w + 6;
            for (int j = 0; j < level_cols; j++) {
                CellDev C{};
                C.level = (short)l;
                float iniX;
                if (i == 0) {
                    iniX = min_b + j * cell_w - 3;
                    ini_x[j] = (int)iniX;
                } else
                    iniX = ini_x[j];
                bool valid = row_valid;
                if (valid && j == level_cols - 1) {
                    hX = max_bx + 3 - iniX;
                    if (hX <= 0) valid = false;
                }
                if (valid && ((int)iniY + (int)hY > L.h || (int)iniX + (int)hX > L.w))
                    return uco_fail(ctx, UCO_E_INVALID, "orb: level %d cell (%d,%d) leaves the image (the reference's "
                                    "cv::Mat::rowRange would assert)", l, i, j);
                C.valid = valid;
                C.x0 = (short)iniX;
                C.y0 = (short)iniY;
                C.w = (short)(valid ? hX : 0);
                C.h = (short)(valid ? hY : 0);
                int iw = std::max(C.w - 6, 0), ih = std::max(C.h - 6, 0);
                C.cand_off = cand_off;
                C.cand_cap = ((iw + 1) / 2) * ((ih + 1) / 2) + 1;
                cand_off += C.cand_cap;
                int smem = valid && iw > 0 && ih > 0 ? fc_smem_bytes(C.w, C.h) : 0;
                s->max_cell_smem = std::max(s->max_cell_smem, smem);
                s->cells.push_back(C);
            }
        }
        L.sel_off = sel_off;
        L.sel_cap = 2 * nd + 2 * L.n_cells + 64;
        sel_off += L.sel_cap;
    }
    if (s->max_cell_smem > 200 * 1024)
        return uco_fail(ctx, UCO_E_INVALID, "orb: a grid cell needs %d bytes of shared memory (max_features too small for "
                        "this image size)", s->max_cell_smem);
    P.frame_bytes = off;
    P.n_cells_total = (int)s->cells.size();
    P.cand_per_frame = cand_off;
    P.sel_per_frame = sel_off;

    const int B = std::max(batch, 1);
    s->in_pitch = (size_t)((w + 255) & ~255);
    UCO_CUDA(ctx, cudaMalloc(&s->d_pyr, (size_t)P.frame_bytes * B));
    {   // tensor maps: level l as a (pitch, h + 2 * ORB_E, frames) u8 tensor inside the pyramid blocks, box = one keypoint window
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        s->use_tma = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess &&
                     P.frame_bytes % 16 == 0 && !getenv("UCO_ORB_NO_TMA");
        for (int l = 0; s->use_tma && l < P.n_levels; l++) {
            const LevelDev& L = P.lv[l];
            const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(L.h + 2 * ORB_E), (cuuint64_t)B};
            const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)P.frame_bytes};
            const cuuint32_t box[3] = {ORB_WIN_COLS, ORB_WIN_ROWS, 1}, estr[3] = {1, 1, 1};
            if (L.off % 16 || L.pitch % 16 ||
                ((EncodeFn)fn)(&s->tmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, s->d_pyr + L.off, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                s->use_tma = false;
        }
        UCO_CUDA(ctx, cudaMalloc(&s->d_tmaps, sizeof(OrbTmaps)));
        UCO_CUDA(ctx, cudaMemcpy(s->d_tmaps, &s->tmaps, sizeof(OrbTmaps), cudaMemcpyHostToDevice));
    }
    UCO_CUDA(ctx, cudaMalloc(&s->d_in, s->in_pitch * h * B));
    UCO_CUDA(ctx, cudaMalloc(&s->d_cells, sizeof(CellDev) * s->cells.size()));
    UCO_CUDA(ctx, cudaMalloc(&s->d_tab_ofs, sizeof(int) * std::max<size_t>(tab_ofs.size(), 1)));
    UCO_CUDA(ctx, cudaMalloc(&s->d_tab_coef, sizeof(short4) * std::max<size_t>(tab_coef.size(), 1)));
    UCO_CUDA(ctx, cudaMalloc(&s->d_cand, sizeof(uint32_t) * (size_t)cand_off * B));
    UCO_CUDA(ctx, cudaMalloc(&s->d_cand_cnt, sizeof(int) * 2 * s->cells.size() * B));
    UCO_CUDA(ctx, cudaMalloc(&s->d_sel, sizeof(uint32_t) * (size_t)sel_off * B));
    UCO_CUDA(ctx, cudaMalloc(&s->d_sel_cnt, sizeof(int) * ORB_MAXL * B));
    UCO_CUDA(ctx, cudaMalloc(&s->d_err, sizeof(int)));
    UCO_CUDA(ctx, cudaMalloc(&s->d_kps, sizeof(uco_keypoint) * (size_t)prm->max_features * B));
    UCO_CUDA(ctx, cudaMalloc(&s->d_desc, (size_t)32 * prm->max_features * B));
    UCO_CUDA(ctx, cudaMalloc(&s->d_nout, sizeof(int) * B));
    UCO_CUDA(ctx, cudaMallocHost(&s->h_nout, sizeof(int) * B));
    UCO_CUDA(ctx, cudaMallocHost(&s->h_err, sizeof(int)));
    UCO_CUDA(ctx, cudaMemsetAsync(s->d_err, 0, sizeof(int), ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(s->d_cells, s->cells.data(), sizeof(CellDev) * s->cells.size(), cudaMemcpyHostToDevice,
                                  ctx->stream));
    if (!tab_ofs.empty()) {
        UCO_CUDA(ctx, cudaMemcpyAsync(s->d_tab_ofs, tab_ofs.data(), sizeof(int) * tab_ofs.size(), cudaMemcpyHostToDevice,
                                      ctx->stream));
        UCO_CUDA(ctx, cudaMemcpyAsync(s->d_tab_coef, tab_coef.data(), sizeof(short4) * tab_coef.size(),
                                      cudaMemcpyHostToDevice, ctx->stream));
    }
    // umax table, ORBextractor.cpp:436-451
    int umax[16];
    {
        const int HP = 15;
        int v, v0, vmax = (int)floorf(HP * sqrtf(2.f) / 2 + 1);
        int vmin = (int)ceilf(HP * sqrtf(2.f) / 2);
        const double hp2 = HP * HP;
        for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt(hp2 - v * v));
        for (v = HP, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyToSymbol(c_pattern, uco_orb_pattern_xy, sizeof uco_orb_pattern_xy));
    UCO_CUDA(ctx, cudaMemcpyToSymbol(c_umax, umax, sizeof umax));
    if (s->max_cell_smem > 40 * 1024)
        UCO_CUDA(ctx, cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->max_cell_smem));
    s->batch_cap = B;
    return UCO_OK;
}

// images already on the device (d_in layout: n frames of h rows, pitch in_pitch); everything asynchronous on ctx->stream
static int orb_run_dev(uco_b200_ctx* ctx, const uint8_t* in_dev, size_t in_pitch, size_t in_frame, int n,
                       uco_keypoint* kps_dev, uint8_t* desc_dev, int* nout_dev) {
    uco_orb_state* s = ctx->orb;
    const PlanDev& P = s->plan;
    cudaStream_t st = ctx->stream;
    const bool prof = ctx->profiling != 0;
    if (prof && !s->ev[0])
        for (auto& e : s->ev) UCO_CUDA(ctx, cudaEventCreate(&e));
    if (prof) cudaEventRecord(s->ev[0], st);
    if (s->pyramid_v1) {   // per-pixel forms (A/B reference)
        dim3 g0((P.lv[0].w + BL_TW - 1) / BL_TW, (P.lv[0].h + BL_TH - 1) / BL_TH, n);
        blur7_kernel<<<g0, 256, 0, st>>>(P, in_dev, in_pitch, in_frame, s->d_pyr, s->prm.blur_first);
    } else {
        const int nwords = (P.lv[0].w + 2 * ORB_E + 3) >> 2, ntx = (nwords + BS_MAXW - 1) / BS_MAXW, wpt = (nwords + ntx - 1) / ntx;
        const int aligned = (((uintptr_t)in_dev | in_pitch | in_frame) & 3) == 0;
        dim3 g0(ntx, (P.lv[0].h + BS_R - 1) / BS_R, n);
        if (s->prm.blur_first) blur7_strip_kernel<true><<<g0, (wpt + 31) & ~31, 0, st>>>(P, in_dev, in_pitch, in_frame, s->d_pyr, wpt, aligned);
        else blur7_strip_kernel<false><<<g0, (wpt + 31) & ~31, 0, st>>>(P, in_dev, in_pitch, in_frame, s->d_pyr, wpt, aligned);
    }
    UCO_LAUNCH_CHECK(ctx);
    if (prof) cudaEventRecord(s->ev[1], st);
    for (int l = 1; l < P.n_levels; l++) {
        if (s->pyramid_v1) {
            dim3 g((P.lv[l].w + RS_T - 1) / RS_T, (P.lv[l].h + RS_T - 1) / RS_T, n);
            resize_cubic_kernel<<<g, 256, 0, st>>>(P, s->d_pyr, l, s->d_tab_ofs, s->d_tab_coef);
        } else {
            dim3 g((P.lv[l].w + 2 * ORB_E + RS2_TW - 1) / RS2_TW, (P.lv[l].h + RS2_TH - 1) / RS2_TH, n);
            resize_cubic_strip_kernel<<<g, 256, 0, st>>>(P, s->d_pyr, l, s->d_tab_ofs, s->d_tab_coef);
        }
        UCO_LAUNCH_CHECK(ctx);
    }
    if (prof) cudaEventRecord(s->ev[2], st);
    fast_cells_kernel<<<dim3(P.n_cells_total, n), 256, s->max_cell_smem, st>>>(P, s->d_pyr, s->d_cells, s->d_cand, s->d_cand_cnt);
    UCO_LAUNCH_CHECK(ctx);
    if (prof) cudaEventRecord(s->ev[3], st);
    int sel_entries = 0;   // largest level list; staged in shared memory next to the cells' candidate lists (static shared memory holds the per-cell tables)
    for (int l = 0; l < P.n_levels; l++) sel_entries = std::max(sel_entries, P.lv[l].sel_cap);
    if (sel_entries > 8192) sel_entries = 0;
    const int cand_budget = 10240;   // 40 KB (three CTAs per SM): a 640x480 level 0 holds a few thousand candidates; a level above the budget works in global memory
    const size_t sel_smem = 4 * (size_t)(cand_budget + sel_entries);
    {
        static size_t configured = 0;   // grow-only attribute (same value from every thread)
        if (sel_smem > configured) {
            UCO_CUDA(ctx, cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
            configured = sel_smem;
        }
    }
    select_kernel<<<dim3(P.n_levels, n), 256, sel_smem, st>>>(P, s->d_cells, s->d_cand, s->d_cand_cnt, s->d_sel, s->d_sel_cnt, s->d_err, cand_budget, sel_entries);
    UCO_LAUNCH_CHECK(ctx);
    if (prof) cudaEventRecord(s->ev[4], st);
    if (s->use_tma)
        orient_describe_kernel<true><<<dim3((P.max_features + 7) / 8, n), 256, 0, st>>>(P, s->d_tmaps, s->d_pyr, s->d_sel, s->d_sel_cnt, kps_dev,
                                                                                      desc_dev, nout_dev, P.max_features);
    else
        orient_describe_kernel<false><<<dim3((P.max_features + 7) / 8, n), 256, 0, st>>>(P, s->d_tmaps, s->d_pyr, s->d_sel, s->d_sel_cnt, kps_dev,
                                                                                       desc_dev, nout_dev, P.max_features);
    UCO_LAUNCH_CHECK(ctx);
    if (prof) {
        cudaEventRecord(s->ev[5], st);
        s->ev_valid = true;
    }
    return UCO_OK;
}

extern "C" {

void uco_b200_orb_default_params(uco_orb_params* p) {
    p->max_features = 4000;   // ucoslamtypes.cpp:31-40 / FeatParams defaults, feature2dserializable.h:34-39
    p->n_levels = 8;
    p->scale_factor = 1.2f;
    p->ini_th_fast = 20;      // ORBextractor.cpp:478-479
    p->min_th_fast = 7;
    p->blur_first = 1;        // ORBextractor.h: _doGaussianBlurAtFirst = true
}

int uco_b200_orb_extract_batch_dev(uco_b200_ctx* ctx, const uint8_t* imgs_dev, int n_imgs, int w, int h, size_t pitch,
                                   size_t frame_stride, const uco_orb_params* prm, uco_keypoint* kps_dev,
                                   uint8_t* desc_dev, int* n_out_dev) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (!prm || n_imgs < 0) return uco_fail(ctx, UCO_E_INVALID, "orb: bad arguments");
    if (n_imgs == 0) return UCO_OK;
    if (!imgs_dev || !kps_dev || !desc_dev || !n_out_dev) return uco_fail(ctx, UCO_E_INVALID, "orb: null pointer");
    if (pitch < (size_t)w) return uco_fail(ctx, UCO_E_INVALID, "orb: pitch below width");
    int rc = orb_prepare(ctx, w, h, prm, n_imgs);
    if (rc != UCO_OK) return rc;
    return orb_run_dev(ctx, imgs_dev, pitch, frame_stride, n_imgs, kps_dev, desc_dev, n_out_dev);
}

int uco_b200_orb_extract_batch(uco_b200_ctx* ctx, const uint8_t* const* imgs, int n_imgs, int w, int h, size_t stride,
                               const uco_orb_params* prm, uco_keypoint* kps, uint8_t* desc, int capacity, int* n_out) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (!prm || n_imgs < 0) return uco_fail(ctx, UCO_E_INVALID, "orb: bad arguments");
    if (n_imgs == 0) return UCO_OK;
    if (!imgs || !kps || !desc || !n_out) return uco_fail(ctx, UCO_E_INVALID, "orb: null pointer");
    if (stride < (size_t)w) return uco_fail(ctx, UCO_E_INVALID, "orb: stride below width");
    if (capacity < prm->max_features)
        return uco_fail(ctx, UCO_E_CAPACITY, "orb: output capacity %d below max_features %d", capacity, prm->max_features);
    int rc = orb_prepare(ctx, w, h, prm, n_imgs);
    if (rc != UCO_OK) return rc;
    uco_orb_state* s = ctx->orb;
    for (int i = 0; i < n_imgs; i++) {
        if (!imgs[i]) return uco_fail(ctx, UCO_E_INVALID, "orb: null image %d", i);
        UCO_CUDA(ctx, cudaMemcpy2DAsync(s->d_in + (size_t)i * s->in_pitch * h, s->in_pitch, imgs[i], stride, w, h,
                                        cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = orb_run_dev(ctx, s->d_in, s->in_pitch, s->in_pitch * h, n_imgs, s->d_kps, s->d_desc, s->d_nout);
    if (rc != UCO_OK) return rc;
    s->last_n = n_imgs; s->last_max_features = prm->max_features;
    const int mf = prm->max_features;
    UCO_CUDA(ctx, cudaMemcpyAsync(s->h_nout, s->d_nout, sizeof(int) * n_imgs, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(s->h_err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpy2DAsync(kps, sizeof(uco_keypoint) * (size_t)capacity, s->d_kps, sizeof(uco_keypoint) * (size_t)mf,
                                    sizeof(uco_keypoint) * (size_t)mf, n_imgs, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpy2DAsync(desc, (size_t)32 * capacity, s->d_desc, (size_t)32 * mf, (size_t)32 * mf, n_imgs,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*s->h_err) return uco_fail(ctx, UCO_E_CAPACITY, "orb: internal selection list overflow");
    for (int i = 0; i < n_imgs; i++) n_out[i] = s->h_nout[i];
    return UCO_OK;
}

// internal (match.cu): the frames of the LAST extraction call, still resident in the context's buffers
int uco_orb_resident(uco_b200_ctx* ctx, const uco_keypoint** d_kps, const uint8_t** d_desc, const int** d_nout, int* max_features, int* n_frames) {
    if (!ctx->orb || !ctx->orb->batch_cap || !ctx->orb->last_n) return uco_fail(ctx, UCO_E_INVALID, "no extraction result is resident in this context");
    *d_kps = ctx->orb->d_kps; *d_desc = ctx->orb->d_desc; *d_nout = ctx->orb->d_nout;
    *max_features = ctx->orb->last_max_features; *n_frames = ctx->orb->last_n;
    return UCO_OK;
}

// internal (track.cu): host frames in, extraction on the device, results LEFT on the device: keypoints of frame i at
// (*d_kps) + i * max_features, descriptors at (*d_desc) + 32 * i * max_features, counts at (*d_nout)[i]; the extractor's error word at *d_err
int uco_orb_extract_keep_dev(uco_b200_ctx* ctx, const uint8_t* const* imgs, int n_imgs, int w, int h, size_t stride, const uco_orb_params* prm,
                             uco_keypoint** d_kps, uint8_t** d_desc, int** d_nout, int** d_err) {
    if (!prm || n_imgs <= 0 || !imgs) return uco_fail(ctx, UCO_E_INVALID, "orb: bad arguments");
    if (stride < (size_t)w) return uco_fail(ctx, UCO_E_INVALID, "orb: stride below width");
    int rc = orb_prepare(ctx, w, h, prm, n_imgs);
    if (rc != UCO_OK) return rc;
    uco_orb_state* s = ctx->orb;
    for (int i = 0; i < n_imgs; i++) {
        if (!imgs[i]) return uco_fail(ctx, UCO_E_INVALID, "orb: null image %d", i);
        UCO_CUDA(ctx, cudaMemcpy2DAsync(s->d_in + (size_t)i * s->in_pitch * h, s->in_pitch, imgs[i], stride, w, h,
                                        cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = orb_run_dev(ctx, s->d_in, s->in_pitch, s->in_pitch * h, n_imgs, s->d_kps, s->d_desc, s->d_nout);
    if (rc != UCO_OK) return rc;
    s->last_n = n_imgs; s->last_max_features = prm->max_features;
    *d_kps = s->d_kps; *d_desc = s->d_desc; *d_nout = s->d_nout; *d_err = s->d_err;
    return UCO_OK;
}

int uco_b200_orb_extract(uco_b200_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, const uco_orb_params* prm,
                         uco_keypoint* kps, uint8_t* desc, int capacity, int* n_out) {
    const uint8_t* one[1] = {img};
    return uco_b200_orb_extract_batch(ctx, one, 1, w, h, stride, prm, kps, desc, capacity, n_out);
}

// device time (ms) of the stages of the LAST extract call, measured with CUDA events on the context stream when profiling
// is enabled: out[0..4] = blur, resize chain, fast_cells, select, orient_describe.  Synchronises the stream.
int uco_b200_orb_last_stage_ms(uco_b200_ctx* ctx, float* out) {
    if (!ctx || !ctx->orb || !ctx->orb->ev_valid) return UCO_E_INVALID;
    UCO_CUDA(ctx, cudaEventSynchronize(ctx->orb->ev[5]));
    for (int i = 0; i < 5; i++) UCO_CUDA(ctx, cudaEventElapsedTime(&out[i], ctx->orb->ev[i], ctx->orb->ev[i + 1]));
    return UCO_OK;
}
// algorithmic bytes per frame of the current plan: out[0] = input image, out[1] = pyramid level pixels (sum w*h),
// out[2] = pyramid bytes incl. the 19-px borders
int uco_b200_orb_plan_bytes(uco_b200_ctx* ctx, uint64_t* out) {
    if (!ctx || !ctx->orb || !ctx->orb->batch_cap) return UCO_E_INVALID;
    const PlanDev& P = ctx->orb->plan;
    out[0] = (uint64_t)ctx->orb->w * ctx->orb->h;
    out[1] = out[2] = 0;
    for (int l = 0; l < P.n_levels; l++) {
        out[1] += (uint64_t)P.lv[l].w * P.lv[l].h;
        out[2] += (uint64_t)(P.lv[l].w + 2 * ORB_E) * (P.lv[l].h + 2 * ORB_E);
    }
    return UCO_OK;
}

// ---- test / inspection hooks (used by the per-stage parity tests) -----------------------------------------------------
int uco_b200_orb_debug_level_info(uco_b200_ctx* ctx, int level, int* w, int* h, int* pitch, int* n_desired, int* rows, int* cols) {
    if (!ctx || !ctx->orb || level < 0 || level >= ctx->orb->plan.n_levels) return UCO_E_INVALID;
    const LevelDev& L = ctx->orb->plan.lv[level];
    *w = L.w; *h = L.h; *pitch = L.pitch; *n_desired = L.n_desired; *rows = L.rows; *cols = L.cols;
    return UCO_OK;
}
// copies the bordered buffer of (frame, level) of the LAST extract call: (h+38) rows of `pitch` bytes
int uco_b200_orb_debug_pyramid(uco_b200_ctx* ctx, int frame, int level, uint8_t* out) {
    if (!ctx || !ctx->orb || level < 0 || level >= ctx->orb->plan.n_levels || frame < 0 || frame >= ctx->orb->batch_cap)
        return UCO_E_INVALID;
    const PlanDev& P = ctx->orb->plan;
    const LevelDev& L = P.lv[level];
    UCO_CUDA(ctx, cudaMemcpyAsync(out, ctx->orb->d_pyr + (size_t)frame * P.frame_bytes + L.off,
                                  (size_t)L.pitch * (L.h + 2 * ORB_E), cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UCO_OK;
}
// selected keypoints of (frame, level) before orientation: packed score<<24 | y<<12 | x ; returns count in *n
int uco_b200_orb_debug_selected(uco_b200_ctx* ctx, int frame, int level, uint32_t* out, int cap, int* n) {
    if (!ctx || !ctx->orb || level < 0 || level >= ctx->orb->plan.n_levels || frame < 0 || frame >= ctx->orb->batch_cap)
        return UCO_E_INVALID;
    const PlanDev& P = ctx->orb->plan;
    const LevelDev& L = P.lv[level];
    int cnt = 0;
    UCO_CUDA(ctx, cudaMemcpyAsync(&cnt, ctx->orb->d_sel_cnt + frame * ORB_MAXL + level, sizeof(int), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    if (cnt > cap) cnt = cap;
    UCO_CUDA(ctx, cudaMemcpyAsync(out, ctx->orb->d_sel + (size_t)frame * P.sel_per_frame + L.sel_off, sizeof(uint32_t) * cnt,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UCO_OK;
}

// raw FAST candidates of grid cell `cell` (index over all levels, row-major inside a level) of frame `frame`, as left by the
// LAST extract call (note: the selection stage permutes / filters these lists in place).  geom = {level,valid,x0,y0,w,h}.
int uco_b200_orb_debug_candidates(uco_b200_ctx* ctx, int frame, int cell, uint32_t* out, int cap, int* counts, int* geom) {
    if (!ctx || !ctx->orb || frame < 0 || frame >= ctx->orb->batch_cap || cell < 0 || cell >= ctx->orb->plan.n_cells_total)
        return UCO_E_INVALID;
    uco_orb_state* s = ctx->orb;
    const CellDev& C = s->cells[cell];
    geom[0] = C.level; geom[1] = C.valid; geom[2] = C.x0; geom[3] = C.y0; geom[4] = C.w; geom[5] = C.h;
    UCO_CUDA(ctx, cudaMemcpyAsync(counts, s->d_cand_cnt + ((size_t)frame * s->plan.n_cells_total + cell) * 2, 2 * sizeof(int),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int n = std::min(std::min(counts[0], cap), C.cand_cap);
    UCO_CUDA(ctx, cudaMemcpyAsync(out, s->d_cand + (size_t)frame * s->plan.cand_per_frame + C.cand_off, sizeof(uint32_t) * n,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UCO_OK;
}

// host-compiled copies of the exact-arithmetic helpers, so CPU-only tests can pin them to libm / OpenCV / libstdc++
int uco_b200_probe_math(int what, const float* in0, const float* in1, int n, float* out0, float* out1) {
    for (int i = 0; i < n; i++) {
        if (what == 0) out0[i] = uco_math::fast_atan2_deg(in0[i], in1[i]);
        else if (what == 1) uco_math::sincos_glibc(in0[i], &out0[i], &out1[i]);
        else return UCO_E_INVALID;
    }
    return UCO_OK;
}
int uco_b200_probe_retain_best(uint32_t* packed, int count, int n_points) {
    return uco_sel::retain_best_truncate(packed, count, n_points);
}

}  // extern "C"
