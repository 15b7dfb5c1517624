// stereo_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).  Plain-C restatement of the association loop of
// FrameExtractor::processStereo (reference: src/utils/frameextractor.cpp:1410-2634; the file is macro-obfuscated, the statements
// followed here are those of its de-obfuscated text, SURVEY.md reading aid) with MapPoint::getDescDistance
// (src/map_types/mappoint.h:146-162,172-177).  The reference cannot be compiled here (OpenCV C++), so PARITY IS UNPINNED BY THE
// REFERENCE for this row; this restatement is cross-checked against an independent numpy / cv2 restatement (cv2.absdiff +
// cv2.sumElems for the two OpenCV calls of the loop) in tests/test_stereo_oracle.py.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <float.h>

typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } kp_t;   // cv::KeyPoint

static float hamm(const uint8_t* a, const uint8_t* b) {   // getHammDescDistance_2: 4 x popcount64, returned as float
    const uint64_t* x = (const uint64_t*)a; const uint64_t* y = (const uint64_t*)b;
    uint64_t n = 0;
    for (int i = 0; i < 4; i++) n += (uint64_t)__builtin_popcountll(x[i] ^ y[i]);
    return (float)n;
}

// returns the number of keypoints that received a depth, -1 where the reference's cv::Mat ROI constructor would throw
int oracle_stereo_depth(const uint8_t* img_l, size_t stride_l, const uint8_t* img_r, size_t stride_r, int cols, int rows,
                        const kp_t* kl, const uint8_t* dl, int nl, const kp_t* kr, const uint8_t* dr, int nr, float maxDescDistance,
                        float bl, float fx, float* depth, int32_t* match) {
    // vector<vector<int>> rowIndices(rows): band of 0 rows around round(y), keypoint order
    int* count = (int*)calloc((size_t)rows + 1, sizeof(int));
    int* start = (int*)calloc((size_t)rows + 2, sizeof(int));
    int* list = (int*)malloc(sizeof(int) * (size_t)(nr > 0 ? nr : 1));
    for (int i = 0; i < nr; i++) {
        double r = 0, y = kr[i].y;
        int lo = (int)round(y - r), hi = (int)round(y + r);
        if (lo < 0) lo = 0;
        if (hi > rows - 1) hi = rows - 1;
        for (int yy = lo; yy <= hi; yy++) count[yy]++;
    }
    for (int y = 0; y < rows; y++) start[y + 1] = start[y] + count[y];
    for (int y = 0; y < rows; y++) count[y] = 0;
    for (int i = 0; i < nr; i++) {
        double r = 0, y = kr[i].y;
        int lo = (int)round(y - r), hi = (int)round(y + r);
        if (lo < 0) lo = 0;
        if (hi > rows - 1) hi = rows - 1;
        for (int yy = lo; yy <= hi; yy++) list[start[yy] + count[yy]++] = i;
    }
    int nmatches = 0, bad = 0;
    for (int i = 0; i < nl; i++) {
        depth[i] = 0;
        if (match) match[i] = -1;
        int y = (int)roundf(kl[i].y);
        if (y < 0 || y >= rows) continue;   // (the reference indexes the bucket vector unchecked)
        int bestIdx = -1;
        double bestDist = DBL_MAX;
        for (int c = start[y]; c < start[y + 1]; c++) {
            int j = list[c];
            if (kr[j].x > kl[i].x || abs(kr[j].octave - kl[i].octave) > 1) continue;
            float d = hamm(dl + (size_t)i * 32, dr + (size_t)j * 32);
            if (d < maxDescDistance) {
                if (d < bestDist) { bestDist = d; bestIdx = j; }
            }
        }
        if (bestIdx == -1) continue;
        if (match) match[i] = bestIdx;
        const int w = 7, hw = w / 2;
        int xl = (int)roundf(kl[i].x), yl = (int)roundf(kl[i].y);
        if (xl < hw || xl + hw >= cols) continue;
        if (yl < hw || yl + hw >= rows) continue;
        int xr = (int)roundf(kr[bestIdx].x), yr = (int)roundf(kr[bestIdx].y);
        if (xr < hw || xr + hw >= cols) continue;
        if (yr < hw || yr + hw >= rows) continue;
        const int L = 7;
        double vDists[2 * 7 + 1];
        int lo = -L > -xr ? -L : -xr, hi = L < cols - 1 - xr ? L : cols - 1 - xr;
        double best = DBL_MAX;
        int bestInc = -1;
        for (int inc = lo; inc <= hi; inc++) {
            int idx = inc + L, xc = xr + inc;
            if (xc - hw < 0 || xc + hw > cols) { bad = 1; break; }   // cv::Mat(Range, Range) asserts
            double sum = 0;
            for (int r = -hw; r < hw; r++)                           // cv::Range(a, b) is [a, b): 6 x 6
                for (int c = -hw; c < hw; c++)
                    sum += abs((int)img_l[(size_t)(yl + r) * stride_l + xl + c] - (int)img_r[(size_t)(yr + r) * stride_r + xc + c]);
            if (sum < best) { best = sum; bestInc = idx; }
            vDists[idx] = sum;
        }
        if (bad) break;
        if (bestInc > lo + L && bestInc < hi + L) {
            double d1 = vDists[bestInc - 1], d2 = vDists[bestInc], d3 = vDists[bestInc + 1];
            double deltaR = 0.5 * (d1 - d3) / (d1 + d3 - 2 * d2) + bestInc - L;
            double xs = kr[bestIdx].x + deltaR;
            depth[i] = (bl * fx) / (kl[i].x - xs);
            nmatches++;
        }
    }
    free(count); free(start); free(list);
    return bad ? -1 : nmatches;
}
