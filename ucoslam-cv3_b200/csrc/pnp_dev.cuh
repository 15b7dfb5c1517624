// pnp_dev.cuh — device-side problem layout of the pose-only LM kernel (pnp.cu), shared with the batched tracker (track.cu), which
// assembles the problems on the device from its match lists.
#pragma once
#include <stdint.h>

struct PnpHead {
    int n, nm, off, moff;  // matches / markers of this problem and their offsets in the concatenated arrays
    float pose44[16];
    double fx, fy, cx, cy, bf, wm;  // wm = WeightedHubber weight of the marker edges (pnpsolver.cpp:298-300)
};
struct PnpOut {
    float pose44[16];
    double pose7[7];
    int n_good;
    int iters[4];
    int pad;
};
struct PnpArrays {
    const float* pts;    // 3 per match
    const float* uv;     // 2 per match
    const float* ur;     // 1
    const float* isig;   // 1
    const uint8_t* flg;  // bit 0 stereo, bit 1 stable
    const float* mpose;  // 16 per marker
    const float* msize;  // 1
    const float* mobs;   // 8
    double* chi2;        // scratch, per match
    uint8_t* active;     // scratch, per match (level 0)
    double* mchi2;       // scratch, per marker
    uint8_t* mrobust;    // scratch, per marker
    uint8_t* bad;        // out, per match
};


struct uco_b200_ctx;
int uco_pnp_launch_dev(uco_b200_ctx* ctx, int n, const PnpHead* heads_dev, const PnpArrays& A, PnpOut* outs_dev);
