// orb.cu — ORB pyramid extractor kernels (K1-K6).  Filled in below.
#include "common.cuh"
void uco_orb_state_free(uco_b200_ctx*) {}
