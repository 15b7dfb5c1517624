// ba.cu — bundle-adjustment kernels (K10-K14).
#include "common.cuh"
void uco_ba_state_free(uco_b200_ctx*) {}
