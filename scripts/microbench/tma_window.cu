// 2-D TMA window loads at arbitrary element coordinates (the orient_describe staging): which box shapes the hardware accepts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_window scripts/microbench/tma_window.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int COLS, int ROWS>
__global__ void __launch_bounds__(256) k(const CUtensorMap* tm, const int* xy, const unsigned char* img, int pitch, int* bad) {
    __shared__ __align__(128) unsigned char patch[8][(COLS * ROWS + 127) / 128 * 128];
    __shared__ __align__(8) unsigned long long bars[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, slot = blockIdx.x * 8 + w;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[w])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int x = xy[2 * slot], y = xy[2 * slot + 1];
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[w])), "r"(COLS * ROWS) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(patch[w])),
                     "l"(tm), "r"(x), "r"(y), "r"(0), "r"(s32(&bars[w]))
                     : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(&bars[w])), "r"(0) : "memory");
    int nb = 0;
    for (int r = 0; r < ROWS; r++)
        for (int c = lane; c < 39; c += 32) nb += patch[w][r * COLS + c] != img[(size_t)(y + r) * pitch + x + c];
    if (nb) atomicAdd(bad, nb);
}
static int g_nblk = 2000, g_fx = -1, g_fy = -1;
template <int COLS, int ROWS> int run(int pitch, int rows, const char* name) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    unsigned char* img; cudaMalloc(&img, (size_t)pitch * rows);
    unsigned char* h = (unsigned char*)malloc((size_t)pitch * rows);
    for (size_t i = 0; i < (size_t)pitch * rows; i++) h[i] = (unsigned char)(rand() & 255);
    cudaMemcpy(img, h, (size_t)pitch * rows, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, 1}, strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
    cuuint32_t box[3] = {COLS, ROWS, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof tm); cudaMemcpy(dtm, &tm, sizeof tm, cudaMemcpyHostToDevice);
    const int nblk = g_nblk, n = nblk * 8;
    int* hxy = (int*)malloc(8 * n);
    for (int i = 0; i < n; i++) { hxy[2 * i] = g_fx >= 0 ? g_fx : rand() % (pitch - 64); hxy[2 * i + 1] = g_fy >= 0 ? g_fy : rand() % (rows - ROWS); }
    int *xy, *bad; cudaMalloc(&xy, 8 * n); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    cudaMemcpy(xy, hxy, 8 * n, cudaMemcpyHostToDevice);
    k<COLS, ROWS><<<nblk, 256>>>(dtm, xy, img, pitch, bad);
    cudaError_t e = cudaDeviceSynchronize();
    int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("%-28s encode %d  run: %s  mismatching bytes %d\n", name, (int)r, cudaGetErrorString(e), hb);
    return e != cudaSuccess;
}
int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    if (argc > 2) g_nblk = atoi(argv[2]);
    if (argc > 3) g_fx = atoi(argv[3]);
    if (argc > 4) g_fy = atoi(argv[4]);
    if (which == 0) return run<48, 39>(704, 518, "box 48 x 39");
    if (which == 1) return run<64, 39>(704, 518, "box 64 x 39");
    if (which == 2) return run<48, 40>(704, 518, "box 48 x 40");
    if (which == 3) return run<64, 40>(704, 518, "box 64 x 40");
    if (which == 4) return run<32, 39>(704, 518, "box 32 x 39");
    return 0;
}
