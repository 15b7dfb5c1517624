/*
 * ucoslam_b200.h — C ABI of the B200 (sm_100a) implementation of UcoSLAM's per-frame tracking hot path.
 *
 * Every entry point replaces one CPU function of the reference (lambdaloop/ucoslam-cv3); the reference
 * interface it stands in for is cited as file:line relative to the reference tree.  No C++ or torch types
 * cross this boundary: plain pointers, sizes and POD structs only.
 *
 * Conventions
 *   - every function returns an int status: 0 = UCO_OK, negative = UCO_E_*; the text of the last error is
 *     kept in the context (uco_b200_last_error).  The C++ adapters in ucoslam-cv3_b200/host/ turn a negative
 *     status into std::runtime_error, which is how the reference reports errors
 *     (src/featureextractors/feature2dserializable.cpp:71, 3rdparty/fbow/fbow/fbow.cpp:52-54).
 *   - pointers are caller-owned HOST buffers unless the function name ends in _dev, in which case they are
 *     device pointers valid on the context's device and the call is asynchronous on the context's stream.
 *   - a context owns one CUDA stream and its workspaces and is NOT thread safe: one context per calling
 *     thread (the reference has one extractor call in flight per System, and its OpenMP matcher callers can
 *     hold one context each).
 *   - there is no CPU fallback: if no CUDA device is usable uco_b200_create fails.
 */
#ifndef UCOSLAM_B200_H
#define UCOSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UCO_OK 0
#define UCO_E_INVALID (-1) /* bad argument */
#define UCO_E_CUDA (-2)    /* CUDA runtime error (text in last_error) */
#define UCO_E_NOMEM (-3)
#define UCO_E_CAPACITY (-4) /* caller-provided output capacity too small */
#define UCO_E_FORMAT (-5)   /* malformed vocabulary / stream */
#define UCO_E_ABORTED (-6)  /* stop flag raised */

typedef struct uco_b200_ctx uco_b200_ctx;

/* ------------------------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------------------------- */
uco_b200_ctx* uco_b200_create(int device, int flags);
void uco_b200_destroy(uco_b200_ctx* ctx);
const char* uco_b200_last_error(const uco_b200_ctx* ctx);
/* the cudaStream_t all work of this context is enqueued on (as void* so the header stays CUDA-free) */
void* uco_b200_stream(uco_b200_ctx* ctx);
int uco_b200_sync(uco_b200_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t uco_b200_launch_count(const uco_b200_ctx* ctx);
/* milliseconds accumulated on the context stream by the kernel class `what` between timing_begin/end.
 * what: 0 = all ORB kernels, 1 = hamming scan, 2 = bow, 3 = BA.  Measured with CUDA events on the stream. */
int uco_b200_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * K7  brute-force 256-bit Hamming k-NN
 *   replaces xflann::Index::build + Index::search with LinearParams
 *     3rdparty/xflann/xflann/index.cpp:45-69,77-103   (build / search / optional sort)
 *     3rdparty/xflann/xflann/impl/linear.h:68-88      (_knnsearch: scan all train rows in order)
 *     3rdparty/xflann/xflann/impl/resultset.h:28-140  (bounded max-heap; strict '<' replacement)
 *     3rdparty/xflann/xflann/impl/distances.h:279-283 (4x popcount64)
 *   which is what FrameMatcher_Flann calls through trainIndex.search (src/utils/framematcher.cpp:239).
 *
 *   q: nq rows of 32 bytes, row pitch q_stride bytes; t: nt rows, pitch t_stride.
 *   idx / dist: nq x k int32, dense.  Rows with fewer than k results are padded with idx = -1, dist = 0
 *   (int32 "quiet_NaN", linear.h:82-85).
 *   order: UCO_KNN_HEAP   -> exactly the array order xflann's max-heap leaves (sorted=false, the tracker's setting)
 *          UCO_KNN_SORTED -> the order after xflann's exchange sort (index.h:119-133, sorted=true)
 * ---------------------------------------------------------------------------------------------------------- */
#define UCO_KNN_HEAP 0
#define UCO_KNN_SORTED 1
#define UCO_KNN_MAX_K 32

int uco_b200_hamming_knn(uco_b200_ctx* ctx, const uint8_t* q, int nq, size_t q_stride, const uint8_t* t, int nt,
                         size_t t_stride, int k, int order, int32_t* idx, int32_t* dist);
/* device-resident variant: q_dev/t_dev dense 32-byte rows, outputs device buffers; asynchronous */
int uco_b200_hamming_knn_dev(uco_b200_ctx* ctx, const uint8_t* q_dev, int nq, const uint8_t* t_dev, int nt, int k,
                             int order, int32_t* idx_dev, int32_t* dist_dev);

#ifdef __cplusplus
}
#endif
#endif /* UCOSLAM_B200_H */
