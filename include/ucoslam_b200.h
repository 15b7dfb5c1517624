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
/* when on, multi-kernel entry points record CUDA events at their stage boundaries (read back with *_last_stage_ms) */
void uco_b200_set_profiling(uco_b200_ctx* ctx, int on);

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

/* a batch of independent (query set, train set) pairs in ONE launch (a clip: frame i against frame i-1).  Pair p reads
 * q_dev + p*q_pair_stride (nq_dev ? min(nq_max, nq_dev[p]) : nq_max rows) and t_dev + p*t_pair_stride, and writes
 * idx_dev/dist_dev + p*nq_max*k.  The optional per-pair row counts are DEVICE arrays (e.g. n_out_dev of
 * uco_b200_orb_extract_batch_dev), so no host round trip is needed; rows >= nq of a pair are left untouched. */
int uco_b200_hamming_knn_batch_dev(uco_b200_ctx* ctx, int n_pairs, const uint8_t* q_dev, size_t q_pair_stride, int nq_max,
                                   const int32_t* nq_dev, const uint8_t* t_dev, size_t t_pair_stride, int nt_max,
                                   const int32_t* nt_dev, int k, int order, int32_t* idx_dev, int32_t* dist_dev);

/* Host-buffer batch of (query set, train set) pairs: one launch, one synchronisation (the host-side twin of the call above, for
 * callers that hold descriptors in cv::Mat rows).  q[i] / t[i]: nq[i] / nt[i] rows of 32 bytes with row strides q_stride / t_stride
 * (the same for every pair); idx[i] / dist[i]: nq[i] x k outputs as uco_b200_hamming_knn.  The chain pattern of tracking
 * (t[i] == q[i-1]) is detected and every descriptor block is uploaded once. */
int uco_b200_hamming_knn_batch(uco_b200_ctx* ctx, int n_pairs, const uint8_t* const* q, const int32_t* nq, size_t q_stride,
                               const uint8_t* const* t, const int32_t* nt, size_t t_stride, int k, int order, int32_t* const* idx,
                               int32_t* const* dist);

/* ------------------------------------------------------------------------------------------------------------
 * K1-K6  ORB pyramid extractor
 *   replaces ucoslam::ORBextractor::detectAndCompute_impl -> compute()
 *     src/featureextractors/ORBextractor.cpp:1139-1149, 1247-1351   (and everything it calls, see csrc/orb.cu)
 *   behind the plugin interface ucoslam::Feature2DSerializable
 *     src/featureextractors/feature2dserializable.h:30-95  (FeatParams :34-61 -> uco_orb_params)
 *   Output is what the reference returns: keypoints in cv::KeyPoint memory layout (28 bytes), level-major, and
 *   N x 32 descriptor bytes; bit-exact against the CPU path (OpenCV 4.13 without IPP, libstdc++ 13, glibc libm).
 * ---------------------------------------------------------------------------------------------------------- */
#define UCO_ORB_MAX_LEVELS 16

typedef struct uco_keypoint { /* == cv::KeyPoint */
    float x, y;     /* pt, level-0 coordinates */
    float size;     /* (int)(31 * scale[octave]) */
    float angle;    /* degrees, [0,360) */
    float response; /* FAST score */
    int32_t octave;
    int32_t class_id; /* -1 */
} uco_keypoint;

typedef struct uco_orb_params {
    int32_t max_features;  /* FeatParams::maxFeatures */
    int32_t n_levels;      /* FeatParams::nOctaveLevels */
    float scale_factor;    /* FeatParams::scaleFactor */
    int32_t ini_th_fast;   /* 20, ORBextractor.cpp:478 */
    int32_t min_th_fast;   /* 7,  ORBextractor.cpp:479 */
    int32_t blur_first;    /* ORBextractor::_doGaussianBlurAtFirst (default true) */
} uco_orb_params;

void uco_b200_orb_default_params(uco_orb_params* p);

/* one frame: img is a host CV_8UC1 image (rows `stride` bytes apart). kps/desc hold `capacity` >= max_features entries. */
int uco_b200_orb_extract(uco_b200_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, const uco_orb_params* prm,
                         uco_keypoint* kps, uint8_t* desc, int capacity, int* n_out);
/* n_imgs frames of identical size in one pass (the throughput path): frame i writes kps[i*capacity ..], desc[i*capacity*32 ..],
 * n_out[i]. */
int uco_b200_orb_extract_batch(uco_b200_ctx* ctx, const uint8_t* const* imgs, int n_imgs, int w, int h, size_t stride,
                               const uco_orb_params* prm, uco_keypoint* kps, uint8_t* desc, int capacity, int* n_out);
/* device-resident, asynchronous: frames at imgs_dev + i*frame_stride, rows `pitch` bytes apart; outputs are device buffers of
 * n_imgs * max_features entries (kps, desc) and n_imgs ints. */
int uco_b200_orb_extract_batch_dev(uco_b200_ctx* ctx, const uint8_t* imgs_dev, int n_imgs, int w, int h, size_t pitch,
                                   size_t frame_stride, const uco_orb_params* prm, uco_keypoint* kps_dev,
                                   uint8_t* desc_dev, int* n_out_dev);

int uco_b200_orb_last_stage_ms(uco_b200_ctx* ctx, float* out5);
int uco_b200_orb_plan_bytes(uco_b200_ctx* ctx, uint64_t* out3);

/* inspection hooks for the per-stage parity tests (state of the last extract call on this context) */
int uco_b200_orb_debug_level_info(uco_b200_ctx* ctx, int level, int* w, int* h, int* pitch, int* n_desired, int* rows,
                                  int* cols);
int uco_b200_orb_debug_pyramid(uco_b200_ctx* ctx, int frame, int level, uint8_t* out);
int uco_b200_orb_debug_selected(uco_b200_ctx* ctx, int frame, int level, uint32_t* out, int cap, int* n);
int uco_b200_orb_debug_candidates(uco_b200_ctx* ctx, int frame, int cell, uint32_t* out, int cap, int* counts, int* geom);
/* host-compiled copies of the exact-arithmetic device helpers (no GPU needed): what = 0 fastAtan2(in0=y, in1=x) -> out0;
 * what = 1 sinf/cosf(in0) -> out0 = sin, out1 = cos.  retain_best: packed score<<24|y<<12|x, returns new count. */
int uco_b200_probe_math(int what, const float* in0, const float* in1, int n, float* out0, float* out1);
int uco_b200_probe_retain_best(uint32_t* packed, int count, int n_points);

/* ------------------------------------------------------------------------------------------------------------
 * K9  bag-of-words transform (fbow vocabulary tree)
 *   replaces fbow::Vocabulary::fromStream / transform(features, level, fBow&, fBow2&)
 *     3rdparty/fbow/fbow/fbow.cpp:180-190, 51-90 ; 3rdparty/fbow/fbow/fbow.h:402-448
 *   as used by KPFrameDataBase::computeBow, src/map_types/keyframedatabase.cpp:310-321 (level 3).
 *   bytes: a complete .fbow stream (e.g. 3rdparty/vocabularies/orb.fbow).  Only CV_8UC1 / 32-byte vocabularies.
 *   transform emits, per descriptor i: word[i] (0xFFFFFFFF = none), weight[i], node[i] = id of the tree node reached at
 *   `level` (0xFFFFFFFF = none).  The host folds them in descriptor order into fBow (word -> sum of weights) and
 *   fBow2 (node -> descriptor indices) exactly as fbow.h:428-436 does (see ucoslam-cv3_b200/host/bow_b200.h).
 *   Errors mirror the reference: n == 0 -> "Vocabulary::transform No input data", bad signature -> UCO_E_FORMAT.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_b200_voc uco_b200_voc;
int uco_b200_bow_load(uco_b200_ctx* ctx, const void* bytes, size_t n, uco_b200_voc** voc);
void uco_b200_bow_free(uco_b200_ctx* ctx, uco_b200_voc* voc);
int uco_b200_bow_info(const uco_b200_voc* voc, uint32_t* k, uint32_t* nblocks, uint32_t* desc_size);
int uco_b200_bow_transform(uco_b200_ctx* ctx, const uco_b200_voc* voc, const uint8_t* desc, int n, size_t stride, int level,
                           uint32_t* word, float* weight, uint32_t* node);
int uco_b200_bow_transform_dev(uco_b200_ctx* ctx, const uco_b200_voc* voc, const uint8_t* desc_dev, int n, int level,
                               uint32_t* word_dev, float* weight_dev, uint32_t* node_dev);


/* ------------------------------------------------------------------------------------------------------------
 * K10-K13  bundle adjustment (local and global): two-stage Levenberg-Marquardt with Schur complement
 *   replaces ucoslam::GlobalOptimizerG2O::{setParams, optimize, getResults}
 *     src/optimization/globaloptimizer_g2o.cpp:77-401, 418-463, 466-538
 *   behind the plugin interface ucoslam::GlobalOptimizer (src/optimization/globaloptimizer.h:28-68), i.e. what g2o does
 *   for that graph:
 *     src/optimization/typesg2o.h:249-325, 338-405                 EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ residual + Jacobians
 *     3rdparty/g2o/g2o/core/base_binary_edge.hpp:83-155            per-edge quadratic form with Huber rho' weighting
 *     3rdparty/g2o/g2o/core/block_solver.hpp:315-443               Schur complement, reduced solve, landmark back-substitution
 *     3rdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:58-175  LM control;  sparse_optimizer.cpp:366-436 outer loop
 *   The caller (the C++ adapter GlobalOptimizerB200, ucoslam-cv3_b200/host/) flattens the Map into this structure exactly
 *   as setParams walks it: one row per keyframe vertex, per map point vertex and per (map point, keyframe) observation.
 *   All arithmetic is f64 on the device; inputs are the reference's f32 containers (cv::Mat CV_32F pose, cv::Point3f,
 *   cv::KeyPoint::pt, vector<float> _InvScaleFactors).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_ba_problem {
    int32_t n_poses, n_points, n_obs;
    const float* poses44;        /* n_poses x 16  Frame::pose_f2g, row-major 4x4 */
    const uint8_t* fixed;        /* n_poses       != 0: vertex fixed (isFixedFrame) */
    const float* points3;        /* n_points x 3  MapPoint::getCoordinates() */
    const int32_t* obs_pose;     /* n_obs         index into poses */
    const int32_t* obs_point;    /* n_obs         index into points */
    const float* obs_uv;         /* n_obs x 2     und_kpts[].pt */
    const float* obs_ur;         /* n_obs         right-image x (kp_ur) of stereo observations; may be NULL when no stereo */
    const uint8_t* obs_stereo;   /* n_obs         != 0: EdgeStereoSE3ProjectXYZ; may be NULL (all monocular) */
    const float* obs_inv_sigma2; /* n_obs         _InvScaleFactors[kp.octave] */
    float fx, fy, cx, cy, bf;    /* ImageParams; bf = baseline * fx */
    int32_t n_iters;             /* ParamSet::nIters: stage 1 runs n_iters LM iterations, stage 2 runs 2 * n_iters */
    /* ArUco markers (globaloptimizer_g2o.cpp:157-170, 304-350; zero / NULL when there are none): one free SE3 vertex per map marker
     * with a valid pose and one MarkerEdge (typesg2o.h:108-167: the 4 corners reprojected through camera * marker, 8 residuals, numeric
     * Jacobian with delta 1e-4, no robust kernel) per (marker, keyframe) observation.  Problems with markers are solved by the sharded
     * solver (one rank when no communicator is given). */
    int32_t n_markers;
    const float* marker_pose44;  /* n_markers x 16   Marker::pose_g2m (global <- marker), row-major 4x4 */
    const float* marker_size;    /* n_markers        Marker::size */
    int32_t n_marker_obs;
    const int32_t* mobs_marker;  /* n_marker_obs     index into the markers */
    const int32_t* mobs_pose;    /* n_marker_obs     index into poses */
    const float* mobs_corners;   /* n_marker_obs x 8 MarkerObservation::und_corners */
    const float* mobs_weight;    /* n_marker_obs     frame_MarkerWeight[frame] (:276-297): information = I8 * weight */
    /* Keyframes taken with different cameras in one window (a map built with one camera and extended with another): every edge carries
     * the ImageParams of ITS keyframe (globaloptimizer_g2o.cpp:233-236, :262-266, :335-338).  NULL: all keyframes use fx..bf above.
     * Problems with this table are solved by the sharded solver (one rank when no communicator is given). */
    const float* pose_cam;       /* n_poses x 5      fx fy cx cy bf (= bl * fx) of each keyframe, or NULL */
    /* The InPlaneMarkers option (globaloptimizer_g2o.cpp:356-401; n_plane = 0: off): the map's reference marker (the valid marker seen by
     * most keyframes, :362-368) is tied to every OTHER marker vertex of the window by one MarkerEdgeX (:37-66: with M = inverse(ref) * other,
     * residuals 10 * (M(0,2), M(1,2), 1 - M(2,2), M(2,3)); g2o's numeric Jacobian, delta 1e-9f; information = I4 * plane_weight, no kernel).
     * plane_ref is the reference's index among the markers above, or -1 when it is not a vertex of this window: it then enters as a FIXED
     * vertex with the pose in plane_ref_pose44 (:371-379). */
    int32_t n_plane;             /* planar edges = markers of the window other than the reference */
    int32_t plane_ref;
    const float* plane_ref_pose44; /* 16             Marker::pose_g2m of the reference marker when plane_ref < 0 */
    const int32_t* plane_other;  /* n_plane          index into the markers */
    double plane_weight;         /* 0.33 * (sum of marker weights x 8 + sum of keypoint weights) / (4 * n_plane)   (:381-382) */
} uco_ba_problem;

typedef struct uco_ba_result {   /* every pointer may be NULL (not wanted) */
    double* pose7;               /* n_poses x 7   qx qy qz qw tx ty tz of the optimised vertices (f64) */
    float* poses44;              /* n_poses x 16  what getResults stores in Frame::pose_f2g (fixed poses: input copied) */
    double* points3;             /* n_points x 3  optimised points (f64; getResults narrows to cv::Point3f) */
    double* obs_chi2;            /* n_obs         chi2 held by each edge at the end */
    uint8_t* obs_level;          /* n_obs         1: edge was excluded after stage 1 (level 1) */
    uint8_t* obs_bad;            /* n_obs         getBadAssociations() predicate, globaloptimizer_g2o.cpp:506-521 */
    double* trace;               /* 64 x 2        per LM iteration: robust chi2, number of LM trials (both stages, in order) */
    int32_t iters[2];            /* iterations executed by stage 1 / stage 2 (SparseOptimizer::optimize return values) */
    float device_ms;             /* device time of the solve between first and last kernel (CUDA events) */
    double* profile;             /* 16 doubles, optional: SM cycles CTA 0 of the window's cluster spent per phase (cluster-resident form) */
    float* marker_poses44;       /* n_markers x 16   optimised Marker::pose_g2m (getResults :526-527); may be NULL */
    double* marker_pose7;        /* n_markers x 7    the same vertices as qx qy qz qw tx ty tz (f64); may be NULL */
    double* mobs_chi2;           /* n_marker_obs     chi2 of each marker edge at the end; may be NULL */
} uco_ba_result;

/* stop: optional one-byte flag (the reference's bool* stopASAP can be passed as it is) polled between LM trials
 * (GlobalOptimizerG2O::optimize(bool* stopASAP), sparse_optimizer.h:189);
 * when raised during stage 1 the solve returns the current estimate with UCO_OK and iters[1] = 0 like the reference. */
int uco_b200_ba_solve(uco_b200_ctx* ctx, const uco_ba_problem* pb, const volatile unsigned char* stop, uco_ba_result* res);
/* n independent problems (one local-BA window per camera / map) solved together; results as n separate uco_b200_ba_solve calls */
int uco_b200_ba_solve_batch(uco_b200_ctx* ctx, int n, const uco_ba_problem* pbs, const volatile unsigned char* stop, uco_ba_result* res);
/* ---- multi-GPU (BASELINE config 5: global BA sharded over the GPUs of a node; config 4: row-sharded descriptor map) -------------
 * One process per GPU.  The process group is the caller's (torch.distributed under torchrun, MPI, ...): rank 0 obtains a 128-byte
 * id, the caller distributes it, every rank creates its communicator on its own context.  Collectives are NCCL over NVLink /
 * NVSwitch, issued on the context's stream.  world == 1 needs no id and no NCCL. */
typedef struct uco_b200_comm uco_b200_comm;
int uco_b200_comm_unique_id(uint8_t* id128);
int uco_b200_comm_create(uco_b200_ctx* ctx, const uint8_t* id128, int rank, int world, uco_b200_comm** out);
void uco_b200_comm_destroy(uco_b200_comm* comm);

/* Hamming k-NN over a ROW-SHARDED train set (config 4: a 10^6-descriptor map split over the GPUs): rank r holds rows
 * [row_base, row_base + nt_shard) of the map, the queries are replicated.  Every rank scans its shard, the per-shard top-k lists
 * are all-gathered (nq x k 64-bit keys per rank) and merged on every rank.  Output rows are in (distance, row index) order with
 * the same distance list as xflann's linear scan and the same rows for every distance below the k-th.  Among rows TIED at the
 * k-th distance xflann keeps whichever its max-heap still holds after later, closer rows evicted the root (resultset.h:64-85) —
 * a function of the scan order that no partition of the scan can reproduce; the merged lists keep the lowest row indices there.
 * comm == NULL: one shard.  With few queries a shard is itself scanned in up to 128 row ranges by different CTAs and merged the same
 * way, so that a single query still streams the map at HBM speed.  Asynchronous on the stream. */
int uco_b200_hamming_knn_sharded_dev(uco_b200_ctx* ctx, uco_b200_comm* comm, const uint8_t* q_dev, int nq, const uint8_t* t_shard_dev,
                                     int nt_shard, int row_base, int k, int32_t* idx_dev, int32_t* dist_dev);
/* the merge step on its own: n_lists (n_lists * k <= 1024) lists of nq x k (global row index, distance) with -1 padding, list-major */
int uco_b200_knn_merge_dev(uco_b200_ctx* ctx, int n_lists, int nq, int k, const int32_t* idx_lists_dev, const int32_t* dist_lists_dev,
                           int32_t* idx_dev, int32_t* dist_dev);

/* GlobalOptimizerG2O::optimize on a problem of any size, optionally sharded: every rank passes the SAME complete problem; the
 * landmarks (with all their observations: the Hll / Hpl columns of block_solver.hpp:329-400) are partitioned over the ranks, each
 * rank linearizes its part and builds its partial Hpp / bp and partial Schur complement, ONE all-reduce per LM trial sums the
 * packed reduced Hessian + right-hand side over NVLink, every rank solves the reduced system redundantly (dense Cholesky; above
 * 170 free keyframes the two-level block-envelope Cholesky of uco_b200_block_solve) and back-substitutes its own landmarks.  The LM control sums (chi2, scale) and the
 * ranks' stop flags are all-reduced too, so every rank takes identical decisions.  Every rank returns the complete result.
 * comm == NULL: single GPU.  Results equal uco_b200_ba_solve's up to the summation order of the Schur complement. */
int uco_b200_ba_solve_sharded(uco_b200_ctx* ctx, uco_b200_comm* comm, const uco_ba_problem* pb, const volatile unsigned char* stop,
                              uco_ba_result* res);
/* host-only inspection hook: landmark ranges of the sharded solver, out[0..world] = boundaries, out[world+1 .. 2 world] = observations per rank */
int uco_b200_probe_ba_partition(const uco_ba_problem* pb, int world, int* out);

/* The linear-solver seam on its own (3rdparty/g2o/g2o/core/linear_solver.h:44-90, LinearSolver<MatrixType>::solve(A, x, b); the
 * reference plugs LinearSolverEigen into it, solvers/eigen/linear_solver_eigen.h:92-123): S x = b for a symmetric positive definite
 * block-sparse S of nb x nb blocks of 6 x 6, given by nblk blocks of its UPPER block triangle (blk_ij[2 b] <= blk_ij[2 b + 1], every
 * diagonal block present, blocks row-major, host buffers).  This is the solver uco_b200_ba_solve_sharded runs on reduced systems of
 * more than 170 free keyframes: a two-level block-envelope Cholesky (csrc/ba_band.cu) — separator levels of the block graph's
 * breadth-first level structure, the pieces between them factored by one thread block each, the separator system by one more.
 * force_k < 0: number of separator levels chosen by the cost model; >= 0: forced (0 = plain envelope Cholesky in one thread block).
 * info8 (optional): {fronts, separator levels, root rows, root window, longest interior front, widest interior window, widest
 * border, 1 if a pivot was not positive (x = 0 then)}. */
int uco_b200_block_solve(uco_b200_ctx* ctx, int nb, int nblk, const int* blk_ij, const double* blocks, const double* rhs, int force_k, double* x,
                         int* info8);
/* device time (ms) of the solver's launches (assemble + fronts + root + back substitution) in the LAST uco_b200_block_solve of the
 * calling thread, measured on a second pass with CUDA events when context profiling is on */
int uco_b200_block_solve_profile(uco_b200_ctx* ctx, float* ms);
/* host-only inspection hook of its planner (no GPU needed): same ordering, fronts and storage map, the same algebra executed by plain
 * host loops.  smem_optin <= 0: 227 KB.  info8[7] = doubles of factor storage.  Returns 0, 1 (a pivot was not positive), < 0 (bad input). */
int uco_b200_probe_block_solve(int nb, int nblk, const int* blk_ij, const double* blocks, const double* rhs, int smem_optin, int force_k, double* x,
                               int* info8);

/* tuning / test knob.  mode 0 (default): windows with <= 38 free keyframes run cluster-resident (one thread-block cluster per
 * window, the whole LM loop in one launch), larger ones as streamed kernels; 1: always streamed; 2: always cluster-resident.
 * cluster_size: CTAs per cluster (power of two <= 16, 0 = 8). */
int uco_b200_ba_set_mode(uco_b200_ctx* ctx, int mode, int cluster_size);
/* worker threads the host-side planner of uco_b200_ba_solve_batch may use per call (0 = automatic: this process's share of the host
 * cores).  A caller that keeps several batches in flight from several mapper threads wants 1; a single caller wants the default. */
int uco_b200_ba_set_host_threads(uco_b200_ctx* ctx, int n_threads);
/* host-only inspection hook (no GPU needed): builds the window structure the solver uses and reports
 * {free poses, Schur blocks, gather units, contributions, chunks, pose-list entries, max observations per chunk, landmarks covered}.
 * cluster_size >= 1000: the plan the cluster-resident kernel gets for a cluster of (cluster_size - 1000) CTAs (fused lists for small windows) */
int uco_b200_probe_ba_plan(const uco_ba_problem* pb, int cluster_size, int* out8);

/* ------------------------------------------------------------------------------------------------------------
 * K14  pose-only optimisation (tracking): 4 rounds x 10 Levenberg-Marquardt iterations on one camera pose
 *   replaces ucoslam::PnPSolver::solvePnp(frame, map, matches_io, pose_io, currentKeyFrame)
 *     src/optimization/pnpsolver.cpp:116-408, src/optimization/pnpsolver.h:33   (tracker thread, 2-3 calls per frame)
 *   i.e. what g2o does for that graph:
 *     src/optimization/typesg2o.h:590-663, 521-588   EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose
 *     src/optimization/typesg2o.h:82-105             WeightedHubberRobustKernel (weight 0.5 for unstable points, x2 stereo)
 *     src/optimization/typesg2o.h:414-470            MarkerEdgeOnlyProject (numeric Jacobian, delta 1e-4; base_binary_edge.hpp:167-232)
 *     3rdparty/g2o/g2o/core/base_unary_edge.hpp:50-80, optimization_algorithm_levenberg.cpp:58-175, sparse_optimizer.cpp:366-436
 *   The caller flattens the matches as solvePnp walks them (pnpsolver.cpp:200-259): one row per cv::DMatch.
 *   n_matches == 0 and n_markers == 0 returns n_good = 0 and the input pose, like the reference (:144).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_pnp_problem {
    int32_t n_matches;
    const float* pose44;         /* 16            estimatedPose, row-major 4x4 (frame <- global) */
    const float* points3;        /* n x 3         map_points[trainIdx].getCoordinates() */
    const float* obs_uv;         /* n x 2         frame.und_kpts[queryIdx].pt */
    const float* obs_ur;         /* n             kpt.pt.x - mbf/depth of stereo matches (pnpsolver.cpp:226); may be NULL */
    const uint8_t* obs_stereo;   /* n             frame.getDepth(queryIdx) > 0; may be NULL (all monocular) */
    const float* obs_inv_sigma2; /* n             1/scaleFactors[kpt.octave] */
    const uint8_t* stable;       /* n             MapPoint::isStable(); may be NULL (all stable) */
    float fx, fy, cx, cy, bf;    /* ImageParams; bf = bl * fx */
    int32_t n_markers;           /* frame markers with a valid map pose seen from the neighbourhood (pnpsolver.cpp:262-281) */
    const float* marker_pose44;  /* n_markers x 16  Marker::pose_g2m */
    const float* marker_size;    /* n_markers */
    const float* marker_corners; /* n_markers x 8   MarkerObservation::und_corners */
} uco_pnp_problem;

typedef struct uco_pnp_result {
    float pose44[16];            /* estimatedPose after the call */
    double pose7[7];             /* the same vertex as qx qy qz qw tx ty tz (f64) */
    int32_t n_good;              /* return value of solvePnp: matches not flagged as outliers */
    int32_t iters[4];            /* LM iterations executed by each of the 4 rounds */
    uint8_t* bad;                /* n_matches, caller-owned, may be NULL: != 0 <-> map_matches[i].imgIdx = -1 */
} uco_pnp_result;

int uco_b200_pose_only(uco_b200_ctx* ctx, const uco_pnp_problem* pb, uco_pnp_result* res);
/* n independent problems (frames / cameras) in one launch, one thread block each */
int uco_b200_pose_only_batch(uco_b200_ctx* ctx, int n, const uco_pnp_problem* pbs, uco_pnp_result* res);

/* ------------------------------------------------------------------------------------------------------------
 * K8  descriptor matcher between two frames: exact k-NN (K7) + the reference's post-filters, on the device
 *   replaces FrameMatcher_Flann::setParams + match / matchEpipolar
 *     src/utils/framematcher.cpp:200-215, 216-322   (k = 10 candidates per query in index order; minDescDist, octave gate,
 *                                                    optional epipolar gate, second-best ratio test)
 *     src/basictypes/misc.cpp:153-185, 105-107      filter_ambiguous_train + remove_unused_matches
 *     src/utils/framematcher.cpp:67-108, 288-316    rotation-consistency histogram (three fullest of 30 bins)
 *     src/basictypes/misc.h:72-81                   epipolarLineSqDist
 *   behind _impl::FrameMatcher_impl (src/utils/framematcher.cpp:31-58).  The candidates come from the EXACT linear search
 *   (K7), not from the reference's 16-check k-means tree, which approximates them.
 *   q_map / t_map: keypoint index of descriptor row i (FrameMatcher::manageMode :160-198 packs the rows of the requested
 *   mode); NULL = identity (MODE_ALL).  Output: cv::DMatch records in the reference's order.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_match { /* == cv::DMatch */
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
} uco_match;

#define UCO_MATCH_MAX_SCALES 32
typedef struct uco_match_params {
    float min_desc_dist;       /* _minDescDist */
    float nn_match_ratio;      /* _nn_match_ratio (0.8 default, 0.6 tracker) */
    int32_t check_orientation; /* _checkOrientation */
    int32_t max_octave_diff;   /* _maxOctaveDiff */
    int32_t use_f12;           /* != 0: epipolar gate with F12 (matchEpipolar with a non-empty FQ2T) */
    float f12[9];              /* row-major 3x3, computeF12(...) of the caller (src/basictypes/misc.cpp:893-920) */
    int32_t n_scales;          /* queryFrame.scaleFactors.size() */
    float scale_factors[UCO_MATCH_MAX_SCALES];
} uco_match_params;

int uco_b200_frame_match(uco_b200_ctx* ctx, const uint8_t* q_desc, int nq, size_t q_stride, const uco_keypoint* q_kps,
                         int n_q_kps, const int32_t* q_map, const uint8_t* t_desc, int nt, size_t t_stride,
                         const uco_keypoint* t_kps, int n_t_kps, const int32_t* t_map, const uco_match_params* prm,
                         uco_match* out, int capacity, int* n_out);
/* a clip in one pass, device resident (outputs of uco_b200_orb_extract_batch_dev): pair p matches the descriptors at
 * q_desc_dev + p*q_pair_stride (bytes; keypoints at q_kps_dev + p*q_kps_pair_stride records) against t_* likewise, all
 * keypoints used (MODE_ALL); nq_dev / nt_dev: optional per-pair row counts.  Writes out_dev[p*nq_max ..] and n_out_dev[p]. */
int uco_b200_frame_match_batch_dev(uco_b200_ctx* ctx, int n_pairs, const uint8_t* q_desc_dev, size_t q_pair_stride,
                                   const uco_keypoint* q_kps_dev, size_t q_kps_pair_stride, int nq_max, const int32_t* nq_dev,
                                   const uint8_t* t_desc_dev, size_t t_pair_stride, const uco_keypoint* t_kps_dev,
                                   size_t t_kps_pair_stride, int nt_max, const int32_t* nt_dev, const uco_match_params* prm,
                                   uco_match* out_dev, int32_t* n_out_dev);

/* the mapper's pattern (new-map-point creation, src/utils/mapmanager.cpp:9972-10065): FrameMatcher::setParams(train = the keyframe)
 * once, then matchEpipolar(query = each of n_frames neighbour keyframes, F12_f).  t_*: the keyframe (descriptor rows of the
 * requested mode + t_map, as in uco_b200_frame_match); q_*[f]: the neighbours; f12: n_frames x 9 (row-major, required when
 * prm->use_f12; prm->f12 is ignored); out[f]: capacity `capacity` >= nq[f] matches, n_out[f].  One upload, two launches, one download. */
int uco_b200_frame_match_multi(uco_b200_ctx* ctx, const uint8_t* t_desc, int nt, size_t t_stride, const uco_keypoint* t_kps, int n_t_kps,
                               const int32_t* t_map, int n_frames, const uint8_t* const* q_desc, const int32_t* nq, size_t q_stride,
                               const uco_keypoint* const* q_kps, const int32_t* n_q_kps, const int32_t* const* q_map, const float* f12,
                               const uco_match_params* prm, uco_match* const* out, int capacity, int32_t* n_out);

/* per-keyframe work of the mapper on DEVICE-RESIDENT frames (rows of a frame-strided buffer: the extractor's batch output):
 * the bag of words of each keyframe (KPFrameDataBase::computeBow, src/map_types/keyframedatabase.cpp:310-321: fbow transform at
 * `bow_level`; voc may be NULL to skip it) and FrameMatcher::setParams(train = keyframe) + matchEpipolar(query = each neighbour)
 * (src/utils/mapmanager.cpp:9972-10065).  kf_frame[n_kf], nb_ptr[n_kf+1], nb_frame[nb_ptr[n_kf]] and f12 (pairs x 9, required
 * when prm->use_f12) are HOST arrays; three launches for the whole batch.  _dev: outputs stay on the device (words / weights /
 * nodes of keyframe j at [j * kp_cap ..), matches of pair e at [e * kp_cap ..), n_match[e]).  The host variant works on the frames
 * of this context's last extraction call (uco_b200_track_frames / uco_b200_orb_extract_batch) with kp_cap = max_features. */
int uco_b200_keyframes_batch_dev(uco_b200_ctx* ctx, const uco_b200_voc* voc, int bow_level, const uco_keypoint* kps_dev, size_t kps_frame_stride,
                                 const uint8_t* desc_dev, size_t desc_frame_stride, const int32_t* n_kp_dev, int kp_cap, int n_frames,
                                 int n_kf, const int32_t* kf_frame, const int32_t* nb_ptr, const int32_t* nb_frame, const float* f12,
                                 const uco_match_params* prm, uint32_t* word_dev, float* weight_dev, uint32_t* node_dev,
                                 uco_match* match_dev, int32_t* n_match_dev);
int uco_b200_keyframes_batch(uco_b200_ctx* ctx, const uco_b200_voc* voc, int bow_level, int n_kf, const int32_t* kf_frame, const int32_t* nb_ptr,
                             const int32_t* nb_frame, const float* f12, const uco_match_params* prm, uint32_t* word, float* weight, uint32_t* node,
                             uco_match* matches, int32_t* n_matches);

/* FrameMatcher_BoW::matchEpipolar (src/utils/framematcher.cpp:407-541): candidates of a query keypoint are the train keypoints under
 * the same level-3 vocabulary node (Frame::bowvector_level, what uco_b200_bow_transform reports as level_node), then the same
 * filters as above.  A frame's fBow2 is passed flattened in std::map order. */
typedef struct uco_bow_index {
    int32_t n_nodes;
    const uint32_t* node_id;   /* n_nodes, ascending (std::map<uint32_t, std::vector<uint32_t>> iteration order) */
    const int32_t* ptr;        /* n_nodes + 1: the keypoint indices of node i are kp[ptr[i] .. ptr[i+1]) in the vector's order */
    const int32_t* kp;
} uco_bow_index;
/* q_desc / t_desc: one 32-byte row per KEYPOINT (Frame::desc); q_usable / t_usable: isUsed(frame, keypoint, mode) per keypoint or NULL;
 * out: capacity >= number of query entries (q_bow->ptr[n_nodes]); matches in the reference's order (node by node). */
int uco_b200_frame_match_bow(uco_b200_ctx* ctx, const uint8_t* q_desc, size_t q_stride, const uco_keypoint* q_kps, int n_q_kps,
                             const uint8_t* q_usable, const uco_bow_index* q_bow, const uint8_t* t_desc, size_t t_stride,
                             const uco_keypoint* t_kps, int n_t_kps, const uint8_t* t_usable, const uco_bow_index* t_bow,
                             const uco_match_params* prm, uco_match* out, int capacity, int* n_out);

/* ------------------------------------------------------------------------------------------------------------
 * K9  projection matcher: local map points -> keypoints of the current frame
 *   replaces ucoslam::Map::matchFrameToMapPoints(used_frames, curframe, pose_f2g, minDescDist, maxRepjDist, markVisible, ...)
 *     src/map.cpp:651-770 (the per-map-point loop from :684 on and filter_ambiguous_query; gathering the map-point list,
 *     :655-672, is container walking and stays with the caller), src/map_types/frame.cpp:102-115 getKeyPointsInRegion,
 *     src/map_types/frame.h:129-136 predictScale, src/map_types/mappoint.h:99,146-162, src/basictypes/misc.cpp:117-150
 *   The reference's best / second-best bookkeeping depends on the ORDER in which Frame::keypoint_kdtree (picoflann) reports the
 *   keypoints of a region, so the frame's kd-tree is an input: flattened nodes in picoflann's numbering.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_kdnode {   /* one picoflann node (picoflann.h:203-215) */
    float divlow, divhigh;    /* internal node: the two sides of the cut */
    int32_t col;              /* internal node: split dimension 0 / 1; leaf: -1 */
    int32_t left, right;      /* children (node indices), -1 for a leaf */
    int32_t leaf_begin, leaf_count; /* leaf: its keypoint indices are leaf_idx[leaf_begin .. leaf_begin + leaf_count) */
} uco_kdnode;

/* host-side: the tree of n 2-d points (first two floats of records stride_bytes apart, e.g. cv::KeyPoint::pt with stride 28),
 * node for node what picoflann::KdTreeIndex<2>::build gives (mean/variance split, leaf size 10, libstdc++ std::sort for the
 * degenerate cuts).  nodes: capacity cap_nodes >= 2 n; leaf_idx: n entries; bbox4 = {x.min, x.max, y.min, y.max}. */
int uco_b200_kdtree_build(const float* xy, size_t stride_bytes, int n, uco_kdnode* nodes, int cap_nodes, int32_t* leaf_idx,
                          double* bbox4, int* n_nodes);
/* host-side: the same structure from the bytes KdTreeIndex::toStream writes (picoflann.h:603-660; Frame::toStream embeds them) */
int uco_b200_kdtree_parse(const void* bytes, size_t n_bytes, uco_kdnode* nodes, int cap_nodes, int32_t* leaf_idx, int cap_leaf,
                          double* bbox4, int* n_nodes, int* n_leaf);

/* device-side build of the same tree (K16, kdtree.cu): one thread block per frame, node for node what uco_b200_kdtree_build and
 * picoflann::KdTreeIndex::build give (Frame::create_kdtree, src/map_types/frame.h:124-127).  *_batch_dev: everything device
 * resident — keypoints of frame f at kps_dev + f * kps_frame_stride (records; pt is read), n_kp_dev[f] <= cap of them; writes
 * nodes_dev + f * node_cap (node_cap >= 2 * (cap / 5) + 2), leaf_idx_dev + f * cap, bbox_dev + 4 f, n_nodes_dev[f].
 * cap <= UCO_KDTREE_DEV_MAX_POINTS (the build lives in shared memory).  _dev: host buffers in and out (one tree). */
#define UCO_KDTREE_DEV_MAX_POINTS 4096
int uco_b200_kdtree_build_batch_dev(uco_b200_ctx* ctx, int n_frames, const uco_keypoint* kps_dev, size_t kps_frame_stride,
                                    const int32_t* n_kp_dev, int cap, uco_kdnode* nodes_dev, int node_cap, int32_t* leaf_idx_dev,
                                    double* bbox_dev, int32_t* n_nodes_dev);
int uco_b200_kdtree_build_dev(uco_b200_ctx* ctx, const float* xy, size_t stride_bytes, int n, uco_kdnode* nodes, int cap_nodes,
                              int32_t* leaf_idx, double* bbox4, int* n_nodes);
/* host-side probe: the libstdc++ std::sort replay the device build uses for degenerate cuts (csrc/sort_exact.h), on indices keyed
 * by keys[idx] */
int uco_b200_probe_sort_indices(uint32_t* idx, int n, const float* keys);

typedef struct uco_mappoints {      /* the candidate map points in the order of smap_ids (map.cpp:655-672) */
    int32_t n;
    const uint32_t* ids;            /* n        MapPoint::id -> DMatch::trainIdx */
    const float* pos;               /* n x 3    getCoordinates() */
    const float* normal;            /* n x 3    getNormal() */
    const float* min_dist;          /* n        getMinDistanceInvariance() */
    const float* max_dist;          /* n        getMaxDistanceInvariance() */
    const uint8_t* desc;            /* n x 32   the map point's descriptor */
} uco_mappoints;

typedef struct uco_frame_view {
    int32_t n_kp;
    const uco_keypoint* kps;        /* n_kp     Frame::und_kpts (pt and octave are read) */
    const uint8_t* desc;            /* n_kp rows of 32 bytes, Frame::desc */
    size_t desc_stride;             /* bytes between rows (0 = 32) */
    int32_t n_nodes;                /* Frame::keypoint_kdtree, flattened */
    const uco_kdnode* nodes;
    const int32_t* leaf_idx;
    double bbox[4];
    int32_t n_levels;               /* Frame::scaleFactors */
    const float* scale_factors;
    float fx, fy, cx, cy;           /* Frame::imageParams.CameraMatrix */
    float min_xy[2], max_xy[2];     /* Frame::minXY / maxXY */
} uco_frame_view;

/* out: capacity mp->n matches (queryIdx = keypoint, trainIdx = map point id, distance = Hamming), in map-point order like the
 * reference; visible (optional, mp->n bytes): 1 where the reference calls MapPoint::setVisible() (map.cpp:711). */
int uco_b200_match_projected(uco_b200_ctx* ctx, const uco_mappoints* mp, const uco_frame_view* fr, const float* pose_f2g,
                             float min_desc_dist, float max_reproj_dist, uco_match* out, int* n_out, uint8_t* visible);

/* ------------------------------------------------------------------------------------------------------------
 * K17  the tracker's per-frame sequence (track.cu)
 *   uco_b200_track_projected: the search by projection from the PREVIOUS frame, System::_11946837405316294395
 *     (src/utils/system.cpp:5921-6456, macro-obfuscated; called first thing in tracking, :6559): every previous-frame keypoint with a
 *     valid, non-bad map point (prev_mp_row[i] >= 0 = row of that point in `mp`) is projected with the current pose guess
 *     (Frame::project(p, true, true), src/map_types/frame.h:140-161); current keypoints of the SAME octave within
 *     proj_dist_thr * scaleFactors[octave] in kd-tree visit order; best / second best (a new best does not demote the old one);
 *     accepted when best < 0.7 * second; filter_ambiguous_query.  dist_thr = maxDescDistance*1.5 at the reference's call site.
 *     out: capacity n_prev (queryIdx = current keypoint, trainIdx = map point id, distance = Hamming), previous-keypoint order.
 *   uco_b200_track_batch[_dev]: the tracker's main branch (System::_11166622111371682966, system.cpp:6460-6960) for n_frames INDEPENDENT
 *     frames (cameras / streams) in eight launches: kd-trees -> that search (maxDescDistance*1.5) -> solvePnp when > 30 matches
 *     -> (> 30 inliers: pose taken, matched points marked seen, radius 4; else matches dropped, radius projDistThr)
 *     -> Map::matchFrameToMapPoints over the rows flagged mp_local that were not seen (maxDescDistance*2) -> append,
 *     filter_ambiguous_query -> solvePnp.  Frame f's arrays start at f * kp_cap / prev_cap / map_cap records.  The keypoints are
 *     the frame's und_kpts (undistort first: uco_b200_undistort_points_dev).  status[f]: bit 0 = the first search found <= 30
 *     matches (the reference then tries FrameMatcher against the reference keyframe, :6600-6700: the caller's), bit 1 = the
 *     first solvePnp kept <= 30 inliers.  matches[f]: imgIdx = 1 inlier / -1 outlier of the final solvePnp (pnpsolver.cpp:398-404).
 *     Markers are not part of the batch (frames with markers go through uco_b200_pose_only).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_track_params {
    float max_desc_dist;        /* Params::maxDescDistance (50 for ORB, ORBextractor.h:105) */
    float proj_dist_thr;        /* Params::projDistThr (15, ucoslamtypes.cpp:49) */
    float fx, fy, cx, cy, bf;   /* ImageParams; bf = bl * fx (stereo only) */
    float min_xy[2], max_xy[2]; /* Frame::minXY / maxXY */
    int32_t n_levels;
    float scale_factors[UCO_MATCH_MAX_SCALES];
} uco_track_params;

#define UCO_TRACK_NO_SYNC 1     /* _dev only: return after queueing the launches (errors of this call surface with the next sync) */
typedef struct uco_track_batch {
    int32_t n_frames, kp_cap, prev_cap, map_cap, flags;
    const uco_keypoint* kps;    /* n_frames x kp_cap   current frames' und_kpts (pt, octave) */
    const uint8_t* desc;        /* n_frames x kp_cap x 32 */
    const int32_t* n_kp;        /* n_frames */
    const float* depth;         /* n_frames x kp_cap   Frame::depth, NULL = monocular */
    const uco_keypoint* prev_kps;   /* n_frames x prev_cap  previous frames' und_kpts (octave) */
    const uint8_t* prev_desc;       /* n_frames x prev_cap x 32 */
    const int32_t* prev_n_kp;       /* n_frames */
    const int32_t* prev_mp_row;     /* n_frames x prev_cap  row of the keypoint's map point in the frame's map block, -1 = none / bad */
    const int32_t* map_n;           /* n_frames             rows used in each map block */
    const uint32_t* mp_id;          /* n_frames x map_cap */
    const float* mp_pos;            /* x 3 */
    const float* mp_normal;         /* x 3 */
    const float* mp_min_dist;
    const float* mp_max_dist;
    const uint8_t* mp_desc;         /* x 32 */
    const uint8_t* mp_stable;       /* MapPoint::isStable(), NULL = all stable */
    const uint8_t* mp_local;        /* 1 = belongs to the local map searched by matchFrameToMapPoints, NULL = all */
    const float* pose_prior;        /* n_frames x 16        the pose guess (frame <- global) */
} uco_track_batch;

typedef struct uco_track_out {
    uco_match* matches;             /* n_frames x kp_cap */
    int32_t* n_matches;             /* n_frames */
    float* pose;                    /* n_frames x 16 */
    int32_t* n_good;                /* n_frames   return value of the final solvePnp */
    int32_t* status;                /* n_frames */
    int32_t* n_tbp;                 /* n_frames   matches of the first search */
    uint8_t* visible;               /* n_frames x map_cap, optional: MapPoint::setVisible() of map.cpp:711 */
} uco_track_out;

int uco_b200_track_batch_dev(uco_b200_ctx* ctx, const uco_track_batch* in_dev, const uco_track_params* prm, const uco_track_out* out_dev);
int uco_b200_track_batch(uco_b200_ctx* ctx, const uco_track_batch* in_host, const uco_track_params* prm, const uco_track_out* out_host);
/* device-resident mirror of the tracking state of n_streams independent streams (what Frame / Map hold for the
 * tracker — the previous frame's und_kpts / desc / ids, src/map_types/frame.h:60-75, and the local map's points,
 * src/map_types/mappoint.h — kept in HBM between calls).  set_prev / set_map replace one stream's block from host arrays.
 * track_frames: one tracking step of every stream through HOST buffers — images in (n_streams pointers, w x h, row stride),
 * ORB extraction (uco_orb_params), kd-trees and the tracker's sequence on the device, out: the frames' keypoints / descriptors /
 * counts (n_streams x orb->max_features records; kps / desc / n_kp may be NULL) and `out` (rows of orb->max_features matches).
 * track_state_step_dev: the same against keypoints / descriptors already on the device (uco_b200_orb_extract_batch_dev). */
typedef struct uco_b200_track_state uco_b200_track_state;
int uco_b200_track_state_create(uco_b200_ctx* ctx, int n_streams, int prev_cap, int map_cap, uco_b200_track_state** out);
void uco_b200_track_state_free(uco_b200_ctx* ctx, uco_b200_track_state* st);
int uco_b200_track_state_set_prev(uco_b200_ctx* ctx, uco_b200_track_state* st, int stream, int n, const uco_keypoint* kps, const uint8_t* desc,
                                  const int32_t* mp_row);
int uco_b200_track_state_set_map(uco_b200_ctx* ctx, uco_b200_track_state* st, int stream, const uco_mappoints* mp, const uint8_t* stable,
                                 const uint8_t* local);
int uco_b200_track_state_step_dev(uco_b200_ctx* ctx, const uco_b200_track_state* st, const uco_keypoint* kps_dev, const uint8_t* desc_dev,
                                  const int32_t* n_kp_dev, int kp_cap, const float* pose_prior_dev, const uco_track_params* prm,
                                  const uco_track_out* out_dev, int flags);
int uco_b200_track_frames(uco_b200_ctx* ctx, const uco_b200_track_state* st, const uint8_t* const* imgs, int w, int h, size_t stride,
                          const uco_orb_params* orb, const uco_track_params* prm, const float* pose_prior, uco_keypoint* kps, uint8_t* desc,
                          int32_t* n_kp, const uco_track_out* out);
int uco_b200_track_projected(uco_b200_ctx* ctx, int n_prev, const uco_keypoint* prev_kps, const uint8_t* prev_desc,
                             const int32_t* prev_mp_row, const uco_mappoints* mp, const uco_frame_view* fr, const float* pose_f2g,
                             float dist_thr, float proj_dist_thr, uco_match* out, int* n_out);

/* ------------------------------------------------------------------------------------------------------------
 * K11  keyframe database: relocalisation / loop-closure candidates by bag of words (BASELINE config 4)
 *   replaces ucoslam::KeyFrameDataBase (KPFrameDataBase) add / del / clear / size / isId / relocalizationCandidates
 *     src/map_types/keyframedatabase.h:31-52, src/map_types/keyframedatabase.cpp:136-171 (add/del/clear),
 *     :195-233 (votes per frame over the inverted word index, maxCommonWords, minCommonWords = max*0.8f, fbow::fBow::score
 *     3rdparty/fbow/fbow/fbow.cpp:192-243 of the frames above it, kept if > minScore), :236-275 (covisibility accumulation over
 *     the 10 best neighbours, 0.75*best gate, optional sort by decreasing accumulated score -> uco_b200_kfdb_rank, host only).
 *   A frame's bag of words is passed as its fBow in map order: `words` strictly ascending, `weights` the float values
 *   (what uco_b200_bow_transform + the host fold produce).  The database is device resident (one context's device).
 *   query: out_* hold the scored frames (the reference's frame_score map) in ascending frame id, capacity `cap`
 *   (UCO_E_CAPACITY with *n_out = needed otherwise); scores are bit-identical doubles; out_common (optional) = votes of
 *   each returned frame; max_common (optional) = maxCommonWords.  excluded = the reference's excludedFrames (unknown ids ignored).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_b200_kfdb uco_b200_kfdb;
int uco_b200_kfdb_create(uco_b200_ctx* ctx, uco_b200_kfdb** out);
void uco_b200_kfdb_free(uco_b200_ctx* ctx, uco_b200_kfdb* db);
int uco_b200_kfdb_clear(uco_b200_ctx* ctx, uco_b200_kfdb* db);
int uco_b200_kfdb_size(const uco_b200_kfdb* db, uint32_t* n_frames, uint64_t* n_words);
int uco_b200_kfdb_has(const uco_b200_kfdb* db, uint32_t frame_id); /* KeyFrameDataBase::isId */
int uco_b200_kfdb_add(uco_b200_ctx* ctx, uco_b200_kfdb* db, uint32_t frame_id, const uint32_t* words, const float* weights, int n);
/* several frames in one call: counts[f] words each, concatenated in `words` / `weights` */
int uco_b200_kfdb_add_batch(uco_b200_ctx* ctx, uco_b200_kfdb* db, int n_frames, const uint32_t* frame_ids, const int32_t* counts,
                            const uint32_t* words, const float* weights);
int uco_b200_kfdb_del(uco_b200_ctx* ctx, uco_b200_kfdb* db, uint32_t frame_id);
int uco_b200_kfdb_query(uco_b200_ctx* ctx, uco_b200_kfdb* db, const uint32_t* words, const float* weights, int n,
                        const uint32_t* excluded, int n_excluded, float min_score, uint32_t* out_frame, double* out_score,
                        uint32_t* out_common, int cap, int* n_out, uint32_t* max_common);
/* device time (ms) of the vote scan [0] and the scoring pass [1] of the LAST query (context profiling on) */
int uco_b200_kfdb_last_ms(const uco_b200_kfdb* db, float* out2);
/* steps 3-4 on the scored frames (ascending ids, as uco_b200_kfdb_query returns them).  nbr_off[n+1] / nbr: for scored frame i
 * the neighbour ids CovisGraph::getNeighborsWeights(frame[i], true) returns (decreasing weight; the first 10 are used).
 * out: capacity n. */
int uco_b200_kfdb_rank(const uint32_t* frame, const double* score, int n, const int32_t* nbr_off, const uint32_t* nbr, int sorted,
                       float min_score, uint32_t* out, int* n_out);

/* ------------------------------------------------------------------------------------------------------------
 * K12  stereo depth association of a rectified pair (BASELINE config 3)
 *   replaces the association loop of ucoslam::FrameExtractor::processStereo
 *     src/utils/frameextractor.cpp:1410-2634 (macro-obfuscated; see csrc/stereo.cu for the de-obfuscated statement list):
 *     right keypoints bucketed by round(y); per left keypoint the first candidate of least Hamming distance among those not to its
 *     right, at most one octave apart and closer than Params::maxDescDistance; 6x6 SAD over offsets -7..7, parabola refinement,
 *     depth = bl*fx / disparity (Frame::depth, 0 = none).
 *   img_l / img_r: the two grey images the keypoints were extracted from (w x h); kps_l = the left frame's UNDISTORTED keypoints
 *   (Frame::und_kpts), kps_r = the right image's raw detections; desc_*: 32-byte rows.
 *   depth: n_l floats; match_r (optional): index of the associated right keypoint or -1 (also set when the SAD stage then rejects
 *   the pair); n_with_depth (optional): number of keypoints that received a depth.
 *   Where the reference would throw (cv::Mat ROI outside the right image: a matched right keypoint within 10 px of the border,
 *   impossible for ORB keypoints) the call fails with UCO_E_INVALID.
 *   _dev: device pointers, dense 32-byte descriptor rows, asynchronous; counters_dev = 2 int32 (keypoints with depth, error flag).
 * ---------------------------------------------------------------------------------------------------------- */
int uco_b200_stereo_depth(uco_b200_ctx* ctx, const uint8_t* img_l, size_t stride_l, const uint8_t* img_r, size_t stride_r, int w, int h,
                          const uco_keypoint* kps_l, const uint8_t* desc_l, size_t desc_l_stride, int n_l, const uco_keypoint* kps_r,
                          const uint8_t* desc_r, size_t desc_r_stride, int n_r, float max_desc_dist, float bl, float fx, float* depth,
                          int32_t* match_r, int* n_with_depth);
int uco_b200_stereo_depth_dev(uco_b200_ctx* ctx, const uint8_t* img_l_dev, size_t pitch_l, const uint8_t* img_r_dev, size_t pitch_r,
                              int w, int h, const uco_keypoint* kps_l_dev, const uint8_t* desc_l_dev, int n_l,
                              const uco_keypoint* kps_r_dev, const uint8_t* desc_r_dev, int n_r, float max_desc_dist, float bl, float fx,
                              float* depth_dev, int32_t* match_dev, int32_t* counters_dev);

/* ------------------------------------------------------------------------------------------------------------
 * K13  two-view triangulation with the reference's acceptance gates (new-map-point creation)
 *   replaces ucoslam::Triangulate(Train, Query, RT_Q2T, matches, maxChi2)   src/basictypes/misc.cpp:921-1040
 *   (and the body of triangulate_ :1042-1160), called from src/utils/mapmanager.cpp:10093 and mapinitializer.cpp:1574.
 *   kps_train / kps_query: Frame::und_kpts of the two frames; matches: cv::DMatch records (trainIdx -> camera 1 = K_train [I|0],
 *   queryIdx -> camera 2 = K_query [R|t] with RT = the 4x4 row-major transform camera 1 -> camera 2);
 *   scale_factors_*: Frame::scaleFactors.  xyz: 3 floats per match in camera-1 coordinates, NaN NaN NaN where the reference
 *   rejects (parallax cosine outside [0, 0.9998], w == 0, non-finite, behind either camera, reprojection chi2 > max_chi2).
 *   Floating point: the reference takes the null vector from OpenCV's float SVD; here it is computed in double, so points agree
 *   to float-SVD accuracy (tolerance stated in tests/test_triangulate_gpu.py), not bit for bit.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_triangulate_params {
    float K_train[4];             /* fx fy cx cy */
    float K_query[4];
    float RT[16];
    int32_t n_levels_train; const float* scale_factors_train;
    int32_t n_levels_query; const float* scale_factors_query;
    float max_chi2;               /* 5.998 default, misc.h:65 */
    /* the mapper's scale-consistency test on the surviving points (new-map-point creation, src/utils/mapmanager.cpp:9772-10788,
     * de-obfuscated): ratioDist = |P - C_train| / |P - C_query|, ratioOctave = sf_train[oct] / sf_query[oct]; rejected if
     * ratioDist * f < ratioOctave or ratioDist > ratioOctave * f with f = 1.5f * Params::scaleFactor, or if a distance is 0.
     * 0 = test off (plain ucoslam::Triangulate). */
    float scale_ratio_factor;
    /* when non-zero the accepted points are returned as g2f_train * p (the mapper's `pose_f2g.inv() * p`, Se3Transform float
     * arithmetic) instead of camera-1 coordinates */
    int32_t to_global;
    float g2f_train[16];
} uco_triangulate_params;
int uco_b200_triangulate(uco_b200_ctx* ctx, const uco_keypoint* kps_train, int n_train, const uco_keypoint* kps_query, int n_query,
                         const uco_match* matches, int n_matches, const uco_triangulate_params* prm, float* xyz, int* n_good);

/* New-map-point creation of one keyframe as a unit:
 *   replaces MapManager::createNewPoints(frame, nn, maxPoints)   src/utils/mapmanager.cpp:9772-10788 (macro-obfuscated; de-obfuscated
 *   with the C preprocessor): FrameMatcher::setParams(frame, MODE_UNASSIGNED, maxDescDistance*2, 0.6, true, INT_MAX), then per neighbour
 *   keyframe f of the covisibility graph (the OpenMP loop :9992): T_f = nb.pose_f2g * frame.pose_f2g.inv(), matchEpipolar(nb,
 *   MODE_UNASSIGNED, T_f), Triangulate(frame, nb, T_f, matches), p = frame.pose_f2g.inv() * xyz, the scale-consistency gate
 *   (ratioDist vs ratioOctave, factor 1.5f * scaleFactor), and the merge into one NewPointInfo per keyframe keypoint.
 *   Inputs as uco_b200_frame_match_multi (t_* = the keyframe with t_map = its unassigned keypoints, q_*[f] = the neighbours with
 *   q_map[f]); f12: n_frames x 9 (computeF12 of the caller, as for the matcher); rt: n_frames x 16, T_f row-major (keyframe camera ->
 *   neighbour camera); K_nb: n_frames x 4 (fx fy cx cy).
 *   Output, merged: point j < *n_points belongs to keyframe keypoint pt_kpt[j] (ascending: std::map order; after a max_points cut in
 *   std::sort order of the descriptor distance), pt_xyz / pt_dist = position (global) and descriptor distance taken from the LAST
 *   neighbour that saw it (what the reference's minimum-octave loop returns: its minimum is never updated), observations
 *   obs_ptr[j] .. obs_ptr[j+1]: (obs_frame = neighbour index f, obs_kpt = its keypoint) in neighbour order; the keyframe's own
 *   observation (frame.idx, pt_kpt[j]) is implied.  Optional per-neighbour output: matches[f] / n_matches[f] (the matcher's lists) and
 *   xyz[f] (3 floats per match, global, NaN = rejected); pass NULL to skip.
 *   One upload, three launches (k-NN, filters, triangulation of all neighbours), one download; the merge runs on the host.
 *   Parity: matches bit-exact (K8); points to the float tolerance of K13 (tests/test_new_points_gpu.py). */
typedef struct uco_new_points_params {
    uco_match_params match;       /* min_desc_dist = maxDescDistance*2, nn_match_ratio 0.6, check_orientation 1, max_octave_diff INT_MAX, use_f12 1 */
    float K_kf[4];                /* fx fy cx cy of the new keyframe */
    float g2f_kf[16];             /* frame.pose_f2g.inv(), row-major 4x4 (keyframe camera -> global) */
    int32_t n_levels_kf; const float* scale_factors_kf;
    int32_t n_levels_nb; const float* scale_factors_nb;   /* the neighbours' (one extractor configuration per map) */
    float max_chi2;               /* Triangulate's default 5.998 (misc.h:65) */
    float scale_ratio_factor;     /* 1.5f * Params::scaleFactor */
    int32_t max_points;           /* maxPoints (< 0: no limit) */
} uco_new_points_params;
int uco_b200_new_points(uco_b200_ctx* ctx, const uint8_t* t_desc, int nt, size_t t_stride, const uco_keypoint* t_kps, int n_t_kps, const int32_t* t_map,
                        int n_frames, const uint8_t* const* q_desc, const int32_t* nq, size_t q_stride, const uco_keypoint* const* q_kps,
                        const int32_t* n_q_kps, const int32_t* const* q_map, const float* f12, const float* rt, const float* K_nb,
                        const uco_new_points_params* prm, int32_t* n_points, int32_t* pt_kpt, float* pt_xyz, float* pt_dist, int32_t* obs_ptr,
                        int32_t* obs_frame, int32_t* obs_kpt, int capacity_points, int capacity_obs, uco_match* const* matches, int32_t* n_matches,
                        float* const* xyz);

/* ------------------------------------------------------------------------------------------------------------
 * Frame streams and the device-resident Frame mirror
 *   replaces / reads Frame::toStream / fromStream   src/map_types/frame.cpp:260-341 — the unit of .map / .slm files (Map::toStream,
 *   src/map.cpp:316-352, writes its keyframes with it; System::saveToFile wraps that) and of the tracker -> mapper queue — together
 *   with the streams of its members: cv::Mat, std::vector (src/basictypes/io_utils.{h,cpp}), MarkerObservation / MarkerPosesIPPE
 *   (src/map_types/marker.cpp:94-113, frame.cpp:372-386), Se3Transform (se3transform.h:178-188), fbow::fBow / fBow2
 *   (3rdparty/fbow/fbow/fbow.cpp:261-303), ImageParams (src/imageparams.cpp:68-84), picoflann::KdTreeIndex (picoflann.h:603-660).
 *   uco_b200_frame_stream_parse gives a VIEW into the caller's bytes (nothing is copied; variable-size members stay opaque byte
 *   ranges: markers, bow_level, kdtree); uco_b200_frame_stream_write produces byte for byte what the reference's toStream writes
 *   for the same field values (out == NULL: size only); uco_b200_kdtree_serialize writes the tree member from a flattened tree.
 *   uco_b200_frame_upload puts what the *_dev entry points consume on the device in one allocation (keypoints, descriptors, map-point
 *   ids, flags, depth, pose, scale factors, the kd-tree flattened as uco_kdnode) so that matcher / tracker / mapper kernels chain on a
 *   keyframe of a loaded map without host vectors in between; uco_b200_frame_dev exposes the device pointers.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct uco_mat_view { int32_t rows, cols, type; const uint8_t* data; } uco_mat_view;   /* OpenCV type code; rows without padding */
typedef struct uco_frame_stream {
    uint32_t idx, fseq_idx; uint8_t frame_flags; int8_t kp_desc_type;        /* DescriptorTypes::Type : int8 (1 = ORB) */
    uco_mat_view desc;
    uint32_t n_und_kpts; const uco_keypoint* und_kpts;
    uint32_t n_kpts; const float* kpts;                                       /* cv::Point2f */
    uint32_t n_depth; const float* depth;
    uint32_t n_ids; const uint32_t* ids;
    uint32_t n_flags; const uint8_t* flags;
    uint32_t n_markers; const uint8_t* markers; uint64_t markers_bytes;       /* the MarkerObservation records as streamed */
    float pose_f2g[16];
    uint32_t n_bow; const uint8_t* bow;                                       /* (uint32 word, float weight) pairs */
    uint32_t n_bow_level; const uint8_t* bow_level; uint64_t bow_level_bytes; /* (uint32 node, uint32 count, count x uint32 keypoint) records */
    uint32_t n_scale_factors; const float* scale_factors;
    uco_mat_view camera_matrix, distortion; int32_t cam_size[2]; float bl, rgb_depthscale;
    uco_mat_view image;
    const uint8_t* kdtree; uint64_t kdtree_bytes;                             /* KdTreeIndex::toStream bytes (uco_b200_kdtree_parse reads them) */
    int32_t min_xy[2], max_xy[2];                                             /* cv::Point */
} uco_frame_stream;
int uco_b200_frame_stream_parse(const uint8_t* bytes, size_t len, uco_frame_stream* view, size_t* consumed);
int uco_b200_frame_stream_write(const uco_frame_stream* view, uint8_t* out, size_t cap, size_t* written);
int uco_b200_kdtree_serialize(const uco_kdnode* nodes, int n_nodes, const int32_t* leaf_idx, const double* bbox4, int n_values, const double* div_val,
                              uint8_t* out, size_t cap, size_t* written);
/* MapPoint::toStream / fromStream (src/map_types/mappoint.cpp:117-175), the other record type of a map file: a view and its writer.
 * The position / normal / distance range / descriptor of the view are what uco_mappoints (K9, K17) holds per row. */
typedef struct uco_mappoint_stream {
    uint32_t id; float pos3d[3];
    uco_mat_view desc;                       /* 1 x 32 bytes for ORB */
    uint32_t n_frames; const uint32_t* frames;   /* (keyframe idx, keypoint index) pairs in std::map order */
    float normal[3];
    uint16_t n_times_seen, n_times_visible; uint8_t flags;
    float max_distance, min_distance;
    uint64_t kf_since_addition; uint32_t last_fidx_seen;
} uco_mappoint_stream;
int uco_b200_mappoint_stream_parse(const uint8_t* bytes, size_t len, uco_mappoint_stream* view, size_t* consumed);
int uco_b200_mappoint_stream_write(const uco_mappoint_stream* view, uint8_t* out, size_t cap, size_t* written);
/* The map-point SECTION of a map file: ReusableContainer<MapPoint>::toStream (src/basictypes/reusablecontainer.h:276-291) over
 * ExpansibleContainer<Pair>::toStream (expansiblecontainer.h:115-129), as Map::toStream writes it (src/map.cpp:316-325).  walk: one pass over the
 * section; slot_offset[i] = byte offset of slot i's MapPoint stream (uco_b200_mappoint_stream_parse reads it), slot_valid[i] = its valid flag
 * (either may be NULL; cap = their length; UCO_E_CAPACITY when the section has more slots, the header is filled anyway).
 * from_container: the valid points as the flat rows of uco_mappoints / uco_b200_track_state_set_map (a device-resident map from a map file without
 * building MapPoint objects).  write: the inverse of walk (n_slots a multiple of 200; unused slots: uco_b200_mappoint_stream_default). */
typedef struct uco_mappoint_container {
    uint32_t n_slots, n_used, n_valid;            /* capacity (chunks x 200), ExpansibleContainer::size(), slots whose flag is set */
    uint32_t n_free; const uint32_t* free_slots;  /* _emptySpaces, in the order the slots were freed (insert() reuses the last one) */
} uco_mappoint_container;
int uco_b200_mappoint_container_walk(const uint8_t* bytes, size_t len, uco_mappoint_container* c, size_t* slot_offset, uint8_t* slot_valid, uint32_t cap,
                                     size_t* consumed);
/* the keyframe section: FrameSet::toStream (src/map_types/frame.cpp:350-355) = int magic 88888 + the same container over Frame streams */
int uco_b200_frame_container_walk(const uint8_t* bytes, size_t len, uco_mappoint_container* c, size_t* slot_offset, uint8_t* slot_valid, uint32_t cap,
                                  size_t* consumed);
/* the keyframe section from already-serialised Frame streams; unused slots take the stream of a default-constructed Frame (copy one from any
 * reference-written keyframe section) */
int uco_b200_frame_container_write(const uco_mappoint_container* c, const uint8_t* const* slot_bytes, const size_t* slot_len, const uint8_t* valid, uint8_t* out,
                                   size_t cap, size_t* written);
int uco_b200_mappoints_from_container(const uint8_t* bytes, size_t len, uint32_t cap, uint32_t* ids, float* pos, float* normal, float* min_dist, float* max_dist,
                                      uint8_t* desc, uint8_t* flags, uint32_t* n_out, size_t* consumed);
int uco_b200_mappoint_container_write(const uco_mappoint_container* c, const uco_mappoint_stream* points, const uint8_t* valid, uint8_t* out, size_t cap,
                                      size_t* written);
void uco_b200_mappoint_stream_default(uco_mappoint_stream* v);
/* The remaining sections of Map::toStream (src/map.cpp:316-325) and the whole stream.  Views: pointers and offsets into the caller's buffer.
 *   keyframe database  KeyFrameDataBase::toStream (keyframedatabase.cpp:335-340, :278-287, :122-124; vocabulary: fbow.cpp:171-178)
 *   markers            toStream__kv_complex over std::map<u32, Marker> (io_utils.h:113-121; Marker::toStream, marker.cpp:41-47)
 *   covisibility graph CovisGraph::toStream (covisgraph.cpp:308-333) */
typedef struct uco_kfdb_stream {
    int32_t type;                              /* 0 DummyDataBase, 1 KPFrameDataBase */
    size_t voc_off, voc_len;                   /* the fbow vocabulary stream (what uco_b200_bow_load takes); 0 / 0 for type 0 */
    uint32_t n_words; size_t words_off;        /* inverted index: n_words records {u32 word, u32 count, count x u32 frame} from words_off */
    uint64_t n_word_frames;                    /* total (word, frame) entries */
    uint32_t n_frames; const uint32_t* frames; /* the database's frame ids, ascending */
} uco_kfdb_stream;
typedef struct uco_marker_stream {
    uint32_t key, id; float pose_g2m[16]; float size;
    uint32_t n_frames; const uint32_t* frames;
    uint32_t dict_len; const char* dict;       /* not NUL-terminated */
} uco_marker_stream;
typedef struct uco_covis_stream {
    uint32_t n_nodes; const uint32_t* nodes;
    uint32_t n_adj; size_t adj_off;            /* n_adj records {u32 node, u32 count, count x u32 neighbour} from adj_off */
    uint64_t n_neighbours;
    uint32_t n_weights; const uint8_t* weights; /* n_weights packed records {u64 key = join(a, b), float weight}, 12 bytes each */
} uco_covis_stream;
typedef struct uco_map_sections {              /* byte ranges of the five sections, in stream order */
    size_t kfdb_off, kfdb_len; uco_kfdb_stream kfdb;
    size_t points_off, points_len; uco_mappoint_container points;
    size_t markers_off, markers_len; uint32_t n_markers;
    size_t frames_off, frames_len; uco_mappoint_container frames;
    size_t covis_off, covis_len; uco_covis_stream covis;
    size_t total_len;
} uco_map_sections;
int uco_b200_kfdb_stream_walk(const uint8_t* bytes, size_t len, uco_kfdb_stream* out, size_t* consumed);
int uco_b200_marker_map_walk(const uint8_t* bytes, size_t len, uint32_t cap, uco_marker_stream* out, uint32_t* n_out, size_t* consumed);
int uco_b200_covis_stream_walk(const uint8_t* bytes, size_t len, uco_covis_stream* out, size_t* consumed);
/* writers / unpackers of the three sections (keyed lists in CSR form: key[i] owns values[ptr[i] .. ptr[i+1]); ptr has n + 1 entries); with the Frame,
 * MapPoint and container writers a whole map file can be written: u64 225237123, then the five sections in the order above */
int uco_b200_marker_map_write(const uco_marker_stream* markers /* ascending key */, uint32_t n, uint8_t* out, size_t cap, size_t* written);
int uco_b200_covis_stream_unpack(const uint8_t* bytes, size_t len, const uco_covis_stream* view, uint32_t* adj_node, uint32_t* adj_ptr, uint32_t* adj_idx,
                                 uint64_t* w_key, float* w);
int uco_b200_covis_stream_write(uint32_t n_nodes, const uint32_t* nodes, uint32_t n_adj, const uint32_t* adj_node, const uint32_t* adj_ptr, const uint32_t* adj_idx,
                                uint32_t n_weights, const uint64_t* w_key, const float* w, uint8_t* out, size_t cap, size_t* written);
int uco_b200_kfdb_stream_unpack(const uint8_t* bytes, size_t len, const uco_kfdb_stream* view, uint32_t* word, uint32_t* word_ptr, uint32_t* word_frames);
int uco_b200_kfdb_stream_write(int32_t type, const uint8_t* voc, size_t voc_len, uint32_t n_words, const uint32_t* word, const uint32_t* word_ptr,
                               const uint32_t* word_frames, uint32_t n_frames, const uint32_t* frames, uint8_t* out, size_t cap, size_t* written);
/* has_file_magic: the buffer is a map FILE (Map::saveToFile, map.cpp:339-345: u64 225237123 first); section offsets are from the buffer start */
int uco_b200_map_stream_walk(const uint8_t* bytes, size_t len, int has_file_magic, uco_map_sections* out);
typedef struct uco_frame_dev {   /* DEVICE pointers (one allocation, owned by the uco_b200_frame) + the small host-side members */
    uint32_t idx, fseq_idx; int32_t n_kp;
    const uco_keypoint* kps; const uint8_t* desc; const uint32_t* ids; const uint8_t* flags; const float* depth;   /* depth: NULL without depth */
    int32_t n_nodes; const uco_kdnode* nodes; int32_t n_leaf; const int32_t* leaf_idx; const double* bbox;
    int32_t n_scale_factors; const float* scale_factors; const float* pose_f2g;
    double bbox_host[4]; float pose_host[16]; float K[4]; int32_t min_xy[2], max_xy[2];
} uco_frame_dev;
typedef struct uco_b200_frame uco_b200_frame;
int uco_b200_frame_upload(uco_b200_ctx* ctx, const uco_frame_stream* view, uco_b200_frame** out);
const uco_frame_dev* uco_b200_frame_dev(const uco_b200_frame* frame);
int uco_b200_frame_download(uco_b200_ctx* ctx, const uco_b200_frame* frame, uco_keypoint* kps, uint8_t* desc, uint32_t* ids, uint8_t* flags, float* depth);
void uco_b200_frame_free(uco_b200_frame* frame);

/* ------------------------------------------------------------------------------------------------------------
 * K14  RANSAC pose from 2D-3D matches (relocalisation / loop-closure candidate scoring)
 *   replaces ucoslam::PnPSolver::solvePnPRansac(frame, map, matches_io, posef2g_io, maxIters)
 *     src/optimization/pnpsolver.cpp:36-114 (4-match samples, cv::solvePnP P3P hypothesis, float reprojection test < 5.99 px^2,
 *     MapPoint::getViewCos >= 0.5, first iteration with the most inliers wins, fewer than 4 inliers -> false).
 *   Per match j: p3d = MapPoint::getCoordinates(), normals = MapPoint::getNormal(), p2d = frame.und_kpts[queryIdx].pt;
 *   cam = fx fy cx cy of the frame.  samples: max_iters x 4 match indices (the caller's random stream), or NULL to draw them from
 *   the counter-based generator seeded with `seed` (the reference's std::random_shuffle / rand() stream is not reproduced).
 *   Result: *n_inliers = 0 means `false` (pose untouched); otherwise pose44 = the winning float 4x4 (world -> camera, what
 *   Se3Transform(rv,tv) holds), inliers = indices of the matches to keep (ascending), best_iter / counts (optional, max_iters;
 *   -1 where P3P had no solution) for inspection.
 *   uco_b200_probe_p3p: host-only, one hypothesis from 4 correspondences exactly as the kernel forms it (returns 1 / 0).
 * ---------------------------------------------------------------------------------------------------------- */
int uco_b200_pnp_ransac(uco_b200_ctx* ctx, const float* p3d, const float* p2d, const float* normals, int n, const float* cam_fxfycxcy,
                        int max_iters, const int32_t* samples, uint64_t seed, float* pose44, int32_t* inliers, int* n_inliers,
                        int32_t* counts, int* best_iter);
int uco_b200_probe_p3p(const double* X4x3, const double* px4x2, const double* K_fxfycxcy, double* R9, double* t3);

/* ------------------------------------------------------------------------------------------------------------
 * K15  point undistortion when a Frame is built (Frame::und_kpts, marker corners, image bounds)
 *   replaces ucoslam::undistortPoints(points_io, ImageParams)   src/basictypes/misc.cpp:269-292
 *   (cv::undistortPoints with its default 5 iterations, then x*fx+cx in float).
 *   pts / out: n points of two floats, `stride` bytes apart (8 for cv::Point2f arrays, 28 for the pt field of cv::KeyPoint
 *   arrays -- in place is allowed); K = fx fy cx cy; dist: the n_dist (0..12; 14 with zero tilt terms) coefficients of
 *   ImageParams::Distorsion (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]]).
 *   uco_b200_probe_undistort: host-only, the same arithmetic compiled for the host (dense points).
 * ---------------------------------------------------------------------------------------------------------- */
int uco_b200_undistort_points(uco_b200_ctx* ctx, const float* pts, size_t in_stride, int n, const float* K_fxfycxcy, const float* dist,
                              int n_dist, float* out, size_t out_stride);
int uco_b200_undistort_points_dev(uco_b200_ctx* ctx, const float* pts_dev, size_t in_stride, int n, const float* K_fxfycxcy, const float* dist,
                                  int n_dist, float* out_dev, size_t out_stride);
int uco_b200_probe_undistort(const float* pts, int n, const float* K_fxfycxcy, const float* dist, int n_dist, float* out);

#ifdef __cplusplus
}
#endif
#endif /* UCOSLAM_B200_H */
