// common.cuh — context object, error plumbing and workspace helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include "../../include/ucoslam_b200.h"
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: a no-op unless a tool (nsys / ncu --nvtx) injects itself
// NVTX range named after the C-ABI entry point for the duration of the call (the reference marks the same boundaries with its
// __UCOSLAM_ADDTIMER__ / __UCOSLAM_TIMER_EVENT__ debug timers, e.g. globaloptimizer_g2o.cpp:411-417, system.cpp's per-frame events)
struct UcoRange {
    explicit UcoRange(const char* name) { nvtxRangePushA(name); }
    ~UcoRange() { nvtxRangePop(); }
    UcoRange(const UcoRange&) = delete;
    UcoRange& operator=(const UcoRange&) = delete;
};
#define UCO_RANGE() UcoRange uco_range_(__func__)

struct uco_dev_buf {
    void* p = nullptr;
    size_t cap = 0;
};

struct uco_orb_state;   // orb.cu
struct uco_ba_state;    // ba.cu

struct uco_b200_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t sleep_event = nullptr;  // blocking-sync event: long waits (a BA batch) sleep instead of spinning a host core
    std::string err;
    uint64_t launches = 0;
    int profiling = 0;  // record per-stage CUDA events
    // grow-only device / pinned workspaces, keyed by slot
    std::vector<uco_dev_buf> dev;
    std::vector<uco_dev_buf> pin;
    uco_orb_state* orb = nullptr;
    uco_ba_state* ba = nullptr;
    int ba_mode = 0;          // 0 auto, 1 streamed kernels (ba.cu), 2 cluster-resident kernel (ba_cluster.cu)
    int ba_cluster_size = 0;  // CTAs per cluster of the cluster-resident solver (0 = default 8)
    void* track_err_dev = nullptr;
    int* kf_err_dev = nullptr;          // error word of the last keyframes_batch_dev launch sequence (match.cu)
    cudaEvent_t stage_event = nullptr;  // recorded after an un-synchronised upload from a pinned staging buffer: the next writer waits on it  // error words of the last track_batch_dev launch sequence (track.cu)
    int ba_host_threads = 0;  // worker threads of the host-side planner per batch call (0 = this process's share of the cores)
};

enum {  // device workspace slots
    WS_KNN_Q = 0, WS_KNN_T, WS_KNN_IDX, WS_KNN_DIST,
    WS_BOW_DESC, WS_BOW_OUT,
    WS_GENERIC0, WS_GENERIC1, WS_GENERIC2, WS_GENERIC3,
    WS_BA, WS_BA_OUT, WS_BA_STOP,
    WS_PNP_IN, WS_PNP_OUT, WS_PNP_SCRATCH,
    WS_MATCH_KNN, WS_MATCH_SCRATCH, WS_MATCH_IN, WS_MATCH_OUT,
    WS_KNN_N, WS_KNN_MERGE, WS_PROJ, WS_PROJ_OUT,
    WS_KFDB_STAGE, WS_KFDB_Q, WS_KFDB_BITMAP, WS_KFDB_OUT,
    WS_STEREO_SOA, WS_STEREO_IN, WS_STEREO_OUT,
    WS_TRI_IN, WS_TRI_OUT,
    WS_RANSAC_IN, WS_RANSAC_OUT,
    WS_UNDISTORT,
    WS_KDTREE, WS_KDTREE_ERR,
    WS_TRACK, WS_TRACK_ERR, WS_TRACK_IN, WS_TRACK_OUT,
    WS_COUNT
};

int uco_fail(uco_b200_ctx* ctx, int code, const char* fmt, ...);
// returns device pointer with at least `bytes` capacity for the slot (grow-only), nullptr on failure (error set)
void* uco_ws(uco_b200_ctx* ctx, int slot, size_t bytes);
void* uco_pinned(uco_b200_ctx* ctx, int slot, size_t bytes);
// wait for everything queued on the context stream with the calling thread ASLEEP (for waits of milliseconds: several mapper
// threads spinning in cudaStreamSynchronize starve the tracker thread on a host with few cores per GPU)
cudaError_t uco_sleep_sync(uco_b200_ctx* ctx);

// comm.cu: collectives on a stream; a null communicator (or world 1) degenerates to a local copy
struct uco_b200_comm;
int uco_comm_rank(const uco_b200_comm* c);
int uco_comm_world(const uco_b200_comm* c);
int uco_comm_allreduce(uco_b200_comm* c, const void* send, void* recv, size_t count, int op /*0 f64 sum, 1 f64 max, 2 u8 sum, 3 i32 sum*/, cudaStream_t s);
int uco_comm_allgather(uco_b200_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t s);

void uco_orb_state_free(uco_b200_ctx* ctx);
void uco_ba_state_free(uco_b200_ctx* ctx);

#define UCO_CUDA(ctx, call)                                                                           \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return uco_fail(ctx, UCO_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                            cudaGetErrorString(e__));                                                 \
    } while (0)

#define UCO_LAUNCH_CHECK(ctx)                                                                         \
    do {                                                                                              \
        (ctx)->launches++;                                                                            \
        cudaError_t e__ = cudaGetLastError();                                                         \
        if (e__ != cudaSuccess)                                                                       \
            return uco_fail(ctx, UCO_E_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,         \
                            cudaGetErrorString(e__));                                                 \
    } while (0)
