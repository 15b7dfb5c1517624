// knn.cu — K7: brute-force 256-bit Hamming k-NN, exact emulation of xflann's linear index.
//
// Replaces (reference, relative to /root/reference):
//   3rdparty/xflann/xflann/impl/linear.h:68-88      scan of all train rows in index order per query
//   3rdparty/xflann/xflann/impl/resultset.h:64-140  bounded max-heap with strict '<' replacement
//   3rdparty/xflann/xflann/impl/distances.h:279-283 4 x popcount64 distance
//   3rdparty/xflann/xflann/index.h:119-133          optional exchange sort of each result row
//
// Layout: train descriptors are dense 32-byte rows in HBM.  A CTA of KNN_WARPS warps owns
// KNN_WARPS*2 queries; the train set streams through a KNN_STAGES-deep ring of shared-memory tiles filled by
// 1-D TMA bulk copies (cp.async.bulk, completion on an mbarrier), so every train byte is read from L2/HBM once per
// CTA with fully coalesced 128-B lines.  Inside a warp each lane owns one train row of the current 32-row
// chunk and computes two distances (two queries held in registers, 8 x __popc each).  The result heap of
// the reference is replayed EXACTLY: candidates that can enter the heap (d < current worst) are found with a
// warp ballot and inserted by lane 0 in train-index order, which is the order the reference's scalar loop
// pushes them, so the final array order (the "heap order" the tracker consumes, framematcher.cpp:239-270)
// is identical, including all tie cases.
#include "common.cuh"
#include <climits>

#define KNN_WARPS 8
#define KNN_QPW 2                 // queries per warp
#define KNN_QPB (KNN_WARPS * KNN_QPW)
#define KNN_TILE_ROWS 256         // train rows per shared-memory tile (8 KB)
#define KNN_STAGES 4

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- exact replay of xflann::impl::ResultSet (resultset.h:64-140) on one row of packed (dist<<32 | idx) -----------
struct Heap {
    unsigned long long* a;  // shared memory, cap entries
    int n;
    int cap;
};
__device__ __forceinline__ int hdist(unsigned long long v) { return (int)(v >> 32); }

__device__ __forceinline__ void heap_sift_to_root(Heap& h, int index) {  // reference name: down()
    while (index > 0) {
        int parent = (index - 1) >> 1;
        unsigned long long vp = h.a[parent], vi = h.a[index];
        if (hdist(vp) < hdist(vi)) {
            h.a[parent] = vi;
            h.a[index] = vp;
            index = parent;
        } else
            break;
    }
}
__device__ __forceinline__ void heap_sift_from_root(Heap& h) {  // reference name: up(0)
    int index = 0;
    for (;;) {
        int l = 2 * index + 1, r = l + 1;
        if (l >= h.n) return;
        unsigned long long vi = h.a[index], vl = h.a[l];
        if (r >= h.n) {
            if (hdist(vi) < hdist(vl)) {
                h.a[index] = vl;
                h.a[l] = vi;
            }
            return;
        }
        unsigned long long vr = h.a[r];
        if (hdist(vr) < hdist(vl)) {
            if (hdist(vi) < hdist(vl)) {
                h.a[index] = vl;
                h.a[l] = vi;
                index = l;
            } else
                return;
        } else {
            if (hdist(vi) < hdist(vr)) {
                h.a[index] = vr;
                h.a[r] = vi;
                index = r;
            } else
                return;
        }
    }
}
__device__ __forceinline__ void heap_push(Heap& h, int dist, int idx) {
    if (h.n >= h.cap) {
        if (dist < hdist(h.a[0])) {
            unsigned long long t = h.a[0];
            h.a[0] = h.a[h.n - 1];
            h.a[h.n - 1] = t;
            h.n--;
            if (h.n > 1) heap_sift_from_root(h);
        } else
            return;
    }
    h.a[h.n] = ((unsigned long long)(unsigned)dist << 32) | (unsigned)idx;
    if (h.n > 0) heap_sift_to_root(h, h.n);
    h.n++;
}

__global__ void __launch_bounds__(KNN_WARPS * 32)
hamming_knn_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int k, int order,
                   int32_t* __restrict__ out_idx, int32_t* __restrict__ out_dist) {
    __shared__ __align__(128) uint4 tile[KNN_STAGES][KNN_TILE_ROWS * 2];
    __shared__ __align__(8) uint64_t full[KNN_STAGES];
    __shared__ unsigned long long heaps[KNN_QPB][UCO_KNN_MAX_K];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (nt + KNN_TILE_ROWS - 1) / KNN_TILE_ROWS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KNN_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < KNN_STAGES && s < ntiles; s++) {
            int rows = min(KNN_TILE_ROWS, nt - s * KNN_TILE_ROWS);
            mbar_expect_tx(&full[s], rows * 32);
            tma_bulk_g2s(tile[s], t + (size_t)s * KNN_TILE_ROWS * 2, rows * 32, &full[s]);
        }
    }

    // the two queries of this warp; lanes with (lane>>2)&1 read the second 16-byte half of their train row first
    // (bank-conflict-free LDS.128 on 32-byte rows), so their query halves are swapped to match.
    const int q0 = (blockIdx.x * KNN_WARPS + warp) * KNN_QPW;
    const int swap = (lane >> 2) & 1;
    uint4 qa[KNN_QPW], qb[KNN_QPW];
    Heap h[KNN_QPW];
    int worst[KNN_QPW];
#pragma unroll
    for (int j = 0; j < KNN_QPW; j++) {
        int qi = min(q0 + j, nq - 1);
        uint4 lo = q[(size_t)qi * 2], hi = q[(size_t)qi * 2 + 1];
        qa[j] = swap ? hi : lo;
        qb[j] = swap ? lo : hi;
        h[j].a = heaps[warp * KNN_QPW + j];
        h[j].n = 0;
        h[j].cap = k;
        worst[j] = INT_MAX;  // heap not full: everything enters
    }

    for (int tl = 0; tl < ntiles; tl++) {
        const int s = tl % KNN_STAGES;
        mbar_wait(&full[s], (tl / KNN_STAGES) & 1);
        const int rows = min(KNN_TILE_ROWS, nt - tl * KNN_TILE_ROWS);
        const uint4* tp = tile[s];
        for (int c = 0; c < rows; c += 32) {
            const int r = c + lane;
            const bool valid = r < rows;
            const int rr = valid ? r : 0;
            uint4 ta = tp[rr * 2 + swap], tb = tp[rr * 2 + (swap ^ 1)];
            int d[KNN_QPW];
#pragma unroll
            for (int j = 0; j < KNN_QPW; j++) {
                d[j] = __popc(ta.x ^ qa[j].x) + __popc(ta.y ^ qa[j].y) + __popc(ta.z ^ qa[j].z) + __popc(ta.w ^ qa[j].w) +
                       __popc(tb.x ^ qb[j].x) + __popc(tb.y ^ qb[j].y) + __popc(tb.z ^ qb[j].z) + __popc(tb.w ^ qb[j].w);
            }
            const int base = tl * KNN_TILE_ROWS + c;
#pragma unroll
            for (int j = 0; j < KNN_QPW; j++) {
                unsigned m = __ballot_sync(0xffffffffu, valid && d[j] < worst[j]);
                while (m) {
                    int b = __ffs(m) - 1;
                    m &= m - 1;
                    int db = __shfl_sync(0xffffffffu, d[j], b);
                    if (db < worst[j]) {  // worst only shrinks: re-test against the current value (warp-uniform)
                        if (lane == 0) {
                            heap_push(h[j], db, base + b);
                            // once full, the reference tests  val.dist < distances[0]
                            worst[j] = (h[j].n >= h[j].cap) ? hdist(h[j].a[0]) : INT_MAX;
                        }
                        worst[j] = __shfl_sync(0xffffffffu, worst[j], 0);
                    }
                }
            }
        }
        __syncthreads();  // every warp is done reading tile[s]
        if (threadIdx.x == 0 && tl + KNN_STAGES < ntiles) {
            int nx = tl + KNN_STAGES;
            int nrows = min(KNN_TILE_ROWS, nt - nx * KNN_TILE_ROWS);
            mbar_expect_tx(&full[s], nrows * 32);
            tma_bulk_g2s(tile[s], t + (size_t)nx * KNN_TILE_ROWS * 2, nrows * 32, &full[s]);
        }
    }

    // write back: heap array order, -1 / 0 padding (linear.h:82-85: int32 quiet_NaN() == 0), optional exchange sort
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < KNN_QPW; j++) {
            const int qi = q0 + j;
            if (qi >= nq) continue;
            unsigned long long* a = h[j].a;
            for (int i = h[j].n; i < k; i++) a[i] = 0x00000000ffffffffull;  // dist 0, idx -1
            if (order == UCO_KNN_SORTED) {                                   // index.h:119-133
                for (int i = 0; i < k - 1; i++) {
                    if ((int)(unsigned)a[i] != -1) {
                        for (int jj = i + 1; jj < k; jj++) {
                            if (hdist(a[i]) > hdist(a[jj])) {
                                unsigned long long tmp = a[i];
                                a[i] = a[jj];
                                a[jj] = tmp;
                            }
                        }
                    }
                }
            }
            for (int i = 0; i < k; i++) {
                out_idx[(size_t)qi * k + i] = (int)(unsigned)a[i];
                out_dist[(size_t)qi * k + i] = hdist(a[i]);
            }
        }
    }
}

}  // namespace

extern "C" int uco_b200_hamming_knn_dev(uco_b200_ctx* ctx, const uint8_t* q_dev, int nq, const uint8_t* t_dev, int nt,
                                        int k, int order, int32_t* idx_dev, int32_t* dist_dev) {
    if (!ctx) return UCO_E_INVALID;
    if (nq < 0 || nt < 0 || k <= 0 || k > UCO_KNN_MAX_K || (order != UCO_KNN_HEAP && order != UCO_KNN_SORTED))
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: bad sizes nq=%d nt=%d k=%d order=%d", nq, nt, k, order);
    if (nq == 0) return UCO_OK;
    if (!q_dev || !idx_dev || !dist_dev || (nt > 0 && !t_dev))
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: null pointer");
    if (((uintptr_t)q_dev | (uintptr_t)t_dev) & 15)
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: descriptor buffers must be 16-byte aligned");
    int grid = (nq + KNN_QPB - 1) / KNN_QPB;
    hamming_knn_kernel<<<grid, KNN_WARPS * 32, 0, ctx->stream>>>((const uint4*)q_dev, nq, (const uint4*)t_dev, nt, k,
                                                                order, idx_dev, dist_dev);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

extern "C" int uco_b200_hamming_knn(uco_b200_ctx* ctx, const uint8_t* q, int nq, size_t q_stride, const uint8_t* t,
                                    int nt, size_t t_stride, int k, int order, int32_t* idx, int32_t* dist) {
    if (!ctx) return UCO_E_INVALID;
    if (nq < 0 || nt < 0 || k <= 0 || k > UCO_KNN_MAX_K)
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: bad sizes nq=%d nt=%d k=%d", nq, nt, k);
    if (nq == 0) return UCO_OK;
    if (!q || !idx || !dist || (nt > 0 && !t)) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: null pointer");
    if (q_stride < 32 || (nt > 0 && t_stride < 32))
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: row stride below 32 bytes");
    uint8_t* dq = (uint8_t*)uco_ws(ctx, WS_KNN_Q, (size_t)nq * 32);
    uint8_t* dt = (uint8_t*)uco_ws(ctx, WS_KNN_T, (size_t)nt * 32);
    int32_t* di = (int32_t*)uco_ws(ctx, WS_KNN_IDX, (size_t)nq * k * 4);
    int32_t* dd = (int32_t*)uco_ws(ctx, WS_KNN_DIST, (size_t)nq * k * 4);
    if (!dq || !dt || !di || !dd) return UCO_E_NOMEM;
    // strided host rows (cv::Mat step) are packed by the 2-D copy
    UCO_CUDA(ctx, cudaMemcpy2DAsync(dq, 32, q, q_stride, 32, nq, cudaMemcpyHostToDevice, ctx->stream));
    if (nt > 0) UCO_CUDA(ctx, cudaMemcpy2DAsync(dt, 32, t, t_stride, 32, nt, cudaMemcpyHostToDevice, ctx->stream));
    int rc = uco_b200_hamming_knn_dev(ctx, dq, nq, dt, nt, k, order, di, dd);
    if (rc != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaMemcpyAsync(idx, di, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(dist, dd, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UCO_OK;
}
