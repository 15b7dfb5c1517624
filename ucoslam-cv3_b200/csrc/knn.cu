// knn.cu — K7: brute-force 256-bit Hamming k-NN, exact emulation of xflann's linear index.
//
// Replaces (reference, relative to /root/reference):
//   3rdparty/xflann/xflann/impl/linear.h:68-88      scan of all train rows in index order per query
//   3rdparty/xflann/xflann/impl/resultset.h:64-140  bounded max-heap with strict '<' replacement
//   3rdparty/xflann/xflann/impl/distances.h:279-283 4 x popcount64 distance
//   3rdparty/xflann/xflann/index.h:119-133          optional exchange sort of each result row
//
// Layout: train descriptors are dense 32-byte rows in HBM.  A CTA of KNN_WARPS warps owns KNN_WARPS queries (one per
// warp); the train set streams through a KNN_STAGES-deep ring of shared-memory tiles filled by
// 1-D TMA bulk copies (cp.async.bulk, completion on an mbarrier), so every train byte is read from L2/HBM once per
// CTA with fully coalesced 128-B lines.  Inside a warp each lane owns one train row of the current 32-row
// chunk and computes its distance to the warp's query (held in registers, 8 x __popc).  The result heap of
// the reference is replayed EXACTLY: candidates that can enter the heap (d < current worst) are found with a
// warp ballot and inserted by lane 0 in train-index order, which is the order the reference's scalar loop
// pushes them, so the final array order (the "heap order" the tracker consumes, framematcher.cpp:239-270)
// is identical, including all tie cases.
#include "common.cuh"
#include <climits>
#include <algorithm>
#include <cstring>
#include <cstdlib>

#define KNN_WARPS 8
#define KNN_TILE_ROWS 256         // train rows per shared-memory tile (8 KB)
#define KNN_STAGES 4

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// ---- exact replay of xflann::impl::ResultSet (resultset.h:64-140) -------------------------------------------------
// The heap lives in REGISTERS, replicated in every lane of the warp (all lanes execute the same uniform updates, so
// there is no broadcast and no shared-memory round trip on the serial path).  An entry is packed as
// dist << 23 | idx  (dist <= 256 needs 9 bits, idx < 2^23); the reference compares distances only.
// All indices are compile-time constants (template recursion), so the array never leaves the register file.
#define HD(v) ((v) >> 23)

template <int K, int I, int N>
__device__ __forceinline__ void sift_from_root(uint32_t (&h)[K]) {  // reference name: up(); N = live size
    constexpr int L = 2 * I + 1, R = 2 * I + 2;
    if constexpr (L < N) {
        if constexpr (R >= N) {
            if (HD(h[I]) < HD(h[L])) { uint32_t t = h[I]; h[I] = h[L]; h[L] = t; }
        } else {
            if (HD(h[R]) < HD(h[L])) {
                if (HD(h[I]) < HD(h[L])) { uint32_t t = h[I]; h[I] = h[L]; h[L] = t; sift_from_root<K, L, N>(h); }
            } else {
                if (HD(h[I]) < HD(h[R])) { uint32_t t = h[I]; h[I] = h[R]; h[R] = t; sift_from_root<K, R, N>(h); }
            }
        }
    }
}
template <int K, int I>
__device__ __forceinline__ void sift_to_root(uint32_t (&h)[K]) {  // reference name: down()
    if constexpr (I > 0) {
        constexpr int P = (I - 1) / 2;
        if (HD(h[P]) < HD(h[I])) { uint32_t t = h[I]; h[I] = h[P]; h[P] = t; sift_to_root<K, P>(h); }
    }
}
// push while the set is not full: n in [0, K)
template <int K, int N>
__device__ __forceinline__ void push_fill(uint32_t (&h)[K], int n, uint32_t v) {
    if constexpr (N < K) {
        if (n == N) { h[N] = v; sift_to_root<K, N>(h); }
        else push_fill<K, N + 1>(h, n, v);
    }
}
// push on a full set, caller has checked dist < HD(h[0])
template <int K>
__device__ __forceinline__ void push_full(uint32_t (&h)[K], uint32_t v) {
    uint32_t t = h[0]; h[0] = h[K - 1]; h[K - 1] = t;    // swap(0, size-1); size--
    if constexpr (K - 1 > 1) sift_from_root<K, 0, K - 1>(h);
    h[K - 1] = v;                                          // append at the freed slot
    sift_to_root<K, K - 1>(h);
}

// 256-bit population count of x[0..7] with FEWER POPC instructions (Harley-Seal): bitwise carry-save adders (a full adder is two
// LOP3s) compress words into a ones plane and carry planes of weight two.  POPC is quarter rate (8 cycles per warp instruction
// on its own pipe), LOP3 runs on the ALU pipe (2 cycles): three adders leave 5 POPC + 14 LOP3 per distance — 40 POPC cycles
// against ~34 ALU cycles, the balance point (the full tree, 4 POPC + 22 LOP3, is ALU bound: measured 71 % ALU pipe, 0.559 ms;
// 8 plain POPCs are POPC bound: 0.648 ms).  Exact integer arithmetic either way.
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(sum) : "r"(a), "r"(b), "r"(c));     // a ^ b ^ c
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(carry) : "r"(a), "r"(b), "r"(c));   // majority
}
__device__ __forceinline__ int popc256_xor(const uint4& ta, const uint4& tb, const uint4& qa, const uint4& qb) {
    const uint32_t x0 = ta.x ^ qa.x, x1 = ta.y ^ qa.y, x2 = ta.z ^ qa.z, x3 = ta.w ^ qa.w, x4 = tb.x ^ qb.x, x5 = tb.y ^ qb.y, x6 = tb.z ^ qb.z,
                   x7 = tb.w ^ qb.w;
    uint32_t s1, c1, s2, c2, s3, c3;
    csa(x0, x1, x2, s1, c1);
    csa(x3, x4, x5, s2, c2);
    csa(s1, s2, x6, s3, c3);
    return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

template <int K>
__global__ void __launch_bounds__(KNN_WARPS * 32)
hamming_knn_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int order,
                   int32_t* __restrict__ out_idx, int32_t* __restrict__ out_dist, const int* __restrict__ nq_dev,
                   const int* __restrict__ nt_dev, size_t q_stride16, size_t t_stride16, size_t out_stride,
                   const int* __restrict__ q_sel, const int* __restrict__ t_sel) {
    // blockIdx.y = (query set, train set) pair of a batch; per-pair row counts may live on the device (e.g. the
    // extractor's n_out), so a whole clip is matched without a host round trip
    {
        const int pair = blockIdx.y;
        const int qp = q_sel ? q_sel[pair] : pair, tp_ = t_sel ? t_sel[pair] : pair;   // optional frame selection: pair p = (frame q_sel[p], frame t_sel[p])
        q += qp * q_stride16;
        t += tp_ * t_stride16;
        out_idx += pair * out_stride;
        out_dist += pair * out_stride;
        if (nq_dev) nq = min(nq, nq_dev[qp]);
        if (nt_dev) nt = min(nt, nt_dev[tp_]);
        if ((int)(blockIdx.x * KNN_WARPS) >= nq) return;
    }
    __shared__ __align__(128) uint4 tile[KNN_STAGES][KNN_TILE_ROWS * 2];
    __shared__ __align__(8) uint64_t full[KNN_STAGES];

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

    // this warp's query; lanes with (lane>>2)&1 read the second 16-byte half of their train row first
    // (bank-conflict-free LDS.128 on 32-byte rows), so their query halves are swapped to match.
    const int qi = blockIdx.x * KNN_WARPS + warp;
    const int swap = (lane >> 2) & 1;
    uint4 qa, qb;
    {
        int qq = min(qi, nq - 1);
        uint4 lo = q[(size_t)qq * 2], hi = q[(size_t)qq * 2 + 1];
        qa = swap ? hi : lo;
        qb = swap ? lo : hi;
    }
    uint32_t h[K];
#pragma unroll
    for (int i = 0; i < K; i++) h[i] = 0;
    int n = 0;
    int worst = INT_MAX;  // set not full: everything enters

    for (int tl = 0; tl < ntiles; tl++) {
        const int s = tl % KNN_STAGES;
        mbar_wait(&full[s], (tl / KNN_STAGES) & 1);
        const int rows = min(KNN_TILE_ROWS, nt - tl * KNN_TILE_ROWS);
        // this lane's row of every 32-row chunk: one address add per chunk; lanes with (lane>>2)&1 read the halves swapped
        const uint4* lp = tile[s] + lane * 2;
        const int base_t = tl * KNN_TILE_ROWS;
#define KNN_DIST(c, valid_)                                                                                              \
    ({                                                                                                                   \
        const uint4 ta = lp[(valid_) ? (c) * 2 + swap : swap - lane * 2], tb = lp[(valid_) ? (c) * 2 + (swap ^ 1) : (swap ^ 1) - lane * 2]; \
        popc256_xor(ta, tb, qa, qb);                                                                                     \
    })
        // candidates of one chunk (mask m, distances d) enter the heap in train-index order, exactly the order of the reference's
        // scalar loop; `worst` only shrinks, so every candidate is re-tested against the current value (warp-uniform)
#define KNN_INSERT(m_, d_, c_)                                                                                           \
    {                                                                                                                    \
        unsigned m = (m_);                                                                                               \
        while (m) {                                                                                                      \
            const int b = __ffs(m) - 1;                                                                                  \
            m &= m - 1;                                                                                                  \
            const int db = __shfl_sync(0xffffffffu, (d_), b);                                                            \
            if (db < worst) {                                                                                            \
                const uint32_t v = ((uint32_t)db << 23) | (uint32_t)(base_t + (c_) + b);                                \
                if (n < K) {                                                                                             \
                    push_fill<K, 0>(h, n, v);                                                                            \
                    n++;                                                                                                 \
                } else {                                                                                                 \
                    push_full<K>(h, v);                                                                                  \
                }                                                                                                        \
                if (n >= K) worst = (int)HD(h[0]); /* once full the reference tests val.dist < distances[0] */          \
            }                                                                                                            \
        }                                                                                                                \
    }
        const int rows64 = rows & ~63;
        int c = 0;
        for (; c < rows64; c += 64) {   // two chunks per round: two independent popc chains, one test for the common "nothing enters" case
            const int d0 = KNN_DIST(c, true), d1 = KNN_DIST(c + 32, true);
            const unsigned m0 = __ballot_sync(0xffffffffu, d0 < worst), m1 = __ballot_sync(0xffffffffu, d1 < worst);
            if ((m0 | m1) == 0) continue;
            KNN_INSERT(m0, d0, c)
            KNN_INSERT(m1 & __ballot_sync(0xffffffffu, d1 < worst), d1, c + 32)
        }
        for (; c < rows; c += 32) {   // ragged tail of the last tile
            const bool valid = c + lane < rows;
            const int d = KNN_DIST(c, valid);
            const unsigned m0 = __ballot_sync(0xffffffffu, valid && d < worst);
            KNN_INSERT(m0, d, c)
        }
#undef KNN_DIST
#undef KNN_INSERT
        __syncthreads();  // every warp is done reading tile[s]
        if (threadIdx.x == 0 && tl + KNN_STAGES < ntiles) {
            int nx = tl + KNN_STAGES;
            int nrows = min(KNN_TILE_ROWS, nt - nx * KNN_TILE_ROWS);
            mbar_expect_tx(&full[s], nrows * 32);
            tma_bulk_g2s(tile[s], t + (size_t)nx * KNN_TILE_ROWS * 2, nrows * 32, &full[s]);
        }
    }

    // write back: heap array order, -1 / 0 padding (linear.h:82-85: int32 quiet_NaN() == 0), optional exchange sort
    if (qi >= nq) return;
    int od[K], oi[K];
#pragma unroll
    for (int i = 0; i < K; i++) {
        od[i] = i < n ? (int)HD(h[i]) : 0;
        oi[i] = i < n ? (int)(h[i] & 0x7fffffu) : -1;
    }
    if (order == UCO_KNN_SORTED) {  // index.h:119-133
#pragma unroll
        for (int i = 0; i < K - 1; i++) {
#pragma unroll
            for (int j = i + 1; j < K; j++) {
                if (oi[i] != -1 && od[i] > od[j]) {
                    int tmp = od[i]; od[i] = od[j]; od[j] = tmp;
                    tmp = oi[i]; oi[i] = oi[j]; oi[j] = tmp;
                }
            }
        }
    }
    // every lane holds the row; lanes 0..K-1 each store one element (coalesced)
    int vd = 0, vi = 0;
#pragma unroll
    for (int i = 0; i < K; i++)
        if (lane == i) { vd = od[i]; vi = oi[i]; }
    if (lane < K) {
        out_idx[(size_t)qi * K + lane] = vi;
        out_dist[(size_t)qi * K + lane] = vd;
    }
}


// ---- lane-per-query form (the default when a launch holds enough queries to fill the machine) ------------------------------------
// Profile of the kernel above at the tracking shape (2000 x 2000): of 113 issued instructions per 32 distances only 47 score;
// the rest replays the reference's heap, once per accepted row (~53 per query) and redundantly in all 32 lanes of the query's
// warp, which makes the kernel issue-bound at ~52 % of the popc roof.  The replay is inherent — the final ARRAY ORDER is a
// function of every accepted push — but 32 heaps can advance per instruction instead of one:
//   * a LANE owns a query (8 registers) and its heap (K registers); a warp scans the train rows of the shared-memory tile one
//     row at a time (two broadcast LDS.128), every lane scoring the row against its own query: ~27 instructions per 32
//     distances, below the 64 cycles the popc pipe needs for them;
//   * rows with d < bound (the lane's heap root at the start of the current 32-row window: stale values are only LARGER, so this
//     is a superset of what the reference accepts) are appended, in train order, to a lane-private ring in shared memory
//     (ring[pos][thread]: the bank is the lane, no conflicts);
//   * at the end of every window the warp replays the rings in lockstep — round r handles the r-th entry of every lane with the
//     exact test d < worst — so one instruction stream serves up to 32 pushes; late windows need 0-2 rounds.
// Identical output to the kernel above (and to xflann's linear index) by construction: every query sees its candidate rows in
// train-index order and every row the reference would accept is in its ring.
#define KLQ_WARPS 4
#define KLQ_WINDOW 32
#define KLQ_TILE_ROWS 128         // 4 KB tiles x 3 stages + the 16 KB ring = 28 KB per CTA: 7 CTAs per SM, so a clip's launch
#define KLQ_STAGES 3              // (16 CTAs per 2000-query frame pair) is resident in one wave

template <int K>
__global__ void __launch_bounds__((KLQ_WARPS + 1) * 32)
hamming_knn_lq_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int order,
                      int32_t* __restrict__ out_idx, int32_t* __restrict__ out_dist, const int* __restrict__ nq_dev,
                      const int* __restrict__ nt_dev, size_t q_stride16, size_t t_stride16, size_t out_stride,
                   const int* __restrict__ q_sel, const int* __restrict__ t_sel) {
    {
        const int pair = blockIdx.y;
        const int qp = q_sel ? q_sel[pair] : pair, tp_ = t_sel ? t_sel[pair] : pair;   // optional frame selection: pair p = (frame q_sel[p], frame t_sel[p])
        q += qp * q_stride16;
        t += tp_ * t_stride16;
        out_idx += pair * out_stride;
        out_dist += pair * out_stride;
        if (nq_dev) nq = min(nq, nq_dev[qp]);
        if (nt_dev) nt = min(nt, nt_dev[tp_]);
        if ((int)(blockIdx.x * KLQ_WARPS * 32) >= nq) return;
    }
    __shared__ __align__(128) uint4 tile[KLQ_STAGES][KLQ_TILE_ROWS * 2];
    __shared__ __align__(8) uint64_t full[KLQ_STAGES], empty[KLQ_STAGES];
    __shared__ uint32_t ring[KLQ_WINDOW][KLQ_WARPS * 32];

    const int ntiles = (nt + KLQ_TILE_ROWS - 1) / KLQ_TILE_ROWS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < KLQ_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], KLQ_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= KLQ_WARPS * 32) {
        // producer warp: one thread streams the train tiles through the ring; a stage is refilled as soon as all scanning warps
        // have released it (empty[s]), so the scanning warps never wait for each other, only for data
        if (threadIdx.x == KLQ_WARPS * 32) {
            for (int tl = 0; tl < ntiles; tl++) {
                const int s = tl % KLQ_STAGES;
                if (tl >= KLQ_STAGES) mbar_wait(&empty[s], ((tl / KLQ_STAGES) - 1) & 1);
                const int rows = min(KLQ_TILE_ROWS, nt - tl * KLQ_TILE_ROWS);
                mbar_expect_tx(&full[s], rows * 32);
                tma_bulk_g2s(tile[s], t + (size_t)tl * KLQ_TILE_ROWS * 2, rows * 32, &full[s]);
            }
        }
        return;
    }
    const int qi = blockIdx.x * KLQ_WARPS * 32 + threadIdx.x;
    uint4 qa, qb;
    {
        const int qq = min(qi, nq - 1);
        qa = q[(size_t)qq * 2];
        qb = q[(size_t)qq * 2 + 1];
    }
    uint32_t h[K];
#pragma unroll
    for (int i = 0; i < K; i++) h[i] = 0;
    int n = 0, worst = INT_MAX;
    uint32_t* my_ring = &ring[0][threadIdx.x];

    for (int tl = 0; tl < ntiles; tl++) {
        const int s = tl % KLQ_STAGES;
        mbar_wait(&full[s], (tl / KLQ_STAGES) & 1);
        const int rows = min(KLQ_TILE_ROWS, nt - tl * KLQ_TILE_ROWS);
        const int base_t = tl * KLQ_TILE_ROWS;
        for (int w0 = 0; w0 < rows; w0 += KLQ_WINDOW) {
            const int wr = min(KLQ_WINDOW, rows - w0);
            const uint4* tp = tile[s] + w0 * 2;
            const int bound = worst;
            int cnt = 0;
#pragma unroll 8
            for (int r = 0; r < wr; r++) {
                const uint4 ta = tp[2 * r], tb = tp[2 * r + 1];   // the whole warp reads the same row: broadcast
                const int d = popc256_xor(ta, tb, qa, qb);
                // branch-free append: the slot is always written, the counter only advances for a candidate (cnt <= r < 32)
                my_ring[cnt * (KLQ_WARPS * 32)] = ((uint32_t)d << 23) | (uint32_t)(base_t + w0 + r);
                cnt += d < bound ? 1 : 0;
            }
            const int rounds = __reduce_max_sync(0xffffffffu, cnt);
            for (int r = 0; r < rounds; r++) {
                if (r < cnt) {
                    const uint32_t v = my_ring[r * (KLQ_WARPS * 32)];
                    if ((int)HD(v) < worst) {
                        if (n < K) {
                            push_fill<K, 0>(h, n, v);
                            n++;
                        } else {
                            push_full<K>(h, v);
                        }
                        if (n >= K) worst = (int)HD(h[0]);
                    }
                }
            }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);   // this warp is done reading tile[s]
    }
    if (qi >= nq) return;
    int od[K], oi[K];
#pragma unroll
    for (int i = 0; i < K; i++) {
        od[i] = i < n ? (int)HD(h[i]) : 0;
        oi[i] = i < n ? (int)(h[i] & 0x7fffffu) : -1;
    }
    if (order == UCO_KNN_SORTED) {  // index.h:119-133
#pragma unroll
        for (int i = 0; i < K - 1; i++) {
#pragma unroll
            for (int j = i + 1; j < K; j++) {
                if (oi[i] != -1 && od[i] > od[j]) {
                    int tmp = od[i]; od[i] = od[j]; od[j] = tmp;
                    tmp = oi[i]; oi[i] = oi[j]; oi[j] = tmp;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < K; i++) {
        out_idx[(size_t)qi * K + i] = oi[i];
        out_dist[(size_t)qi * K + i] = od[i];
    }
}

typedef void (*knn_fn)(const uint4*, int, const uint4*, int, int, int32_t*, int32_t*, const int*, const int*, size_t, size_t,
                       size_t, const int*, const int*);
template <int K>
struct KnnTable {
    static void fill(knn_fn* f, knn_fn* w) {
        f[K] = hamming_knn_kernel<K>;
        if constexpr (K <= 16) w[K] = hamming_knn_lq_kernel<K>;
        else w[K] = nullptr;
        KnnTable<K - 1>::fill(f, w);
    }
};
template <>
struct KnnTable<0> {
    static void fill(knn_fn*, knn_fn*) {}
};
struct KnnTables {
    knn_fn warp_per_query[UCO_KNN_MAX_K + 1], specialised[UCO_KNN_MAX_K + 1];
    KnnTables() { KnnTable<UCO_KNN_MAX_K>::fill(warp_per_query, specialised); }
};

}  // namespace

static int knn_launch(uco_b200_ctx* ctx, const uint8_t* q_dev, int nq, const uint8_t* t_dev, int nt, int k, int order,
                      int32_t* idx_dev, int32_t* dist_dev, int n_pairs, const int* nq_dev, const int* nt_dev, size_t q_stride,
                      size_t t_stride, const int* q_sel = nullptr, const int* t_sel = nullptr) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (nq < 0 || nt < 0 || n_pairs < 0 || k <= 0 || k > UCO_KNN_MAX_K || (order != UCO_KNN_HEAP && order != UCO_KNN_SORTED))
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: bad sizes nq=%d nt=%d k=%d order=%d", nq, nt, k, order);
    if (nq == 0 || n_pairs == 0) return UCO_OK;
    if (!q_dev || !idx_dev || !dist_dev || (nt > 0 && !t_dev))
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: null pointer");
    if ((((uintptr_t)q_dev | (uintptr_t)t_dev) & 15) || (q_stride & 15) || (t_stride & 15))
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: descriptor buffers must be 16-byte aligned");
    if (nt >= (1 << 23))
        return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: train set above %d rows, shard it", (1 << 23) - 1);
    if (n_pairs > 65535) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn: more than 65535 pairs in one call");
    static const KnnTables tables;   // thread-safe initialisation (C++11 magic static): contexts on several threads may race here
    static const bool force_old = getenv("UCO_KNN_WARP_PER_QUERY") != nullptr;   // A/B switch for measurements
    // lane-per-query needs 32 queries per warp: only when the launch still fills the machine that way (a clip's frame pairs);
    // a single 2000-query scan of a large map keeps the warp-per-query form (8x more warps)
    if ((size_t)n_pairs * ((nq + 31) / 32) >= (size_t)ctx->sm_count * 8 && tables.specialised[k] && !force_old) {
        dim3 grid((nq + KLQ_WARPS * 32 - 1) / (KLQ_WARPS * 32), n_pairs);
        tables.specialised[k]<<<grid, (KLQ_WARPS + 1) * 32, 0, ctx->stream>>>((const uint4*)q_dev, nq, (const uint4*)t_dev, nt, order, idx_dev, dist_dev,
                                                                      nq_dev, nt_dev, q_stride / 16, t_stride / 16, (size_t)nq * k, q_sel, t_sel);
    } else {
        dim3 grid((nq + KNN_WARPS - 1) / KNN_WARPS, n_pairs);
        tables.warp_per_query[k]<<<grid, KNN_WARPS * 32, 0, ctx->stream>>>((const uint4*)q_dev, nq, (const uint4*)t_dev, nt, order, idx_dev,
                                                                            dist_dev, nq_dev, nt_dev, q_stride / 16, t_stride / 16, (size_t)nq * k, q_sel, t_sel);
    }
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

// used by match.cu (K8 = K7 + post-filters)
int uco_knn_launch_internal(uco_b200_ctx* ctx, const uint8_t* q_dev, int nq, const uint8_t* t_dev, int nt, int k, int order,
                            int32_t* idx_dev, int32_t* dist_dev, int n_pairs, const int* nq_dev, const int* nt_dev,
                            size_t q_stride, size_t t_stride) {
    return knn_launch(ctx, q_dev, nq, t_dev, nt, k, order, idx_dev, dist_dev, n_pairs, nq_dev, nt_dev, q_stride, t_stride);
}
// pair p = (query frame q_sel[p], train frame t_sel[p]) of a frame-strided buffer; nq_dev / nt_dev are indexed by FRAME
int uco_knn_launch_selected(uco_b200_ctx* ctx, const uint8_t* q_dev, int nq, const uint8_t* t_dev, int nt, int k, int order,
                            int32_t* idx_dev, int32_t* dist_dev, int n_pairs, const int* nq_dev, const int* nt_dev,
                            size_t q_stride, size_t t_stride, const int* q_sel, const int* t_sel) {
    return knn_launch(ctx, q_dev, nq, t_dev, nt, k, order, idx_dev, dist_dev, n_pairs, nq_dev, nt_dev, q_stride, t_stride, q_sel, t_sel);
}

extern "C" int uco_b200_hamming_knn_dev(uco_b200_ctx* ctx, const uint8_t* q_dev, int nq, const uint8_t* t_dev, int nt,
                                        int k, int order, int32_t* idx_dev, int32_t* dist_dev) {
    return knn_launch(ctx, q_dev, nq, t_dev, nt, k, order, idx_dev, dist_dev, 1, nullptr, nullptr, 0, 0);
}

extern "C" int uco_b200_hamming_knn_batch_dev(uco_b200_ctx* ctx, int n_pairs, const uint8_t* q_dev, size_t q_pair_stride,
                                              int nq_max, const int32_t* nq_dev, const uint8_t* t_dev,
                                              size_t t_pair_stride, int nt_max, const int32_t* nt_dev, int k, int order,
                                              int32_t* idx_dev, int32_t* dist_dev) {
    return knn_launch(ctx, q_dev, nq_max, t_dev, nt_max, k, order, idx_dev, dist_dev, n_pairs, nq_dev, nt_dev,
                      q_pair_stride, t_pair_stride);
}

extern "C" int uco_b200_hamming_knn(uco_b200_ctx* ctx, const uint8_t* q, int nq, size_t q_stride, const uint8_t* t,
                                    int nt, size_t t_stride, int k, int order, int32_t* idx, int32_t* dist) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
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

// Host-buffer batch: n_pairs (query set, train set) pairs in ONE launch and one synchronisation (the per-call fixed cost of
// uco_b200_hamming_knn, ~60 us of copies + launch + sync, is what a 64-frame clip pays 64 times otherwise).  Copies are
// coalesced: a run of pairs whose host buffers are contiguous in the device layout is one transfer, and the chain pattern of
// tracking (the train set of pair i is the query set of pair i-1) is detected so every descriptor block is uploaded once.
extern "C" int uco_b200_hamming_knn_batch(uco_b200_ctx* ctx, int n_pairs, const uint8_t* const* q, const int32_t* nq,
                                          size_t q_stride, const uint8_t* const* t, const int32_t* nt, size_t t_stride, int k,
                                          int order, int32_t* const* idx, int32_t* const* dist) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (n_pairs < 0 || k <= 0 || k > UCO_KNN_MAX_K) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_batch: bad sizes n_pairs=%d k=%d", n_pairs, k);
    if (n_pairs == 0) return UCO_OK;
    if (!q || !nq || !t || !nt || !idx || !dist) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_batch: null pointer");
    if (q_stride < 32 || t_stride < 32) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_batch: row stride below 32 bytes");
    int nq_max = 0, nt_max = 0;
    for (int i = 0; i < n_pairs; i++) {
        if (nq[i] < 0 || nt[i] < 0) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_batch: negative row count in pair %d", i);
        if ((nq[i] > 0 && (!q[i] || !idx[i] || !dist[i])) || (nt[i] > 0 && !t[i]))
            return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_batch: null buffer in pair %d", i);
        nq_max = std::max(nq_max, nq[i]);
        nt_max = std::max(nt_max, nt[i]);
    }
    if (nq_max == 0) return UCO_OK;
    bool chain = q_stride == t_stride;   // train set of pair i == query set of pair i-1
    for (int i = 1; i < n_pairs && chain; i++) chain = t[i] == q[i - 1] && nt[i] == nq[i - 1];
    const int slot_rows = chain ? std::max(nq_max, nt_max) : 0;
    const size_t qs = (size_t)(chain ? slot_rows : nq_max) * 32, ts = (size_t)(chain ? slot_rows : std::max(nt_max, 1)) * 32;
    uint8_t *dq, *dt;
    if (chain) {
        uint8_t* d = (uint8_t*)uco_ws(ctx, WS_KNN_Q, qs * (size_t)(n_pairs + 1));
        if (!d) return UCO_E_NOMEM;
        dt = d;
        dq = d + qs;
    } else {
        dq = (uint8_t*)uco_ws(ctx, WS_KNN_Q, qs * (size_t)n_pairs);
        dt = (uint8_t*)uco_ws(ctx, WS_KNN_T, ts * (size_t)n_pairs);
    }
    const size_t os = (size_t)nq_max * k;  // output entries per pair
    int32_t* di = (int32_t*)uco_ws(ctx, WS_KNN_IDX, os * 4 * (size_t)n_pairs);
    int32_t* dd = (int32_t*)uco_ws(ctx, WS_KNN_DIST, os * 4 * (size_t)n_pairs);
    int32_t* dn = (int32_t*)uco_ws(ctx, WS_KNN_N, 8 * (size_t)n_pairs);
    int32_t* hn = (int32_t*)uco_pinned(ctx, WS_KNN_N, 8 * (size_t)n_pairs);
    if (!dq || !dt || !di || !dd || !dn || !hn) return UCO_E_NOMEM;
    cudaStream_t s = ctx->stream;
    memcpy(hn, nq, 4 * (size_t)n_pairs);
    memcpy(hn + n_pairs, nt, 4 * (size_t)n_pairs);
    UCO_CUDA(ctx, cudaMemcpyAsync(dn, hn, 8 * (size_t)n_pairs, cudaMemcpyHostToDevice, s));
    // uploads: runs of host blocks that are already laid out like the device slots become one copy
    auto upload = [&](uint8_t* dbase, size_t slot, const uint8_t* const* hp, const int32_t* rows, size_t stride, int n) -> int {
        for (int i = 0; i < n;) {
            if (rows[i] == 0) { i++; continue; }
            int j = i + 1;
            if (stride == 32)
                while (j < n && hp[j] == hp[j - 1] + slot && (size_t)rows[j - 1] * 32 == slot) j++;
            if (j - i > 1 || stride == 32) {
                const size_t bytes = slot * (size_t)(j - 1 - i) + (size_t)rows[j - 1] * 32;
                UCO_CUDA(ctx, cudaMemcpyAsync(dbase + slot * i, hp[i], bytes, cudaMemcpyHostToDevice, s));
            } else {
                UCO_CUDA(ctx, cudaMemcpy2DAsync(dbase + slot * i, 32, hp[i], stride, 32, rows[i], cudaMemcpyHostToDevice, s));
            }
            i = j;
        }
        return UCO_OK;
    };
    int rc;
    if (chain) {
        if ((rc = upload(dt, qs, t, nt, t_stride, 1)) != UCO_OK) return rc;
        if ((rc = upload(dq, qs, q, nq, q_stride, n_pairs)) != UCO_OK) return rc;
    } else {
        if ((rc = upload(dq, qs, q, nq, q_stride, n_pairs)) != UCO_OK) return rc;
        if ((rc = upload(dt, ts, t, nt, t_stride, n_pairs)) != UCO_OK) return rc;
    }
    rc = knn_launch(ctx, dq, nq_max, dt, nt_max, k, order, di, dd, n_pairs, dn, dn + n_pairs, qs, ts);
    if (rc != UCO_OK) return rc;
    auto download = [&](const int32_t* dbase, int32_t* const* hp) -> int {
        for (int i = 0; i < n_pairs;) {
            if (nq[i] == 0) { i++; continue; }
            int j = i + 1;
            while (j < n_pairs && nq[j - 1] == nq_max && nq[j] > 0 && hp[j] == hp[j - 1] + os) j++;
            const size_t n_int = os * (size_t)(j - 1 - i) + (size_t)nq[j - 1] * k;
            UCO_CUDA(ctx, cudaMemcpyAsync(hp[i], dbase + os * i, n_int * 4, cudaMemcpyDeviceToHost, s));
            i = j;
        }
        return UCO_OK;
    };
    if ((rc = download(di, idx)) != UCO_OK) return rc;
    if ((rc = download(dd, dist)) != UCO_OK) return rc;
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    return UCO_OK;
}

// ---- row-sharded train set (BASELINE config 4: a 10^6-descriptor map split over the GPUs of a node) ------------------------------
// Every rank scans its own rows (same kernel), the per-shard top-k lists travel once (all-gather of nq x k 64-bit keys per rank) and
// every rank merges them.  A key is  distance << 32 | global row index ; the merged rows come out in (distance, index) order with
// xflann's distance list and xflann's rows below the k-th distance.  Which of the rows TIED at the k-th distance survive in xflann
// depends on its heap layout when closer rows arrive later (resultset.h:64-85), i.e. on the scan order: per-shard lists cannot
// reproduce that, so the merge keeps the lowest row indices among those ties (SURVEY.md 8(c)(ii) defines parity on sorted lists).
namespace {
// list l (per_list entries each) holds row indices relative to base + l * list_rows
__global__ void knn_pack_keys_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ dist, int n, int base, int per_list,
                                     int list_rows, unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = idx[i];
    keys[i] = id < 0 ? ~0ull : ((unsigned long long)(unsigned)dist[i] << 32) | (unsigned)(id + base + (i / per_list) * list_rows);
}
// one warp per query: n_lists x k candidate keys (list l of query q at keys[(l * nq + q) * k]), k rounds of warp-wide minimum
__global__ void __launch_bounds__(256) knn_merge_kernel(const unsigned long long* __restrict__ keys, int n_lists, int nq, int k,
                                                        int32_t* __restrict__ out_idx, int32_t* __restrict__ out_dist) {
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const int total = n_lists * k;   // <= 1024
    unsigned long long mine[32];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const int e = j * 32 + lane;
        mine[j] = e < total ? keys[((size_t)(e / k) * nq + q) * k + e % k] : ~0ull;
    }
    for (int r = 0; r < k; r++) {
        unsigned long long m = ~0ull;
#pragma unroll
        for (int j = 0; j < 32; j++) m = mine[j] < m ? mine[j] : m;
        unsigned long long w = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, w, o);
            w = t < w ? t : w;
        }
        if (w != ~0ull && m == w) {  // keys are unique (distinct rows): exactly one lane owns the winner
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (mine[j] == w) mine[j] = ~0ull;
        }
        if (lane == 0) {
            out_idx[(size_t)q * k + r] = w == ~0ull ? -1 : (int32_t)(unsigned)w;
            out_dist[(size_t)q * k + r] = w == ~0ull ? 0 : (int32_t)(w >> 32);
        }
    }
}
}  // namespace

extern "C" int uco_b200_knn_merge_dev(uco_b200_ctx* ctx, int n_lists, int nq, int k, const int32_t* idx_lists_dev,
                                      const int32_t* dist_lists_dev, int32_t* idx_dev, int32_t* dist_dev) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (n_lists <= 0 || n_lists * k > 1024 || nq < 0 || k <= 0 || k > UCO_KNN_MAX_K || !idx_lists_dev || !dist_lists_dev || !idx_dev || !dist_dev)
        return uco_fail(ctx, UCO_E_INVALID, "knn_merge: bad arguments");
    if (nq == 0) return UCO_OK;
    const size_t n = (size_t)n_lists * nq * k;
    unsigned long long* keys = (unsigned long long*)uco_ws(ctx, WS_KNN_MERGE, 8 * n);
    if (!keys) return UCO_E_NOMEM;
    knn_pack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(idx_lists_dev, dist_lists_dev, (int)n, 0, (int)n, 0, keys);
    UCO_LAUNCH_CHECK(ctx);
    knn_merge_kernel<<<(nq + 7) / 8, 256, 0, ctx->stream>>>(keys, n_lists, nq, k, idx_dev, dist_dev);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

extern "C" int uco_b200_hamming_knn_sharded_dev(uco_b200_ctx* ctx, uco_b200_comm* comm, const uint8_t* q_dev, int nq,
                                                const uint8_t* t_shard_dev, int nt_shard, int row_base, int k, int32_t* idx_dev,
                                                int32_t* dist_dev) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    const int R = uco_comm_world(comm), rank = uco_comm_rank(comm);
    if (R > 32) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_sharded: more than 32 ranks");
    // every argument is checked before anything is sized from it or any rank enters a collective (an early return on one rank
    // only — e.g. out of memory below — is NOT collective: the caller has to tear the communicator down, see include/ucoslam_b200.h)
    if (k <= 0 || k > UCO_KNN_MAX_K || (size_t)R * k > 1024) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_sharded: k=%d with %d ranks (need 1 <= k <= %d, ranks * k <= 1024)", k, R, UCO_KNN_MAX_K);
    if (nt_shard < 0 || row_base < 0) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_sharded: negative shard size / row base");
    if (nq > 0 && (!q_dev || !idx_dev || !dist_dev || (nt_shard > 0 && !t_shard_dev))) return uco_fail(ctx, UCO_E_INVALID, "hamming_knn_sharded: null pointer");
    if (nq <= 0) return nq == 0 ? UCO_OK : uco_fail(ctx, UCO_E_INVALID, "hamming_knn_sharded: nq < 0");
    const size_t n = (size_t)nq * k;
    // few queries against many rows: one CTA per 8 queries would leave the GPU idle, so the shard is itself scanned in S row ranges
    // (grid.y) and the S lists are merged first — the same step as across ranks
    int S = 1;
    const int qcta = (nq + KNN_WARPS - 1) / KNN_WARPS;
    if (qcta < 2 * ctx->sm_count && nt_shard >= 2 * 8192) S = std::min(std::min(std::min(128, 1024 / k), nt_shard / 8192), (2 * ctx->sm_count + qcta - 1) / qcta);
    const int chunk = S > 1 ? (((nt_shard + S - 1) / S + 255) & ~255) : nt_shard;
    if (S > 1) S = (nt_shard + chunk - 1) / chunk;
    int32_t* li = (int32_t*)uco_ws(ctx, WS_KNN_IDX, 4 * n * (size_t)S);
    int32_t* ld = (int32_t*)uco_ws(ctx, WS_KNN_DIST, 4 * n * (size_t)S);
    unsigned long long* keys = (unsigned long long*)uco_ws(ctx, WS_KNN_MERGE, 8 * n * (size_t)(R + S));
    if (!li || !ld || !keys) return UCO_E_NOMEM;
    int rc;
    if (S == 1) {
        rc = knn_launch(ctx, q_dev, nq, t_shard_dev, nt_shard, k, UCO_KNN_SORTED, li, ld, 1, nullptr, nullptr, 0, 0);
        if (rc != UCO_OK) return rc;
        knn_pack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(li, ld, (int)n, row_base, (int)n, 0, keys + n * rank);
        UCO_LAUNCH_CHECK(ctx);
    } else {
        int32_t* dn = (int32_t*)uco_ws(ctx, WS_KNN_N, 4 * 128);
        int32_t* hn = (int32_t*)uco_pinned(ctx, WS_KNN_N, 4 * 128);
        if (!dn || !hn) return UCO_E_NOMEM;
        UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // hn may still be in flight from the previous call
        for (int i = 0; i < S; i++) hn[i] = std::min(chunk, nt_shard - i * chunk);
        UCO_CUDA(ctx, cudaMemcpyAsync(dn, hn, 4 * (size_t)S, cudaMemcpyHostToDevice, ctx->stream));
        rc = knn_launch(ctx, q_dev, nq, t_shard_dev, chunk, k, UCO_KNN_SORTED, li, ld, S, nullptr, dn, 0, (size_t)chunk * 32);
        if (rc != UCO_OK) return rc;
        unsigned long long* part = keys + n * (size_t)R;
        knn_pack_keys_kernel<<<(unsigned)((n * S + 255) / 256), 256, 0, ctx->stream>>>(li, ld, (int)(n * S), row_base, (int)n, chunk, part);
        UCO_LAUNCH_CHECK(ctx);
        // merged shard list -> (idx, dist) -> this rank's slot of the exchange buffer
        knn_merge_kernel<<<(nq + 7) / 8, 256, 0, ctx->stream>>>(part, S, nq, k, li, ld);
        UCO_LAUNCH_CHECK(ctx);
        knn_pack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(li, ld, (int)n, 0, (int)n, 0, keys + n * rank);
        UCO_LAUNCH_CHECK(ctx);
    }
    if ((rc = uco_comm_allgather(comm, keys + n * rank, keys, 8 * n, ctx->stream)) != UCO_OK) return rc;
    knn_merge_kernel<<<(nq + 7) / 8, 256, 0, ctx->stream>>>(keys, R, nq, k, idx_dev, dist_dev);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}
