// kfdb.cu — K11: keyframe database query (relocalisation / loop-closure candidates) over device-resident bags of words.
//
// Replaces (reference, relative to /root/reference):
//   src/map_types/keyframedatabase.cpp:150-171  KPFrameDataBase::add / del (inverted index word -> set of frames)
//   src/map_types/keyframedatabase.cpp:195-233  relocalizationCandidates steps 1-2: for every database frame the number of
//                                               query words it contains, maxCommonWords, minCommonWords = max*0.8f, and
//                                               fbow::fBow::score for the frames above it, kept if > minScore
//   src/map_types/keyframedatabase.cpp:236-275  steps 3-4 (covisibility accumulation, 0.75*best, sort): host, uco_b200_kfdb_rank
//   3rdparty/fbow/fbow/fbow.cpp:192-243         fBow::score: sum of v_i*w_i (float product, double sum, ascending word id),
//                                               1 - sqrt(1 - s), clamped at s >= 1
//
// The reference walks an inverted index (std::map<word, std::set<frame>>) and counts votes into a std::map<frame, nobs>:
// pointer chasing whose cost grows with the database.  Here the database is the FORWARD index: every keyframe's ascending word
// list lives in one HBM arena (u32 words, 16-byte aligned segments padded with 0xFFFFFFFF; f32 weights in a parallel arena), and a
// query is one streaming pass over the word arena: the query's words are a bitmap (max word / 8 bytes: 128 KB for the shipped 10^6
// word vocabulary, L1/L2 resident), one warp per keyframe tests 4 words per lane per 16-byte load and the popcount of hits IS the
// reference's vote count.  Algorithmic bytes: 4 B per stored word (weights are only read for the few frames that pass the
// 0.8*max gate), so the scan is HBM-bound: 36 M words (20 k keyframes) = 144 MB = 22 us at the measured copy bandwidth.
// The scores of the surviving frames are summed by one warp per frame in ascending word order (ballot + ordered shuffle), so the
// doubles are bit-identical to fBow::score.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <utility>
#include <vector>

struct uco_b200_kfdb {
    int device = 0;
    // device arenas (entries; every segment starts at a multiple of 4 entries)
    uint32_t* d_words = nullptr;
    float* d_weights = nullptr;
    uint64_t cap = 0, used = 0, dead = 0;
    // per slot: (segment start / 4, length); length 0 = deleted
    uint2* d_slots = nullptr;
    uint32_t* d_frame = nullptr;
    uint32_t* d_nobs = nullptr;
    uint8_t* d_excl = nullptr;
    uint32_t slot_cap = 0;
    std::vector<uint2> slots;
    std::vector<uint32_t> frame;
    std::vector<uint8_t> alive;   // a live frame may have no words (length 0) too
    std::unordered_map<uint32_t, uint32_t> slot_of;
    uint32_t live = 0;
    float last_ms[2] = {0, 0};
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};

namespace {

struct KfHit {
    uint32_t frame, common;
    double score;
};

__device__ __forceinline__ unsigned kf_test(uint32_t w, const uint32_t* __restrict__ bitmap, uint32_t max_word) {
    return (w <= max_word) ? ((__ldg(bitmap + (w >> 5)) >> (w & 31)) & 1u) : 0u;
}

__global__ void kfdb_mark_kernel(const uint32_t* __restrict__ qwords, int nq, uint32_t* bitmap, const uint32_t* __restrict__ excl_slots,
                                 int n_excl, uint8_t* excl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) atomicOr(bitmap + (qwords[i] >> 5), 1u << (qwords[i] & 31));
    if (i < n_excl) excl[excl_slots[i]] = 1;
}

// step 1 (keyframedatabase.cpp:206-217): votes per database frame + the maximum; one warp per frame, grid-stride
__global__ void __launch_bounds__(256) kfdb_count_kernel(const uint4* __restrict__ words4, const uint2* __restrict__ slots,
                                                         const uint8_t* __restrict__ excl, int n_slots,
                                                         const uint32_t* __restrict__ bitmap, uint32_t max_word,
                                                         uint32_t* __restrict__ nobs, uint32_t* __restrict__ d_max) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t local_max = 0;
    for (int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < n_slots; slot += nwarps) {
        const uint2 s = slots[slot];
        uint32_t cnt = 0;
        if (s.y != 0 && !excl[slot]) {
            const uint4* p = words4 + s.x;
            const int n4 = (int)((s.y + 3) >> 2);
            int i = lane;
            for (; i + 32 < n4; i += 64) {   // two 16-byte loads in flight per lane
                const uint4 a = __ldcs(p + i), b = __ldcs(p + i + 32);
                cnt += kf_test(a.x, bitmap, max_word) + kf_test(a.y, bitmap, max_word) + kf_test(a.z, bitmap, max_word) +
                       kf_test(a.w, bitmap, max_word);
                cnt += kf_test(b.x, bitmap, max_word) + kf_test(b.y, bitmap, max_word) + kf_test(b.z, bitmap, max_word) +
                       kf_test(b.w, bitmap, max_word);
            }
            if (i < n4) {
                const uint4 a = __ldcs(p + i);
                cnt += kf_test(a.x, bitmap, max_word) + kf_test(a.y, bitmap, max_word) + kf_test(a.z, bitmap, max_word) +
                       kf_test(a.w, bitmap, max_word);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) nobs[slot] = cnt;
        local_max = max(local_max, cnt);
    }
    if (lane == 0 && local_max) atomicMax(d_max, local_max);
}

// the same scan with the query bitmap staged in shared memory (vocabularies up to KFDB_SMEM_BITMAP_BYTES * 8 words: the shipped
// 10^6-word vocabulary needs 122 KB): one persistent CTA of 16 warps per SM, warps take keyframes from a global ticket counter.
// Through L1 every word test is its own 32-byte sector request (ncu on the kernel above: 37.7 M sector requests per query against
// 4.7 M for the word stream itself, warps waiting in lg_throttle); in shared memory it is a bank-conflicted 4-byte read.
#define KFDB_SMEM_BITMAP_BYTES (200 * 1024)
// word -> bit; indices past the bitmap (the 0xFFFFFFFF padding, words above the query's largest) are clamped onto a zero word
__device__ __forceinline__ unsigned kf_test_s(uint32_t w, const uint32_t* bm, uint32_t zero_idx) {
    return (bm[min(w >> 5, zero_idx)] >> (w & 31)) & 1u;
}
__device__ __forceinline__ unsigned kf_test4_s(const uint4 a, const uint32_t* bm, uint32_t zero_idx) {
    return kf_test_s(a.x, bm, zero_idx) + kf_test_s(a.y, bm, zero_idx) + kf_test_s(a.z, bm, zero_idx) + kf_test_s(a.w, bm, zero_idx);
}
#define KFDB_SCAN_THREADS 512
#define KFDB_SCAN_LOADS 8   // 16-byte loads per lane and round: a round covers 32 * 8 * 4 = 1024 words of a keyframe
__global__ void __launch_bounds__(KFDB_SCAN_THREADS, 1) kfdb_count_smem_kernel(const uint4* __restrict__ words4, const uint2* __restrict__ slots,
                                                                  const uint8_t* __restrict__ excl, int n_slots,
                                                                  const uint32_t* __restrict__ bitmap, uint32_t max_word,
                                                                  uint32_t* __restrict__ nobs, uint32_t* __restrict__ d_max,
                                                                  uint32_t* __restrict__ ticket, int static_rounds, int streaming) {
    extern __shared__ uint4 bm4[];
    const uint32_t* bm = (const uint32_t*)bm4;
    const int bm16 = (int)(((max_word >> 5) + 4) >> 2);          // bitmap length in 16-byte units (the workspace is padded)
    const uint32_t zero_idx = (uint32_t)bm16 * 4;                // one more 16-byte unit of zeros behind it
    const int lane = threadIdx.x & 31;
    const uint4 pad = make_uint4(~0u, ~0u, ~0u, ~0u);
    constexpr int R = 32 * KFDB_SCAN_LOADS;
    const int n_warps = (int)(gridDim.x * blockDim.x) >> 5;
    const int wid = (int)(blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int turn = 0;
    auto ld = [&](const uint4* q) { return streaming ? __ldcs(q) : __ldg(q); };
    // Software pipeline, two keyframes deep: while a round of 8 x 16 bytes per lane is tested, the next round (of this keyframe or
    // the first of the next one) is already in flight in a second register set, and the ticket / header of the keyframe after
    // that are on their way.  State: (slot, s) streams now; (nslot, ns) is next, its header already here.
    // keyframes are dealt round-robin for the first static_rounds turns of a warp (no traffic on the ticket word), the rest are
    // drawn from the ticket counter so that the tail balances
    auto take = [&]() {
        const int k = turn++;
        if (k < static_rounds) return wid + k * n_warps;
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(ticket, 1u);
        return static_rounds * n_warps + __shfl_sync(0xffffffffu, t, 0);
    };
    auto header = [&](int sl) {
        uint2 h = make_uint2(0, 0);
        if (sl < n_slots) {
            h = slots[sl];
            if (excl[sl]) h.y = 0;
        }
        return h;
    };
    int slot = take();
    int nslot = take();
    uint2 s = header(slot), ns = header(nslot);
    uint4 v[KFDB_SCAN_LOADS], nv[KFDB_SCAN_LOADS];
    {
        const uint4* p = words4 + s.x;
        const int n4 = (int)((s.y + 3) >> 2);
#pragma unroll
        for (int k = 0; k < KFDB_SCAN_LOADS; k++) v[k] = (lane + 32 * k < n4) ? ld(p + lane + 32 * k) : pad;
    }
    for (int i = threadIdx.x; i <= bm16; i += blockDim.x) bm4[i] = (i < bm16) ? __ldg((const uint4*)bitmap + i) : make_uint4(0, 0, 0, 0);
    __syncthreads();
    uint32_t local_max = 0;
    while (slot < n_slots) {
        const int nnslot = take();
        const uint4* p = words4 + s.x;
        const int n4 = (int)((s.y + 3) >> 2);
        uint32_t cnt = 0;
        for (int base = R; base < n4; base += R) {   // further rounds of this keyframe
#pragma unroll
            for (int k = 0; k < KFDB_SCAN_LOADS; k++) nv[k] = (base + lane + 32 * k < n4) ? ld(p + base + lane + 32 * k) : pad;
#pragma unroll
            for (int k = 0; k < KFDB_SCAN_LOADS; k++) cnt += kf_test4_s(v[k], bm, zero_idx);
#pragma unroll
            for (int k = 0; k < KFDB_SCAN_LOADS; k++) v[k] = nv[k];
        }
        {   // first round of the next keyframe goes out before the last round of this one is tested
            const uint4* np = words4 + ns.x;
            const int nn4 = (int)((ns.y + 3) >> 2);
#pragma unroll
            for (int k = 0; k < KFDB_SCAN_LOADS; k++) nv[k] = (lane + 32 * k < nn4) ? ld(np + lane + 32 * k) : pad;
        }
#pragma unroll
        for (int k = 0; k < KFDB_SCAN_LOADS; k++) cnt += kf_test4_s(v[k], bm, zero_idx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) nobs[slot] = cnt;
        local_max = max(local_max, cnt);
        const uint2 nns = header(nnslot);
        slot = nslot; s = ns;
        nslot = nnslot; ns = nns;
#pragma unroll
        for (int k = 0; k < KFDB_SCAN_LOADS; k++) v[k] = nv[k];
    }
    if (lane == 0 && local_max) atomicMax(d_max, local_max);
}

// step 2a (keyframedatabase.cpp:222-226): the frames with more than 0.8*max votes, as a work list
__global__ void __launch_bounds__(256) kfdb_select_kernel(const uint32_t* __restrict__ nobs, int n_slots, const uint32_t* __restrict__ d_max,
                                                          uint32_t* __restrict__ sel, uint32_t* __restrict__ n_sel) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t max_common = *d_max;
    const uint32_t min_common = __float2uint_rz(__fmul_rn(__uint2float_rn(max_common), 0.8f));   // uint32_t = maxCommonWords*0.8f
    const bool take = slot < n_slots && max_common != 0 && nobs[slot] > min_common;
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(n_sel, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) sel[base + __popc(m & ((1u << lane) - 1))] = (uint32_t)slot;
}

// step 2b (keyframedatabase.cpp:227-232): fBow::score of the listed frames.  One CTA per frame: all threads find the common words
// and their float products in parallel and park them IN WORD ORDER in shared memory (ballot + block scan), then one thread adds them
// into the double one by one -- the only part of fbow.cpp:209 whose order matters.
#define KFDB_SCORE_THREADS 1024
// a + p[0] + p[1] + ... one rounded double addition at a time, in index order; the loads run ahead of the dependent add chain
__device__ __forceinline__ double kf_ordered_sum(double a, const float* p, uint32_t n) {
    uint32_t k = 0;
    for (; k + 8 <= n; k += 8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; j++) t[j] = p[k + j];
#pragma unroll
        for (int j = 0; j < 8; j++) a = __dadd_rn(a, (double)t[j]);
    }
    for (; k < n; k++) a = __dadd_rn(a, (double)p[k]);
    return a;
}
#define KFDB_SCORE_CAP 4096
__global__ void __launch_bounds__(KFDB_SCORE_THREADS) kfdb_score_kernel(const uint32_t* __restrict__ words, const float* __restrict__ weights,
                                                         const uint2* __restrict__ slots, const uint32_t* __restrict__ frame,
                                                         const uint32_t* __restrict__ nobs,
                                                         const uint32_t* __restrict__ bitmap, uint32_t max_word,
                                                         const uint32_t* __restrict__ qwords, const float* __restrict__ qweights, int nq,
                                                         const uint32_t* __restrict__ sel, const uint32_t* __restrict__ n_sel,
                                                         float min_score, KfHit* __restrict__ out, int cap, uint32_t* __restrict__ n_out) {
    __shared__ float prod[KFDB_SCORE_CAP];
    __shared__ uint32_t warp_cnt[KFDB_SCORE_THREADS / 32];
    __shared__ uint32_t fill;
    __shared__ double acc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t total = *n_sel;
    for (uint32_t it = blockIdx.x; it < total; it += gridDim.x) {
        const uint32_t slot = sel[it];
        const uint2 s = slots[slot];
        const uint32_t* w = words + (size_t)s.x * 4;
        const float* wt = weights + (size_t)s.x * 4;
        if (threadIdx.x == 0) { fill = 0; acc = 0.0; }
        __syncthreads();
        for (uint32_t base = 0; base < s.y; base += KFDB_SCORE_THREADS) {
            const uint32_t i = base + threadIdx.x;
            float p = 0.f;
            bool hit = false;
            if (i < s.y) {
                const uint32_t wi = w[i];
                if (kf_test(wi, bitmap, max_word)) {
                    int lo = 0, hi = nq - 1;   // the word is in the query: find its weight
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (__ldg(qwords + mid) < wi) lo = mid + 1; else hi = mid;
                    }
                    p = __fmul_rn(__ldg(qweights + lo), wt[i]);   // float product (fBow values are floats), fbow.cpp:209
                    hit = true;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) warp_cnt[warp] = __popc(m);
            __syncthreads();
            uint32_t before = 0, tile = 0;
#pragma unroll
            for (int k = 0; k < KFDB_SCORE_THREADS / 32; k++) {
                const uint32_t c = warp_cnt[k];
                before += (k < warp) ? c : 0u;
                tile += c;
            }
            const uint32_t start = fill;
            if (start + tile > KFDB_SCORE_CAP) {   // flush what is parked (uniform branch: every thread sees the same counters)
                __syncthreads();
                if (threadIdx.x == 0) {
                    acc = kf_ordered_sum(acc, prod, start);
                    fill = 0;
                }
                __syncthreads();
            }
            const uint32_t at = (start + tile > KFDB_SCORE_CAP) ? 0u : start;
            if (hit) prod[at + before + __popc(m & ((1u << lane) - 1))] = p;
            __syncthreads();
            if (threadIdx.x == 0) fill = at + tile;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double score = kf_ordered_sum(acc, prod, fill);                         // ascending word order, like the two map iterators
            const double si = (score >= 1.0) ? 1.0 : 1.0 - sqrt(1.0 - score);             // fbow.cpp:237-240
            if (si > (double)min_score) {
                const uint32_t k = atomicAdd(n_out, 1u);
                if (k < (uint32_t)cap) out[k] = KfHit{frame[slot], nobs[slot], si};
            }
        }
        __syncthreads();
    }
}

int kfdb_grow_arena(uco_b200_ctx* ctx, uco_b200_kfdb* db, uint64_t need) {
    if (need <= db->cap) return UCO_OK;
    uint64_t ncap = std::max<uint64_t>(need + need / 2, 1u << 20);
    uint32_t* nw = nullptr;
    float* nf = nullptr;
    if (cudaMalloc(&nw, ncap * 4) != cudaSuccess || cudaMalloc(&nf, ncap * 4) != cudaSuccess) {
        cudaGetLastError();
        if (nw) cudaFree(nw);
        return uco_fail(ctx, UCO_E_NOMEM, "kfdb: cudaMalloc of %llu entries failed", (unsigned long long)ncap);
    }
    if (db->used) {
        UCO_CUDA(ctx, cudaMemcpyAsync(nw, db->d_words, db->used * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        UCO_CUDA(ctx, cudaMemcpyAsync(nf, db->d_weights, db->used * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(db->d_words);
    cudaFree(db->d_weights);
    db->d_words = nw;
    db->d_weights = nf;
    db->cap = ncap;
    return UCO_OK;
}

int kfdb_grow_slots(uco_b200_ctx* ctx, uco_b200_kfdb* db, uint32_t need) {
    if (need <= db->slot_cap) return UCO_OK;
    uint32_t ncap = std::max<uint32_t>(need + need / 2, 1024);
    uint2* ns = nullptr;
    uint32_t *nfr = nullptr, *nn = nullptr;
    uint8_t* ne = nullptr;
    if (cudaMalloc(&ns, (size_t)ncap * 8) != cudaSuccess || cudaMalloc(&nfr, (size_t)ncap * 4) != cudaSuccess ||
        cudaMalloc(&nn, (size_t)ncap * 4) != cudaSuccess || cudaMalloc(&ne, ncap) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(ns); cudaFree(nfr); cudaFree(nn); cudaFree(ne);
        return uco_fail(ctx, UCO_E_NOMEM, "kfdb: cudaMalloc of %u slots failed", ncap);
    }
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(db->d_slots); cudaFree(db->d_frame); cudaFree(db->d_nobs); cudaFree(db->d_excl);
    db->d_slots = ns; db->d_frame = nfr; db->d_nobs = nn; db->d_excl = ne;
    db->slot_cap = ncap;
    if (!db->slots.empty()) {   // the host mirror is authoritative
        UCO_CUDA(ctx, cudaMemcpyAsync(ns, db->slots.data(), db->slots.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        UCO_CUDA(ctx, cudaMemcpyAsync(nfr, db->frame.data(), db->frame.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return UCO_OK;
}

// drop the segments of deleted frames (run when more than half of the arena is dead)
int kfdb_compact(uco_b200_ctx* ctx, uco_b200_kfdb* db) {
    uint64_t live_entries = 0;
    for (auto& s : db->slots) live_entries += (uint64_t)((s.y + 3) & ~3u);
    const uint64_t ncap = std::max<uint64_t>(live_entries + live_entries / 2, 1u << 20);
    uint32_t* nw = nullptr;
    float* nf = nullptr;
    if (cudaMalloc(&nw, ncap * 4) != cudaSuccess || cudaMalloc(&nf, ncap * 4) != cudaSuccess) {
        cudaGetLastError();
        if (nw) cudaFree(nw);
        return UCO_OK;   // no room to compact now: keep the sparse arena
    }
    std::vector<uint2> nslots;
    std::vector<uint32_t> nframe;
    uint64_t pos = 0;
    for (size_t i = 0; i < db->slots.size(); i++) {
        const uint2 s = db->slots[i];
        if (!db->alive[i]) continue;
        const uint64_t len4 = (s.y + 3) & ~3u;
        if (len4) UCO_CUDA(ctx, cudaMemcpyAsync(nw + pos, db->d_words + (uint64_t)s.x * 4, len4 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        if (len4) UCO_CUDA(ctx, cudaMemcpyAsync(nf + pos, db->d_weights + (uint64_t)s.x * 4, len4 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        nslots.push_back(make_uint2((uint32_t)(pos / 4), s.y));
        nframe.push_back(db->frame[i]);
        pos += len4;
    }
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(db->d_words);
    cudaFree(db->d_weights);
    db->d_words = nw; db->d_weights = nf; db->cap = ncap; db->used = pos; db->dead = 0;
    db->slots.swap(nslots);
    db->frame.swap(nframe);
    db->alive.assign(db->frame.size(), 1);
    db->slot_of.clear();
    for (size_t i = 0; i < db->frame.size(); i++) db->slot_of[db->frame[i]] = (uint32_t)i;
    if (!db->slots.empty()) {
        UCO_CUDA(ctx, cudaMemcpyAsync(db->d_slots, db->slots.data(), db->slots.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        UCO_CUDA(ctx, cudaMemcpyAsync(db->d_frame, db->frame.data(), db->frame.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return UCO_OK;
}

}  // namespace

extern "C" {

int uco_b200_kfdb_create(uco_b200_ctx* ctx, uco_b200_kfdb** out) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!out) return uco_fail(ctx, UCO_E_INVALID, "kfdb_create: null pointer");
    uco_b200_kfdb* db = new uco_b200_kfdb();
    db->device = ctx->device;
    *out = db;
    return UCO_OK;
}

void uco_b200_kfdb_free(uco_b200_ctx* ctx, uco_b200_kfdb* db) {
    if (!db) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(db->d_words); cudaFree(db->d_weights); cudaFree(db->d_slots); cudaFree(db->d_frame); cudaFree(db->d_nobs);
    cudaFree(db->d_excl);
    for (auto& e : db->ev)
        if (e) cudaEventDestroy(e);
    delete db;
}

int uco_b200_kfdb_clear(uco_b200_ctx* ctx, uco_b200_kfdb* db) {
    if (!ctx) return UCO_E_INVALID;
    if (!db) return uco_fail(ctx, UCO_E_INVALID, "kfdb_clear: no database");
    cudaSetDevice(ctx->device);
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    db->slots.clear(); db->frame.clear(); db->alive.clear(); db->slot_of.clear();
    db->used = db->dead = 0;
    db->live = 0;
    return UCO_OK;
}

int uco_b200_kfdb_size(const uco_b200_kfdb* db, uint32_t* n_frames, uint64_t* n_words) {
    if (!db) return UCO_E_INVALID;
    if (n_frames) *n_frames = db->live;
    if (n_words) {
        uint64_t t = 0;
        for (auto& s : db->slots) t += s.y;
        *n_words = t;
    }
    return UCO_OK;
}

int uco_b200_kfdb_has(const uco_b200_kfdb* db, uint32_t frame_id) { return db && db->slot_of.count(frame_id) ? 1 : 0; }

int uco_b200_kfdb_add_batch(uco_b200_ctx* ctx, uco_b200_kfdb* db, int n_frames, const uint32_t* frame_ids, const int32_t* counts,
                            const uint32_t* words, const float* weights) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!db || n_frames < 0 || (n_frames && (!frame_ids || !counts))) return uco_fail(ctx, UCO_E_INVALID, "kfdb_add: null pointer");
    if (db->device != ctx->device) return uco_fail(ctx, UCO_E_INVALID, "kfdb_add: the database lives on another device");
    // validate: ids new and distinct, word lists strictly ascending (they are std::map keys in the reference), ids below 2^32-1
    uint64_t total = 0, padded = 0;
    {
        std::unordered_map<uint32_t, int> seen;
        const uint32_t* w = words;
        for (int f = 0; f < n_frames; f++) {
            if (counts[f] < 0) return uco_fail(ctx, UCO_E_INVALID, "kfdb_add: negative word count");
            if (db->slot_of.count(frame_ids[f]) || seen.count(frame_ids[f]))
                return uco_fail(ctx, UCO_E_INVALID, "kfdb_add: frame %u is already in the database", frame_ids[f]);
            seen[frame_ids[f]] = 1;
            if (counts[f] && (!words || !weights)) return uco_fail(ctx, UCO_E_INVALID, "kfdb_add: null pointer");
            for (int i = 0; i < counts[f]; i++) {
                if (w[i] == 0xFFFFFFFFu || (i && w[i] <= w[i - 1]))
                    return uco_fail(ctx, UCO_E_INVALID, "kfdb_add: the words of frame %u are not strictly ascending", frame_ids[f]);
            }
            w += counts[f];
            total += (uint64_t)counts[f];
            padded += ((uint64_t)counts[f] + 3) & ~3ull;
        }
    }
    if (n_frames == 0) return UCO_OK;
    if ((db->used + padded) / 4 > 0xFFFFFFFFull) return uco_fail(ctx, UCO_E_CAPACITY, "kfdb_add: arena index overflow");
    int rc = kfdb_grow_arena(ctx, db, db->used + padded);
    if (rc != UCO_OK) return rc;
    rc = kfdb_grow_slots(ctx, db, (uint32_t)db->slots.size() + (uint32_t)n_frames);
    if (rc != UCO_OK) return rc;
    // stage the padded segments in pinned memory, one transfer per array
    uint32_t* hw = (uint32_t*)uco_pinned(ctx, WS_KFDB_STAGE, padded * 8 + 16);
    if (!hw) return UCO_E_NOMEM;
    float* hf = (float*)(hw + padded);
    uint64_t pos = 0, src = 0;
    const size_t first_slot = db->slots.size();
    for (int f = 0; f < n_frames; f++) {
        const uint64_t len4 = ((uint64_t)counts[f] + 3) & ~3ull;
        if (counts[f]) {
            memcpy(hw + pos, words + src, (size_t)counts[f] * 4);
            memcpy(hf + pos, weights + src, (size_t)counts[f] * 4);
        }
        for (uint64_t i = counts[f]; i < len4; i++) {
            hw[pos + i] = 0xFFFFFFFFu;
            hf[pos + i] = 0.f;
        }
        db->slots.push_back(make_uint2((uint32_t)((db->used + pos) / 4), (uint32_t)counts[f]));
        db->frame.push_back(frame_ids[f]);
        db->alive.push_back(1);
        db->slot_of[frame_ids[f]] = (uint32_t)(db->slots.size() - 1);
        pos += len4;
        src += counts[f];
    }
    (void)total;
    if (padded) {
        UCO_CUDA(ctx, cudaMemcpyAsync(db->d_words + db->used, hw, padded * 4, cudaMemcpyHostToDevice, ctx->stream));
        UCO_CUDA(ctx, cudaMemcpyAsync(db->d_weights + db->used, hf, padded * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    UCO_CUDA(ctx, cudaMemcpyAsync(db->d_slots + first_slot, db->slots.data() + first_slot, (size_t)n_frames * 8, cudaMemcpyHostToDevice,
                                  ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(db->d_frame + first_slot, db->frame.data() + first_slot, (size_t)n_frames * 4, cudaMemcpyHostToDevice,
                                  ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    db->used += padded;
    db->live += (uint32_t)n_frames;
    return UCO_OK;
}

int uco_b200_kfdb_add(uco_b200_ctx* ctx, uco_b200_kfdb* db, uint32_t frame_id, const uint32_t* words, const float* weights, int n) {
    const int32_t cnt = n;
    return uco_b200_kfdb_add_batch(ctx, db, 1, &frame_id, &cnt, words, weights);
}

int uco_b200_kfdb_del(uco_b200_ctx* ctx, uco_b200_kfdb* db, uint32_t frame_id) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!db) return uco_fail(ctx, UCO_E_INVALID, "kfdb_del: no database");
    auto it = db->slot_of.find(frame_id);
    if (it == db->slot_of.end()) return uco_fail(ctx, UCO_E_INVALID, "kfdb_del: frame %u is not in the database", frame_id);
    const uint32_t slot = it->second;
    db->dead += (db->slots[slot].y + 3) & ~3u;
    db->slots[slot].y = 0;
    db->alive[slot] = 0;
    db->slot_of.erase(it);
    db->live--;
    UCO_CUDA(ctx, cudaMemcpyAsync(db->d_slots + slot, &db->slots[slot], 8, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (db->dead * 2 > db->used && db->used > (1u << 16)) return kfdb_compact(ctx, db);
    return UCO_OK;
}

int uco_b200_kfdb_query(uco_b200_ctx* ctx, uco_b200_kfdb* db, const uint32_t* words, const float* weights, int n,
                        const uint32_t* excluded, int n_excluded, float min_score, uint32_t* out_frame, double* out_score,
                        uint32_t* out_common, int cap, int* n_out, uint32_t* max_common) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!db || !n_out || n < 0 || n_excluded < 0 || cap < 0) return uco_fail(ctx, UCO_E_INVALID, "kfdb_query: bad argument");
    if ((n && (!words || !weights)) || (n_excluded && !excluded) || (cap && (!out_frame || !out_score)))
        return uco_fail(ctx, UCO_E_INVALID, "kfdb_query: null pointer");
    if (db->device != ctx->device) return uco_fail(ctx, UCO_E_INVALID, "kfdb_query: the database lives on another device");
    for (int i = 0; i < n; i++)
        if (words[i] == 0xFFFFFFFFu || (i && words[i] <= words[i - 1]))
            return uco_fail(ctx, UCO_E_INVALID, "kfdb_query: the query words are not strictly ascending");
    *n_out = 0;
    if (max_common) *max_common = 0;
    db->last_ms[0] = db->last_ms[1] = 0.f;
    const int n_slots = (int)db->slots.size();
    if (n == 0 || n_slots == 0 || db->live == 0) return UCO_OK;   // frame_nobs.size()==0 -> {}  (keyframedatabase.cpp:221)
    const uint32_t max_word = words[n - 1];
    const size_t bm_words = (((size_t)(max_word >> 5) + 1) + 3) & ~(size_t)3;   // whole 16-byte units (the scan copies it as uint4)
    std::vector<uint32_t> excl_slots;
    for (int i = 0; i < n_excluded; i++) {
        auto it = db->slot_of.find(excluded[i]);
        if (it != db->slot_of.end()) excl_slots.push_back(it->second);
    }
    const int ne = (int)excl_slots.size();
    // query words | weights | excluded slots in one pinned block -> one device block
    const size_t qbytes = (size_t)n * 8 + (size_t)ne * 4;
    uint8_t* hq = (uint8_t*)uco_pinned(ctx, WS_KFDB_Q, qbytes);
    uint8_t* dq = (uint8_t*)uco_ws(ctx, WS_KFDB_Q, qbytes);
    uint32_t* bitmap = (uint32_t*)uco_ws(ctx, WS_KFDB_BITMAP, bm_words * 4);
    const int hit_cap = n_slots;
    uint8_t* dout = (uint8_t*)uco_ws(ctx, WS_KFDB_OUT, 16 + (size_t)hit_cap * sizeof(KfHit));
    uint8_t* hout = (uint8_t*)uco_pinned(ctx, WS_KFDB_OUT, 16 + (size_t)hit_cap * sizeof(KfHit));
    uint32_t* d_sel = (uint32_t*)uco_ws(ctx, WS_KFDB_STAGE, (size_t)n_slots * 4);
    if (!hq || !dq || !bitmap || !dout || !hout || !d_sel) return UCO_E_NOMEM;
    memcpy(hq, words, (size_t)n * 4);
    memcpy(hq + (size_t)n * 4, weights, (size_t)n * 4);
    if (ne) memcpy(hq + (size_t)n * 8, excl_slots.data(), (size_t)ne * 4);
    const uint32_t* d_qw = (const uint32_t*)dq;
    const float* d_qf = (const float*)(dq + (size_t)n * 4);
    const uint32_t* d_ex = (const uint32_t*)(dq + (size_t)n * 8);
    uint32_t* d_max = (uint32_t*)dout;          // [0] max votes, [1] number of hits, [2] scan ticket, [3] number of selected frames
    uint32_t* d_nhit = d_max + 1;
    uint32_t* d_ticket = d_max + 2;
    uint32_t* d_nsel = d_max + 3;
    KfHit* d_hits = (KfHit*)(dout + 16);
    const bool prof = ctx->profiling != 0;
    cudaEvent_t* ev = db->ev;   // owned by the database (created once): no leak on the error returns below
    if (prof && !ev[0])
        for (int k = 0; k < 3; k++) cudaEventCreate(&ev[k]);
    UCO_CUDA(ctx, cudaMemcpyAsync(dq, hq, qbytes, cudaMemcpyHostToDevice, ctx->stream));
    UCO_CUDA(ctx, cudaMemsetAsync(bitmap, 0, bm_words * 4, ctx->stream));
    UCO_CUDA(ctx, cudaMemsetAsync(db->d_excl, 0, n_slots, ctx->stream));
    UCO_CUDA(ctx, cudaMemsetAsync(dout, 0, 16, ctx->stream));
    const int nm = std::max(n, ne);
    kfdb_mark_kernel<<<(nm + 255) / 256, 256, 0, ctx->stream>>>(d_qw, n, bitmap, d_ex, ne, db->d_excl);
    UCO_LAUNCH_CHECK(ctx);
    if (prof) cudaEventRecord(ev[0], ctx->stream);
    if (bm_words * 4 + 16 <= KFDB_SMEM_BITMAP_BYTES) {
        // per device (a process may drive several GPUs), so set on every call: it is a host-side table update
        UCO_CUDA(ctx, cudaFuncSetAttribute(kfdb_count_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KFDB_SMEM_BITMAP_BYTES));
        const int blocks = std::min((n_slots + KFDB_SCAN_THREADS / 32 - 1) / (KFDB_SCAN_THREADS / 32), ctx->sm_count);
        const int n_warps = blocks * (KFDB_SCAN_THREADS / 32);
        // three quarters of the keyframes are dealt round-robin, the tail is ticketed (measured on B200, 20 k keyframes: all tickets
        // 42.0 us, all dealt 40.0 us, 3/4 + tickets 38.9 us; ld.global.cs vs ld.global.nc makes no difference)
        const int static_rounds = (int)((long long)n_slots * 3 / 4 / n_warps), streaming = 1;
        kfdb_count_smem_kernel<<<blocks, KFDB_SCAN_THREADS, bm_words * 4 + 16, ctx->stream>>>((const uint4*)db->d_words, db->d_slots, db->d_excl, n_slots,
                                                                           bitmap, max_word, db->d_nobs, d_max, d_ticket, static_rounds,
                                                                           streaming);
    } else {
        const int blocks = std::min((n_slots + 7) / 8, ctx->sm_count * 8);
        kfdb_count_kernel<<<blocks, 256, 0, ctx->stream>>>((const uint4*)db->d_words, db->d_slots, db->d_excl, n_slots, bitmap, max_word,
                                                           db->d_nobs, d_max);
    }
    UCO_LAUNCH_CHECK(ctx);
    if (prof) cudaEventRecord(ev[1], ctx->stream);
    kfdb_select_kernel<<<(n_slots + 255) / 256, 256, 0, ctx->stream>>>(db->d_nobs, n_slots, d_max, d_sel, d_nsel);
    UCO_LAUNCH_CHECK(ctx);
    kfdb_score_kernel<<<std::min(n_slots, ctx->sm_count * 2), KFDB_SCORE_THREADS, 0, ctx->stream>>>(
        db->d_words, db->d_weights, db->d_slots, db->d_frame, db->d_nobs, bitmap, max_word, d_qw, d_qf, n, d_sel, d_nsel, min_score,
        d_hits, hit_cap, d_nhit);
    UCO_LAUNCH_CHECK(ctx);
    if (prof) cudaEventRecord(ev[2], ctx->stream);
    // the common case returns a handful of frames: fetch the counters and the first hits in one transfer, the rest if needed
    const int first = std::min(hit_cap, 256);
    UCO_CUDA(ctx, cudaMemcpyAsync(hout, dout, 16 + (size_t)first * sizeof(KfHit), cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t nhit = ((uint32_t*)hout)[1];
    if (max_common) *max_common = ((uint32_t*)hout)[0];
    if ((int)nhit > first) {
        UCO_CUDA(ctx, cudaMemcpyAsync(hout + 16 + (size_t)first * sizeof(KfHit), dout + 16 + (size_t)first * sizeof(KfHit),
                                      (size_t)(nhit - first) * sizeof(KfHit), cudaMemcpyDeviceToHost, ctx->stream));
        UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (prof) {
        cudaEventElapsedTime(&db->last_ms[0], ev[0], ev[1]);
        cudaEventElapsedTime(&db->last_ms[1], ev[1], ev[2]);
    }
    *n_out = (int)nhit;
    if ((int)nhit > cap) return uco_fail(ctx, UCO_E_CAPACITY, "kfdb_query: %u frames scored, capacity %d", nhit, cap);
    KfHit* h = (KfHit*)(hout + 16);
    std::sort(h, h + nhit, [](const KfHit& a, const KfHit& b) { return a.frame < b.frame; });   // std::map<frame, score> order
    for (uint32_t i = 0; i < nhit; i++) {
        out_frame[i] = h[i].frame;
        out_score[i] = h[i].score;
        if (out_common) out_common[i] = h[i].common;
    }
    return UCO_OK;
}

int uco_b200_kfdb_last_ms(const uco_b200_kfdb* db, float* out2) {
    if (!db || !out2) return UCO_E_INVALID;
    out2[0] = db->last_ms[0];
    out2[1] = db->last_ms[1];
    return UCO_OK;
}

// steps 3-4 of relocalizationCandidates (keyframedatabase.cpp:236-275) on the scored frames; pure host arithmetic on a handful of
// entries.  nbr lists = what CovisGraph::getNeighborsWeights(frame, true) returns for each scored frame (ids, decreasing weight).
int uco_b200_kfdb_rank(const uint32_t* frame, const double* score, int n, const int32_t* nbr_off, const uint32_t* nbr, int sorted,
                       float min_score, uint32_t* out, int* n_out) {
    if (n < 0 || !n_out || (n && (!frame || !score || !out))) return UCO_E_INVALID;
    *n_out = 0;
    if (n == 0) return UCO_OK;                       // :236
    if (n == 1) {                                    // :237
        out[0] = frame[0];
        *n_out = 1;
        return UCO_OK;
    }
    if (!nbr_off) return UCO_E_INVALID;
    std::vector<std::pair<uint32_t, double>> acc;
    acc.reserve(n);
    double best = min_score;                         // :240
    for (int i = 0; i < n; i++) {
        double a = score[i];
        const int lo = nbr_off[i], hi = std::min(nbr_off[i + 1], nbr_off[i] + 10);   // the 10 best neighbours, :249
        for (int k = lo; k < hi; k++) {
            const uint32_t* it = std::lower_bound(frame, frame + n, nbr[k]);         // frame_score.find
            if (it != frame + n && *it == nbr[k]) a += score[it - frame];
        }
        acc.push_back(std::make_pair(frame[i], a));
        if (a > best) best = a;
    }
    const double min_retain = 0.75f * best;          // :262
    acc.erase(std::remove_if(acc.begin(), acc.end(), [&](const std::pair<uint32_t, double>& v) { return v.second < min_retain; }),
              acc.end());
    if (sorted)
        std::sort(acc.begin(), acc.end(),
                  [&](const std::pair<uint32_t, double>& a, const std::pair<uint32_t, double>& b) { return a.second > b.second; });
    for (size_t i = 0; i < acc.size(); i++) out[i] = acc[i].first;
    *n_out = (int)acc.size();
    return UCO_OK;
}

}  // extern "C"
