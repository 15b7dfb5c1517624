// bow.cu — K9: bag-of-words transform of 32-byte ORB descriptors through an fbow vocabulary tree.
//
// Replaces (reference, relative to /root/reference):
//   3rdparty/fbow/fbow/fbow.cpp:51-90     Vocabulary::transform(features, level, fBow&, fBow2&)
//   3rdparty/fbow/fbow/fbow.h:402-448     _transform2<L1_32bytes>: descend from block 0, at every block take the child with
//                                         the FIRST minimum Hamming distance, stop at a leaf (word id + weight)
//   3rdparty/fbow/fbow/fbow.cpp:180-190   Vocabulary::fromStream (the .fbow stream is uploaded as it is)
// as called by KPFrameDataBase::computeBow (src/map_types/keyframedatabase.cpp:310-321, level 3).
//
// Layout: the vocabulary's block array is copied to HBM byte for byte (blocks of block_size bytes: u16 N | u16 isLeaf |
// u32 parent | k features of 32 B | k x (u32 id_or_child, f32 weight)); at 45 MB the shipped ORB vocabulary stays resident in
// the 126 MB L2 after first touch.  A group of 16 lanes walks one descriptor: lane c holds child c's distance (8 x __popc),
// the first-minimum is a 4-step shuffle reduction on (dist << 8 | c), so one level costs one dependent L2 round trip.
// The kernel emits per DESCRIPTOR (word, weight, level-node); folding them into the reference's std::map containers in
// descriptor order (so that float weight sums are bit-identical) is done by the host adapter.
#include "common.cuh"
#include <cmath>
#include <cstring>

struct uco_b200_voc {
    uint8_t* d_data = nullptr;
    uint32_t alignment = 0, nblocks = 0, k = 0;
    uint64_t desc_size_wp = 0, block_size = 0, feature_off = 0, child_off = 0, total_size = 0;
    int32_t desc_type = 0, desc_size = 0;
    int nbits = 0;
};

namespace {

struct VocDev {
    const uint8_t* data;
    unsigned block_size, feature_off, child_off, desc_wp, k, nblocks;
    int nbits;
};

#define BOW_GROUP 16
#define BOW_MAX_DEPTH 64

__global__ void __launch_bounds__(256) bow_transform_kernel(const __grid_constant__ VocDev V, const uint2* __restrict__ desc,
                                                            int n, int store_level, uint32_t* __restrict__ word,
                                                            float* __restrict__ weight, uint32_t* __restrict__ node,
                                                            int* __restrict__ err, const int* __restrict__ seg_sel, int seg_rows,
                                                            size_t seg_stride8, const int* __restrict__ seg_n) {
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / BOW_GROUP;
    const int c = threadIdx.x & (BOW_GROUP - 1);
    const unsigned gmask = 0xFFFFu << (threadIdx.x & 16);   // the two 16-lane groups of a warp leave the loop independently
    // plain form: descriptor gid of a dense array.  Segmented form (seg_sel): output slot gid = (segment gid / seg_rows, row
    // gid % seg_rows); the segment's descriptors are those of frame seg_sel[segment] in a frame-strided buffer, seg_n[frame] of them
    bool active = gid < n;
    const int f = gid < n ? gid : n - 1;   // inactive groups shadow the last descriptor so that shuffles stay convergent
    const uint2* src = desc + (size_t)f * 4;
    if (seg_sel) {
        const int fr = seg_sel[f / seg_rows], r = f % seg_rows, nr = seg_n ? min(seg_n[fr], seg_rows) : seg_rows;
        active = active && r < nr;
        src = desc + (size_t)fr * seg_stride8 + (size_t)(r < nr ? r : 0) * 4;
    }
    uint2 q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = src[i];
    uint32_t block = 0, level = 0, cur_node = 0, best_idx = 0;
    uint32_t out_word = 0xFFFFFFFFu, out_node = 0xFFFFFFFFu;
    float out_w = 0.f;
    for (int it = 0; it < BOW_MAX_DEPTH; it++) {
        const uint8_t* b = V.data + (size_t)block * V.block_size;
        const unsigned N = *(const unsigned short*)b;
        unsigned key = 0xFFFFFFFFu;
        for (unsigned c0 = 0; c0 < N; c0 += BOW_GROUP) {   // k <= 16 in practice: one pass
            unsigned cc = c0 + c;
            if (cc < N) {
                const uint2* ft = (const uint2*)(b + V.feature_off + (size_t)cc * V.desc_wp);
                unsigned d = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint2 t = ft[i];
                    d += __popc(t.x ^ q[i].x) + __popc(t.y ^ q[i].y);
                }
                key = min(key, (d << 16) | cc);             // strict '<' in child order == min over (dist, child)
            }
        }
#pragma unroll
        for (int o = BOW_GROUP / 2; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(gmask, key, o, BOW_GROUP));
        if (key != 0xFFFFFFFFu) best_idx = key & 0xffffu;   // an empty block keeps the previous index (fbow.h:421-426)
        if (level == (uint32_t)store_level) out_node = cur_node;
        const uint2 info = *(const uint2*)(b + V.child_off + 8u * best_idx);
        if (info.x & 0x80000000u) {
            out_word = info.x & 0x7FFFFFFFu;
            out_w = __uint_as_float(info.y);
            if (level < (uint32_t)store_level) out_node = cur_node;
            break;
        }
        block = info.x & 0x7FFFFFFFu;
        cur_node = (cur_node << V.nbits) | best_idx;
        level++;
        if (block == 0) break;
        if (block >= V.nblocks || it == BOW_MAX_DEPTH - 1) {
            if (c == 0) atomicExch(err, 1);
            break;
        }
    }
    if (active && c == 0) {
        word[gid] = out_word;
        weight[gid] = out_w;
        node[gid] = out_node;
    }
}

}  // namespace

extern "C" {

int uco_b200_bow_load(uco_b200_ctx* ctx, const void* bytes, size_t n, uco_b200_voc** voc_out) {
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (!bytes || !voc_out) return uco_fail(ctx, UCO_E_INVALID, "bow_load: null pointer");
    const uint8_t* p = (const uint8_t*)bytes;
    uint64_t sig = 0;
    if (n < 128) return uco_fail(ctx, UCO_E_FORMAT, "Vocabulary::fromStream invalid signature");
    memcpy(&sig, p, 8);
    if (sig != 55824124ull) return uco_fail(ctx, UCO_E_FORMAT, "Vocabulary::fromStream invalid signature");
    uco_b200_voc* v = new uco_b200_voc();
    p += 8;
    memcpy(&v->alignment, p + 52, 4);
    memcpy(&v->nblocks, p + 56, 4);
    memcpy(&v->desc_size_wp, p + 64, 8);
    memcpy(&v->block_size, p + 72, 8);
    memcpy(&v->feature_off, p + 80, 8);
    memcpy(&v->child_off, p + 88, 8);
    memcpy(&v->total_size, p + 96, 8);
    memcpy(&v->desc_type, p + 104, 4);
    memcpy(&v->desc_size, p + 108, 4);
    memcpy(&v->k, p + 112, 4);
    auto bad = [&](const char* why) {
        delete v;
        return uco_fail(ctx, UCO_E_FORMAT, "bow_load: %s", why);
    };
    if (v->desc_type != 0 || v->desc_size != 32) return bad("only CV_8UC1 32-byte (ORB) vocabularies are supported");
    if (v->nblocks == 0 || v->k == 0 || v->k > 65535) return bad("empty vocabulary");
    if (v->total_size != v->block_size * v->nblocks || n < 128 + v->total_size) return bad("truncated stream");
    if ((v->block_size & 7) || (v->feature_off & 7) || (v->desc_size_wp & 7) || (v->child_off & 7))
        return bad("blocks are not 8-byte aligned");
    if (v->feature_off + v->k * v->desc_size_wp > v->block_size || v->child_off + 8ull * v->k > v->block_size)
        return bad("inconsistent block layout");
    v->nbits = (int)ceil(log2((double)v->k));
    cudaError_t e = cudaMalloc(&v->d_data, v->total_size);
    if (e != cudaSuccess) {
        delete v;
        return uco_fail(ctx, UCO_E_NOMEM, "bow_load: cudaMalloc(%llu) -> %s", (unsigned long long)v->total_size,
                        cudaGetErrorString(e));
    }
    e = cudaMemcpyAsync(v->d_data, (const uint8_t*)bytes + 128, v->total_size, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(v->d_data);
        delete v;
        return uco_fail(ctx, UCO_E_CUDA, "bow_load: upload -> %s", cudaGetErrorString(e));
    }
    *voc_out = v;
    return UCO_OK;
}

void uco_b200_bow_free(uco_b200_ctx* ctx, uco_b200_voc* voc) {
    if (!voc) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(voc->d_data);
    delete voc;
}

int uco_b200_bow_info(const uco_b200_voc* voc, uint32_t* k, uint32_t* nblocks, uint32_t* desc_size) {
    if (!voc) return UCO_E_INVALID;
    if (k) *k = voc->k;
    if (nblocks) *nblocks = voc->nblocks;
    if (desc_size) *desc_size = (uint32_t)voc->desc_size;
    return UCO_OK;
}

int uco_b200_bow_transform_dev(uco_b200_ctx* ctx, const uco_b200_voc* voc, const uint8_t* desc_dev, int n, int level,
                               uint32_t* word_dev, float* weight_dev, uint32_t* node_dev) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (!voc) return uco_fail(ctx, UCO_E_INVALID, "bow_transform: no vocabulary");
    if (n <= 0) return uco_fail(ctx, UCO_E_INVALID, "Vocabulary::transform No input data");   // fbow.cpp:52
    if (!desc_dev || !word_dev || !weight_dev || !node_dev) return uco_fail(ctx, UCO_E_INVALID, "bow_transform: null pointer");
    if ((uintptr_t)desc_dev & 7) return uco_fail(ctx, UCO_E_INVALID, "bow_transform: descriptors must be 8-byte aligned");
    int* err = (int*)uco_ws(ctx, WS_GENERIC0, sizeof(int));
    if (!err) return UCO_E_NOMEM;
    UCO_CUDA(ctx, cudaMemsetAsync(err, 0, sizeof(int), ctx->stream));
    VocDev V{voc->d_data, (unsigned)voc->block_size, (unsigned)voc->feature_off, (unsigned)voc->child_off,
             (unsigned)voc->desc_size_wp, voc->k, voc->nblocks, voc->nbits};
    const int groups_per_block = 256 / BOW_GROUP;
    bow_transform_kernel<<<(n + groups_per_block - 1) / groups_per_block, 256, 0, ctx->stream>>>(
        V, (const uint2*)desc_dev, n, level, word_dev, weight_dev, node_dev, err, nullptr, 0, 0, nullptr);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}
}  // extern "C"

// internal (match.cu, keyframes batch): the descriptors of n_seg selected frames of a frame-strided device buffer in ONE launch;
// outputs at [segment * seg_rows + row]; the error word is the caller's
int uco_bow_transform_segments(uco_b200_ctx* ctx, const uco_b200_voc* voc, const uint8_t* desc_dev, size_t frame_stride_bytes, const int* n_dev,
                               const int* seg_sel_dev, int n_seg, int seg_rows, int level, uint32_t* word_dev, float* weight_dev,
                               uint32_t* node_dev, int* err_dev) {
    if (!voc) return uco_fail(ctx, UCO_E_INVALID, "bow_transform: no vocabulary");
    if ((uintptr_t)desc_dev & 7 || frame_stride_bytes & 7) return uco_fail(ctx, UCO_E_INVALID, "bow_transform: descriptors must be 8-byte aligned");
    VocDev V{voc->d_data, (unsigned)voc->block_size, (unsigned)voc->feature_off, (unsigned)voc->child_off,
             (unsigned)voc->desc_size_wp, voc->k, voc->nblocks, voc->nbits};
    const int n = n_seg * seg_rows, groups_per_block = 256 / BOW_GROUP;
    bow_transform_kernel<<<(n + groups_per_block - 1) / groups_per_block, 256, 0, ctx->stream>>>(
        V, (const uint2*)desc_dev, n, level, word_dev, weight_dev, node_dev, err_dev, seg_sel_dev, seg_rows, frame_stride_bytes / 8, n_dev);
    UCO_LAUNCH_CHECK(ctx);
    return UCO_OK;
}

extern "C" {

int uco_b200_bow_transform(uco_b200_ctx* ctx, const uco_b200_voc* voc, const uint8_t* desc, int n, size_t stride, int level,
                           uint32_t* word, float* weight, uint32_t* node) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);  // the calling thread may be a new one (mapper / tracker threads): bind it to the context's GPU
    if (!voc) return uco_fail(ctx, UCO_E_INVALID, "bow_transform: no vocabulary");
    if (n <= 0) return uco_fail(ctx, UCO_E_INVALID, "Vocabulary::transform No input data");
    if (!desc || !word || !weight || !node) return uco_fail(ctx, UCO_E_INVALID, "bow_transform: null pointer");
    if (stride < 32) return uco_fail(ctx, UCO_E_INVALID, "bow_transform: row stride below 32 bytes");
    uint8_t* dd = (uint8_t*)uco_ws(ctx, WS_BOW_DESC, (size_t)n * 32);
    uint8_t* out = (uint8_t*)uco_ws(ctx, WS_BOW_OUT, (size_t)n * 12);
    if (!dd || !out) return UCO_E_NOMEM;
    uint32_t* dw = (uint32_t*)out;
    float* dwt = (float*)(out + (size_t)n * 4);
    uint32_t* dn = (uint32_t*)(out + (size_t)n * 8);
    UCO_CUDA(ctx, cudaMemcpy2DAsync(dd, 32, desc, stride, 32, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = uco_b200_bow_transform_dev(ctx, voc, dd, n, level, dw, dwt, dn);
    if (rc != UCO_OK) return rc;
    int herr = 0;
    UCO_CUDA(ctx, cudaMemcpyAsync(word, dw, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(weight, dwt, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(node, dn, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaMemcpyAsync(&herr, (int*)ctx->dev[WS_GENERIC0].p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    UCO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (herr) return uco_fail(ctx, UCO_E_FORMAT, "bow_transform: malformed vocabulary (cycle or block index out of range)");
    return UCO_OK;
}

}  // extern "C"
