/*
 * bow_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle). Plain-C restatement of the reference's bag-of-words transform:
 *   /root/reference/3rdparty/fbow/fbow/fbow.cpp:169-190   Vocabulary::fromStream  (signature 55824124, params struct, block data)
 *   /root/reference/3rdparty/fbow/fbow/fbow.h:125-133      params layout (120 bytes)
 *   /root/reference/3rdparty/fbow/fbow/fbow.h:137-194      block layout: u16 N | u16 isLeaf | u32 parent | features | (id_or_child, weight)[]
 *   /root/reference/3rdparty/fbow/fbow/fbow.h:402-448      _transform2<L1_32bytes>: per descriptor descend by first-minimum Hamming
 *   /root/reference/3rdparty/fbow/fbow/fbow.h:343-350      L1_32bytes distance (4 x popcount64)
 * Output is per DESCRIPTOR (word id, weight, level-node id); folding into fBow (map<word, sum of weights in descriptor order>)
 * and fBow2 (map<node, descriptor indices>) is a host loop in descriptor order, as in the reference.
 * Parity pin: checked against the reference's own fbow compiled from /root/reference (oracle/_ref/libref_fbow.so) on the
 * shipped vocabulary 3rdparty/vocabularies/orb.fbow and on synthetic vocabularies; golden vectors in tests/golden/bow_*.npz.
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

typedef struct {
    uint32_t alignment, nblocks;
    uint64_t desc_size_wp, block_size, feature_off, child_off, total_size;
    int32_t desc_type, desc_size;
    uint32_t k;
    const uint8_t* data;
} voc_t;

/* returns 0 on success */
int oracle_bow_parse(const uint8_t* bytes, size_t n, voc_t* v) {
    uint64_t sig;
    if (n < 128) return -1;
    memcpy(&sig, bytes, 8);
    if (sig != 55824124ull) return -2;
    const uint8_t* p = bytes + 8;
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
    if (n < 128 + v->total_size) return -3;
    v->data = bytes + 128;
    return 0;
}

static uint32_t dist32(const uint8_t* a, const uint8_t* b) {
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
           __builtin_popcountll(x[3] ^ y[3]);
}

/* word[i] / node[i] = 0xFFFFFFFF when the descriptor contributes no word / no level-node entry */
int oracle_bow_transform(const uint8_t* voc_bytes, size_t voc_n, const uint8_t* desc, int n, size_t stride, int store_level,
                         uint32_t* word, float* weight, uint32_t* node) {
    voc_t v;
    int rc = oracle_bow_parse(voc_bytes, voc_n, &v);
    if (rc) return rc;
    if (v.desc_type != 0 || v.desc_size != 32) return -4;
    const int nbits = (int)ceil(log2((double)v.k));
    uint32_t best_idx = 0; /* persists across iterations exactly like best_dist_idx.second */
    for (int f = 0; f < n; f++) {
        const uint8_t* feat = desc + (size_t)f * stride;
        uint32_t block = 0, level = 0, cur_node = 0;
        word[f] = 0xFFFFFFFFu; weight[f] = 0.f; node[f] = 0xFFFFFFFFu;
        for (;;) {
            const uint8_t* b = v.data + (uint64_t)block * v.block_size;
            uint16_t N; memcpy(&N, b, 2);
            uint64_t best = 0xFFFFFFFFull;
            for (int c = 0; c < N; c++) {
                uint64_t d = dist32(feat, b + v.feature_off + (uint64_t)c * v.desc_size_wp);
                if (d < best) { best = d; best_idx = (uint32_t)c; }
            }
            if (level == (uint32_t)store_level) node[f] = cur_node;
            uint32_t id; float w;
            memcpy(&id, b + v.child_off + 8ull * best_idx, 4);
            memcpy(&w, b + v.child_off + 8ull * best_idx + 4, 4);
            if (id & 0x80000000u) {
                word[f] = id & 0x7FFFFFFFu; weight[f] = w;
                if (level < (uint32_t)store_level) node[f] = cur_node;
                break;
            }
            block = id & 0x7FFFFFFFu;
            cur_node = (cur_node << nbits) | best_idx;
            level++;
            if (block == 0) break; /* while( !isleaf && getId()!=0 ) */
        }
    }
    return 0;
}

/* fBow::score, fbow.cpp:192-243 on sorted (id, weight) lists */
double oracle_bow_score(const uint32_t* id1, const float* w1, int n1, const uint32_t* id2, const float* w2, int n2) {
    int i = 0, j = 0; double score = 0;
    while (i < n1 && j < n2) {
        if (id1[i] == id2[j]) { score += w1[i] * w2[j]; i++; j++; }
        else if (id1[i] < id2[j]) { while (i < n1 && id1[i] < id2[j]) i++; }
        else { while (j < n2 && id2[j] < id1[i]) j++; }
    }
    if (score >= 1) score = 1.0; else score = 1.0 - sqrt(1.0 - score);
    return score;
}
