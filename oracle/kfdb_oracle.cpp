// kfdb_oracle.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).  CPU restatement of the reference's keyframe
// database query on flat arrays, container for container, so that iteration orders (std::map / std::set) and the libstdc++
// std::sort tie behaviour are the reference's:
//   src/map_types/keyframedatabase.cpp:150-160   add: word_frames_[word].insert(frame)
//   src/map_types/keyframedatabase.cpp:195-276   relocalizationCandidates (votes, 0.8*max gate, fBow::score, covisibility sum, 0.75*best)
//   3rdparty/fbow/fbow/fbow.cpp:192-243          fBow::score
//   src/map_types/covisgraph.cpp:167-182         getNeighborsWeights(idx, sorted=true)
// Pinned against the reference's own keyframedatabase.cpp + covisgraph.cpp + fbow compiled unchanged (oracle/_ref/libref_kfdb.so,
// oracle/ref_kfdb_wrap.cpp) in tests/test_kfdb_oracle.py; goldens tests/golden/kfdb_ref.npz.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <set>
#include <utility>
#include <vector>

namespace {
typedef std::map<uint32_t, float> Bow;

double bow_score(const Bow& v1, const Bow& v2) {   // fbow.cpp:192-243
    Bow::const_iterator a = v1.begin(), b = v2.begin();
    double score = 0;
    while (a != v1.end() && b != v2.end()) {
        if (a->first == b->first) {
            score += a->second * b->second;   // float product, double accumulation
            ++a;
            ++b;
        } else if (a->first < b->first) {
            while (a != v1.end() && a->first < b->first) ++a;
        } else {
            while (b != v2.end() && b->first < a->first) ++b;
        }
    }
    if (score >= 1) return 1.0;
    return 1.0 - sqrt(1.0 - score);
}
}  // namespace

extern "C" {

double oracle_bow_score(const uint32_t* ids1, const float* w1, int n1, const uint32_t* ids2, const float* w2, int n2) {
    Bow a, b;
    for (int i = 0; i < n1; i++) a[ids1[i]] = w1[i];
    for (int i = 0; i < n2; i++) b[ids2[i]] = w2[i];
    return bow_score(a, b);
}

// Database: n_frames frames, frame f owns words[off[f]..off[f+1]) (any order; folded into a map like fBow).
// Covisibility graph: undirected edges (edge_a, edge_b) with final weights edge_w.
// Outputs: scored frames (the reference's frame_score map, ascending id) and the returned candidate list.
int oracle_kfdb_candidates(int n_frames, const uint32_t* frame_ids, const int64_t* off, const uint32_t* words, const float* weights,
                           const uint32_t* q_words, const float* q_weights, int nq, const uint32_t* excluded, int n_excluded,
                           float minScore, int sorted, int n_edges, const uint32_t* edge_a, const uint32_t* edge_b,
                           const float* edge_w, uint32_t* scored_frame, double* scored_score, uint32_t* scored_common, int* n_scored,
                           uint32_t* max_common, uint32_t* cand, int* n_cand) {
    std::map<uint32_t, std::set<uint32_t>> word_frames;   // keyframedatabase.cpp:157
    std::map<uint32_t, Bow> bows;
    for (int f = 0; f < n_frames; f++) {
        Bow& b = bows[frame_ids[f]];
        for (int64_t i = off[f]; i < off[f + 1]; i++) b[words[i]] = weights[i];
        for (auto& w : b) word_frames[w.first].insert(frame_ids[f]);
    }
    Bow q;
    for (int i = 0; i < nq; i++) q[q_words[i]] = q_weights[i];
    std::set<uint32_t> excludedFrames(excluded, excluded + n_excluded);
    std::map<uint32_t, std::set<uint32_t>> graph;          // covisgraph: _mgraph
    std::map<std::pair<uint32_t, uint32_t>, float> gw;
    for (int e = 0; e < n_edges; e++) {
        graph[edge_a[e]].insert(edge_b[e]);
        graph[edge_b[e]].insert(edge_a[e]);
        gw[std::make_pair(std::min(edge_a[e], edge_b[e]), std::max(edge_a[e], edge_b[e]))] = edge_w[e];
    }
    *n_scored = 0;
    *n_cand = 0;
    *max_common = 0;

    std::map<uint32_t, uint32_t> frame_nobs;               // :203-217
    uint32_t maxCommonWords = 0;
    for (auto& w : q) {
        auto it = word_frames.find(w.first);
        if (it == word_frames.end()) continue;
        for (auto f : it->second) {
            if (excludedFrames.count(f)) continue;
            uint32_t& n = frame_nobs[f];
            n++;
            if (n > maxCommonWords) maxCommonWords = n;
        }
    }
    *max_common = maxCommonWords;
    if (frame_nobs.empty()) return 0;                      // :221
    uint32_t minCommonWords = maxCommonWords * 0.8f;       // :222
    std::map<uint32_t, double> frame_score;
    for (auto& fn : frame_nobs) {
        if (fn.second > minCommonWords) {
            double si = bow_score(q, bows[fn.first]);
            if (si > minScore) {
                frame_score[fn.first] = si;
                scored_frame[*n_scored] = fn.first;
                scored_score[*n_scored] = si;
                scored_common[*n_scored] = fn.second;
                (*n_scored)++;
            }
        }
    }
    if (frame_score.empty()) return 0;                     // :236
    if (frame_score.size() == 1) {                         // :237
        cand[0] = frame_score.begin()->first;
        *n_cand = 1;
        return 0;
    }
    std::vector<std::pair<uint32_t, double>> frame_scoreCovis;
    double bestAccScore = minScore;
    for (auto& fs : frame_score) {
        double accScore = fs.second;
        std::vector<std::pair<uint32_t, float>> n_weight;  // covisgraph.cpp:167-182
        auto g = graph.find(fs.first);
        if (g != graph.end()) {
            for (auto n : g->second)
                n_weight.push_back(std::make_pair(n, gw[std::make_pair(std::min(fs.first, n), std::max(fs.first, n))]));
            std::sort(n_weight.begin(), n_weight.end(),
                      [](const std::pair<uint32_t, float>& a, const std::pair<uint32_t, float>& b) { return a.second > b.second; });
        }
        n_weight.resize(std::min(n_weight.size(), size_t(10)));
        for (auto& nw : n_weight) {
            auto it = frame_score.find(nw.first);
            if (it != frame_score.end()) accScore += it->second;
        }
        frame_scoreCovis.push_back(std::make_pair(fs.first, accScore));
        if (accScore > bestAccScore) bestAccScore = accScore;
    }
    double minScoreToRetain = 0.75f * bestAccScore;        // :262
    frame_scoreCovis.erase(std::remove_if(frame_scoreCovis.begin(), frame_scoreCovis.end(),
                                          [&](const std::pair<uint32_t, double>& v) { return v.second < minScoreToRetain; }),
                           frame_scoreCovis.end());
    if (sorted)
        std::sort(frame_scoreCovis.begin(), frame_scoreCovis.end(),
                  [&](const std::pair<uint32_t, double>& a, const std::pair<uint32_t, double>& b) { return a.second > b.second; });
    for (auto& fs : frame_scoreCovis) cand[(*n_cand)++] = fs.first;
    return 0;
}

// CovisGraph::getNeighborsWeights(idx, true) restated for the adapters' tests: neighbour ids in decreasing weight
int oracle_covis_neighbors(int n_edges, const uint32_t* edge_a, const uint32_t* edge_b, const float* edge_w, uint32_t idx,
                           uint32_t* out, int cap) {
    std::set<uint32_t> nb;
    std::map<std::pair<uint32_t, uint32_t>, float> gw;
    for (int e = 0; e < n_edges; e++) {
        if (edge_a[e] == idx) nb.insert(edge_b[e]);
        if (edge_b[e] == idx) nb.insert(edge_a[e]);
        gw[std::make_pair(std::min(edge_a[e], edge_b[e]), std::max(edge_a[e], edge_b[e]))] = edge_w[e];
    }
    std::vector<std::pair<uint32_t, float>> n_weight;
    for (auto n : nb) n_weight.push_back(std::make_pair(n, gw[std::make_pair(std::min(idx, n), std::max(idx, n))]));
    std::sort(n_weight.begin(), n_weight.end(),
              [](const std::pair<uint32_t, float>& a, const std::pair<uint32_t, float>& b) { return a.second > b.second; });
    int k = 0;
    for (auto& nw : n_weight)
        if (k < cap) out[k++] = nw.first;
    return (int)n_weight.size();
}
}
