// ref_match_wrap.cpp — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own frame matcher: src/utils/framematcher.cpp
// is compiled UNCHANGED, where it lies under /root/reference (oracle/Makefile; output to oracle/_ref/), against container stand-ins
// for OpenCV / Frame (oracle/shim2), together with the reference's own match filters (src/basictypes/misc.cpp:105-185, cut out by
// oracle/gen_ref_extract.py).  Used to PIN oracle/match_oracle.c and the CUDA matcher (tests/golden/match_ref.npz).
//
// One preprocessor substitution, stated here because it decides what is being pinned: FrameMatcher_Flann::setParams builds an
// xflann HKMeans(32,0) index that is searched with 16 checks (framematcher.cpp:213,239) — an APPROXIMATE 10-NN whose misses depend
// on a seeded shuffle.  The product searches exactly, so to compare the POST-FILTERS (:246-322) on identical candidates
// the file is compiled with that one constructor call swapped for xflann's own exact index, `LinearParams()`; xflann.h is included
// first so that only the call site inside framematcher.cpp sees the macro.  FrameMatcher_BoW (:407-541) uses no index and is
// compiled as is.
#include <xflann/xflann.h>
#include <fbow/fbow.h>
#define HKMeansParams(a, b) LinearParams()
#include <utils/framematcher.cpp>
#undef HKMeansParams

namespace ucoslam {
#include "gen/misc_filters.inc"
static cv::Mat g_F12;   // the fundamental matrix is an INPUT of the stage under test (computeF12, misc.cpp:893-920, is OpenCV matrix algebra)
cv::Mat computeF12(const cv::Mat&, const cv::Mat&, const cv::Mat&, const cv::Mat&) { return g_F12.clone(); }
}  // namespace ucoslam

namespace {
struct KP { float x, y, size, angle, response; int octave, class_id; };
void fill_frame(ucoslam::Frame& f, int n, const KP* kps, const unsigned char* desc, const uint32_t* ids, const unsigned char* flags,
                const float* scale_factors, int n_scales, int n_nodes, const uint32_t* node_id, const int32_t* node_ptr, const int32_t* node_kp) {
    f.und_kpts.resize(n);
    memcpy((void*)f.und_kpts.data(), kps, sizeof(KP) * (size_t)n);
    f.desc = cv::Mat(n, 32, CV_8UC1);
    if (n) memcpy(f.desc.ptr<uchar>(0), desc, 32 * (size_t)n);
    f.ids.assign(n, std::numeric_limits<uint32_t>::max());
    if (ids) f.ids.assign(ids, ids + n);
    f.flags.assign(n, Flag());
    for (int i = 0; flags && i < n; i++) f.flags[i].v = flags[i];
    f.scaleFactors.assign(scale_factors, scale_factors + n_scales);
    f.imageParams.CameraMatrix = cv::Mat::eye(3, 3, CV_32F);
    f.bowvector_level = std::make_shared<fbow::fBow2>();
    for (int k = 0; k < n_nodes; k++) {
        std::vector<uint32_t>& v = (*f.bowvector_level)[node_id[k]];
        for (int e = node_ptr[k]; e < node_ptr[k + 1]; e++) v.push_back((uint32_t)node_kp[e]);
    }
}
}  // namespace

extern "C" {
// type: 1 = FrameMatcher::TYPE_FLANN, 2 = TYPE_BOW; modes: FrameMatcher::Mode values; f12: 9 floats or NULL (match / matchEpipolar with
// an empty FQ2T).  *_node_*: the frames' fBow2 flattened in std::map order (BoW matcher only).  Returns the number of matches.
int ref_frame_match(int type, int nt, const KP* t_kps, const unsigned char* t_desc, const uint32_t* t_ids, const unsigned char* t_flags, int t_mode,
                    int t_nodes, const uint32_t* t_node_id, const int32_t* t_node_ptr, const int32_t* t_node_kp,
                    int nq, const KP* q_kps, const unsigned char* q_desc, const uint32_t* q_ids, const unsigned char* q_flags, int q_mode,
                    int q_nodes, const uint32_t* q_node_id, const int32_t* q_node_ptr, const int32_t* q_node_kp,
                    const float* scale_factors, int n_scales, float min_desc_dist, float ratio, int check_orientation, int max_octave_diff,
                    const float* f12, cv::DMatch* out, int cap) {
    try {
        ucoslam::Frame T, Q;
        fill_frame(T, nt, t_kps, t_desc, t_ids, t_flags, scale_factors, n_scales, t_nodes, t_node_id, t_node_ptr, t_node_kp);
        fill_frame(Q, nq, q_kps, q_desc, q_ids, q_flags, scale_factors, n_scales, q_nodes, q_node_id, q_node_ptr, q_node_kp);
        ucoslam::FrameMatcher fm((ucoslam::FrameMatcher::Type)type);
        fm.setParams(T, (ucoslam::FrameMatcher::Mode)t_mode, min_desc_dist, ratio, check_orientation != 0, max_octave_diff);
        cv::Mat FQ2T;
        if (f12) {
            ucoslam::g_F12 = cv::Mat(3, 3, CV_32F);
            memcpy(ucoslam::g_F12.ptr<float>(0), f12, 36);
            FQ2T = cv::Mat::eye(4, 4, CV_32F);
        }
        std::vector<cv::DMatch> m = f12 ? fm.matchEpipolar(Q, (ucoslam::FrameMatcher::Mode)q_mode, FQ2T) : fm.match(Q, (ucoslam::FrameMatcher::Mode)q_mode);
        if ((int)m.size() > cap) return -2;
        for (size_t i = 0; i < m.size(); i++) out[i] = m[i];
        return (int)m.size();
    } catch (std::exception& e) {
        fprintf(stderr, "ref_frame_match: %s\n", e.what());
        return -1;
    }
}
// the reference's filter_ambiguous_query / filter_ambiguous_train on a match list, in place; returns the new count
int ref_filter_ambiguous(cv::DMatch* m, int n, int train) {
    std::vector<cv::DMatch> v(m, m + n);
    if (train) ucoslam::filter_ambiguous_train(v); else ucoslam::filter_ambiguous_query(v);
    for (size_t i = 0; i < v.size(); i++) m[i] = v[i];
    return (int)v.size();
}
}
