// stl_helper.cpp — TEST INFRASTRUCTURE ONLY.  Exposes the host C++ runtime's own algorithms to the Python ORB oracle so that
// the oracle uses the REAL library behaviour the reference gets, not a restatement:
//   * cv::KeyPointsFilter::retainBest (OpenCV features2d, called at /root/reference/src/featureextractors/ORBextractor.cpp:1053,1071)
//     = std::nth_element(begin, begin+n-1, end, response-greater) ; std::partition(begin+n, end, response >= boundary) ; resize
//     followed by the reference's own truncation to n.  libstdc++'s introselect decides which tied keypoints survive and
//     in which order, so it is called here directly.
//   * cosf / sinf of the platform libm (ORBextractor.cpp:119).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

struct KP { float x, y, size, angle, response; int octave, class_id; };  // cv::KeyPoint layout, 28 bytes

extern "C" {
// returns the new element count; kps is permuted in place exactly as the library does
int stl_retain_best(KP* kps, int count, int n_points) {
    if (n_points >= 0 && count > n_points) {
        if (n_points == 0) return 0;
        std::nth_element(kps, kps + n_points - 1, kps + count,
                         [](const KP& a, const KP& b) { return a.response > b.response; });
        float ambiguous = kps[n_points - 1].response;
        KP* new_end = std::partition(kps + n_points, kps + count, [ambiguous](const KP& k) { return k.response >= ambiguous; });
        return (int)(new_end - kps);
    }
    return count;
}
// std::sort of indices by keys[idx] with the host C++ runtime's own algorithm (what picoflann's kd-tree build calls for degenerate
// cuts, /root/reference/src/basictypes/picoflann.h:310-318): the unstable introsort decides the order of equal keys
void stl_sort_indices(uint32_t* idx, int n, const float* keys) {
    std::sort(idx, idx + n, [keys](const uint32_t& a, const uint32_t& b) { return keys[a] < keys[b]; });
}
float stl_cosf(float x) { return std::cos(x); }
float stl_sinf(float x) { return std::sin(x); }
void stl_sincosf_array(const float* x, int n, float* c, float* s) {
    for (int i = 0; i < n; i++) { c[i] = std::cos(x[i]); s[i] = std::sin(x[i]); }
}
}
