// frame_matcher_b200.h — the reference's descriptor-matcher seam: an _impl::FrameMatcher_impl
// (src/utils/framematcher.cpp:31-58: setParams(trainFrame, mode, minDescDist, ratio, checkOrientation, maxOctaveDiff), match,
// matchEpipolar) whose work is uco_b200_frame_match: exact 10-NN + the post-filters of FrameMatcher_Flann::matchEpipolar
// (:228-322) on the device.  The class is declared inside framematcher.cpp in the reference; paste it next to FrameMatcher_Flann
// and select it in FrameMatcher::FrameMatcher(Type) (:110-125) with a new Type value.  setParams only records the train frame (there
// is no index to build); match* are re-entrant per object as long as each calling thread owns its own instance, which is how the
// OpenMP callers use FrameMatcher (src/utils/mapmanager.cpp:9992-10065, src/utils/system.cpp:5026-5078).
// Compiled and driven next to the reference's own code by tests/adapters/adapter_world_test.cpp (oracle/shim2 stand-ins).
#pragma once
#include <vector>
#include "uco_b200_cxx.h"

namespace ucoslam { namespace _impl {

class FrameMatcher_B200 : public FrameMatcher_impl {
public:
    explicit FrameMatcher_B200(int device = 0) : _ctx(device) {}
    void setParams(const Frame& trainFrame, FrameMatcher::Mode mode, float minDescDist, float nn_match_ratio, bool checkOrientation,
                   int maxOctaveDiff) override {
        _train = &trainFrame;
        _trainImageParams = trainFrame.imageParams;        // as FrameMatcher_Flann::setParams: getFund12 reads it
        _prm = uco_match_params{};
        _prm.min_desc_dist = minDescDist; _prm.nn_match_ratio = nn_match_ratio;
        _prm.check_orientation = checkOrientation; _prm.max_octave_diff = maxOctaveDiff;
        rows(trainFrame, mode, _tRows, _tDesc);
    }
    std::vector<cv::DMatch> match(const Frame& queryFrame, FrameMatcher::Mode mode) override { return run(queryFrame, mode, cv::Mat()); }
    // FQ2T is the SE3 matrix train -> query; the fundamental matrix comes from the reference's own getFund12 / computeF12
    // (framematcher.cpp:58-64, misc.cpp:893-920), exactly as in FrameMatcher_Flann::matchEpipolar (:234)
    std::vector<cv::DMatch> matchEpipolar(const Frame& queryFrame, FrameMatcher::Mode mode, const cv::Mat& FQ2T) override {
        return run(queryFrame, mode, getFund12(_trainImageParams.CameraMatrix, queryFrame.imageParams.CameraMatrix, FQ2T));
    }

private:
    // FrameMatcher::manageMode (:160-198): the descriptor rows of the requested mode and the keypoint each one belongs to
    static void rows(const Frame& f, FrameMatcher::Mode mode, std::vector<int32_t>& map, std::vector<uint8_t>& desc) {
        map.clear();
        for (size_t i = 0; i < f.ids.size(); i++) {
            const bool assigned = f.ids[i] != std::numeric_limits<uint32_t>::max();
            // MODE_ALL takes every keypoint (:168-173); the two selective modes also drop FLAG_NONMAXIMA keypoints (:174-196)
            if (mode == FrameMatcher::MODE_ALL ||
                (!f.flags[i].is(Frame::FLAG_NONMAXIMA) && ((mode == FrameMatcher::MODE_ASSIGNED && assigned) || (mode == FrameMatcher::MODE_UNASSIGNED && !assigned))))
                map.push_back((int32_t)i);
        }
        desc.resize(32 * map.size());
        for (size_t r = 0; r < map.size(); r++) memcpy(&desc[32 * r], f.desc.ptr<uchar>(map[r]), 32);
    }
    std::vector<cv::DMatch> run(const Frame& q, FrameMatcher::Mode mode, const cv::Mat& F) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(uco_keypoint) && sizeof(cv::DMatch) == sizeof(uco_match), "layouts");
        std::vector<int32_t> qRows;
        std::vector<uint8_t> qDesc;
        rows(q, mode, qRows, qDesc);
        uco_match_params prm = _prm;
        prm.use_f12 = !F.empty();
        if (prm.use_f12) { cv::Mat f32; F.convertTo(f32, CV_32F); memcpy(prm.f12, f32.ptr<float>(0), 36); }
        prm.n_scales = (int)q.scaleFactors.size();
        for (int i = 0; i < prm.n_scales && i < UCO_MATCH_MAX_SCALES; i++) prm.scale_factors[i] = q.scaleFactors[i];
        std::vector<cv::DMatch> out(qRows.size());
        int n = 0;
        _ctx.check(uco_b200_frame_match(_ctx.get(), qDesc.data(), (int)qRows.size(), 32, reinterpret_cast<const uco_keypoint*>(q.und_kpts.data()),
                                        (int)q.und_kpts.size(), qRows.data(), _tDesc.data(), (int)_tRows.size(), 32,
                                        reinterpret_cast<const uco_keypoint*>(_train->und_kpts.data()), (int)_train->und_kpts.size(),
                                        _tRows.data(), &prm, reinterpret_cast<uco_match*>(out.data()), (int)out.size(), &n));
        out.resize(n);
        return out;
    }
public:
    // FrameMatcher_BoW::matchEpipolar (:407-541) on the same object: candidates from the frames' bowvector_level (fBow2)
    std::vector<cv::DMatch> matchEpipolarBoW(const Frame& q, FrameMatcher::Mode qMode, FrameMatcher::Mode tMode, const cv::Mat& F12) {
        struct Flat { std::vector<uint32_t> id; std::vector<int32_t> ptr, kp; std::vector<uint8_t> usable; };
        auto flatten = [](const Frame& f, FrameMatcher::Mode mode) {
            Flat o;
            o.ptr.push_back(0);
            for (const auto& kv : *f.bowvector_level) {            // std::map: ascending node id
                o.id.push_back(kv.first);
                for (auto k : kv.second) o.kp.push_back((int32_t)k);
                o.ptr.push_back((int32_t)o.kp.size());
            }
            o.usable.resize(f.und_kpts.size());
            for (size_t i = 0; i < f.und_kpts.size(); i++) {       // isUsed(frame, idx, mode), framematcher.cpp:543-556
                const bool assigned = f.ids[i] != std::numeric_limits<uint32_t>::max();
                o.usable[i] = !f.flags[i].is(Frame::FLAG_NONMAXIMA) &&
                              (mode == FrameMatcher::MODE_ALL || (mode == FrameMatcher::MODE_ASSIGNED) == assigned);
            }
            return o;
        };
        const Flat fq = flatten(q, qMode), ft = flatten(*_train, tMode);
        uco_bow_index qb{(int32_t)fq.id.size(), fq.id.data(), fq.ptr.data(), fq.kp.data()};
        uco_bow_index tb{(int32_t)ft.id.size(), ft.id.data(), ft.ptr.data(), ft.kp.data()};
        uco_match_params prm = _prm;
        prm.use_f12 = !F12.empty();
        if (prm.use_f12) { cv::Mat f32; F12.convertTo(f32, CV_32F); memcpy(prm.f12, f32.ptr<float>(0), 36); }
        prm.n_scales = (int)q.scaleFactors.size();
        for (int i = 0; i < prm.n_scales && i < UCO_MATCH_MAX_SCALES; i++) prm.scale_factors[i] = q.scaleFactors[i];
        std::vector<cv::DMatch> out(fq.kp.size());
        int n = 0;
        _ctx.check(uco_b200_frame_match_bow(_ctx.get(), q.desc.ptr<uchar>(0), q.desc.step[0], reinterpret_cast<const uco_keypoint*>(q.und_kpts.data()),
                                            (int)q.und_kpts.size(), fq.usable.data(), &qb, _train->desc.ptr<uchar>(0), _train->desc.step[0],
                                            reinterpret_cast<const uco_keypoint*>(_train->und_kpts.data()), (int)_train->und_kpts.size(),
                                            ft.usable.data(), &tb, &prm, reinterpret_cast<uco_match*>(out.data()), (int)out.size(), &n));
        out.resize(n);
        return out;
    }

private:
    uco_b200::Context _ctx;
    const Frame* _train = nullptr;
    uco_match_params _prm{};
    std::vector<int32_t> _tRows;
    std::vector<uint8_t> _tDesc;
};

}}  // namespace ucoslam::_impl
