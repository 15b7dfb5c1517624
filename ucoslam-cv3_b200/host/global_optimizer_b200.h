// global_optimizer_b200.h — the reference's bundle-adjustment seam: a ucoslam::GlobalOptimizer
// (src/optimization/globaloptimizer.h:28-68) whose optimize() is uco_b200_ba_solve.  setParams walks the Map with the rules
// of GlobalOptimizerG2O::setParams (src/optimization/globaloptimizer_g2o.cpp:98-176: used / fixed frames, points with >= 2
// observations or stereo and not bad, frames that only observe a used point enter as fixed) and flattens it into
// uco_ba_problem; getResults writes poses, points and the bad-association list back as :466-538 does.  The object keeps
// its own copy of inputs and results, because the mapper thread calls setParams + optimize WITHOUT the map lock and the
// tracker thread calls getResults later (src/utils/mapmanager.cpp:11361-11405, :1267-1305); *stopASAP is forwarded.
// ArUco markers travel too (marker vertices, MarkerEdges, the per-keyframe marker weights of :276-297), so do windows whose keyframes
// were taken with different cameras (one fx fy cx cy bf row per keyframe: uco_ba_problem::pose_cam) and the InPlaneMarkers option
// (:356-401: reference marker choice, planar edges, their weight).  There is no CPU fallback: errors of the C ABI become
// std::runtime_error, like every other error path of the reference's plugins.
// Selected with Params::global_optimizer = "b200" once registered in GlobalOptimizer::create.
// Compiled and driven against the reference's own globaloptimizer.h by tests/adapters/adapter_world_test.cpp (container stand-ins for
// Map / Frame / OpenCV, oracle/shim2); inside the reference tree it compiles against the real headers.
#pragma once
#include <algorithm>
#include <cstring>
#include <limits>
#include <map>
#include <vector>
#include "optimization/globaloptimizer.h"
#include "uco_b200_cxx.h"

namespace ucoslam {

class GlobalOptimizerB200 : public GlobalOptimizer {
public:
    explicit GlobalOptimizerB200(int device = 0) : _ctx(device) {}
    string getName() const override { return "b200"; }

    void setParams(std::shared_ptr<Map> map, const ParamSet& ps) override {
        _params = ps;
        const uint32_t INVALID = std::numeric_limits<uint32_t>::max();
        std::vector<uint32_t> frameSlot(map->keyframes.capacity(), INVALID), pointSlot(map->map_points.capacity(), INVALID);
        std::vector<char> fixedKind(map->keyframes.capacity(), 0);  // 0 free, 1 fixed without points, 2 fixed with points
        _frameIds.clear(); _pointIds.clear();
        auto useFrame = [&](uint32_t f, char kind) {
            if (frameSlot[f] == INVALID) { frameSlot[f] = (uint32_t)_frameIds.size(); _frameIds.push_back(f); fixedKind[f] = kind; }
        };
        if (_params.used_frames.empty()) for (auto& f : map->keyframes) useFrame(f.idx, 0);
        else for (auto f : _params.used_frames) useFrame(f, 0);
        if (_params.fixFirstFrame && frameSlot[map->keyframes.front().idx] != INVALID) fixedKind[map->keyframes.front().idx] = 2;
        for (auto f : _params.fixed_frames) if (frameSlot[f] != INVALID) fixedKind[f] = 2;
        bool markers = false, mixedCameras = false;
        std::map<uint32_t, uint32_t> markerSlot;
        _markerIds.clear();
        const size_t nInitial = _frameIds.size();
        for (size_t k = 0; k < nInitial; k++) {           // :135-176 (frames added as observers are not walked for points)
            const uint32_t f = _frameIds[k];
            for (auto pid : map->keyframes[f].ids) {
                if (pid == INVALID || pointSlot[pid] != INVALID) continue;
                MapPoint& mp = map->map_points[pid];
                if ((mp.frames.size() < 2 && !mp.isStereo()) || mp.isBad()) { pointSlot[pid] = INVALID - 1; continue; }
                pointSlot[pid] = (uint32_t)_pointIds.size();
                _pointIds.push_back(pid);
                for (const auto& fi : mp.frames) useFrame(fi.first, 1);
            }
            for (auto& m : map->keyframes[f].markers) {    // :157-170: valid map markers, all their frames join as fixed observers
                if (markerSlot.count(m.id) || !map->map_markers[m.id].pose_g2m.isValid()) continue;
                markerSlot[m.id] = (uint32_t)_markerIds.size();
                _markerIds.push_back(m.id);
                for (auto fid : map->map_markers[m.id].frames) useFrame(fid, 1);
                markers = true;
            }
        }
        const Frame& f0 = map->keyframes[_frameIds.front()];
        for (auto f : _frameIds) {
            const ImageParams& ip = map->keyframes[f].imageParams;
            if (ip.fx() != f0.imageParams.fx() || ip.fy() != f0.imageParams.fy() || ip.cx() != f0.imageParams.cx() ||
                ip.cy() != f0.imageParams.cy() || ip.bl != f0.imageParams.bl) mixedCameras = true;
        }
        // the reference emits marker vertices / edges in ascending marker id (std::map, globaloptimizer_g2o.cpp:306-348)
        std::sort(_markerIds.begin(), _markerIds.end());
        for (size_t i = 0; i < _markerIds.size(); i++) markerSlot[_markerIds[i]] = (uint32_t)i;
        // ---- flatten (own copy: the map may change before optimize()/getResults())
        const size_t P = _frameIds.size(), N = _pointIds.size();
        _poses.resize(16 * P); _fixed.resize(P); _points.resize(3 * N);
        _obsPose.clear(); _obsPoint.clear(); _obsUV.clear(); _obsUR.clear(); _obsStereo.clear(); _obsInv.clear();
        std::vector<float> invScale;
        for (auto s : map->keyframes.front().scaleFactors) invScale.push_back(1. / s);   // :95-96
        for (size_t k = 0; k < P; k++) {
            cv::Mat T = map->keyframes[_frameIds[k]].pose_f2g;
            for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) _poses[16 * k + 4 * r + c] = T.at<float>(r, c);
            _fixed[k] = fixedKind[_frameIds[k]] != 0;
        }
        for (size_t j = 0; j < N; j++) {
            MapPoint& mp = map->map_points[_pointIds[j]];
            const cv::Point3f p = mp.getCoordinates();
            _points[3 * j] = p.x; _points[3 * j + 1] = p.y; _points[3 * j + 2] = p.z;
            for (const auto& fi : mp.frames) {             // :228-276
                if (frameSlot[fi.first] == INVALID) continue;
                const Frame& fr = map->keyframes[fi.first];
                const cv::KeyPoint& kp = fr.und_kpts[fi.second];
                const float depth = fr.getDepth(fi.second);
                _obsPose.push_back((int32_t)frameSlot[fi.first]);
                _obsPoint.push_back((int32_t)j);
                _obsUV.push_back(kp.pt.x); _obsUV.push_back(kp.pt.y);
                const float mbf = fr.imageParams.bl * fr.imageParams.fx();
                _obsStereo.push_back(depth > 0);
                _obsUR.push_back(depth > 0 ? kp.pt.x - mbf / depth : 0.f);
                _obsInv.push_back(invScale[kp.octave]);
            }
        }
        // markers (:304-350) and the per-keyframe weight of their 8 residuals (:276-297)
        _mkPose.clear(); _mkSize.clear(); _moMarker.clear(); _moPose.clear(); _moCorners.clear(); _moWeight.clear();
        double totalMarkerWeight = 0;
        if (markers) {
            std::vector<double> kpw(P, 0.0);
            for (size_t o = 0; o < _obsPose.size(); o++) kpw[_obsPose[o]] += (_obsStereo[o] ? 3 : 2) * _obsInv[o];
            for (auto mid : _markerIds) {
                const Marker& mk = map->map_markers[mid];
                const float* g = mk.pose_g2m.ptr<float>(0);
                _mkPose.insert(_mkPose.end(), g, g + 16);
                _mkSize.push_back(mk.size);
                for (auto fid : mk.frames) {
                    const Frame& fr = map->keyframes[fid];
                    const uint32_t slot = frameSlot[fid];
                    double w = 1;
                    if (kpw[slot] > 40 && fr.markers.size() > 0)
                        w = _params.markersOptWeight * std::min(1., double(fr.markers.size()) / _params.minMarkersForMaxWeight) * kpw[slot] /
                            double(fr.markers.size() * 8);
                    _moMarker.push_back((int32_t)markerSlot[mid]);
                    _moPose.push_back((int32_t)slot);
                    for (const auto& c : fr.getMarker(mid).und_corners) { _moCorners.push_back(c.x); _moCorners.push_back(c.y); }
                    _moWeight.push_back((float)w);
                    totalMarkerWeight += w * 8;
                }
            }
        }
        // InPlaneMarkers (:356-401): the valid map marker seen by most keyframes is the reference (first one in id order on a tie); every
        // OTHER marker vertex of the window gets one planar edge; a reference outside the window enters as a fixed vertex
        _planeOther.clear(); _planeRef44.clear();
        int planeRef = -1;
        double planeWeight = 0;
        {
            int nValid = 0;
            for (const auto& m : map->map_markers) if (m.second.pose_g2m.isValid()) nValid++;
            if (_params.InPlaneMarkers && nValid >= 2) {
                std::pair<uint32_t, uint32_t> best(INVALID, 0);
                for (const auto& m : map->map_markers)
                    if (m.second.pose_g2m.isValid() && m.second.frames.size() > best.second) best = {m.first, (uint32_t)m.second.frames.size()};
                if (best.first != INVALID) {
                    size_t nInfo = _markerIds.size();
                    if (markerSlot.count(best.first) && std::binary_search(_markerIds.begin(), _markerIds.end(), best.first)) planeRef = (int)markerSlot[best.first];
                    else {
                        nInfo++;
                        const float* g = map->map_markers[best.first].pose_g2m.ptr<float>(0);
                        _planeRef44.assign(g, g + 16);
                    }
                    for (size_t i = 0; i < _markerIds.size(); i++)
                        if (_markerIds[i] != best.first) _planeOther.push_back((int32_t)i);
                    double totalKpWeight = 0;
                    {
                        std::vector<double> kpw(P, 0.0);
                        for (size_t o = 0; o < _obsPose.size(); o++) kpw[_obsPose[o]] += (_obsStereo[o] ? 3 : 2) * _obsInv[o];
                        for (double v : kpw) totalKpWeight += v;
                    }
                    if (nInfo > 1) planeWeight = 0.33 * (totalMarkerWeight + totalKpWeight) / double(4 * (nInfo - 1));   // :381-382
                }
            }
        }
        _pb = uco_ba_problem{};
        _pb.n_markers = (int32_t)_mkSize.size(); _pb.marker_pose44 = _mkPose.data(); _pb.marker_size = _mkSize.data();
        _pb.n_marker_obs = (int32_t)_moMarker.size(); _pb.mobs_marker = _moMarker.data(); _pb.mobs_pose = _moPose.data();
        _pb.mobs_corners = _moCorners.data(); _pb.mobs_weight = _moWeight.data();
        if (!_planeOther.empty()) {
            _pb.n_plane = (int32_t)_planeOther.size(); _pb.plane_other = _planeOther.data(); _pb.plane_ref = planeRef;
            _pb.plane_ref_pose44 = planeRef < 0 ? _planeRef44.data() : nullptr; _pb.plane_weight = planeWeight;
        }
        _pb.n_poses = (int32_t)P; _pb.n_points = (int32_t)N; _pb.n_obs = (int32_t)_obsPose.size();
        _pb.poses44 = _poses.data(); _pb.fixed = _fixed.data(); _pb.points3 = _points.data();
        _pb.obs_pose = _obsPose.data(); _pb.obs_point = _obsPoint.data(); _pb.obs_uv = _obsUV.data(); _pb.obs_ur = _obsUR.data();
        _pb.obs_stereo = _obsStereo.data(); _pb.obs_inv_sigma2 = _obsInv.data();
        _pb.fx = f0.imageParams.fx(); _pb.fy = f0.imageParams.fy(); _pb.cx = f0.imageParams.cx(); _pb.cy = f0.imageParams.cy();
        _pb.bf = f0.imageParams.bl * f0.imageParams.fx();
        _poseCam.clear();
        if (mixedCameras) {   // every edge carries the ImageParams of its keyframe (:233-236, :262-266, :335-338)
            for (auto f : _frameIds) {
                const ImageParams& ip = map->keyframes[f].imageParams;
                const float c[5] = {ip.fx(), ip.fy(), ip.cx(), ip.cy(), ip.bl * ip.fx()};
                _poseCam.insert(_poseCam.end(), c, c + 5);
            }
            _pb.pose_cam = _poseCam.data();
        }
        _pb.n_iters = _params.nIters;
    }

    void optimize(bool* stopASAP = nullptr) override {
        _outPoses.resize(16 * (size_t)_pb.n_poses); _outPoints.resize(3 * (size_t)_pb.n_points); _outBad.resize(_pb.n_obs);
        uco_ba_result res{};
        res.poses44 = _outPoses.data(); res.points3 = _outPoints.data(); res.obs_bad = _outBad.data();
        _outMarkers.resize(16 * (size_t)_pb.n_markers);
        res.marker_poses44 = _outMarkers.data();
        static_assert(sizeof(bool) == 1, "the ABI polls a one-byte flag");
        // the mapper may flip *stopASAP from another thread (MapManager::stop, mapmanager.cpp:1614-1625): the solver polls it
        _ctx.check(uco_b200_ba_solve(_ctx.get(), &_pb, reinterpret_cast<const volatile unsigned char*>(stopASAP), &res));
    }

    void getResults(std::shared_ptr<Map> map) override {
        for (size_t k = 0; k < _frameIds.size(); k++) {   // :483-492
            if (_fixed[k]) continue;
            cv::Mat T(4, 4, CV_32F);
            std::memcpy(T.data, &_outPoses[16 * k], 64);
            map->keyframes[_frameIds[k]].pose_f2g = T.clone();
        }
        _badAssociations.clear();
        size_t o = 0;
        for (size_t j = 0; j < _pointIds.size(); j++) {   // :497-523 (observations were pushed point by point)
            map->map_points[_pointIds[j]].setCoordinates(
                cv::Point3f((float)_outPoints[3 * j], (float)_outPoints[3 * j + 1], (float)_outPoints[3 * j + 2]));
            for (; o < _obsPoint.size() && (size_t)_obsPoint[o] == j; o++)
                if (_outBad[o]) _badAssociations.push_back(std::make_pair(_pointIds[j], _frameIds[_obsPose[o]]));
        }
        for (size_t m = 0; m < _markerIds.size(); m++) {  // :526-527
            cv::Mat T(4, 4, CV_32F);
            std::memcpy(T.data, &_outMarkers[16 * m], 64);
            map->map_markers[_markerIds[m]].pose_g2m = T.clone();
        }
        for (auto pid : _pointIds) map->updatePointNormalAndDistances(pid);   // :533-536
    }

    void optimize(std::shared_ptr<Map> map, const ParamSet& p = ParamSet()) override {
        setParams(map, p);
        optimize();
        getResults(map);
    }
    vector<std::pair<uint32_t, uint32_t>> getBadAssociations() override {
        return _badAssociations;
    }

    const uco_ba_problem& problem() const { return _pb; }   // the flattened window (tests compare it with the reference's g2o)

protected:
    void saveToStream_impl(std::ostream&) override {}
    void readFromStream_impl(std::istream&) override {}

private:
    uco_b200::Context _ctx;
    ParamSet _params;
    uco_ba_problem _pb{};
    std::vector<uint32_t> _frameIds, _pointIds, _markerIds;
    std::vector<float> _poses, _points, _obsUV, _obsUR, _obsInv, _outPoses, _mkPose, _mkSize, _moCorners, _moWeight, _outMarkers, _poseCam;
    std::vector<int32_t> _moMarker, _moPose, _planeOther;
    std::vector<float> _planeRef44;
    std::vector<uint8_t> _fixed, _obsStereo, _outBad;
    std::vector<int32_t> _obsPose, _obsPoint;
    std::vector<double> _outPoints;
    vector<std::pair<uint32_t, uint32_t>> _badAssociations;
};

}  // namespace ucoslam
