// TEST INFRASTRUCTURE ONLY — stand-in for src/map.h: the containers the projection matchers / solvePnp / the BA adapter walk, with
// the reference's names (map.h:60-110): map_points / keyframes as id-indexed containers offering is() / operator[] / capacity() /
// iteration (the reference's ReusableContainer, basictypes/reusablecontainer.h), map_markers, and Map::getMapPointsInFrames with the
// reference's semantics (map.h:202-235: ascending ids of the non-bad points observed by the given, non-bad keyframes).
#pragma once
#include <mutex>
#include <set>
#include "map_types/mappoint.h"
#include "map_types/frame.h"
#include "map_types/marker.h"
namespace ucoslam {
template <typename T> class IdContainer {   // is() / operator[] / capacity() / size() / add(id) / iteration in ascending id
public:
    bool is(uint32_t id) const { return id < _used.size() && _used[id]; }
    T& operator[](uint32_t id) { if (!is(id)) throw std::runtime_error("IdContainer: no such element"); return _data[id]; }
    const T& operator[](uint32_t id) const { if (!is(id)) throw std::runtime_error("IdContainer: no such element"); return _data[id]; }
    size_t capacity() const { return _data.size(); }
    size_t size() const { size_t n = 0; for (auto u : _used) n += u; return n; }
    T& add(uint32_t id) { if (id >= _data.size()) { _data.resize(id + 1); _used.resize(id + 1, 0); } _used[id] = 1; return _data[id]; }
    T& front() { for (size_t i = 0; i < _used.size(); i++) if (_used[i]) return _data[i]; throw std::runtime_error("IdContainer: empty"); }
    struct iterator {
        IdContainer* c; size_t i;
        void skip() { while (i < c->_used.size() && !c->_used[i]) i++; }
        T& operator*() { return c->_data[i]; }
        iterator& operator++() { i++; skip(); return *this; }
        bool operator!=(const iterator& o) const { return i != o.i; }
    };
    iterator begin() { iterator it{this, 0}; it.skip(); return it; }
    iterator end() { return iterator{this, _used.size()}; }
private:
    std::vector<T> _data;
    std::vector<char> _used;
};
class Map {
public:
    IdContainer<MapPoint> map_points;
    IdContainer<Frame> keyframes;
    std::map<uint32_t, Marker> map_markers;
    std::map<uint32_t, std::set<uint32_t>> neighbors;   // stands for TheKpGraph
    std::set<uint32_t> getNeighborKeyFrames(uint32_t idx, bool includeIdx) { std::set<uint32_t> s = neighbors[idx]; if (includeIdx) s.insert(idx); return s; }
    int nNormalUpdates = 0;
    void updatePointNormalAndDistances(uint32_t) { nNormalUpdates++; }
    std::vector<cv::DMatch> matchFrameToMapPoints(const std::vector<uint32_t>& used_frames, Frame& curframe, const cv::Mat& pose_f2g, float minDescDist,
                                                  float maxRepjDist, bool markMapPointsAsVisible, bool useAllPoints = false,
                                                  std::set<uint32_t> excludedPoints = {});
    template <typename Iterator>
    std::vector<uint32_t> getMapPointsInFrames(Iterator fstart, Iterator fend, const std::set<uint32_t>& excludedPoints = {}) {
        std::vector<char> used(map_points.capacity(), 0);
        for (auto f = fstart; f != fend; f++) {
            if (!keyframes.is(*f) || keyframes[*f].isBad()) continue;
            for (auto id : keyframes[*f].ids)
                if (id != std::numeric_limits<uint32_t>::max()) used[id] = 1;
        }
        std::vector<uint32_t> out;
        for (size_t i = 0; i < used.size(); i++)
            if (used[i] && map_points.is(i) && !map_points[i].isBad() && !excludedPoints.count(i)) out.push_back(i);
        return out;
    }
};
}
