// keyframe_database_b200.h — the reference's keyframe-database seam: KFDataBaseVirtual / KPFrameDataBase
// (src/map_types/keyframedatabase.cpp:33-100, selected in KeyFrameDataBase::loadFromFile :329-333).  Same member functions,
// argument meaning and error behaviour; the votes and fBow::score of relocalizationCandidates (:195-233) run on the device-resident
// database (uco_b200_kfdb_*), computeBow (:310-321) on the device vocabulary, the covisibility accumulation (:236-275) through
// uco_b200_kfdb_rank fed by the reference's own CovisGraph.  The stream format of toStream_ / fromStream_ (:278-302) is the
// reference's (vocabulary, inverted word index, frame ids), so .map files stay interchangeable.
// KFDataBaseVirtual is private to keyframedatabase.cpp: to plug this class in, include this header there after that class with
//   #define UCO_B200_KFDB_BASE : public KFDataBaseVirtual
// and construct it in KeyFrameDataBase::loadFromFile / fromStream.  Needs fbow.h, the reference's Frame /
// FrameSet / CovisGraph (or oracle/shim for the container-only Frame), and the C ABI.
#pragma once
#include <istream>
#include <map>
#include <ostream>
#include <set>
#include <vector>
#include <fbow/fbow.h>
#include "map_types/frame.h"
#include "map_types/covisgraph.h"
#include "basictypes/io_utils.h"
#include "basictypes/hash.h"
#include "bow_b200.h"
#include "uco_b200_cxx.h"
#ifndef UCO_B200_KFDB_BASE
#define UCO_B200_KFDB_BASE
#endif

namespace uco_b200 {

class KPFrameDataBaseB200 UCO_B200_KFDB_BASE {
public:
    explicit KPFrameDataBaseB200(int device = 0) : _ctx(device), _dvoc(device) { _ctx.check(uco_b200_kfdb_create(_ctx.get(), &_db)); }
    ~KPFrameDataBaseB200() { uco_b200_kfdb_free(_ctx.get(), _db); }

    void loadFromFile(const std::string& filename) {              // :136-139
        _voc.readFromFile(filename);
        _dvoc.fromVocabulary(_voc);
        clear();
    }
    bool isEmpty() const { return frames.size() == 0; }
    size_t size() const { return frames.size(); }
    bool isId(uint32_t id) const { return frames.count(id) != 0; }
    const std::set<uint32_t> getFrames() const { return frames; }
    void clear() {                                                 // :140-146
        frames.clear();
        _frame_words.clear();
        _words.clear();
        _stale = false;
        _ctx.check(uco_b200_kfdb_clear(_ctx.get(), _db));
    }

    bool add(ucoslam::Frame& f) {                                  // :150-160
        if (_voc.size() == 0) return false;
        computeBow(f);
        register_frame(f.idx, *f.bowvector);
        return true;
    }
    bool del(const ucoslam::Frame& f) { return del(f.idx); }      // :162-170
    bool del(uint32_t fidx) {                                      // :171-178
        if (_voc.size() == 0) return false;
        if (!frames.count(fidx)) return true;                      // the reference erases nothing in that case
        if (!_stale) _ctx.check(uco_b200_kfdb_del(_ctx.get(), _db, fidx));
        _frame_words.erase(fidx);
        frames.erase(fidx);
        return true;
    }

    bool computeBow(ucoslam::Frame& f) {                           // :310-321
        if (_voc.size() == 0 || f.desc.rows == 0) return false;
        if (_voc.getDescSize() != uint32_t(f.desc.cols))
            throw std::runtime_error("FrameDataBase::computeBow Vocabulary and descriptor employed have different sizes. May be you are using a wrong descriptor type");
        if (_voc.getDescType() != uint32_t(f.desc.type()))
            throw std::runtime_error("FrameDataBase::computeBow Vocabulary and descriptor employed have different types. May be you are using a wrong descriptor type");
        _dvoc.transform(f.desc, 3, *f.bowvector, *f.bowvector_level);
        return true;
    }

    float score(ucoslam::Frame& f, ucoslam::Frame& f2) {           // :304-308 (two map walks on the host: nothing to offload)
        if (f.bowvector->size() == 0) computeBow(f);
        if (f2.bowvector->size() == 0) computeBow(f2);
        return fbow::fBow::score(*f.bowvector, *f2.bowvector);
    }

    std::vector<uint32_t> relocalizationCandidates(ucoslam::Frame& frame, ucoslam::FrameSet& fset, ucoslam::CovisGraph& covisgraph,
                                                   bool sorted = true, float minScore = 0,
                                                   const std::set<uint32_t>& excludedFrames = {}) {
        if (_voc.size() == 0) throw std::runtime_error("no vocabulary");                       // :198
        if (frame.bowvector->size() == 0) computeBow(frame);
        if (_stale) rebuild(fset);
        flatten(*frame.bowvector, _qw, _qf);
        std::vector<uint32_t> exc(excludedFrames.begin(), excludedFrames.end());
        const int cap = (int)frames.size();
        _of.resize(cap); _os.resize(cap);
        int n = 0;
        _ctx.check(uco_b200_kfdb_query(_ctx.get(), _db, _qw.data(), _qf.data(), (int)_qw.size(), exc.data(), (int)exc.size(), minScore,
                                       _of.data(), _os.data(), nullptr, cap, &n, nullptr));
        if (n == 0) return {};
        std::vector<int32_t> off(n + 1, 0);
        std::vector<uint32_t> nbr;
        if (n > 1)
            for (int i = 0; i < n; i++) {                                                      // :245-249
                auto nw = covisgraph.getNeighborsWeights(_of[i], true);
                for (size_t k = 0; k < nw.size() && k < 10; k++) nbr.push_back(nw[k].first);
                off[i + 1] = (int32_t)nbr.size();
            }
        std::vector<uint32_t> out(n);
        int no = 0;
        if (uco_b200_kfdb_rank(_of.data(), _os.data(), n, off.data(), nbr.data(), sorted ? 1 : 0, minScore, out.data(), &no) != UCO_OK)
            throw std::runtime_error("ucoslam_b200: kfdb_rank rejected its arguments");
        out.resize(no);
        return out;
    }

    void toStream_(std::iostream& str) const {                     // :278-287
        _voc.toStream(str);
        const std::map<uint32_t, std::set<uint32_t>> word_frames = inverted();
        ucoslam::io_write<uint32_t>(word_frames.size(), str);
        for (const auto& w : word_frames) {
            ucoslam::io_write<uint32_t>(w.first, str);
            ucoslam::toStream__(w.second, str);
        }
        ucoslam::toStream__(frames, str);
    }
    // the stream holds words per frame but no weights: the device database is refilled from the frames' bowvectors at the next query
    void fromStream_(std::istream& str) {                          // :290-302
        _voc.fromStream(str);
        _dvoc.fromVocabulary(_voc);
        clear();
        const int s = ucoslam::io_read<uint32_t>(str);
        for (int i = 0; i < s; i++) {
            const uint32_t word = ucoslam::io_read<uint32_t>(str);
            std::set<uint32_t> fr;
            ucoslam::fromStream__(fr, str);
            _words.insert(word);
            for (uint32_t f : fr) _frame_words[f].push_back(word);   // words arrive ascending: the lists stay sorted
        }
        ucoslam::fromStream__(frames, str);
        _stale = true;
    }
    uint64_t getSignature() const {                                // :184-193
        const std::map<uint32_t, std::set<uint32_t>> word_frames = inverted();
        ucoslam::Hash sig;
        for (auto wf : word_frames) {
            sig += wf.first;
            sig.add(wf.second.begin(), wf.second.end());
        }
        sig.add(frames.begin(), frames.end());
        sig += _voc.hash();
        return sig;
    }

private:
    // the reference's word_frames_: a word keeps its (possibly empty) entry after its last frame is deleted (:162-170)
    std::map<uint32_t, std::set<uint32_t>> inverted() const {
        std::map<uint32_t, std::set<uint32_t>> word_frames;
        for (uint32_t w : _words) word_frames[w];
        for (const auto& fw : _frame_words)
            for (uint32_t w : fw.second) word_frames[w].insert(fw.first);
        return word_frames;
    }
    static void flatten(const fbow::fBow& b, std::vector<uint32_t>& w, std::vector<float>& f) {
        w.clear(); f.clear();
        for (const auto& e : b) { w.push_back(e.first); f.push_back((float)e.second); }
    }
    void register_frame(uint32_t idx, const fbow::fBow& bow) {
        flatten(bow, _qw, _qf);
        if (!_stale) _ctx.check(uco_b200_kfdb_add(_ctx.get(), _db, idx, _qw.data(), _qf.data(), (int)_qw.size()));
        _frame_words[idx] = _qw;
        _words.insert(_qw.begin(), _qw.end());
        frames.insert(idx);
    }
    void rebuild(ucoslam::FrameSet& fset) {
        _ctx.check(uco_b200_kfdb_clear(_ctx.get(), _db));
        for (uint32_t f : frames) {
            flatten(*fset[f].bowvector, _qw, _qf);
            _ctx.check(uco_b200_kfdb_add(_ctx.get(), _db, f, _qw.data(), _qf.data(), (int)_qw.size()));
        }
        _stale = false;
    }

    Context _ctx;
    VocabularyB200 _dvoc;
    mutable fbow::Vocabulary _voc;     // kept for toStream / hash / descriptor checks (fbow's accessors are not const)
    uco_b200_kfdb* _db = nullptr;
    std::set<uint32_t> frames;
    std::map<uint32_t, std::vector<uint32_t>> _frame_words;   // what the reference keeps inverted as word_frames_
    std::set<uint32_t> _words;                                // every word registered since the last clear()
    bool _stale = false;
    std::vector<uint32_t> _qw, _of;
    std::vector<float> _qf;
    std::vector<double> _os;
};

}  // namespace uco_b200
