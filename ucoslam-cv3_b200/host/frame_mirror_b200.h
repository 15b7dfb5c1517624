// frame_mirror_b200.h — a ucoslam::Frame on the device.  The reference already has ONE complete, versioned description
// of a Frame: its stream (Frame::toStream, src/map_types/frame.cpp:260-302 — what .map / .slm files and Map::toStream hold).  The
// mirror is built from exactly those bytes: toStream -> uco_b200_frame_stream_parse (a view, no copy) -> uco_b200_frame_upload (one
// allocation: keypoints, descriptors, map-point ids, flags, depth, pose, scale factors, the kd-tree flattened), so a keyframe of a map
// loaded from disk reaches the matcher / tracker / mapper kernels without per-field marshalling code that could drift from the class.
// dev() gives the device pointers the *_dev entry points take (uco_b200_frame_match_batch_dev, uco_b200_keyframes_batch_dev,
// uco_b200_track_*_dev, uco_b200_kdtree_*).  store() writes the per-keypoint arrays a device stage changed (ids, flags) back.
// Needs the reference's headers and OpenCV C++ (see the note in orb_extractor_b200.h).
#pragma once
#include <sstream>
#include <string>
#include "uco_b200_cxx.h"

namespace ucoslam {

class FrameMirrorB200 {
public:
    FrameMirrorB200(uco_b200::Context& ctx, const Frame& f) : _ctx(ctx) {
        std::stringstream ss;
        f.toStream(ss);
        _bytes = ss.str();
        size_t used = 0;
        if (uco_b200_frame_stream_parse(reinterpret_cast<const uint8_t*>(_bytes.data()), _bytes.size(), &_view, &used) != UCO_OK)
            throw std::runtime_error("FrameMirrorB200: Frame::toStream produced a stream the parser rejects");
        _ctx.check(uco_b200_frame_upload(_ctx.get(), &_view, &_frame));
    }
    ~FrameMirrorB200() { uco_b200_frame_free(_frame); }
    FrameMirrorB200(const FrameMirrorB200&) = delete;
    FrameMirrorB200& operator=(const FrameMirrorB200&) = delete;
    const uco_frame_dev& dev() const { return *uco_b200_frame_dev(_frame); }
    const uco_frame_stream& view() const { return _view; }
    // ids / flags as the device holds them now -> the Frame (Frame::ids, Frame::flags)
    void store(Frame& f) const {
        std::vector<uint8_t> fl(f.flags.size());
        _ctx.check(uco_b200_frame_download(_ctx.get(), _frame, nullptr, nullptr, f.ids.data(), fl.data(), nullptr));
        for (size_t i = 0; i < fl.size(); i++) f.flags[i].v = fl[i];
    }
private:
    uco_b200::Context& _ctx;
    std::string _bytes;
    uco_frame_stream _view{};
    uco_b200_frame* _frame = nullptr;
};

}  // namespace ucoslam
