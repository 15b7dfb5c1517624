// frame_stream.cu — SURVEY 8(f)4: the reference's Frame as it travels in .map / .slm files and between the tracker and the mapper,
// and its device-resident mirror.
//
// Restates the byte layout of (reference, relative to /root/reference):
//   src/map_types/frame.cpp:260-302, 304-341      Frame::toStream / fromStream (magic 134243 ... 134244)
//   src/basictypes/io_utils.h:45-66, 171-183      toStream__(std::vector<T>) (u32 count + raw elements), toStream__ts (u32 count + toStream each)
//   src/basictypes/io_utils.cpp:21-52             toStream__(cv::Mat): rows, cols, type (ints), then the rows without padding
//   src/basictypes/io_utils.cpp:54-65             toStream__(std::string)
//   src/map_types/marker.cpp:94-113               MarkerObservation::toStream (dict_info, corners, und_corners, ssize, id, poses)
//   src/map_types/frame.cpp:372-386               MarkerPosesIPPE::toStream (two matrices, errs[2], err_ratio)
//   src/basictypes/se3transform.h:178-188         Se3Transform::toStream (signature 928511272 + 16 floats)
//   3rdparty/fbow/fbow/fbow.cpp:261-303           fBow::toStream (u32 count + (u32 word, float weight) pairs), fBow2::toStream
//   src/map_types/mappoint.cpp:117-175            MapPoint::toStream / fromStream (magic 123200)
//   src/imageparams.cpp:68-84                     ImageParams::toStream (CameraMatrix, Distorsion, CamSize, bl, rgb_depthscale)
//   src/basictypes/picoflann.h:603-660            KdTreeIndex::toStream (walked here only to find its length; uco_b200_kdtree_parse reads it)
// The parser yields a VIEW: pointers into the caller's buffer, nothing copied.  The writer produces, byte for byte, what the
// reference's toStream writes for the same field values (tests/test_frame_stream.py compares with the reference's own statements
// compiled against container stand-ins).  The device mirror holds what the *_dev entry points consume — keypoints, descriptors,
// map-point ids, flags, depth, pose, scale factors and the flattened kd-tree — so that matcher / tracker / mapper kernels chain on a
// keyframe loaded from a map file without host vectors in between.
#include "common.cuh"
#include <cstring>
#include <vector>

namespace {
struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    bool need(size_t k) { if ((size_t)(end - p) < k) ok = false; return ok; }
    template <class T> T get() { T v{}; if (need(sizeof(T))) { memcpy(&v, p, sizeof(T)); p += sizeof(T); } return v; }
    const uint8_t* skip(size_t k) { const uint8_t* q = p; if (need(k)) p += k; return q; }
};
size_t elem_size(int type) {   // OpenCV type codes: depth = type & 7, channels = (type >> 3) + 1
    static const int depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return (size_t)depth_bytes[type & 7] * (size_t)((type >> 3) + 1);
}
bool read_mat(Reader& R, uco_mat_view& m) {
    m.rows = R.get<int32_t>(); m.cols = R.get<int32_t>(); m.type = R.get<int32_t>();
    m.data = nullptr;
    if (!R.ok || m.rows < 0 || m.cols < 0 || m.type < 0 || m.type > 511) return R.ok = false;
    if ((int64_t)m.rows * m.cols > 0) m.data = R.skip((size_t)m.rows * m.cols * elem_size(m.type));
    return R.ok;
}
bool skip_mat(Reader& R) { uco_mat_view m; return read_mat(R, m); }
template <class T> bool read_vec(Reader& R, uint32_t& n, const T*& ptr, size_t elem) {
    n = R.get<uint32_t>();
    ptr = (const T*)R.skip((size_t)n * elem);
    return R.ok;
}
struct Writer {
    uint8_t* p; size_t cap, n = 0;
    void put(const void* src, size_t k) { if (p && n + k <= cap && k) memcpy(p + n, src, k); n += k; }
    template <class T> void val(T v) { put(&v, sizeof(T)); }
    void mat(const uco_mat_view& m) {
        const bool empty = !m.data || (int64_t)m.rows * m.cols <= 0;
        val<int32_t>(empty ? 0 : m.rows); val<int32_t>(empty ? 0 : m.cols); val<int32_t>(empty ? 0 : m.type);
        if (!empty) put(m.data, (size_t)m.rows * m.cols * elem_size(m.type));
    }
};
}  // namespace

extern "C" {

int uco_b200_frame_stream_parse(const uint8_t* bytes, size_t len, uco_frame_stream* v, size_t* consumed) {
    if (!bytes || !v) return UCO_E_INVALID;
    memset(v, 0, sizeof *v);
    Reader R{bytes, bytes + len};
    if (R.get<int32_t>() != 134243) return UCO_E_INVALID;
    v->idx = R.get<uint32_t>(); v->fseq_idx = R.get<uint32_t>(); v->frame_flags = R.get<uint8_t>(); v->kp_desc_type = R.get<int8_t>();
    read_mat(R, v->desc);
    read_vec(R, v->n_und_kpts, v->und_kpts, sizeof(uco_keypoint));
    read_vec(R, v->n_kpts, v->kpts, 8);
    read_vec(R, v->n_depth, v->depth, 4);
    read_vec(R, v->n_ids, v->ids, 4);
    read_vec(R, v->n_flags, v->flags, 1);
    v->n_markers = R.get<uint32_t>();
    v->markers = R.p;
    for (uint32_t i = 0; i < v->n_markers && R.ok; i++) {   // MarkerObservation::toStream
        R.skip(R.get<uint32_t>());                        // dict_info
        R.skip((size_t)R.get<uint32_t>() * 8);            // corners
        R.skip((size_t)R.get<uint32_t>() * 8);            // und_corners
        R.skip(8);                                        // ssize, id
        skip_mat(R); skip_mat(R);                         // poses.sols
        R.skip(24);                                       // errs[2], err_ratio
    }
    v->markers_bytes = (uint64_t)(R.p - v->markers);
    if (R.get<uint32_t>() != 928511272u) return UCO_E_INVALID;   // Se3Transform signature
    if (R.need(64)) { memcpy(v->pose_f2g, R.p, 64); R.p += 64; }
    read_vec(R, v->n_bow, v->bow, 8);
    v->n_bow_level = R.get<uint32_t>();
    v->bow_level = R.p;
    for (uint32_t i = 0; i < v->n_bow_level && R.ok; i++) { R.skip(4); R.skip((size_t)R.get<uint32_t>() * 4); }
    v->bow_level_bytes = (uint64_t)(R.p - v->bow_level);
    read_vec(R, v->n_scale_factors, v->scale_factors, 4);
    read_mat(R, v->camera_matrix); read_mat(R, v->distortion);
    v->cam_size[0] = R.get<int32_t>(); v->cam_size[1] = R.get<int32_t>(); v->bl = R.get<float>(); v->rgb_depthscale = R.get<float>();
    read_mat(R, v->image);
    v->kdtree = R.p;
    {   // KdTreeIndex::toStream: dims, nValues, bbox, nodes
        R.skip(8);
        const uint64_t nb = R.get<uint64_t>();
        if (nb > 16) R.ok = false;
        R.skip(16 * (size_t)nb);
        const uint64_t k = R.get<uint64_t>();
        for (uint64_t i = 0; i < k && R.ok; i++) { R.skip(8 + 2 + 4 + 4 + 8 + 8); R.skip((size_t)R.get<uint64_t>() * 4); }
    }
    v->kdtree_bytes = (uint64_t)(R.p - v->kdtree);
    v->min_xy[0] = R.get<int32_t>(); v->min_xy[1] = R.get<int32_t>(); v->max_xy[0] = R.get<int32_t>(); v->max_xy[1] = R.get<int32_t>();
    if (R.get<int32_t>() != 134244 || !R.ok) return UCO_E_INVALID;
    if (consumed) *consumed = (size_t)(R.p - bytes);
    return UCO_OK;
}

/* out == NULL: only the size is computed */
int uco_b200_frame_stream_write(const uco_frame_stream* v, uint8_t* out, size_t cap, size_t* written) {
    if (!v || !written) return UCO_E_INVALID;
    Writer W{out, cap};
    W.val<int32_t>(134243);
    W.val(v->idx); W.val(v->fseq_idx); W.val(v->frame_flags); W.val(v->kp_desc_type);
    W.mat(v->desc);
    W.val(v->n_und_kpts); W.put(v->und_kpts, sizeof(uco_keypoint) * (size_t)v->n_und_kpts);
    W.val(v->n_kpts); W.put(v->kpts, 8 * (size_t)v->n_kpts);
    W.val(v->n_depth); W.put(v->depth, 4 * (size_t)v->n_depth);
    W.val(v->n_ids); W.put(v->ids, 4 * (size_t)v->n_ids);
    W.val(v->n_flags); W.put(v->flags, (size_t)v->n_flags);
    W.val(v->n_markers); W.put(v->markers, (size_t)v->markers_bytes);
    W.val<uint32_t>(928511272u); W.put(v->pose_f2g, 64);
    W.val(v->n_bow); W.put(v->bow, 8 * (size_t)v->n_bow);
    W.val(v->n_bow_level); W.put(v->bow_level, (size_t)v->bow_level_bytes);
    W.val(v->n_scale_factors); W.put(v->scale_factors, 4 * (size_t)v->n_scale_factors);
    W.mat(v->camera_matrix); W.mat(v->distortion);
    W.val(v->cam_size[0]); W.val(v->cam_size[1]); W.val(v->bl); W.val(v->rgb_depthscale);
    W.mat(v->image);
    W.put(v->kdtree, (size_t)v->kdtree_bytes);
    W.val(v->min_xy[0]); W.val(v->min_xy[1]); W.val(v->max_xy[0]); W.val(v->max_xy[1]);
    W.val<int32_t>(134244);
    *written = W.n;
    if (out && W.n > cap) return UCO_E_CAPACITY;
    return UCO_OK;
}

/* MapPoint::toStream / fromStream (src/map_types/mappoint.cpp:117-175): magic 123200, id, pos3d, descriptor (cv::Mat), the observing
 * frames (std::map<uint32,uint32> as u32 count + pairs), normal, nTimesSeen, nTimesVisible (uint16), flags (uint8), max / min
 * distance, kfSinceAddition (uint64), lastFIdxSeen. */
int uco_b200_mappoint_stream_parse(const uint8_t* bytes, size_t len, uco_mappoint_stream* v, size_t* consumed) {
    if (!bytes || !v) return UCO_E_INVALID;
    memset(v, 0, sizeof *v);
    Reader R{bytes, bytes + len};
    if (R.get<int32_t>() != 123200) return UCO_E_INVALID;
    v->id = R.get<uint32_t>();
    for (int k = 0; k < 3; k++) v->pos3d[k] = R.get<float>();
    read_mat(R, v->desc);
    v->n_frames = R.get<uint32_t>();
    v->frames = (const uint32_t*)R.skip(8 * (size_t)v->n_frames);
    for (int k = 0; k < 3; k++) v->normal[k] = R.get<float>();
    v->n_times_seen = R.get<uint16_t>(); v->n_times_visible = R.get<uint16_t>(); v->flags = R.get<uint8_t>();
    v->max_distance = R.get<float>(); v->min_distance = R.get<float>();
    v->kf_since_addition = R.get<uint64_t>(); v->last_fidx_seen = R.get<uint32_t>();
    if (!R.ok) return UCO_E_INVALID;
    if (consumed) *consumed = (size_t)(R.p - bytes);
    return UCO_OK;
}
int uco_b200_mappoint_stream_write(const uco_mappoint_stream* v, uint8_t* out, size_t cap, size_t* written) {
    if (!v || !written) return UCO_E_INVALID;
    Writer W{out, cap};
    W.val<int32_t>(123200);
    W.val(v->id); W.put(v->pos3d, 12);
    W.mat(v->desc);
    W.val(v->n_frames); W.put(v->frames, 8 * (size_t)v->n_frames);
    W.put(v->normal, 12);
    W.val(v->n_times_seen); W.val(v->n_times_visible); W.val(v->flags); W.val(v->max_distance); W.val(v->min_distance);
    W.val(v->kf_since_addition); W.val(v->last_fidx_seen);
    *written = W.n;
    if (out && W.n > cap) return UCO_E_CAPACITY;
    return UCO_OK;
}

/* The map-point section of a map file: Map::toStream writes `map_points.toStream(str)` (src/map.cpp:316-325), a ReusableContainer<MapPoint>
 * (src/basictypes/reusablecontainer.h:276-291): int64 magic 123299999, u32 count + the indices of the free slots (in the order they were freed;
 * insert() reuses the LAST one), then ExpansibleContainer<Pair>::toStream (expansiblecontainer.h:115-129): u64 signature 13218888, u64 number
 * of chunks, every slot of every chunk (chunks hold 200 slots: the container's default, which its reader also assumes) as {bool valid,
 * MapPoint stream}, then int curBuffer, curElm, chunk size.  Slots past size() = curBuffer * chunk + curElm hold default-constructed points. */
enum { MP_CHUNK = 200 };
static int container_walk(const uint8_t* bytes, size_t len, bool frames, uco_mappoint_container* c, size_t* slot_offset, uint8_t* slot_valid, uint32_t cap,
                          size_t* consumed) {
    if (!bytes || !c) return UCO_E_INVALID;
    memset(c, 0, sizeof *c);
    Reader R{bytes, bytes + len};
    if (frames && R.get<int32_t>() != 88888) return UCO_E_INVALID;   // FrameSet::toStream, frame.cpp:350-355
    if (R.get<int64_t>() != 123299999) return UCO_E_INVALID;
    c->n_free = R.get<uint32_t>();
    c->free_slots = (const uint32_t*)R.skip(4 * (size_t)c->n_free);
    if (!R.ok || R.get<uint64_t>() != 13218888) return UCO_E_INVALID;
    const uint64_t chunks = R.get<uint64_t>();
    if (!R.ok || chunks > (uint64_t)(len / MP_CHUNK) + 1) return UCO_E_INVALID;   // every slot takes more than one byte
    c->n_slots = (uint32_t)(chunks * MP_CHUNK);
    uint32_t n_valid = 0;
    for (uint32_t i = 0; i < c->n_slots; i++) {
        const uint8_t valid = R.get<uint8_t>();
        if (!R.ok) return UCO_E_INVALID;
        size_t used = 0;
        if (frames) {
            uco_frame_stream v;
            if (uco_b200_frame_stream_parse(R.p, (size_t)(R.end - R.p), &v, &used) != UCO_OK) return UCO_E_INVALID;
        } else {
            uco_mappoint_stream v;
            if (uco_b200_mappoint_stream_parse(R.p, (size_t)(R.end - R.p), &v, &used) != UCO_OK) return UCO_E_INVALID;
        }
        if (i < cap) {
            if (slot_offset) slot_offset[i] = (size_t)(R.p - bytes);
            if (slot_valid) slot_valid[i] = valid;
        }
        n_valid += valid != 0;
        R.skip(used);
    }
    const int32_t cur_buffer = R.get<int32_t>(), cur_elm = R.get<int32_t>(), chunk = R.get<int32_t>();
    if (!R.ok || chunk != MP_CHUNK || cur_buffer < 0 || cur_elm < 0 || cur_elm > MP_CHUNK || (uint64_t)cur_buffer >= chunks) return UCO_E_INVALID;
    c->n_used = (uint32_t)cur_buffer * MP_CHUNK + (uint32_t)cur_elm;
    c->n_valid = n_valid;
    if (consumed) *consumed = (size_t)(R.p - bytes);
    return c->n_slots > cap && (slot_offset || slot_valid) ? UCO_E_CAPACITY : UCO_OK;
}
int uco_b200_mappoint_container_walk(const uint8_t* bytes, size_t len, uco_mappoint_container* c, size_t* slot_offset, uint8_t* slot_valid, uint32_t cap,
                                     size_t* consumed) {
    return container_walk(bytes, len, false, c, slot_offset, slot_valid, cap, consumed);
}
/* The keyframe section of a map file: FrameSet::toStream (src/map_types/frame.cpp:350-355) = int magic 88888 + the same container over Frame
 * streams; slot_offset[i] is where uco_b200_frame_stream_parse (and from there uco_b200_frame_upload) reads keyframe slot i. */
int uco_b200_frame_container_walk(const uint8_t* bytes, size_t len, uco_mappoint_container* c, size_t* slot_offset, uint8_t* slot_valid, uint32_t cap,
                                  size_t* consumed) {
    return container_walk(bytes, len, true, c, slot_offset, slot_valid, cap, consumed);
}
/* the valid points of the section as the flat rows uco_mappoints / uco_b200_track_state_set_map take (ids = MapPoint::id); every output may be NULL */
int uco_b200_mappoints_from_container(const uint8_t* bytes, size_t len, uint32_t cap, uint32_t* ids, float* pos, float* normal, float* min_dist, float* max_dist,
                                      uint8_t* desc, uint8_t* flags, uint32_t* n_out, size_t* consumed) {
    if (!bytes || !n_out) return UCO_E_INVALID;
    uco_mappoint_container c;
    int rc = uco_b200_mappoint_container_walk(bytes, len, &c, nullptr, nullptr, 0, consumed);
    if (rc != UCO_OK) return rc;
    *n_out = c.n_valid;
    if (c.n_valid > cap) return UCO_E_CAPACITY;
    const uint8_t* p = bytes + 8 + 4 + 4 * (size_t)c.n_free + 16;
    uint32_t k = 0;
    for (uint32_t i = 0; i < c.n_slots; i++) {
        const uint8_t valid = *p++;
        uco_mappoint_stream v;
        size_t used = 0;
        uco_b200_mappoint_stream_parse(p, (size_t)(bytes + len - p), &v, &used);
        p += used;
        if (!valid) continue;
        if (desc && (v.desc.rows != 1 || (size_t)v.desc.cols * elem_size(v.desc.type) != 32)) return UCO_E_INVALID;   // ORB rows only
        if (ids) ids[k] = v.id;
        if (pos) memcpy(pos + 3 * (size_t)k, v.pos3d, 12);
        if (normal) memcpy(normal + 3 * (size_t)k, v.normal, 12);
        if (min_dist) min_dist[k] = v.min_distance;
        if (max_dist) max_dist[k] = v.max_distance;
        if (desc) memcpy(desc + 32 * (size_t)k, v.desc.data, 32);
        if (flags) flags[k] = v.flags;
        k++;
    }
    return UCO_OK;
}
/* the writer: n_slots must be a multiple of 200; slot i is written as {valid[i], points[i]} (the caller passes default points for unused slots:
 * uco_b200_mappoint_stream_default) */
int uco_b200_mappoint_container_write(const uco_mappoint_container* c, const uco_mappoint_stream* points, const uint8_t* valid, uint8_t* out, size_t cap,
                                      size_t* written) {
    if (!c || !written || c->n_slots % MP_CHUNK || c->n_slots == 0 || c->n_used > c->n_slots || (c->n_slots && (!points || !valid)) || (c->n_free && !c->free_slots))
        return UCO_E_INVALID;
    Writer W{out, cap};
    W.val<int64_t>(123299999);
    W.val(c->n_free); W.put(c->free_slots, 4 * (size_t)c->n_free);
    W.val<uint64_t>(13218888); W.val<uint64_t>(c->n_slots / MP_CHUNK);
    for (uint32_t i = 0; i < c->n_slots; i++) {
        W.val<uint8_t>(valid[i] ? 1 : 0);
        size_t n = 0;
        int rc = uco_b200_mappoint_stream_write(points + i, out && W.n <= cap ? out + W.n : nullptr, out && W.n <= cap ? cap - W.n : 0, &n);
        if (rc != UCO_OK && rc != UCO_E_CAPACITY) return rc;
        W.n += n;
    }
    // ExpansibleContainer::push_back (expansiblecontainer.h:44-52): after k >= 1 pushes curBuffer = (k - 1) / chunk, curElm = (k - 1) % chunk + 1
    W.val<int32_t>(c->n_used ? (int32_t)((c->n_used - 1) / MP_CHUNK) : 0);
    W.val<int32_t>(c->n_used ? (int32_t)((c->n_used - 1) % MP_CHUNK) + 1 : 0);
    W.val<int32_t>(MP_CHUNK);
    *written = W.n;
    if (out && W.n > cap) return UCO_E_CAPACITY;
    return UCO_OK;
}
/* the keyframe section from already-serialised Frame streams (uco_b200_frame_stream_write output, or ranges of another map file): int 88888 + the
 * container framing around slot_bytes[i] (slot_len[i] bytes each).  Unused slots need the stream of a default-constructed Frame: copy one from any
 * reference-written keyframe section (its slots past n_used). */
int uco_b200_frame_container_write(const uco_mappoint_container* c, const uint8_t* const* slot_bytes, const size_t* slot_len, const uint8_t* valid, uint8_t* out,
                                   size_t cap, size_t* written) {
    if (!c || !written || c->n_slots % MP_CHUNK || c->n_slots == 0 || c->n_used > c->n_slots || !slot_bytes || !slot_len || !valid || (c->n_free && !c->free_slots))
        return UCO_E_INVALID;
    Writer W{out, cap};
    W.val<int32_t>(88888);
    W.val<int64_t>(123299999);
    W.val(c->n_free); W.put(c->free_slots, 4 * (size_t)c->n_free);
    W.val<uint64_t>(13218888); W.val<uint64_t>(c->n_slots / MP_CHUNK);
    for (uint32_t i = 0; i < c->n_slots; i++) {
        if (!slot_bytes[i]) return UCO_E_INVALID;
        W.val<uint8_t>(valid[i] ? 1 : 0);
        W.put(slot_bytes[i], slot_len[i]);
    }
    W.val<int32_t>(c->n_used ? (int32_t)((c->n_used - 1) / MP_CHUNK) : 0);
    W.val<int32_t>(c->n_used ? (int32_t)((c->n_used - 1) % MP_CHUNK) + 1 : 0);
    W.val<int32_t>(MP_CHUNK);
    *written = W.n;
    return out && W.n > cap ? UCO_E_CAPACITY : UCO_OK;
}
/* what a default-constructed MapPoint streams as (mappoint.h:111-131 member initialisers): the content of never-used slots */
void uco_b200_mappoint_stream_default(uco_mappoint_stream* v) {
    memset(v, 0, sizeof *v);
    v->id = 0xffffffffu;
    v->last_fidx_seen = 0xffffffffu;
    v->max_distance = 1.17549435e-38f;   // std::numeric_limits<float>::min()
    v->min_distance = 3.40282347e+38f;   // std::numeric_limits<float>::max()
}

/* ---- the other sections of Map::toStream (src/map.cpp:316-325), and the whole stream ------------------------------------------------------------
 * keyframe database: KeyFrameDataBase::toStream (keyframedatabase.cpp:335-340) = int type, then for type 1 KPFrameDataBase::toStream_ (:278-287): the
 *   fbow vocabulary (fbow.cpp:171-178: u64 55824124, the 120-byte params block, _total_size bytes), u32 count + per word {u32 word, set<u32> of frames},
 *   set<u32> of frame ids; for type 0 (DummyDataBase, :122-124) the set of frame ids alone.  A set<u32> is u32 count + the values (io_utils.h:52-58).
 * markers: toStream__kv_complex over std::map<u32, Marker> (io_utils.h:113-121): u32 count + per marker {u32 key, Marker::toStream (marker.cpp:41-47):
 *   u32 id, Se3Transform (u32 928511272 + 16 floats), float size, set<u32> frames, string (u32 length + bytes)}.
 * covisibility graph: CovisGraph::toStream (covisgraph.cpp:308-333): set<u32> nodes, u32 count + per node {u32 id, set<u32> neighbours}, u32 count +
 *   per edge {u64 key, float weight}. */
static bool skip_set(Reader& R, uint32_t* n = nullptr, const uint32_t** v = nullptr) {
    const uint32_t k = R.get<uint32_t>();
    const uint8_t* p = R.skip(4 * (size_t)k);
    if (n) *n = k;
    if (v) *v = (const uint32_t*)p;
    return R.ok;
}
int uco_b200_kfdb_stream_walk(const uint8_t* bytes, size_t len, uco_kfdb_stream* o, size_t* consumed) {
    if (!bytes || !o) return UCO_E_INVALID;
    memset(o, 0, sizeof *o);
    Reader R{bytes, bytes + len};
    o->type = R.get<int32_t>();
    if (!R.ok || (o->type != 0 && o->type != 1)) return UCO_E_INVALID;
    if (o->type == 1) {
        o->voc_off = (size_t)(R.p - bytes);
        if (R.get<uint64_t>() != 55824124ull) return UCO_E_INVALID;
        const uint8_t* prm = R.skip(120);
        if (!R.ok) return UCO_E_INVALID;
        uint64_t total = 0;
        memcpy(&total, prm + 96, 8);
        R.skip((size_t)total);
        if (!R.ok) return UCO_E_INVALID;
        o->voc_len = (size_t)(R.p - bytes) - o->voc_off;
        o->n_words = R.get<uint32_t>();
        o->words_off = (size_t)(R.p - bytes);
        for (uint32_t w = 0; w < o->n_words && R.ok; w++) {
            R.get<uint32_t>();
            uint32_t k = 0;
            skip_set(R, &k);
            o->n_word_frames += k;
        }
    }
    if (!skip_set(R, &o->n_frames, &o->frames)) return UCO_E_INVALID;
    if (consumed) *consumed = (size_t)(R.p - bytes);
    return UCO_OK;
}
int uco_b200_marker_map_walk(const uint8_t* bytes, size_t len, uint32_t cap, uco_marker_stream* out, uint32_t* n_out, size_t* consumed) {
    if (!bytes || !n_out) return UCO_E_INVALID;
    Reader R{bytes, bytes + len};
    const uint32_t n = R.get<uint32_t>();
    if (!R.ok || n > len / 80) return UCO_E_INVALID;     // a marker record takes at least 4 + 4 + 68 + 4 + 4 + 4 bytes
    *n_out = n;
    for (uint32_t i = 0; i < n; i++) {
        uco_marker_stream m;
        memset(&m, 0, sizeof m);
        m.key = R.get<uint32_t>();
        m.id = R.get<uint32_t>();
        if (R.get<uint32_t>() != 928511272u) return UCO_E_INVALID;
        const uint8_t* pose = R.skip(64);
        if (!R.ok) return UCO_E_INVALID;
        memcpy(m.pose_g2m, pose, 64);
        m.size = R.get<float>();
        skip_set(R, &m.n_frames, &m.frames);
        m.dict_len = R.get<uint32_t>();
        m.dict = (const char*)R.skip(m.dict_len);
        if (!R.ok) return UCO_E_INVALID;
        if (out && i < cap) out[i] = m;
    }
    if (consumed) *consumed = (size_t)(R.p - bytes);
    return out && n > cap ? UCO_E_CAPACITY : UCO_OK;
}
int uco_b200_covis_stream_walk(const uint8_t* bytes, size_t len, uco_covis_stream* o, size_t* consumed) {
    if (!bytes || !o) return UCO_E_INVALID;
    memset(o, 0, sizeof *o);
    Reader R{bytes, bytes + len};
    if (!skip_set(R, &o->n_nodes, &o->nodes)) return UCO_E_INVALID;
    o->n_adj = R.get<uint32_t>();
    o->adj_off = (size_t)(R.p - bytes);
    for (uint32_t i = 0; i < o->n_adj && R.ok; i++) {
        R.get<uint32_t>();
        uint32_t k = 0;
        skip_set(R, &k);
        o->n_neighbours += k;
    }
    o->n_weights = R.get<uint32_t>();
    o->weights = R.skip(12 * (size_t)o->n_weights);     // packed {u64 key, float weight} records (12 bytes each, unaligned)
    if (!R.ok) return UCO_E_INVALID;
    if (consumed) *consumed = (size_t)(R.p - bytes);
    return UCO_OK;
}
/* ---- writers / unpackers of those three sections: with the Frame, MapPoint and container writers above a whole map file can be written here ------ */
int uco_b200_marker_map_write(const uco_marker_stream* m, uint32_t n, uint8_t* out, size_t cap, size_t* written) {
    if (!written || (n && !m)) return UCO_E_INVALID;
    Writer W{out, cap};
    W.val(n);
    for (uint32_t i = 0; i < n; i++) {                       // the caller passes the markers in ascending key order (std::map iteration)
        if (i && m[i].key <= m[i - 1].key) return UCO_E_INVALID;
        W.val(m[i].key); W.val(m[i].id);
        W.val<uint32_t>(928511272u); W.put(m[i].pose_g2m, 64);
        W.val(m[i].size);
        W.val(m[i].n_frames); W.put(m[i].frames, 4 * (size_t)m[i].n_frames);
        W.val(m[i].dict_len); W.put(m[i].dict, m[i].dict_len);
    }
    *written = W.n;
    return out && W.n > cap ? UCO_E_CAPACITY : UCO_OK;
}
/* lists in CSR form: rec_key[i] owns values[ptr[i] .. ptr[i+1]) */
static void write_keyed_sets(Writer& W, uint32_t n, const uint32_t* key, const uint32_t* ptr, const uint32_t* values) {
    W.val(n);
    for (uint32_t i = 0; i < n; i++) {
        W.val(key[i]);
        W.val<uint32_t>(ptr[i + 1] - ptr[i]);
        W.put(values + ptr[i], 4 * (size_t)(ptr[i + 1] - ptr[i]));
    }
}
static bool read_keyed_sets(const uint8_t* p, const uint8_t* end, uint32_t n, uint32_t* key, uint32_t* ptr, uint32_t* values) {
    Reader R{p, end};
    uint32_t at = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t k = R.get<uint32_t>(), c = R.get<uint32_t>();
        const uint8_t* v = R.skip(4 * (size_t)c);
        if (!R.ok) return false;
        if (key) key[i] = k;
        if (ptr) ptr[i] = at;
        if (values) memcpy(values + at, v, 4 * (size_t)c);
        at += c;
    }
    if (ptr) ptr[n] = at;
    return true;
}
int uco_b200_covis_stream_unpack(const uint8_t* bytes, size_t len, const uco_covis_stream* v, uint32_t* adj_node, uint32_t* adj_ptr, uint32_t* adj_idx,
                                 uint64_t* w_key, float* w) {
    if (!bytes || !v || v->adj_off > len) return UCO_E_INVALID;
    if (!read_keyed_sets(bytes + v->adj_off, bytes + len, v->n_adj, adj_node, adj_ptr, adj_idx)) return UCO_E_INVALID;
    for (uint32_t i = 0; i < v->n_weights; i++) {
        if (w_key) memcpy(w_key + i, v->weights + 12 * (size_t)i, 8);
        if (w) memcpy(w + i, v->weights + 12 * (size_t)i + 8, 4);
    }
    return UCO_OK;
}
int uco_b200_covis_stream_write(uint32_t n_nodes, const uint32_t* nodes, uint32_t n_adj, const uint32_t* adj_node, const uint32_t* adj_ptr, const uint32_t* adj_idx,
                                uint32_t n_weights, const uint64_t* w_key, const float* w, uint8_t* out, size_t cap, size_t* written) {
    if (!written || (n_nodes && !nodes) || (n_adj && (!adj_node || !adj_ptr || !adj_idx)) || (n_weights && (!w_key || !w))) return UCO_E_INVALID;
    Writer W{out, cap};
    W.val(n_nodes); W.put(nodes, 4 * (size_t)n_nodes);
    write_keyed_sets(W, n_adj, adj_node, adj_ptr, adj_idx);
    W.val(n_weights);
    for (uint32_t i = 0; i < n_weights; i++) { W.val(w_key[i]); W.val(w[i]); }
    *written = W.n;
    return out && W.n > cap ? UCO_E_CAPACITY : UCO_OK;
}
int uco_b200_kfdb_stream_unpack(const uint8_t* bytes, size_t len, const uco_kfdb_stream* v, uint32_t* word, uint32_t* word_ptr, uint32_t* word_frames) {
    if (!bytes || !v || v->words_off > len) return UCO_E_INVALID;
    return read_keyed_sets(bytes + v->words_off, bytes + len, v->n_words, word, word_ptr, word_frames) ? UCO_OK : UCO_E_INVALID;
}
int uco_b200_kfdb_stream_write(int32_t type, const uint8_t* voc, size_t voc_len, uint32_t n_words, const uint32_t* word, const uint32_t* word_ptr,
                               const uint32_t* word_frames, uint32_t n_frames, const uint32_t* frames, uint8_t* out, size_t cap, size_t* written) {
    if (!written || (type != 0 && type != 1) || (type == 1 && (!voc || voc_len < 128)) || (n_words && (!word || !word_ptr || !word_frames)) || (n_frames && !frames))
        return UCO_E_INVALID;
    Writer W{out, cap};
    W.val(type);
    if (type == 1) {
        W.put(voc, voc_len);
        write_keyed_sets(W, n_words, word, word_ptr, word_frames);
    }
    W.val(n_frames); W.put(frames, 4 * (size_t)n_frames);
    *written = W.n;
    return out && W.n > cap ? UCO_E_CAPACITY : UCO_OK;
}

/* Map::toStream (src/map.cpp:316-325): keyframe database, map points, markers, keyframes, covisibility graph - in that order; a map FILE
 * (Map::saveToFile, :339-345) puts the u64 magic 225237123 in front (has_file_magic != 0). */
int uco_b200_map_stream_walk(const uint8_t* bytes, size_t len, int has_file_magic, uco_map_sections* o) {
    if (!bytes || !o) return UCO_E_INVALID;
    memset(o, 0, sizeof *o);
    size_t at = 0, used = 0;
    if (has_file_magic) {
        uint64_t sig = 0;
        if (len < 8) return UCO_E_INVALID;
        memcpy(&sig, bytes, 8);
        if (sig != 225237123ull) return UCO_E_INVALID;
        at = 8;
    }
    int rc;
    o->kfdb_off = at;
    if ((rc = uco_b200_kfdb_stream_walk(bytes + at, len - at, &o->kfdb, &used)) != UCO_OK) return rc;
    at += used; o->kfdb_len = used;
    o->points_off = at;
    if ((rc = uco_b200_mappoint_container_walk(bytes + at, len - at, &o->points, nullptr, nullptr, 0, &used)) != UCO_OK) return rc;
    at += used; o->points_len = used;
    o->markers_off = at;
    if ((rc = uco_b200_marker_map_walk(bytes + at, len - at, 0, nullptr, &o->n_markers, &used)) != UCO_OK) return rc;
    at += used; o->markers_len = used;
    o->frames_off = at;
    if ((rc = uco_b200_frame_container_walk(bytes + at, len - at, &o->frames, nullptr, nullptr, 0, &used)) != UCO_OK) return rc;
    at += used; o->frames_len = used;
    o->covis_off = at;
    if ((rc = uco_b200_covis_stream_walk(bytes + at, len - at, &o->covis, &used)) != UCO_OK) return rc;
    at += used; o->covis_len = used;
    o->total_len = at;
    return UCO_OK;
}

/* KdTreeIndex::toStream bytes from the flattened tree (the inverse of uco_b200_kdtree_parse): what a frame whose tree was built on
 * the device (uco_b200_kdtree_build_batch_dev) writes into its stream.  div_val: the node's split value (picoflann keeps it as a
 * double next to the two float sides); pass NULL to write (divlow + divhigh) / 2. */
int uco_b200_kdtree_serialize(const uco_kdnode* nodes, int n_nodes, const int32_t* leaf_idx, const double* bbox4, int n_values, const double* div_val,
                              uint8_t* out, size_t cap, size_t* written) {
    if (!written || n_nodes < 0 || (n_nodes && (!nodes || !leaf_idx || !bbox4))) return UCO_E_INVALID;
    Writer W{out, cap};
    W.val<int32_t>(2); W.val<int32_t>(n_values);
    W.val<uint64_t>(2);
    double bb[4] = {0, 0, 0, 0};
    if (bbox4) memcpy(bb, bbox4, 32);
    W.put(bb, 32);
    W.val<uint64_t>((uint64_t)n_nodes);
    for (int i = 0; i < n_nodes; i++) {
        const uco_kdnode& N = nodes[i];
        const bool leaf = N.col < 0;
        W.val<double>(leaf ? 0.0 : (div_val ? div_val[i] : 0.5 * ((double)N.divlow + (double)N.divhigh)));
        W.val<uint16_t>(leaf ? 0 : (uint16_t)N.col);
        W.val<float>(leaf ? 0.f : N.divhigh); W.val<float>(leaf ? 0.f : N.divlow);
        W.val<int64_t>(leaf ? -1 : N.left); W.val<int64_t>(leaf ? -1 : N.right);
        W.val<uint64_t>(leaf ? (uint64_t)N.leaf_count : 0);
        if (leaf) W.put(leaf_idx + N.leaf_begin, 4 * (size_t)N.leaf_count);
    }
    *written = W.n;
    if (out && W.n > cap) return UCO_E_CAPACITY;
    return UCO_OK;
}

}  // extern "C"

// ---- device mirror ---------------------------------------------------------------------------------------------------------------------
struct uco_b200_frame {
    uco_b200_ctx* ctx;
    uint8_t* block = nullptr;      // one allocation
    uco_frame_dev d;
};

extern "C" {

int uco_b200_frame_upload(uco_b200_ctx* ctx, const uco_frame_stream* v, uco_b200_frame** out) {
    UCO_RANGE();
    if (!ctx) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!v || !out) return uco_fail(ctx, UCO_E_INVALID, "frame_upload: null argument");
    *out = nullptr;
    const uint32_t n = v->n_und_kpts;
    if (v->desc.data && (v->desc.rows != (int)n || v->desc.cols != 32 || elem_size(v->desc.type) != 1))
        return uco_fail(ctx, UCO_E_INVALID, "frame_upload: %d x %d descriptors of type %d for %u keypoints (expected n x 32 bytes: ORB)", v->desc.rows, v->desc.cols, v->desc.type, n);
    if ((v->n_ids && v->n_ids != n) || (v->n_flags && v->n_flags != n) || (v->n_depth && v->n_depth != n))
        return uco_fail(ctx, UCO_E_INVALID, "frame_upload: per-keypoint arrays disagree with %u keypoints", n);
    // the kd-tree of the stream, flattened
    std::vector<uco_kdnode> nodes(2 * (size_t)n + 4);
    std::vector<int32_t> leaf(n + 1);
    double bbox[4] = {0, 0, 0, 0};
    int n_nodes = 0, n_leaf = 0;
    if (v->kdtree && v->kdtree_bytes) {
        const int rc = uco_b200_kdtree_parse(v->kdtree, (size_t)v->kdtree_bytes, nodes.data(), (int)nodes.size(), leaf.data(), (int)leaf.size(), bbox, &n_nodes, &n_leaf);
        if (rc != UCO_OK) return uco_fail(ctx, UCO_E_INVALID, "frame_upload: malformed kd-tree stream");
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t b) { const size_t o = off; off = al(off + b); return o; };
    const size_t o_kp = take(sizeof(uco_keypoint) * (size_t)n), o_desc = take(32 * (size_t)n), o_ids = take(4 * (size_t)n), o_fl = take(n),
                 o_dep = take(4 * (size_t)n), o_nodes = take(sizeof(uco_kdnode) * (size_t)n_nodes), o_leaf = take(4 * (size_t)n_leaf),
                 o_sf = take(4 * (size_t)v->n_scale_factors), o_misc = take(256);
    uco_b200_frame* f = new uco_b200_frame();
    f->ctx = ctx;
    if (cudaMalloc(&f->block, off) != cudaSuccess) { delete f; return uco_fail(ctx, UCO_E_NOMEM, "frame_upload: cudaMalloc of %zu bytes failed", off); }
    uint8_t* h = (uint8_t*)uco_pinned(ctx, WS_GENERIC1, off);
    if (!h) { cudaFree(f->block); delete f; return UCO_E_NOMEM; }
    memset(h, 0, off);
    if (n) memcpy(h + o_kp, v->und_kpts, sizeof(uco_keypoint) * (size_t)n);
    if (n && v->desc.data) memcpy(h + o_desc, v->desc.data, 32 * (size_t)n);
    if (v->n_ids) memcpy(h + o_ids, v->ids, 4 * (size_t)n); else memset(h + o_ids, 0xff, 4 * (size_t)n);   // no map point: uint32 max
    if (v->n_flags) memcpy(h + o_fl, v->flags, n);
    if (v->n_depth) memcpy(h + o_dep, v->depth, 4 * (size_t)n);
    if (n_nodes) memcpy(h + o_nodes, nodes.data(), sizeof(uco_kdnode) * (size_t)n_nodes);
    if (n_leaf) memcpy(h + o_leaf, leaf.data(), 4 * (size_t)n_leaf);
    if (v->n_scale_factors) memcpy(h + o_sf, v->scale_factors, 4 * (size_t)v->n_scale_factors);
    memcpy(h + o_misc, v->pose_f2g, 64);
    memcpy(h + o_misc + 64, bbox, 32);
    cudaError_t e = cudaMemcpyAsync(f->block, h, off, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(f->block); delete f; return uco_fail(ctx, UCO_E_CUDA, "frame_upload: %s", cudaGetErrorString(e)); }
    uco_frame_dev& d = f->d;
    memset(&d, 0, sizeof d);
    d.idx = v->idx; d.fseq_idx = v->fseq_idx; d.n_kp = (int32_t)n;
    d.kps = (const uco_keypoint*)(f->block + o_kp); d.desc = f->block + o_desc; d.ids = (const uint32_t*)(f->block + o_ids);
    d.flags = f->block + o_fl; d.depth = v->n_depth ? (const float*)(f->block + o_dep) : nullptr;
    d.n_nodes = n_nodes; d.nodes = (const uco_kdnode*)(f->block + o_nodes); d.n_leaf = n_leaf; d.leaf_idx = (const int32_t*)(f->block + o_leaf);
    d.n_scale_factors = (int32_t)v->n_scale_factors; d.scale_factors = (const float*)(f->block + o_sf);
    d.pose_f2g = (const float*)(f->block + o_misc); d.bbox = (const double*)(f->block + o_misc + 64);
    memcpy(d.bbox_host, bbox, 32);
    memcpy(d.pose_host, v->pose_f2g, 64);
    if (v->camera_matrix.data && v->camera_matrix.rows == 3 && v->camera_matrix.cols == 3 && v->camera_matrix.type == 5) {
        const float* K = (const float*)v->camera_matrix.data;
        d.K[0] = K[0]; d.K[1] = K[4]; d.K[2] = K[2]; d.K[3] = K[5];
    }
    d.min_xy[0] = v->min_xy[0]; d.min_xy[1] = v->min_xy[1]; d.max_xy[0] = v->max_xy[0]; d.max_xy[1] = v->max_xy[1];
    *out = f;
    return UCO_OK;
}

const uco_frame_dev* uco_b200_frame_dev(const uco_b200_frame* f) { return f ? &f->d : nullptr; }

/* the per-keypoint arrays of the mirror back to the host (after a device-side stage changed them: map-point ids, flags); any pointer may be NULL */
int uco_b200_frame_download(uco_b200_ctx* ctx, const uco_b200_frame* f, uco_keypoint* kps, uint8_t* desc, uint32_t* ids, uint8_t* flags, float* depth) {
    if (!ctx || !f) return UCO_E_INVALID;
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)f->d.n_kp;
    cudaStream_t s = ctx->stream;
    if (kps && n) UCO_CUDA(ctx, cudaMemcpyAsync(kps, f->d.kps, sizeof(uco_keypoint) * n, cudaMemcpyDeviceToHost, s));
    if (desc && n) UCO_CUDA(ctx, cudaMemcpyAsync(desc, f->d.desc, 32 * n, cudaMemcpyDeviceToHost, s));
    if (ids && n) UCO_CUDA(ctx, cudaMemcpyAsync(ids, f->d.ids, 4 * n, cudaMemcpyDeviceToHost, s));
    if (flags && n) UCO_CUDA(ctx, cudaMemcpyAsync(flags, f->d.flags, n, cudaMemcpyDeviceToHost, s));
    if (depth && n && f->d.depth) UCO_CUDA(ctx, cudaMemcpyAsync(depth, f->d.depth, 4 * n, cudaMemcpyDeviceToHost, s));
    UCO_CUDA(ctx, cudaStreamSynchronize(s));
    return UCO_OK;
}

void uco_b200_frame_free(uco_b200_frame* f) {
    if (!f) return;
    if (f->block) { cudaSetDevice(f->ctx->device); cudaFree(f->block); }
    delete f;
}

}  // extern "C"
