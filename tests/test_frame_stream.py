"""Frame streams: csrc/frame_stream.cu against the reference's OWN Frame::toStream / fromStream statements
(/root/reference/src/map_types/frame.cpp:260-341 and the streams of its members), compiled by oracle/Makefile into
oracle/_ref/libref_frame.so on container stand-ins, and against the committed golden stream tests/golden/frame_stream.bin written
by that library (tests/golden/make_frame_golden.py).  Bit-exact: these are bytes."""
import ctypes, os
import numpy as np
import pytest
import ucoslam_b200
from ucoslam_b200 import frame_stream_parse, frame_stream_write, view_array, KP_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "libref_frame.so")


def make_fields(seed, n_kp=300, n_markers=2, img=(12, 16), depth=True):
    rng = np.random.default_rng(seed)
    kp = np.zeros(n_kp, KP_DTYPE)
    kp["x"], kp["y"] = rng.uniform(0, 640, n_kp), rng.uniform(0, 480, n_kp)
    kp["size"], kp["angle"], kp["response"] = 31, rng.uniform(0, 360, n_kp), rng.integers(5, 200, n_kp)
    kp["octave"], kp["class_id"] = rng.integers(0, 8, n_kp), -1
    ids = np.where(rng.random(n_kp) < 0.4, rng.integers(0, 5000, n_kp), 0xFFFFFFFF).astype(np.uint32)
    node_kp = rng.permutation(n_kp).astype(np.uint32)
    cuts = np.sort(rng.choice(np.arange(1, max(n_kp, 2)), min(9, max(n_kp - 1, 0)), replace=False)) if n_kp > 1 else np.zeros(0, int)
    node_ptr = np.concatenate([[0], cuts, [n_kp]]).astype(np.int32)
    pose = np.eye(4, dtype=np.float32); pose[:3, 3] = rng.normal(0, 1, 3)
    return dict(idx=7, fseq_idx=123, frame_flags=4, kp=kp, desc=rng.integers(0, 256, (n_kp, 32), dtype=np.uint8),
                kpts=np.c_[kp["x"], kp["y"]].astype(np.float32) + np.float32(0.25), depth=rng.uniform(0.5, 5, n_kp if depth else 0).astype(np.float32),
                ids=ids, flags=rng.integers(0, 4, n_kp).astype(np.uint8), marker_id=rng.integers(0, 250, n_markers).astype(np.int32),
                marker_f=rng.uniform(0, 400, (n_markers, 17)).astype(np.float32), marker_d=rng.normal(0, 1, (n_markers, 19)),
                pose=pose, bow_word=np.sort(rng.choice(100000, 60, replace=False)).astype(np.uint32), bow_weight=rng.random(60).astype(np.float32),
                node_id=np.sort(rng.choice(1000, len(node_ptr) - 1, replace=False)).astype(np.uint32), node_ptr=node_ptr, node_kp=node_kp,
                sf=(np.float32(1.2) ** np.arange(8)).astype(np.float32), K=np.array([[525, 0, 319.5], [0, 525, 239.5], [0, 0, 1]], np.float32),
                dist=rng.normal(0, 0.01, 5).astype(np.float32), cam=(640, 480), bl=0.12, depthscale=0.001,
                img=rng.integers(0, 256, img, dtype=np.uint8), min_xy=(-3, -2), max_xy=(644, 483))


def ref_stream(fd, build_tree=True):
    lib = ctypes.CDLL(REF)
    lib.ref_frame_to_stream.restype = ctypes.c_long
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(ctypes.c_void_p)
    out = np.zeros(4 << 20, np.uint8)
    keep = [np.ascontiguousarray(fd[k]) for k in ("kp", "desc", "kpts", "depth", "ids", "flags", "marker_id", "marker_f", "marker_d", "pose", "bow_word",
                                                    "bow_weight", "node_id", "node_ptr", "node_kp", "sf", "K", "dist", "img")]
    A = [ctypes.c_void_p(a.ctypes.data) for a in keep]
    n = lib.ref_frame_to_stream(ctypes.c_uint32(fd["idx"]), ctypes.c_uint32(fd["fseq_idx"]), ctypes.c_ubyte(fd["frame_flags"]), len(fd["kp"]), A[0], A[1], A[2],
                                len(fd["depth"]), A[3], A[4], A[5], len(fd["marker_id"]), A[6], A[7], A[8], A[9], len(fd["bow_word"]), A[10], A[11],
                                len(fd["node_id"]), A[12], A[13], A[14], len(fd["sf"]), A[15], A[16], len(fd["dist"]), A[17], fd["cam"][0], fd["cam"][1],
                                ctypes.c_float(fd["bl"]), ctypes.c_float(fd["depthscale"]), fd["img"].shape[0], fd["img"].shape[1], A[18], int(build_tree),
                                fd["min_xy"][0], fd["min_xy"][1], fd["max_xy"][0], fd["max_xy"][1], ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(out)))
    assert n > 0
    return out[:n].copy()


def ref_roundtrip(buf):
    lib = ctypes.CDLL(REF)
    lib.ref_frame_roundtrip.restype = ctypes.c_long
    out = np.zeros(len(buf) + 1024, np.uint8)
    n = lib.ref_frame_roundtrip(ctypes.c_void_p(buf.ctypes.data), ctypes.c_long(len(buf)), ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(out)))
    return None if n < 0 else out[:n].copy()


def check_view(v, fd):
    n = len(fd["kp"])
    assert (v.idx, v.fseq_idx, v.frame_flags, v.kp_desc_type) == (fd["idx"], fd["fseq_idx"], fd["frame_flags"], 1)
    assert (v.desc.rows, v.desc.cols, v.desc.type) == ((n, 32, 0) if n else (0, 0, 0))
    assert np.array_equal(view_array(v.desc.data, 32 * n, np.uint8).reshape(n, 32), fd["desc"])
    assert v.n_und_kpts == n and view_array(v.und_kpts, n, KP_DTYPE).tobytes() == fd["kp"].tobytes()
    assert np.array_equal(view_array(v.kpts, 2 * v.n_kpts, np.float32).reshape(-1, 2), fd["kpts"])
    assert np.array_equal(view_array(v.depth, v.n_depth, np.float32), fd["depth"])
    assert np.array_equal(view_array(v.ids, v.n_ids, np.uint32), fd["ids"])
    assert np.array_equal(view_array(v.flags, v.n_flags, np.uint8), fd["flags"])
    assert v.n_markers == len(fd["marker_id"])
    assert np.array_equal(np.array(v.pose_f2g[:], np.float32).reshape(4, 4), fd["pose"])
    bow = view_array(v.bow, v.n_bow, np.dtype([("w", "<u4"), ("v", "<f4")]))
    assert np.array_equal(bow["w"], fd["bow_word"]) and np.array_equal(bow["v"], fd["bow_weight"])
    assert v.n_bow_level == len(fd["node_id"])
    lv = view_array(v.bow_level, v.bow_level_bytes // 4, np.uint32)
    at = 0
    for k in range(v.n_bow_level):
        cnt = int(lv[at + 1])
        assert lv[at] == fd["node_id"][k] and np.array_equal(lv[at + 2:at + 2 + cnt], fd["node_kp"][fd["node_ptr"][k]:fd["node_ptr"][k + 1]])
        at += 2 + cnt
    assert np.array_equal(view_array(v.scale_factors, v.n_scale_factors, np.float32), fd["sf"])
    assert np.array_equal(view_array(v.camera_matrix.data, 9, np.float32).reshape(3, 3), fd["K"])
    assert np.array_equal(view_array(v.distortion.data, v.distortion.cols, np.float32), fd["dist"])
    assert tuple(v.cam_size[:]) == fd["cam"] and v.bl == np.float32(fd["bl"]) and v.rgb_depthscale == np.float32(fd["depthscale"])
    assert (v.image.rows, v.image.cols) == fd["img"].shape and np.array_equal(view_array(v.image.data, fd["img"].size, np.uint8).reshape(fd["img"].shape), fd["img"])
    assert tuple(v.min_xy[:]) == fd["min_xy"] and tuple(v.max_xy[:]) == fd["max_xy"]


needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_frame.so not built (python __graft_entry__.py where /root/reference exists)")


@needs_ref
@pytest.mark.parametrize("kw", [dict(seed=1), dict(seed=2, n_kp=2000, n_markers=0, img=(0, 0)), dict(seed=3, n_kp=1, n_markers=5), dict(seed=4, n_kp=0, depth=False),
                                dict(seed=5, n_kp=700, depth=False)])
def test_parse_and_write_against_the_reference_statements(kw):
    fd = make_fields(**kw)
    ref = ref_stream(fd)
    v, used = frame_stream_parse(ref)
    assert used == len(ref)
    check_view(v, fd)
    assert np.array_equal(frame_stream_write(v), ref)                       # byte for byte what the reference wrote
    # the kd-tree member parses with the projection matcher's reader and agrees with the library's own build of the same points
    if len(fd["kp"]) > 10:
        lib = ucoslam_b200.load()
        n = len(fd["kp"])
        nodes, leaf, bbox = np.zeros((2 * n + 4, 7), np.int32), np.zeros(n + 1, np.int32), np.zeros(4)
        nn, nl = ctypes.c_int(), ctypes.c_int()
        assert lib.uco_b200_kdtree_parse(v.kdtree, v.kdtree_bytes, nodes.ctypes.data, len(nodes), leaf.ctypes.data, len(leaf), bbox.ctypes.data,
                                         ctypes.addressof(nn), ctypes.addressof(nl)) == 0
        nodes2, leaf2, bbox2, n2 = np.zeros_like(nodes), np.zeros_like(leaf), np.zeros(4), ctypes.c_int()
        xy = np.ascontiguousarray(np.c_[fd["kp"]["x"], fd["kp"]["y"]], np.float32)
        assert lib.uco_b200_kdtree_build(xy.ctypes.data, 8, n, nodes2.ctypes.data, len(nodes2), leaf2.ctypes.data, bbox2.ctypes.data, ctypes.addressof(n2)) == 0
        k = nn.value
        assert k == n2.value and nl.value == n and np.array_equal(nodes[:k, :5], nodes2[:k, :5]) and np.array_equal(nodes[:k, 6], nodes2[:k, 6])
        for a, b in zip(nodes[:k], nodes2[:k]):      # the two flatten the leaves' index lists in different orders; the lists agree
            assert np.array_equal(leaf[a[5]:a[5] + a[6]], leaf2[b[5]:b[5] + b[6]])
        assert np.array_equal(bbox, bbox2)
        # ... and serialises back to a stream the reader takes to the same tree
        cap = int(v.kdtree_bytes) + 64
        out, wn = np.zeros(cap, np.uint8), ctypes.c_size_t()
        assert lib.uco_b200_kdtree_serialize(nodes.ctypes.data, nn.value, leaf.ctypes.data, bbox.ctypes.data, n, None, out.ctypes.data, cap, ctypes.addressof(wn)) == 0
        assert wn.value == v.kdtree_bytes
        nodes3, leaf3, bbox3 = np.zeros_like(nodes), np.zeros_like(leaf), np.zeros(4)
        assert lib.uco_b200_kdtree_parse(out.ctypes.data, wn.value, nodes3.ctypes.data, len(nodes3), leaf3.ctypes.data, len(leaf3), bbox3.ctypes.data,
                                         ctypes.addressof(nn), ctypes.addressof(nl)) == 0
        assert np.array_equal(nodes3, nodes) and np.array_equal(leaf3, leaf) and np.array_equal(bbox3, bbox)


@needs_ref
def test_the_reference_reads_what_the_codec_writes():
    """fields changed through the view (new map-point ids, a new pose, no image: what a mapper does to a keyframe) -> the reference's
    fromStream accepts the bytes and its toStream gives them back unchanged"""
    fd = make_fields(6, n_kp=400)
    v, _ = frame_stream_parse(ref_stream(fd))
    ids = (np.arange(400, dtype=np.uint32) * 3)
    v.ids = ids.ctypes.data
    pose = np.eye(4, dtype=np.float32); pose[0, 3] = 2.5
    v.pose_f2g = (ctypes.c_float * 16)(*pose.reshape(-1))
    v.image.data = None
    mine = frame_stream_write(v)
    back = ref_roundtrip(mine)
    assert back is not None and np.array_equal(back, mine)
    fd2 = dict(fd, ids=ids, pose=pose, img=np.zeros((0, 0), np.uint8))
    assert np.array_equal(ref_stream(fd2), mine)


def test_golden_stream():
    """the committed stream (written by the reference's statements, tests/golden/make_frame_golden.py) parses to the committed fields
    and is reproduced byte for byte — runs where /root/reference does not exist"""
    ref = np.fromfile(os.path.join(GOLD, "frame_stream.bin"), np.uint8)
    fd = make_fields(1)
    v, used = frame_stream_parse(ref)
    assert used == len(ref)
    check_view(v, fd)
    assert np.array_equal(frame_stream_write(v), ref)


def test_malformed_streams_are_rejected():
    ref = np.fromfile(os.path.join(GOLD, "frame_stream.bin"), np.uint8)
    for cut in (0, 3, 17, 200, len(ref) - 1):
        with pytest.raises(ucoslam_b200.UcoError):
            frame_stream_parse(ref[:cut].copy() if cut else np.zeros(4, np.uint8))
    bad = ref.copy(); bad[0] ^= 1
    with pytest.raises(ucoslam_b200.UcoError):
        frame_stream_parse(bad)
    bad = ref.copy(); bad[-1] ^= 1                                            # closing magic
    with pytest.raises(ucoslam_b200.UcoError):
        frame_stream_parse(bad)


@pytest.mark.gpu
def test_device_mirror_chains_into_the_matcher():
    """two keyframes loaded from streams live on the device; the device-pointer matcher on the mirrors gives what the host-buffer
    matcher gives on the same frames; per-keypoint arrays come back unchanged"""
    import torch
    ctx = ucoslam_b200.Context(0)
    ref = np.fromfile(os.path.join(GOLD, "frame_stream.bin"), np.uint8)
    fd = make_fields(1)
    va, _ = frame_stream_parse(ref)
    # second frame: the same keypoints seen again (descriptor noise, small motion) through the writer
    rng = np.random.default_rng(9)
    kp2 = fd["kp"].copy(); kp2["x"] += rng.normal(0, 1, len(kp2)).astype(np.float32); kp2["y"] += rng.normal(0, 1, len(kp2)).astype(np.float32)
    d2 = fd["desc"].copy(); fl = rng.integers(0, 256, (len(kp2), 8))
    for j in range(8):
        d2[np.arange(len(kp2)), fl[:, j] >> 3] ^= (1 << (fl[:, j] & 7)).astype(np.uint8)
    vb, _ = frame_stream_parse(ref)
    vb.und_kpts, vb.desc.data, vb.kdtree, vb.kdtree_bytes = kp2.ctypes.data, d2.ctypes.data, None, 0
    ha, da = ctx.frame_upload(va)
    hb, db = ctx.frame_upload(vb)
    n = len(kp2)
    assert da.n_kp == n and db.n_kp == n and da.n_nodes > 0 and db.n_nodes == 0
    kps, desc, ids, flags, depth = ctx.frame_download(ha, n)
    assert kps.tobytes() == fd["kp"].tobytes() and np.array_equal(desc, fd["desc"]) and np.array_equal(ids, fd["ids"]) and np.array_equal(flags, fd["flags"])
    assert np.array_equal(depth, fd["depth"])
    prm = ucoslam_b200.MatchParams(80.0, 0.8, True, 1)
    want = ctx.frame_match(d2, kp2, fd["desc"], fd["kp"], prm)
    out = torch.zeros((n, 16), dtype=torch.uint8, device="cuda"); n_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    ctx.frame_match_batch_dev(1, db.desc, 0, db.kps, 0, n, None, da.desc, 0, da.kps, 0, n, None, prm, out.data_ptr(), n_out.data_ptr())
    ctx.sync()
    got = out.cpu().numpy().view(ucoslam_b200.MATCH_DTYPE).reshape(-1)[:int(n_out.item())]
    assert len(want) > n // 2 and got.tobytes() == want.tobytes()
    ctx.frame_free(ha); ctx.frame_free(hb)
    ctx.close()


# ---- MapPoint streams (the other record type of a map file) ---------------------------------------------------------------------------
def _mp_fields(seed, n_frames=5, with_desc=True):
    rng = np.random.default_rng(seed)
    fr = np.sort(rng.choice(500, n_frames, replace=False)).astype(np.uint32)
    return dict(id=int(rng.integers(0, 1 << 20)), pos=rng.normal(0, 2, 3).astype(np.float32), desc=rng.integers(0, 256, 32, dtype=np.uint8) if with_desc else None,
                frames=np.c_[fr, rng.integers(0, 2000, n_frames)].astype(np.uint32), normal=rng.normal(0, 1, 3).astype(np.float32), seen=int(rng.integers(0, 60000)),
                visible=int(rng.integers(0, 60000)), flags=int(rng.integers(0, 8)), maxd=float(np.float32(rng.uniform(1, 9))), mind=float(np.float32(rng.uniform(0.1, 1))),
                kf_since=int(rng.integers(0, 1 << 40)), last_seen=int(rng.integers(0, 1 << 30)))


def _ref_mappoint(fd):
    lib = ctypes.CDLL(REF)
    lib.ref_mappoint_to_stream.restype = ctypes.c_long
    out = np.zeros(4096, np.uint8)
    fr = np.ascontiguousarray(fd["frames"])
    n = lib.ref_mappoint_to_stream(ctypes.c_uint32(fd["id"]), ctypes.c_void_p(fd["pos"].ctypes.data), None if fd["desc"] is None else ctypes.c_void_p(fd["desc"].ctypes.data),
                                   len(fr), ctypes.c_void_p(fr.ctypes.data), ctypes.c_void_p(fd["normal"].ctypes.data), fd["seen"], fd["visible"],
                                   ctypes.c_ubyte(fd["flags"]), ctypes.c_float(fd["maxd"]), ctypes.c_float(fd["mind"]), ctypes.c_ulonglong(fd["kf_since"]),
                                   ctypes.c_uint32(fd["last_seen"]), ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(out)))
    assert n > 0
    return out[:n].copy()


@needs_ref
@pytest.mark.parametrize("kw", [dict(seed=1), dict(seed=2, n_frames=0), dict(seed=3, n_frames=40, with_desc=False)])
def test_mappoint_stream_against_the_reference_statements(kw):
    fd = _mp_fields(**kw)
    ref = _ref_mappoint(fd)
    lib = ucoslam_b200.load()
    v, used = ucoslam_b200.MapPointStream(), ctypes.c_size_t()
    assert lib.uco_b200_mappoint_stream_parse(ref.ctypes.data, len(ref), ctypes.addressof(v), ctypes.addressof(used)) == 0 and used.value == len(ref)
    assert v.id == fd["id"] and np.array_equal(np.array(v.pos3d[:], np.float32), fd["pos"]) and np.array_equal(np.array(v.normal[:], np.float32), fd["normal"])
    assert np.array_equal(view_array(v.frames, 2 * v.n_frames, np.uint32).reshape(-1, 2), fd["frames"])
    if fd["desc"] is not None:
        assert (v.desc.rows, v.desc.cols, v.desc.type) == (1, 32, 0) and np.array_equal(view_array(v.desc.data, 32, np.uint8), fd["desc"])
    else:
        assert v.desc.rows == 0 and not v.desc.data
    assert (v.n_times_seen, v.n_times_visible, v.flags, v.kf_since_addition, v.last_fidx_seen) == (fd["seen"], fd["visible"], fd["flags"], fd["kf_since"], fd["last_seen"])
    assert v.max_distance == np.float32(fd["maxd"]) and v.min_distance == np.float32(fd["mind"])
    out, n = np.zeros(len(ref) + 8, np.uint8), ctypes.c_size_t()
    assert lib.uco_b200_mappoint_stream_write(ctypes.addressof(v), out.ctypes.data, len(out), ctypes.addressof(n)) == 0
    assert np.array_equal(out[:n.value], ref)
    # a moved point written by the codec is read back by the reference unchanged
    v.pos3d = (ctypes.c_float * 3)(1.5, -2.0, 0.25)
    assert lib.uco_b200_mappoint_stream_write(ctypes.addressof(v), out.ctypes.data, len(out), ctypes.addressof(n)) == 0
    rlib = ctypes.CDLL(REF); rlib.ref_mappoint_roundtrip.restype = ctypes.c_long
    back = np.zeros(len(out) + 64, np.uint8)
    m = rlib.ref_mappoint_roundtrip(ctypes.c_void_p(out.ctypes.data), ctypes.c_long(n.value), ctypes.c_void_p(back.ctypes.data), ctypes.c_long(len(back)))
    assert m == n.value and np.array_equal(back[:m], out[:m])
    assert np.array_equal(_ref_mappoint(dict(fd, pos=np.array([1.5, -2.0, 0.25], np.float32))), out[:m])


def test_mappoint_golden_and_malformed():
    ref = np.fromfile(os.path.join(GOLD, "mappoint_stream.bin"), np.uint8)
    fd = _mp_fields(1)
    lib = ucoslam_b200.load()
    v, used = ucoslam_b200.MapPointStream(), ctypes.c_size_t()
    assert lib.uco_b200_mappoint_stream_parse(ref.ctypes.data, len(ref), ctypes.addressof(v), ctypes.addressof(used)) == 0 and used.value == len(ref)
    assert v.id == fd["id"] and np.array_equal(view_array(v.frames, 2 * v.n_frames, np.uint32).reshape(-1, 2), fd["frames"])
    out, n = np.zeros(len(ref), np.uint8), ctypes.c_size_t()
    assert lib.uco_b200_mappoint_stream_write(ctypes.addressof(v), out.ctypes.data, len(out), ctypes.addressof(n)) == 0 and np.array_equal(out, ref)
    assert lib.uco_b200_mappoint_stream_parse(ref.ctypes.data, len(ref) - 3, ctypes.addressof(v), None) != 0
    bad = ref.copy(); bad[0] ^= 1
    assert lib.uco_b200_mappoint_stream_parse(bad.ctypes.data, len(bad), ctypes.addressof(v), None) != 0


# ---- the map-point SECTION of a map file: ReusableContainer<MapPoint> (Map::toStream, map.cpp:316-325) ------------------------------------
def _ref_container(points, erase=(), again=()):
    """the reference's own ReusableContainer / ExpansibleContainer headers (unchanged) driven as Map drives them: insert all of `points`, erase the slots
    in `erase` (in that order), insert `again` (they take the freed slots, last freed first); returns the bytes the reference writes"""
    lib = ctypes.CDLL(REF)
    lib.ref_mappoint_container.restype = ctypes.c_long
    streams = [_ref_mappoint(fd) for fd in list(points) + list(again)]
    blob = np.concatenate(streams)
    lens = np.array([len(s) for s in streams], np.int64)
    er = np.array(list(erase), np.uint32)
    out = np.zeros(len(blob) + 400 * 200 + 4096, np.uint8)
    n = lib.ref_mappoint_container(ctypes.c_void_p(blob.ctypes.data), ctypes.c_void_p(lens.ctypes.data), len(points), ctypes.c_void_p(er.ctypes.data), len(er),
                                   len(again), ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(out)))
    assert n > 0
    return out[:n].copy()


def _container_case():
    pts = [dict(_mp_fields(100 + i, n_frames=i % 7), id=i) for i in range(450)]      # three chunks of 200 slots, the last one half used
    erase = [7, 399, 0, 211, 449]
    again = [dict(_mp_fields(900 + i), id=1000 + i) for i in range(2)]               # reuse slots 449 and 211 (last freed first)
    return pts, erase, again


@needs_ref
def test_mappoint_container_against_the_reference_headers():
    from ucoslam_b200 import mappoint_container_walk, mappoints_from_container, MapPointStream, MapPointContainer
    pts, erase, again = _container_case()
    ref = _ref_container(pts, erase, again)
    lib = ucoslam_b200.load()
    c, off, valid, used = mappoint_container_walk(ref)
    assert used == len(ref) and (c.n_slots, c.n_used, c.n_valid, c.n_free) == (600, 450, 447, 3)
    assert list(np.ctypeslib.as_array(ctypes.cast(c.free_slots, ctypes.POINTER(ctypes.c_uint32)), (3,))) == [7, 399, 0]
    expect = {i: p for i, p in enumerate(pts)}
    expect[449], expect[211] = again[0], again[1]
    views = (MapPointStream * c.n_slots)()
    for i in range(c.n_slots):
        u = ctypes.c_size_t()
        assert lib.uco_b200_mappoint_stream_parse(ref.ctypes.data + int(off[i]), len(ref) - int(off[i]), ctypes.addressof(views[i]), ctypes.addressof(u)) == 0
        assert bool(valid[i]) == (i < 450 and i not in (7, 399, 0))
        if i < 450:
            assert views[i].id == expect[i]["id"] and np.allclose(list(views[i].pos3d), expect[i]["pos"])
        else:                                       # never used: what a default-constructed MapPoint streams as
            d = MapPointStream()
            lib.uco_b200_mappoint_stream_default(ctypes.addressof(d))
            assert (views[i].id, views[i].n_frames, views[i].max_distance, views[i].min_distance, views[i].last_fidx_seen) == \
                   (d.id, 0, d.max_distance, d.min_distance, d.last_fidx_seen) and views[i].desc.rows == 0
    # the valid points as flat rows, in slot order
    mp = mappoints_from_container(ref)
    order = [i for i in range(450) if i not in (7, 399, 0)]
    assert list(mp["ids"]) == [expect[i]["id"] for i in order]
    assert np.array_equal(mp["pos"], np.array([expect[i]["pos"] for i in order])) and np.array_equal(mp["desc"], np.array([expect[i]["desc"] for i in order]))
    assert np.array_equal(mp["normal"], np.array([expect[i]["normal"] for i in order]))
    assert np.array_equal(mp["max_dist"], np.array([expect[i]["maxd"] for i in order], np.float32)) and list(mp["flags"]) == [expect[i]["flags"] for i in order]
    # the writer reproduces the reference's bytes, and the reference reads them back
    out, n = np.zeros(len(ref) + 64, np.uint8), ctypes.c_size_t()
    assert lib.uco_b200_mappoint_container_write(ctypes.addressof(c), ctypes.addressof(views), valid.ctypes.data, out.ctypes.data, len(out), ctypes.addressof(n)) == 0
    assert n.value == len(ref) and np.array_equal(out[:n.value], ref)
    assert lib.uco_b200_mappoint_container_write(ctypes.addressof(c), ctypes.addressof(views), valid.ctypes.data, out.ctypes.data, 100, ctypes.addressof(n)) == -4
    rlib = ctypes.CDLL(REF)
    rlib.ref_mappoint_container_roundtrip.restype = ctypes.c_long
    back, nv = np.zeros(len(ref) + 64, np.uint8), ctypes.c_long()
    m = rlib.ref_mappoint_container_roundtrip(ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(ref)), ctypes.c_void_p(back.ctypes.data), ctypes.c_long(len(back)),
                                              ctypes.byref(nv))
    assert m == len(ref) and np.array_equal(back[:m], ref) and nv.value == 447


@needs_ref
@pytest.mark.parametrize("n", [0, 1, 200, 201, 400])
def test_mappoint_container_chunk_edges(n):
    """0 points, exactly one / two full chunks, one past a full chunk: curBuffer / curElm as ExpansibleContainer::push_back leaves them"""
    from ucoslam_b200 import mappoint_container_walk, MapPointStream
    pts = [dict(_mp_fields(300 + i, n_frames=1), id=i) for i in range(n)]
    if n == 0:
        lib = ctypes.CDLL(REF); lib.ref_mappoint_container.restype = ctypes.c_long
        out = np.zeros(200 * 80 + 256, np.uint8)
        k = lib.ref_mappoint_container(None, None, 0, None, 0, 0, ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(out)))
        ref = out[:k].copy()
    else:
        ref = _ref_container(pts)
    c, off, valid, used = mappoint_container_walk(ref)
    assert used == len(ref) and c.n_used == n and c.n_valid == n and c.n_slots == 200 * max(1, (n + 199) // 200) and int(valid.sum()) == n
    lib = ucoslam_b200.load()
    views = (MapPointStream * c.n_slots)()
    for i in range(c.n_slots):
        assert lib.uco_b200_mappoint_stream_parse(ref.ctypes.data + int(off[i]), len(ref) - int(off[i]), ctypes.addressof(views[i]), None) == 0
    out, w = np.zeros(len(ref), np.uint8), ctypes.c_size_t()
    assert lib.uco_b200_mappoint_container_write(ctypes.addressof(c), ctypes.addressof(views), valid.ctypes.data, out.ctypes.data, len(out), ctypes.addressof(w)) == 0
    assert np.array_equal(out, ref)


def test_mappoint_container_golden_and_malformed():
    from ucoslam_b200 import mappoint_container_walk, mappoints_from_container
    ref = np.fromfile(os.path.join(GOLD, "mappoint_container.bin"), np.uint8)
    c, off, valid, used = mappoint_container_walk(ref)
    assert used == len(ref) and (c.n_slots, c.n_used, c.n_valid, c.n_free) == (200, 6, 5, 1)
    mp = mappoints_from_container(ref)
    assert list(mp["ids"]) == [0, 1, 3, 4, 5] and mp["desc"].shape == (5, 32)
    for bad in (ref[:-5], ref[:40], np.concatenate([np.zeros(8, np.uint8), ref[8:]])):
        with pytest.raises(ucoslam_b200.UcoError):
            mappoint_container_walk(np.ascontiguousarray(bad))
    worse = ref.copy()
    worse[8 + 4 + 4 + 8] = 77          # the chunk count
    with pytest.raises(ucoslam_b200.UcoError):
        mappoint_container_walk(worse)


# ---- the keyframe SECTION of a map file: FrameSet (int 88888 + ReusableContainer<Frame>) -----------------------------------------------------
def _ref_frame_container(fields, erase=(), again=()):
    lib = ctypes.CDLL(REF)
    lib.ref_frame_container.restype = ctypes.c_long
    streams = [ref_stream(fd) for fd in list(fields) + list(again)]
    blob = np.concatenate(streams)
    lens = np.array([len(s) for s in streams], np.int64)
    er = np.array(list(erase), np.uint32)
    out = np.zeros(len(blob) + (4 << 20), np.uint8)
    n = lib.ref_frame_container(ctypes.c_void_p(blob.ctypes.data), ctypes.c_void_p(lens.ctypes.data), len(fields), ctypes.c_void_p(er.ctypes.data), len(er),
                                len(again), ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(out)))
    assert n > 0
    return out[:n].copy(), streams


@needs_ref
def test_keyframe_container_against_the_reference_headers():
    """four keyframes inserted, one erased, one more inserted into the freed slot: every slot's Frame stream is found (the 196 never-used slots hold
    default-constructed Frames, which must parse too) and the live slots carry, byte for byte, the streams that went in"""
    from ucoslam_b200 import mappoint_container_walk
    fields = [dict(make_fields(20 + i, n_kp=60 + 10 * i, n_markers=i % 2), idx=i) for i in range(4)]
    again = [dict(make_fields(30, n_kp=45, n_markers=0), idx=9)]
    ref, streams = _ref_frame_container(fields, erase=[1], again=again)
    c, off, valid, used = mappoint_container_walk(ref, frames=True)
    assert used == len(ref) and (c.n_slots, c.n_used, c.n_valid, c.n_free) == (200, 4, 4, 0) and list(valid[:5]) == [1, 1, 1, 1, 0]
    want = [streams[0], streams[4], streams[2], streams[3]]
    for i in range(4):
        v, n = frame_stream_parse(ref[int(off[i]):])
        assert n == len(want[i]) and np.array_equal(ref[int(off[i]):int(off[i]) + n], want[i]) and v.idx == (9 if i == 1 else i)
    v, n = frame_stream_parse(ref[int(off[150]):])             # a default-constructed Frame
    assert v.n_und_kpts == 0 and v.n_ids == 0 and v.desc.rows == 0
    # the writer rebuilds the section from the slots' byte ranges (a changed keyframe would go in as frame_stream_write output)
    lib = ucoslam_b200.load()
    ends = list(off[1:] - 1) + [used - 12]                       # a slot ends where the next flag byte starts; the last one before curBuffer / curElm / chunk
    lens = np.array([int(e) - int(o) for o, e in zip(off, ends)], np.uint64)
    ptrs = (ctypes.c_void_p * c.n_slots)(*[ref.ctypes.data + int(o) for o in off])
    out, n = np.zeros(len(ref), np.uint8), ctypes.c_size_t()
    assert lib.uco_b200_frame_container_write(ctypes.addressof(c), ctypes.addressof(ptrs), lens.ctypes.data, valid.ctypes.data, out.ctypes.data, len(out),
                                              ctypes.addressof(n)) == 0 and n.value == len(ref) and np.array_equal(out, ref)
    with pytest.raises(ucoslam_b200.UcoError):
        mappoint_container_walk(ref, frames=False)              # a keyframe section is not a map-point section
    with pytest.raises(ucoslam_b200.UcoError):
        mappoint_container_walk(np.ascontiguousarray(ref[:-7]), frames=True)


# ---- the other sections of Map::toStream and the whole stream --------------------------------------------------------------------------------
def _ref_bytes(fn, h):
    out = np.zeros(8 << 20, np.uint8)
    fn.restype = ctypes.c_long
    n = fn(ctypes.c_void_p(h), ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(out)))
    assert n > 0
    return out[:n].copy()


def _ref_sections(tmp_path):
    """keyframe-database and covisibility-graph sections written by the reference's own keyframedatabase.cpp / covisgraph.cpp (compiled unchanged,
    libref_kfdb.so) on a small synthetic vocabulary, the marker section by its Marker::toStream statements under its own toStream__kv_complex"""
    import oracle_py
    from ucoslam_b200 import workload
    voc = workload.synth_vocabulary_full(seed=4, k=6, depth=3)
    vp = os.path.join(str(tmp_path), "voc.fbow")
    with open(vp, "wb") as f:
        f.write(voc)
    db = oracle_py.RefKeyFrameDataBase(vp)
    rng = np.random.default_rng(3)
    n_words = 0
    for idx in (0, 1, 2, 5, 9):
        ids, w = db.add(idx, rng.integers(0, 256, (300, 32), dtype=np.uint8))
        n_words += len(ids)
    db.delete(2)
    for a, b, w in ((0, 1, 31.0), (1, 5, 12.0), (5, 9, 40.0), (0, 9, 7.0)):
        db.covis_edge(a, b, w)
    kf = _ref_bytes(db.lib.ref_kfdb_to_stream, db.h)
    cv = _ref_bytes(db.lib.ref_covis_to_stream, db.h)
    lib = ctypes.CDLL(REF)
    lib.ref_marker_map_to_stream.restype = ctypes.c_long
    ids = np.array([17, 3, 250], np.uint32); size = np.array([0.2, 0.15, 0.3], np.float32)
    pose = rng.normal(0, 1, (3, 16)).astype(np.float32); nfr = np.array([2, 0, 3], np.int32); fr = np.array([0, 5, 1, 5, 9], np.uint32)
    out = np.zeros(4096, np.uint8)
    n = lib.ref_marker_map_to_stream(3, ctypes.c_void_p(ids.ctypes.data), ctypes.c_void_p(size.ctypes.data), ctypes.c_void_p(pose.ctypes.data),
                                     ctypes.c_void_p(nfr.ctypes.data), ctypes.c_void_p(fr.ctypes.data), ctypes.c_void_p(out.ctypes.data), ctypes.c_long(len(out)))
    assert n > 0
    return dict(kfdb=kf, covis=cv, markers=out[:n].copy(), voc=np.frombuffer(voc, np.uint8), marker_in=(ids, size, pose, nfr, fr))


@needs_ref
def test_map_sections_against_the_reference(tmp_path):
    from ucoslam_b200 import KfdbStream, CovisStream, MarkerStream, map_stream_walk
    lib = ucoslam_b200.load()
    S = _ref_sections(tmp_path)
    used = ctypes.c_size_t()
    # keyframe database: type 1, the vocabulary stream is the vocabulary file, frame ids of the frames still in the database
    k = KfdbStream()
    assert lib.uco_b200_kfdb_stream_walk(S["kfdb"].ctypes.data, len(S["kfdb"]), ctypes.addressof(k), ctypes.addressof(used)) == 0 and used.value == len(S["kfdb"])
    assert k.type == 1 and np.array_equal(S["kfdb"][k.voc_off:k.voc_off + k.voc_len], S["voc"])
    assert list(np.ctypeslib.as_array(ctypes.cast(k.frames, ctypes.POINTER(ctypes.c_uint32)), (k.n_frames,))) == [0, 1, 5, 9]
    assert k.n_words > 0 and k.n_word_frames >= k.n_words
    # covisibility graph
    c = CovisStream()
    assert lib.uco_b200_covis_stream_walk(S["covis"].ctypes.data, len(S["covis"]), ctypes.addressof(c), ctypes.addressof(used)) == 0 and used.value == len(S["covis"])
    assert list(np.ctypeslib.as_array(ctypes.cast(c.nodes, ctypes.POINTER(ctypes.c_uint32)), (c.n_nodes,))) == [0, 1, 5, 9]
    assert c.n_adj == 4 and c.n_neighbours == 8 and c.n_weights == 4
    w = np.frombuffer(ctypes.string_at(c.weights, 12 * c.n_weights), np.dtype([("key", "<u8"), ("w", "<f4")]))
    assert sorted(w["w"].tolist()) == [7.0, 12.0, 31.0, 40.0]
    # markers: std::map order (ascending key), every field
    ids, size, pose, nfr, fr = S["marker_in"]
    mk, n = (MarkerStream * 3)(), ctypes.c_uint32()
    assert lib.uco_b200_marker_map_walk(S["markers"].ctypes.data, len(S["markers"]), 3, ctypes.addressof(mk), ctypes.addressof(n), ctypes.addressof(used)) == 0
    assert n.value == 3 and used.value == len(S["markers"]) and [m.key for m in mk] == [3, 17, 250] and [m.id for m in mk] == [3, 17, 250]
    order = [1, 0, 2]
    starts = np.concatenate([[0], np.cumsum(nfr)])
    for m, i in zip(mk, order):
        assert m.size == size[i] and np.array_equal(np.array(list(m.pose_g2m), np.float32), pose[i]) and m.n_frames == nfr[i]
        assert sorted(fr[starts[i]:starts[i + 1]].tolist()) == list(np.ctypeslib.as_array(ctypes.cast(m.frames, ctypes.POINTER(ctypes.c_uint32)), (m.n_frames,))) if m.n_frames else True
        assert ctypes.string_at(m.dict, m.dict_len) == b"ARUCO_MIP_36h12"
    # the whole stream in the order of Map::toStream (map.cpp:316-325), as a file (Map::saveToFile's magic in front)
    pts = _ref_container([dict(_mp_fields(70 + i, n_frames=2), id=i) for i in range(5)], erase=[3])
    frs, _ = _ref_frame_container([dict(make_fields(40 + i, n_kp=50, n_markers=1), idx=i) for i in range(3)])
    blob = np.concatenate([np.frombuffer(np.uint64(225237123).tobytes(), np.uint8), S["kfdb"], pts, S["markers"], frs, S["covis"]])
    o = map_stream_walk(blob, has_file_magic=True)
    assert (o.kfdb_off, o.kfdb_len) == (8, len(S["kfdb"])) and (o.points_off, o.points_len) == (8 + len(S["kfdb"]), len(pts))
    assert (o.markers_len, o.frames_len, o.covis_len, o.total_len) == (len(S["markers"]), len(frs), len(S["covis"]), len(blob))
    assert (o.points.n_used, o.points.n_valid, o.frames.n_valid, o.n_markers, o.kfdb.n_frames, o.covis.n_weights) == (5, 4, 3, 3, 4, 4)
    o2 = map_stream_walk(np.ascontiguousarray(blob[8:]))
    assert o2.total_len == len(blob) - 8 and o2.frames_off == o.frames_off - 8
    for bad in (blob[:-3], blob[:len(blob) // 2], np.concatenate([blob[:8], blob[12:]])):
        with pytest.raises(ucoslam_b200.UcoError):
            map_stream_walk(np.ascontiguousarray(bad), has_file_magic=True)


def test_map_file_golden():
    """tests/golden/map_file.bin: a complete small map file whose five sections were written by the reference's own code (make_frame_golden.py)"""
    from ucoslam_b200 import map_stream_walk, mappoint_container_walk, mappoints_from_container
    blob = np.fromfile(os.path.join(GOLD, "map_file.bin"), np.uint8)
    o = map_stream_walk(blob, has_file_magic=True)
    assert o.total_len == len(blob) and o.kfdb.type == 1 and o.kfdb.n_frames == 4 and o.n_markers == 3
    assert (o.points.n_slots, o.points.n_used, o.points.n_valid, o.points.n_free) == (200, 5, 4, 1) and (o.frames.n_used, o.frames.n_valid) == (3, 3)
    assert (o.covis.n_nodes, o.covis.n_weights) == (4, 4)
    mp = mappoints_from_container(blob[o.points_off:o.points_off + o.points_len])
    assert list(mp["ids"]) == [0, 1, 2, 4]
    c, off, valid, used = mappoint_container_walk(blob[o.frames_off:o.frames_off + o.frames_len], frames=True)
    v, n = frame_stream_parse(blob[o.frames_off + int(off[1]):])
    assert v.idx == 1 and v.n_und_kpts == 50
    # the vocabulary section loads as a vocabulary (host-side header check: signature, block geometry)
    voc = blob[o.kfdb_off + o.kfdb.voc_off:o.kfdb_off + o.kfdb.voc_off + o.kfdb.voc_len]
    assert int(np.frombuffer(voc[:8].tobytes(), np.uint64)[0]) == 55824124


@pytest.mark.gpu
def test_map_file_becomes_device_resident_state():
    """the reading side of SURVEY 8(f)4 end to end on the GPU: from the bytes of a map file (tests/golden/map_file.bin, sections written by the
    reference's own code) the vocabulary is loaded into the device BoW transform, every keyframe becomes a device mirror and the map points become the
    rows the projection matcher takes - no Frame / MapPoint / Vocabulary object is built on the way"""
    from ucoslam_b200 import map_stream_walk, mappoint_container_walk, mappoints_from_container, workload
    ctx = ucoslam_b200.Context(0)
    blob = np.fromfile(os.path.join(GOLD, "map_file.bin"), np.uint8)
    o = map_stream_walk(blob, has_file_magic=True)
    # vocabulary: the same words / weights as the vocabulary the database was created with
    vb = blob[o.kfdb_off + o.kfdb.voc_off:o.kfdb_off + o.kfdb.voc_off + o.kfdb.voc_len]
    v1, v2 = ctx.bow_load(vb), ctx.bow_load(np.frombuffer(workload.synth_vocabulary_full(seed=4, k=6, depth=3), np.uint8))
    desc = np.random.default_rng(2).integers(0, 256, (500, 32), dtype=np.uint8)
    a, b = ctx.bow_transform(v1, desc, 2), ctx.bow_transform(v2, desc, 2)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(np.unique(a[0])) > 20
    # keyframes: every valid slot -> device mirror -> the per-keypoint arrays come back as the fields that were written
    sec = blob[o.frames_off:o.frames_off + o.frames_len]
    c, off, valid, _ = mappoint_container_walk(sec, frames=True)
    handles = []
    for i in np.nonzero(valid)[0]:
        view, _ = frame_stream_parse(sec[int(off[i]):])
        h, d = ctx.frame_upload(view)
        fd = make_fields(40 + int(i), n_kp=50, n_markers=1)
        kps, dsc, ids, flags, depth = ctx.frame_download(h, d.n_kp)
        assert d.idx == i and d.n_kp == 50 and kps.tobytes() == fd["kp"].tobytes() and np.array_equal(dsc, fd["desc"]) and np.array_equal(ids, fd["ids"])
        handles.append(h)
    assert len(handles) == 3
    # map points: the valid slots as matcher rows
    mp = mappoints_from_container(blob[o.points_off:o.points_off + o.points_len])
    assert list(mp["ids"]) == [0, 1, 2, 4] and mp["pos"].shape == (4, 3)
    for h in handles:
        ctx.frame_free(h)
    ctx.bow_free(v1); ctx.bow_free(v2)
    ctx.close()


@needs_ref
def test_map_section_writers_reproduce_the_reference_bytes(tmp_path):
    """marker map, covisibility graph and keyframe database: unpack the reference-written section, write it again from the unpacked arrays -> the
    same bytes (so a map file can be written here section by section)"""
    from ucoslam_b200 import KfdbStream, CovisStream, MarkerStream
    lib = ucoslam_b200.load()
    S = _ref_sections(tmp_path)
    n, used = ctypes.c_size_t(), ctypes.c_size_t()
    P = lambda a: a.ctypes.data
    # markers
    mk, cnt = (MarkerStream * 3)(), ctypes.c_uint32()
    assert lib.uco_b200_marker_map_walk(P(S["markers"]), len(S["markers"]), 3, ctypes.addressof(mk), ctypes.addressof(cnt), None) == 0
    out = np.zeros(len(S["markers"]), np.uint8)
    assert lib.uco_b200_marker_map_write(ctypes.addressof(mk), 3, P(out), len(out), ctypes.addressof(n)) == 0 and np.array_equal(out, S["markers"])
    mk[1].key = 1                                               # not ascending any more: std::map could not have produced it
    assert lib.uco_b200_marker_map_write(ctypes.addressof(mk), 3, P(out), len(out), ctypes.addressof(n)) != 0
    # covisibility graph
    c = CovisStream()
    assert lib.uco_b200_covis_stream_walk(P(S["covis"]), len(S["covis"]), ctypes.addressof(c), None) == 0
    an, ap, ai = np.zeros(c.n_adj, np.uint32), np.zeros(c.n_adj + 1, np.uint32), np.zeros(c.n_neighbours, np.uint32)
    wk, ww = np.zeros(c.n_weights, np.uint64), np.zeros(c.n_weights, np.float32)
    assert lib.uco_b200_covis_stream_unpack(P(S["covis"]), len(S["covis"]), ctypes.addressof(c), P(an), P(ap), P(ai), P(wk), P(ww)) == 0
    assert list(an) == [0, 1, 5, 9] and list(ai[ap[0]:ap[1]]) == [1, 9] and sorted(ww.tolist()) == [7.0, 12.0, 31.0, 40.0]
    nodes = np.ctypeslib.as_array(ctypes.cast(c.nodes, ctypes.POINTER(ctypes.c_uint32)), (c.n_nodes,)).copy()
    out = np.zeros(len(S["covis"]), np.uint8)
    assert lib.uco_b200_covis_stream_write(c.n_nodes, P(nodes), c.n_adj, P(an), P(ap), P(ai), c.n_weights, P(wk), P(ww), P(out), len(out), ctypes.addressof(n)) == 0
    assert n.value == len(out) and np.array_equal(out, S["covis"])
    # keyframe database
    k = KfdbStream()
    assert lib.uco_b200_kfdb_stream_walk(P(S["kfdb"]), len(S["kfdb"]), ctypes.addressof(k), None) == 0
    wd, wp, wf = np.zeros(k.n_words, np.uint32), np.zeros(k.n_words + 1, np.uint32), np.zeros(k.n_word_frames, np.uint32)
    assert lib.uco_b200_kfdb_stream_unpack(P(S["kfdb"]), len(S["kfdb"]), ctypes.addressof(k), P(wd), P(wp), P(wf)) == 0
    assert (np.diff(wd.astype(np.int64)) > 0).all() and set(wf.tolist()) <= {0, 1, 5, 9}
    frames = np.ctypeslib.as_array(ctypes.cast(k.frames, ctypes.POINTER(ctypes.c_uint32)), (k.n_frames,)).copy()
    voc = np.ascontiguousarray(S["kfdb"][k.voc_off:k.voc_off + k.voc_len])
    out = np.zeros(len(S["kfdb"]), np.uint8)
    assert lib.uco_b200_kfdb_stream_write(1, P(voc), len(voc), k.n_words, P(wd), P(wp), P(wf), k.n_frames, P(frames), P(out), len(out), ctypes.addressof(n)) == 0
    assert n.value == len(out) and np.array_equal(out, S["kfdb"])


def test_map_file_walkers_survive_damage():
    """truncations and byte flips of the golden map file: every walker answers OK or an error, never reads outside the buffer it was given
    (the buffer is copied to its exact size, so an overrun shows under the allocator / sanitizer rather than by luck)"""
    from ucoslam_b200 import MapSections, MapPointContainer
    lib = ucoslam_b200.load()
    blob = np.fromfile(os.path.join(GOLD, "map_file.bin"), np.uint8)
    rng = np.random.default_rng(12)
    o = MapSections()
    n_ok = 0
    for t in range(400):
        b = blob.copy()
        if t % 2:
            b = np.ascontiguousarray(b[:int(rng.integers(0, len(b)))])
        else:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        if len(b) == 0:
            continue
        rc = lib.uco_b200_map_stream_walk(b.ctypes.data, len(b), 1, ctypes.addressof(o))
        assert rc in (0, -1, -2, -3, -4, -5, -6)
        if rc == 0:
            n_ok += 1
            assert o.total_len <= len(b)
            n = ctypes.c_uint32()
            lib.uco_b200_mappoints_from_container(b.ctypes.data + o.points_off, o.points_len, 0, None, None, None, None, None, None, None, ctypes.addressof(n), None)
    assert n_ok > 20        # flips inside payload bytes (descriptors, poses, weights) leave a well-formed file
