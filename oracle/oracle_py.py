"""TEST INFRASTRUCTURE ONLY — Python loaders for the CPU oracles (oracle/_build/liboracle.so, oracle/_ref/*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes, os, subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_sz = ctypes.c_size_t


def build_oracle():
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def build_ref():
    """Compile the reference's own sources (only where /root/reference exists)."""
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_oracle = None


def load_oracle():
    global _oracle
    if _oracle is None:
        p = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(p):
            build_oracle()
        _oracle = ctypes.CDLL(p)
    return _oracle


def load_ref(name):
    p = os.path.join(HERE, "_ref", name)
    return ctypes.CDLL(p) if os.path.exists(p) else None


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def hamming_knn(q, t, k, order=0):
    """oracle/knn_oracle.c on (nq,32)/(nt,32) uint8 arrays -> (idx, dist) int32 (nq,k)."""
    lib = load_oracle()
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), k), np.int32)
    dist = np.empty((len(q), k), np.int32)
    lib.oracle_hamming_knn(_p(q), len(q), _sz(32), _p(t), len(t), _sz(32), k, order, _p(idx), _p(dist))
    return idx, dist


def ref_xflann_knn(q, t, k, kind=0, max_checks=-1, sorted_=0):
    """The reference's xflann (oracle/_ref/libref_xflann.so). kind 0 = linear (exact), 1 = HKMeans(32,0)."""
    lib = load_ref("libref_xflann.so")
    if lib is None:
        return None
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), k), np.int32)
    dist = np.empty((len(q), k), np.int32)
    rc = lib.ref_xflann_knn(_p(q), len(q), _p(t), len(t), k, kind, max_checks, sorted_, _p(idx), _p(dist))
    if rc != 0:
        raise RuntimeError("ref_xflann_knn failed rc=%d" % rc)
    return idx, dist


def synth_descriptors(seed, nt, nq, max_flips=40):
    """uniform 256-bit train rows; queries = train rows with 0..max_flips random bit flips."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    if nt == 0:
        return t, rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    src = rng.integers(0, nt, nq)
    q = t[src].copy()
    nflip = rng.integers(0, max_flips + 1, nq)
    for i in range(nq):
        bits = rng.integers(0, 256, nflip[i])
        np.bitwise_xor.at(q[i], bits >> 3, (1 << (bits & 7)).astype(np.uint8))
    return t, q


# ---- bag of words (fbow) -------------------------------------------------------------------------------------------
REF_VOC_PATH = os.path.join(HERE, "_ref", "orb.fbow")   # the reference's shipped vocabulary (copied by `make ref`)


def ref_voc_bytes():
    if not os.path.exists(REF_VOC_PATH):
        return None
    return np.fromfile(REF_VOC_PATH, dtype=np.uint8)


def synth_vocabulary(seed, k=10, depth=4, desc_size=32, leaf_prob=0.08, partial_prob=0.1, weight_scale=1.0):
    """A seeded vocabulary file image in fbow's stream format (fbow.cpp:160-190, fbow.h:125-194): a k-ary tree of
    `depth` levels of internal blocks, with some early leaves and some blocks holding fewer than k nodes."""
    rng = np.random.default_rng(seed)
    feature_off, desc_wp = 8, 32
    child_off = feature_off + k * desc_wp
    block_size = child_off + 8 * k
    block_size = (block_size + 7) // 8 * 8
    blocks = []   # list of dict(n, parent, feats, infos)
    words = [0]

    def make_block(level, parent):
        b = len(blocks)
        n = k if rng.random() > partial_prob else int(rng.integers(1, k + 1))
        blk = dict(n=n, parent=parent, feats=rng.integers(0, 256, (k, desc_size), dtype=np.uint8), ids=[0] * k, w=[0.0] * k,
                   leaf=0)
        blocks.append(blk)
        for c in range(n):
            if level == depth - 1 or rng.random() < leaf_prob:
                blk["ids"][c] = 0x80000000 | words[0]
                blk["w"][c] = float(np.float32(np.float32(rng.random() * 3 + 0.01) * np.float32(weight_scale)))
                words[0] += 1
            else:
                blk["ids"][c] = make_block(level + 1, b)
        return b

    make_block(0, 0)
    nb = len(blocks)
    data = np.zeros(nb * block_size, np.uint8)
    for i, blk in enumerate(blocks):
        o = i * block_size
        data[o:o + 2] = np.frombuffer(np.uint16(blk["n"]).tobytes(), np.uint8)
        data[o + 4:o + 8] = np.frombuffer(np.uint32(blk["parent"]).tobytes(), np.uint8)
        data[o + feature_off:o + feature_off + k * desc_wp] = blk["feats"].reshape(-1)
        info = np.zeros(k, dtype=[("id", "<u4"), ("w", "<f4")])
        info["id"] = blk["ids"]
        info["w"] = blk["w"]
        data[o + child_off:o + child_off + 8 * k] = np.frombuffer(info.tobytes(), np.uint8)
    import struct
    hdr = bytearray(128)
    struct.pack_into("<Q", hdr, 0, 55824124)
    hdr[8:11] = b"orb"
    struct.pack_into("<II", hdr, 8 + 52, 8, nb)
    struct.pack_into("<5Q", hdr, 8 + 64, desc_wp, block_size, feature_off, child_off, len(data))
    struct.pack_into("<iiI", hdr, 8 + 104, 0, desc_size, k)
    return np.concatenate([np.frombuffer(bytes(hdr), np.uint8), data])


def bow_transform(voc_bytes, desc, level):
    """oracle/bow_oracle.c: per-descriptor (word, weight, node)."""
    lib = load_oracle()
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    voc_bytes = np.ascontiguousarray(voc_bytes, np.uint8)
    n = len(desc)
    word = np.empty(n, np.uint32)
    weight = np.empty(n, np.float32)
    node = np.empty(n, np.uint32)
    rc = lib.oracle_bow_transform(_p(voc_bytes), _sz(len(voc_bytes)), _p(desc), n, _sz(32), int(level), _p(word),
                                  _p(weight), _p(node))
    if rc != 0:
        raise RuntimeError("oracle_bow_transform rc=%d" % rc)
    return word, weight, node


def fold_bow(word, weight, node):
    """What the host adapter does: fBow = map word -> sum of weights (f32, descriptor order); fBow2 = map node -> indices."""
    bow, bow2 = {}, {}
    for i in range(len(word)):
        if word[i] != 0xFFFFFFFF:
            bow[int(word[i])] = np.float32(bow.get(int(word[i]), np.float32(0)) + np.float32(weight[i]))
        if node[i] != 0xFFFFFFFF:
            bow2.setdefault(int(node[i]), []).append(i)
    ids = np.array(sorted(bow), np.uint32)
    w = np.array([bow[int(i)] for i in ids], np.float32)
    n2 = np.array([k for k in sorted(bow2) for _ in bow2[k]], np.uint32)
    f2 = np.array([x for k in sorted(bow2) for x in bow2[k]], np.uint32)
    return ids, w, n2, f2


class RefVocabulary:
    """The reference's fbow::Vocabulary (oracle/_ref/libref_fbow.so)."""

    def __init__(self, voc_bytes):
        self.lib = load_ref("libref_fbow.so")
        if self.lib is None:
            raise RuntimeError("oracle/_ref/libref_fbow.so not built")
        self.lib.ref_fbow_load_bytes.restype = ctypes.c_void_p
        self.lib.ref_fbow_score.restype = ctypes.c_double
        voc_bytes = np.ascontiguousarray(voc_bytes, np.uint8)
        self.h = self.lib.ref_fbow_load_bytes(_p(voc_bytes), _sz(len(voc_bytes)))
        if not self.h:
            raise RuntimeError("fbow rejected the vocabulary")

    def transform(self, desc, level):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        ids = np.empty(n, np.uint32); w = np.empty(n, np.float32); n2 = np.empty(n, np.uint32); f2 = np.empty(n, np.uint32)
        nb, nb2 = ctypes.c_int(), ctypes.c_int()
        rc = self.lib.ref_fbow_transform(ctypes.c_void_p(self.h), _p(desc), n, int(level), _p(ids), _p(w), ctypes.byref(nb),
                                         _p(n2), _p(f2), ctypes.byref(nb2))
        if rc != 0:
            raise RuntimeError("fbow transform threw")
        return ids[:nb.value], w[:nb.value], n2[:nb2.value], f2[:nb2.value]

    def close(self):
        if self.h:
            self.lib.ref_fbow_free(ctypes.c_void_p(self.h))
            self.h = None


# ---- bundle adjustment -----------------------------------------------------------------------------------------------
from ucoslam_b200.synth import synth_ba_problem  # seeded workload generator shared with bench.py (not part of the oracle)


BA_INPUT_KEYS = ("poses44", "fixed", "points3", "obs_pose", "obs_point", "obs_uv", "obs_ur", "obs_stereo", "obs_inv_sigma2",
                 "fx", "fy", "cx", "cy", "bf")


def ba_problem_from_golden(g, name):
    pb = {k: g["%s_in_%s" % (name, k)] for k in BA_INPUT_KEYS}
    for k in ("fx", "fy", "cx", "cy", "bf"):
        pb[k] = float(pb[k])
    return pb


def ba_optimize(pb, n_iters):
    """oracle/ba_oracle.c (plain-C restatement) on a problem dict from synth_ba_problem."""
    return _ba_call(load_oracle().oracle_ba_optimize, pb, n_iters)


def ref_ba_optimize(pb, n_iters):
    """The reference's g2o bundle adjustment (oracle/_ref/libref_g2o.so) on a problem dict from synth_ba_problem."""
    lib = load_ref("libref_g2o.so")
    if lib is None:
        return None
    if len(pb.get("plane_other", ())):   # InPlaneMarkers: MarkerEdgeX restated on the reference's g2o (ref_g2o_wrap.cpp PlanarMarkerEdge)
        return _ba_call_markers(lib.ref_ba_optimize_planar, pb, n_iters)
    if pb.get("pose_cam") is not None:   # one camera per keyframe: the per-edge ImageParams of globaloptimizer_g2o.cpp:233-236, :262-266, :335-338
        return _ba_call_markers(lib.ref_ba_optimize_cams, pb, n_iters)
    if len(pb.get("marker_size", ())):
        return _ba_call_markers(lib.ref_ba_optimize_markers, pb, n_iters)
    return _ba_call(lib.ref_ba_optimize, pb, n_iters)


def _ba_call_markers(fn, pb, n_iters):
    """ref_ba_optimize_markers: the reference's g2o + its own MarkerEdge class on a problem with ArUco markers"""
    P, N, M = len(pb["fixed"]), len(pb["points3"]), len(pb["obs_pose"])
    if not len(pb.get("marker_size", ())):
        pb = dict(pb, marker_pose44=np.zeros((0, 16), np.float32), marker_size=np.zeros(0, np.float32), mobs_marker=np.zeros(0, np.int32),
                  mobs_pose=np.zeros(0, np.int32), mobs_corners=np.zeros((0, 8), np.float32), mobs_weight=np.zeros(0, np.float32))
    nm, nmo = len(pb["marker_size"]), len(pb["mobs_marker"])
    out = dict(pose7=np.zeros((P, 7)), pose44=np.zeros((P, 16), np.float32), point3=np.zeros((N, 3)), chi2=np.zeros(M),
               level=np.zeros(M, np.uint8), bad=np.zeros(M, np.uint8), iters=np.zeros(2, np.int32), trace=np.zeros((64, 2)),
               marker_pose7=np.zeros((nm, 7)), marker_pose44=np.zeros((nm, 16), np.float32), mobs_chi2=np.zeros(nmo))
    f = lambda v: ctypes.c_float(v)
    A = lambda k, dt: np.ascontiguousarray(pb[k], dt)
    keep = [A("marker_pose44", np.float32), A("marker_size", np.float32), A("mobs_marker", np.int32), A("mobs_pose", np.int32),
            A("mobs_corners", np.float32), A("mobs_weight", np.float32)]
    fn(P, _p(pb["poses44"]), _p(pb["fixed"]), N, _p(pb["points3"]), M, _p(pb["obs_pose"]), _p(pb["obs_point"]), _p(pb["obs_uv"]),
       _p(pb["obs_ur"]), _p(pb["obs_stereo"]), _p(pb["obs_inv_sigma2"]), f(pb["fx"]), f(pb["fy"]), f(pb["cx"]), f(pb["cy"]), f(pb["bf"]),
       int(n_iters), _p(out["pose7"]), _p(out["pose44"]), _p(out["point3"]), _p(out["chi2"]), _p(out["level"]), _p(out["bad"]),
       _p(out["iters"]), _p(out["trace"]), nm, _p(keep[0]), _p(keep[1]), nmo, _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(keep[5]),
       _p(out["marker_pose7"]), _p(out["marker_pose44"]), _p(out["mobs_chi2"]),
       *([_p(np.ascontiguousarray(pb["pose_cam"], np.float32))] if pb.get("pose_cam") is not None else ([None] if len(pb.get("plane_other", ())) else [])),
       *([len(pb["plane_other"]), int(pb["plane_ref"]),
          _p(np.ascontiguousarray(pb["plane_ref_pose44"], np.float32)) if int(pb["plane_ref"]) < 0 else None,
          _p(np.ascontiguousarray(pb["plane_other"], np.int32)), ctypes.c_double(float(pb["plane_weight"]))] if len(pb.get("plane_other", ())) else []))
    return out


def _ba_call(fn, pb, n_iters):
    P, N, M = len(pb["fixed"]), len(pb["points3"]), len(pb["obs_pose"])
    out = dict(pose7=np.zeros((P, 7)), pose44=np.zeros((P, 16), np.float32), point3=np.zeros((N, 3)), chi2=np.zeros(M),
               level=np.zeros(M, np.uint8), bad=np.zeros(M, np.uint8), iters=np.zeros(2, np.int32), trace=np.zeros((64, 2)))
    c = ctypes
    f = lambda v: c.c_float(v)
    fn(P, _p(pb["poses44"]), _p(pb["fixed"]), N, _p(pb["points3"]), M, _p(pb["obs_pose"]),
       _p(pb["obs_point"]), _p(pb["obs_uv"]), _p(pb["obs_ur"]), _p(pb["obs_stereo"]),
                        _p(pb["obs_inv_sigma2"]), f(pb["fx"]), f(pb["fy"]), f(pb["cx"]), f(pb["cy"]), f(pb["bf"]),
                        int(n_iters), _p(out["pose7"]), _p(out["pose44"]), _p(out["point3"]), _p(out["chi2"]),
                        _p(out["level"]), _p(out["bad"]), _p(out["iters"]), _p(out["trace"]))
    return out


# ---- pose-only optimisation (PnPSolver::solvePnp) ------------------------------------------------------------------------
from ucoslam_b200.synth import synth_pnp_problem

PNP_INPUT_KEYS = ("pose44", "points3", "obs_uv", "obs_ur", "obs_stereo", "obs_inv_sigma2", "stable", "fx", "fy", "cx", "cy", "bf",
                  "marker_pose44", "marker_size", "marker_corners")


def pnp_problem_from_golden(g, name):
    pb = {k: g["%s_in_%s" % (name, k)] for k in PNP_INPUT_KEYS}
    for k in ("fx", "fy", "cx", "cy", "bf"):
        pb[k] = float(pb[k])
    return pb


def pose_only(pb):
    """oracle/ba_oracle.c oracle_pose_only (plain-C restatement of pnpsolver.cpp:116-408)."""
    return _pnp_call(load_oracle().oracle_pose_only, pb)


def ref_pose_only(pb):
    """The reference's own edge classes + g2o (oracle/_ref/libref_g2o.so ref_pose_only)."""
    lib = load_ref("libref_g2o.so")
    if lib is None:
        return None
    return _pnp_call(lib.ref_pose_only, pb)


def _pnp_call(fn, pb):
    n, nm = len(pb["points3"]), len(pb["marker_size"])
    out = dict(pose44=np.zeros(16, np.float32), pose7=np.zeros(7), bad=np.zeros(max(n, 1), np.uint8), iters=np.zeros(4, np.int32))
    f = lambda v: ctypes.c_float(v)
    fn.restype = ctypes.c_int
    out["n_good"] = fn(_p(pb["pose44"]), n, _p(pb["points3"]), _p(pb["obs_uv"]), _p(pb["obs_ur"]), _p(pb["obs_stereo"]),
                       _p(pb["obs_inv_sigma2"]), _p(pb["stable"]), f(pb["fx"]), f(pb["fy"]), f(pb["cx"]), f(pb["cy"]), f(pb["bf"]),
                       nm, _p(pb["marker_pose44"]), _p(pb["marker_size"]), _p(pb["marker_corners"]),
                       _p(out["pose44"]), _p(out["pose7"]), _p(out["bad"]), _p(out["iters"]))
    out["bad"] = out["bad"][:n]
    return out


# ---- frame matcher post-filters (FrameMatcher_Flann::matchEpipolar on exact k-NN) -----------------------------------------
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])


def synth_match_frames(seed, nt=2000, nq=2000, max_flips=40, rot=25.0, w=640, h=480):
    """Two synthetic frames' keypoints + descriptors: query keypoint i re-observes train keypoint perm[i] (descriptor with a few
    bit flips, angle turned by ~rot degrees, octave within +-1, position moved by a small shift); 20 % of the queries are new."""
    rng = np.random.default_rng(seed)
    t_desc = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t_kps = np.zeros(nt, KP_DTYPE)
    t_kps["x"], t_kps["y"] = rng.uniform(20, w - 20, nt), rng.uniform(20, h - 20, nt)
    t_kps["octave"] = rng.integers(0, 8, nt)
    t_kps["angle"] = rng.uniform(0, 360, nt)
    t_kps["size"] = 31 * np.float32(1.2) ** t_kps["octave"]
    t_kps["class_id"] = -1
    src = rng.integers(0, nt, nq)
    q_desc = t_desc[src].copy()
    for i in range(nq):
        nf = int(rng.integers(0, max_flips + 1))
        bits = rng.integers(0, 256, nf)
        for b in bits:
            q_desc[i, b >> 3] ^= np.uint8(1 << (b & 7))
    fresh = rng.random(nq) < 0.2
    q_desc[fresh] = rng.integers(0, 256, (int(fresh.sum()), 32), dtype=np.uint8)
    q_kps = t_kps[src].copy()
    q_kps["x"] += rng.normal(3, 1, nq).astype(np.float32)
    q_kps["y"] += rng.normal(-2, 1, nq).astype(np.float32)
    q_kps["octave"] = np.clip(q_kps["octave"] + rng.integers(-1, 2, nq) * (rng.random(nq) < 0.3), 0, 7)
    ang = t_kps["angle"][src] - rot + rng.normal(0, 6, nq) + (rng.random(nq) < 0.1) * rng.uniform(0, 360, nq)
    q_kps["angle"] = np.mod(ang, 360).astype(np.float32)
    return q_desc, q_kps, t_desc, t_kps


def frame_match(q_desc, q_kps, t_desc, t_kps, min_desc_dist=50.0, ratio=0.8, check_orientation=True, max_octave_diff=1, F12=None,
                scale_factors=None, q_map=None, t_map=None):
    """oracle/match_oracle.c on (n,32) uint8 descriptors and KP_DTYPE keypoints -> MATCH_DTYPE records."""
    lib = load_oracle()
    q_desc = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
    t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
    q_kps, t_kps = np.ascontiguousarray(q_kps, KP_DTYPE), np.ascontiguousarray(t_kps, KP_DTYPE)
    sf = np.asarray(scale_factors if scale_factors is not None else [np.float32(1.2) ** i for i in range(8)], np.float32)
    f12 = None if F12 is None else np.ascontiguousarray(F12, np.float32).reshape(9)
    qm = None if q_map is None else np.ascontiguousarray(q_map, np.int32)
    tm = None if t_map is None else np.ascontiguousarray(t_map, np.int32)
    out = np.zeros(max(len(q_desc), 1), MATCH_DTYPE)
    P = lambda a: None if a is None else _p(a)
    n = lib.oracle_frame_match(_p(q_desc), len(q_desc), _sz(32), _p(q_kps), P(qm), _p(t_desc), len(t_desc), _sz(32), _p(t_kps), P(tm),
                               ctypes.c_float(min_desc_dist), ctypes.c_float(ratio), int(check_orientation), int(max_octave_diff),
                               P(f12), _p(sf), len(sf), _p(out))
    return out[:n].copy()


def ref_frame_match_multi(t_desc, t_kps, q_descs, q_kpss, min_desc_dist=50.0, ratio=0.8, check_orientation=True, max_octave_diff=1,
                          scale_factors=None):
    """FrameMatcher_Flann as the reference runs it (bench.py's CPU arm): setParams builds the reference's own xflann HKMeans(32,0)
    index over the train frame ONCE (framematcher.cpp:200-215), every neighbour searches it with 16 checks (:239), then the
    post-filters (oracle/match_oracle.c on that 10-NN table).  Returns the match lists, or None where oracle/_ref was not built."""
    ref = load_ref("libref_xflann.so")
    if ref is None:
        return None
    lib = load_oracle()
    t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
    t_kps = np.ascontiguousarray(t_kps, KP_DTYPE)
    sf = np.asarray(scale_factors if scale_factors is not None else [np.float32(1.2) ** i for i in range(8)], np.float32)
    ref.ref_xflann_build.restype = ctypes.c_void_p
    h = ref.ref_xflann_build(_p(t_desc), len(t_desc), 1)
    outs = []
    for q_desc, q_kps in zip(q_descs, q_kpss):
        q_desc = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
        q_kps = np.ascontiguousarray(q_kps, KP_DTYPE)
        idx, dist = np.zeros((len(q_desc), 10), np.int32), np.zeros((len(q_desc), 10), np.int32)
        ref.ref_xflann_search(ctypes.c_void_p(h), _p(q_desc), len(q_desc), 10, 16, 0, _p(idx), _p(dist))
        out = np.zeros(max(len(q_desc), 1), MATCH_DTYPE)
        n = lib.oracle_frame_match_knn(_p(idx), _p(dist), len(q_desc), _p(q_kps), None, _p(t_kps), None, ctypes.c_float(min_desc_dist),
                                       ctypes.c_float(ratio), int(check_orientation), int(max_octave_diff), None, _p(sf), len(sf), _p(out))
        outs.append(out[:n].copy())
    ref.ref_xflann_free(ctypes.c_void_p(h))
    return outs


def frame_match_bow(q_desc, q_kps, q_bow, t_desc, t_kps, t_bow, min_desc_dist=50.0, ratio=0.8, check_orientation=True, max_octave_diff=1,
                    F12=None, scale_factors=None, q_usable=None, t_usable=None):
    """oracle/match_oracle.c oracle_frame_match_bow; q_bow / t_bow = (node_id u32 ascending, ptr i32, kp i32)"""
    lib = load_oracle()
    q_desc = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
    t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
    q_kps, t_kps = np.ascontiguousarray(q_kps, KP_DTYPE), np.ascontiguousarray(t_kps, KP_DTYPE)
    sf = np.asarray(scale_factors if scale_factors is not None else [np.float32(1.2) ** i for i in range(8)], np.float32)
    f12 = None if F12 is None else np.ascontiguousarray(F12, np.float32).reshape(9)
    qb = [np.ascontiguousarray(q_bow[0], np.uint32), np.ascontiguousarray(q_bow[1], np.int32), np.ascontiguousarray(q_bow[2], np.int32)]
    tb = [np.ascontiguousarray(t_bow[0], np.uint32), np.ascontiguousarray(t_bow[1], np.int32), np.ascontiguousarray(t_bow[2], np.int32)]
    qu = None if q_usable is None else np.ascontiguousarray(q_usable, np.uint8)
    tu = None if t_usable is None else np.ascontiguousarray(t_usable, np.uint8)
    out = np.zeros(max(len(qb[2]), 1), MATCH_DTYPE)
    P = lambda a: None if a is None else _p(a)
    n = lib.oracle_frame_match_bow(_p(q_desc), _p(q_kps), P(qu), len(qb[0]), _p(qb[0]), _p(qb[1]), _p(qb[2]), _p(t_desc), _p(t_kps), P(tu),
                                   len(tb[0]), _p(tb[0]), _p(tb[1]), _p(tb[2]), ctypes.c_float(min_desc_dist), ctypes.c_float(ratio),
                                   int(check_orientation), int(max_octave_diff), P(f12), _p(sf), len(sf), _p(out))
    return out[:n].copy()


def frame_match_bow_py(q_desc, q_kps, q_bow, t_desc, t_kps, t_bow, min_desc_dist=50.0, ratio=0.8, check_orientation=True, max_octave_diff=1,
                       q_usable=None, t_usable=None):
    """independent pure-Python restatement of FrameMatcher_BoW::matchEpipolar (no epipolar gate) used to cross-check the C oracle"""
    tnode = {int(n): i for i, n in enumerate(t_bow[0])}
    matches = []
    for qi, node in enumerate(q_bow[0]):
        ti = tnode.get(int(node))
        if ti is None:
            continue
        for qidx in q_bow[2][q_bow[1][qi]:q_bow[1][qi + 1]]:
            if q_usable is not None and not q_usable[qidx]:
                continue
            best, best2, bt, o2 = np.float32(min_desc_dist), np.float32(3.4e38), -1, -1
            for tidx in t_bow[2][t_bow[1][ti]:t_bow[1][ti + 1]]:
                if t_usable is not None and not t_usable[tidx]:
                    continue
                if abs(int(t_kps["octave"][tidx]) - int(q_kps["octave"][qidx])) > max_octave_diff:
                    continue
                d = np.float32(int(np.unpackbits(q_desc[qidx] ^ t_desc[tidx]).sum()))
                if d < best:
                    best, bt = d, int(tidx)
                else:
                    best2, o2 = d, int(t_kps["octave"][tidx])
            if bt >= 0 and not (o2 == int(q_kps["octave"][qidx]) and best > np.float32(best2 * np.float32(ratio))):
                matches.append([int(qidx), bt, -1, float(best)])
    # filter_ambiguous_train
    used = {}
    for i, m in enumerate(matches):
        j = used.get(m[1])
        if j is None:
            used[m[1]] = i
        elif matches[j][3] > m[3]:
            matches[j][1] = -1
            used[m[1]] = i
        else:
            m[1] = -1
    matches = [m for m in matches if m[1] != -1]
    if check_orientation:
        bins = []
        for m in matches:
            rot = np.float32(t_kps["angle"][m[1]]) - np.float32(q_kps["angle"][m[0]])
            if rot < 0:
                rot = np.float32(rot + np.float32(360.0))
            b = int(np.floor(np.float32(rot * np.float32(1.0 / 30.0)) + np.float32(0.5)))   # roundf for non-negative values
            bins.append(0 if b == 30 else b)
        cnt = np.bincount(bins, minlength=30) if bins else np.zeros(30, int)
        max1 = max2 = max3 = 0
        i1 = i2 = i3 = -1
        for i in range(30):
            s_ = int(cnt[i])
            if s_ > max1:
                max3, max2, max1, i3, i2, i1 = max2, max1, s_, i2, i1, i
            elif s_ > max2:
                max3, max2, i3, i2 = max2, s_, i2, i
            elif s_ > max3:
                max3, i3 = s_, i
        if max2 < np.float32(0.1) * np.float32(max1):
            i2 = i3 = -1
        elif max3 < np.float32(0.1) * np.float32(max1):
            i3 = -1
        matches = [m for m, b in zip(matches, bins) if b in (i1, i2, i3)]
    out = np.zeros(len(matches), MATCH_DTYPE)
    for i, m in enumerate(matches):
        out[i] = tuple(m)
    return out


# ---- projection matcher (row a12): kd-tree restatement, the reference's own picoflann, Map::matchFrameToMapPoints ------------------
_stl = None


def load_stl():
    global _stl
    if _stl is None:
        p = os.path.join(HERE, "_build", "liboracle_stl.so")
        if not os.path.exists(p):
            build_oracle()
        _stl = ctypes.CDLL(p)
    return _stl


def kdtree_build(xy):
    """restated picoflann build over (n,2) f32 points -> dict(nodes (k,5) i32 [col,left,right,leaf_begin,leaf_count], div (k,2) f32
    [divlow, divhigh], div_val (k,) f64, leaf_idx (n,) i32, bbox (4,) f64)"""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    n = len(xy)
    cap = 2 * n + 4
    ni, nf, nd = np.zeros((cap, 5), np.int32), np.zeros((cap, 2), np.float32), np.zeros(cap, np.float64)
    leaf, bbox = np.zeros(max(n, 1), np.int32), np.zeros(4, np.float64)
    k = load_stl().oracle_kdtree_build(_p(xy), 2, n, _p(ni), _p(nf), _p(nd), _p(leaf), _p(bbox), cap)
    assert k >= 0
    return dict(nodes=ni[:k].copy(), div=nf[:k].copy(), div_val=nd[:k].copy(), leaf_idx=leaf[:n].copy(), bbox=bbox)


def kdtree_radius(xy, queries, radii):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    queries = np.ascontiguousarray(queries, np.float32).reshape(-1, 2)
    radii = np.ascontiguousarray(radii, np.float32)
    nq, cap = len(queries), max(1, len(queries) * len(xy))
    ptr, idx = np.zeros(nq + 1, np.int32), np.zeros(cap, np.int32)
    tot = load_stl().oracle_kdtree_radius(_p(xy), 2, len(xy), _p(queries), _p(radii), nq, _p(ptr), _p(idx), cap)
    assert tot >= 0
    return ptr, idx[:tot].copy()


def ref_picoflann_stream(xy):
    """KdTreeIndex::toStream bytes of the reference's own kd-tree over the points, or None where oracle/_ref was not built"""
    lib = load_ref("libref_picoflann.so")
    if lib is None:
        return None
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    cap = 1024 + 200 * max(1, len(xy))
    buf = np.zeros(cap, np.uint8)
    lib.ref_picoflann_stream.restype = ctypes.c_long
    n = lib.ref_picoflann_stream(_p(xy), len(xy), _p(buf), ctypes.c_long(cap))
    assert n >= 0
    return buf[:n].tobytes()


def ref_picoflann_radius(xy, queries, radii):
    lib = load_ref("libref_picoflann.so")
    if lib is None:
        return None
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    queries = np.ascontiguousarray(queries, np.float32).reshape(-1, 2)
    radii = np.ascontiguousarray(radii, np.float32)
    nq, cap = len(queries), max(1, len(queries) * len(xy))
    ptr, idx = np.zeros(nq + 1, np.int32), np.zeros(cap, np.int32)
    tot = lib.ref_picoflann_radius(_p(xy), len(xy), _p(queries), _p(radii), nq, _p(ptr), _p(idx), cap)
    assert tot >= 0
    return ptr, idx[:tot].copy()


def parse_picoflann_stream(b):
    """the byte format of picoflann.h:603-660 (Index::toStream / Node::toStream) -> the same dict as kdtree_build"""
    import struct
    o = 0
    dims, nvalues = struct.unpack_from("<ii", b, o); o += 8
    (nb,) = struct.unpack_from("<Q", b, o); o += 8
    bbox = np.frombuffer(b, np.float64, 2 * nb, o).copy(); o += 16 * nb
    (k,) = struct.unpack_from("<Q", b, o); o += 8
    nodes, div, div_val, leaf = [], [], [], []
    for _ in range(k):
        (dv,) = struct.unpack_from("<d", b, o); o += 8
        (col,) = struct.unpack_from("<H", b, o); o += 2
        dh, dl = struct.unpack_from("<ff", b, o); o += 8
        l, r = struct.unpack_from("<qq", b, o); o += 16
        (s,) = struct.unpack_from("<Q", b, o); o += 8
        ids = np.frombuffer(b, np.int32, s, o); o += 4 * s
        nodes.append((col, l, r, len(leaf), s)); div.append((dl, dh)); div_val.append(dv); leaf.extend(ids.tolist())
    assert o == len(b)
    return dict(nodes=np.array(nodes, np.int32).reshape(-1, 5), div=np.array(div, np.float32).reshape(-1, 2), div_val=np.array(div_val),
                leaf_idx=np.array(leaf, np.int32), bbox=bbox, n_values=nvalues)


MATCH_DT = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])


def match_projected(sc, min_desc_dist, max_reproj_dist):
    """Map::matchFrameToMapPoints on a scene dict (ucoslam_b200.synth.synth_projection_scene): (matches[MATCH_DT], visible u8)"""
    A = lambda k, dt: np.ascontiguousarray(sc[k], dt)
    ids, pos, nrm = A("mp_id", np.uint32), A("mp_pos", np.float32), A("mp_normal", np.float32)
    dmin, dmax, mdesc = A("mp_min_dist", np.float32), A("mp_max_dist", np.float32), A("mp_desc", np.uint8)
    kxy, koct, kdesc = A("kp_xy", np.float32), A("kp_octave", np.int32), A("kp_desc", np.uint8)
    sf, pose = A("scale_factors", np.float32), A("pose44", np.float32)
    mn, mx = A("min_xy", np.float32), A("max_xy", np.float32)
    m = len(ids)
    out, vis = np.zeros(max(m, 1), MATCH_DT), np.zeros(max(m, 1), np.uint8)
    f = load_stl().oracle_match_projected
    f.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_float] * 4 + \
                 [ctypes.c_void_p] * 3 + [ctypes.c_float] * 2 + [ctypes.c_void_p] * 2
    n = f(m, _p(ids), _p(pos), _p(nrm), _p(dmin), _p(dmax), _p(mdesc), len(kxy), _p(kxy), _p(koct), _p(kdesc), _p(sf), len(sf),
          sc["fx"], sc["fy"], sc["cx"], sc["cy"], _p(mn), _p(mx), _p(pose), min_desc_dist, max_reproj_dist, _p(out), _p(vis))
    return out[:n].copy(), vis[:m].copy()


def track_projected(sc, dist_thr, proj_dist_thr):
    """System::_11946837405316294395 (search by projection from the previous frame, src/utils/system.cpp:5921-6456) on a scene dict
    (ucoslam_b200.synth.synth_track_scene) -> matches[MATCH_DT]"""
    A = lambda k, dt: np.ascontiguousarray(sc[k], dt)
    ids, pos = A("mp_id", np.uint32), A("mp_pos", np.float32)
    kxy, koct, kdesc = A("kp_xy", np.float32), A("kp_octave", np.int32), A("kp_desc", np.uint8)
    poct, pdesc, prow = A("prev_octave", np.int32), A("prev_desc", np.uint8), A("prev_mp_row", np.int32)
    sf, pose = A("scale_factors", np.float32), A("pose44", np.float32)
    mn, mx = A("min_xy", np.float32), A("max_xy", np.float32)
    out = np.zeros(max(len(poct), 1), MATCH_DT)
    f = load_stl().oracle_track_projected
    f.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_float] * 4 + \
                 [ctypes.c_void_p] * 3 + [ctypes.c_float] * 2 + [ctypes.c_void_p]
    n = f(len(poct), _p(poct), _p(pdesc), _p(prow), _p(ids), _p(pos), len(kxy), _p(kxy), _p(koct), _p(kdesc), _p(sf), len(sf),
          sc["fx"], sc["fy"], sc["cx"], sc["cy"], _p(mn), _p(mx), _p(pose), dist_thr, proj_dist_thr, _p(out))
    return out[:n].copy()


def ref_match_projected(sc, min_desc_dist, max_reproj_dist):
    """The REFERENCE's own Map::matchFrameToMapPoints (oracle/_ref/libref_project.so: src/map.cpp:651-770 and the helpers it calls,
    compiled from the reference's statements against container stand-ins); map point ids are the row numbers.  None where
    oracle/_ref was not built."""
    lib = load_ref("libref_project.so")
    if lib is None:
        return None
    A = lambda k, dt: np.ascontiguousarray(sc[k], dt)
    pos, nrm = A("mp_pos", np.float32), A("mp_normal", np.float32)
    dmin, dmax, mdesc = A("mp_min_dist", np.float32), A("mp_max_dist", np.float32), A("mp_desc", np.uint8)
    kxy, koct, kdesc = A("kp_xy", np.float32), A("kp_octave", np.int32), A("kp_desc", np.uint8)
    sf, pose = A("scale_factors", np.float32), A("pose44", np.float32)
    mn, mx = A("min_xy", np.float32), A("max_xy", np.float32)
    m = len(pos)
    out, vis = np.zeros(max(m, 1), MATCH_DT), np.zeros(max(m, 1), np.uint8)
    f = lib.ref_match_projected
    f.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_float] * 4 + \
                 [ctypes.c_void_p] * 3 + [ctypes.c_float] * 2 + [ctypes.c_void_p] * 2
    n = f(m, _p(pos), _p(nrm), _p(dmin), _p(dmax), _p(mdesc), len(kxy), _p(kxy), _p(koct), _p(kdesc), _p(sf), len(sf),
          sc["fx"], sc["fy"], sc["cx"], sc["cy"], _p(mn), _p(mx), _p(pose), min_desc_dist, max_reproj_dist, _p(out), _p(vis))
    assert n >= 0
    return out[:n].copy(), vis[:m].copy()


def ref_track_projected(sc, dist_thr, proj_dist_thr):
    """The REFERENCE's own search by projection from the previous frame (System::_11946837405316294395, src/utils/system.cpp:5921-6456,
    de-obfuscated with the C preprocessor and compiled into oracle/_ref/libref_project.so).  None where oracle/_ref was not built."""
    lib = load_ref("libref_project.so")
    if lib is None:
        return None
    A = lambda k, dt: np.ascontiguousarray(sc[k], dt)
    ids, pos = A("mp_id", np.uint32), A("mp_pos", np.float32)
    kxy, koct, kdesc = A("kp_xy", np.float32), A("kp_octave", np.int32), A("kp_desc", np.uint8)
    poct, pdesc, prow = A("prev_octave", np.int32), A("prev_desc", np.uint8), A("prev_mp_row", np.int32)
    sf, pose = A("scale_factors", np.float32), A("pose44", np.float32)
    mn, mx = A("min_xy", np.float32), A("max_xy", np.float32)
    out = np.zeros(max(len(poct), 1), MATCH_DT)
    f = lib.ref_track_projected
    f.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 4 + \
                 [ctypes.c_int] + [ctypes.c_float] * 4 + [ctypes.c_void_p] * 3 + [ctypes.c_float] * 2 + [ctypes.c_void_p]
    n = f(len(poct), _p(poct), _p(pdesc), _p(prow), len(ids), _p(ids), _p(pos), len(kxy), _p(kxy), _p(koct), _p(kdesc), _p(sf), len(sf),
          sc["fx"], sc["fy"], sc["cx"], sc["cy"], _p(mn), _p(mx), _p(pose), dist_thr, proj_dist_thr, _p(out))
    assert n >= 0
    return out[:n].copy()


def ref_frame_match(q_desc, q_kps, t_desc, t_kps, min_desc_dist=50.0, ratio=0.8, check_orientation=True, max_octave_diff=1, F12=None,
                    scale_factors=None, kind="flann", q_bow=None, t_bow=None, q_ids=None, t_ids=None, q_flags=None, t_flags=None, q_mode=0, t_mode=0):
    """The REFERENCE's own FrameMatcher (oracle/_ref/libref_match.so: src/utils/framematcher.cpp compiled unchanged, with its index
    type swapped to xflann's exact LinearParams so that the post-filters see the product's candidates; kind="bow": FrameMatcher_BoW
    as is).  q_bow / t_bow = (node_id, ptr, kp).  None where oracle/_ref was not built."""
    lib = load_ref("libref_match.so")
    if lib is None:
        return None
    q_desc = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
    t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
    q_kps, t_kps = np.ascontiguousarray(q_kps, KP_DTYPE), np.ascontiguousarray(t_kps, KP_DTYPE)
    sf = np.asarray(scale_factors if scale_factors is not None else [np.float32(1.2) ** i for i in range(8)], np.float32)
    f12 = None if F12 is None else np.ascontiguousarray(F12, np.float32).reshape(9)
    P = lambda a, dt=None: None if a is None else _p(np.ascontiguousarray(a, dt))
    keep = []

    def bow(b):
        if b is None:
            return 0, None, None, None
        arrs = [np.ascontiguousarray(b[0], np.uint32), np.ascontiguousarray(b[1], np.int32), np.ascontiguousarray(b[2], np.int32)]
        keep.extend(arrs)
        return len(arrs[0]), _p(arrs[0]), _p(arrs[1]), _p(arrs[2])
    tb, qb = bow(t_bow), bow(q_bow)
    opt = [None if a is None else np.ascontiguousarray(a, dt) for a, dt in ((t_ids, np.uint32), (t_flags, np.uint8), (q_ids, np.uint32), (q_flags, np.uint8))]
    out = np.zeros(max(len(q_kps), 1), MATCH_DTYPE)
    fn = lib.ref_frame_match
    fn.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int] * 2 + [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 4 + \
                  [ctypes.c_int] * 2 + [ctypes.c_void_p] * 3 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    n = fn(1 if kind == "flann" else 2, len(t_kps), _p(t_kps), _p(t_desc), P(opt[0]), P(opt[1]), t_mode, *tb,
           len(q_kps), _p(q_kps), _p(q_desc), P(opt[2]), P(opt[3]), q_mode, *qb, _p(sf), len(sf), min_desc_dist, ratio, int(check_orientation),
           int(max_octave_diff), P(f12), _p(out), len(out))
    assert n >= 0
    return out[:n].copy()


def filter_ambiguous_query(matches):
    m = np.ascontiguousarray(matches, MATCH_DT).copy()
    n = load_stl().oracle_filter_ambiguous_query(_p(m), len(m))
    return m[:n].copy()


def track_frame(sc, max_desc_dist=50.0, proj_dist_thr=15.0, pnp=None):
    """The tracker's main branch (System::_11166622111371682966, src/utils/system.cpp:6460-6960) on one scene dict, assembled from the
    stage oracles exactly as the reference sequences them: search by projection (maxDescDistance*1.5) -> solvePnp when > 30 matches
    -> on > 30 inliers take the pose, mark the matched points seen, radius 4, else drop the matches, radius projDistThr ->
    matchFrameToMapPoints over the local, unseen points (maxDescDistance*2) -> append -> filter_ambiguous_query -> solvePnp.
    pnp: the pose-only solver to use (default: the plain-C restatement; bench's CPU arm passes the reference's g2o)."""
    pnp = pnp or pose_only
    f32 = np.float32
    sf = np.asarray(sc["scale_factors"], f32)
    kxy, koct = np.asarray(sc["kp_xy"], f32).reshape(-1, 2), np.asarray(sc["kp_octave"], np.int32)
    id2row = {int(v): i for i, v in enumerate(sc["mp_id"])}
    stable = np.asarray(sc.get("mp_stable", np.ones(len(sc["mp_id"]), np.uint8)), np.uint8)
    local = np.asarray(sc.get("mp_local", np.ones(len(sc["mp_id"]), np.uint8)), np.uint8)

    def solve(matches, pose):
        rows = np.array([id2row[int(t)] for t in matches["trainIdx"]], np.int64)
        q = matches["queryIdx"]
        inv = (1.0 / sf[koct[q]].astype(np.float64)).astype(f32)
        pb = dict(pose44=np.asarray(pose, f32).reshape(16).copy(), points3=np.asarray(sc["mp_pos"], f32)[rows], obs_uv=kxy[q],
                  obs_ur=np.zeros(len(q), f32), obs_stereo=np.zeros(len(q), np.uint8), obs_inv_sigma2=inv, stable=stable[rows],
                  fx=sc["fx"], fy=sc["fy"], cx=sc["cx"], cy=sc["cy"], bf=float(sc.get("bf", 0.0)),
                  marker_pose44=np.zeros((0, 16), f32), marker_size=np.zeros(0, f32), marker_corners=np.zeros((0, 8), f32))
        return pnp(pb)

    thr1 = float(f32(float(f32(max_desc_dist)) * 1.5))
    thr2 = float(f32(max_desc_dist) * f32(2))
    pose = np.asarray(sc["pose44"], f32).reshape(16).copy()
    m1 = track_projected(sc, thr1, proj_dist_thr)
    n_tbp, status, n_in1 = len(m1), 0, 0
    if len(m1) > 30:
        r1 = solve(m1, pose)
        n_in1 = int(r1["n_good"])
        if n_in1 > 30:
            pose = np.asarray(r1["pose44"], f32).reshape(16).copy()
            m1["imgIdx"] = np.where(r1["bad"] != 0, -1, 1)
        else:
            status |= 2
    else:
        status |= 1
    reproj = 4.0 if n_in1 > 30 else proj_dist_thr
    keep = local.astype(bool).copy()
    if n_in1 > 30:
        for t in m1["trainIdx"]:
            keep[id2row[int(t)]] = False        # lastFIdxSeen == fseq_idx, map.cpp:659-667
    else:
        m1 = m1[:0]
    rows2 = np.nonzero(keep)[0]
    sc2 = dict(sc, pose44=pose)
    for k in ("mp_id", "mp_pos", "mp_normal", "mp_min_dist", "mp_max_dist", "mp_desc"):
        sc2[k] = np.asarray(sc[k])[rows2]
    m2, vis2 = match_projected(sc2, thr2, reproj)
    visible = np.zeros(len(sc["mp_id"]), np.uint8)
    visible[rows2] = vis2
    allm = filter_ambiguous_query(np.concatenate([m1, m2]))
    out = dict(matches=allm, pose44=pose, n_good=0, status=status, n_tbp=n_tbp, visible=visible)
    if len(allm):
        r2 = solve(allm, pose)
        allm["imgIdx"] = np.where(r2["bad"] != 0, -1, 1)
        out.update(pose44=np.asarray(r2["pose44"], f32).reshape(16), n_good=int(r2["n_good"]), iters=r2["iters"])
    return out


# ---- keyframe database: relocalisation / loop-closure candidates ------------------------------------------------
def synth_places(seed, n_places=12, views_per_place=5, n_desc=400, flip_bits=6, replace_frac=0.25):
    """Seeded descriptor sets of keyframes that revisit a few places: every place has n_desc base descriptors, a view flips up to
    flip_bits bits in each and replaces replace_frac of them by fresh random rows.  Returns (list of (n_desc,32) u8, place of each)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (n_places, n_desc, 32), dtype=np.uint8)
    frames, place = [], []
    for v in range(views_per_place):
        for p in range(n_places):
            d = base[p].copy()
            bits = np.unpackbits(d, axis=1)
            for r in range(n_desc):
                nf = int(rng.integers(0, flip_bits + 1))
                if nf:
                    bits[r, rng.choice(256, nf, replace=False)] ^= 1
            d = np.packbits(bits, axis=1)
            rep = rng.random(n_desc) < replace_frac
            d[rep] = rng.integers(0, 256, (int(rep.sum()), 32), dtype=np.uint8)
            frames.append(d)
            place.append(p)
    return frames, np.array(place)


def synth_covis(seed, frame_ids, place, extra=0.1):
    """Seeded covisibility edges: views of one place are connected (integer-valued float weights with ties), plus a few random
    edges.  Returns (edge_a, edge_b, edge_w)."""
    rng = np.random.default_rng(seed)
    ea, eb, ew = [], [], []
    n = len(frame_ids)
    for i in range(n):
        for j in range(i + 1, n):
            if place[i] == place[j] or rng.random() < extra:
                ea.append(frame_ids[i]); eb.append(frame_ids[j]); ew.append(float(rng.integers(20, 26)))
    return np.array(ea, np.uint32), np.array(eb, np.uint32), np.array(ew, np.float32)


def bow_score(ids1, w1, ids2, w2):
    """oracle/kfdb_oracle.cpp: fBow::score restated."""
    lib = load_stl()
    lib.oracle_bow_score.restype = ctypes.c_double
    ids1 = np.ascontiguousarray(ids1, np.uint32); w1 = np.ascontiguousarray(w1, np.float32)
    ids2 = np.ascontiguousarray(ids2, np.uint32); w2 = np.ascontiguousarray(w2, np.float32)
    return lib.oracle_bow_score(_p(ids1), _p(w1), len(ids1), _p(ids2), _p(w2), len(ids2))


def _csr(bows):
    off = np.zeros(len(bows) + 1, np.int64)
    for i, (ids, w) in enumerate(bows):
        off[i + 1] = off[i] + len(ids)
    words = np.concatenate([np.asarray(b[0], np.uint32) for b in bows]) if bows else np.zeros(0, np.uint32)
    weights = np.concatenate([np.asarray(b[1], np.float32) for b in bows]) if bows else np.zeros(0, np.float32)
    return off, np.ascontiguousarray(words, np.uint32), np.ascontiguousarray(weights, np.float32)


def kfdb_candidates(frame_ids, bows, q_bow, excluded=(), min_score=0.0, sorted_=True, edges=None):
    """oracle/kfdb_oracle.cpp: KPFrameDataBase::relocalizationCandidates restated.  bows: list of (ids, weights) per database
    frame; q_bow: (ids, weights); edges: (edge_a, edge_b, edge_w) of the covisibility graph.
    Returns dict(scored_frame, scored_score, scored_common, max_common, candidates)."""
    lib = load_stl()
    frame_ids = np.ascontiguousarray(frame_ids, np.uint32)
    off, words, weights = _csr(bows)
    qw = np.ascontiguousarray(q_bow[0], np.uint32); qf = np.ascontiguousarray(q_bow[1], np.float32)
    exc = np.ascontiguousarray(list(excluded), np.uint32)
    ea, eb, ew = edges if edges is not None else (np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    n = len(frame_ids)
    sf = np.zeros(max(n, 1), np.uint32); ss = np.zeros(max(n, 1), np.float64); sc = np.zeros(max(n, 1), np.uint32)
    cand = np.zeros(max(n, 1), np.uint32)
    ns, nc, mc = ctypes.c_int(), ctypes.c_int(), ctypes.c_uint32()
    lib.oracle_kfdb_candidates(n, _p(frame_ids), _p(off), _p(words), _p(weights), _p(qw), _p(qf), len(qw), _p(exc), len(exc),
                               ctypes.c_float(min_score), int(bool(sorted_)), len(ea), _p(ea), _p(eb), _p(ew), _p(sf), _p(ss), _p(sc),
                               ctypes.byref(ns), ctypes.byref(mc), _p(cand), ctypes.byref(nc))
    return dict(scored_frame=sf[:ns.value].copy(), scored_score=ss[:ns.value].copy(), scored_common=sc[:ns.value].copy(),
                max_common=int(mc.value), candidates=cand[:nc.value].copy())


def covis_neighbors(edges, idx, cap=4096):
    """CovisGraph::getNeighborsWeights(idx, true) restated: neighbour ids by decreasing weight."""
    lib = load_stl()
    ea, eb, ew = edges
    out = np.zeros(cap, np.uint32)
    n = lib.oracle_covis_neighbors(len(ea), _p(ea), _p(eb), _p(ew), ctypes.c_uint32(int(idx)), _p(out), cap)
    return out[:min(n, cap)].copy()


class RefKeyFrameDataBase:
    """The reference's own KeyFrameDataBase + CovisGraph + fbow (oracle/_ref/libref_kfdb.so) fed with descriptors."""

    def __init__(self, voc_path):
        self.lib = load_ref("libref_kfdb.so")
        if self.lib is None:
            raise RuntimeError("oracle/_ref/libref_kfdb.so not built")
        self.lib.ref_kfdb_create.restype = ctypes.c_void_p
        self.lib.ref_kfdb_score.restype = ctypes.c_float
        self.h = self.lib.ref_kfdb_create(voc_path.encode())
        if not self.h:
            raise RuntimeError("the reference rejected the vocabulary file")

    def add(self, idx, desc):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        ids = np.empty(len(desc), np.uint32); w = np.empty(len(desc), np.float32); nb = ctypes.c_int()
        rc = self.lib.ref_kfdb_add(ctypes.c_void_p(self.h), ctypes.c_uint32(int(idx)), _p(desc), len(desc), _p(ids), _p(w),
                                   ctypes.byref(nb))
        if rc != 0:
            raise RuntimeError("ref_kfdb_add rc=%d" % rc)
        return ids[:nb.value].copy(), w[:nb.value].copy()

    def delete(self, idx):
        rc = self.lib.ref_kfdb_del(ctypes.c_void_p(self.h), ctypes.c_uint32(int(idx)))
        if rc != 0:
            raise RuntimeError("ref_kfdb_del rc=%d" % rc)

    def covis_edge(self, a, b, w):
        self.lib.ref_kfdb_covis_edge(ctypes.c_void_p(self.h), ctypes.c_uint32(int(a)), ctypes.c_uint32(int(b)), ctypes.c_float(w))

    def query(self, desc, sorted_=True, min_score=0.0, excluded=()):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        exc = np.ascontiguousarray(list(excluded), np.uint32)
        out = np.zeros(65536, np.uint32)
        ids = np.empty(len(desc), np.uint32); w = np.empty(len(desc), np.float32); nb = ctypes.c_int()
        n = self.lib.ref_kfdb_query(ctypes.c_void_p(self.h), _p(desc), len(desc), int(bool(sorted_)), ctypes.c_float(min_score),
                                    _p(exc), len(exc), _p(out), len(out), _p(ids), _p(w), ctypes.byref(nb))
        if n < 0:
            raise RuntimeError("ref_kfdb_query rc=%d" % n)
        return out[:n].copy(), (ids[:nb.value].copy(), w[:nb.value].copy())

    def add_bow(self, idx, ids, w):
        ids = np.ascontiguousarray(ids, np.uint32); w = np.ascontiguousarray(w, np.float32)
        rc = self.lib.ref_kfdb_add_bow(ctypes.c_void_p(self.h), ctypes.c_uint32(int(idx)), _p(ids), _p(w), len(ids))
        if rc != 0:
            raise RuntimeError("ref_kfdb_add_bow rc=%d" % rc)

    def query_bow(self, ids, w, sorted_=True, min_score=0.0, excluded=()):
        ids = np.ascontiguousarray(ids, np.uint32); w = np.ascontiguousarray(w, np.float32)
        exc = np.ascontiguousarray(list(excluded), np.uint32)
        out = np.zeros(65536, np.uint32)
        n = self.lib.ref_kfdb_query_bow(ctypes.c_void_p(self.h), _p(ids), _p(w), len(ids), int(bool(sorted_)), ctypes.c_float(min_score),
                                        _p(exc), len(exc), _p(out), len(out))
        if n < 0:
            raise RuntimeError("ref_kfdb_query_bow rc=%d" % n)
        return out[:n].copy()

    def score(self, a, b):
        return float(self.lib.ref_kfdb_score(ctypes.c_void_p(self.h), ctypes.c_uint32(int(a)), ctypes.c_uint32(int(b))))

    def close(self):
        if self.h:
            self.lib.ref_kfdb_free(ctypes.c_void_p(self.h))
            self.h = None


# ---- stereo depth association ----------------------------------------------------------------------------------
def stereo_depth(sc, max_desc_dist=50.0):
    """oracle/stereo_oracle.c on a synth_stereo scene -> (depth f32 (n_l,), match i32 (n_l,), n_with_depth)"""
    lib = load_oracle()
    il, ir = np.ascontiguousarray(sc["img_l"]), np.ascontiguousarray(sc["img_r"])
    h, w = il.shape
    kl, kr = np.ascontiguousarray(sc["kps_l"]), np.ascontiguousarray(sc["kps_r"])
    dl, dr = np.ascontiguousarray(sc["desc_l"]), np.ascontiguousarray(sc["desc_r"])
    depth = np.zeros(len(kl), np.float32); match = np.full(len(kl), -1, np.int32)
    n = lib.oracle_stereo_depth(_p(il), _sz(il.strides[0]), _p(ir), _sz(ir.strides[0]), w, h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr),
                                len(kr), ctypes.c_float(max_desc_dist), ctypes.c_float(sc["bl"]), ctypes.c_float(sc["fx"]), _p(depth),
                                _p(match))
    return depth, match, n


def stereo_depth_py(sc, max_desc_dist=50.0):
    """Independent restatement of the same loop in numpy, with the loop's two OpenCV calls (cv::absdiff, cv::sum) made through cv2."""
    import cv2
    il, ir = sc["img_l"], sc["img_r"]
    rows, cols = il.shape
    kl, kr, dl, dr = sc["kps_l"], sc["kps_r"], sc["desc_l"], sc["desc_r"]
    rnd = lambda v: int(np.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1))      # std::round: half away from zero
    buckets = [[] for _ in range(rows)]
    for j in range(len(kr)):
        y = rnd(np.float64(kr["y"][j]))
        for yy in range(max(0, y), min(rows - 1, y) + 1):
            buckets[yy].append(j)
    depth = np.zeros(len(kl), np.float32); match = np.full(len(kl), -1, np.int32)
    pc = np.array([bin(i).count("1") for i in range(256)], np.int32)
    md = np.float32(max_desc_dist)
    nm = 0
    for i in range(len(kl)):
        y = rnd(kl["y"][i])
        if not (0 <= y < rows):
            continue
        best, bj = None, -1
        for j in buckets[y]:
            if kr["x"][j] > kl["x"][i] or abs(int(kr["octave"][j]) - int(kl["octave"][i])) > 1:
                continue
            d = np.float32(pc[dl[i] ^ dr[j]].sum())
            if d < md and (best is None or d < best):
                best, bj = d, j
        if bj < 0:
            continue
        match[i] = bj
        xl, yl, xr, yr = rnd(kl["x"][i]), rnd(kl["y"][i]), rnd(kr["x"][bj]), rnd(kr["y"][bj])
        if xl < 3 or xl + 3 >= cols or yl < 3 or yl + 3 >= rows or xr < 3 or xr + 3 >= cols or yr < 3 or yr + 3 >= rows:
            continue
        lo, hi = max(-7, -xr), min(7, cols - 1 - xr)
        pl = il[yl - 3:yl + 3, xl - 3:xl + 3]
        sads = {}
        for inc in range(lo, hi + 1):
            xc = xr + inc
            if xc - 3 < 0 or xc + 3 > cols:
                return None
            sads[inc + 7] = float(cv2.sumElems(cv2.absdiff(np.ascontiguousarray(pl), np.ascontiguousarray(ir[yr - 3:yr + 3, xc - 3:xc + 3])))[0])
        b = min(sads, key=lambda k: (sads[k], k))
        if lo + 7 < b < hi + 7:
            d1, d2, d3 = np.float64(sads[b - 1]), np.float64(sads[b]), np.float64(sads[b + 1])
            with np.errstate(all="ignore"):
                off = np.float64(0.5) * (d1 - d3) / (d1 + d3 - 2 * d2) + b - 7
                xs = np.float64(kr["x"][bj]) + off
                depth[i] = np.float32(np.float64(np.float32(sc["bl"]) * np.float32(sc["fx"])) / (np.float64(kl["x"][i]) - xs))
            nm += 1
    return depth, match, nm


# ---- two-view triangulation ------------------------------------------------------------------------------------
def triangulate_py(sc, max_chi2=5.998, scale_ratio_factor=0.0, g2f_train=None):
    """ucoslam::Triangulate (src/basictypes/misc.cpp:921-1040) restated on numpy float32 with OpenCV's own float SVD through cv2
    (cv2.SVDecomp with MODIFY_A | FULL_UV, the call of :931).  Returns (xyz f32 (n,3) NaN = rejected, n_good, margin (n,)): margin is
    the distance of the closest gate to its threshold (relative), so that tests can leave out matches a float rounding could flip."""
    import cv2
    f32 = np.float32
    k1, k2, m = sc["kps_train"], sc["kps_query"], sc["matches"]
    fx1, fy1, cx1, cy1 = [f32(x) for x in sc["K_train"]]
    fx2, fy2, cx2, cy2 = [f32(x) for x in sc["K_query"]]
    RT = np.asarray(sc["RT"], f32)
    R, t = RT[:3, :3], RT[:3, 3]
    inv1 = [f32(1) / (f32(s) * f32(s)) for s in sc["sf_train"]]
    inv2 = [f32(1) / (f32(s) * f32(s)) for s in sc["sf_query"]]
    P1 = np.zeros((3, 4), f32); P1[0, 0], P1[1, 1], P1[0, 2], P1[1, 2], P1[2, 2] = fx1, fy1, cx1, cy1, 1
    K2 = np.array([[fx2, 0, cx2], [0, fy2, cy2], [0, 0, 1]], f32)
    P2 = (K2.astype(np.float64) @ np.c_[R, t].astype(np.float64)).astype(f32)
    n = len(m)
    out = np.full((n, 3), np.nan, f32)
    margin = np.full(n, np.inf)
    good = 0
    for i in range(n):
        a, b = k1[m["trainIdx"][i]], k2[m["queryIdx"][i]]
        x1 = np.array([(a["x"] - cx1) * (f32(1) / fx1), (a["y"] - cy1) * (f32(1) / fy1), 1], f32)
        x2 = np.array([(b["x"] - cx2) * (f32(1) / fx2), (b["y"] - cy2) * (f32(1) / fy2), 1], f32)
        r1 = (x1 * f32(1.0 / np.linalg.norm(x1.astype(np.float64)))).astype(f32)
        r2 = (R.T.astype(np.float64) @ (x2 * f32(1.0 / np.linalg.norm(x2.astype(np.float64)))).astype(np.float64)).astype(f32)
        cosp = float(r1.astype(np.float64) @ r2.astype(np.float64))
        margin[i] = min(abs(cosp - 0.9998) / 2e-6, abs(cosp) / 2e-6)
        if cosp < 0 or cosp > 0.9998:
            continue
        A = np.empty((4, 4), f32)
        A[0] = a["x"] * P1[2] - P1[0]; A[1] = a["y"] * P1[2] - P1[1]
        A[2] = b["x"] * P2[2] - P2[0]; A[3] = b["y"] * P2[2] - P2[1]
        w_, u_, vt = cv2.SVDecomp(A, flags=cv2.SVD_MODIFY_A | cv2.SVD_FULL_UV)
        x = vt[3].astype(f32)
        if x[3] == 0:
            continue
        p = (x[:3] / x[3]).astype(f32)
        if not np.isfinite(p).all():
            continue
        margin[i] = min(margin[i], abs(float(p[2])) / 1e-4)
        if p[2] <= 0:
            continue
        p2 = ((R.astype(np.float64) @ p.astype(np.float64)).astype(f32) + t).astype(f32)
        margin[i] = min(margin[i], abs(float(p2[2])) / 1e-4)
        if p2[2] <= 0:
            continue
        iz = f32(1) / p[2]
        px, py = fx1 * p[0] * iz + cx1, fy1 * p[1] * iz + cy1
        chi = inv1[a["octave"]] * ((px - a["x"]) * (px - a["x"]) + (py - a["y"]) * (py - a["y"]))
        margin[i] = min(margin[i], abs(float(chi) - max_chi2) / (0.005 * max_chi2))
        if chi > f32(max_chi2):
            continue
        iz2 = f32(1) / p2[2]
        qx, qy = fx2 * p2[0] * iz2 + cx2, fy2 * p2[1] * iz2 + cy2
        chi = inv2[b["octave"]] * ((qx - b["x"]) * (qx - b["x"]) + (qy - b["y"]) * (qy - b["y"]))
        margin[i] = min(margin[i], abs(float(chi) - max_chi2) / (0.005 * max_chi2))
        if chi > f32(max_chi2):
            continue
        if scale_ratio_factor:      # the mapper's scale-consistency test (mapmanager.cpp new-point creation), in global coordinates
            G = np.asarray(np.eye(4) if g2f_train is None else g2f_train, f32)
            Pg = np.array([G[r, 0] * p[0] + G[r, 1] * p[1] + G[r, 2] * p[2] + G[r, 3] for r in range(3)], f32)
            c1 = G[:3, 3]                                            # train camera centre
            Rq, tq = RT[:3, :3], RT[:3, 3]
            cq_train = -(Rq.T.astype(np.float64) @ tq.astype(np.float64)).astype(f32)   # query centre in train coordinates
            c2 = np.array([G[r, 0] * cq_train[0] + G[r, 1] * cq_train[1] + G[r, 2] * cq_train[2] + G[r, 3] for r in range(3)], f32)
            d1 = f32(np.sqrt(((Pg - c1).astype(np.float64) ** 2).sum())); d2 = f32(np.sqrt(((Pg - c2).astype(np.float64) ** 2).sum()))
            if d1 == 0 or d2 == 0:
                continue
            rd = d1 / d2
            ro = f32(sc["sf_train"][a["octave"]]) / f32(sc["sf_query"][b["octave"]])
            fct = f32(scale_ratio_factor)
            margin[i] = min(margin[i], abs(float(rd * fct) - float(ro)) / (1e-4 * float(ro)), abs(float(rd) - float(ro * fct)) / (1e-4 * float(ro * fct)))
            if rd * fct < ro or rd > ro * fct:
                continue
        if g2f_train is not None:
            G = np.asarray(g2f_train, f32)
            p = np.array([G[r, 0] * p[0] + G[r, 1] * p[1] + G[r, 2] * p[2] + G[r, 3] for r in range(3)], f32)
        out[i] = p
        good += 1
    return out, good, margin


def std_sort_indices(keys):
    """the permutation std::sort (host libstdc++, through oracle/stl_helper.cpp) applies to 0..n-1 under keys[a] < keys[b]"""
    keys = np.ascontiguousarray(keys, np.float32)
    idx = np.arange(len(keys), dtype=np.uint32)
    load_stl().stl_sort_indices(_p(idx), len(idx), _p(keys))
    return idx.astype(np.int64)


def new_points_py(sc, max_points=-1):
    """MapManager::createNewPoints (src/utils/mapmanager.cpp:9772-10788, de-obfuscated with the C preprocessor) restated over the
    restatements of its callees: per neighbour f matchEpipolar (frame_match: oracle/match_oracle.c, pinned to the reference's own
    framematcher.cpp) on the MODE_UNASSIGNED rows, Triangulate + scale consistency + global frame (triangulate_py), then the merge:
    std::map keyed by the keyframe keypoint, elements in neighbour order, position / distance of the LAST element (the reference's
    minimum-octave loop never updates its minimum), std::sort on the distance + resize above max_points (through the host library's
    own std::sort: oracle/stl_helper.cpp, so ties fall the same way).
    Returns dict(kpt, xyz, dist, obs_ptr, obs_frame, obs_kpt, matches (per neighbour), xyz_pairs, margin (per neighbour))."""
    groups = {}
    matches, xyzs, margins = [], [], []
    for f in range(len(sc["q_desc"])):
        qm, tm = np.asarray(sc["q_map"][f], np.int32), np.asarray(sc["t_map"], np.int32)
        if len(qm) and len(tm):
            m = frame_match(np.ascontiguousarray(sc["q_desc"][f][qm]), sc["q_kps"][f], np.ascontiguousarray(sc["t_desc"][tm]), sc["t_kps"],
                            min_desc_dist=sc["min_desc_dist"], ratio=sc["ratio"], check_orientation=True, max_octave_diff=2 ** 31 - 1, F12=sc["f12"][f],
                            scale_factors=sc["sf_nb"], q_map=qm, t_map=tm)
        else:
            m = np.zeros(0, MATCH_DTYPE)
        tv = dict(kps_train=sc["t_kps"], kps_query=sc["q_kps"][f], matches=m, K_train=sc["K_kf"], K_query=sc["K_nb"][f], RT=sc["rt"][f],
                  sf_train=sc["sf_kf"], sf_query=sc["sf_nb"])
        xyz, _, margin = triangulate_py(tv, sc.get("max_chi2", 5.998), sc["scale_ratio_factor"], sc["g2f_kf"])
        matches.append(m); xyzs.append(xyz); margins.append(margin)
        for i in range(len(m)):
            if not np.isnan(xyz[i, 0]):
                groups.setdefault(int(m["trainIdx"][i]), []).append((f, int(m["queryIdx"][i]), float(m["distance"][i]), xyz[i]))
    keys = sorted(groups)
    if max_points >= 0 and len(keys) > max_points:
        d = np.array([groups[k][-1][2] for k in keys], np.float32)
        order = std_sort_indices(d)[:max_points]
        keys = [keys[i] for i in order]
    kpt = np.array(keys, np.int32)
    xyz = np.array([groups[k][-1][3] for k in keys], np.float32).reshape(-1, 3)
    dist = np.array([groups[k][-1][2] for k in keys], np.float32)
    ptr = np.zeros(len(keys) + 1, np.int32)
    ptr[1:] = np.cumsum([len(groups[k]) for k in keys])
    ofr = np.array([r[0] for k in keys for r in groups[k]], np.int32)
    okp = np.array([r[1] for k in keys for r in groups[k]], np.int32)
    return dict(kpt=kpt, xyz=xyz, dist=dist, obs_ptr=ptr, obs_frame=ofr, obs_kpt=okp, matches=matches, xyz_pairs=xyzs, margin=margins)


# ---- RANSAC P3P pose -------------------------------------------------------------------------------------------
def pnp_ransac_py(sc, samples):
    """PnPSolver::solvePnPRansac (src/optimization/pnpsolver.cpp:36-114) restated with the reference's own OpenCV call
    (cv2.solvePnP, SOLVEPNP_P3P, 4 points) on GIVEN 4-match samples (the reference draws them with std::random_shuffle) and its float
    inlier test.  Returns dict(ok, pose44 f32, inliers, counts (per iteration, -1 = no P3P solution), best_iter,
    borderline (per iteration: matches whose test is within rounding of a threshold))."""
    import cv2
    f32 = np.float32
    p3 = np.asarray(sc["p3d"], f32); p2 = np.asarray(sc["p2d"], f32); nr = np.asarray(sc["normals"], f32)
    fx, fy, cx, cy = [f32(x) for x in sc["cam"]]
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], f32)
    n = len(p3)
    counts = np.full(len(samples), -1, np.int32); border = np.zeros(len(samples), np.int32)
    best, best_inl, best_pose, best_it = -1, None, None, -1
    p3d64, p2d64 = p3.astype(np.float64), p2.astype(np.float64)
    for it, smp in enumerate(samples):
        ok, rv, tv = cv2.solvePnP(p3d64[smp].reshape(-1, 1, 3), p2d64[smp].reshape(-1, 1, 2), Km, np.zeros((1, 5), f32),
                                  flags=cv2.SOLVEPNP_P3P)
        if not ok or rv.size != 3 or tv.size != 3:
            continue
        M = np.eye(4, dtype=f32)
        M[:3, :3] = cv2.Rodrigues(rv.astype(f32).reshape(3, 1))[0].astype(f32)      # Se3Transform(rv, tv): float 4x4
        M[:3, 3] = tv.astype(f32).ravel()
        x = M[0, 0] * p3[:, 0] + M[0, 1] * p3[:, 1] + M[0, 2] * p3[:, 2] + M[0, 3]
        y = M[1, 0] * p3[:, 0] + M[1, 1] * p3[:, 1] + M[1, 2] * p3[:, 2] + M[1, 3]
        z = M[2, 0] * p3[:, 0] + M[2, 1] * p3[:, 1] + M[2, 2] * p3[:, 2] + M[2, 3]
        with np.errstate(all="ignore"):
            z = (1.0 / z.astype(np.float64)).astype(f32)
            rx = (((fx * x) * z) + cx).astype(np.float64); ry = (((fy * y) * z) + cy).astype(np.float64)
            dx = (p2d64[:, 0] - rx).astype(f32); dy = (p2d64[:, 1] - ry).astype(f32)
            d2 = dx * dx + dy * dy
            c = -np.array([M[0, 3] * M[0, 0] + M[1, 3] * M[1, 0] + M[2, 3] * M[2, 0], M[0, 3] * M[0, 1] + M[1, 3] * M[1, 1] + M[2, 3] * M[2, 1],
                           M[0, 3] * M[0, 2] + M[1, 3] * M[1, 2] + M[2, 3] * M[2, 2]], f32)
            v = (c[None, :] - p3).astype(f32)
            inv = 1.0 / np.sqrt((v.astype(np.float64) ** 2).sum(axis=1))
            v = (v.astype(np.float64) * inv[:, None]).astype(f32)
            vc = v[:, 0] * nr[:, 0] + v[:, 1] * nr[:, 1] + v[:, 2] * nr[:, 2]
        inl = (d2 < f32(5.99)) & ~(vc < f32(0.5))
        counts[it] = int(inl.sum())
        border[it] = int(((np.abs(d2 - 5.99) < 0.02) | ((d2 < 5.99) & (np.abs(vc - 0.5) < 1e-4))).sum())
        if counts[it] > best:
            best, best_inl, best_pose, best_it = counts[it], np.nonzero(inl)[0].astype(np.int32), M, it
    ok = best >= 4
    return dict(ok=ok, pose44=best_pose if ok else None, inliers=best_inl if ok else np.zeros(0, np.int32), counts=counts,
                best_iter=best_it if ok else -1, borderline=border)


# ---- point undistortion ----------------------------------------------------------------------------------------
def undistort_points_py(pts, K, dist):
    """ucoslam::undistortPoints (src/basictypes/misc.cpp:269-292) with the reference's own OpenCV call made through cv2:
    cv2.undistortPoints (default termination), then x*fx+cx in float.  pts (n,2) f32, K = fx fy cx cy, dist = coefficients."""
    import cv2
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    K = np.asarray(K, np.float32)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    d = np.asarray(dist, np.float32).reshape(-1)
    if len(pts) == 0:
        return pts.copy()
    n = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, d if len(d) else None).reshape(-1, 2).astype(np.float32)
    return np.c_[n[:, 0] * K[0] + K[2], n[:, 1] * K[1] + K[3]].astype(np.float32)
