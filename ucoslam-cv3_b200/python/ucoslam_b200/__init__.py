"""ctypes binding of libucoslam_b200.so (the C ABI declared in include/ucoslam_b200.h).

Used by tests/, bench.py and __graft_entry__.py only: the product is the shared library plus the C++ adapters in
ucoslam-cv3_b200/host/.  There is NO CPU fallback here: if the library is missing or no CUDA device can be opened the
calls raise.
"""
import ctypes, os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libucoslam_b200.so")

UCO_KNN_HEAP, UCO_KNN_SORTED = 0, 1
UCO_KDTREE_DEV_MAX_POINTS = 4096
UCO_TRACK_NO_SYNC = 1
_c = ctypes
_vp, _i, _sz, _u64 = _c.c_void_p, _c.c_int, _c.c_size_t, _c.c_uint64

# name -> (restype, argtypes); must list every symbol include/ucoslam_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "uco_b200_create": (_vp, [_i, _i]),
    "uco_b200_destroy": (None, [_vp]),
    "uco_b200_last_error": (_c.c_char_p, [_vp]),
    "uco_b200_stream": (_vp, [_vp]),
    "uco_b200_sync": (_i, [_vp]),
    "uco_b200_launch_count": (_u64, [_vp]),
    "uco_b200_version": (_i, []),
    "uco_b200_hamming_knn": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _vp, _vp]),
    "uco_b200_hamming_knn_dev": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "uco_b200_hamming_knn_batch_dev": (_i, [_vp, _i, _vp, _sz, _i, _vp, _vp, _sz, _i, _vp, _i, _i, _vp, _vp]),
    "uco_b200_hamming_knn_batch": (_i, [_vp, _i, _vp, _vp, _sz, _vp, _vp, _sz, _i, _i, _vp, _vp]),
    "uco_b200_hamming_knn_sharded_dev": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "uco_b200_knn_merge_dev": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "uco_b200_frame_match_bow": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _vp, _vp, _sz, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "uco_b200_kdtree_build": (_i, [_vp, _sz, _i, _vp, _i, _vp, _vp, _vp]),
    "uco_b200_kdtree_parse": (_i, [_vp, _sz, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "uco_b200_match_projected": (_i, [_vp, _vp, _vp, _vp, _c.c_float, _c.c_float, _vp, _vp, _vp]),
    "uco_b200_kdtree_build_batch_dev": (_i, [_vp, _i, _vp, _sz, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "uco_b200_kdtree_build_dev": (_i, [_vp, _vp, _sz, _i, _vp, _i, _vp, _vp, _vp]),
    "uco_b200_probe_sort_indices": (_i, [_vp, _i, _vp]),
    "uco_b200_track_batch_dev": (_i, [_vp, _vp, _vp, _vp]),
    "uco_b200_track_batch": (_i, [_vp, _vp, _vp, _vp]),
    "uco_b200_track_state_create": (_i, [_vp, _i, _i, _i, _vp]),
    "uco_b200_track_state_free": (None, [_vp, _vp]),
    "uco_b200_track_state_set_prev": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "uco_b200_track_state_set_map": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "uco_b200_track_state_step_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i]),
    "uco_b200_track_frames": (_i, [_vp, _vp, _vp, _i, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "uco_b200_track_projected": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _c.c_float, _c.c_float, _vp, _vp]),
    "uco_b200_ba_set_host_threads": (_i, [_vp, _i]),
    "uco_b200_comm_unique_id": (_i, [_vp]),
    "uco_b200_comm_create": (_i, [_vp, _vp, _i, _i, _vp]),
    "uco_b200_comm_destroy": (None, [_vp]),
    "uco_b200_ba_solve_sharded": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "uco_b200_probe_ba_partition": (_i, [_vp, _i, _vp]),
    "uco_b200_frame_stream_parse": (_i, [_vp, _c.c_size_t, _vp, _vp]),
    "uco_b200_frame_stream_write": (_i, [_vp, _vp, _c.c_size_t, _vp]),
    "uco_b200_mappoint_stream_parse": (_i, [_vp, _c.c_size_t, _vp, _vp]),
    "uco_b200_mappoint_stream_write": (_i, [_vp, _vp, _c.c_size_t, _vp]),
    "uco_b200_mappoint_container_walk": (_i, [_vp, _c.c_size_t, _vp, _vp, _vp, _c.c_uint32, _vp]),
    "uco_b200_frame_container_walk": (_i, [_vp, _c.c_size_t, _vp, _vp, _vp, _c.c_uint32, _vp]),
    "uco_b200_kfdb_stream_walk": (_i, [_vp, _c.c_size_t, _vp, _vp]),
    "uco_b200_marker_map_walk": (_i, [_vp, _c.c_size_t, _c.c_uint32, _vp, _vp, _vp]),
    "uco_b200_covis_stream_walk": (_i, [_vp, _c.c_size_t, _vp, _vp]),
    "uco_b200_map_stream_walk": (_i, [_vp, _c.c_size_t, _i, _vp]),
    "uco_b200_frame_container_write": (_i, [_vp, _vp, _vp, _vp, _vp, _c.c_size_t, _vp]),
    "uco_b200_marker_map_write": (_i, [_vp, _c.c_uint32, _vp, _c.c_size_t, _vp]),
    "uco_b200_covis_stream_unpack": (_i, [_vp, _c.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp]),
    "uco_b200_covis_stream_write": (_i, [_c.c_uint32, _vp, _c.c_uint32, _vp, _vp, _vp, _c.c_uint32, _vp, _vp, _vp, _c.c_size_t, _vp]),
    "uco_b200_kfdb_stream_unpack": (_i, [_vp, _c.c_size_t, _vp, _vp, _vp, _vp]),
    "uco_b200_kfdb_stream_write": (_i, [_c.c_int32, _vp, _c.c_size_t, _c.c_uint32, _vp, _vp, _vp, _c.c_uint32, _vp, _vp, _c.c_size_t, _vp]),
    "uco_b200_mappoints_from_container": (_i, [_vp, _c.c_size_t, _c.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "uco_b200_mappoint_container_write": (_i, [_vp, _vp, _vp, _vp, _c.c_size_t, _vp]),
    "uco_b200_mappoint_stream_default": (None, [_vp]),
    "uco_b200_kdtree_serialize": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _c.c_size_t, _vp]),
    "uco_b200_frame_upload": (_i, [_vp, _vp, _vp]),
    "uco_b200_frame_dev": (_vp, [_vp]),
    "uco_b200_frame_download": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "uco_b200_frame_free": (None, [_vp]),
    "uco_b200_new_points": (_i, [_vp] + [_vp, _i, _c.c_size_t, _vp, _i, _vp] + [_i, _vp, _vp, _c.c_size_t, _vp, _vp, _vp] + [_vp, _vp, _vp, _vp] + [_vp] * 7 + [_i, _i] + [_vp, _vp, _vp]),
    "uco_b200_block_solve": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "uco_b200_block_solve_profile": (_i, [_vp, _vp]),
    "uco_b200_probe_block_solve": (_i, [_i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "uco_b200_orb_default_params": (None, [_vp]),
    "uco_b200_orb_extract": (_i, [_vp, _vp, _i, _i, _sz, _vp, _vp, _vp, _i, _vp]),
    "uco_b200_orb_extract_batch": (_i, [_vp, _vp, _i, _i, _i, _sz, _vp, _vp, _vp, _i, _vp]),
    "uco_b200_orb_extract_batch_dev": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _vp, _vp, _vp, _vp]),
    "uco_b200_set_profiling": (None, [_vp, _i]),
    "uco_b200_orb_last_stage_ms": (_i, [_vp, _vp]),
    "uco_b200_orb_plan_bytes": (_i, [_vp, _vp]),
    "uco_b200_orb_debug_level_info": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "uco_b200_orb_debug_pyramid": (_i, [_vp, _i, _i, _vp]),
    "uco_b200_orb_debug_selected": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "uco_b200_orb_debug_candidates": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "uco_b200_bow_load": (_i, [_vp, _vp, _sz, _vp]),
    "uco_b200_bow_free": (None, [_vp, _vp]),
    "uco_b200_bow_info": (_i, [_vp, _vp, _vp, _vp]),
    "uco_b200_bow_transform": (_i, [_vp, _vp, _vp, _i, _sz, _i, _vp, _vp, _vp]),
    "uco_b200_bow_transform_dev": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "uco_b200_ba_solve": (_i, [_vp, _vp, _vp, _vp]),
    "uco_b200_ba_solve_batch": (_i, [_vp, _i, _vp, _vp, _vp]),
    "uco_b200_ba_set_mode": (_i, [_vp, _i, _i]),
    "uco_b200_probe_ba_plan": (_i, [_vp, _i, _vp]),
    "uco_b200_pose_only": (_i, [_vp, _vp, _vp]),
    "uco_b200_frame_match": (_i, [_vp, _vp, _i, _sz, _vp, _i, _vp, _vp, _i, _sz, _vp, _i, _vp, _vp, _vp, _i, _vp]),
    "uco_b200_frame_match_batch_dev": (_i, [_vp, _i, _vp, _sz, _vp, _sz, _i, _vp, _vp, _sz, _vp, _sz, _i, _vp, _vp, _vp, _vp]),
    "uco_b200_pose_only_batch": (_i, [_vp, _i, _vp, _vp]),
    "uco_b200_keyframes_batch_dev": (_i, [_vp, _vp, _i, _vp, _sz, _vp, _sz, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "uco_b200_keyframes_batch": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "uco_b200_frame_match_multi": (_i, [_vp, _vp, _i, _sz, _vp, _i, _vp, _i, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "uco_b200_probe_math": (_i, [_i, _vp, _vp, _i, _vp, _vp]),
    "uco_b200_probe_retain_best": (_i, [_vp, _i, _i]),
    "uco_b200_stereo_depth": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _vp, _sz, _i, _vp, _vp, _sz, _i, _c.c_float, _c.c_float,
                                   _c.c_float, _vp, _vp, _vp]),
    "uco_b200_stereo_depth_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _c.c_float, _c.c_float, _c.c_float,
                                       _vp, _vp, _vp]),
    "uco_b200_triangulate": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "uco_b200_pnp_ransac": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "uco_b200_probe_p3p": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "uco_b200_undistort_points": (_i, [_vp, _vp, _sz, _i, _vp, _vp, _i, _vp, _sz]),
    "uco_b200_undistort_points_dev": (_i, [_vp, _vp, _sz, _i, _vp, _vp, _i, _vp, _sz]),
    "uco_b200_probe_undistort": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "uco_b200_kfdb_create": (_i, [_vp, _vp]),
    "uco_b200_kfdb_free": (None, [_vp, _vp]),
    "uco_b200_kfdb_clear": (_i, [_vp, _vp]),
    "uco_b200_kfdb_size": (_i, [_vp, _vp, _vp]),
    "uco_b200_kfdb_has": (_i, [_vp, _c.c_uint32]),
    "uco_b200_kfdb_add": (_i, [_vp, _vp, _c.c_uint32, _vp, _vp, _i]),
    "uco_b200_kfdb_add_batch": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "uco_b200_kfdb_del": (_i, [_vp, _vp, _c.c_uint32]),
    "uco_b200_kfdb_query": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _c.c_float, _vp, _vp, _vp, _i, _vp, _vp]),
    "uco_b200_kfdb_last_ms": (_i, [_vp, _vp]),
    "uco_b200_kfdb_rank": (_i, [_vp, _vp, _i, _vp, _vp, _i, _c.c_float, _vp, _vp]),
}

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])  # uco_keypoint == cv::KeyPoint


class OrbParams(ctypes.Structure):
    _fields_ = [("max_features", _c.c_int32), ("n_levels", _c.c_int32), ("scale_factor", _c.c_float),
                ("ini_th_fast", _c.c_int32), ("min_th_fast", _c.c_int32), ("blur_first", _c.c_int32)]

    def __init__(self, max_features=2000, n_levels=8, scale_factor=1.2, ini_th=20, min_th=7, blur_first=True):
        super().__init__(max_features, n_levels, scale_factor, ini_th, min_th, int(blur_first))


def probe_fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    assert load().uco_b200_probe_math(0, _p(y), _p(x), len(y), _p(out), None) == 0
    return out


def probe_sincos(a):
    a = np.ascontiguousarray(a, np.float32)
    s = np.empty_like(a); c = np.empty_like(a)
    assert load().uco_b200_probe_math(1, _p(a), None, len(a), _p(s), _p(c)) == 0
    return s, c


def probe_undistort(pts, K, dist):
    """host-only: ucoslam::undistortPoints on (n,2) f32 pixel coordinates as the kernel computes it"""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2); K = np.ascontiguousarray(K, np.float32)
    dist = np.ascontiguousarray(dist, np.float32).reshape(-1)
    out = np.empty_like(pts)
    if load().uco_b200_probe_undistort(_p(pts), len(pts), _p(K), _p(dist), len(dist), _p(out)) != 0:
        raise UcoError("probe_undistort failed")
    return out


def probe_p3p(X4, px4, K):
    """host-only: (R (3,3), t (3,)) of the hypothesis the RANSAC kernel forms from 4 correspondences, or None"""
    X4 = np.ascontiguousarray(X4, np.float64); px4 = np.ascontiguousarray(px4, np.float64); K = np.ascontiguousarray(K, np.float64)
    R = np.zeros(9, np.float64); t = np.zeros(3, np.float64)
    ok = load().uco_b200_probe_p3p(_p(X4), _p(px4), _p(K), _p(R), _p(t))
    return (R.reshape(3, 3), t) if ok else None


def probe_retain_best(packed, n_points):
    packed = np.ascontiguousarray(packed, np.uint32).copy()
    n = load().uco_b200_probe_retain_best(_p(packed), len(packed), int(n_points))
    return packed[:n]


_lib = None


KDNODE_DTYPE = np.dtype([("divlow", "<f4"), ("divhigh", "<f4"), ("col", "<i4"), ("left", "<i4"), ("right", "<i4"),
                         ("leaf_begin", "<i4"), ("leaf_count", "<i4")])  # uco_kdnode


class MapPoints(ctypes.Structure):  # uco_mappoints
    _fields_ = [("n", _c.c_int32), ("ids", _vp), ("pos", _vp), ("normal", _vp), ("min_dist", _vp), ("max_dist", _vp), ("desc", _vp)]


class FrameView(ctypes.Structure):  # uco_frame_view
    _fields_ = [("n_kp", _c.c_int32), ("kps", _vp), ("desc", _vp), ("desc_stride", _c.c_size_t), ("n_nodes", _c.c_int32), ("nodes", _vp),
                ("leaf_idx", _vp), ("bbox", _c.c_double * 4), ("n_levels", _c.c_int32), ("scale_factors", _vp), ("fx", _c.c_float),
                ("fy", _c.c_float), ("cx", _c.c_float), ("cy", _c.c_float), ("min_xy", _c.c_float * 2), ("max_xy", _c.c_float * 2)]


def kdtree_build(xy):
    """host-side uco_b200_kdtree_build over (n,2) f32 points -> (nodes[KDNODE_DTYPE], leaf_idx i32, bbox f64[4])"""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    n = len(xy)
    nodes, leaf, bbox, k = np.zeros(2 * n + 2, KDNODE_DTYPE), np.zeros(max(n, 1), np.int32), np.zeros(4), ctypes.c_int(0)
    if load().uco_b200_kdtree_build(_p(xy), 8, n, _p(nodes), len(nodes), _p(leaf), _p(bbox), ctypes.byref(k)) != 0:
        raise UcoError("kdtree_build failed")
    return nodes[:k.value].copy(), leaf[:n].copy(), bbox


def probe_sort_indices(keys):
    """libstdc++ std::sort replay (csrc/sort_exact.h) of the indices 0..n-1 by keys[idx]"""
    keys = np.ascontiguousarray(keys, np.float32)
    idx = np.arange(len(keys), dtype=np.uint32)
    rc = load().uco_b200_probe_sort_indices(_p(idx), len(keys), _p(keys))
    assert rc == 0
    return idx


class TrackParams(ctypes.Structure):  # uco_track_params
    _fields_ = [("max_desc_dist", ctypes.c_float), ("proj_dist_thr", ctypes.c_float), ("fx", ctypes.c_float), ("fy", ctypes.c_float),
                ("cx", ctypes.c_float), ("cy", ctypes.c_float), ("bf", ctypes.c_float), ("min_xy", ctypes.c_float * 2),
                ("max_xy", ctypes.c_float * 2), ("n_levels", ctypes.c_int32), ("scale_factors", ctypes.c_float * 32)]

    def __init__(self, sc=None, max_desc_dist=50.0, proj_dist_thr=15.0):
        super().__init__()
        self.max_desc_dist, self.proj_dist_thr = max_desc_dist, proj_dist_thr
        if sc is not None:
            self.fx, self.fy, self.cx, self.cy, self.bf = sc["fx"], sc["fy"], sc["cx"], sc["cy"], sc.get("bf", 0.0)
            self.min_xy[:] = [float(v) for v in sc["min_xy"]]
            self.max_xy[:] = [float(v) for v in sc["max_xy"]]
            sf = np.asarray(sc["scale_factors"], np.float32)
            self.n_levels = len(sf)
            for i, v in enumerate(sf):
                self.scale_factors[i] = float(v)


class TrackBatch(ctypes.Structure):  # uco_track_batch
    _fields_ = [("n_frames", ctypes.c_int32), ("kp_cap", ctypes.c_int32), ("prev_cap", ctypes.c_int32), ("map_cap", ctypes.c_int32),
                ("flags", ctypes.c_int32)] + [(k, ctypes.c_void_p) for k in (
                    "kps", "desc", "n_kp", "depth", "prev_kps", "prev_desc", "prev_n_kp", "prev_mp_row", "map_n", "mp_id", "mp_pos",
                    "mp_normal", "mp_min_dist", "mp_max_dist", "mp_desc", "mp_stable", "mp_local", "pose_prior")]


class TrackOut(ctypes.Structure):  # uco_track_out
    _fields_ = [(k, ctypes.c_void_p) for k in ("matches", "n_matches", "pose", "n_good", "status", "n_tbp", "visible")]


def kdtree_parse(stream_bytes):
    """host-side uco_b200_kdtree_parse of KdTreeIndex::toStream bytes -> (nodes, leaf_idx, bbox)"""
    buf = np.frombuffer(stream_bytes, np.uint8)
    cap = len(buf) // 4 + 4
    nodes, leaf, bbox = np.zeros(cap, KDNODE_DTYPE), np.zeros(cap, np.int32), np.zeros(4)
    k, nl = ctypes.c_int(0), ctypes.c_int(0)
    if load().uco_b200_kdtree_parse(_p(buf), len(buf), _p(nodes), cap, _p(leaf), cap, _p(bbox), ctypes.byref(k), ctypes.byref(nl)) != 0:
        raise UcoError("kdtree_parse failed")
    return nodes[:k.value].copy(), leaf[:nl.value].copy(), bbox


def probe_block_solve(nb, blk_ij, blocks, rhs, force_k=-1, smem_optin=0, solve=True):
    """host-only: planner + plain host execution of the two-level block-envelope Cholesky; returns (x or None, info8)"""
    blk_ij = np.ascontiguousarray(blk_ij, np.int32).reshape(-1, 2)
    blocks = np.ascontiguousarray(blocks, np.float64).reshape(-1, 36)
    rhs = np.ascontiguousarray(rhs, np.float64).reshape(-1)
    x = np.zeros(6 * nb)
    info = np.zeros(8, np.int32)
    rc = load().uco_b200_probe_block_solve(nb, len(blk_ij), _p(blk_ij), _p(blocks), _p(rhs), smem_optin, force_k, _p(x) if solve else None, _p(info))
    if rc < 0:
        raise UcoError("probe_block_solve: bad input (%d)" % rc)
    return (x if solve and rc == 0 else None), info


def probe_ba_partition(pb, world):
    """host-only: (boundaries[world+1], observations per rank[world]) of the sharded solver's landmark partition"""
    cp, cr, keep, out = Context.ba_pack(pb, 1)
    o = np.zeros(2 * world + 1, np.int32)
    if load().uco_b200_probe_ba_partition(ctypes.addressof(cp), world, _p(o)) != 0:
        raise UcoError("probe_ba_partition failed")
    return o[:world + 1].copy(), o[world + 1:].copy()


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libucoslam_b200.so not built: run python ucoslam-cv3_b200/build.py (%s)" % LIB_PATH)
        lib = ctypes.CDLL(os.environ.get("UCO_B200_LIB", LIB_PATH))   # the override loads a build variant for A/B measurements
        for name, (res, args) in SIGNATURES.items():
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


class BaProblem(ctypes.Structure):  # uco_ba_problem
    _fields_ = [("n_poses", _c.c_int32), ("n_points", _c.c_int32), ("n_obs", _c.c_int32), ("poses44", _vp), ("fixed", _vp),
                ("points3", _vp), ("obs_pose", _vp), ("obs_point", _vp), ("obs_uv", _vp), ("obs_ur", _vp), ("obs_stereo", _vp),
                ("obs_inv_sigma2", _vp), ("fx", _c.c_float), ("fy", _c.c_float), ("cx", _c.c_float), ("cy", _c.c_float),
                ("bf", _c.c_float), ("n_iters", _c.c_int32), ("n_markers", _c.c_int32), ("marker_pose44", _vp), ("marker_size", _vp),
                ("n_marker_obs", _c.c_int32), ("mobs_marker", _vp), ("mobs_pose", _vp), ("mobs_corners", _vp), ("mobs_weight", _vp),
                ("pose_cam", _vp), ("n_plane", _c.c_int32), ("plane_ref", _c.c_int32), ("plane_ref_pose44", _vp), ("plane_other", _vp),
                ("plane_weight", _c.c_double)]


class BaResult(ctypes.Structure):  # uco_ba_result
    _fields_ = [("pose7", _vp), ("poses44", _vp), ("points3", _vp), ("obs_chi2", _vp), ("obs_level", _vp), ("obs_bad", _vp),
                ("trace", _vp), ("iters", _c.c_int32 * 2), ("device_ms", _c.c_float), ("profile", _vp), ("marker_poses44", _vp),
                ("marker_pose7", _vp), ("mobs_chi2", _vp)]


MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])  # uco_match == cv::DMatch


class MatchParams(ctypes.Structure):  # uco_match_params
    _fields_ = [("min_desc_dist", _c.c_float), ("nn_match_ratio", _c.c_float), ("check_orientation", _c.c_int32),
                ("max_octave_diff", _c.c_int32), ("use_f12", _c.c_int32), ("f12", _c.c_float * 9), ("n_scales", _c.c_int32),
                ("scale_factors", _c.c_float * 32)]

    def __init__(self, min_desc_dist=50.0, ratio=0.8, check_orientation=True, max_octave_diff=1, F12=None, scale_factors=None):
        super().__init__(min_desc_dist, ratio, int(check_orientation), max_octave_diff, int(F12 is not None))
        if F12 is not None:
            self.f12 = (_c.c_float * 9)(*[float(v) for v in np.asarray(F12, np.float32).reshape(9)])
        sf = np.asarray(scale_factors if scale_factors is not None else [np.float32(1.2) ** i for i in range(8)], np.float32)
        self.n_scales = len(sf)
        self.scale_factors = (_c.c_float * 32)(*([float(v) for v in sf] + [1.0] * (32 - len(sf))))


class BowIndex(ctypes.Structure):  # uco_bow_index
    _fields_ = [("n_nodes", _c.c_int32), ("node_id", _vp), ("ptr", _vp), ("kp", _vp)]


def bow_index(level_node, usable=None):
    """a frame's fBow2 flattened in std::map order from the per-keypoint level-3 node ids (what uco_b200_bow_transform reports):
    (node_id u32 ascending, ptr i32, kp i32 in keypoint order inside a node)"""
    level_node = np.asarray(level_node, np.uint32)
    order = np.argsort(level_node, kind="stable").astype(np.int32)
    ids, counts = np.unique(level_node, return_counts=True)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return ids.astype(np.uint32), ptr, order


class PnpProblem(ctypes.Structure):  # uco_pnp_problem
    _fields_ = [("n_matches", _c.c_int32), ("pose44", _vp), ("points3", _vp), ("obs_uv", _vp), ("obs_ur", _vp), ("obs_stereo", _vp),
                ("obs_inv_sigma2", _vp), ("stable", _vp), ("fx", _c.c_float), ("fy", _c.c_float), ("cx", _c.c_float),
                ("cy", _c.c_float), ("bf", _c.c_float), ("n_markers", _c.c_int32), ("marker_pose44", _vp), ("marker_size", _vp),
                ("marker_corners", _vp)]


class PnpResult(ctypes.Structure):  # uco_pnp_result
    _fields_ = [("pose44", _c.c_float * 16), ("pose7", _c.c_double * 7), ("n_good", _c.c_int32), ("iters", _c.c_int32 * 4),
                ("bad", _vp)]


class TriangulateParams(ctypes.Structure):  # uco_triangulate_params
    _fields_ = [("K_train", _c.c_float * 4), ("K_query", _c.c_float * 4), ("RT", _c.c_float * 16), ("n_levels_train", _c.c_int32),
                ("scale_factors_train", _vp), ("n_levels_query", _c.c_int32), ("scale_factors_query", _vp), ("max_chi2", _c.c_float),
                ("scale_ratio_factor", _c.c_float), ("to_global", _c.c_int32), ("g2f_train", _c.c_float * 16)]


class MatView(ctypes.Structure):  # uco_mat_view
    _fields_ = [("rows", _c.c_int32), ("cols", _c.c_int32), ("type", _c.c_int32), ("data", _vp)]


class FrameStream(ctypes.Structure):  # uco_frame_stream: a view into a Frame::toStream byte range
    _fields_ = [("idx", _c.c_uint32), ("fseq_idx", _c.c_uint32), ("frame_flags", _c.c_uint8), ("kp_desc_type", _c.c_int8), ("desc", MatView),
                ("n_und_kpts", _c.c_uint32), ("und_kpts", _vp), ("n_kpts", _c.c_uint32), ("kpts", _vp), ("n_depth", _c.c_uint32), ("depth", _vp),
                ("n_ids", _c.c_uint32), ("ids", _vp), ("n_flags", _c.c_uint32), ("flags", _vp), ("n_markers", _c.c_uint32), ("markers", _vp),
                ("markers_bytes", _c.c_uint64), ("pose_f2g", _c.c_float * 16), ("n_bow", _c.c_uint32), ("bow", _vp), ("n_bow_level", _c.c_uint32),
                ("bow_level", _vp), ("bow_level_bytes", _c.c_uint64), ("n_scale_factors", _c.c_uint32), ("scale_factors", _vp),
                ("camera_matrix", MatView), ("distortion", MatView), ("cam_size", _c.c_int32 * 2), ("bl", _c.c_float), ("rgb_depthscale", _c.c_float),
                ("image", MatView), ("kdtree", _vp), ("kdtree_bytes", _c.c_uint64), ("min_xy", _c.c_int32 * 2), ("max_xy", _c.c_int32 * 2)]


class MapPointStream(ctypes.Structure):  # uco_mappoint_stream
    _fields_ = [("id", _c.c_uint32), ("pos3d", _c.c_float * 3), ("desc", MatView), ("n_frames", _c.c_uint32), ("frames", _vp),
                ("normal", _c.c_float * 3), ("n_times_seen", _c.c_uint16), ("n_times_visible", _c.c_uint16), ("flags", _c.c_uint8),
                ("max_distance", _c.c_float), ("min_distance", _c.c_float), ("kf_since_addition", _c.c_uint64), ("last_fidx_seen", _c.c_uint32)]


class MapPointContainer(ctypes.Structure):  # uco_mappoint_container
    _fields_ = [("n_slots", _c.c_uint32), ("n_used", _c.c_uint32), ("n_valid", _c.c_uint32), ("n_free", _c.c_uint32), ("free_slots", _vp)]


class KfdbStream(ctypes.Structure):  # uco_kfdb_stream
    _fields_ = [("type", _c.c_int32), ("voc_off", _c.c_size_t), ("voc_len", _c.c_size_t), ("n_words", _c.c_uint32), ("words_off", _c.c_size_t),
                ("n_word_frames", _c.c_uint64), ("n_frames", _c.c_uint32), ("frames", _vp)]


class MarkerStream(ctypes.Structure):  # uco_marker_stream
    _fields_ = [("key", _c.c_uint32), ("id", _c.c_uint32), ("pose_g2m", _c.c_float * 16), ("size", _c.c_float), ("n_frames", _c.c_uint32), ("frames", _vp),
                ("dict_len", _c.c_uint32), ("dict", _vp)]


class CovisStream(ctypes.Structure):  # uco_covis_stream
    _fields_ = [("n_nodes", _c.c_uint32), ("nodes", _vp), ("n_adj", _c.c_uint32), ("adj_off", _c.c_size_t), ("n_neighbours", _c.c_uint64),
                ("n_weights", _c.c_uint32), ("weights", _vp)]


class MapSections(ctypes.Structure):  # uco_map_sections
    _fields_ = [("kfdb_off", _c.c_size_t), ("kfdb_len", _c.c_size_t), ("kfdb", KfdbStream), ("points_off", _c.c_size_t), ("points_len", _c.c_size_t),
                ("points", MapPointContainer), ("markers_off", _c.c_size_t), ("markers_len", _c.c_size_t), ("n_markers", _c.c_uint32),
                ("frames_off", _c.c_size_t), ("frames_len", _c.c_size_t), ("frames", MapPointContainer), ("covis_off", _c.c_size_t), ("covis_len", _c.c_size_t),
                ("covis", CovisStream), ("total_len", _c.c_size_t)]


def map_stream_walk(buf, has_file_magic=False):
    """uco_map_sections of a Map::toStream byte range (has_file_magic: a map file written by Map::saveToFile)"""
    o = MapSections()
    rc = load().uco_b200_map_stream_walk(buf.ctypes.data, len(buf), int(has_file_magic), ctypes.addressof(o))
    if rc != 0:
        raise UcoError("map_stream_walk: malformed stream (%d)" % rc)
    return o


def mappoint_container_walk(buf, frames=False):
    """(header, slot offsets, slot valid flags, bytes consumed) of the map-point section (frames=True: the keyframe section, FrameSet) at the start of
    `buf` (uint8 array)"""
    lib = load()
    c, used = MapPointContainer(), ctypes.c_size_t()
    walk = lib.uco_b200_frame_container_walk if frames else lib.uco_b200_mappoint_container_walk
    rc = walk(buf.ctypes.data, len(buf), ctypes.addressof(c), None, None, 0, ctypes.addressof(used))
    if rc != 0:
        raise UcoError("mappoint_container_walk: malformed section (%d)" % rc)
    off, valid = np.zeros(c.n_slots, np.uint64), np.zeros(c.n_slots, np.uint8)
    rc = walk(buf.ctypes.data, len(buf), ctypes.addressof(c), _p(off), _p(valid), c.n_slots, ctypes.addressof(used))
    if rc != 0:
        raise UcoError("mappoint_container_walk failed (%d)" % rc)
    return c, off, valid, used.value


def mappoints_from_container(buf):
    """the valid map points of the section as the arrays of uco_mappoints (+ flags): dict(ids, pos, normal, min_dist, max_dist, desc, flags)"""
    lib = load()
    n = ctypes.c_uint32()
    rc = lib.uco_b200_mappoints_from_container(buf.ctypes.data, len(buf), 0, None, None, None, None, None, None, None, ctypes.addressof(n), None)
    if rc not in (0, -4) and n.value == 0:
        raise UcoError("mappoints_from_container: malformed section (%d)" % rc)
    k = n.value
    out = dict(ids=np.zeros(k, np.uint32), pos=np.zeros((k, 3), np.float32), normal=np.zeros((k, 3), np.float32), min_dist=np.zeros(k, np.float32),
               max_dist=np.zeros(k, np.float32), desc=np.zeros((k, 32), np.uint8), flags=np.zeros(k, np.uint8))
    rc = lib.uco_b200_mappoints_from_container(buf.ctypes.data, len(buf), k, _p(out["ids"]), _p(out["pos"]), _p(out["normal"]), _p(out["min_dist"]),
                                               _p(out["max_dist"]), _p(out["desc"]), _p(out["flags"]), ctypes.addressof(n), None)
    if rc != 0:
        raise UcoError("mappoints_from_container failed (%d)" % rc)
    return out


class FrameDev(ctypes.Structure):  # uco_frame_dev
    _fields_ = [("idx", _c.c_uint32), ("fseq_idx", _c.c_uint32), ("n_kp", _c.c_int32), ("kps", _vp), ("desc", _vp), ("ids", _vp), ("flags", _vp),
                ("depth", _vp), ("n_nodes", _c.c_int32), ("nodes", _vp), ("n_leaf", _c.c_int32), ("leaf_idx", _vp), ("bbox", _vp),
                ("n_scale_factors", _c.c_int32), ("scale_factors", _vp), ("pose_f2g", _vp), ("bbox_host", _c.c_double * 4),
                ("pose_host", _c.c_float * 16), ("K", _c.c_float * 4), ("min_xy", _c.c_int32 * 2), ("max_xy", _c.c_int32 * 2)]


def frame_stream_parse(buf):
    """view of the Frame::toStream bytes at the start of `buf` (uint8 array; it must outlive the view) -> (FrameStream, consumed)"""
    buf = np.ascontiguousarray(buf, np.uint8)
    v = FrameStream()
    n = _c.c_size_t(0)
    if load().uco_b200_frame_stream_parse(_p(buf), len(buf), ctypes.addressof(v), ctypes.addressof(n)) != 0:
        raise UcoError("frame_stream_parse: not a Frame stream")
    v._keep = buf
    return v, int(n.value)


def frame_stream_write(view):
    n = _c.c_size_t(0)
    lib = load()
    if lib.uco_b200_frame_stream_write(ctypes.addressof(view), None, 0, ctypes.addressof(n)) != 0:
        raise UcoError("frame_stream_write failed")
    out = np.zeros(n.value, np.uint8)
    if lib.uco_b200_frame_stream_write(ctypes.addressof(view), _p(out), len(out), ctypes.addressof(n)) != 0:
        raise UcoError("frame_stream_write failed")
    return out


def view_array(ptr, count, dtype):
    """numpy view of `count` elements of `dtype` at address `ptr` (a field of a FrameStream)"""
    if not ptr or count == 0:
        return np.zeros(0, dtype)
    dt = np.dtype(dtype)
    return np.frombuffer((ctypes.c_uint8 * (count * dt.itemsize)).from_address(ptr), dt, count)


class NewPointsParams(ctypes.Structure):  # uco_new_points_params
    _fields_ = [("match", MatchParams), ("K_kf", _c.c_float * 4), ("g2f_kf", _c.c_float * 16), ("n_levels_kf", _c.c_int32), ("scale_factors_kf", _vp),
                ("n_levels_nb", _c.c_int32), ("scale_factors_nb", _vp), ("max_chi2", _c.c_float), ("scale_ratio_factor", _c.c_float),
                ("max_points", _c.c_int32)]


class UcoError(RuntimeError):
    pass


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class Context:
    """One uco_b200_ctx: a CUDA stream plus workspaces. Not thread safe (one per calling thread)."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = self.lib.uco_b200_create(device, 0)
        if not self.h:
            raise UcoError("uco_b200_create failed: no usable CUDA device (there is no CPU fallback)")

    def close(self):
        if self.h:
            self.lib.uco_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise UcoError("uco_b200 error %d: %s" % (rc, self.lib.uco_b200_last_error(self.h).decode()))

    @property
    def stream(self):
        return self.lib.uco_b200_stream(self.h)

    def sync(self):
        self._chk(self.lib.uco_b200_sync(self.h))

    def launch_count(self):
        return int(self.lib.uco_b200_launch_count(self.h))

    # -- K15 -----------------------------------------------------------------------------------------------------
    def undistort_points(self, pts, K, dist):
        """(n,2) f32 pixel coordinates -> undistorted pixel coordinates (ucoslam::undistortPoints)"""
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2); K = np.ascontiguousarray(K, np.float32)
        dist = np.ascontiguousarray(dist, np.float32).reshape(-1)
        out = np.empty_like(pts)
        self._chk(self.lib.uco_b200_undistort_points(self.h, _p(pts), 8, len(pts), _p(K), _p(dist), len(dist), _p(out), 8))
        return out

    def undistort_keypoints(self, kps, K, dist):
        """Frame::und_kpts from the extracted keypoints: a copy of kps with pt undistorted (the other fields untouched)"""
        und = np.ascontiguousarray(kps).copy()
        K = np.ascontiguousarray(K, np.float32); dist = np.ascontiguousarray(dist, np.float32).reshape(-1)
        self._chk(self.lib.uco_b200_undistort_points(self.h, _p(und), und.dtype.itemsize, len(und), _p(K), _p(dist), len(dist), _p(und),
                                                     und.dtype.itemsize))
        return und

    # -- K14 -----------------------------------------------------------------------------------------------------
    def pnp_ransac(self, sc, max_iters, samples=None, seed=0):
        """sc: dict(p3d (n,3) f32, p2d (n,2) f32, normals (n,3) f32, cam (fx fy cx cy)).  Returns dict(ok, pose44, inliers, counts, best_iter)"""
        p3 = np.ascontiguousarray(sc["p3d"], np.float32); p2 = np.ascontiguousarray(sc["p2d"], np.float32)
        nr = np.ascontiguousarray(sc["normals"], np.float32); cam = np.ascontiguousarray(sc["cam"], np.float32)
        n = len(p3)
        smp = None if samples is None else np.ascontiguousarray(samples, np.int32)
        pose = np.zeros((4, 4), np.float32); inl = np.zeros(max(n, 1), np.int32); counts = np.zeros(max(max_iters, 1), np.int32)
        ni, bi = ctypes.c_int(), ctypes.c_int()
        self._chk(self.lib.uco_b200_pnp_ransac(self.h, _p(p3), _p(p2), _p(nr), n, _p(cam), int(max_iters), _p(smp), int(seed), _p(pose),
                                               _p(inl), ctypes.addressof(ni), _p(counts), ctypes.addressof(bi)))
        return dict(ok=ni.value > 0, pose44=pose, inliers=inl[:ni.value].copy(), counts=counts[:max_iters].copy(), best_iter=int(bi.value))

    # -- K13 -----------------------------------------------------------------------------------------------------
    def triangulate(self, sc, max_chi2=5.998, scale_ratio_factor=0.0, g2f_train=None):
        """sc: dict(kps_train, kps_query, matches (MATCH_DTYPE), K_train, K_query (fx fy cx cy), RT (4,4), sf_train, sf_query)
        -> (xyz f32 (n,3) with NaN rows where rejected, n_good)"""
        k1, k2 = np.ascontiguousarray(sc["kps_train"]), np.ascontiguousarray(sc["kps_query"])
        m = np.ascontiguousarray(sc["matches"])
        s1, s2 = np.ascontiguousarray(sc["sf_train"], np.float32), np.ascontiguousarray(sc["sf_query"], np.float32)
        prm = TriangulateParams()
        prm.K_train[:] = [float(x) for x in sc["K_train"]]; prm.K_query[:] = [float(x) for x in sc["K_query"]]
        prm.RT[:] = [float(x) for x in np.asarray(sc["RT"], np.float32).reshape(-1)]
        prm.n_levels_train, prm.scale_factors_train = len(s1), _p(s1)
        prm.n_levels_query, prm.scale_factors_query = len(s2), _p(s2)
        prm.max_chi2 = max_chi2
        prm.scale_ratio_factor = scale_ratio_factor
        prm.to_global = int(g2f_train is not None)
        prm.g2f_train[:] = [float(x) for x in np.asarray(np.eye(4) if g2f_train is None else g2f_train, np.float32).reshape(-1)]
        xyz = np.zeros((len(m), 3), np.float32)
        n = ctypes.c_int()
        self._chk(self.lib.uco_b200_triangulate(self.h, _p(k1), len(k1), _p(k2), len(k2), _p(m), len(m), ctypes.addressof(prm), _p(xyz),
                                                ctypes.addressof(n)))
        return xyz, int(n.value)

    # -- K12 -----------------------------------------------------------------------------------------------------
    def stereo_depth(self, sc, max_desc_dist=50.0):
        """sc: dict(img_l, img_r, kps_l, desc_l, kps_r, desc_r, bl, fx) -> (depth f32 (n_l,), match i32 (n_l,), n_with_depth)"""
        il, ir = np.asarray(sc["img_l"]), np.asarray(sc["img_r"])
        h, w = il.shape
        kl, kr = np.ascontiguousarray(sc["kps_l"]), np.ascontiguousarray(sc["kps_r"])
        dl, dr = np.asarray(sc["desc_l"]), np.asarray(sc["desc_r"])
        depth = np.zeros(len(kl), np.float32); match = np.full(len(kl), -1, np.int32)
        n = ctypes.c_int()
        self._chk(self.lib.uco_b200_stereo_depth(self.h, _p(il), il.strides[0], _p(ir), ir.strides[0], w, h, _p(kl), _p(dl),
                                                 dl.strides[0] if len(dl) else 32, len(kl), _p(kr), _p(dr),
                                                 dr.strides[0] if len(dr) else 32, len(kr), float(max_desc_dist), float(sc["bl"]),
                                                 float(sc["fx"]), _p(depth), _p(match), ctypes.addressof(n)))
        return depth, match, int(n.value)

    # -- K7 ------------------------------------------------------------------------------------------------------
    def hamming_knn(self, q, t, k, order=UCO_KNN_HEAP):
        """q: (nq,32) uint8 host rows (may be strided), t: (nt,32). Returns (idx, dist) int32 (nq,k)."""
        q = np.asarray(q)
        t = np.asarray(t)
        assert q.dtype == np.uint8 and t.dtype == np.uint8 and q.ndim == 2 and t.ndim == 2
        assert q.shape[1] == 32 and t.shape[1] == 32
        assert q.strides[1] == 1 and t.strides[1] == 1
        nq, nt = q.shape[0], t.shape[0]
        idx = np.empty((nq, k), np.int32)
        dist = np.empty((nq, k), np.int32)
        qs = q.strides[0] if nq > 0 else 32
        ts = t.strides[0] if nt > 0 else 32
        self._chk(self.lib.uco_b200_hamming_knn(self.h, _p(q), nq, qs, _p(t), nt, ts, k, order, _p(idx), _p(dist)))
        return idx, dist

    # -- K10-K13 -------------------------------------------------------------------------------------------------
    def hamming_knn_batch(self, qs, ts, k, order=UCO_KNN_HEAP):
        """host-buffer batch: qs[i] against ts[i] for every pair in one launch; returns lists of (idx, dist) arrays"""
        qs = [np.ascontiguousarray(q, np.uint8).reshape(-1, 32) for q in qs]
        ts = [t if any(t is q for q in qs) else np.ascontiguousarray(t, np.uint8).reshape(-1, 32) for t in ts]
        n = len(qs)
        idx = [np.empty((len(q), k), np.int32) for q in qs]
        dist = [np.empty((len(q), k), np.int32) for q in qs]
        VP = ctypes.c_void_p * n
        nq = np.array([len(q) for q in qs], np.int32)
        nt = np.array([len(t) for t in ts], np.int32)
        self._chk(self.lib.uco_b200_hamming_knn_batch(self.h, n, VP(*[q.ctypes.data for q in qs]), _p(nq), 32,
                                                      VP(*[t.ctypes.data for t in ts]), _p(nt), 32, k, order,
                                                      VP(*[a.ctypes.data for a in idx]), VP(*[a.ctypes.data for a in dist])))
        return idx, dist

    # -- K9 ------------------------------------------------------------------------------------------------------
    def match_projected(self, sc, min_desc_dist, max_reproj_dist, tree=None):
        """Map::matchFrameToMapPoints on a scene dict (synth.synth_projection_scene keys); tree = (nodes, leaf_idx, bbox) of the frame's
        kd-tree (default: built here).  Returns (matches[MATCH_DTYPE], visible u8[m])."""
        A = lambda k, dt: np.ascontiguousarray(sc[k], dt)
        ids, pos, nrm = A("mp_id", np.uint32), A("mp_pos", np.float32), A("mp_normal", np.float32)
        dmin, dmax, mdesc = A("mp_min_dist", np.float32), A("mp_max_dist", np.float32), A("mp_desc", np.uint8)
        kdesc, sf, pose = A("kp_desc", np.uint8), A("scale_factors", np.float32), A("pose44", np.float32)
        kxy, koct = A("kp_xy", np.float32).reshape(-1, 2), A("kp_octave", np.int32)
        kps = np.zeros(len(kxy), KP_DTYPE)
        kps["x"], kps["y"], kps["octave"] = kxy[:, 0], kxy[:, 1], koct
        nodes, leaf, bbox = tree if tree is not None else kdtree_build(kxy)
        nodes, leaf = np.ascontiguousarray(nodes), np.ascontiguousarray(leaf, np.int32)
        mp = MapPoints(len(ids), _p(ids), _p(pos), _p(nrm), _p(dmin), _p(dmax), _p(mdesc))
        fr = FrameView(len(kps), _p(kps), _p(kdesc), 32, len(nodes), _p(nodes), _p(leaf), (ctypes.c_double * 4)(*bbox), len(sf), _p(sf),
                       sc["fx"], sc["fy"], sc["cx"], sc["cy"], (ctypes.c_float * 2)(*sc["min_xy"]), (ctypes.c_float * 2)(*sc["max_xy"]))
        m = len(ids)
        out, vis, n = np.zeros(max(m, 1), MATCH_DTYPE), np.zeros(max(m, 1), np.uint8), ctypes.c_int(0)
        self._chk(self.lib.uco_b200_match_projected(self.h, ctypes.addressof(mp), ctypes.addressof(fr), _p(pose), min_desc_dist,
                                                    max_reproj_dist, _p(out), ctypes.byref(n), _p(vis)))
        return out[:n.value].copy(), vis[:m].copy()

    # -- K16 / K17 ----------------------------------------------------------------------------------------------
    def kdtree_build_dev(self, xy):
        """the frame's kd-tree built on the device: (nodes[KDNODE_DTYPE], leaf_idx i32, bbox f64[4]) like kdtree_build"""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        n = len(xy)
        cap = 2 * n + 4
        nodes, leaf, bbox, k = np.zeros(cap, KDNODE_DTYPE), np.zeros(max(n, 1), np.int32), np.zeros(4), ctypes.c_int(0)
        self._chk(self.lib.uco_b200_kdtree_build_dev(self.h, _p(xy), 8, n, _p(nodes), cap, _p(leaf), _p(bbox), ctypes.byref(k)))
        return nodes[:k.value].copy(), leaf[:n].copy(), bbox

    def track_projected(self, sc, dist_thr, proj_dist_thr, tree=None):
        """System's search by projection from the previous frame on a scene dict (synth.synth_track_scene keys) -> matches"""
        A = lambda k, dt: np.ascontiguousarray(sc[k], dt)
        ids, pos = A("mp_id", np.uint32), A("mp_pos", np.float32)
        kdesc, sf, pose = A("kp_desc", np.uint8), A("scale_factors", np.float32), A("pose44", np.float32)
        kxy, koct = A("kp_xy", np.float32).reshape(-1, 2), A("kp_octave", np.int32)
        kps = np.zeros(len(kxy), KP_DTYPE)
        kps["x"], kps["y"], kps["octave"] = kxy[:, 0], kxy[:, 1], koct
        pkps = np.zeros(len(sc["prev_octave"]), KP_DTYPE)
        pkps["octave"] = sc["prev_octave"]
        pdesc, prow = A("prev_desc", np.uint8), A("prev_mp_row", np.int32)
        nodes, leaf, bbox = tree if tree is not None else kdtree_build(kxy)
        nodes, leaf = np.ascontiguousarray(nodes), np.ascontiguousarray(leaf, np.int32)
        mp = MapPoints(len(ids), _p(ids), _p(pos), None, None, None, None)
        fr = FrameView(len(kps), _p(kps), _p(kdesc), 32, len(nodes), _p(nodes), _p(leaf), (ctypes.c_double * 4)(*bbox), len(sf), _p(sf),
                       sc["fx"], sc["fy"], sc["cx"], sc["cy"], (ctypes.c_float * 2)(*sc["min_xy"]), (ctypes.c_float * 2)(*sc["max_xy"]))
        out, n = np.zeros(max(len(pkps), 1), MATCH_DTYPE), ctypes.c_int(0)
        self._chk(self.lib.uco_b200_track_projected(self.h, len(pkps), _p(pkps), _p(pdesc), _p(prow), ctypes.addressof(mp),
                                                    ctypes.addressof(fr), _p(pose), dist_thr, proj_dist_thr, _p(out), ctypes.byref(n)))
        return out[:n.value].copy()

    @staticmethod
    def track_pack(scs):
        """flatten scene dicts (synth.synth_track_scene) into the fixed-stride arrays of uco_track_batch"""
        F = len(scs)
        kc = max(len(s["kp_octave"]) for s in scs)
        pc = max(len(s["prev_octave"]) for s in scs)
        mc = max(len(s["mp_id"]) for s in scs)
        a = dict(kps=np.zeros((F, kc), KP_DTYPE), desc=np.zeros((F, kc, 32), np.uint8), n_kp=np.zeros(F, np.int32),
                 prev_kps=np.zeros((F, pc), KP_DTYPE), prev_desc=np.zeros((F, pc, 32), np.uint8), prev_n_kp=np.zeros(F, np.int32),
                 prev_mp_row=np.full((F, pc), -1, np.int32), map_n=np.zeros(F, np.int32), mp_id=np.zeros((F, mc), np.uint32),
                 mp_pos=np.zeros((F, mc, 3), np.float32), mp_normal=np.zeros((F, mc, 3), np.float32), mp_min_dist=np.zeros((F, mc), np.float32),
                 mp_max_dist=np.zeros((F, mc), np.float32), mp_desc=np.zeros((F, mc, 32), np.uint8), mp_stable=np.ones((F, mc), np.uint8),
                 mp_local=np.ones((F, mc), np.uint8), pose_prior=np.zeros((F, 16), np.float32))
        for f, s in enumerate(scs):
            nk, npv, nm = len(s["kp_octave"]), len(s["prev_octave"]), len(s["mp_id"])
            xy = np.asarray(s["kp_xy"], np.float32).reshape(-1, 2)
            a["kps"][f, :nk]["x"], a["kps"][f, :nk]["y"], a["kps"][f, :nk]["octave"] = xy[:, 0], xy[:, 1], s["kp_octave"]
            a["desc"][f, :nk] = s["kp_desc"]; a["n_kp"][f] = nk
            a["prev_kps"][f, :npv]["octave"] = s["prev_octave"]; a["prev_desc"][f, :npv] = s["prev_desc"]; a["prev_n_kp"][f] = npv
            a["prev_mp_row"][f, :npv] = s["prev_mp_row"]; a["map_n"][f] = nm; a["mp_id"][f, :nm] = s["mp_id"]
            a["mp_pos"][f, :nm] = s["mp_pos"]; a["mp_normal"][f, :nm] = s["mp_normal"]; a["mp_min_dist"][f, :nm] = s["mp_min_dist"]
            a["mp_max_dist"][f, :nm] = s["mp_max_dist"]; a["mp_desc"][f, :nm] = s["mp_desc"]
            if "mp_stable" in s: a["mp_stable"][f, :nm] = s["mp_stable"]
            if "mp_local" in s: a["mp_local"][f, :nm] = s["mp_local"]
            a["pose_prior"][f] = np.asarray(s["pose44"], np.float32).reshape(16)
        return a, (F, kc, pc, mc)

    def track_batch(self, scs, prm=None, packed=None):
        """the tracker's main branch for a batch of independent frames (host buffers) -> list of dicts(matches, pose44, n_good, status, n_tbp, visible)"""
        a, (F, kc, pc, mc) = packed or self.track_pack(scs)
        prm = prm or TrackParams(scs[0])
        tb = TrackBatch(F, kc, pc, mc, 0, *[_p(a[k]) if k in a else None for k in (
            "kps", "desc", "n_kp", "depth", "prev_kps", "prev_desc", "prev_n_kp", "prev_mp_row", "map_n", "mp_id", "mp_pos", "mp_normal",
            "mp_min_dist", "mp_max_dist", "mp_desc", "mp_stable", "mp_local", "pose_prior")])
        o = dict(matches=np.zeros((F, kc), MATCH_DTYPE), n_matches=np.zeros(F, np.int32), pose=np.zeros((F, 16), np.float32),
                 n_good=np.zeros(F, np.int32), status=np.zeros(F, np.int32), n_tbp=np.zeros(F, np.int32), visible=np.zeros((F, mc), np.uint8))
        to = TrackOut(*[_p(o[k]) for k in ("matches", "n_matches", "pose", "n_good", "status", "n_tbp", "visible")])
        self._chk(self.lib.uco_b200_track_batch(self.h, ctypes.addressof(tb), ctypes.addressof(prm), ctypes.addressof(to)))
        return [dict(matches=o["matches"][f, :o["n_matches"][f]].copy(), pose44=o["pose"][f].copy(), n_good=int(o["n_good"][f]),
                     status=int(o["status"][f]), n_tbp=int(o["n_tbp"][f]), visible=o["visible"][f, :a["map_n"][f]].copy()) for f in range(F)]

    # -- multi-GPU ----------------------------------------------------------------------------------------------
    def comm_create(self, rank=0, world=1, broadcast=None):
        """communicator of this context; `broadcast(bytes_or_None) -> bytes` distributes rank 0's 128-byte id (see shard.make_comm)"""
        idb = None
        if world > 1:
            buf = (ctypes.c_uint8 * 128)()
            if rank == 0:
                if self.lib.uco_b200_comm_unique_id(buf) != 0:
                    raise UcoError("uco_b200_comm_unique_id failed (NCCL not loadable)")
            data = broadcast(bytes(buf) if rank == 0 else None)
            idb = (ctypes.c_uint8 * 128).from_buffer_copy(data)
        out = ctypes.c_void_p()
        self._chk(self.lib.uco_b200_comm_create(self.h, idb, rank, world, ctypes.byref(out)))
        return out

    def comm_destroy(self, comm):
        self.lib.uco_b200_comm_destroy(comm)

    def hamming_knn_sharded_dev(self, comm, q_dev, nq, t_shard_dev, nt_shard, row_base, k, idx_dev, dist_dev):
        self._chk(self.lib.uco_b200_hamming_knn_sharded_dev(self.h, comm, q_dev, nq, t_shard_dev, nt_shard, row_base, k, idx_dev, dist_dev))

    def knn_merge_dev(self, n_lists, nq, k, idx_lists_dev, dist_lists_dev, idx_dev, dist_dev):
        self._chk(self.lib.uco_b200_knn_merge_dev(self.h, n_lists, nq, k, idx_lists_dev, dist_lists_dev, idx_dev, dist_dev))

    def block_solve(self, nb, blk_ij, blocks, rhs, force_k=-1):
        """S x = b, S symmetric positive definite block sparse (upper block triangle given); returns (x, info8)"""
        blk_ij = np.ascontiguousarray(blk_ij, np.int32).reshape(-1, 2)
        blocks = np.ascontiguousarray(blocks, np.float64).reshape(-1, 36)
        rhs = np.ascontiguousarray(rhs, np.float64).reshape(-1)
        x = np.zeros(6 * nb)
        info = np.zeros(8, np.int32)
        self._chk(self.lib.uco_b200_block_solve(self.h, nb, len(blk_ij), _p(blk_ij), _p(blocks), _p(rhs), force_k, _p(x), _p(info)))
        return x, info

    def block_solve_ms(self):
        ms = np.zeros(1, np.float32)
        self._chk(self.lib.uco_b200_block_solve_profile(self.h, _p(ms)))
        return float(ms[0])

    def ba_solve_sharded(self, pb, n_iters, comm=None, stop=None):
        """uco_b200_ba_solve_sharded: any problem size; with a communicator the landmarks are sharded over its ranks"""
        cp, cr, keep, out = self.ba_pack(pb, n_iters)
        self._chk(self.lib.uco_b200_ba_solve_sharded(self.h, comm, ctypes.addressof(cp), _p(stop), ctypes.addressof(cr)))
        out["iters"] = np.array(list(cr.iters), np.int32)
        out["device_ms"] = float(cr.device_ms)
        return out

    def ba_set_mode(self, mode=0, cluster_size=0):
        self._chk(self.lib.uco_b200_ba_set_mode(self.h, mode, cluster_size))

    def ba_set_host_threads(self, n):
        self._chk(self.lib.uco_b200_ba_set_host_threads(self.h, n))

    @staticmethod
    def ba_pack(pb, n_iters):
        """(uco_ba_problem, uco_ba_result, keep-alive input arrays, output dict) for a problem dict (see ba_solve)."""
        A = lambda k, dt: np.ascontiguousarray(pb[k], dtype=dt)
        a = dict(poses44=A("poses44", np.float32), fixed=A("fixed", np.uint8), points3=A("points3", np.float32),
                 obs_pose=A("obs_pose", np.int32), obs_point=A("obs_point", np.int32), obs_uv=A("obs_uv", np.float32),
                 obs_ur=A("obs_ur", np.float32), obs_stereo=A("obs_stereo", np.uint8), obs_inv_sigma2=A("obs_inv_sigma2", np.float32))
        P, N, M = len(a["fixed"]), len(a["points3"]), len(a["obs_pose"])
        out = dict(pose7=np.zeros((P, 7)), pose44=np.zeros((P, 16), np.float32), point3=np.zeros((N, 3)), chi2=np.zeros(M),
                   level=np.zeros(M, np.uint8), bad=np.zeros(M, np.uint8), trace=np.zeros((64, 2)), profile=np.zeros(16))
        cp = BaProblem(P, N, M, _p(a["poses44"]), _p(a["fixed"]), _p(a["points3"]), _p(a["obs_pose"]), _p(a["obs_point"]),
                       _p(a["obs_uv"]), _p(a["obs_ur"]), _p(a["obs_stereo"]), _p(a["obs_inv_sigma2"]), pb["fx"], pb["fy"],
                       pb["cx"], pb["cy"], pb["bf"], int(n_iters))
        cr = BaResult(_p(out["pose7"]), _p(out["pose44"]), _p(out["point3"]), _p(out["chi2"]), _p(out["level"]), _p(out["bad"]),
                      _p(out["trace"]))
        cr.profile = _p(out["profile"])
        if len(pb.get("marker_size", ())):   # ArUco markers
            a.update(marker_pose44=A("marker_pose44", np.float32), marker_size=A("marker_size", np.float32), mobs_marker=A("mobs_marker", np.int32),
                     mobs_pose=A("mobs_pose", np.int32), mobs_corners=A("mobs_corners", np.float32), mobs_weight=A("mobs_weight", np.float32))
            nm, nmo = len(a["marker_size"]), len(a["mobs_marker"])
            cp.n_markers, cp.marker_pose44, cp.marker_size = nm, _p(a["marker_pose44"]), _p(a["marker_size"])
            cp.n_marker_obs, cp.mobs_marker, cp.mobs_pose = nmo, _p(a["mobs_marker"]), _p(a["mobs_pose"])
            cp.mobs_corners, cp.mobs_weight = _p(a["mobs_corners"]), _p(a["mobs_weight"])
            out.update(marker_pose44=np.zeros((nm, 16), np.float32), marker_pose7=np.zeros((nm, 7)), mobs_chi2=np.zeros(nmo))
            cr.marker_poses44, cr.marker_pose7, cr.mobs_chi2 = _p(out["marker_pose44"]), _p(out["marker_pose7"]), _p(out["mobs_chi2"])
        if len(pb.get("plane_other", ())):   # InPlaneMarkers: reference marker (index, or -1 + its fixed pose) tied to the other markers
            a["plane_other"] = A("plane_other", np.int32)
            cp.n_plane, cp.plane_ref, cp.plane_other, cp.plane_weight = len(a["plane_other"]), int(pb["plane_ref"]), _p(a["plane_other"]), float(pb["plane_weight"])
            if int(pb["plane_ref"]) < 0:
                a["plane_ref_pose44"] = A("plane_ref_pose44", np.float32)
                cp.plane_ref_pose44 = _p(a["plane_ref_pose44"])
        if pb.get("pose_cam") is not None:   # one camera per keyframe (fx fy cx cy bf)
            a["pose_cam"] = A("pose_cam", np.float32)
            assert a["pose_cam"].shape == (P, 5)
            cp.pose_cam = _p(a["pose_cam"])
        return cp, cr, a, out

    def ba_solve(self, pb, n_iters, stop=None):
        """pb: dict with the uco_ba_problem arrays (poses44 f32 (P,16), fixed u8, points3 f32 (N,3), obs_pose/obs_point i32,
        obs_uv f32 (M,2), obs_ur f32, obs_stereo u8, obs_inv_sigma2 f32, fx fy cx cy bf).  Returns a dict like the oracle's."""
        cp, cr, keep, out = self.ba_pack(pb, n_iters)
        self._chk(self.lib.uco_b200_ba_solve(self.h, ctypes.addressof(cp), _p(stop), ctypes.addressof(cr)))
        out["iters"] = np.array(list(cr.iters), np.int32)
        out["device_ms"] = float(cr.device_ms)
        return out

    def ba_solve_batch(self, pbs, n_iters, stop=None, packed=None):
        """several independent problems in one call; `packed` = result of a previous ba_pack_batch to skip the packing"""
        packed = packed or self.ba_pack_batch(pbs, n_iters)
        cps, crs, keep, outs = packed
        self._chk(self.lib.uco_b200_ba_solve_batch(self.h, len(outs), ctypes.addressof(cps), _p(stop), ctypes.addressof(crs)))
        for i, o in enumerate(outs):
            o["iters"] = np.array(list(crs[i].iters), np.int32)
            o["device_ms"] = float(crs[i].device_ms)
        return outs

    def ba_pack_batch(self, pbs, n_iters):
        packs = [self.ba_pack(pb, n_iters) for pb in pbs]
        cps = (BaProblem * len(packs))(*[p[0] for p in packs])
        crs = (BaResult * len(packs))(*[p[1] for p in packs])
        return cps, crs, [p[2] for p in packs], [p[3] for p in packs]

    # -- K8 ------------------------------------------------------------------------------------------------------
    def frame_match(self, q_desc, q_kps, t_desc, t_kps, prm, q_map=None, t_map=None):
        """FrameMatcher_Flann::setParams(train) + matchEpipolar(query): (n,) MATCH_DTYPE records in the reference's order."""
        q_desc = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
        t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
        q_kps, t_kps = np.ascontiguousarray(q_kps, KP_DTYPE), np.ascontiguousarray(t_kps, KP_DTYPE)
        qm = None if q_map is None else np.ascontiguousarray(q_map, np.int32)
        tm = None if t_map is None else np.ascontiguousarray(t_map, np.int32)
        out = np.zeros(max(len(q_desc), 1), MATCH_DTYPE)
        n = _c.c_int(0)
        self._chk(self.lib.uco_b200_frame_match(self.h, _p(q_desc), len(q_desc), 32, _p(q_kps), len(q_kps), _p(qm), _p(t_desc),
                                                len(t_desc), 32, _p(t_kps), len(t_kps), _p(tm), ctypes.addressof(prm), _p(out),
                                                len(out), ctypes.addressof(n)))
        return out[:n.value].copy()

    def frame_match_multi(self, t_desc, t_kps, q_descs, q_kpss, prm, f12=None, t_map=None, q_maps=None):
        """FrameMatcher::setParams(train = the keyframe) + matchEpipolar(query = each neighbour, F12_f): list of match arrays"""
        t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
        t_kps = np.ascontiguousarray(t_kps, KP_DTYPE)
        qd = [np.ascontiguousarray(d, np.uint8).reshape(-1, 32) for d in q_descs]
        qk = [np.ascontiguousarray(k, KP_DTYPE) for k in q_kpss]
        F = len(qd)
        VP = ctypes.c_void_p * max(F, 1)
        nq = np.array([len(d) for d in qd], np.int32)
        nk = np.array([len(k) for k in qk], np.int32)
        tm = None if t_map is None else np.ascontiguousarray(t_map, np.int32)
        qm = None if q_maps is None else [None if m is None else np.ascontiguousarray(m, np.int32) for m in q_maps]
        qmp = None if qm is None else VP(*[None if m is None else m.ctypes.data for m in qm])
        f12a = None if f12 is None else np.ascontiguousarray(f12, np.float32).reshape(F, 9)
        cap = max(1, int(nq.max()) if F else 1)
        outs = [np.zeros(cap, MATCH_DTYPE) for _ in range(F)]
        n_out = np.zeros(max(F, 1), np.int32)
        self._chk(self.lib.uco_b200_frame_match_multi(
            self.h, _p(t_desc), len(t_desc), 32, _p(t_kps), len(t_kps), _p(tm), F, ctypes.cast(VP(*[d.ctypes.data for d in qd]), ctypes.c_void_p),
            _p(nq), 32, ctypes.cast(VP(*[k.ctypes.data for k in qk]), ctypes.c_void_p), _p(nk),
            None if qmp is None else ctypes.cast(qmp, ctypes.c_void_p), _p(f12a), ctypes.addressof(prm),
            ctypes.cast(VP(*[o.ctypes.data for o in outs]), ctypes.c_void_p), cap, _p(n_out)))
        return [outs[f][:n_out[f]].copy() for f in range(F)]

    def frame_upload(self, view):
        """device-resident mirror of a parsed Frame stream -> (handle, FrameDev with DEVICE pointers); free with frame_free"""
        hnd = ctypes.c_void_p()
        self._chk(self.lib.uco_b200_frame_upload(self.h, ctypes.addressof(view), ctypes.addressof(hnd)))
        dev = FrameDev.from_address(self.lib.uco_b200_frame_dev(hnd))
        return hnd, dev

    def frame_download(self, hnd, n):
        kps, desc = np.zeros(n, KP_DTYPE), np.zeros((n, 32), np.uint8)
        ids, flags, depth = np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.zeros(n, np.float32)
        self._chk(self.lib.uco_b200_frame_download(self.h, hnd, _p(kps), _p(desc), _p(ids), _p(flags), _p(depth)))
        return kps, desc, ids, flags, depth

    def frame_free(self, hnd):
        self.lib.uco_b200_frame_free(hnd)

    def new_points(self, sc, max_points=-1, per_pair=False):
        """MapManager::createNewPoints on a scene dict (synth.synth_new_points_scene): returns dict(kpt, xyz, dist, obs_ptr, obs_frame, obs_kpt
        [, matches, xyz_pairs])"""
        t_desc = np.ascontiguousarray(sc["t_desc"], np.uint8).reshape(-1, 32)
        t_kps = np.ascontiguousarray(sc["t_kps"], KP_DTYPE)
        tm = np.ascontiguousarray(sc["t_map"], np.int32)
        qd = [np.ascontiguousarray(d, np.uint8).reshape(-1, 32) for d in sc["q_desc"]]
        qk = [np.ascontiguousarray(k, KP_DTYPE) for k in sc["q_kps"]]
        qm = [np.ascontiguousarray(m, np.int32) for m in sc["q_map"]]
        F = len(qd)
        VP = ctypes.c_void_p * max(F, 1)
        rows_t = np.ascontiguousarray(t_desc[tm])
        rows_q = [np.ascontiguousarray(d[m]) for d, m in zip(qd, qm)]
        nq = np.array([len(d) for d in rows_q], np.int32)
        nk = np.array([len(k) for k in qk], np.int32)
        sf1, sf2 = np.ascontiguousarray(sc["sf_kf"], np.float32), np.ascontiguousarray(sc["sf_nb"], np.float32)
        prm = NewPointsParams()
        prm.match = MatchParams(sc["min_desc_dist"], sc["ratio"], True, 2 ** 31 - 1, np.eye(3), sf2)
        prm.K_kf[:] = [float(v) for v in sc["K_kf"]]
        prm.g2f_kf[:] = [float(v) for v in np.asarray(sc["g2f_kf"], np.float32).reshape(-1)]
        prm.n_levels_kf, prm.scale_factors_kf = len(sf1), _p(sf1)
        prm.n_levels_nb, prm.scale_factors_nb = len(sf2), _p(sf2)
        prm.max_chi2, prm.scale_ratio_factor, prm.max_points = sc.get("max_chi2", 5.998), sc["scale_ratio_factor"], max_points
        f12 = np.ascontiguousarray(sc["f12"], np.float32).reshape(F, 9)
        rt = np.ascontiguousarray(sc["rt"], np.float32).reshape(F, 16)
        Knb = np.ascontiguousarray(sc["K_nb"], np.float32).reshape(F, 4)
        cap_p, cap_o = max(len(tm), 1), max(int(nq.sum()), 1)
        n_pts = ctypes.c_int32(0)
        kpt, xyz, dist = np.zeros(cap_p, np.int32), np.zeros((cap_p, 3), np.float32), np.zeros(cap_p, np.float32)
        optr, ofr, okp = np.zeros(cap_p + 1, np.int32), np.zeros(cap_o, np.int32), np.zeros(cap_o, np.int32)
        cap = max(1, int(nq.max()) if F else 1)
        ms = [np.zeros(cap, MATCH_DTYPE) for _ in range(F)] if per_pair else None
        xs = [np.zeros((cap, 3), np.float32) for _ in range(F)] if per_pair else None
        nm = np.zeros(max(F, 1), np.int32)
        self._chk(self.lib.uco_b200_new_points(
            self.h, _p(rows_t), len(rows_t), 32, _p(t_kps), len(t_kps), _p(tm), F, ctypes.cast(VP(*[d.ctypes.data for d in rows_q]), ctypes.c_void_p), _p(nq), 32,
            ctypes.cast(VP(*[k.ctypes.data for k in qk]), ctypes.c_void_p), _p(nk), ctypes.cast(VP(*[m.ctypes.data for m in qm]), ctypes.c_void_p),
            _p(f12), _p(rt), _p(Knb), ctypes.addressof(prm), ctypes.addressof(n_pts), _p(kpt), _p(xyz), _p(dist), _p(optr), _p(ofr), _p(okp), cap_p, cap_o,
            None if ms is None else ctypes.cast(VP(*[m.ctypes.data for m in ms]), ctypes.c_void_p), _p(nm),
            None if xs is None else ctypes.cast(VP(*[x.ctypes.data for x in xs]), ctypes.c_void_p)))
        n = int(n_pts.value)
        out = dict(kpt=kpt[:n].copy(), xyz=xyz[:n].copy(), dist=dist[:n].copy(), obs_ptr=optr[:n + 1].copy(), obs_frame=ofr[:optr[n]].copy(),
                   obs_kpt=okp[:optr[n]].copy())
        if per_pair:
            out["matches"] = [ms[f][:nm[f]].copy() for f in range(F)]
            out["xyz_pairs"] = [xs[f][:nm[f]].copy() for f in range(F)]
        return out

    def keyframes_batch(self, voc, groups, prm, max_features, f12=None, level=3):
        """per-keyframe work on the frames of this context's last extraction call: groups = [(kf_frame, [neighbour frames])];
        -> (list of (word, weight, node) per keyframe, list of lists of match arrays per keyframe)"""
        kf = np.array([g[0] for g in groups], np.int32)
        ptr = np.zeros(len(groups) + 1, np.int32)
        ptr[1:] = np.cumsum([len(g[1]) for g in groups])
        nb = np.array([f for g in groups for f in g[1]], np.int32) if ptr[-1] else np.zeros(1, np.int32)
        npairs, mf = int(ptr[-1]), max_features
        word, wgt, node = (np.zeros((len(groups), mf), np.uint32), np.zeros((len(groups), mf), np.float32), np.zeros((len(groups), mf), np.uint32))
        m, nm = np.zeros((max(npairs, 1), mf), MATCH_DTYPE), np.zeros(max(npairs, 1), np.int32)
        f12a = None if f12 is None else np.ascontiguousarray(f12, np.float32).reshape(npairs, 9)
        self._chk(self.lib.uco_b200_keyframes_batch(self.h, voc, level, len(groups), _p(kf), _p(ptr), _p(nb), _p(f12a), ctypes.addressof(prm),
                                                    _p(word), _p(wgt), _p(node), _p(m), _p(nm)))
        bows = [(word[j], wgt[j], node[j]) for j in range(len(groups))]
        matches = [[m[e, :nm[e]].copy() for e in range(ptr[j], ptr[j + 1])] for j in range(len(groups))]
        return bows, matches

    def frame_match_bow(self, q_desc, q_kps, q_bow, t_desc, t_kps, t_bow, prm, q_usable=None, t_usable=None):
        """FrameMatcher_BoW::matchEpipolar; q_bow / t_bow = (node_id, ptr, kp) as bow_index() gives"""
        q_desc = np.ascontiguousarray(q_desc, np.uint8).reshape(-1, 32)
        t_desc = np.ascontiguousarray(t_desc, np.uint8).reshape(-1, 32)
        q_kps, t_kps = np.ascontiguousarray(q_kps, KP_DTYPE), np.ascontiguousarray(t_kps, KP_DTYPE)
        keep = [np.ascontiguousarray(a) for a in (*q_bow, *t_bow)]
        qb = BowIndex(len(keep[0]), _p(keep[0]), _p(keep[1]), _p(keep[2]))
        tb = BowIndex(len(keep[3]), _p(keep[3]), _p(keep[4]), _p(keep[5]))
        qu = None if q_usable is None else np.ascontiguousarray(q_usable, np.uint8)
        tu = None if t_usable is None else np.ascontiguousarray(t_usable, np.uint8)
        out = np.zeros(max(len(keep[2]), 1), MATCH_DTYPE)
        n = _c.c_int(0)
        self._chk(self.lib.uco_b200_frame_match_bow(self.h, _p(q_desc), 32, _p(q_kps), len(q_kps), _p(qu), ctypes.addressof(qb), _p(t_desc), 32,
                                                    _p(t_kps), len(t_kps), _p(tu), ctypes.addressof(tb), ctypes.addressof(prm), _p(out), len(out),
                                                    ctypes.addressof(n)))
        return out[:n.value].copy()

    def frame_match_batch_dev(self, n_pairs, q_desc_dev, q_stride, q_kps_dev, q_kps_stride, nq_max, nq_dev, t_desc_dev, t_stride,
                              t_kps_dev, t_kps_stride, nt_max, nt_dev, prm, out_dev, n_out_dev):
        self._chk(self.lib.uco_b200_frame_match_batch_dev(self.h, n_pairs, _p(q_desc_dev), q_stride, _p(q_kps_dev), q_kps_stride,
                                                          nq_max, _p(nq_dev), _p(t_desc_dev), t_stride, _p(t_kps_dev), t_kps_stride,
                                                          nt_max, _p(nt_dev), ctypes.addressof(prm), _p(out_dev), _p(n_out_dev)))

    # -- K14 -----------------------------------------------------------------------------------------------------
    @staticmethod
    def pnp_pack(pbs):
        """(uco_pnp_problem[], uco_pnp_result[], keep-alive arrays, bad arrays) for problem dicts (see synth_pnp_problem)."""
        keep, cps, crs, bads = [], [], [], []
        for pb in pbs:
            A = lambda k, dt: np.ascontiguousarray(pb[k], dtype=dt)
            a = dict(pose44=A("pose44", np.float32), points3=A("points3", np.float32), obs_uv=A("obs_uv", np.float32),
                     obs_ur=A("obs_ur", np.float32), obs_stereo=A("obs_stereo", np.uint8), obs_inv_sigma2=A("obs_inv_sigma2", np.float32),
                     stable=A("stable", np.uint8), marker_pose44=A("marker_pose44", np.float32), marker_size=A("marker_size", np.float32),
                     marker_corners=A("marker_corners", np.float32))
            n, nm = len(a["points3"]), len(a["marker_size"])
            bad = np.zeros(max(n, 1), np.uint8)
            cps.append(PnpProblem(n, _p(a["pose44"]), _p(a["points3"]), _p(a["obs_uv"]), _p(a["obs_ur"]), _p(a["obs_stereo"]),
                                  _p(a["obs_inv_sigma2"]), _p(a["stable"]), pb["fx"], pb["fy"], pb["cx"], pb["cy"], pb["bf"], nm,
                                  _p(a["marker_pose44"]), _p(a["marker_size"]), _p(a["marker_corners"])))
            r = PnpResult()
            r.bad = _p(bad)
            crs.append(r)
            keep.append(a)
            bads.append(bad[:n])
        return (PnpProblem * len(cps))(*cps), (PnpResult * len(crs))(*crs), keep, bads

    def pose_only_batch(self, pbs, packed=None):
        """PnPSolver::solvePnp for several frames in one launch; returns dicts like oracle_py.pose_only."""
        cps, crs, keep, bads = packed or self.pnp_pack(pbs)
        self._chk(self.lib.uco_b200_pose_only_batch(self.h, len(bads), ctypes.addressof(cps), ctypes.addressof(crs)))
        return [dict(pose44=np.array(list(crs[i].pose44), np.float32), pose7=np.array(list(crs[i].pose7)), n_good=int(crs[i].n_good),
                     iters=np.array(list(crs[i].iters), np.int32), bad=bads[i].copy()) for i in range(len(bads))]

    def pose_only(self, pb):
        cps, crs, keep, bads = self.pnp_pack([pb])
        self._chk(self.lib.uco_b200_pose_only(self.h, ctypes.addressof(cps), ctypes.addressof(crs)))
        return dict(pose44=np.array(list(crs[0].pose44), np.float32), pose7=np.array(list(crs[0].pose7)), n_good=int(crs[0].n_good),
                    iters=np.array(list(crs[0].iters), np.int32), bad=bads[0].copy())

    # -- K9 ------------------------------------------------------------------------------------------------------
    def bow_load(self, voc_bytes):
        voc_bytes = np.ascontiguousarray(voc_bytes, np.uint8)
        h = ctypes.c_void_p()
        self._chk(self.lib.uco_b200_bow_load(self.h, _p(voc_bytes), len(voc_bytes), ctypes.addressof(h)))
        return h

    def bow_free(self, voc):
        self.lib.uco_b200_bow_free(self.h, voc)

    def bow_transform(self, voc, desc, level):
        """desc: (n,32) uint8 host rows. Returns per-descriptor (word, weight, node)."""
        desc = np.asarray(desc)
        n = desc.shape[0]
        word = np.empty(n, np.uint32); weight = np.empty(n, np.float32); node = np.empty(n, np.uint32)
        stride = desc.strides[0] if n else 32
        self._chk(self.lib.uco_b200_bow_transform(self.h, voc, _p(desc), n, stride, level, _p(word), _p(weight), _p(node)))
        return word, weight, node

    def bow_transform_dev(self, voc, desc_dev, n, level, word_dev, weight_dev, node_dev):
        self._chk(self.lib.uco_b200_bow_transform_dev(self.h, voc, desc_dev, n, level, word_dev, weight_dev, node_dev))

    def hamming_knn_batch_dev(self, n_pairs, q_dev, q_stride, nq_max, nq_dev, t_dev, t_stride, nt_max, nt_dev, k, order,
                              idx_dev, dist_dev):
        self._chk(self.lib.uco_b200_hamming_knn_batch_dev(self.h, n_pairs, q_dev, q_stride, nq_max, nq_dev, t_dev, t_stride,
                                                          nt_max, nt_dev, k, order, idx_dev, dist_dev))

    # -- K1-K6 ---------------------------------------------------------------------------------------------------
    def orb_extract(self, img, prm=None):
        """img: (h,w) uint8 host image (rows may be strided). Returns (keypoints[KP_DTYPE], desc[N,32])."""
        k, d, n = self.orb_extract_batch([img], prm)
        return k[0][:n[0]], d[0][:n[0]]

    def orb_extract_batch(self, imgs, prm=None):
        prm = prm or OrbParams()
        n = len(imgs)
        h, w = imgs[0].shape
        stride = imgs[0].strides[0]
        for im in imgs:
            assert im.dtype == np.uint8 and im.shape == (h, w) and im.strides == (stride, 1)
        cap = prm.max_features
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        nout = np.zeros(n, np.int32)
        ptrs = (ctypes.c_void_p * n)(*[im.ctypes.data for im in imgs])
        self._chk(self.lib.uco_b200_orb_extract_batch(self.h, ctypes.cast(ptrs, _vp), n, w, h, stride,
                                                      ctypes.addressof(prm), _p(kps), _p(desc), cap, _p(nout)))
        return kps, desc, nout

    def orb_extract_batch_dev(self, imgs_dev, n, w, h, pitch, frame_stride, prm, kps_dev, desc_dev, nout_dev):
        self._chk(self.lib.uco_b200_orb_extract_batch_dev(self.h, imgs_dev, n, w, h, pitch, frame_stride,
                                                          ctypes.addressof(prm), kps_dev, desc_dev, nout_dev))

    def set_profiling(self, on):
        self.lib.uco_b200_set_profiling(self.h, int(on))

    def orb_last_stage_ms(self):
        out = np.zeros(5, np.float32)
        self._chk(self.lib.uco_b200_orb_last_stage_ms(self.h, _p(out)))
        return dict(zip(["blur", "resize", "fast_cells", "select", "orient_describe"], out.tolist()))

    def orb_plan_bytes(self):
        out = np.zeros(3, np.uint64)
        self._chk(self.lib.uco_b200_orb_plan_bytes(self.h, _p(out)))
        return dict(zip(["input", "pyramid_px", "pyramid_bordered"], [int(v) for v in out]))

    def orb_level_info(self, level):
        v = [ctypes.c_int() for _ in range(6)]
        self._chk(self.lib.uco_b200_orb_debug_level_info(self.h, level, *[ctypes.addressof(x) for x in v]))
        return dict(zip(["w", "h", "pitch", "n_desired", "rows", "cols"], [x.value for x in v]))

    def orb_pyramid_level(self, frame, level):
        """Bordered buffer (h+38, w+38) of a level after the last extract call."""
        li = self.orb_level_info(level)
        buf = np.empty((li["h"] + 38, li["pitch"]), np.uint8)
        self._chk(self.lib.uco_b200_orb_debug_pyramid(self.h, frame, level, _p(buf)))
        return buf[:, :li["w"] + 38]

    def orb_selected(self, frame, level):
        out = np.empty(65536, np.uint32)
        n = ctypes.c_int()
        self._chk(self.lib.uco_b200_orb_debug_selected(self.h, frame, level, _p(out), len(out), ctypes.addressof(n)))
        return out[:n.value].copy()

    def orb_candidates(self, frame, cell):
        out = np.empty(16384, np.uint32)
        counts = np.zeros(2, np.int32)
        geom = np.zeros(6, np.int32)
        self._chk(self.lib.uco_b200_orb_debug_candidates(self.h, frame, cell, _p(out), len(out), _p(counts), _p(geom)))
        return out[:counts[0]].copy(), counts, geom

    def hamming_knn_dev(self, q_dev, nq, t_dev, nt, k, order, idx_dev, dist_dev):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous on the context stream."""
        self._chk(self.lib.uco_b200_hamming_knn_dev(self.h, q_dev, nq, t_dev, nt, k, order, idx_dev, dist_dev))


class TrackState:
    """uco_b200_track_state: device-resident mirror of the tracking state (previous frame + map block) of n independent streams."""

    def __init__(self, ctx, n_streams, prev_cap, map_cap):
        self.ctx, self.n, self.prev_cap, self.map_cap = ctx, n_streams, prev_cap, map_cap
        h = ctypes.c_void_p()
        ctx._chk(ctx.lib.uco_b200_track_state_create(ctx.h, n_streams, prev_cap, map_cap, ctypes.addressof(h)))
        self.h = h

    def close(self):
        if self.h:
            self.ctx.lib.uco_b200_track_state_free(self.ctx.h, self.h)
            self.h = None

    def set_scene(self, stream, sc):
        """previous frame + map block of one stream from a scene dict (synth.synth_track_scene keys)"""
        A = lambda k, dt: np.ascontiguousarray(sc[k], dt)
        pk = np.zeros(len(sc["prev_octave"]), KP_DTYPE)
        pk["octave"] = sc["prev_octave"]
        pd, pr = A("prev_desc", np.uint8), A("prev_mp_row", np.int32)
        c = self.ctx
        c._chk(c.lib.uco_b200_track_state_set_prev(c.h, self.h, stream, len(pk), _p(pk), _p(pd), _p(pr)))
        ids, pos, nrm = A("mp_id", np.uint32), A("mp_pos", np.float32), A("mp_normal", np.float32)
        dmin, dmax, mdesc = A("mp_min_dist", np.float32), A("mp_max_dist", np.float32), A("mp_desc", np.uint8)
        mp = MapPoints(len(ids), _p(ids), _p(pos), _p(nrm), _p(dmin), _p(dmax), _p(mdesc))
        st = A("mp_stable", np.uint8) if "mp_stable" in sc else None
        lo = A("mp_local", np.uint8) if "mp_local" in sc else None
        c._chk(c.lib.uco_b200_track_state_set_map(c.h, self.h, stream, ctypes.addressof(mp), _p(st) if st is not None else None,
                                                  _p(lo) if lo is not None else None))

    def track_frames(self, imgs, orb_prm, prm, pose_prior):
        """one step of every stream through host buffers: imgs (n, h, w) u8 -> (kps, desc, n_kp, list of result dicts)"""
        c = self.ctx
        imgs = np.ascontiguousarray(imgs, np.uint8)
        F, h, w = imgs.shape
        assert F == self.n
        mf = orb_prm.max_features
        ptrs = (ctypes.c_void_p * F)(*[imgs[i].ctypes.data for i in range(F)])
        prior = np.ascontiguousarray(pose_prior, np.float32).reshape(F, 16)
        kps, desc, nkp = np.zeros((F, mf), KP_DTYPE), np.zeros((F, mf, 32), np.uint8), np.zeros(F, np.int32)
        o = dict(matches=np.zeros((F, mf), MATCH_DTYPE), n_matches=np.zeros(F, np.int32), pose=np.zeros((F, 16), np.float32),
                 n_good=np.zeros(F, np.int32), status=np.zeros(F, np.int32), n_tbp=np.zeros(F, np.int32))
        to = TrackOut(*[_p(o[k]) for k in ("matches", "n_matches", "pose", "n_good", "status", "n_tbp")], None)
        c._chk(c.lib.uco_b200_track_frames(c.h, self.h, ctypes.cast(ptrs, ctypes.c_void_p), w, h, w, ctypes.addressof(orb_prm),
                                           ctypes.addressof(prm), _p(prior), _p(kps), _p(desc), _p(nkp), ctypes.addressof(to)))
        res = [dict(matches=o["matches"][f, :o["n_matches"][f]].copy(), pose44=o["pose"][f].copy(), n_good=int(o["n_good"][f]),
                    status=int(o["status"][f]), n_tbp=int(o["n_tbp"][f])) for f in range(F)]
        return kps, desc, nkp, res


class KeyFrameDataBase:
    """Host-side mirror of ucoslam::KeyFrameDataBase (src/map_types/keyframedatabase.h:31-52) over the device-resident database:
    add / delete / clear / size / is_id / relocalization_candidates with the reference's argument meaning.  A frame is passed as its
    bag of words (ids ascending, float weights): what Vocabulary::transform + the host fold produce.  `neighbors(frame_id)` stands
    for CovisGraph::getNeighborsWeights(frame_id, true): neighbour ids by decreasing weight."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.h = ctypes.c_void_p()
        ctx._chk(ctx.lib.uco_b200_kfdb_create(ctx.h, ctypes.addressof(self.h)))

    def close(self):
        if self.h:
            self.ctx.lib.uco_b200_kfdb_free(self.ctx.h, self.h)
            self.h = None

    def add(self, frame_id, words, weights):
        words = np.ascontiguousarray(words, np.uint32); weights = np.ascontiguousarray(weights, np.float32)
        self.ctx._chk(self.ctx.lib.uco_b200_kfdb_add(self.ctx.h, self.h, int(frame_id), _p(words), _p(weights), len(words)))

    def add_batch(self, frame_ids, bows):
        frame_ids = np.ascontiguousarray(frame_ids, np.uint32)
        counts = np.array([len(b[0]) for b in bows], np.int32)
        words = np.ascontiguousarray(np.concatenate([np.asarray(b[0], np.uint32) for b in bows]) if len(bows) else np.zeros(0, np.uint32))
        weights = np.ascontiguousarray(np.concatenate([np.asarray(b[1], np.float32) for b in bows]) if len(bows) else np.zeros(0, np.float32))
        self.ctx._chk(self.ctx.lib.uco_b200_kfdb_add_batch(self.ctx.h, self.h, len(frame_ids), _p(frame_ids), _p(counts), _p(words),
                                                           _p(weights)))

    def delete(self, frame_id):
        self.ctx._chk(self.ctx.lib.uco_b200_kfdb_del(self.ctx.h, self.h, int(frame_id)))

    def clear(self):
        self.ctx._chk(self.ctx.lib.uco_b200_kfdb_clear(self.ctx.h, self.h))

    def size(self):
        n, w = ctypes.c_uint32(), ctypes.c_uint64()
        self.ctx.lib.uco_b200_kfdb_size(self.h, ctypes.addressof(n), ctypes.addressof(w))
        return int(n.value), int(w.value)

    def is_id(self, frame_id):
        return bool(self.ctx.lib.uco_b200_kfdb_has(self.h, int(frame_id)))

    def query(self, words, weights, excluded=(), min_score=0.0):
        """steps 1-2: dict(frame, score, common, max_common) of the scored frames, ascending frame id"""
        words = np.ascontiguousarray(words, np.uint32); weights = np.ascontiguousarray(weights, np.float32)
        exc = np.ascontiguousarray(list(excluded), np.uint32)
        cap = max(self.size()[0], 1)
        if getattr(self, "_cap", 0) < cap:      # output buffers are kept between queries (a caller's std::vector would be too)
            self._cap = cap
            self._out = (np.zeros(cap, np.uint32), np.zeros(cap, np.float64), np.zeros(cap, np.uint32))
        fr, sc, cm = self._out
        n, mc = ctypes.c_int(), ctypes.c_uint32()
        self.ctx._chk(self.ctx.lib.uco_b200_kfdb_query(self.ctx.h, self.h, _p(words), _p(weights), len(words), _p(exc), len(exc),
                                                       float(min_score), _p(fr), _p(sc), _p(cm), cap, ctypes.addressof(n),
                                                       ctypes.addressof(mc)))
        return dict(frame=fr[:n.value].copy(), score=sc[:n.value].copy(), common=cm[:n.value].copy(), max_common=int(mc.value))

    def last_ms(self):
        out = np.zeros(2, np.float32)
        self.ctx.lib.uco_b200_kfdb_last_ms(self.h, _p(out))
        return out

    def relocalization_candidates(self, words, weights, neighbors, sorted_=True, min_score=0.0, excluded=()):
        q = self.query(words, weights, excluded, min_score)
        return rank_candidates(q["frame"], q["score"], neighbors, sorted_, min_score)


def rank_candidates(frame, score, neighbors, sorted_=True, min_score=0.0):
    """uco_b200_kfdb_rank: covisibility accumulation, 0.75*best gate, optional sort (keyframedatabase.cpp:236-275)."""
    frame = np.ascontiguousarray(frame, np.uint32); score = np.ascontiguousarray(score, np.float64)
    n = len(frame)
    lists = [np.asarray(neighbors(int(f)), np.uint32) for f in frame] if n > 1 else [np.zeros(0, np.uint32)] * n
    off = np.zeros(n + 1, np.int32)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    nbr = np.ascontiguousarray(np.concatenate(lists) if n else np.zeros(0, np.uint32), np.uint32)
    out = np.zeros(max(n, 1), np.uint32)
    no = ctypes.c_int()
    rc = load().uco_b200_kfdb_rank(_p(frame), _p(score), n, _p(off), _p(nbr), int(bool(sorted_)), float(min_score), _p(out),
                                   ctypes.addressof(no))
    if rc != 0:
        raise UcoError("uco_b200_kfdb_rank rc=%d" % rc)
    return out[:no.value].copy()
