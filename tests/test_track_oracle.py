"""CPU tests (no GPU): the libstdc++ std::sort replay used by the device kd-tree build against the real std::sort, and the
tracking-sequence oracle's internal consistency (search by projection + filter against a brute-force restatement in numpy)."""
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from ucoslam_b200 import synth


@pytest.mark.parametrize("n,levels", [(0, 1), (1, 1), (2, 1), (15, 3), (16, 2), (17, 2), (40, 3), (100, 4), (333, 5), (1000, 7), (5000, 2), (5000, 50)])
def test_sort_replay_matches_libstdcxx(n, levels):
    rng = np.random.default_rng(n * 31 + levels)
    for rep in range(6):
        keys = rng.integers(0, levels, n).astype(np.float32) * np.float32(1.25)   # massive ties: the order of equal keys is the point
        if rep == 4:
            keys = np.sort(keys)
        if rep == 5:
            keys = np.sort(keys)[::-1].copy()
        ref = np.arange(n, dtype=np.uint32)
        oracle_py.load_stl().stl_sort_indices(oracle_py._p(ref), n, oracle_py._p(keys))
        got = ucoslam_b200.probe_sort_indices(keys)
        assert np.array_equal(got, ref)


def test_sort_replay_heapsort_fallback():
    # a "median-of-3 killer" drives introsort into its depth limit (heap sort path)
    n = 4096
    keys = np.zeros(n, np.float32)
    k = n // 2
    for i in range(1, k + 1):
        if i & 1:
            keys[i - 1] = i
            keys[i] = k + i
        keys[k + i - 1] = 2 * i
    ref = np.arange(n, dtype=np.uint32)
    oracle_py.load_stl().stl_sort_indices(oracle_py._p(ref), n, oracle_py._p(keys))
    assert np.array_equal(ucoslam_b200.probe_sort_indices(keys), ref)


def _tbp_numpy(sc, dist_thr, proj):
    """independent restatement: radius search results in the oracle's own kd-tree visit order, bookkeeping in numpy"""
    f32 = np.float32
    pose = np.asarray(sc["pose44"], f32).reshape(4, 4)
    kxy, koct, kdesc = np.asarray(sc["kp_xy"], f32), np.asarray(sc["kp_octave"]), np.asarray(sc["kp_desc"], np.uint8)
    sf = np.asarray(sc["scale_factors"], f32)
    qs, rad, who = [], [], []
    for i, r in enumerate(sc["prev_mp_row"]):
        if r < 0:
            continue
        P = np.asarray(sc["mp_pos"], f32)[r]
        rz = f32(f32(f32(P[0] * pose[2, 0]) + f32(P[1] * pose[2, 1])) + f32(P[2] * pose[2, 2])) + pose[2, 3]
        if rz < 0:
            continue
        rx = f32(f32(f32(P[0] * pose[0, 0]) + f32(P[1] * pose[0, 1])) + f32(P[2] * pose[0, 2])) + pose[0, 3]
        ry = f32(f32(f32(P[0] * pose[1, 0]) + f32(P[1] * pose[1, 1])) + f32(P[2] * pose[1, 2])) + pose[1, 3]
        iz = f32(1.0 / np.float64(rz)) if rz != 0 else f32(np.inf)
        x = f32(f32(f32(f32(sc["fx"]) * rx) * iz) + f32(sc["cx"]))
        y = f32(f32(f32(f32(sc["fy"]) * ry) * iz) + f32(sc["cy"]))
        if not (x >= sc["min_xy"][0] and y >= sc["min_xy"][1] and x < sc["max_xy"][0] and y < sc["max_xy"][1]):
            continue
        qs.append((x, y)); rad.append(f32(f32(proj) * sf[sc["prev_octave"][i]])); who.append(i)
    ptr, idx = oracle_py.kdtree_radius(kxy, np.array(qs, f32).reshape(-1, 2), np.array(rad, f32))
    out = []
    for k, i in enumerate(who):
        best, best2, bk = f32(np.float64(f32(dist_thr)) + 0.01), f32(np.finfo(f32).max), -1
        for kp in idx[ptr[k]:ptr[k + 1]]:
            if koct[kp] != sc["prev_octave"][i]:
                continue
            d = f32(np.unpackbits(kdesc[kp] ^ np.asarray(sc["prev_desc"], np.uint8)[i]).sum())
            if d < best:
                best, bk = d, kp
            elif d < best2:
                best2 = d
        if bk >= 0 and np.float64(best) < 0.7 * np.float64(best2):
            out.append((bk, int(sc["mp_id"][sc["prev_mp_row"][i]]), -1, best))
    m = np.array(out, oracle_py.MATCH_DT)
    # filter_ambiguous_query: per keypoint the smallest distance, the earlier entry on ties; order kept
    keep = np.ones(len(m), bool)
    first = {}
    for j, mm in enumerate(m):
        q = int(mm["queryIdx"])
        if q not in first:
            first[q] = j
        elif m[first[q]]["distance"] > mm["distance"]:
            keep[first[q]] = False
            first[q] = j
        else:
            keep[j] = False
    return m[keep]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_track_projected_oracle_vs_numpy(seed):
    sc = synth.synth_track_scene(seed, n_kp=600, n_mp=900, n_prev=500)
    a = oracle_py.track_projected(sc, 75.0, 15.0)
    b = _tbp_numpy(sc, 75.0, 15.0)
    assert len(a) > 50
    assert np.array_equal(a, b)


def test_track_frame_oracle_recovers_pose():
    sc = synth.synth_track_scene(5)
    r = oracle_py.track_frame(sc)
    T = sc["pose_true44"].reshape(4, 4)
    assert r["status"] == 0 and r["n_tbp"] > 30 and r["n_good"] > 500
    assert np.abs(r["pose44"].reshape(4, 4) - T).max() < 0.2 * np.abs(sc["pose44"].reshape(4, 4) - T).max()
    # every keypoint at most once after the final filter
    assert len(np.unique(r["matches"]["queryIdx"])) == len(r["matches"])


def test_track_frame_oracle_fallback_flag():
    sc = synth.synth_track_scene(6, n_kp=300, n_mp=400, n_prev=20)
    r = oracle_py.track_frame(sc)
    assert r["status"] & 1 and r["n_tbp"] <= 30
