"""Parity of the sm_100a projection matcher (uco_b200_match_projected, through the C ABI) with the restatement of
Map::matchFrameToMapPoints: identical cv::DMatch records in the reference's order and identical setVisible() flags, on the golden
scenes and on seeded scenes incl. the edge cases (no map points, no keypoints, nothing visible, every map point on one keypoint)."""
import os
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from ucoslam_b200.synth import synth_projection_scene

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "project_match.npz")


def check(ctx, sc, min_desc=50.0, max_reproj=15.0, tree=None):
    got, gv = ctx.match_projected(sc, min_desc, max_reproj, tree)
    ref, rv = oracle_py.match_projected(sc, min_desc, max_reproj)
    assert np.array_equal(gv, rv)
    assert np.array_equal(got, ref)
    return got


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_matches_golden(ctx, name):
    g = np.load(GOLD)
    sc = {k[len(name) + 1:]: g[k] for k in g.files if k.startswith(name + "_") and not k.startswith(name + "_out_")}
    for k in ("fx", "fy", "cx", "cy"):
        sc[k] = float(sc[k])
    got, vis = ctx.match_projected(sc, float(g[name + "_out_min_desc"]), float(g[name + "_out_max_reproj"]))
    assert np.array_equal(got, g[name + "_out_matches"]) and np.array_equal(vis, g[name + "_out_visible"])


@pytest.mark.parametrize("seed,kw,thr", [(1, {}, (50.0, 15.0)), (2, dict(n_kp=4000, n_mp=5000), (50.0, 15.0)), (3, dict(n_kp=300, n_mp=2000, dup_frac=0.5), (80.0, 30.0)),
                                          (4, dict(n_kp=2000, n_mp=3000, clutter=0.9), (100.0, 40.0)), (5, dict(n_kp=12, n_mp=50), (256.0, 100.0)),
                                          (6, dict(n_kp=2000, n_mp=3000), (20.0, 2.5))])
def test_matches_oracle(ctx, seed, kw, thr):
    got = check(ctx, synth_projection_scene(seed, **kw), *thr)
    if seed < 5:
        assert len(got) > 50


def test_tree_from_the_references_stream(ctx):
    """the tree handed over as the bytes Frame::toStream holds (KdTreeIndex::toStream), parsed by the library"""
    sc = synth_projection_scene(7)
    stream = oracle_py.ref_picoflann_stream(sc["kp_xy"])
    if stream is None:
        pytest.skip("oracle/_ref/libref_picoflann.so not built")
    check(ctx, sc, tree=ucoslam_b200.kdtree_parse(stream))


def test_edge_cases(ctx):
    sc = synth_projection_scene(8, n_kp=500, n_mp=400)
    empty = dict(sc)
    for k in ("mp_id", "mp_pos", "mp_normal", "mp_min_dist", "mp_max_dist", "mp_desc"):
        empty[k] = sc[k][:0]
    got, vis = ctx.match_projected(empty, 50.0, 15.0)
    assert len(got) == 0 and len(vis) == 0
    nokp = dict(sc)
    for k in ("kp_xy", "kp_octave", "kp_desc"):
        nokp[k] = sc[k][:0]
    got, vis = ctx.match_projected(nokp, 50.0, 15.0)
    assert len(got) == 0 and not vis.any()
    away = dict(sc)
    away["mp_normal"] = -sc["mp_normal"]          # every point faces away: viewCos < 0.5
    check(ctx, away)
    one = dict(sc)                                 # every map point carries keypoint 0's descriptor and projects onto it
    one["mp_desc"] = np.tile(sc["kp_desc"][0], (len(sc["mp_id"]), 1))
    check(ctx, one, 256.0, 1000.0)
    bad = dict(sc)
    nodes, leaf, bbox = ucoslam_b200.kdtree_build(sc["kp_xy"])
    leaf = leaf.copy(); leaf[0] = 10 ** 6
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.match_projected(bad, 50.0, 15.0, (nodes, leaf, bbox))


def test_map_points_from_a_map_file_section(ctx):
    """the candidate map points of the projection matcher taken from the map-point SECTION of a map file (ReusableContainer<MapPoint>,
    Map::toStream) without building MapPoint objects: the scene's points are written as a container with erased slots in between, read back as flat
    rows (uco_b200_mappoints_from_container) and matched; same matches as from the scene's own arrays"""
    import ctypes
    from ucoslam_b200 import MapPointStream, MapPointContainer, MatView, mappoints_from_container
    lib = ucoslam_b200.load()
    sc = synth_projection_scene(11, n_kp=1200, n_mp=900)
    m = len(sc["mp_id"])
    n_slots = 200 * ((m + 60 + 199) // 200)
    rng = np.random.default_rng(1)
    slots = np.sort(rng.choice(m + 60, m, replace=False))            # the live points sit in these slots, erased ones in between
    views, valid = (MapPointStream * n_slots)(), np.zeros(n_slots, np.uint8)
    for i in range(n_slots):
        lib.uco_b200_mappoint_stream_default(ctypes.addressof(views[i]))
    desc = np.ascontiguousarray(sc["mp_desc"], np.uint8)
    for k, s in enumerate(slots):
        v = views[int(s)]
        v.id = int(sc["mp_id"][k])
        for a in range(3):
            v.pos3d[a], v.normal[a] = float(sc["mp_pos"][k][a]), float(sc["mp_normal"][k][a])
        v.min_distance, v.max_distance = float(sc["mp_min_dist"][k]), float(sc["mp_max_dist"][k])
        v.desc = MatView(1, 32, 0, desc[k].ctypes.data)
        valid[int(s)] = 1
    free = np.array([i for i in range(m + 60) if not valid[i]], np.uint32)
    c = MapPointContainer(n_slots, m + 60, m, len(free), free.ctypes.data)
    need = ctypes.c_size_t()
    assert lib.uco_b200_mappoint_container_write(ctypes.addressof(c), ctypes.addressof(views), valid.ctypes.data, None, 0, ctypes.addressof(need)) == 0
    buf = np.zeros(need.value, np.uint8)
    assert lib.uco_b200_mappoint_container_write(ctypes.addressof(c), ctypes.addressof(views), valid.ctypes.data, buf.ctypes.data, len(buf), ctypes.addressof(need)) == 0
    mp = mappoints_from_container(buf)
    assert len(mp["ids"]) == m and np.array_equal(mp["ids"], np.asarray(sc["mp_id"], np.uint32)) and np.array_equal(mp["desc"], desc)
    sc2 = dict(sc, mp_id=mp["ids"], mp_pos=mp["pos"], mp_normal=mp["normal"], mp_min_dist=mp["min_dist"], mp_max_dist=mp["max_dist"], mp_desc=mp["desc"])
    a, va = ctx.match_projected(sc, 50.0, 15.0)
    b, vb = ctx.match_projected(sc2, 50.0, 15.0)
    assert len(a) > 100 and a.tobytes() == b.tobytes() and np.array_equal(va, vb)
