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
