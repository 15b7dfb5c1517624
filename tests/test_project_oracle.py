"""CPU checks of the projection matcher's test infrastructure and host logic:
  - the kd-tree restatement in oracle/project_oracle.cpp and the product's host-side builder / stream parser are node-for-node
    what the reference's own picoflann.h builds (oracle/_ref/libref_picoflann.so, compiled from /root/reference),
  - radius searches report keypoints in the reference's visit order,
  - the golden vectors of Map::matchFrameToMapPoints (tests/golden/project_match.npz, made by tests/golden/make_golden.py from the
    oracle) are reproduced by the oracle (guards the fixture)."""
import os
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from ucoslam_b200.synth import synth_projection_scene

GOLD = os.path.join(os.path.dirname(__file__), "golden", "project_match.npz")


def point_sets():
    rng = np.random.default_rng(0)
    for n in (1, 5, 10, 11, 25, 200, 2000, 5000):
        xy = rng.uniform(0, 640, (n, 2)).astype(np.float32)
        if n >= 200:
            xy[::7] = xy[3]                  # exact duplicates
        if n >= 25:
            xy[:, 0] = np.round(xy[:, 0])    # many ties along one axis (the std::sort fallback)
        yield xy
    yield np.full((64, 2), 7.5, np.float32)  # all points identical
    yield synth_projection_scene(3)["kp_xy"]


def as_dict(nodes, leaf, bbox):
    return dict(nodes=np.stack([nodes["col"], nodes["left"], nodes["right"], nodes["leaf_begin"], nodes["leaf_count"]], 1).astype(np.int32),
                div=np.stack([nodes["divlow"], nodes["divhigh"]], 1), leaf_idx=leaf, bbox=bbox)


def same_tree(a, b):
    leaf = a["nodes"][:, 1] < 0
    assert np.array_equal(a["nodes"][:, 1:3], b["nodes"][:, 1:3])
    assert np.array_equal(a["nodes"][~leaf, 0], b["nodes"][~leaf, 0]) and np.array_equal(a["div"][~leaf], b["div"][~leaf])
    assert np.array_equal(a["nodes"][leaf, 4], b["nodes"][leaf, 4])
    for (ba, ca), (bb, cb) in zip(a["nodes"][leaf, 3:], b["nodes"][leaf, 3:]):   # each leaf owns the same keypoints in the same order
        assert np.array_equal(a["leaf_idx"][ba:ba + ca], b["leaf_idx"][bb:bb + cb])
    assert np.array_equal(np.asarray(a["bbox"]), np.asarray(b["bbox"]))


@pytest.mark.parametrize("k", range(10))
def test_kdtree_restatement_and_product_builder_equal_the_references_picoflann(k):
    xy = list(point_sets())[k]
    stream = oracle_py.ref_picoflann_stream(xy)
    if stream is None:
        pytest.skip("oracle/_ref/libref_picoflann.so not built")
    ref = oracle_py.parse_picoflann_stream(stream)
    same_tree(oracle_py.kdtree_build(xy), ref)                       # oracle restatement
    same_tree(as_dict(*ucoslam_b200.kdtree_build(xy)), ref)          # product: host-side builder
    same_tree(as_dict(*ucoslam_b200.kdtree_parse(stream)), ref)      # product: parser of the reference's stream format
    rng = np.random.default_rng(k)
    q = rng.uniform(-50, 700, (300, 2)).astype(np.float32)
    r = rng.uniform(0.5, 60, 300).astype(np.float32)
    a, b = oracle_py.kdtree_radius(xy, q, r), oracle_py.ref_picoflann_radius(xy, q, r)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_kdtree_build_without_the_reference():
    """the product's builder against the restatement (runs where oracle/_ref does not exist)"""
    for xy in point_sets():
        same_tree(as_dict(*ucoslam_b200.kdtree_build(xy)), oracle_py.kdtree_build(xy))
    nodes, leaf, bbox = ucoslam_b200.kdtree_build(np.zeros((0, 2), np.float32))
    assert len(nodes) == 0


def test_oracle_reproduces_golden():
    g = np.load(GOLD)
    for name in ("a", "b", "c"):
        sc = {k[len(name) + 1:]: g[k] for k in g.files if k.startswith(name + "_") and not k.startswith(name + "_out_")}
        for k in ("fx", "fy", "cx", "cy"):
            sc[k] = float(sc[k])
        m, vis = oracle_py.match_projected(sc, float(g[name + "_out_min_desc"]), float(g[name + "_out_max_reproj"]))
        assert np.array_equal(m, g[name + "_out_matches"]) and np.array_equal(vis, g[name + "_out_visible"])
        assert len(m) > 100
