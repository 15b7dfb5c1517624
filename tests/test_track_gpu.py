"""Parity of the sm_100a tracking sequence (through the C ABI) with the CPU restatements:
  * uco_b200_kdtree_build_dev (device build of Frame::keypoint_kdtree) == uco_b200_kdtree_build (host restatement, itself pinned node
    for node to the reference's picoflann.h in tests/test_project_oracle.py): identical node arrays, leaf lists and boxes;
  * uco_b200_track_projected == oracle_track_projected (System's search by projection from the previous frame): identical cv::DMatch
    lists in the reference's order;
  * uco_b200_track_batch == oracle_py.track_frame (the tracker's main branch assembled from the stage oracles): identical match lists
    incl. the inlier flags, n_tbp / status, visible flags; poses to 1e-5 (f32 results of an f64 LM whose sums are ordered differently).
"""
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from ucoslam_b200 import synth

pytestmark = pytest.mark.gpu


def _points(seed, n, kind):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        xy = rng.uniform([16, 16], [624, 464], (n, 2))
    elif kind == "grid":          # many equal coordinates: degenerate cuts -> the std::sort path and its tie order
        xy = np.stack([rng.integers(2, 40, n) * 16.0, rng.integers(2, 30, n) * 16.0], 1)
    elif kind == "line":
        xy = np.stack([rng.uniform(16, 600, n), np.full(n, 100.0)], 1)
    elif kind == "clumps":
        c = rng.uniform([50, 50], [590, 430], (max(1, n // 40), 2))
        xy = c[rng.integers(0, len(c), n)] + rng.normal(0, 1.5, (n, 2))
    elif kind == "same":
        xy = np.full((n, 2), 77.25)
    return xy.astype(np.float32)


@pytest.mark.parametrize("n", [0, 1, 9, 10, 11, 20, 21, 35, 100, 199, 200, 201, 1000, 2000, 4000, 4096])
@pytest.mark.parametrize("kind", ["uniform", "grid", "clumps"])
def test_kdtree_dev_equals_host(ctx, n, kind):
    xy = _points(n * 7 + len(kind), n, kind)
    hn, hl, hb = ucoslam_b200.kdtree_build(xy)
    dn, dl, db = ctx.kdtree_build_dev(xy)
    assert len(hn) == len(dn)
    assert np.array_equal(hl, dl)
    assert np.array_equal(hn, dn)
    assert np.array_equal(np.asarray(hb), np.asarray(db))


@pytest.mark.parametrize("kind,n", [("line", 700), ("same", 300), ("same", 25)])
def test_kdtree_dev_degenerate(ctx, kind, n):
    xy = _points(5, n, kind)
    hn, hl, hb = ucoslam_b200.kdtree_build(xy)
    dn, dl, db = ctx.kdtree_build_dev(xy)
    assert np.array_equal(hl, dl) and np.array_equal(hn, dn) and np.array_equal(np.asarray(hb), np.asarray(db))


def test_kdtree_dev_on_orb_keypoints(ctx):
    rng = np.random.default_rng(11)
    img = np.kron(rng.integers(0, 256, (30, 40)), np.ones((16, 16))).astype(np.uint8)
    img = np.clip(img.astype(np.int32) + rng.integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
    kps, _ = ctx.orb_extract(img, ucoslam_b200.OrbParams(2000))
    xy = np.stack([kps["x"], kps["y"]], 1)
    hn, hl, hb = ucoslam_b200.kdtree_build(xy)
    dn, dl, db = ctx.kdtree_build_dev(xy)
    assert len(hn) > 100 and np.array_equal(hl, dl) and np.array_equal(hn, dn) and np.array_equal(np.asarray(hb), np.asarray(db))


def test_kdtree_dev_too_many_points(ctx):
    xy = _points(1, ucoslam_b200.UCO_KDTREE_DEV_MAX_POINTS + 1, "uniform")
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.kdtree_build_dev(xy)


@pytest.mark.parametrize("seed,kw,thr", [(1, {}, (75.0, 15.0)), (2, dict(n_kp=4000, n_mp=5000, n_prev=3500), (75.0, 15.0)),
                                          (3, dict(n_kp=300, n_mp=2000, n_prev=1500, dup_frac=0.5), (120.0, 30.0)),
                                          (4, dict(clutter=0.9), (150.0, 40.0)), (5, dict(n_kp=12, n_mp=50, n_prev=40), (256.0, 100.0)),
                                          (6, {}, (30.0, 2.5))])
def test_track_projected_matches_oracle(ctx, seed, kw, thr):
    sc = synth.synth_track_scene(seed, **kw)
    got = ctx.track_projected(sc, *thr)
    ref = oracle_py.track_projected(sc, *thr)
    assert np.array_equal(got, ref)
    if seed < 5:
        assert len(got) > 30


def test_track_projected_with_device_tree(ctx):
    sc = synth.synth_track_scene(9)
    tree = ctx.kdtree_build_dev(sc["kp_xy"])
    assert np.array_equal(ctx.track_projected(sc, 75.0, 15.0, tree), oracle_py.track_projected(sc, 75.0, 15.0))


def test_track_projected_empty(ctx):
    sc = synth.synth_track_scene(2, n_kp=100, n_mp=100, n_prev=50)
    sc["prev_mp_row"][:] = -1
    assert len(ctx.track_projected(sc, 75.0, 15.0)) == 0
    sc2 = dict(sc, prev_octave=sc["prev_octave"][:0], prev_desc=sc["prev_desc"][:0], prev_mp_row=sc["prev_mp_row"][:0])
    assert len(ctx.track_projected(sc2, 75.0, 15.0)) == 0


def _check_frame(got, ref):
    assert got["n_tbp"] == ref["n_tbp"] and got["status"] == ref["status"]
    assert np.array_equal(got["visible"], ref["visible"])
    assert np.array_equal(got["matches"], ref["matches"])
    assert got["n_good"] == ref["n_good"]
    assert np.abs(got["pose44"] - ref["pose44"]).max() < 1e-5


def test_track_batch_matches_oracle(ctx):
    scs = [synth.synth_track_scene(20 + i, n_kp=[2000, 1500, 1000, 600][i % 4], n_mp=[3000, 2500, 900, 1200][i % 4],
                                   n_prev=[1800, 700, 900, 400][i % 4]) for i in range(8)]
    got = ctx.track_batch(scs)
    for g, sc in zip(got, scs):
        ref = oracle_py.track_frame(sc)
        _check_frame(g, ref)
        assert ref["n_good"] > 100


def test_track_batch_branches(ctx):
    """frames whose first search finds <= 30 matches (status bit 0), whose first solvePnp fails (bit 1), and an empty map block"""
    few = synth.synth_track_scene(31, n_kp=800, n_mp=1000, n_prev=25)
    bad = synth.synth_track_scene(32, n_kp=800, n_mp=1000, n_prev=600, prior_noise=(0.004, 0.01))
    bad["mp_pos"] = bad["mp_pos"].copy()
    rows = bad["prev_mp_row"][bad["prev_mp_row"] >= 0]
    bad["mp_pos"][rows] += np.random.default_rng(1).normal(0, 0.004, (len(rows), 3)).astype(np.float32)   # stage-1 geometry off
    nothing = synth.synth_track_scene(33, n_kp=500, n_mp=600, n_prev=300)
    nothing["prev_mp_row"][:] = -1
    nothing["mp_local"][:] = 0
    ok = synth.synth_track_scene(34)
    scs = [few, bad, nothing, ok]
    got = ctx.track_batch(scs)
    refs = [oracle_py.track_frame(s) for s in scs]
    assert refs[0]["status"] & 1
    assert refs[2]["n_good"] == 0 and len(refs[2]["matches"]) == 0
    for g, r in zip(got, refs):
        _check_frame(g, r)


def test_track_batch_is_deterministic(ctx):
    scs = [synth.synth_track_scene(40 + i) for i in range(3)]
    a, b = ctx.track_batch(scs), ctx.track_batch(scs)
    for x, y in zip(a, b):
        assert np.array_equal(x["matches"], y["matches"]) and np.array_equal(x["pose44"], y["pose44"])
