"""Stereo depth association (FrameExtractor::processStereo's loop).  The reference needs OpenCV C++ and cannot be
compiled here: parity for this row is UNPINNED BY THE REFERENCE.  What is checked: the plain-C restatement (oracle/stereo_oracle.c)
against an independent numpy restatement that makes the loop's two OpenCV calls (absdiff, sum) through cv2, bit for bit."""
import numpy as np
import pytest
import oracle_py
from ucoslam_b200.synth import synth_stereo

CASES = {"vga": dict(seed=1), "small": dict(seed=2, w=320, h=240, n=600, max_disp=30.0),
         "cluttered": dict(seed=3, n=2500, outlier_frac=0.4, tie_frac=0.3), "odd": dict(seed=5, w=333, h=201, n=400, max_disp=20.0)}


def edge_scene():
    """keypoints on and beyond the image border, rows outside the image, an empty right side"""
    sc = synth_stereo(6, w=160, h=120, n=200, max_disp=12.0)
    kl, kr = sc["kps_l"], sc["kps_r"]
    kl["y"][:5] = [-3.0, 119.6, 0.4, 2.4, 117.0]
    kl["x"][5:9] = [1.0, 2.5, 157.0, 156.4]
    kr["y"][:4] = [-0.6, 119.5, 119.4, 200.0]
    return sc


@pytest.mark.parametrize("name", list(CASES))
def test_c_oracle_matches_numpy_cv2_restatement(name):
    sc = synth_stereo(**CASES[name])
    d, m, n = oracle_py.stereo_depth(sc)
    d2, m2, n2 = oracle_py.stereo_depth_py(sc)
    assert n == n2 and n > 50 and (m >= 0).sum() > n            # some associations die in the SAD stage
    assert np.array_equal(m, m2) and np.array_equal(d.view(np.uint32), d2.view(np.uint32))
    assert (d[d != 0] > 0).all()


def test_threshold_and_edges():
    sc = edge_scene()
    for md in (50.0, 20.0, 1.0, 300.0):
        d, m, n = oracle_py.stereo_depth(sc, md)
        d2, m2, n2 = oracle_py.stereo_depth_py(sc, md)
        assert n == n2 and np.array_equal(m, m2) and np.array_equal(d.view(np.uint32), d2.view(np.uint32))
    empty = dict(sc)
    empty["kps_r"], empty["desc_r"] = sc["kps_r"][:0], sc["desc_r"][:0]
    d, m, n = oracle_py.stereo_depth(empty)
    assert n == 0 and not d.any() and (m == -1).all()


def test_depth_recovers_the_rendered_disparity():
    """sanity of the synthetic pair itself: where a depth comes out it is bl*fx/disparity of that image row (within the keypoint noise)"""
    sc = synth_stereo(1)
    d, m, n = oracle_py.stereo_depth(sc)
    h = sc["img_l"].shape[0]
    rows = np.clip(np.round(sc["kps_l"]["y"]).astype(int), 0, h - 1)
    disp = 8.0 + (48.0 - 8.0) * (0.5 + 0.5 * np.sin(rows / h * 2 * np.pi))
    ok = d != 0
    got = sc["bl"] * sc["fx"] / d[ok]
    assert np.median(np.abs(got - disp[ok])) < 0.75
