"""Point undistortion (ucoslam::undistortPoints).  The checker is the reference's own OpenCV call made through cv2
(cv2.undistortPoints + the float rescale of misc.cpp:283-290); the product's arithmetic (undistort_math.h, compiled for the host
behind uco_b200_probe_undistort) must reproduce it bit for bit."""
import numpy as np
import pytest
import oracle_py
import ucoslam_b200

K = np.array([525.3, 517.8, 319.5, 239.5], np.float32)
DISTS = {"none": [], "k1k2p1p2": [-0.28, 0.07, 0.0002, 0.00002], "five": [0.1, -0.2, 0.001, -0.002, 0.05],
         "rational": [0.2, -0.3, 0.001, 0.002, 0.1, 0.01, -0.02, 0.003],
         "prism": [0.1, -0.2, 0.001, -0.002, 0.05, 0.01, 0.02, 0.003, 0.001, -0.001, 0.002, 0.0005],
         "extreme": [-2.5, 0.5, 0, 0, 0]}           # icdist < 0 for points far from the centre: OpenCV returns the input ray


def points(n=4000, seed=0):
    rng = np.random.default_rng(seed)
    p = np.c_[rng.uniform(-50, 700, n), rng.uniform(-50, 530, n)].astype(np.float32)
    p[:4] = [[0, 0], [640, 480], [319.5, 239.5], [1e4, -1e4]]
    return p


@pytest.mark.parametrize("name", list(DISTS))
def test_host_arithmetic_matches_opencv(name):
    p = points()
    want = oracle_py.undistort_points_py(p, K, DISTS[name])
    got = ucoslam_b200.probe_undistort(p, K, DISTS[name])
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    if name == "none":
        assert np.abs(got[:3] - p[:3]).max() < 1e-3
