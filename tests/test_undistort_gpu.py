"""undistort.cu through the C ABI against the reference's own OpenCV call (cv2.undistortPoints + float rescale): bit-exact, for
dense point arrays and in place on the pt field of keypoint records (Frame::und_kpts)."""
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from test_undistort_oracle import K, DISTS, points

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(DISTS))
def test_matches_opencv(ctx, name):
    p = points()
    want = oracle_py.undistort_points_py(p, K, DISTS[name])
    got = ctx.undistort_points(p, K, DISTS[name])
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


def test_keypoint_records_and_edges(ctx):
    p = points(1000, 3)
    kps = np.zeros(len(p), ucoslam_b200.KP_DTYPE)
    kps["x"], kps["y"] = p[:, 0], p[:, 1]
    kps["octave"] = np.arange(len(p)) % 8; kps["angle"] = 33.0; kps["response"] = 7.0; kps["size"] = 31.0; kps["class_id"] = -1
    und = ctx.undistort_keypoints(kps, K, DISTS["five"])
    want = oracle_py.undistort_points_py(p, K, DISTS["five"])
    assert np.array_equal(und["x"].view(np.uint32), want[:, 0].view(np.uint32)) and np.array_equal(und["y"].view(np.uint32), want[:, 1].view(np.uint32))
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(und[f], kps[f])
    assert ctx.undistort_points(np.zeros((0, 2), np.float32), K, DISTS["five"]).shape == (0, 2)
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.undistort_points(p, K, [0.1] * 14)          # tilt terms are not supported
