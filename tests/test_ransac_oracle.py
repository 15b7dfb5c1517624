"""RANSAC P3P pose (PnPSolver::solvePnPRansac).  CPU-side checks: the product's three-point solver (p3p_math.h,
compiled for the host behind uco_b200_probe_p3p) against OpenCV's own cv::solvePnP(SOLVEPNP_P3P) through cv2 -- the call the
reference makes at pnpsolver.cpp:67 -- and the consistency of the cv2-backed restatement of the RANSAC loop."""
import numpy as np
import pytest
import cv2
import oracle_py
import ucoslam_b200
from ucoslam_b200.synth import synth_reloc_matches, _rodrigues

K = np.array([525.0, 525.0, 319.5, 239.5])
KM = np.array([[525.0, 0, 319.5], [0, 525.0, 239.5], [0, 0, 1]])


def make_samples(seed, n, iters):
    rng = np.random.default_rng(seed)
    return np.array([rng.choice(n, 4, replace=False) for _ in range(iters)], np.int32)


def test_p3p_matches_opencv():
    rng = np.random.default_rng(0)
    exist_diff, worst, big = 0, 0.0, 0
    for trial in range(1500):
        R = _rodrigues(rng.uniform(-0.5, 0.5, 3)); t = rng.uniform(-1, 1, 3) + np.array([0, 0, 1.0])
        X = rng.uniform(-2, 2, (4, 3)) + np.array([0, 0, 6.0])
        Xc = X @ R.T + t
        uv = Xc[:, :2] / Xc[:, 2:] * 525.0 + np.array([319.5, 239.5])
        if trial % 2:
            uv += rng.normal(0, 0.5, (4, 2))
        ok, rv, tv = cv2.solvePnP(X.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), KM, None, flags=cv2.SOLVEPNP_P3P)
        mine = ucoslam_b200.probe_p3p(X, uv, K)
        if ok != (mine is not None):
            exist_diff += 1
            continue
        if ok:
            d = max(np.abs(cv2.Rodrigues(rv)[0] - mine[0]).max(), np.abs(tv.ravel() - mine[1]).max())
            worst = max(worst, d)
            big += d > 1e-6
            if trial % 2 == 0:      # noise-free: the generating pose is one of the solutions and the 4th point selects it
                assert np.abs(mine[0] - R).max() < 1e-5 and np.abs(mine[1] - t).max() < 1e-4
            # a rotation, and the three sample points reproject exactly
            assert np.abs(mine[0] @ mine[0].T - np.eye(3)).max() < 1e-10 and abs(np.linalg.det(mine[0]) - 1) < 1e-10
            Xm = X[:3] @ mine[0].T + mine[1]
            assert np.abs(Xm[:, :2] / Xm[:, 2:] * 525.0 + np.array([319.5, 239.5]) - uv[:3]).max() < 1e-4
    assert exist_diff <= 5 and big <= 8 and worst < 1e-4    # stated tolerance: poses agree to 1e-6, rare near-degenerate samples aside


def test_p3p_degenerate_inputs():
    X = np.array([[0, 0, 5.0], [1, 0, 5.0], [2, 0, 5.0], [0, 1, 5.0]])      # first three collinear
    uv = X[:, :2] / X[:, 2:] * 525.0 + np.array([319.5, 239.5])
    assert ucoslam_b200.probe_p3p(X, uv, K) is None
    X[1] = X[0]                                                             # coincident points
    assert ucoslam_b200.probe_p3p(X, uv, K) is None


@pytest.mark.parametrize("seed", [1, 2])
def test_restated_loop_finds_the_pose(seed):
    sc = synth_reloc_matches(seed)
    smp = make_samples(seed, len(sc["p3d"]), 200)
    r = oracle_py.pnp_ransac_py(sc, smp)
    assert r["ok"] and r["counts"].max() == len(r["inliers"]) and r["counts"][r["best_iter"]] == r["counts"].max()
    assert (r["counts"][:r["best_iter"]] < r["counts"].max()).all()          # the FIRST maximiser
    assert len(r["inliers"]) > 0.4 * len(sc["p3d"])
    assert np.abs(r["pose44"][:3, :3] - sc["pose_gt"][:3, :3]).max() < 0.01 and np.abs(r["pose44"][:3, 3] - sc["pose_gt"][:3, 3]).max() < 0.05
    few = dict(sc, p3d=sc["p3d"][:3], p2d=sc["p2d"][:3], normals=sc["normals"][:3])
    assert not oracle_py.pnp_ransac_py(few, np.zeros((0, 4), np.int32))["ok"]
