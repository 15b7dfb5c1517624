"""ransac.cu through the C ABI against the cv2-backed restatement of PnPSolver::solvePnPRansac on the SAME 4-match samples.
Floating point, stated tolerance: per iteration the inlier count may differ by the number of matches whose test lies within rounding
of a threshold (the restatement reports them); the winner must be a maximiser
of the restatement's counts within that slack, its pose must agree with the restatement's pose of the same iteration to 1e-5."""
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from ucoslam_b200.synth import synth_reloc_matches
from test_ransac_oracle import make_samples

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n,iters", [(1, 400, 300), (2, 400, 300), (3, 1500, 500), (4, 60, 100)])
def test_matches_restatement(ctx, seed, n, iters):
    sc = synth_reloc_matches(seed, n=n)
    smp = make_samples(seed, n, iters)
    want = oracle_py.pnp_ransac_py(sc, smp)
    got = ctx.pnp_ransac(sc, iters, smp)
    assert got["ok"] == want["ok"]
    # "no P3P solution" (-1 here; OpenCV answers true with a NaN pose for samples without a real root, which then has no inliers)
    # and "no inliers" are the same outcome for the loop
    diff = np.abs(np.maximum(got["counts"], 0) - np.maximum(want["counts"], 0))
    assert (diff <= want["borderline"] + 1).all() and (diff == 0).mean() > 0.95
    b = got["best_iter"]
    assert got["counts"][b] == got["counts"].max() and (got["counts"][:b] < got["counts"][b]).all()      # first maximiser of its own counts
    assert want["counts"][b] >= want["counts"].max() - want["borderline"].max() - 1
    assert len(got["inliers"]) == got["counts"][b]
    # the pose of that iteration as the restatement computes it
    ref_b = oracle_py.pnp_ransac_py(sc, smp[b:b + 1])
    assert np.abs(got["pose44"] - ref_b["pose44"]).max() < 1e-5
    sym = np.setxor1d(got["inliers"], ref_b["inliers"])
    assert len(sym) <= want["borderline"][b] + 1
    assert np.abs(got["pose44"][:3, 3] - sc["pose_gt"][:3, 3]).max() < 0.05


def test_internal_sampler_and_edges(ctx):
    sc = synth_reloc_matches(5, n=500)
    r1 = ctx.pnp_ransac(sc, 400, None, seed=11)
    r2 = ctx.pnp_ransac(sc, 400, None, seed=11)
    r3 = ctx.pnp_ransac(sc, 400, None, seed=12)
    assert r1["ok"] and np.array_equal(r1["counts"], r2["counts"]) and not np.array_equal(r1["counts"], r3["counts"])
    for r in (r1, r3):
        assert len(r["inliers"]) > 200 and np.abs(r["pose44"][:3, 3] - sc["pose_gt"][:3, 3]).max() < 0.05
    few = dict(sc, p3d=sc["p3d"][:3], p2d=sc["p2d"][:3], normals=sc["normals"][:3])
    assert not ctx.pnp_ransac(few, 50)["ok"]                                   # fewer than 4 matches -> false
    junk = dict(sc, p2d=np.random.default_rng(0).uniform(0, 480, sc["p2d"].shape).astype(np.float32))
    assert not ctx.pnp_ransac(junk, 50, None, seed=1)["ok"] or len(ctx.pnp_ransac(junk, 50, None, seed=1)["inliers"]) < 12
