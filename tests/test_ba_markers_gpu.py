"""Bundle adjustment with ArUco markers (globaloptimizer_g2o.cpp:304-350): free marker vertices + MarkerEdges with numeric Jacobians,
solved by the streamed / sharded solver, against golden vectors produced by the reference's own g2o and its OWN MarkerEdge class
(tests/golden/ba_markers_g2o.npz) and against the live reference where oracle/_ref exists.  Tolerances: see check_markers."""
import os, sys
import numpy as np
import pytest
import oracle_py
from test_ba_oracle import check_ba
from ucoslam_b200.synth import add_markers, synth_global_ba

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import BA_MARKER_CASES, BA_MARKER_KEYS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ba_markers_g2o.npz")


def check_markers(got, ref, slack=1.0):
    """MarkerEdge narrows its projections to float and differentiates them numerically with delta = 1e-4 (typesg2o.h:121,157-161):
    a last-bit difference in the f64 pose algebra can flip a float rounding (3e-5 px), which the 1/(2 delta) = 5000 factor turns into
    a 1e-4-relative difference of a Jacobian entry.  The two optimisers therefore follow paths that differ at the 1e-6 level in
    chi2 (checked at EVERY iteration: a wrong sign or weight would show in the first step): same iteration and LM-trial counts,
    keyframes to 1e-4, points to 1e-3, and the weakly constrained marker poses (a 15-25 cm square seen from 2-3 m) to 3e-3 m / rad on
    a ~5 m scene; identical outlier sets away from the gates."""
    assert np.array_equal(got["iters"], ref["iters"])
    n = int(ref["iters"].sum())
    assert np.array_equal(got["trace"][:n, 1], ref["trace"][:n, 1]), "LM trials per iteration"
    assert np.allclose(got["trace"][:n, 0], ref["trace"][:n, 0], rtol=5e-5)
    d = dict(pose=np.abs(got["pose7"] - ref["pose7"]).max(), point=np.abs(got["point3"] - ref["point3"]).max(),
             marker=np.abs(got["marker_pose7"] - ref["marker_pose7"]).max())
    print("max differences vs g2o:", d)
    assert d["pose"] < 1e-4 * slack and d["point"] < 1e-3 * slack and d["marker"] < 3e-3 * slack
    assert np.abs(got["marker_pose44"] - ref["marker_pose44"]).max() < 6e-3 * slack
    print("marker edge chi2: sum", got["mobs_chi2"].sum(), ref["mobs_chi2"].sum(), "max |d|", np.abs(got["mobs_chi2"] - ref["mobs_chi2"]).max(),
          "max", ref["mobs_chi2"].max())
    assert abs(got["mobs_chi2"].sum() - ref["mobs_chi2"].sum()) < 2e-3 * min(slack, 10.0) * ref["mobs_chi2"].sum() + 1e-3
    assert np.allclose(got["mobs_chi2"], ref["mobs_chi2"], rtol=5e-2 * min(slack, 4.0), atol=2e-2 * min(slack, 4.0))
    near_gate = np.minimum(np.abs(ref["chi2"] - 5.99), np.abs(ref["chi2"] - 7.815)) < 0.05
    assert np.allclose(got["chi2"][~near_gate], ref["chi2"][~near_gate], rtol=2e-2, atol=1e-3)
    assert np.array_equal(got["level"][~near_gate], ref["level"][~near_gate])
    assert np.array_equal(got["bad"][~near_gate], ref["bad"][~near_gate])


@pytest.mark.parametrize("name", list(BA_MARKER_CASES))
def test_markers_match_reference_golden(ctx, name):
    g = np.load(GOLD)
    pb = {k: g["%s_in_%s" % (name, k)] for k in oracle_py.BA_INPUT_KEYS + BA_MARKER_KEYS}
    for k in ("fx", "fy", "cx", "cy", "bf"):
        pb[k] = float(pb[k])
    ref = {k[len(name) + 5:]: g[k] for k in g.files if k.startswith(name + "_out_")}
    iters = BA_MARKER_CASES[name][2]
    check_markers(ctx.ba_solve_sharded(pb, iters), ref)
    check_markers(ctx.ba_solve(pb, iters), ref)          # the plain entry point routes marker problems to the same solver
    assert np.abs(ref["marker_pose44"] - pb["marker_pose44"]).max() > 1e-3   # the markers moved


def test_markers_on_a_loop_graph_match_live_reference(ctx):
    pb = add_markers(synth_global_ba(8, n_kf=40, n_points=1500), seed=9, n_markers=6)
    ref = oracle_py.ref_ba_optimize(pb, 5)
    if ref is None:
        pytest.skip("oracle/_ref/libref_g2o.so not built")
    got = ctx.ba_solve_sharded(pb, 5)
    # a 40-keyframe loop held by two fixed keyframes: the two optimisers must agree on chi2 at EVERY iteration and on the iteration /
    # LM-trial counts; between equally good solutions there is centimetre-level play along the loop, so states are compared loosely
    assert np.array_equal(got["iters"], ref["iters"])
    n = int(ref["iters"].sum())
    assert np.array_equal(got["trace"][:n, 1], ref["trace"][:n, 1])
    assert np.allclose(got["trace"][:n, 0], ref["trace"][:n, 0], rtol=5e-5)
    assert np.abs(got["pose7"] - ref["pose7"]).max() < 2e-2 and np.abs(got["marker_pose7"] - ref["marker_pose7"]).max() < 5e-2
    assert abs(got["mobs_chi2"].sum() - ref["mobs_chi2"].sum()) < 0.05 * ref["mobs_chi2"].sum()
    assert (got["level"] != ref["level"]).mean() < 0.01
    gt = pb["marker_gt"][:, :3, 3]
    err_got = np.abs(got["marker_pose44"].reshape(-1, 4, 4)[:, :3, 3] - gt).max()
    err_ref = np.abs(ref["marker_pose44"].reshape(-1, 4, 4)[:, :3, 3] - gt).max()
    assert abs(err_got - err_ref) < 0.01      # as far from the ground truth as the reference ends up


def test_marker_input_errors(ctx):
    import ucoslam_b200
    pb = add_markers(oracle_py.synth_ba_problem(seed=64, n_poses=5, n_fixed=1, n_points=80), seed=3, n_markers=2)
    bad = dict(pb)
    bad["mobs_marker"] = pb["mobs_marker"].copy()
    bad["mobs_marker"][0] = 7
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.ba_solve_sharded(bad, 3)
