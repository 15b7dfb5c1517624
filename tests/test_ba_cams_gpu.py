"""Bundle adjustment over a window whose keyframes were taken with DIFFERENT cameras (uco_ba_problem::pose_cam): every edge carries the
ImageParams of its keyframe, as GlobalOptimizerG2O::setParams sets them per edge (globaloptimizer_g2o.cpp:233-236, :262-266, :335-338).
Against golden vectors produced by the reference's own g2o + typesg2o.h (tests/golden/ba_cams_g2o.npz, make_golden.py ba_cams) and
against the live reference where oracle/_ref exists.  Tolerances: check_ba (poses 1e-7, points 1e-5, identical LM decisions) for the
keypoint-only windows, check_markers for the window with ArUco markers (numeric Jacobians, see test_ba_markers_gpu.py)."""
import os, sys
import numpy as np
import pytest
import oracle_py
from test_ba_oracle import check_ba
from test_ba_markers_gpu import check_markers
from ucoslam_b200.synth import mix_cameras, synth_global_ba

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import BA_CAM_CASES, BA_MARKER_KEYS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ba_cams_g2o.npz")


def _case(name):
    g = np.load(GOLD)
    keys = oracle_py.BA_INPUT_KEYS + ("pose_cam",) + (BA_MARKER_KEYS if BA_CAM_CASES[name][1] else ())
    pb = {k: g["%s_in_%s" % (name, k)] for k in keys}
    for k in ("fx", "fy", "cx", "cy", "bf"):
        pb[k] = float(pb[k])
    ref = {k[len(name) + 5:]: g[k] for k in g.files if k.startswith(name + "_out_")}
    return pb, ref, BA_CAM_CASES[name][3]


@pytest.mark.parametrize("name", list(BA_CAM_CASES))
def test_mixed_cameras_match_reference_golden(ctx, name):
    pb, ref, iters = _case(name)
    check = check_markers if BA_CAM_CASES[name][1] else check_ba
    check(ctx.ba_solve(pb, iters), ref)            # the plain entry point routes the window to the solver that reads the per-keyframe table
    check(ctx.ba_solve_sharded(pb, iters), ref)
    assert len(np.unique(pb["pose_cam"][:, 0])) == 2


def test_the_camera_table_matters(ctx):
    """the same observations solved with ONE camera for every keyframe end somewhere else: the table is read, not ignored"""
    pb, ref, iters = _case("cam_mono")
    one = dict(pb)
    one.pop("pose_cam")
    got = ctx.ba_solve(one, iters)
    assert np.abs(got["pose7"] - ref["pose7"]).max() > 1e-3


def test_equal_rows_equal_the_single_camera_solve(ctx):
    """a table that repeats fx..bf for every keyframe gives bit-identical results to the single-camera problem (sharded solver)"""
    pb, _, iters = _case("cam_mono")
    one = dict(pb)
    one.pop("pose_cam")
    same = dict(one, pose_cam=np.tile(np.array([pb["fx"], pb["fy"], pb["cx"], pb["cy"], pb["bf"]], np.float32), (len(pb["fixed"]), 1)))
    a, b = ctx.ba_solve_sharded(one, iters), ctx.ba_solve(same, iters)
    assert np.array_equal(a["pose7"], b["pose7"]) and np.array_equal(a["point3"], b["point3"]) and np.array_equal(a["iters"], b["iters"])


def test_mixed_cameras_in_a_batch(ctx):
    """a batch mixing ordinary windows (cluster-resident solver) with a two-camera window"""
    from ucoslam_b200.synth import synth_ba_problem
    pb, ref, iters = _case("cam_stereo")
    plain = synth_ba_problem(9, n_poses=6, n_fixed=1, n_points=150)
    outs = ctx.ba_solve_batch([plain, pb, plain], iters)
    check_ba(outs[1], ref)
    assert np.array_equal(outs[0]["pose7"], outs[2]["pose7"])
    check_ba(outs[0], oracle_py.ba_optimize(plain, iters))


def test_mixed_cameras_on_a_loop_graph_match_live_reference(ctx):
    pb = mix_cameras(synth_global_ba(8, n_kf=40, n_points=1500), seed=11)
    ref = oracle_py.ref_ba_optimize(pb, 5)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    check_ba(ctx.ba_solve(pb, 5), ref, tol_pose=1e-6, tol_pt=1e-4, tol_chi=1e-5)
