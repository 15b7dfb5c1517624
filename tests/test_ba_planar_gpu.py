"""Bundle adjustment with the InPlaneMarkers option (globaloptimizer_g2o.cpp:356-401): one planar edge (MarkerEdgeX, :37-66) between the
map's reference marker and every other marker of the window, the reference inside the window (a free vertex: marker-to-marker blocks in
the reduced system) or outside it (a fixed vertex).  Against golden vectors from the reference's g2o with the edge class restated in
oracle/ref_g2o_wrap.cpp (tests/golden/ba_planar_g2o.npz) and the live reference where oracle/_ref exists.
Tolerances: check_markers (test_ba_markers_gpu.py).  g2o differentiates this edge numerically with delta = 1e-9f, so ITS Jacobian carries
1e-7-relative noise that depends on the last bits of Eigen's 4x4 inverse; the product evaluates the same residuals in closed form, the two
solutions agree to the marker tolerances below, not to 1e-7."""
import os, sys
import numpy as np
import pytest
import oracle_py
from test_ba_markers_gpu import check_markers
from ucoslam_b200.synth import add_markers, add_plane_edges, synth_global_ba

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import BA_PLANE_CASES, BA_PLANE_KEYS, BA_MARKER_KEYS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ba_planar_g2o.npz")


def _case(name):
    g = np.load(GOLD)
    pb = {k: g["%s_in_%s" % (name, k)] for k in oracle_py.BA_INPUT_KEYS + BA_MARKER_KEYS + BA_PLANE_KEYS if "%s_in_%s" % (name, k) in g.files}
    for k in ("fx", "fy", "cx", "cy", "bf", "plane_weight"):
        pb[k] = float(pb[k])
    pb["plane_ref"] = int(pb["plane_ref"])
    ref = {k[len(name) + 5:]: g[k] for k in g.files if k.startswith(name + "_out_")}
    return pb, ref, BA_PLANE_CASES[name][3]


def _planarity(marker_pose44, ref44):
    Ri = np.linalg.inv(np.asarray(ref44, np.float64).reshape(4, 4))
    M = np.array([Ri @ np.asarray(m, np.float64).reshape(4, 4) for m in marker_pose44])
    return np.abs(np.c_[M[:, 0, 2], M[:, 1, 2], 1 - M[:, 2, 2], M[:, 2, 3]]).max()


@pytest.mark.parametrize("name", list(BA_PLANE_CASES))
def test_planar_edges_match_reference_golden(ctx, name):
    pb, ref, iters = _case(name)
    got = ctx.ba_solve(pb, iters)
    check_markers(got, ref)
    check_markers(ctx.ba_solve_sharded(pb, iters), ref)
    # the option does what it is for: the markers end closer to one plane than without the edges
    plain = {k: v for k, v in pb.items() if not k.startswith("plane_")}
    ref44 = got["marker_pose44"][pb["plane_ref"]] if pb["plane_ref"] >= 0 else pb["plane_ref_pose44"]
    free = ctx.ba_solve(plain, iters)
    ref44_free = free["marker_pose44"][pb["plane_ref"]] if pb["plane_ref"] >= 0 else pb["plane_ref_pose44"]
    assert _planarity(got["marker_pose44"], ref44) < _planarity(free["marker_pose44"], ref44_free)


def test_planar_edges_on_a_loop_graph_match_live_reference(ctx):
    pb = add_plane_edges(add_markers(synth_global_ba(8, n_kf=40, n_points=1500), seed=9, n_markers=6, coplanar=True), True)
    ref = oracle_py.ref_ba_optimize(pb, 5)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    got = ctx.ba_solve(pb, 5)
    # as test_markers_on_a_loop_graph_match_live_reference: a 40-keyframe loop held by two fixed keyframes has centimetre-level play along the
    # loop between equally good solutions; the two optimisers must agree on chi2 at EVERY iteration and on the iteration / LM-trial counts
    assert np.array_equal(got["iters"], ref["iters"])
    n = int(ref["iters"].sum())
    assert np.array_equal(got["trace"][:n, 1], ref["trace"][:n, 1])
    assert np.allclose(got["trace"][:n, 0], ref["trace"][:n, 0], rtol=5e-5)
    assert np.abs(got["pose7"] - ref["pose7"]).max() < 2e-3 and np.abs(got["marker_pose7"] - ref["marker_pose7"]).max() < 5e-3


def test_malformed_planar_arrays_are_rejected(ctx):
    pb, _, iters = _case("plane_in")
    bad = dict(pb, plane_other=np.array([pb["plane_ref"]], np.int32))     # an edge from the reference to itself
    with pytest.raises(Exception):
        ctx.ba_solve(bad, iters)
    bad = dict(pb, plane_other=np.array([99], np.int32))
    with pytest.raises(Exception):
        ctx.ba_solve(bad, iters)
