"""Pins oracle_pose_only (oracle/ba_oracle.c) to the reference's pose-only optimisation: golden vectors produced by the
reference's own g2o + typesg2o.h edge classes under PnPSolver::solvePnp's schedule (tests/golden/make_golden.py pnp ->
pnp_g2o.npz), and a live comparison where oracle/_ref exists.  Tolerances (f64 vs f64): pose 1e-9 without markers; 1e-6 with
markers (their numeric Jacobian differences float-rounded projections, so a last-bit change flips a 3e-5 px quantum);
identical LM iteration counts per round, inlier count and outlier flags (away from the chi2 gate)."""
import os, sys
import numpy as np
import pytest
import oracle_py

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pnp_g2o.npz")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import PNP_CASES


def check_pnp(got, ref, markers=False):
    tol = 1e-6 if markers else 1e-9
    assert np.array_equal(got["iters"], ref["iters"]), (got["iters"], ref["iters"])
    assert np.abs(got["pose7"] - ref["pose7"]).max() < tol
    assert np.abs(got["pose44"] - ref["pose44"]).max() < 1e-6
    assert int(got["n_good"]) == int(ref["n_good"])
    assert np.array_equal(got["bad"], ref["bad"])


def golden_ref(g, name):
    return {k[len(name) + 5:]: g[k] for k in g.files if k.startswith(name + "_out_")}


@pytest.mark.parametrize("name", list(PNP_CASES))
def test_oracle_matches_reference_golden(name):
    g = np.load(GOLD)
    pb = oracle_py.pnp_problem_from_golden(g, name)
    check_pnp(oracle_py.pose_only(pb), golden_ref(g, name), markers=len(pb["marker_size"]) > 0)


def test_golden_inputs_are_reproducible():
    g = np.load(GOLD)
    for name, kw in PNP_CASES.items():
        pb = oracle_py.synth_pnp_problem(**kw)
        for k in oracle_py.PNP_INPUT_KEYS:
            assert np.array_equal(np.asarray(pb[k]), g["%s_in_%s" % (name, k)]), (name, k)


def test_empty_problem_returns_zero():
    pb = oracle_py.synth_pnp_problem(seed=9, n_matches=0, n_markers=0)
    r = oracle_py.pose_only(pb)
    assert r["n_good"] == 0 and len(r["bad"]) == 0


def test_oracle_matches_live_reference():
    if oracle_py.load_ref("libref_g2o.so") is None:
        pytest.skip("oracle/_ref not built")
    for kw in (dict(seed=31, n_matches=1200, stereo_frac=0.2, outlier_frac=0.3), dict(seed=32, n_matches=100, n_markers=5),
               dict(seed=33, n_matches=2000, unstable_frac=0.8)):
        pb = oracle_py.synth_pnp_problem(**kw)
        check_pnp(oracle_py.pose_only(pb), oracle_py.ref_pose_only(pb), markers=kw.get("n_markers", 0) > 0)
