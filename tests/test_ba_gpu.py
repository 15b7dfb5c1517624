"""Parity of the sm_100a bundle adjustment (uco_b200_ba_solve, through the C ABI) against
  - the golden vectors produced by the reference's own g2o + typesg2o.h (tests/golden/ba_g2o.npz),
  - the plain-C oracle (oracle/ba_oracle.c) on seeded problems of BASELINE config-2 size (12 KF window, ~15k observations),
  - the reference itself where oracle/_ref was built.
Tolerances are those of tests/test_ba_oracle.py (f64 vs f64; only the summation order differs): poses 1e-7, points 1e-5
absolute on a ~5 m scene, identical iteration / LM-trial counts, identical outlier and bad-association sets away from the gate."""
import os, sys
import numpy as np
import pytest
import oracle_py
from test_ba_oracle import check_ba, GOLD, BA_CASES

pytestmark = pytest.mark.gpu


MODES = {"streamed": (1, 0), "cluster8": (2, 8), "cluster4": (2, 4), "cluster16": (2, 16), "cluster1": (2, 1)}


@pytest.fixture(params=list(MODES))
def bctx(ctx, request):
    """the context with the BA solver forced into one of its two forms (streamed kernels / cluster-resident kernel)"""
    ctx.ba_set_mode(*MODES[request.param])
    yield ctx
    ctx.ba_set_mode(0, 0)


@pytest.mark.parametrize("name", list(BA_CASES))
def test_ba_matches_reference_golden(bctx, name):
    g = np.load(GOLD)
    pb = oracle_py.ba_problem_from_golden(g, name)
    got = bctx.ba_solve(pb, BA_CASES[name][1])
    ref = {k[len(name) + 5:]: g[k] for k in g.files if k.startswith(name + "_out_")}
    check_ba(got, ref)


@pytest.mark.parametrize("kw,iters", [
    (dict(seed=21, n_poses=12, n_fixed=2, n_points=2000), 5),                       # config 2: local BA window
    (dict(seed=22, n_poses=12, n_fixed=2, n_points=1500, stereo_frac=0.4), 5),
    (dict(seed=23, n_poses=30, n_fixed=5, n_points=1200, outlier_frac=0.05), 10),
    (dict(seed=24, n_poses=3, n_fixed=3, n_points=50), 5),                          # every pose fixed: points only
    (dict(seed=25, n_poses=45, n_fixed=1, n_points=600), 3),                        # reduced system too big for shared memory
])
def test_ba_matches_oracle(ctx, kw, iters):
    pb = oracle_py.synth_ba_problem(**kw)
    ref = oracle_py.ref_ba_optimize(pb, iters) or oracle_py.ba_optimize(pb, iters)
    for mode in ((1, 0), (0, 0)):  # streamed, then automatic (cluster-resident when the reduced system fits)
        ctx.ba_set_mode(*mode)
        got = ctx.ba_solve(pb, iters)
        ctx.ba_set_mode(0, 0)
        check_ba(got, ref)


def test_ba_batch_equals_single_solves(ctx):
    """a batch is one launch with one cluster per window; every window gets exactly the result of a single solve"""
    pbs = [oracle_py.synth_ba_problem(40 + i, n_poses=6 + 2 * i, n_fixed=1 + i % 2, n_points=150 + 60 * i, stereo_frac=0.1 * i)
           for i in range(5)]
    singles = [ctx.ba_solve(pb, 5) for pb in pbs]
    batch = ctx.ba_solve_batch(pbs, 5)
    for a, b in zip(singles, batch):
        for k in ("pose7", "pose44", "point3", "chi2", "level", "bad", "trace", "iters"):
            assert np.array_equal(a[k], b[k]), k
    with pytest.raises(Exception):
        ctx.ba_set_mode(2, 8)
        try:
            ctx.ba_solve(oracle_py.synth_ba_problem(25, n_poses=45, n_fixed=1, n_points=100), 1)  # 44 free KFs: not cluster-resident
        finally:
            ctx.ba_set_mode(0, 0)


def test_ba_is_bitwise_reproducible(bctx):
    pb = oracle_py.synth_ba_problem(31, n_poses=10, n_fixed=2, n_points=800, stereo_frac=0.2)
    a, b = bctx.ba_solve(pb, 5), bctx.ba_solve(pb, 5)
    for k in ("pose7", "point3", "chi2", "trace"):
        assert np.array_equal(a[k], b[k]), k


def test_ba_observation_order_does_not_matter_much(ctx):
    """the library sorts observations by landmark itself; a shuffled input gives the same answer to round-off"""
    pb = oracle_py.synth_ba_problem(32, n_poses=8, n_fixed=2, n_points=500)
    a = ctx.ba_solve(pb, 5)
    perm = np.random.default_rng(0).permutation(len(pb["obs_pose"]))
    pb2 = dict(pb)
    for k in ("obs_pose", "obs_point", "obs_uv", "obs_ur", "obs_stereo", "obs_inv_sigma2"):
        pb2[k] = np.ascontiguousarray(pb[k][perm])
    b = ctx.ba_solve(pb2, 5)
    assert np.abs(a["pose7"] - b["pose7"]).max() < 1e-9
    assert np.abs(a["chi2"][perm] - b["chi2"]).max() < 1e-6
    assert np.array_equal(a["level"][perm], b["level"])


def test_ba_stop_flag_and_errors(bctx):
    import ucoslam_b200
    ctx = bctx
    pb = oracle_py.synth_ba_problem(33, n_poses=6, n_fixed=1, n_points=200)
    stop = np.ones(1, np.uint8)  # the reference's bool stopASAP
    out = ctx.ba_solve(pb, 5, stop=stop)
    assert out["iters"].tolist() == [0, 0]
    assert np.abs(out["pose44"] - pb["poses44"]).max() < 1e-6  # nothing moved
    bad = dict(pb)
    bad["obs_pose"] = pb["obs_pose"].copy()
    bad["obs_pose"][3] = 99
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.ba_solve(bad, 5)
