"""Parity of the sm_100a pose-only optimisation (uco_b200_pose_only{,_batch}, through the C ABI) against
  - the golden vectors produced by the reference's own g2o + typesg2o.h under solvePnp's schedule (tests/golden/pnp_g2o.npz),
  - the plain-C oracle on seeded problems of tracker size (up to 2000 matches), and the reference where oracle/_ref exists.
Tolerances are those of tests/test_pnp_oracle.py: pose 1e-9 (1e-6 with marker edges), identical LM iteration counts per
round, identical inlier count and outlier flags."""
import numpy as np
import pytest
import oracle_py
from test_pnp_oracle import check_pnp, golden_ref, GOLD, PNP_CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(PNP_CASES))
def test_pnp_matches_reference_golden(ctx, name):
    g = np.load(GOLD)
    pb = oracle_py.pnp_problem_from_golden(g, name)
    check_pnp(ctx.pose_only(pb), golden_ref(g, name), markers=len(pb["marker_size"]) > 0)


@pytest.mark.parametrize("kw", [
    dict(seed=41, n_matches=2000, outlier_frac=0.15),                     # config 2: one tracked frame
    dict(seed=42, n_matches=1500, stereo_frac=0.5, unstable_frac=0.6),
    dict(seed=43, n_matches=257, n_markers=20),                           # more markers than 16-lane groups
    dict(seed=44, n_matches=9),                                           # below the 10-inlier stop
    dict(seed=45, n_matches=700, pose_noise=(0.2, 8.0), outlier_frac=0.3),
])
def test_pnp_matches_oracle(ctx, kw):
    pb = oracle_py.synth_pnp_problem(**kw)
    ref = oracle_py.ref_pose_only(pb) or oracle_py.pose_only(pb)
    check_pnp(ctx.pose_only(pb), ref, markers=kw.get("n_markers", 0) > 0)


def test_pnp_empty(ctx):
    pb = oracle_py.synth_pnp_problem(seed=9, n_matches=0, n_markers=0)
    r = ctx.pose_only(pb)
    assert r["n_good"] == 0 and np.array_equal(r["pose44"], pb["pose44"])


def test_pnp_batch_equals_single_and_is_reproducible(ctx):
    """a batch is one launch with one thread block per frame; results are those of single calls, bit for bit, run to run"""
    pbs = [oracle_py.synth_pnp_problem(seed=60 + i, n_matches=100 + 300 * i, stereo_frac=0.1 * i, n_markers=i % 3) for i in range(7)]
    pbs.insert(3, oracle_py.synth_pnp_problem(seed=9, n_matches=0, n_markers=0))
    singles = [ctx.pose_only(pb) for pb in pbs]
    for _ in range(2):
        batch = ctx.pose_only_batch(pbs)
        for a, b in zip(singles, batch):
            for k in ("pose7", "pose44", "iters", "bad"):
                assert np.array_equal(a[k], b[k]), k
            assert a["n_good"] == b["n_good"]


def test_pnp_recovers_ground_truth(ctx):
    """size-independent property: from a perturbed start the optimised pose returns to the generating pose (px noise 0.5)"""
    pb = oracle_py.synth_pnp_problem(seed=77, n_matches=2000, outlier_frac=0.2)
    r = ctx.pose_only(pb)
    assert np.abs(r["pose44"].reshape(4, 4) - pb["pose_gt"]).max() < 5e-3
    assert r["n_good"] > 1500
