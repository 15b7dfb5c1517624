"""Host-side structure of a BA window (no GPU): the planner behind uco_b200_ba_solve must enumerate exactly the block
structure g2o's BlockSolver::buildStructure would (3rdparty/g2o/g2o/core/block_solver.hpp:103-312): one Schur contribution
per (landmark, pose pair i <= j both free), one 6x6 block per co-observing pose pair, chunks of whole landmarks."""
import ctypes
import numpy as np
import pytest
import ucoslam_b200
from ucoslam_b200.synth import synth_ba_problem


@pytest.mark.parametrize("kw", [dict(seed=1, n_poses=12, n_fixed=2, n_points=400), dict(seed=2, n_poses=30, n_fixed=5, n_points=300),
                                dict(seed=3, n_poses=4, n_fixed=4, n_points=50), dict(seed=4, n_poses=6, n_fixed=0, n_points=1)])
@pytest.mark.parametrize("cluster", [1, 8, 16])
def test_plan_counts(kw, cluster):
    lib = ucoslam_b200.load()
    pb = synth_ba_problem(**kw)
    cp, cr, keep, out = ucoslam_b200.Context.ba_pack(pb, 5)
    o = np.zeros(8, np.int32)
    assert lib.uco_b200_probe_ba_plan(ctypes.addressof(cp), cluster, o.ctypes.data) == 0
    free = pb["fixed"] == 0
    pf = int(free.sum())
    obs_free = free[pb["obs_pose"]]
    kf = np.bincount(pb["obs_point"][obs_free], minlength=len(pb["points3"]))
    pairs = set()
    for l in range(len(pb["points3"])):
        ps = sorted(pb["obs_pose"][(pb["obs_point"] == l) & obs_free])
        pairs.update((a, b) for i, a in enumerate(ps) for b in ps[i:])
    pairs.update((p, p) for p in np.nonzero(free)[0])
    assert o[0] == pf
    assert o[1] == len(pairs)                              # Schur blocks (upper triangle incl. diagonal)
    assert o[3] == int((kf * (kf + 1) // 2).sum())         # contributions
    assert o[2] >= o[1] - pf and o[2] * 64 >= o[3]         # units cover them
    assert o[5] == int(obs_free.sum())
    assert o[6] <= 512 and o[7] == len(pb["points3"])      # chunks fit a CTA and cover every landmark


def test_plan_rejects_bad_indices():
    lib = ucoslam_b200.load()
    pb = synth_ba_problem(5, n_poses=5, n_fixed=1, n_points=20)
    pb["obs_point"] = pb["obs_point"].copy()
    pb["obs_point"][0] = 10 ** 6
    cp, cr, keep, out = ucoslam_b200.Context.ba_pack(pb, 5)
    assert lib.uco_b200_probe_ba_plan(ctypes.addressof(cp), 8, None) == -1
