"""uco_b200_new_points (one keyframe against its neighbours: epipolar matching, triangulation, the mapper's gates and merge, SURVEY.md
8f rank 2) against the restatement of MapManager::createNewPoints (oracle_py.new_points_py).

Parity: the match lists are bit-exact (the matcher is, tests/test_match_gpu.py).  Triangulated points differ from the cv2 float-SVD
restatement by float-SVD accuracy: relative 2e-4 of the depth on points away from the gates (as tests/test_triangulate_gpu.py); a
match within rounding of a gate (margin < 1, see oracle_py.triangulate_py) may fall on either side and is excluded from the
point-set comparison."""
import numpy as np
import pytest
import ucoslam_b200
import oracle_py
from ucoslam_b200.synth import synth_new_points_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ucoslam_b200.Context(0)
    yield c
    c.close()


def same_matches(a, b):
    return len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in ("queryIdx", "trainIdx", "distance"))


@pytest.mark.parametrize("seed,kw", [(3, dict(n_kp=800, n_nb=4)), (11, dict(n_kp=2000, n_nb=20)), (12, dict(n_kp=1200, n_nb=1)),
                                      (13, dict(n_kp=1500, n_nb=7, assigned_frac=0.8))])
def test_unit_matches_restatement(ctx, seed, kw):
    sc = synth_new_points_scene(seed, **kw)
    ref = oracle_py.new_points_py(sc)
    got = ctx.new_points(sc, per_pair=True)
    F = len(sc["q_desc"])
    unsure = set()
    for f in range(F):
        assert same_matches(got["matches"][f], ref["matches"][f])
        g, r, mg = got["xyz_pairs"][f], ref["xyz_pairs"][f], ref["margin"][f]
        firm = mg >= 1.0
        assert np.array_equal(np.isnan(g[firm, 0]), np.isnan(r[firm, 0]))
        both = firm & ~np.isnan(r[:, 0])
        depth = np.linalg.norm(r[both] - np.asarray(sc["g2f_kf"])[:3, 3], axis=1)
        assert np.all(np.linalg.norm(g[both] - r[both], axis=1) <= 2e-4 * depth + 1e-5)
        unsure |= set(ref["matches"][f]["trainIdx"][~firm].tolist())
    # merged points: identical set of keyframe keypoints / observations wherever no borderline match is involved
    sel_g = ~np.isin(got["kpt"], list(unsure))
    sel_r = ~np.isin(ref["kpt"], list(unsure))
    assert np.array_equal(got["kpt"][sel_g], ref["kpt"][sel_r])
    assert np.array_equal(got["dist"][sel_g], ref["dist"][sel_r])
    for jg, jr in zip(np.nonzero(sel_g)[0], np.nonzero(sel_r)[0]):
        a, b = slice(got["obs_ptr"][jg], got["obs_ptr"][jg + 1]), slice(ref["obs_ptr"][jr], ref["obs_ptr"][jr + 1])
        assert np.array_equal(got["obs_frame"][a], ref["obs_frame"][b]) and np.array_equal(got["obs_kpt"][a], ref["obs_kpt"][b])
    assert sel_r.sum() > 0.5 * max(len(sc["t_map"]), 1) * 0.1
    assert np.all(np.diff(got["kpt"]) > 0)


def test_max_points_cut_and_empty(ctx):
    sc = synth_new_points_scene(4, n_kp=600, n_nb=3)
    ref = oracle_py.new_points_py(sc, max_points=50)
    got = ctx.new_points(sc, max_points=50)
    assert len(got["kpt"]) == 50
    unsure = np.concatenate([ref["matches"][f]["trainIdx"][ref["margin"][f] < 1.0] for f in range(3)])
    if not np.isin(np.concatenate([got["kpt"], ref["kpt"]]), unsure).any():
        assert np.array_equal(got["kpt"], ref["kpt"]) and np.array_equal(got["dist"], ref["dist"])
    assert np.all(np.diff(got["dist"]) >= 0)
    sc["t_map"] = np.zeros(0, np.int32)
    got = ctx.new_points(sc)
    assert len(got["kpt"]) == 0 and got["obs_ptr"].tolist() == [0]
