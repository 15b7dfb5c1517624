"""triangulate.cu through the C ABI against the cv2-backed restatement of ucoslam::Triangulate.  Stated tolerance (floating point):
accepted points agree to 2e-5 relative (the reference's own float SVD differs from an exact null vector by up to ~3e-6 on these
scenes), accept / reject decisions are identical for every match whose closest gate is not within rounding of its threshold."""
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from ucoslam_b200.synth import synth_two_view
from test_triangulate_oracle import CASES

pytestmark = pytest.mark.gpu
REL_TOL = 2e-5


@pytest.mark.parametrize("name", list(CASES))
def test_matches_restatement(ctx, name):
    sc = synth_two_view(**CASES[name])
    want, good, margin = oracle_py.triangulate_py(sc)
    got, n = ctx.triangulate(sc)
    a, b = ~np.isnan(want[:, 0]), ~np.isnan(got[:, 0])
    clear = margin >= 1.0
    assert clear.mean() > 0.95
    assert np.array_equal(a[clear], b[clear])
    assert abs(n - good) <= (~clear).sum() and n == b.sum()
    both = a & b
    rel = np.linalg.norm(got[both].astype(np.float64) - want[both], axis=1) / np.linalg.norm(want[both].astype(np.float64), axis=1)
    assert rel.max() < REL_TOL
    assert np.isnan(got[~b]).all()


def test_scale_consistency_and_global_frame(ctx):
    """the mapper's follow-up to Triangulate (new-map-point creation): distance-ratio vs octave-ratio gate, points in global coordinates"""
    from ucoslam_b200.synth import _rodrigues
    sc = synth_two_view(1)
    G = np.eye(4); G[:3, :3] = _rodrigues(np.array([0.3, -0.2, 0.1])); G[:3, 3] = [1.0, -2.0, 0.5]
    plain, n_plain = ctx.triangulate(sc)
    for factor in (1.8, 1.15):                      # 1.5 * scaleFactor = 1.8 for the default pyramid; a tight one that rejects
        want, good, margin = oracle_py.triangulate_py(sc, 5.998, factor, G)
        got, n = ctx.triangulate(sc, 5.998, factor, G)
        a, b = ~np.isnan(want[:, 0]), ~np.isnan(got[:, 0])
        clear = margin >= 1.0
        assert clear.mean() > 0.95 and np.array_equal(a[clear], b[clear]) and n == b.sum()
        both = a & b
        rel = np.linalg.norm(got[both].astype(np.float64) - want[both], axis=1) / np.linalg.norm(want[both].astype(np.float64), axis=1)
        assert rel.max() < REL_TOL
        assert n <= n_plain
    assert n < 0.8 * n_plain                        # the tight factor does reject
    loose, n_loose = ctx.triangulate(sc, 5.998, 1.8, G)
    back = (loose[~np.isnan(loose[:, 0])].astype(np.float64) - G[:3, 3]) @ G[:3, :3]
    assert np.abs(back - plain[~np.isnan(loose[:, 0])]).max() < 1e-4


def test_edges(ctx):
    sc = synth_two_view(7, n=64)
    empty = dict(sc, matches=sc["matches"][:0])
    xyz, n = ctx.triangulate(empty)
    assert n == 0 and xyz.shape == (0, 3)
    ident = dict(sc, RT=np.eye(4, dtype=np.float32), kps_query=sc["kps_train"].copy())
    ident["matches"] = sc["matches"].copy(); ident["matches"]["queryIdx"] = ident["matches"]["trainIdx"]
    xyz, n = ctx.triangulate(ident)
    assert n == 0 and np.isnan(xyz).all()
    bad = dict(sc, matches=sc["matches"].copy()); bad["matches"]["queryIdx"][3] = 10_000
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.triangulate(bad)
    # a looser gate accepts a superset
    tight, n1 = ctx.triangulate(sc, 2.0)
    loose, n2 = ctx.triangulate(sc, 50.0)
    assert n2 >= n1 and not (np.isnan(loose[:, 0]) & ~np.isnan(tight[:, 0])).any()
