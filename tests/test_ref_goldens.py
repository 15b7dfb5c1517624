"""The restatement oracles and the CUDA path against golden vectors produced by the REFERENCE'S OWN code (tests/golden/make_golden.py
project / match: src/utils/framematcher.cpp compiled unchanged; Map::matchFrameToMapPoints, the tracker's search by projection and the
helpers they call compiled from the reference's statements — oracle/ref_match_wrap.cpp, oracle/ref_project_wrap.cpp).  CPU part: the
oracles reproduce the goldens bit for bit.  GPU part (-m gpu): so does the library, through the C ABI."""
import os
import numpy as np
import pytest
import oracle_py
import ucoslam_b200

G = os.path.join(os.path.dirname(__file__), "golden")


def _scene(g, name):
    sc = {k[len(name) + 1:]: g[k] for k in g.files if k.startswith(name + "_") and not k.startswith(name + "_out_")}
    for k in ("fx", "fy", "cx", "cy", "bf"):
        if k in sc:
            sc[k] = float(sc[k])
    return sc


def _match_case(g, name):
    kw = dict(min_desc_dist=float(g[name + "_min_desc_dist"]), ratio=float(g[name + "_ratio"]), check_orientation=bool(g[name + "_check_orientation"]),
              max_octave_diff=int(g[name + "_max_octave_diff"]))
    if name + "_F12" in g.files:
        kw["F12"] = g[name + "_F12"]
    return g[name + "_q"], g[name + "_qk"], g[name + "_t"], g[name + "_tk"], kw, g[name + "_out"]


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_match_projected_equals_reference(name):
    g = np.load(os.path.join(G, "project_match.npz"))
    m, vis = oracle_py.match_projected(_scene(g, name), float(g[name + "_out_min_desc"]), float(g[name + "_out_max_reproj"]))
    assert np.array_equal(m, g[name + "_out_matches"]) and np.array_equal(vis, g[name + "_out_visible"])


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_track_projected_equals_reference(name):
    g = np.load(os.path.join(G, "track_projected.npz"))
    m = oracle_py.track_projected(_scene(g, name), float(g[name + "_out_dist_thr"]), float(g[name + "_out_proj_thr"]))
    assert np.array_equal(m, g[name + "_out_matches"]) and len(m) > 100


@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e"])
def test_oracle_frame_match_equals_reference(name):
    q, qk, t, tk, kw, out = _match_case(np.load(os.path.join(G, "match_ref.npz")), name)
    assert np.array_equal(oracle_py.frame_match(q, qk, t, tk, **kw), out) and len(out) > 30


@pytest.mark.parametrize("name", ["p", "q", "r", "s"])
def test_oracle_frame_match_bow_equals_reference(name):
    g = np.load(os.path.join(G, "match_ref.npz"))
    q, qk, t, tk, kw, out = _match_case(g, name)
    qb, tb = tuple(g["%s_qb%d" % (name, i)] for i in range(3)), tuple(g["%s_tb%d" % (name, i)] for i in range(3))
    m = oracle_py.frame_match_bow(q, qk, qb, t, tk, tb, q_usable=1 - g[name + "_qflags"], t_usable=1 - g[name + "_tflags"], **kw)
    assert np.array_equal(m, out) and len(out) > 20


def test_reference_libs_agree_when_present():
    """where oracle/_ref exists (the build container), the goldens are reproduced by the reference libraries themselves"""
    if oracle_py.load_ref("libref_project.so") is None or oracle_py.load_ref("libref_match.so") is None:
        pytest.skip("oracle/_ref not built")
    g = np.load(os.path.join(G, "track_projected.npz"))
    assert np.array_equal(oracle_py.ref_track_projected(_scene(g, "a"), float(g["a_out_dist_thr"]), float(g["a_out_proj_thr"])), g["a_out_matches"])
    q, qk, t, tk, kw, out = _match_case(np.load(os.path.join(G, "match_ref.npz")), "b")
    assert np.array_equal(oracle_py.ref_frame_match(q, qk, t, tk, **kw), out)


# ---- the library (C ABI, sm_100a) against the same reference-made vectors -------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_cuda_track_projected_equals_reference(ctx, name):
    g = np.load(os.path.join(G, "track_projected.npz"))
    m = ctx.track_projected(_scene(g, name), float(g[name + "_out_dist_thr"]), float(g[name + "_out_proj_thr"]))
    assert np.array_equal(m, g[name + "_out_matches"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c", "d", "e"])
def test_cuda_frame_match_equals_reference(ctx, name):
    q, qk, t, tk, kw, out = _match_case(np.load(os.path.join(G, "match_ref.npz")), name)
    prm = ucoslam_b200.MatchParams(kw["min_desc_dist"], kw["ratio"], kw["check_orientation"], kw["max_octave_diff"], kw.get("F12"))
    assert np.array_equal(ctx.frame_match(q, qk, t, tk, prm), out)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["p", "q", "r", "s"])
def test_cuda_frame_match_bow_equals_reference(ctx, name):
    g = np.load(os.path.join(G, "match_ref.npz"))
    q, qk, t, tk, kw, out = _match_case(g, name)
    qb, tb = tuple(g["%s_qb%d" % (name, i)] for i in range(3)), tuple(g["%s_tb%d" % (name, i)] for i in range(3))
    prm = ucoslam_b200.MatchParams(kw["min_desc_dist"], kw["ratio"], kw["check_orientation"], kw["max_octave_diff"], kw.get("F12"))
    m = ctx.frame_match_bow(q, qk, qb, t, tk, tb, prm, q_usable=1 - g[name + "_qflags"], t_usable=1 - g[name + "_tflags"])
    assert np.array_equal(m, out)
