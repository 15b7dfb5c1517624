"""Pins oracle/bow_oracle.c to the reference's fbow: golden vectors produced by fbow::Vocabulary::transform (on the shipped
orb.fbow and on seeded synthetic vocabularies), and a live comparison where oracle/_ref exists."""
import os
import numpy as np
import pytest
import oracle_py

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bow_fbow.npz")
SYNTH = {"s1": dict(seed=1), "s2": dict(seed=2, k=7, depth=3), "s5": dict(seed=5, k=16, depth=3, leaf_prob=0.3)}


def _voc(name):
    if name == "orb":
        v = oracle_py.ref_voc_bytes()
        if v is None:
            pytest.skip("oracle/_ref/orb.fbow not present (copied from the reference by `make -C oracle ref`)")
        return v
    return oracle_py.synth_vocabulary(**SYNTH[name])


@pytest.mark.parametrize("name", ["orb", "s1", "s2", "s5"])
@pytest.mark.parametrize("level", [0, 3, 7])
def test_oracle_matches_reference_golden(name, level):
    g = np.load(GOLD)
    got = oracle_py.fold_bow(*oracle_py.bow_transform(_voc(name), g["desc"], level))
    for k, v in zip(("ids", "w", "n2", "f2"), got):
        ref = g["%s_L%d_%s" % (name, level, k)]
        assert np.array_equal(v.view(np.uint32), ref.view(np.uint32)), k


def test_oracle_matches_live_reference():
    if oracle_py.load_ref("libref_fbow.so") is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    desc = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    for voc in (oracle_py.synth_vocabulary(21, k=10, depth=5, leaf_prob=0.02), oracle_py.synth_vocabulary(22, k=3, depth=6)):
        R = oracle_py.RefVocabulary(voc)
        for level in (1, 3):
            a = oracle_py.fold_bow(*oracle_py.bow_transform(voc, desc, level))
            b = R.transform(desc, level)
            assert all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(a, b))
        R.close()


def test_score_matches_reference():
    lib = oracle_py.load_oracle()
    import ctypes
    lib.oracle_bow_score.restype = ctypes.c_double
    voc = oracle_py.synth_vocabulary(1)
    rng = np.random.default_rng(2)
    d1 = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    d2 = d1.copy(); d2[250:] = rng.integers(0, 256, (250, 32), dtype=np.uint8)
    a = oracle_py.fold_bow(*oracle_py.bow_transform(voc, d1, 3))
    b = oracle_py.fold_bow(*oracle_py.bow_transform(voc, d2, 3))
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    s = lib.oracle_bow_score(p(a[0]), p(a[1]), len(a[0]), p(b[0]), p(b[1]), len(b[0]))
    assert 0 <= s <= 1
    if oracle_py.load_ref("libref_fbow.so") is not None:
        R = oracle_py.RefVocabulary(voc)
        assert s == R.lib.ref_fbow_score(p(a[0]), p(a[1]), len(a[0]), p(b[0]), p(b[1]), len(b[0]))
