"""Parity of the sm_100a bag-of-words transform (through the C ABI) with the oracle and with the reference's fbow."""
import os
import numpy as np
import pytest
import oracle_py, ucoslam_b200

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bow_fbow.npz")
VOCS = {"s1": dict(seed=1), "s2": dict(seed=2, k=7, depth=3), "s5": dict(seed=5, k=16, depth=3, leaf_prob=0.3),
        "deep": dict(seed=21, k=10, depth=6, leaf_prob=0.02), "k20": dict(seed=8, k=20, depth=3), "k2": dict(seed=9, k=2, depth=9)}


@pytest.mark.parametrize("name", list(VOCS))
def test_gpu_matches_oracle_per_descriptor(ctx, name):
    voc_bytes = oracle_py.synth_vocabulary(**VOCS[name])
    voc = ctx.bow_load(voc_bytes)
    rng = np.random.default_rng(5)
    for n in (1, 15, 16, 17, 2000):
        desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        for level in (0, 2, 3, 8):
            got = ctx.bow_transform(voc, desc, level)
            ref = oracle_py.bow_transform(voc_bytes, desc, level)
            for a, b in zip(got, ref):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ctx.bow_free(voc)


@pytest.mark.parametrize("name", ["orb", "s1", "s2", "s5"])
def test_gpu_folded_matches_reference_golden(ctx, name):
    g = np.load(GOLD)
    if name == "orb":
        voc_bytes = oracle_py.ref_voc_bytes()
        if voc_bytes is None:
            pytest.skip("oracle/_ref/orb.fbow not present")
    else:
        voc_bytes = oracle_py.synth_vocabulary(**VOCS[name])
    voc = ctx.bow_load(voc_bytes)
    for level in (0, 3, 7):
        got = oracle_py.fold_bow(*ctx.bow_transform(voc, g["desc"], level))
        for k, v in zip(("ids", "w", "n2", "f2"), got):
            ref = g["%s_L%d_%s" % (name, level, k)]
            assert np.array_equal(v.view(np.uint32), ref.view(np.uint32)), (name, level, k)
    ctx.bow_free(voc)


def test_gpu_strided_rows_and_errors(ctx):
    voc_bytes = oracle_py.synth_vocabulary(1)
    voc = ctx.bow_load(voc_bytes)
    rng = np.random.default_rng(1)
    wide = rng.integers(0, 256, (300, 48), dtype=np.uint8)
    got = ctx.bow_transform(voc, wide[:, :32], 3)
    ref = oracle_py.bow_transform(voc_bytes, np.ascontiguousarray(wide[:, :32]), 3)
    assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(got, ref))
    with pytest.raises(ucoslam_b200.UcoError, match="No input data"):          # fbow.cpp:52
        ctx.bow_transform(voc, wide[:0, :32], 3)
    ctx.bow_free(voc)
    bad = voc_bytes.copy(); bad[0] ^= 1
    with pytest.raises(ucoslam_b200.UcoError, match="invalid signature"):      # fbow.cpp:183
        ctx.bow_load(bad)
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.bow_load(voc_bytes[:1000])
