"""Pins oracle/knn_oracle.c to the reference: golden vectors produced by the reference's xflann (linear index), and,
where oracle/_ref exists, a live comparison on fresh seeded inputs."""
import os
import numpy as np
import pytest
import oracle_py

GOLD = os.path.join(os.path.dirname(__file__), "golden", "knn_xflann_linear.npz")


@pytest.mark.parametrize("case", list("abcde"))
@pytest.mark.parametrize("order", [0, 1])
def test_oracle_matches_reference_golden(case, order):
    g = np.load(GOLD)
    idx, dist = oracle_py.hamming_knn(g[case + "_q"], g[case + "_t"], int(g[case + "_k"]), order)
    assert np.array_equal(idx, g["%s_idx%d" % (case, order)])
    assert np.array_equal(dist, g["%s_dist%d" % (case, order)])


def test_oracle_matches_live_reference():
    if oracle_py.load_ref("libref_xflann.so") is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    for seed, nt, nq, k in [(21, 1500, 40, 10), (22, 9, 12, 10), (23, 64, 64, 32)]:
        t, q = oracle_py.synth_descriptors(seed, nt, nq)
        for order in (0, 1):
            a = oracle_py.hamming_knn(q, t, k, order)
            b = oracle_py.ref_xflann_knn(q, t, k, 0, -1, order)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_oracle_properties():
    t, q = oracle_py.synth_descriptors(5, 400, 30)
    idx, dist = oracle_py.hamming_knn(q, t, 10, 1)
    assert (np.diff(dist, axis=1) >= 0).all()
    full = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(2)
    assert np.array_equal(np.sort(full, axis=1)[:, :10], dist)
    assert np.array_equal(full[np.arange(30)[:, None], idx], dist)
