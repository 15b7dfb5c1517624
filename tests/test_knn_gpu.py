"""Parity of the sm_100a Hamming k-NN kernel (through the C ABI) with the oracle: bit-exact indices and distances,
in the reference's heap order and in its sorted order."""
import os
import numpy as np
import pytest
import oracle_py

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "knn_xflann_linear.npz")


@pytest.mark.parametrize("case", list("abcde"))
@pytest.mark.parametrize("order", [0, 1])
def test_gpu_matches_reference_golden(ctx, case, order):
    g = np.load(GOLD)
    idx, dist = ctx.hamming_knn(g[case + "_q"], g[case + "_t"], int(g[case + "_k"]), order)
    assert np.array_equal(idx, g["%s_idx%d" % (case, order)])
    assert np.array_equal(dist, g["%s_dist%d" % (case, order)])


@pytest.mark.parametrize("nq,nt,k", [(2000, 2000, 10), (1, 1, 1), (17, 255, 10), (16, 256, 10), (15, 257, 3),
                                      (100, 1025, 32), (5, 3, 10), (2000, 4000, 10), (33, 5000, 7)])
@pytest.mark.parametrize("order", [0, 1])
def test_gpu_matches_oracle(ctx, nq, nt, k, order):
    t, q = oracle_py.synth_descriptors(1000 + nq + nt + k, nt, nq)
    a = ctx.hamming_knn(q, t, k, order)
    b = oracle_py.hamming_knn(q, t, k, order)
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1], b[1])


def test_gpu_strided_rows_and_empty(ctx):
    t, q = oracle_py.synth_descriptors(3, 500, 70)
    qs = np.zeros((70, 48), np.uint8); qs[:, :32] = q        # cv::Mat rows with step > cols
    ts = np.zeros((500, 64), np.uint8); ts[:, :32] = t
    a = ctx.hamming_knn(qs[:, :32], ts[:, :32], 10)
    b = oracle_py.hamming_knn(q, t, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    i, d = ctx.hamming_knn(q[:0], t, 10)
    assert i.shape == (0, 10)
    i, d = ctx.hamming_knn(q, t[:0], 4)                       # empty train set: all padding
    assert (i == -1).all() and (d == 0).all()


def test_gpu_bad_arguments(ctx):
    import ucoslam_b200
    t, q = oracle_py.synth_descriptors(3, 50, 7)
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.hamming_knn(q, t, 0)
    with pytest.raises(ucoslam_b200.UcoError):
        ctx.hamming_knn(q, t, 33)


def test_gpu_large_scan_properties(ctx):
    """Full-size property check (config 4 shape scaled to what the oracle cannot do in seconds): distances sorted,
    each (idx, dist) pair consistent, and the k-th distance equals the true k-th order statistic."""
    t, q = oracle_py.synth_descriptors(77, 200000, 256)
    idx, dist = ctx.hamming_knn(q, t, 10, 1)
    assert (np.diff(dist, axis=1) >= 0).all()
    d_chk = np.unpackbits(q[:, None, :] ^ t[idx], axis=2).sum(2)
    assert np.array_equal(d_chk, dist)
    for i in range(0, 256, 37):
        full = np.unpackbits(q[i][None, :] ^ t, axis=1).sum(1)
        assert np.array_equal(np.sort(full)[:10], dist[i])


def test_gpu_batch_pairs_with_device_counts(ctx):
    """One launch over a clip: pair p = (frame p+1 vs frame p), per-pair row counts read from device memory."""
    import torch, ucoslam_b200
    F, N, k = 5, 600, 10
    rng = np.random.default_rng(9)
    desc = rng.integers(0, 256, (F, N, 32), dtype=np.uint8)
    counts = np.array([600, 512, 77, 600, 1], np.int32)
    stream = torch.cuda.ExternalStream(ctx.stream)
    with torch.cuda.stream(stream):
        d = torch.from_numpy(desc).cuda()
        c = torch.from_numpy(counts).cuda()
        idx = torch.full((F - 1, N, k), -7, dtype=torch.int32, device="cuda")
        dist = torch.full((F - 1, N, k), -7, dtype=torch.int32, device="cuda")
        ctx.hamming_knn_batch_dev(F - 1, d[1].data_ptr(), N * 32, N, c[1:].data_ptr(), d[0].data_ptr(), N * 32, N,
                                  c.data_ptr(), k, ucoslam_b200.UCO_KNN_HEAP, idx.data_ptr(), dist.data_ptr())
        ctx.sync()
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    for p in range(F - 1):
        nq, nt = counts[p + 1], counts[p]
        oi, od = oracle_py.hamming_knn(desc[p + 1][:nq], desc[p][:nt], k, 0)
        assert np.array_equal(idx[p][:nq], oi) and np.array_equal(dist[p][:nq], od)
        assert (idx[p][nq:] == -7).all()


@pytest.mark.gpu
def test_host_batch_matches_single_calls(ctx):
    """uco_b200_hamming_knn_batch (one launch for many host-buffer pairs) == uco_b200_hamming_knn per pair, for the tracking
    chain (train of pair i is the query of pair i-1, ragged row counts) and for unrelated pairs."""
    rng = np.random.default_rng(11)
    sets = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (300, 257, 0, 300, 64, 1)]
    block = rng.integers(0, 256, (4, 300, 32), dtype=np.uint8)   # contiguous equal-sized blocks: the coalesced-copy path
    for qs, ts in (([sets[i] for i in range(1, 6)], [sets[i - 1] for i in range(1, 6)]),
                   ([sets[0], sets[3], sets[4]], [sets[1], sets[4], sets[2]]),
                   ([block[i] for i in range(4)], [block[i - 1] for i in range(4)])):
        idx, dist = ctx.hamming_knn_batch(qs, ts, 10)
        for q, t, i_b, d_b in zip(qs, ts, idx, dist):
            if len(q) == 0:
                continue
            i_s, d_s = ctx.hamming_knn(q, t.reshape(-1, 32), 10)
            assert np.array_equal(i_b, i_s) and np.array_equal(d_b, d_s)
