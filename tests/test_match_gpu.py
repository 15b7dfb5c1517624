"""Parity of the sm_100a frame matcher (uco_b200_frame_match{,_batch_dev}: exact k-NN + FrameMatcher_Flann's post-filters)
against oracle/match_oracle.c on seeded frames: bit-exact cv::DMatch records (indices, order, distances)."""
import numpy as np
import pytest
import torch
import oracle_py
import ucoslam_b200

pytestmark = pytest.mark.gpu


def same(a, b):
    return len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in ("queryIdx", "trainIdx", "imgIdx", "distance"))


F = np.array([[0, -1e-6, 2e-4], [1e-6, 0, -3e-3], [-2e-4, 3.1e-3, 0.01]], np.float32)


@pytest.mark.parametrize("seed,nt,nq,kw", [
    (1, 2000, 2000, dict()),                                                     # config 2: two 2000-keypoint frames
    (2, 2000, 1777, dict(check_orientation=False)),
    (3, 1500, 2000, dict(ratio=0.6, max_octave_diff=0)),                         # the tracker's ratio
    (4, 4000, 4000, dict(min_desc_dist=100.0)),                                  # config 3 size
    (5, 2000, 2000, dict(F12=F)),                                                # epipolar gate
    (6, 7, 40, dict(min_desc_dist=300.0)),                                       # fewer train rows than k
    (7, 1200, 3000, dict(min_desc_dist=60.0, check_orientation=True, max_octave_diff=7)),
])
def test_match_equals_oracle(ctx, seed, nt, nq, kw):
    q, qk, t, tk = oracle_py.synth_match_frames(seed, nt=nt, nq=nq)
    if seed in (3, 7):
        t[nt // 2:] = t[:nt - nt // 2]                                           # duplicated rows: ties + ratio test
    ref = oracle_py.frame_match(q, qk, t, tk, **kw)
    prm = ucoslam_b200.MatchParams(kw.get("min_desc_dist", 50.0), kw.get("ratio", 0.8), kw.get("check_orientation", True),
                                   kw.get("max_octave_diff", 1), kw.get("F12"))
    got = ctx.frame_match(q, qk, t, tk, prm)
    assert same(got, ref)
    assert len(ref) > 0


def test_match_with_index_maps(ctx):
    """MODE_ASSIGNED-style subsets: descriptor row i belongs to keypoint map[i] (framematcher.cpp:160-198)"""
    q, qk, t, tk = oracle_py.synth_match_frames(11, nt=1500, nq=1500)
    rng = np.random.default_rng(5)
    qm = np.sort(rng.choice(1500, 900, replace=False)).astype(np.int32)
    tm = np.sort(rng.choice(1500, 1100, replace=False)).astype(np.int32)
    ref = oracle_py.frame_match(q[qm], qk, t[tm], tk, q_map=qm, t_map=tm)
    got = ctx.frame_match(q[qm], qk, t[tm], tk, ucoslam_b200.MatchParams(), q_map=qm, t_map=tm)
    assert same(got, ref) and len(ref) > 100


def test_match_empty(ctx):
    q, qk, t, tk = oracle_py.synth_match_frames(4, nt=50, nq=20)
    assert len(ctx.frame_match(q[:0], qk, t, tk, ucoslam_b200.MatchParams())) == 0
    assert len(ctx.frame_match(q, qk, t[:0], tk, ucoslam_b200.MatchParams())) == 0


def test_match_batch_dev_equals_oracle(ctx):
    """a clip: frame p+1 against frame p for every p, one launch pair, ragged keypoint counts from device arrays"""
    n, cap = 6, 800
    frames = [oracle_py.synth_match_frames(20 + i, nt=cap, nq=cap)[2:] for i in range(n)]
    counts = np.array([800, 640, 800, 1, 333, 800], np.int32)
    desc = np.zeros((n, cap, 32), np.uint8)
    kps = np.zeros((n, cap), ucoslam_b200.KP_DTYPE)
    for i, (d, k) in enumerate(frames):
        desc[i], kps[i] = d, k
        if i:  # make frame i a re-observation of frame i-1
            m = min(counts[i], counts[i - 1])
            desc[i, :m] = desc[i - 1, :m]
            desc[i, :m, 3] ^= 0x11
            kps[i]["angle"][:m] = np.mod(kps[i - 1]["angle"][:m] + 10, 360)
            kps[i]["octave"][:m] = kps[i - 1]["octave"][:m]
    d_desc = torch.from_numpy(desc).cuda()
    d_kps = torch.from_numpy(kps.view(np.uint8).reshape(n, cap * 28)).cuda()
    d_cnt = torch.from_numpy(counts).cuda()
    out = torch.zeros((n - 1) * cap * 16, dtype=torch.uint8, device="cuda")
    nout = torch.zeros(n - 1, dtype=torch.int32, device="cuda")
    prm = ucoslam_b200.MatchParams()
    torch.cuda.synchronize()
    ctx.frame_match_batch_dev(n - 1, d_desc.data_ptr() + cap * 32, cap * 32, d_kps.data_ptr() + cap * 28, cap, cap,
                              d_cnt.data_ptr() + 4, d_desc.data_ptr(), cap * 32, d_kps.data_ptr(), cap, cap, d_cnt.data_ptr(), prm,
                              out.data_ptr(), nout.data_ptr())
    ctx.sync()
    res = out.cpu().numpy().view(ucoslam_b200.MATCH_DTYPE).reshape(n - 1, cap)
    cnt = nout.cpu().numpy()
    for p in range(n - 1):
        nq, nt = counts[p + 1], counts[p]
        ref = oracle_py.frame_match(desc[p + 1, :nq], kps[p + 1], desc[p, :nt], kps[p])
        assert cnt[p] == len(ref) and same(res[p, :cnt[p]], ref), p


@pytest.mark.parametrize("seed,nt,nq,bits,kw", [
    (11, 2000, 2000, 9, dict()),
    (12, 2000, 1777, 10, dict(check_orientation=False)),
    (13, 1500, 2000, 6, dict(ratio=0.6, max_octave_diff=0)),          # few, crowded nodes: long candidate lists, many ties
    (14, 4000, 4000, 10, dict(min_desc_dist=100.0)),
    (15, 2000, 2000, 8, dict(F12=F)),
    (16, 30, 40, 2, dict(min_desc_dist=300.0)),
])
def test_bow_match_equals_oracle(ctx, seed, nt, nq, bits, kw):
    """uco_b200_frame_match_bow against the restatement of FrameMatcher_BoW::matchEpipolar, with usable masks (matcher modes)"""
    from test_match_oracle import _bow_frames
    q, qk, qb, t, tk, tb = _bow_frames(seed, nt, nq, bits)
    rng = np.random.default_rng(seed)
    qu, tu = (rng.random(nq) > 0.15).astype(np.uint8), (rng.random(nt) > 0.15).astype(np.uint8)
    prm = ucoslam_b200.MatchParams(kw.get("min_desc_dist", 50.0), kw.get("ratio", 0.8), kw.get("check_orientation", True),
                                   kw.get("max_octave_diff", 1), kw.get("F12"))
    for masks in ((None, None), (qu, tu)):
        ref = oracle_py.frame_match_bow(q, qk, qb, t, tk, tb, q_usable=masks[0], t_usable=masks[1], **kw)
        got = ctx.frame_match_bow(q, qk, qb, t, tk, tb, prm, q_usable=masks[0], t_usable=masks[1])
        assert same(got, ref)
        assert len(ref) > 0


def test_bow_match_disjoint_vocabulary_nodes(ctx):
    """no common node -> no match; empty inputs -> no match"""
    from test_match_oracle import _bow_frames
    q, qk, qb, t, tk, tb = _bow_frames(17, 300, 300, 6)
    tb2 = (tb[0] + np.uint32(1), tb[1], tb[2])
    assert len(ctx.frame_match_bow(q, qk, qb, t, tk, tb2, ucoslam_b200.MatchParams())) == 0
    assert len(oracle_py.frame_match_bow(q, qk, qb, t, tk, tb2)) == 0


def test_match_multi_equals_oracle(ctx):
    """the mapper's pattern: one keyframe (train) against several neighbours (queries), each with its own F12, ragged sizes and
    MODE_UNASSIGNED-style row maps; every list must equal the per-pair oracle"""
    rng = np.random.default_rng(9)
    _, _, t, tk = oracle_py.synth_match_frames(50, nt=1800, nq=10)
    tm = np.sort(rng.choice(1800, 1300, replace=False)).astype(np.int32)
    qs, qks, qms, f12s = [], [], [], []
    for f, nq in enumerate([1500, 900, 2000, 0, 37]):
        q, qk, _, _ = oracle_py.synth_match_frames(60 + f, nt=10, nq=max(nq, 1))
        # queries = noisy copies of train rows so that matches exist
        if nq:
            src = rng.integers(0, 1800, nq)
            q = t[src].copy()
            flips = rng.integers(0, 256, (nq, 12))
            for j in range(12):
                q[np.arange(nq), flips[:, j] >> 3] ^= (1 << (flips[:, j] & 7)).astype(np.uint8)
            qk = qk[:nq]
            qk["octave"] = tk["octave"][src]
            qk["angle"] = tk["angle"][src]
            qk["x"], qk["y"] = tk["x"][src] + rng.normal(0, 1, nq).astype(np.float32), tk["y"][src] + rng.normal(0, 1, nq).astype(np.float32)
        else:
            q, qk = q[:0], qk[:0]
        qm = None if f % 2 else np.sort(rng.choice(max(nq, 1), nq * 2 // 3, replace=False)).astype(np.int32) if nq else None
        qs.append(q if qm is None else q[qm]); qks.append(qk); qms.append(qm)
        f12s.append(F * np.float32(1 + 0.1 * f))
    for use_f in (False, True):
        prm = ucoslam_b200.MatchParams(100.0, 0.6, True, 100, F if use_f else None)
        got = ctx.frame_match_multi(t[tm], tk, qs, qks, prm, f12=np.array(f12s) if use_f else None, t_map=tm, q_maps=qms)
        tot = 0
        for f in range(len(qs)):
            ref = oracle_py.frame_match(qs[f], qks[f], t[tm], tk, min_desc_dist=100.0, ratio=0.6, check_orientation=True, max_octave_diff=100,
                                        F12=f12s[f] if use_f else None, q_map=qms[f], t_map=tm) if len(qs[f]) else np.zeros(0, ucoslam_b200.MATCH_DTYPE)
            assert same(got[f], ref)
            tot += len(ref)
        assert tot > 50


def test_keyframes_batch_on_resident_frames(ctx):
    """the mapper's per-keyframe work on the frames of the last extraction call: bags of words equal uco_b200_bow_transform of the
    downloaded descriptors, match lists equal the per-pair oracle (train = keyframe, query = neighbour)"""
    from ucoslam_b200 import workload
    rng = np.random.default_rng(3)
    tex = workload.texture(5, 1024)
    cam = workload.Camera()
    imgs = np.stack([workload.render(cam, tex, workload.gt_pose(i, 0.3)) for i in range(6)])
    prm = ucoslam_b200.OrbParams(1000)
    kps, desc, n = ctx.orb_extract_batch(list(imgs), prm)
    voc_bytes = oracle_py.synth_vocabulary(3, k=10, depth=4)
    voc = ctx.bow_load(voc_bytes)
    groups = [(5, [0, 1, 2, 3, 4]), (2, [0, 1]), (3, [])]
    mprm = ucoslam_b200.MatchParams(100.0, 0.6, True, 1 << 30)
    bows, matches = ctx.keyframes_batch(voc, groups, mprm, 1000)
    for j, (kf, nbs) in enumerate(groups):
        w, wt, nd = ctx.bow_transform(voc, desc[kf][:n[kf]], 3)
        assert np.array_equal(bows[j][0][:n[kf]], w) and np.array_equal(bows[j][1][:n[kf]], wt) and np.array_equal(bows[j][2][:n[kf]], nd)
        for e, f in enumerate(nbs):
            ref = oracle_py.frame_match(desc[f][:n[f]], kps[f][:n[f]], desc[kf][:n[kf]], kps[kf][:n[kf]], min_desc_dist=100.0, ratio=0.6,
                                        check_orientation=True, max_octave_diff=1 << 30)
            assert same(matches[j][e], ref) and len(ref) > 20
    ctx.bow_free(voc)
