"""CPU checks of oracle/match_oracle.c (the restatement of FrameMatcher_Flann::matchEpipolar's post-filters on exact k-NN).
The reference cannot be compiled for this row (OpenCV C++ + Frame/Map classes) and holds no golden vectors for it: PARITY
UNPINNED.  What can be checked without it: the invariants the reference's code guarantees (framematcher.cpp:228-322,
misc.cpp:153-185) and an independent pure-Python restatement on a small case."""
import numpy as np
import oracle_py


def py_frame_match(q_desc, q_kps, t_desc, t_kps, min_dd, ratio, check_or, max_oct):
    idx, dist = oracle_py.hamming_knn(q_desc, t_desc, 10, 0)
    matches = []
    for i in range(len(q_desc)):
        best, best2, bt, o2 = np.float32(min_dd), np.float32(np.finfo(np.float32).max), -1, -1
        for j in range(10):
            if idx[i, j] < 0:
                continue
            d = np.float32(dist[i, j])
            if d > np.float32(min_dd):
                continue
            if d < best2:
                t = idx[i, j]
                if abs(int(t_kps["octave"][t]) - int(q_kps["octave"][i])) > max_oct:
                    continue
                if d < best:
                    best, bt = d, t
                else:
                    best2, o2 = d, int(t_kps["octave"][t])
        if bt != -1 and not (o2 == int(q_kps["octave"][i]) and best > np.float32(best2 * np.float32(ratio))):
            matches.append([i, bt, float(best)])
    used = {}
    for k, m in enumerate(matches):
        if m[1] not in used:
            used[m[1]] = k
        elif matches[used[m[1]]][2] > m[2]:
            matches[used[m[1]]][1] = -1
            used[m[1]] = k
        else:
            m[1] = -1
    matches = [m for m in matches if m[1] != -1]
    if check_or:
        bins = []
        for m in matches:
            rot = np.float32(t_kps["angle"][m[1]]) - np.float32(q_kps["angle"][m[0]])
            if rot < 0:
                rot = np.float32(rot + np.float32(360))
            v = float(np.float32(rot * (np.float32(1.0) / np.float32(30))))
            b = int(np.floor(v + 0.5))   # roundf: half away from zero (v >= 0)
            bins.append(0 if b == 30 else b)
        cnt = np.bincount(bins, minlength=30)
        m1 = m2 = m3 = 0
        i1 = i2 = i3 = -1
        for i in range(30):
            s = cnt[i]
            if s > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
            elif s > m2:
                m3, m2, i3, i2 = m2, s, i2, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < np.float32(0.1) * np.float32(m1):
            i2 = i3 = -1
        elif m3 < np.float32(0.1) * np.float32(m1):
            i3 = -1
        matches = [m for m, b in zip(matches, bins) if b in (i1, i2, i3)]
    return matches


def test_oracle_against_python_restatement():
    for seed, kw in ((1, dict(min_dd=50.0, ratio=0.8, check_or=True, max_oct=1)), (2, dict(min_dd=80.0, ratio=0.6, check_or=False, max_oct=0))):
        q, qk, t, tk = oracle_py.synth_match_frames(seed, nt=300, nq=250)
        t[220:] = t[:80]                                     # duplicated train rows: second-best ties, ratio test fires
        tk["octave"][220:260] = tk["octave"][:40]
        got = oracle_py.frame_match(q, qk, t, tk, kw["min_dd"], kw["ratio"], kw["check_or"], kw["max_oct"])
        ref = py_frame_match(q, qk, t, tk, **kw)
        assert len(got) == len(ref) and len(ref) > 20
        assert [(int(m["queryIdx"]), int(m["trainIdx"]), float(m["distance"])) for m in got] == [tuple(m) for m in ref]


def test_invariants():
    q, qk, t, tk = oracle_py.synth_match_frames(3)
    m = oracle_py.frame_match(q, qk, t, tk)
    assert len(m) > 500
    assert len(np.unique(m["trainIdx"])) == len(m)               # filter_ambiguous_train
    assert np.all(np.diff(m["queryIdx"]) > 0)                    # query order kept by the stable removals
    assert np.all(m["distance"] < 50) and np.all(m["imgIdx"] == -1)
    assert np.all(np.abs(tk["octave"][m["trainIdx"]] - qk["octave"][m["queryIdx"]]) <= 1)
    rot = tk["angle"][m["trainIdx"]] - qk["angle"][m["queryIdx"]]
    rot = np.where(rot < 0, rot + np.float32(360), rot)
    assert len(np.unique(np.floor(rot * (np.float32(1) / np.float32(30)) + 0.5))) <= 3


def test_empty_inputs():
    q, qk, t, tk = oracle_py.synth_match_frames(4, nt=50, nq=20)
    assert len(oracle_py.frame_match(q[:0], qk, t, tk)) == 0
    assert len(oracle_py.frame_match(q, qk, t[:0], tk)) == 0


def _bow_frames(seed, nt, nq, bits=9):
    """two frames whose keypoints are filed under synthetic level-3 nodes: node id = the first `bits` bits of the descriptor, so
    re-observed keypoints usually (not always: flipped bits) share a node, as with a real vocabulary"""
    import ucoslam_b200
    q, qk, t, tk = oracle_py.synth_match_frames(seed, nt=nt, nq=nq)
    node = lambda d: ((d[:, 0].astype(np.uint32) << 8 | d[:, 1]) >> (16 - bits)).astype(np.uint32) * 16 + 3
    return q, qk, ucoslam_b200.bow_index(node(q)), t, tk, ucoslam_b200.bow_index(node(t))


def test_bow_matcher_oracle_equals_python_restatement():
    """oracle_frame_match_bow (C) against an independent pure-Python restatement of FrameMatcher_BoW::matchEpipolar"""
    for seed, nt, nq, kw in ((1, 700, 700, {}), (2, 500, 650, dict(ratio=0.6, max_octave_diff=0)), (3, 600, 400, dict(check_orientation=False, min_desc_dist=80.0))):
        q, qk, qb, t, tk, tb = _bow_frames(seed, nt, nq, bits=6)
        rng = np.random.default_rng(seed)
        qu, tu = (rng.random(nq) > 0.1).astype(np.uint8), (rng.random(nt) > 0.1).astype(np.uint8)
        a = oracle_py.frame_match_bow(q, qk, qb, t, tk, tb, q_usable=qu, t_usable=tu, **kw)
        b = oracle_py.frame_match_bow_py(q, qk, qb, t, tk, tb, q_usable=qu, t_usable=tu, **kw)
        assert len(a) > 20 and len(a) == len(b)
        for k in ("queryIdx", "trainIdx", "imgIdx", "distance"):
            assert np.array_equal(a[k], b[k]), k
