"""Pins oracle/kfdb_oracle.cpp (keyframe database query) to the reference: golden candidate lists produced by the
reference's own keyframedatabase.cpp + covisgraph.cpp + fbow (tests/golden/make_golden.py kfdb), a live comparison where oracle/_ref
exists, and the product's host-only ranking step (uco_b200_kfdb_rank, no device needed) against the same lists."""
import os
import numpy as np
import pytest
import oracle_py

GOLD = os.path.join(os.path.dirname(__file__), "golden", "kfdb_ref.npz")
CASES = ["s", "t", "orb"]


def load_case(g, name):
    ids, off, words, weights = g[name + "_ids"], g[name + "_off"], g[name + "_words"], g[name + "_weights"]
    bows = [(words[off[i]:off[i + 1]], weights[off[i]:off[i + 1]]) for i in range(len(ids))]
    edges = (g[name + "_ea"], g[name + "_eb"], g[name + "_ew"])
    deleted = set(int(x) for x in g[name + "_deleted"])
    queries = []
    for q in range(int(g[name + "_nq"])):
        k = "%s_q%d" % (name, q)
        sorted_, ms, phase = g[k + "_prm"]
        queries.append(dict(bow=(g[k + "_words"], g[k + "_weights"]), sorted=bool(sorted_), min_score=float(ms), phase=int(phase),
                            excluded=[int(x) for x in g[k + "_exc"]], cand=g[k + "_cand"]))
    return ids, bows, edges, deleted, queries


def db_at_phase(ids, bows, deleted, phase):
    keep = [i for i in range(len(ids)) if phase == 0 or int(ids[i]) not in deleted]
    return ids[keep], [bows[i] for i in keep]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    ids, bows, edges, deleted, queries = load_case(np.load(GOLD), name)
    n_nonempty = 0
    for q in queries:
        fid, fb = db_at_phase(ids, bows, deleted, q["phase"])
        r = oracle_py.kfdb_candidates(fid, fb, q["bow"], q["excluded"], q["min_score"], q["sorted"], edges)
        assert np.array_equal(r["candidates"], q["cand"])
        n_nonempty += len(q["cand"]) > 1
    assert n_nonempty >= 8   # the fixtures exercise the covisibility / sort path, not only the trivial exits


@pytest.mark.parametrize("name", CASES)
def test_pair_scores_match_reference_golden(name):
    g = np.load(GOLD)
    ids, bows, _, _, _ = load_case(g, name)
    by_id = {int(i): b for i, b in zip(ids, bows)}
    for (a, b), want in zip(g[name + "_pairs"], g[name + "_pair_score"]):
        got = np.float32(oracle_py.bow_score(*by_id[int(a)], *by_id[int(b)]))   # KeyFrameDataBase::score returns float
        assert got.view(np.uint32) == want.view(np.uint32)
    if name != "orb":
        assert len(set(g[name + "_pair_score"].tolist())) > 3   # not saturated


def test_score_matches_live_fbow():
    lib = oracle_py.load_ref("libref_fbow.so")
    if lib is None:
        pytest.skip("oracle/_ref not built")
    import ctypes
    lib.ref_fbow_score.restype = ctypes.c_double
    rng = np.random.default_rng(5)
    for t in range(50):
        n1, n2 = int(rng.integers(0, 400)), int(rng.integers(1, 400))
        a = np.sort(rng.choice(1000, n1, replace=False)).astype(np.uint32)
        b = np.sort(rng.choice(1000, n2, replace=False)).astype(np.uint32)
        scale = 10.0 ** rng.uniform(-3, 0)
        wa = (rng.random(n1) * scale).astype(np.float32); wb = (rng.random(n2) * scale).astype(np.float32)
        want = lib.ref_fbow_score(oracle_py._p(a), oracle_py._p(wa), n1, oracle_py._p(b), oracle_py._p(wb), n2)
        assert oracle_py.bow_score(a, wa, b, wb) == want


def test_oracle_matches_live_reference(tmp_path):
    if oracle_py.load_ref("libref_kfdb.so") is None:
        pytest.skip("oracle/_ref not built")
    path = str(tmp_path / "v.fbow")
    oracle_py.synth_vocabulary(31, k=9, depth=4, weight_scale=0.01).tofile(path)
    ref = oracle_py.RefKeyFrameDataBase(path)
    frames, place = oracle_py.synth_places(41, n_places=8, views_per_place=6, n_desc=350, replace_frac=0.4)
    ids = np.random.default_rng(3).permutation(len(frames)).astype(np.uint32) + 5
    bows = [ref.add(i, d) for i, d in zip(ids, frames)]
    edges = oracle_py.synth_covis(7, ids, place, extra=0.3)
    for a, b, w in zip(*edges):
        ref.covis_edge(a, b, w)
    qs, _ = oracle_py.synth_places(43, n_places=8, views_per_place=1, n_desc=350, replace_frac=0.4)
    for qd in qs:
        for sorted_ in (True, False):
            for ms in (0.0, 0.02):
                cand, qb = ref.query(qd, sorted_, ms, [int(ids[0])])
                r = oracle_py.kfdb_candidates(ids, bows, qb, [int(ids[0])], ms, sorted_, edges)
                assert np.array_equal(cand, r["candidates"])
    ref.close()


@pytest.mark.parametrize("name", CASES)
def test_host_rank_step_matches_reference_golden(name):
    """uco_b200_kfdb_rank (product, host arithmetic only) fed with the oracle's scored frames reproduces the reference's lists."""
    import ucoslam_b200
    ids, bows, edges, deleted, queries = load_case(np.load(GOLD), name)
    for q in queries:
        fid, fb = db_at_phase(ids, bows, deleted, q["phase"])
        r = oracle_py.kfdb_candidates(fid, fb, q["bow"], q["excluded"], q["min_score"], q["sorted"], edges)
        got = ucoslam_b200.rank_candidates(r["scored_frame"], r["scored_score"], lambda f: oracle_py.covis_neighbors(edges, f),
                                           q["sorted"], q["min_score"])
        assert np.array_equal(got, q["cand"])


def test_host_rank_step_fuzz():
    """uco_b200_kfdb_rank against a direct restatement of keyframedatabase.cpp:236-275 on random scored frames / neighbour lists
    (distinct accumulated scores, so the order does not depend on std::sort's treatment of ties)"""
    import ucoslam_b200
    rng = np.random.default_rng(8)
    for trial in range(200):
        n = int(rng.integers(0, 40))
        frame = np.sort(rng.choice(500, n, replace=False)).astype(np.uint32)
        score = rng.uniform(0.01, 1.0, n)
        nbrs = {int(f): rng.permutation(rng.choice(500, int(rng.integers(0, 25)), replace=False)).astype(np.uint32) for f in frame}
        sorted_, ms = bool(trial % 2), float(rng.choice([0.0, 0.3]))
        got = ucoslam_b200.rank_candidates(frame, score, lambda f: nbrs[f], sorted_, ms)
        # restatement
        if n == 0:
            want = []
        elif n == 1:
            want = [int(frame[0])]
        else:
            sc = {int(f): float(s) for f, s in zip(frame, score)}
            acc, best = [], np.float64(np.float32(ms))
            for f in frame:
                a = sc[int(f)]
                for nb in nbrs[int(f)][:10]:
                    if int(nb) in sc:
                        a += sc[int(nb)]
                acc.append((int(f), a))
                best = max(best, a)
            keep = [(f, a) for f, a in acc if not a < np.float64(np.float32(0.75)) * best]
            if sorted_:
                keep.sort(key=lambda fa: -fa[1])
            want = [f for f, _ in keep]
        assert got.tolist() == want


def test_edge_cases_against_live_reference():
    """empty database, disjoint words, a single scored frame (returned whatever its score), a minScore above every score,
    every frame excluded -- the reference's early exits (keyframedatabase.cpp:221,236,237)"""
    if oracle_py.load_ref("libref_kfdb.so") is None or not os.path.exists(oracle_py.REF_VOC_PATH):
        pytest.skip("oracle/_ref not built")
    ref = oracle_py.RefKeyFrameDataBase(oracle_py.REF_VOC_PATH)
    q = (np.array([5, 9, 100, 2000], np.uint32), np.array([0.1, 0.2, 0.05, 0.3], np.float32))
    bows, ids = [], []

    def both(excluded=(), min_score=0.0):
        a = ref.query_bow(q[0], q[1], True, min_score, excluded)
        b = oracle_py.kfdb_candidates(np.array(ids, np.uint32), bows, q, excluded, min_score, True, None)["candidates"]
        assert np.array_equal(a, b)
        return a

    assert len(both()) == 0                                             # empty database
    for i, (w, f) in enumerate([([1, 2, 3], [0.5, 0.5, 0.5]), ([7, 8], [0.1, 0.1])]):
        ids.append(10 + i); bows.append((np.array(w, np.uint32), np.array(f, np.float32))); ref.add_bow(ids[-1], *bows[-1])
    assert len(both()) == 0                                             # no common word
    ids.append(20); bows.append((np.array([5, 9, 77], np.uint32), np.array([0.2, 0.1, 0.4], np.float32))); ref.add_bow(20, *bows[-1])
    assert both().tolist() == [20]                                      # one scored frame
    assert len(both(min_score=0.9)) == 0                                # nothing above minScore
    ids.append(21); bows.append((np.array([5, 9, 100], np.uint32), np.array([0.3, 0.3, 0.3], np.float32))); ref.add_bow(21, *bows[-1])
    ids.append(22); bows.append((np.array([9, 100, 2000], np.uint32), np.array([0.01, 0.01, 0.01], np.float32))); ref.add_bow(22, *bows[-1])
    r = both()
    assert 21 in r.tolist()
    assert both(excluded=[21]).tolist() != r.tolist()
    assert len(both(excluded=[20, 21, 22, 10, 11])) == 0                # everything excluded
    ref.close()
