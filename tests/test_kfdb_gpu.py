"""Parity of the device-resident keyframe database (kfdb.cu, through the C ABI) with the oracle and with the reference's golden
candidate lists: scored frames, vote counts and scores bit-exact (doubles), candidate lists identical."""
import os
import numpy as np
import pytest
import oracle_py
import ucoslam_b200
from test_kfdb_oracle import GOLD, CASES, load_case, db_at_phase

pytestmark = pytest.mark.gpu


def check_query(db, fid, fb, q_bow, excluded, min_score, edges):
    want = oracle_py.kfdb_candidates(fid, fb, q_bow, excluded, min_score, True, edges)
    got = db.query(q_bow[0], q_bow[1], excluded, min_score)
    assert got["max_common"] == want["max_common"]
    assert np.array_equal(got["frame"], want["scored_frame"])
    assert np.array_equal(got["common"], want["scored_common"])
    assert np.array_equal(got["score"].view(np.uint64), want["scored_score"].view(np.uint64))
    return got


@pytest.mark.parametrize("name", CASES)
def test_matches_reference_golden(ctx, name):
    ids, bows, edges, deleted, queries = load_case(np.load(GOLD), name)
    db = ucoslam_b200.KeyFrameDataBase(ctx)
    for i, b in zip(ids, bows):          # one by one, like KeyFrameDataBase::add at keyframe insertion
        db.add(i, *b)
    assert db.size() == (len(ids), sum(len(b[0]) for b in bows))
    phase = 0
    for q in queries:
        if q["phase"] == 1 and phase == 0:
            for d in deleted:
                db.delete(d)
                assert not db.is_id(d)
            phase = 1
        fid, fb = db_at_phase(ids, bows, deleted, q["phase"])
        assert db.size()[0] == len(fid)
        check_query(db, fid, fb, q["bow"], q["excluded"], q["min_score"], edges)
        cand = db.relocalization_candidates(q["bow"][0], q["bow"][1], lambda f: oracle_py.covis_neighbors(edges, f), q["sorted"],
                                            q["min_score"], q["excluded"])
        assert np.array_equal(cand, q["cand"])
    db.close()


def random_bows(rng, n_frames, n_words, vocab, n_places=40, scale=0.01):
    """frames that share most words with one of n_places word sets"""
    places = [np.sort(rng.choice(vocab, n_words, replace=False)) for _ in range(n_places)]
    bows = []
    for f in range(n_frames):
        base = places[int(rng.integers(n_places))]
        keep = base[rng.random(len(base)) < rng.uniform(0.3, 0.9)]
        extra = rng.choice(vocab, int(rng.integers(0, n_words // 2)), replace=False)
        w = np.unique(np.concatenate([keep, extra])).astype(np.uint32)
        bows.append((w, (rng.random(len(w)) * scale).astype(np.float32)))
    return bows, places


@pytest.mark.parametrize("vocab,n_frames", [(1_000_000, 3000), (6_000_000, 1200)])   # shared-memory bitmap / global bitmap scan
def test_large_database_batch_add_and_edge_cases(ctx, vocab, n_frames):
    rng = np.random.default_rng(17)
    bows, places = random_bows(rng, n_frames, 600, vocab)
    bows[5] = (np.zeros(0, np.uint32), np.zeros(0, np.float32))                    # a frame without words
    bows[6] = (np.array([0, vocab - 1], np.uint32), np.array([0.5, 0.25], np.float32))
    ids = (rng.permutation(len(bows)) * 2 + 1).astype(np.uint32)
    db = ucoslam_b200.KeyFrameDataBase(ctx)
    db.add_batch(ids[:n_frames // 2], bows[:n_frames // 2])
    db.add_batch(ids[n_frames // 2:], bows[n_frames // 2:])
    assert db.size()[0] == n_frames
    empty = (np.zeros(0, np.uint32), np.zeros(0, np.float32))
    assert len(db.query(*empty)["frame"]) == 0
    for t in range(6):
        base = places[t]
        qw = np.unique(np.concatenate([base[rng.random(len(base)) < 0.8], rng.choice(vocab, 200, replace=False)])).astype(np.uint32)
        if t == 5:
            qw = np.unique(np.concatenate([qw, [0, vocab - 1]])).astype(np.uint32)
        q = (qw, (rng.random(len(qw)) * 0.01).astype(np.float32))
        exc = [int(x) for x in ids[rng.integers(0, len(ids), 20)]] + [4_000_000_000]
        got = check_query(db, ids, bows, q, exc if t % 2 else [], 0.0 if t < 3 else 0.001, None)
        assert len(got["frame"]) > 3
    # delete two thirds (forces the arena compaction), then the remaining database must answer like a fresh one
    dead = set(int(x) for x in ids[rng.random(len(ids)) < 0.67])
    for d in dead:
        db.delete(d)
    keep = [i for i in range(len(ids)) if int(ids[i]) not in dead]
    kid, kb = ids[keep], [bows[i] for i in keep]
    assert db.size() == (len(kid), sum(len(b[0]) for b in kb))
    q = (places[1].astype(np.uint32), (rng.random(len(places[1])) * 0.01).astype(np.float32))
    check_query(db, kid, kb, q, [], 0.0, None)
    db.add(ids[keep[0]] + 1, *bows[0])                      # ids are odd: +1 is new
    check_query(db, np.append(kid, ids[keep[0]] + 1).astype(np.uint32), kb + [bows[0]], q, [], 0.0, None)
    db.clear()
    assert db.size() == (0, 0) and len(db.query(*q)["frame"]) == 0
    db.close()


def test_saturated_scores_and_errors(ctx):
    db = ucoslam_b200.KeyFrameDataBase(ctx)
    w = np.arange(10, dtype=np.uint32)
    db.add(1, w, np.full(10, 3.0, np.float32))             # sum of products >= 1 -> score 1.0 (fbow.cpp:237)
    db.add(2, w[:3], np.full(3, 0.1, np.float32))
    r = db.query(w, np.full(10, 2.0, np.float32))
    assert r["frame"].tolist() == [1] and r["score"][0] == 1.0 and r["max_common"] == 10
    with pytest.raises(ucoslam_b200.UcoError):
        db.add(1, w, np.ones(10, np.float32))              # already there
    with pytest.raises(ucoslam_b200.UcoError):
        db.add(3, w[::-1].copy(), np.ones(10, np.float32))   # not ascending
    with pytest.raises(ucoslam_b200.UcoError):
        db.delete(99)
    db.close()
