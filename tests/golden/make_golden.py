"""Generate the committed golden fixtures from the REFERENCE's own code (oracle/_ref, compiled from /root/reference).

Run in the build container only (where /root/reference exists):  python tests/golden/make_golden.py
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import oracle_py


def knn():
    oracle_py.build_ref()
    cases = {}
    for name, (seed, nt, nq, k, flips) in {
        "a": (7, 2000, 64, 10, 40),     # the tracker's shape (k=10), many exact ties
        "b": (8, 300, 33, 32, 8),       # k = max
        "c": (9, 5, 20, 10, 3),         # fewer train rows than k: -1 / 0 padding
        "d": (10, 1000, 50, 2, 0),      # duplicates: distance-0 ties
        "e": (11, 777, 40, 1, 20),
    }.items():
        t, q = oracle_py.synth_descriptors(seed, nt, nq, flips)
        if name == "d":
            t[500:] = t[:500]  # every row appears twice
        for order in (0, 1):
            idx, dist = oracle_py.ref_xflann_knn(q, t, k, 0, -1, order)
            cases["%s_idx%d" % (name, order)] = idx
            cases["%s_dist%d" % (name, order)] = dist
        cases[name + "_t"], cases[name + "_q"], cases[name + "_k"] = t, q, np.int32(k)
    np.savez_compressed(os.path.join(HERE, "knn_xflann_linear.npz"), **cases)
    print("knn golden written", {k: v.shape for k, v in cases.items() if k.endswith("idx0")})


def bow():
    """fbow::Vocabulary::transform of the reference on (a) its shipped orb.fbow and (b) seeded synthetic vocabularies."""
    oracle_py.build_ref()
    out = {}
    rng = np.random.default_rng(77)
    desc = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    out["desc"] = desc
    vocs = {"orb": oracle_py.ref_voc_bytes(), "s1": oracle_py.synth_vocabulary(1), "s2": oracle_py.synth_vocabulary(2, k=7, depth=3),
            "s5": oracle_py.synth_vocabulary(5, k=16, depth=3, leaf_prob=0.3)}
    for name, voc in vocs.items():
        R = oracle_py.RefVocabulary(voc)
        for level in (0, 3, 7):
            ids, w, n2, f2 = R.transform(desc, level)
            for k, v in zip(("ids", "w", "n2", "f2"), (ids, w, n2, f2)):
                out["%s_L%d_%s" % (name, level, k)] = v
        R.close()
    np.savez_compressed(os.path.join(HERE, "bow_fbow.npz"), **out)
    print("bow golden written", len(out))


BA_CASES = {  # name -> synth_ba_problem kwargs, n_iters   (also imported by the tests)
    "mono": (dict(seed=3, n_poses=6, n_fixed=1, n_points=160), 5),
    "stereo": (dict(seed=4, n_poses=5, n_fixed=2, n_points=120, stereo_frac=0.5), 5),
    "outliers": (dict(seed=5, n_poses=8, n_fixed=2, n_points=200, outlier_frac=0.15, stereo_frac=0.2), 10),
    "allfree": (dict(seed=6, n_poses=4, n_fixed=0, n_points=100, pose_noise=(0.03, 1.5), point_noise=0.1), 5),
}


def ba():
    """GlobalOptimizerG2O's graph + two-stage LM run by the reference's own g2o / typesg2o.h (oracle/ref_g2o_wrap.cpp)."""
    oracle_py.build_ref()
    out = {}
    for name, (kw, iters) in BA_CASES.items():
        pb = oracle_py.synth_ba_problem(**kw)
        r = oracle_py.ref_ba_optimize(pb, iters)
        for k in oracle_py.BA_INPUT_KEYS:
            out["%s_in_%s" % (name, k)] = np.asarray(pb[k])
        for k, v in r.items():
            out["%s_out_%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "ba_g2o.npz"), **out)
    print("ba golden written", {n: int(out[n + "_out_iters"].sum()) for n in BA_CASES})


BA_MARKER_CASES = {  # name -> (synth_ba_problem kwargs, add_markers kwargs, n_iters)   (also imported by the tests)
    "mk_mono": (dict(seed=61, n_poses=8, n_fixed=1, n_points=250), dict(seed=5, n_markers=3), 5),
    "mk_stereo": (dict(seed=62, n_poses=6, n_fixed=2, n_points=200, stereo_frac=0.4), dict(seed=6, n_markers=2, size=0.15), 5),
    "mk_many": (dict(seed=63, n_poses=10, n_fixed=1, n_points=150, outlier_frac=0.08), dict(seed=7, n_markers=8, corner_sigma=0.6), 10),
}
BA_MARKER_KEYS = ("marker_pose44", "marker_size", "mobs_marker", "mobs_pose", "mobs_corners", "mobs_weight")


def ba_markers():
    """the same graph with ArUco markers: the reference's g2o + its OWN MarkerEdge class (typesg2o.h:108-167), numeric Jacobians"""
    oracle_py.build_ref()
    from ucoslam_b200.synth import add_markers
    out = {}
    for name, (kw, mkw, iters) in BA_MARKER_CASES.items():
        pb = add_markers(oracle_py.synth_ba_problem(**kw), **mkw)
        r = oracle_py.ref_ba_optimize(pb, iters)
        for k in oracle_py.BA_INPUT_KEYS + BA_MARKER_KEYS:
            out["%s_in_%s" % (name, k)] = np.asarray(pb[k])
        for k, v in r.items():
            out["%s_out_%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "ba_markers_g2o.npz"), **out)
    print("ba marker golden written", {n: (int(out[n + "_out_iters"].sum()), len(out[n + "_in_mobs_marker"])) for n in BA_MARKER_CASES})


BA_CAM_CASES = {  # name -> (synth_ba_problem kwargs, add_markers kwargs or None, mix_cameras kwargs, n_iters)   (also imported by the tests)
    "cam_mono": (dict(seed=71, n_poses=8, n_fixed=1, n_points=250), None, dict(seed=3), 5),
    "cam_stereo": (dict(seed=72, n_poses=7, n_fixed=2, n_points=220, stereo_frac=0.4, outlier_frac=0.06), None, dict(seed=4, frac=0.4), 5),
    "cam_markers": (dict(seed=73, n_poses=8, n_fixed=1, n_points=250), dict(seed=5, n_markers=3), dict(seed=5), 5),
}


def ba_cams():
    """windows whose keyframes were taken with two cameras: every edge carries the ImageParams of its keyframe (globaloptimizer_g2o.cpp:
    233-236, :262-266, :335-338); the reference's g2o + its own edge classes through ref_ba_optimize_cams"""
    oracle_py.build_ref()
    from ucoslam_b200.synth import add_markers, mix_cameras
    out = {}
    for name, (kw, mkw, ckw, iters) in BA_CAM_CASES.items():
        pb = oracle_py.synth_ba_problem(**kw)
        if mkw:
            pb = add_markers(pb, **mkw)
        pb = mix_cameras(pb, **ckw)
        r = oracle_py.ref_ba_optimize(pb, iters)
        for k in oracle_py.BA_INPUT_KEYS + ("pose_cam",) + (BA_MARKER_KEYS if mkw else ()):
            out["%s_in_%s" % (name, k)] = np.asarray(pb[k])
        for k, v in r.items():
            out["%s_out_%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "ba_cams_g2o.npz"), **out)
    print("ba mixed-camera golden written", {n: int(out[n + "_out_iters"].sum()) for n in BA_CAM_CASES})


BA_PLANE_CASES = {  # name -> (synth_ba_problem kwargs, add_markers kwargs, reference marker inside the window?, n_iters)   (also imported by the tests)
    "plane_in": (dict(seed=81, n_poses=8, n_fixed=1, n_points=250), dict(seed=5, n_markers=4, coplanar=True), True, 5),
    "plane_out": (dict(seed=82, n_poses=8, n_fixed=1, n_points=250), dict(seed=6, n_markers=5, coplanar=True), False, 5),
    "plane_stereo": (dict(seed=83, n_poses=6, n_fixed=2, n_points=200, stereo_frac=0.4), dict(seed=7, n_markers=3, size=0.15, coplanar=True), True, 5),
}
BA_PLANE_KEYS = ("plane_ref", "plane_other", "plane_weight", "plane_ref_pose44")


def ba_planar():
    """the InPlaneMarkers option (globaloptimizer_g2o.cpp:356-401): MarkerEdgeX restated on the reference's g2o (ref_g2o_wrap.cpp), g2o's
    own numeric Jacobians; reference marker inside / outside the window"""
    oracle_py.build_ref()
    from ucoslam_b200.synth import add_markers, add_plane_edges
    out = {}
    for name, (kw, mkw, inw, iters) in BA_PLANE_CASES.items():
        pb = add_plane_edges(add_markers(oracle_py.synth_ba_problem(**kw), **mkw), inw)
        r = oracle_py.ref_ba_optimize(pb, iters)
        for k in oracle_py.BA_INPUT_KEYS + BA_MARKER_KEYS + BA_PLANE_KEYS:
            if k in pb:
                out["%s_in_%s" % (name, k)] = np.asarray(pb[k])
        for k, v in r.items():
            out["%s_out_%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "ba_planar_g2o.npz"), **out)
    print("ba planar golden written", {n: int(out[n + "_out_iters"].sum()) for n in BA_PLANE_CASES})


PNP_CASES = {  # name -> synth_pnp_problem kwargs   (also imported by the tests)
    "mono": dict(seed=1, n_matches=800),
    "stereo": dict(seed=2, n_matches=600, stereo_frac=0.5),
    "markers": dict(seed=3, n_matches=300, n_markers=3),
    "few": dict(seed=4, n_matches=40, outlier_frac=0.5),
    "markers_only": dict(seed=5, n_matches=0, n_markers=2),
    "hopeless": dict(seed=6, n_matches=12, outlier_frac=0.9),            # < 10 inliers after round 0: early exit (:383)
    "big": dict(seed=7, n_matches=1500, stereo_frac=0.3, n_markers=2, outlier_frac=0.2),
    "far": dict(seed=8, n_matches=500, pose_noise=(0.15, 6.0), outlier_frac=0.25),   # rejected LM trials
}


def pnp():
    """PnPSolver::solvePnp's graph + 4-round schedule run by the reference's own g2o / typesg2o.h (oracle/ref_g2o_wrap.cpp)."""
    oracle_py.build_ref()
    out = {}
    for name, kw in PNP_CASES.items():
        pb = oracle_py.synth_pnp_problem(**kw)
        r = oracle_py.ref_pose_only(pb)
        for k in oracle_py.PNP_INPUT_KEYS:
            out["%s_in_%s" % (name, k)] = np.asarray(pb[k])
        for k, v in r.items():
            out["%s_out_%s" % (name, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "pnp_g2o.npz"), **out)
    print("pnp golden written", {n: (int(out[n + "_out_n_good"]), out[n + "_out_iters"].tolist()) for n in PNP_CASES})


def project():
    """Map::matchFrameToMapPoints and the tracker's search by projection from the previous frame, BY THE REFERENCE ITSELF: its own
    statements (src/map.cpp:651-770, src/utils/system.cpp:5921-6456, the Frame / MapPoint / Se3Transform helpers they call, picoflann)
    compiled into oracle/_ref/libref_project.so (oracle/ref_project_wrap.cpp, oracle/gen_ref_extract.py).  The restatement
    oracle/project_oracle.cpp is checked against it in the same breath.  Scenes are small so the fixture stays small."""
    oracle_py.build_ref()
    from ucoslam_b200.synth import synth_projection_scene, synth_track_scene
    out = {}
    for name, (kw, thr) in {"a": (dict(seed=11, n_kp=1200, n_mp=1500), (50.0, 15.0)), "b": (dict(seed=12, n_kp=600, n_mp=900, dup_frac=0.4), (80.0, 30.0)),
                            "c": (dict(seed=13, n_kp=800, n_mp=1000, clutter=0.8), (100.0, 40.0))}.items():
        sc = synth_projection_scene(**kw)
        sc["mp_id"] = np.arange(len(sc["mp_id"]), dtype=np.uint32)      # the reference wrapper names the points by their row
        ref = oracle_py.parse_picoflann_stream(oracle_py.ref_picoflann_stream(sc["kp_xy"]))
        mine = oracle_py.kdtree_build(sc["kp_xy"])
        assert all(np.array_equal(ref[k], mine[k]) for k in ("nodes", "div", "leaf_idx"))
        m, vis = oracle_py.ref_match_projected(sc, *thr)
        m2, vis2 = oracle_py.match_projected(sc, *thr)
        assert np.array_equal(m, m2) and np.array_equal(vis, vis2), "restatement differs from the reference"
        for k, v in sc.items():
            out["%s_%s" % (name, k)] = np.asarray(v)
        out[name + "_out_matches"], out[name + "_out_visible"] = m, vis
        out[name + "_out_min_desc"], out[name + "_out_max_reproj"] = np.float32(thr[0]), np.float32(thr[1])
    np.savez_compressed(os.path.join(HERE, "project_match.npz"), **out)
    print("projection golden written (reference-compiled)", {n: len(out[n + "_out_matches"]) for n in "abc"})
    out = {}
    for name, (kw, thr) in {"a": (dict(seed=21, n_kp=1200, n_mp=1500, n_prev=1000), (75.0, 15.0)),
                            "b": (dict(seed=22, n_kp=500, n_mp=900, n_prev=800, dup_frac=0.4), (120.0, 30.0)),
                            "c": (dict(seed=23, n_kp=800, n_mp=1000, n_prev=700, clutter=0.8), (40.0, 4.0))}.items():
        sc = synth_track_scene(**kw)
        m = oracle_py.ref_track_projected(sc, *thr)
        assert np.array_equal(m, oracle_py.track_projected(sc, *thr)), "restatement differs from the reference"
        for k, v in sc.items():
            out["%s_%s" % (name, k)] = np.asarray(v)
        out[name + "_out_matches"] = m
        out[name + "_out_dist_thr"], out[name + "_out_proj_thr"] = np.float32(thr[0]), np.float32(thr[1])
    np.savez_compressed(os.path.join(HERE, "track_projected.npz"), **out)
    print("search-by-projection golden written (reference-compiled)", {n: len(out[n + "_out_matches"]) for n in "abc"})


def match():
    """FrameMatcher_Flann (on xflann's exact index, see oracle/ref_match_wrap.cpp) and FrameMatcher_BoW BY THE REFERENCE ITSELF:
    src/utils/framematcher.cpp compiled unchanged into oracle/_ref/libref_match.so with the reference's own match filters."""
    oracle_py.build_ref()
    import ucoslam_b200
    F = np.array([[0, -1e-6, 2e-4], [1e-6, 0, -3e-3], [-2e-4, 3.1e-3, 0.01]], np.float32)
    out = {}
    for name, (seed, nt, nq, kw) in {"a": (1, 1200, 1200, {}), "b": (3, 900, 1100, dict(ratio=0.6, max_octave_diff=0)),
                                     "c": (5, 1000, 1000, dict(F12=F)), "d": (7, 700, 1500, dict(min_desc_dist=60.0, max_octave_diff=7)),
                                     "e": (9, 800, 800, dict(check_orientation=False, min_desc_dist=90.0))}.items():
        q, qk, t, tk = oracle_py.synth_match_frames(seed, nt=nt, nq=nq)
        if name in "bd":
            t[nt // 2:] = t[:nt - nt // 2]                          # duplicated rows: ties + ratio test
        m = oracle_py.ref_frame_match(q, qk, t, tk, **kw)
        assert np.array_equal(m, oracle_py.frame_match(q, qk, t, tk, **kw)), "restatement differs from the reference"
        out.update({name + "_q": q, name + "_qk": qk, name + "_t": t, name + "_tk": tk, name + "_out": m})
        for k, v in dict(min_desc_dist=50.0, ratio=0.8, check_orientation=True, max_octave_diff=1).items():
            out["%s_%s" % (name, k)] = np.float32(kw.get(k, v))
        if "F12" in kw:
            out[name + "_F12"] = kw["F12"]
    node = lambda d, bits: ((d[:, 0].astype(np.uint32) << 8 | d[:, 1]) >> (16 - bits)).astype(np.uint32) * 16 + 3
    for name, (seed, nt, nq, kw) in {"p": (1, 700, 700, {}), "q": (2, 500, 650, dict(ratio=0.6, max_octave_diff=0)),
                                     "r": (3, 600, 400, dict(check_orientation=False, min_desc_dist=80.0)), "s": (4, 600, 600, dict(F12=F))}.items():
        q, qk, t, tk = oracle_py.synth_match_frames(seed, nt=nt, nq=nq)
        qb, tb = ucoslam_b200.bow_index(node(q, 6)), ucoslam_b200.bow_index(node(t, 6))
        rng = np.random.default_rng(seed)
        qflags, tflags = (rng.random(nq) < 0.1).astype(np.uint8), (rng.random(nt) < 0.1).astype(np.uint8)   # FLAG_NONMAXIMA
        m = oracle_py.ref_frame_match(q, qk, t, tk, kind="bow", q_bow=qb, t_bow=tb, q_flags=qflags, t_flags=tflags, **kw)
        m2 = oracle_py.frame_match_bow(q, qk, qb, t, tk, tb, q_usable=1 - qflags, t_usable=1 - tflags, **kw)
        assert np.array_equal(m, m2), "BoW restatement differs from the reference"
        out.update({name + "_q": q, name + "_qk": qk, name + "_t": t, name + "_tk": tk, name + "_out": m, name + "_qflags": qflags, name + "_tflags": tflags})
        for i, a in enumerate(qb):
            out["%s_qb%d" % (name, i)] = np.asarray(a)
        for i, a in enumerate(tb):
            out["%s_tb%d" % (name, i)] = np.asarray(a)
        for k, v in dict(min_desc_dist=50.0, ratio=0.8, check_orientation=True, max_octave_diff=1).items():
            out["%s_%s" % (name, k)] = np.float32(kw.get(k, v))
        if "F12" in kw:
            out[name + "_F12"] = kw["F12"]
    np.savez_compressed(os.path.join(HERE, "match_ref.npz"), **out)
    print("frame-matcher golden written (reference-compiled)", {n: len(out[n + "_out"]) for n in "abcdepqrs"})


KFDB_CASES = {  # name -> (vocabulary kwargs or None = the shipped orb.fbow, synth_places kwargs)
    "s": (dict(seed=5, k=10, depth=4, weight_scale=0.02), dict(seed=1, n_places=12, views_per_place=5, n_desc=400)),
    "t": (dict(seed=6, k=8, depth=5, leaf_prob=0.02, weight_scale=0.004), dict(seed=2, n_places=5, views_per_place=8, n_desc=700, replace_frac=0.5)),
    "orb": (None, dict(seed=3, n_places=6, views_per_place=4, n_desc=300)),
}


def kfdb():
    """KeyFrameDataBase::relocalizationCandidates of the reference (its own keyframedatabase.cpp + covisgraph.cpp + fbow compiled
    unchanged, oracle/_ref/libref_kfdb.so) on seeded keyframes that revisit a few places."""
    import tempfile
    oracle_py.build_ref()
    out = {}
    for name, (vkw, pkw) in KFDB_CASES.items():
        if vkw is None:
            path = oracle_py.REF_VOC_PATH
        else:
            path = os.path.join(tempfile.mkdtemp(), "v.fbow")
            oracle_py.synth_vocabulary(**vkw).tofile(path)
        ref = oracle_py.RefKeyFrameDataBase(path)
        frames, place = oracle_py.synth_places(**pkw)
        ids = (np.arange(len(frames)) * 3 + 7).astype(np.uint32)
        ids[::7] += 1000          # insertion order is not id order
        bows = [ref.add(i, d) for i, d in zip(ids, frames)]
        ea, eb, ew = oracle_py.synth_covis(pkw["seed"] + 100, ids, place)
        for a, b, w in zip(ea, eb, ew):
            ref.covis_edge(a, b, w)
        off, words, weights = oracle_py._csr(bows)
        out[name + "_ids"], out[name + "_off"], out[name + "_words"], out[name + "_weights"] = ids, off, words, weights
        out[name + "_ea"], out[name + "_eb"], out[name + "_ew"] = ea, eb, ew
        qkw = dict(pkw); qkw["seed"] += 50; qkw["views_per_place"] = 1
        qframes, _ = oracle_py.synth_places(**qkw)
        nq = 0
        deleted = []
        for phase in range(2):
            for qi, qd in enumerate(qframes[:4]):
                for sorted_ in (1, 0):
                    for ms, exc in ((0.0, []), (0.05, [int(ids[2]), int(ids[9]), 123456])):
                        cand, qbow = ref.query(qd, sorted_, ms, exc)
                        k = "%s_q%d" % (name, nq)
                        out[k + "_words"], out[k + "_weights"] = qbow
                        out[k + "_prm"] = np.array([sorted_, ms, phase], np.float64)
                        out[k + "_exc"] = np.array(exc, np.uint32)
                        out[k + "_cand"] = cand
                        nq += 1
            if phase == 0:      # the second phase queries after deletions
                deleted = [int(x) for x in ids[1::4]]
                for d in deleted:
                    ref.delete(d)
        out[name + "_deleted"] = np.array(deleted, np.uint32)
        out[name + "_nq"] = np.int32(nq)
        # KeyFrameDataBase::score (float) of a few pairs still in the database
        alive = [int(x) for x in ids if int(x) not in deleted]
        pairs = [(alive[i], alive[(i * 5 + 3) % len(alive)]) for i in range(12)]
        out[name + "_pairs"] = np.array(pairs, np.uint32)
        out[name + "_pair_score"] = np.array([ref.score(a, b) for a, b in pairs], np.float32)
        ref.close()
    np.savez_compressed(os.path.join(HERE, "kfdb_ref.npz"), **out)
    print("kfdb golden written", len(out), {n: int(out[n + "_nq"]) for n in KFDB_CASES})


if __name__ == "__main__":
    which = sys.argv[1:] or ["knn", "bow", "ba", "ba_markers", "ba_cams", "ba_planar", "pnp", "project", "match", "kfdb"]
    for w in which:
        globals()[w]()
