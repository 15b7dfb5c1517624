"""New-map-point creation as a unit (MapManager::createNewPoints, /root/reference/src/utils/mapmanager.cpp:9772-10788, de-obfuscated):
the restatement over the pinned matcher oracle and the cv2-SVD triangulation restatement, on a synthetic keyframe + neighbours."""
import numpy as np
import oracle_py
from ucoslam_b200.synth import synth_new_points_scene


def test_restatement_recovers_the_scene():
    sc = synth_new_points_scene(3, n_kp=800, n_nb=4)
    ref = oracle_py.new_points_py(sc)
    n = len(ref["kpt"])
    assert n > 100
    assert np.all(np.diff(ref["kpt"]) > 0)                                 # std::map order
    assert np.isin(ref["kpt"], sc["t_map"]).all()                          # MODE_UNASSIGNED: only keypoints without a map point
    # the points are where the scene put them (correct matches dominate)
    pid = sc["t_point"][ref["kpt"]]
    ok = pid >= 0
    err = np.linalg.norm(ref["xyz"][ok] - sc["points_gt"][pid[ok]], axis=1)
    depth = np.linalg.norm(sc["points_gt"][pid[ok]] - np.linalg.inv(sc["kf_pose"])[:3, 3], axis=1)
    assert np.median(err / depth) < 0.05
    # observations: neighbour order inside a point, unassigned keypoints of that neighbour, at most one per neighbour
    for j in range(n):
        fr = ref["obs_frame"][ref["obs_ptr"][j]:ref["obs_ptr"][j + 1]]
        kp = ref["obs_kpt"][ref["obs_ptr"][j]:ref["obs_ptr"][j + 1]]
        assert len(fr) >= 1 and np.all(np.diff(fr) > 0)
        for f, k in zip(fr, kp):
            assert k in sc["q_map"][f]
    assert ref["obs_ptr"][-1] > n                                          # some points are seen by several neighbours


def test_max_points_keeps_the_smallest_distances():
    sc = synth_new_points_scene(4, n_kp=600, n_nb=3)
    full = oracle_py.new_points_py(sc)
    cut = oracle_py.new_points_py(sc, max_points=50)
    assert len(cut["kpt"]) == 50
    assert np.all(np.diff(cut["dist"]) >= 0)
    assert cut["dist"].max() <= np.sort(full["dist"])[49]


def test_empty_inputs():
    sc = synth_new_points_scene(5, n_kp=300, n_nb=2)
    sc["t_map"] = np.zeros(0, np.int32)
    ref = oracle_py.new_points_py(sc)
    assert len(ref["kpt"]) == 0 and ref["obs_ptr"].tolist() == [0]
