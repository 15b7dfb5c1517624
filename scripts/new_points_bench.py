"""New-map-point creation of one keyframe as a unit (MapManager::createNewPoints): a keyframe of ~2000
keypoints against 20 covisible neighbours.  Device path through the host-buffer C ABI (upload + 3 launches + download + host merge
inside the timed region) and, with --ref, the CPU restatement (pinned C matcher oracle + cv2-SVD triangulation, one core) beside it.

    python scripts/new_points_bench.py [--ref]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import numpy as np
import ucoslam_b200
from ucoslam_b200.synth import synth_new_points_scene

ctx = ucoslam_b200.Context(0)
sc = synth_new_points_scene(21, n_kp=2600, n_nb=20, assigned_frac=0.35)
for _ in range(3):
    out = ctx.new_points(sc)
l0 = ctx.launch_count()
reps = 20
t0 = time.perf_counter()
for _ in range(reps):
    out = ctx.new_points(sc)
ms = (time.perf_counter() - t0) / reps * 1e3
launches = (ctx.launch_count() - l0) / reps
rows_q = int(sum(len(m) for m in sc["q_map"]))
line = {"workload": "new-map-point creation: keyframe with %d unassigned keypoints (of %d) x %d neighbours (%d unassigned keypoints in all)" %
        (len(sc["t_map"]), len(sc["t_kps"]), len(sc["q_desc"]), rows_q),
        "ms_per_keyframe_host_buffers": ms, "launches_per_keyframe": launches, "new_points": int(len(out["kpt"])),
        "observations": int(out["obs_ptr"][-1]), "hamming_pairs": int(rows_q * len(sc["t_map"]))}
if "--ref" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    t = time.perf_counter()
    ref = oracle_py.new_points_py(sc)
    line["cpu_restatement_ms"] = (time.perf_counter() - t) * 1e3
    line["cpu_restatement"] = "C matcher oracle (exact linear 10-NN + filters) + numpy/cv2-SVD triangulation, 1 core"
    line["same_point_count_up_to_borderline"] = abs(len(ref["kpt"]) - len(out["kpt"])) <= 3
print(json.dumps(line), flush=True)
ctx.close()
