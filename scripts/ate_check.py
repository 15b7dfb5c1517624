"""ATE of the tracking chain (ORB extract -> projection matcher -> pose-only LM) on a rendered 100-frame 640x480 clip with exact
ground truth, CUDA path vs CPU oracle path (BASELINE metric: 'ATE vs ref').  One JSON line.   python scripts/ate_check.py [n_frames]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ucoslam_b200, oracle_py, orb_oracle
from ucoslam_b200 import chain
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
tex = chain.texture()
gt = np.array([chain.gt_pose(i) for i in range(n)])
frames = [chain.render(tex, T) for T in gt]
ctx = ucoslam_b200.Context(0)
prm = ucoslam_b200.OrbParams(2000)
ext = lambda im: ctx.orb_extract(im, prm)
chain.track(frames[:3], gt[0], ext, ctx.match_projected, ctx.pose_only)   # warm-up
t0 = time.perf_counter()
gp, gs = chain.track(frames, gt[0], ext, ctx.match_projected, ctx.pose_only)
t_gpu = time.perf_counter() - t0
t0 = time.perf_counter()
cp, cs = chain.track(frames, gt[0], lambda im: orb_oracle.extract(im, 2000), oracle_py.match_projected, oracle_py.pose_only)
t_cpu = time.perf_counter() - t0
print(json.dumps({"workload": "tracking chain over %d rendered 640x480 frames (plane scene, exact ground truth), 2000 ORB kpts/frame" % n,
                  "ate_m_cuda_path": chain.ate(gp, gt), "ate_m_cpu_oracle_path": chain.ate(cp, gt),
                  "max_abs_pose_entry_difference": float(np.abs(gp - cp).max()), "identical_match_and_inlier_counts": gs == cs,
                  "min_inliers": int(min(g for _, g in gs)), "mean_matches": float(np.mean([m for m, _ in gs])),
                  "frames_per_s_cuda_path_single_frame_calls_incl_python": (n - 1) / t_gpu, "frames_per_s_cpu_oracle_path": (n - 1) / t_cpu}))
os._exit(0)
