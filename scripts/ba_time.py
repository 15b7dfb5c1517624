"""Wall / device time of uco_b200_ba_solve on a config-2 sized local-BA window next to the reference's g2o on the host."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, ucoslam_b200, oracle_py
ctx = ucoslam_b200.Context(0)
for kw in (dict(seed=21, n_poses=12, n_fixed=2, n_points=2000), dict(seed=23, n_poses=30, n_fixed=5, n_points=3000)):
    pb = oracle_py.synth_ba_problem(**kw)
    for _ in range(3):
        out = ctx.ba_solve(pb, 5)
    n0 = ctx.launch_count()
    t = time.perf_counter()
    for _ in range(10):
        out = ctx.ba_solve(pb, 5)
    wall = (time.perf_counter() - t) / 10
    launches = (ctx.launch_count() - n0) / 10
    t = time.perf_counter(); ref = oracle_py.ref_ba_optimize(pb, 5); tr = time.perf_counter() - t
    print("obs %d iters %s trials %d: wall %.3f ms device %.3f ms launches %d | g2o %.1f ms" % (
        len(pb["obs_pose"]), out["iters"], int(out["trace"][:, 1].sum()), wall * 1e3, out["device_ms"], launches, tr * 1e3))
