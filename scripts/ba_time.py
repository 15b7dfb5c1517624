"""Wall / device time of the BA solver forms on a config-2 sized local-BA window next to the reference's g2o on the host."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, ucoslam_b200, oracle_py
ctx = ucoslam_b200.Context(0)
quick = "--quick" in sys.argv
for kw in (dict(seed=21, n_poses=12, n_fixed=2, n_points=2000), dict(seed=23, n_poses=30, n_fixed=5, n_points=3000)):
    pbs = [oracle_py.synth_ba_problem(**dict(kw, seed=kw["seed"] + 100 * i)) for i in range(18)]
    pb = pbs[0]
    if not quick:
        t = time.perf_counter(); ref = oracle_py.ref_ba_optimize(pb, 5); tr = time.perf_counter() - t
        print("obs %d  g2o %.1f ms" % (len(pb["obs_pose"]), tr * 1e3))
    for name, mode in (("streamed", (1, 0)), ("cluster4", (2, 4)), ("cluster8", (2, 8)), ("cluster16", (2, 16))):
        if quick and name != "cluster8":
            continue
        ctx.ba_set_mode(*mode)
        for nb in (1, 8, 18):
            if name == "streamed" and nb > 1:
                continue
            packed = ctx.ba_pack_batch(pbs[:nb], 5)
            for _ in range(2):
                outs = ctx.ba_solve_batch(None, 5, packed=packed)
            reps = 5
            t = time.perf_counter()
            for _ in range(reps):
                outs = ctx.ba_solve_batch(None, 5, packed=packed)
            wall = (time.perf_counter() - t) / reps
            out = outs[0]
            print("  %-9s batch %2d: wall %.3f ms  device %.3f ms  (%.3f ms / window)  iters %s trials %d" % (
                name, nb, wall * 1e3, out["device_ms"], wall * 1e3 / nb, out["iters"], int(out["trace"][:, 1].sum())))
            if nb == 1 and out["profile"].sum() > 0:
                names = ["init", "err0", "lin_obs", "lin_sum", "prep", "gather", "solve", "update", "errors", "decide", "results", "flag",
                         "solve:assemble", "solve:factor", "solve:backsub", "gather:warp0"]
                print("      phase us (CTA 0 @1.965 GHz): " + "  ".join("%s %.0f" % (n, c / 1965.0) for n, c in zip(names, out["profile"])))
