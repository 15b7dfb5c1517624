"""A/B of the pyramid kernels on the GPU: stage times of a 128-frame 640x480 batch (and a 32-frame 720p batch) with the strip forms and,
under UCO_ORB_PYRAMID_V1=1 (a child process: the switch is read when the plan is made), with the per-pixel forms; the child also dumps
every level of frame 0 so the parent can compare the two builds byte for byte.   python scripts/orb_pyramid_ab.py"""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def run(tag):
    import ucoslam_b200, orb_oracle as oo
    ctx = ucoslam_b200.Context(0)
    ctx.set_profiling(True)
    out = {}
    for (w, h, nf, n) in ((640, 480, 2000, 128), (1280, 720, 4000, 32), (641, 479, 1500, 8)):
        frames = [oo.synth_frame(i % 8, w, h) for i in range(n)]
        prm = ucoslam_b200.OrbParams(nf)
        ts = []
        for it in range(6):
            ctx.orb_extract_batch(frames, prm)
            ts.append(ctx.orb_last_stage_ms())
        med = {k: float(np.median([t[k] for t in ts[2:]])) for k in ts[0]}
        out["%dx%d x%d" % (w, h, n)] = med
        lv = [ctx.orb_pyramid_level(0, l) for l in range(8)]
        np.savez("/tmp/pyr_%s_%d.npz" % (tag, w), *lv)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print(json.dumps(run(sys.argv[1])))
        sys.exit(0)
    res = {}
    for tag, env in (("strip", {}), ("v1", {"UCO_ORB_PYRAMID_V1": "1"})):
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, __file__, tag], env=e, capture_output=True, text=True)
        if p.returncode != 0:
            print(tag, "FAILED", p.stderr[-2000:]); continue
        res[tag] = json.loads(p.stdout.strip().splitlines()[-1])
    for k in res.get("strip", {}):
        print(k)
        for tag in res:
            print("  %-6s" % tag, " ".join("%s %.3f" % (a, b) for a, b in res[tag][k].items()))
    for w in (640, 1280, 641):
        try:
            a, b = np.load("/tmp/pyr_strip_%d.npz" % w), np.load("/tmp/pyr_v1_%d.npz" % w)
            print("levels of frame 0 at width %d: differing bytes per level" % w, [int((a[k] != b[k]).sum()) for k in a.files])
        except Exception as ex:
            print("compare failed", ex)
