import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ucoslam_b200, orb_oracle as oo
ctx = ucoslam_b200.Context(0)
img = oo.synth_frame(0)
prm = ucoslam_b200.OrbParams(2000)
K, D = ctx.orb_extract(img, prm)
oK, oD, inter = oo.extract(img, want_intermediates=True)
out = {}
for l, lv in enumerate(inter["levels"]):
    got = ctx.orb_selected(0, l)
    ref = lv["selected"]
    refp = (ref["response"].astype(np.uint32) << 24) | (ref["y"].astype(np.uint32) << 12) | ref["x"].astype(np.uint32)
    print("level", l, "n", len(got), len(ref), "same order", np.array_equal(got, refp), "same set", set(got.tolist()) == set(refp.tolist()),
          "only got", len(set(got.tolist()) - set(refp.tolist())), "only ref", len(set(refp.tolist()) - set(got.tolist())))
    out["got%d" % l] = got; out["ref%d" % l] = refp
lv0 = inter["levels"][0]
bad = 0
for ci, ((i, j), c) in enumerate(sorted(lv0["candidates"].items())):
    cell = i * lv0["grid"]["level_cols"] + j
    got, counts, geom = ctx.orb_candidates(0, cell)
    x0, y0 = lv0["ini_x_col"][j], lv0["ini_y_row"][i]
    ref = sorted(((int(k["response"]) << 24) | ((int(k["y"]) + y0) << 12) | (int(k["x"]) + x0)) for k in c)
    g = sorted(got.tolist())
    if g != ref:
        bad += 1
        if bad < 4:
            dec = lambda a: [(v >> 24, (v >> 12) & 0xfff, v & 0xfff) for v in a]
            print("cell", i, j, "geom", geom, "counts", counts, "ref n", len(ref)); print(" got", dec(g)[:8]); print(" ref", dec(ref)[:8])
print("level0 candidate cells bad:", bad)
print("final", len(K), len(oK), "desc equal", np.array_equal(D, oD) if len(K) == len(oK) else None)
np.savez(os.path.join(ROOT, "gpurun_out", "dbg_orb.npz"), K=K, D=D, oK=oK, oD=oD, **out)
