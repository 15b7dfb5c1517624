"""BASELINE config 3 (stereo half): depth association of one 1280x720 rectified pair with 4000 ORB keypoints per image, keypoints
and descriptors from the device extractor.  One JSON line: host-call time of uco_b200_stereo_depth (images + keypoints uploaded
inside the call), the plain-C restatement of the reference's loop on one host core, and parity."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ucoslam_b200
from ucoslam_b200.synth import synth_stereo

ctx = ucoslam_b200.Context(0)
sc = synth_stereo(9, w=1280, h=720, n=10, max_disp=64.0)
prm = ucoslam_b200.OrbParams(4000)
kl, dl = ctx.orb_extract(sc["img_l"], prm)
kr, dr = ctx.orb_extract(sc["img_r"], prm)
pair = dict(sc, kps_l=kl, desc_l=dl, kps_r=kr, desc_r=dr, fx=np.float32(1050.0))
for _ in range(5):
    d, m, n = ctx.stereo_depth(pair)
t = []
for _ in range(50):
    t0 = time.perf_counter(); d, m, n = ctx.stereo_depth(pair); t.append((time.perf_counter() - t0) * 1e3)
line = {"workload": "config3 (stereo half): depth association, 1280x720 pair, %d / %d ORB keypoints" % (len(kl), len(kr)),
        "ms_per_pair_host_call": float(np.median(t)), "keypoints_with_depth": int(n), "associated": int((m >= 0).sum())}
import oracle_py
tc = []
for _ in range(5):
    t0 = time.perf_counter(); od, om, on = oracle_py.stereo_depth(pair); tc.append((time.perf_counter() - t0) * 1e3)
line["cpu_port_ms_per_pair"] = float(np.median(tc)); line["cpu_cores"] = 1
line["bit_exact_vs_oracle"] = bool(np.array_equal(od.view(np.uint32), d.view(np.uint32)) and np.array_equal(om, m) and on == n)
print(json.dumps(line), flush=True)
