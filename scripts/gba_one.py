import os, sys
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import ucoslam_b200
from ucoslam_b200.synth import synth_global_ba
ctx = ucoslam_b200.Context(0)
pb = synth_global_ba(42)
for _ in range(2):
    out = ctx.ba_solve_sharded(pb, 5)
print(out["device_ms"])
os._exit(0)
