"""Two config-5 global-BA solves on one GPU (the second one is warm), for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gba_launches.csv python scripts/gba_one.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import ucoslam_b200
from ucoslam_b200.synth import synth_global_ba
ctx = ucoslam_b200.Context(0)
pb = synth_global_ba(42)
for _ in range(2):
    out = ctx.ba_solve_sharded(pb, 5)
print(out["device_ms"])
os._exit(0)
