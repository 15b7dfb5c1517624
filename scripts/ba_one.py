"""One config-2 sized local-BA solve in the cluster-resident form (target of ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ucoslam_b200
from ucoslam_b200.synth import synth_ba_problem
ctx = ucoslam_b200.Context(0)
ctx.ba_set_mode(2, 8)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pbs = [synth_ba_problem(21 + 100 * i, n_poses=12, n_fixed=2, n_points=2000) for i in range(nb)]
for _ in range(3):
    out = ctx.ba_solve_batch(pbs, 5)
print(out[0]["device_ms"], out[0]["iters"])
