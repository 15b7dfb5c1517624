"""Launch times of the two-level block-envelope Cholesky on BASELINE config 5's keyframe graph for a range of separator counts.
Run under ncu (--metrics gpu__time_duration.sum --clock-control none): the launch list, in order, holds per force_k one assemble,
one band_front (interior fronts), one band_front (root) and one band_back launch.  Prints the planner's numbers per force_k."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ucoslam_b200
from test_block_solve import graphs, block_system

ctx = ucoslam_b200.Context(0)
ctx.set_profiling(1)
rng = np.random.default_rng(1)
nb, edges = graphs(rng)["ring_498_w5"]
ij, blocks, S = block_system(rng, nb, edges)
b = rng.normal(0, 1, 6 * nb)
ref = np.linalg.solve(S, b)
for k in [int(v) for v in (sys.argv[1:] or "0 2 102 4 104 6 106 8 108 12 112 16 116".split())]:
    x, info = ctx.block_solve(nb, ij, blocks, b, force_k=k)
    print("force_k", k, "info", info.tolist(), "err", float(np.abs(x - ref).max()), "device_us", round(1e3 * ctx.block_solve_ms(), 1), flush=True)
ctx.close()
