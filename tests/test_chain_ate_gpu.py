"""'Matched ATE' (BASELINE metric): the tracking chain ORB extract -> projection matcher -> pose-only LM over a rendered clip with
exact ground truth, run once through the CUDA path (C ABI) and once through the CPU oracles.  Same matches, same inlier counts,
the same trajectory to float round-off, and therefore the same absolute trajectory error."""
import numpy as np
import pytest
import oracle_py, orb_oracle
from ucoslam_b200 import chain
import ucoslam_b200

pytestmark = pytest.mark.gpu


def test_gpu_chain_tracks_like_the_cpu_path(ctx):
    n = 12
    tex = chain.texture()
    gt = np.array([chain.gt_pose(i) for i in range(n)])
    frames = [chain.render(tex, T) for T in gt]
    prm = ucoslam_b200.OrbParams(2000)
    gpu_poses, gpu_stats = chain.track(frames, gt[0], lambda im: ctx.orb_extract(im, prm), ctx.match_projected, ctx.pose_only)
    cpu_poses, cpu_stats = chain.track(frames, gt[0], lambda im: orb_oracle.extract(im, 2000), oracle_py.match_projected, oracle_py.pose_only)
    assert gpu_stats == cpu_stats                                   # identical match and inlier counts in every frame
    assert min(g for _, g in gpu_stats) > 300                       # and the chain really tracks
    assert np.abs(gpu_poses - cpu_poses).max() < 1e-5               # f32 poses of a ~2 m scene: round-off of the f64 LM only
    a_gpu, a_cpu = chain.ate(gpu_poses, gt), chain.ate(cpu_poses, gt)
    assert a_gpu < 0.01 and abs(a_gpu - a_cpu) <= 0.01 * max(a_cpu, 1e-6) + 1e-6   # ATE difference <= 1 %
