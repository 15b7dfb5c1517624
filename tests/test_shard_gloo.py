"""The N>1 host logic on CPU: two gloo ranks shard units without overlap, see distinct seeds, and agree on the max / sum of
their step times (what bench.py computes its whole-job throughput from)."""
import os, subprocess, sys, json
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, os.path.join(%r, "ucoslam-cv3_b200", "python"))
from ucoslam_b200 import shard
rank, world = shard.init("gloo")
b, e = shard.shard_range(37, rank, world)
seeds = [shard.unit_seed(500, rank, i) for i in range(e - b)]
shard.barrier()
mx = shard.max_over_ranks(10.0 + rank)
sm = shard.sum_over_ranks(e - b)
print(json.dumps({"rank": rank, "world": world, "range": [b, e], "seeds": seeds, "max": mx, "sum": sm}))
shard.finalize()
'''


def test_shard_range_is_a_partition():
    sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
    from ucoslam_b200 import shard
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


@pytest.mark.timeout(180)
def test_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=150)
        assert p.returncode == 0, e
        outs.append(json.loads(o.strip().splitlines()[-1]))
    outs.sort(key=lambda d: d["rank"])
    assert outs[0]["range"] == [0, 19] and outs[1]["range"] == [19, 37]
    assert not set(outs[0]["seeds"]) & set(outs[1]["seeds"])
    assert outs[0]["max"] == outs[1]["max"] == 11.0
    assert outs[0]["sum"] == outs[1]["sum"] == 37.0


def test_ba_landmark_partition_covers_and_balances():
    """host logic of the sharded global BA (uco_b200_probe_ba_partition): contiguous landmark ranges that tile [0, N) and carry
    equal shares of the observations; two gloo ranks derive the same boundaries independently (every rank plans the full graph)."""
    sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
    import numpy as np, ucoslam_b200
    from ucoslam_b200.synth import synth_global_ba
    pb = synth_global_ba(9, n_kf=50, n_points=4000)
    M, N = len(pb["obs_pose"]), len(pb["points3"])
    for world in (1, 2, 3, 8):
        L, cnt = ucoslam_b200.probe_ba_partition(pb, world)
        assert L[0] == 0 and L[-1] == N and np.all(np.diff(L) >= 0)
        assert cnt.sum() == M and cnt.max() - cnt.min() <= 12      # within two landmarks' worth of observations
        per_lm = np.bincount(pb["obs_point"], minlength=N)
        assert [int(per_lm[L[r]:L[r + 1]].sum()) for r in range(world)] == cnt.tolist()
