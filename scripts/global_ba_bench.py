"""BASELINE config 5: global BA, 500 keyframes / 50k landmarks / ~300k observations, landmarks sharded over N GPUs with one NCCL
all-reduce of the packed reduced Hessian per LM trial.  One JSON line on rank 0.

    python scripts/global_ba_bench.py                                         # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/global_ba_bench.py [--ref]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import numpy as np, torch
import ucoslam_b200
from ucoslam_b200 import shard
from ucoslam_b200.synth import synth_global_ba

rank, world, local = shard.env_rank_world()
torch.cuda.set_device(local)
shard.init("nccl", torch.device("cuda", local))
ctx = ucoslam_b200.Context(local)
comm = shard.make_comm(ctx, "cuda") if world > 1 else None
pb = synth_global_ba(42)                       # 500 KF on a 50 m loop, 50 000 landmarks, 6 observations each
for _ in range(2):
    out = ctx.ba_solve_sharded(pb, 5, comm=comm)
reps = 5
dev = []
shard.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    out = ctx.ba_solve_sharded(pb, 5, comm=comm)
    dev.append(out["device_ms"])
torch.cuda.synchronize()
shard.barrier()
wall = shard.max_over_ranks((time.perf_counter() - t0) / reps * 1e3, "cuda")
dev_ms = shard.max_over_ranks(float(np.mean(dev)), "cuda")
if rank == 0:
    trials = int(out["trace"][:, 1].sum())
    line = {"workload": "config5: global BA 500 KF / 50k landmarks / %d observations, nIters=5 (5 + <=10 LM iterations)" % len(pb["obs_pose"]),
            "n_gpus": world, "ms_per_solve_wall": wall, "ms_per_solve_device": dev_ms, "lm_iterations": out["iters"].tolist(), "lm_trials": trials,
            "ms_per_trial": dev_ms / max(1, trials), "allreduce_bytes_per_trial": float(out["profile"][3]), "schur_blocks": int(out["profile"][2]),
            "host_ms": {"structure_plan_upload": float(out["profile"][4]), "lm_loop_wall": float(out["profile"][5]), "download_scatter": float(out["profile"][6])},
            "landmarks_per_rank": int(out["profile"][0]), "observations_per_rank": int(out["profile"][1]),
            "final_chi2": float(out["trace"][int(out["iters"].sum()) - 1, 0])}
    if "--ref" in sys.argv:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py
        t = time.perf_counter()
        ref = oracle_py.ref_ba_optimize(pb, 5)
        if ref is not None:
            line["cpu_reference_g2o_ms"] = (time.perf_counter() - t) * 1e3
            line["cpu_reference_cores"] = 1
            line["max_pose_diff_vs_g2o"] = float(np.abs(ref["pose7"] - out["pose7"]).max())
            line["iters_vs_g2o"] = [ref["iters"].tolist(), out["iters"].tolist()]
    print(json.dumps(line), flush=True)
if comm is not None:
    ctx.comm_destroy(comm)
shard.barrier()
shard.finalize()
sys.stdout.flush()
os._exit(0)
