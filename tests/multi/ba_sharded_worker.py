"""torchrun worker: the sharded global BA on WORLD_SIZE GPUs equals the single-GPU solve (run by tests/test_ba_sharded_gpu.py
and by hand:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi/ba_sharded_worker.py [n_kf n_points])."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import numpy as np, torch
import ucoslam_b200
from ucoslam_b200 import shard
from ucoslam_b200.synth import synth_global_ba

rank, world, local = shard.env_rank_world()
torch.cuda.set_device(local)
shard.init("nccl", torch.device("cuda", local))
ctx = ucoslam_b200.Context(local)
comm = shard.make_comm(ctx, "cuda")
sizes = [(60, 3000), (240, 20000)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
ok = True
for n_kf, n_pts in sizes:
    pb = synth_global_ba(7, n_kf=n_kf, n_points=n_pts)
    one = ctx.ba_solve_sharded(pb, 5)                 # this rank alone
    shard.barrier()
    t0 = time.perf_counter()
    many = ctx.ba_solve_sharded(pb, 5, comm=comm)     # landmarks sharded over the ranks, all-reduce per LM trial
    dt = time.perf_counter() - t0
    dp = float(np.abs(one["pose7"] - many["pose7"]).max())
    dx = float(np.abs(one["point3"] - many["point3"]).max())
    same = np.array_equal(one["iters"], many["iters"]) and np.array_equal(one["trace"][:, 1], many["trace"][:, 1])
    flags = np.array_equal(one["level"], many["level"]) and np.array_equal(one["bad"], many["bad"])
    good = same and dp < 1e-8 and dx < 1e-7 and flags
    ok = ok and good
    print("rank %d/%d  %d KF %d obs: 1 GPU %.1f ms, %d GPUs %.1f ms (wall %.1f)  |dpose| %.2e |dpoint| %.2e iters %s trials-equal %s flags-equal %s  shard %s"
          % (rank, world, n_kf, len(pb["obs_pose"]), one["device_ms"], world, many["device_ms"], dt * 1e3, dp, dx, many["iters"].tolist(), same, flags,
             many["profile"][:4].tolist()), flush=True)
tot = shard.sum_over_ranks(0.0 if ok else 1.0, "cuda")
if rank == 0:
    print("SHARDED_OK" if tot == 0 else "SHARDED_FAIL", flush=True)
ctx.comm_destroy(comm)
shard.barrier()
shard.finalize()
sys.stdout.flush()
os._exit(0 if tot == 0 else 1)
