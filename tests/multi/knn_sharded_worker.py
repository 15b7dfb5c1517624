"""torchrun worker: k-NN over a map row-sharded across WORLD_SIZE GPUs (all-gather of per-shard top-k + merge) equals the scan of
the whole map on one GPU.  Also the config-4 timing: python -m torch.distributed.run --nproc-per-node N ... knn_sharded_worker.py bench"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import numpy as np, torch
import ucoslam_b200
from ucoslam_b200 import shard

rank, world, local = shard.env_rank_world()
torch.cuda.set_device(local)
shard.init("nccl", torch.device("cuda", local))
ctx = ucoslam_b200.Context(local)
stream = torch.cuda.ExternalStream(ctx.stream, device=local)
comm = shard.make_comm(ctx, "cuda") if world > 1 else None
bench = len(sys.argv) > 1 and sys.argv[1] == "bench"
nt, nq, k = (1_000_000, 2000, 10) if bench else (200_003, 300, 10)
rng = np.random.default_rng(1234)                            # the same map and queries on every rank
t_h = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
q_h = t_h[rng.integers(0, nt, nq)].copy()
q_h[:, :2] ^= 0x5A
t, q = torch.from_numpy(t_h).cuda(), torch.from_numpy(q_h).cuda()
b, e = shard.shard_range(nt, rank, world)
full_i = torch.empty((nq, k), dtype=torch.int32, device="cuda"); full_d = torch.empty_like(full_i)
sh_i = torch.empty_like(full_i); sh_d = torch.empty_like(full_i)
torch.cuda.synchronize()
ctx.hamming_knn_sharded_dev(None, q.data_ptr(), nq, t.data_ptr(), nt, 0, k, full_i.data_ptr(), full_d.data_ptr())
ctx.hamming_knn_sharded_dev(comm, q.data_ptr(), nq, t[b:].data_ptr(), e - b, b, k, sh_i.data_ptr(), sh_d.data_ptr())
ctx.sync()
# identical distance lists; identical rows below the k-th distance; rows tied at the k-th distance are a function of the scan order
# in the reference (its heap evicts whichever tie sits at the root), so there the merged lists keep the lowest row indices instead
below = full_d < full_d[:, -1:]
ok = bool(torch.equal(full_d, sh_d) and torch.equal(full_i[below], sh_i[below]))
tie = (~below).nonzero()
dq = q[tie[:, 0]].view(torch.int64) ^ t[sh_i[~below].long()].view(torch.int64)
cnt = sum(((dq >> s) & 1) for s in range(64)).sum(1) if len(tie) else torch.zeros(0, device="cuda")
ok = ok and bool(torch.equal(cnt.to(torch.int32), sh_d[~below]))
line = {}
if bench:
    def timed(fn, reps=10):
        fn(); ctx.sync(); shard.barrier()
        with torch.cuda.stream(stream):
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            z.record(stream)
        ctx.sync()
        return shard.max_over_ranks(a.elapsed_time(z) / reps, "cuda")
    ms_full = timed(lambda: ctx.hamming_knn_sharded_dev(None, q.data_ptr(), nq, t.data_ptr(), nt, 0, k, full_i.data_ptr(), full_d.data_ptr()))
    ms_sh = timed(lambda: ctx.hamming_knn_sharded_dev(comm, q.data_ptr(), nq, t[b:].data_ptr(), e - b, b, k, sh_i.data_ptr(), sh_d.data_ptr()))
    ms_q1 = timed(lambda: ctx.hamming_knn_sharded_dev(None, q.data_ptr(), 8, t.data_ptr(), nt, 0, k, full_i.data_ptr(), full_d.data_ptr()))
    popc_peak = 16 * 148 * 1.965e9
    line = {"workload": "config4: %d query descriptors against a %d-row map, k=%d" % (nq, nt, k), "n_gpus": world,
            "ms_one_gpu_whole_map": ms_full, "ms_sharded": ms_sh, "queries_per_s_sharded": nq / (ms_sh * 1e-3),
            "popc32_per_s_one_gpu": nq * nt * 8.0 / (ms_full * 1e-3), "frac_of_popc_peak_one_gpu": nq * nt * 8.0 / (ms_full * 1e-3) / popc_peak,
            "ms_8_queries_whole_map": ms_q1, "gbs_8_queries": nt * 32 / (ms_q1 * 1e-3) / 1e9,
            "allgather_bytes_per_rank": nq * k * 8}
bad = shard.sum_over_ranks(0.0 if ok else 1.0, "cuda")
if rank == 0:
    if line:
        print(json.dumps(line), flush=True)
    print("KNN_SHARDED_OK" if bad == 0 else "KNN_SHARDED_FAIL", flush=True)
if comm is not None:
    ctx.comm_destroy(comm)
shard.barrier()
shard.finalize()
sys.stdout.flush()
os._exit(0 if bad == 0 else 1)
