"""Multi-GPU plumbing of the hot path: independent units (frames / cameras / BA windows) shard across ranks with NO data-path
collective; torch.distributed is used only for the rendezvous, the barrier around the timed region and the
max-over-ranks of the step time.  Works with backend "nccl" (one process per GPU) and "gloo" (CPU tests)."""
import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_units, rank, world):
    """contiguous, balanced [begin, end) of n_units for this rank (the first n_units % world ranks get one more)"""
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def unit_seed(base_seed, rank, index):
    """seed of the index-th local unit of a rank: distinct across ranks so every GPU works on its own stream of frames"""
    return base_seed + 1000003 * rank + index


def init(backend, device=None):
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """the slowest rank's time: what a whole-job throughput is computed from"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def make_comm(ctx, device="cpu"):
    """NCCL communicator of a library context over the ranks of the initialised process group: rank 0's 128-byte id travels through
    torch.distributed (the plumbing); the data path then talks NCCL directly from the library (uco_b200_comm_*)."""
    import torch
    import torch.distributed as dist
    rank, world, _ = env_rank_world()

    def bcast(data):
        t = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            t = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    return ctx.comm_create(rank, world, bcast if world > 1 else None)


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
