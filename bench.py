#!/usr/bin/env python
"""bench.py — frames/s of the per-frame tracking hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F]

A step is one pass of the hot path over one batch of F synthetic frames of BASELINE config 2
(640x480 mono, 2000 ORB keypoints/frame, 8 levels, local-BA window 10 KF).  Stages implemented so far are listed in
config.stages; every stage runs through the C ABI of libucoslam_b200.so (no CPU fallback).
  value  : frames/s with the step's inputs already resident in HBM (device-timed with CUDA events on the context stream,
           L2 flushed between steps, max over ranks)
  e2e    : the same work driven through the host-buffer C-ABI calls (H2D + D2H copies inside the timed region)
  roofline / cpu_baseline : see DESIGN.md "Measurement"
Multi-GPU: frames shard across ranks (independent units, no data-path collective) -> weak scaling.
"""
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))

KPTS, K_NN = 2000, 10
WORKLOAD = "config2: 640x480 mono tracking, 2000 ORB kpts/frame, local-BA window=10 KF"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for ln in out.strip().splitlines():
                    self.rows.append([c.strip() for c in ln.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def synth_frames_descriptors(n_frames, seed):
    """Seeded synthetic ORB descriptors of consecutive frames: frame f+1 re-observes frame f's features with a few
    flipped bits."""
    import numpy as np
    rng = np.random.default_rng(seed)
    cur = rng.integers(0, 256, (KPTS, 32), dtype=np.uint8)
    out = [cur]
    for _ in range(n_frames):
        nxt = cur[rng.permutation(KPTS)].copy()
        flips = rng.integers(0, 256, (KPTS, 12))
        mask = np.zeros((KPTS, 32), np.uint8)
        for j in range(12):
            np.bitwise_xor.at(mask, (np.arange(KPTS), flips[:, j] >> 3), (1 << (flips[:, j] & 7)).astype(np.uint8))
        nxt ^= mask
        out.append(nxt)
        cur = nxt
    return np.stack(out)  # (n_frames+1, KPTS, 32)


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np, oracle_py
    frames = min(args.frames, 8)
    desc = synth_frames_descriptors(frames, 1234)
    have_ref = oracle_py.load_ref("libref_xflann.so") is not None

    def one_step():
        for f in range(frames):
            if have_ref:   # what FrameMatcher_Flann does: build HKMeans(32,0) on the train frame, search k=10, 16 checks
                oracle_py.ref_xflann_knn(desc[f + 1], desc[f], K_NN, 1, 16, 0)
            else:
                oracle_py.hamming_knn(desc[f + 1], desc[f], K_NN, 0)

    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    fps = frames * args.steps / dt
    kind = "reference" if have_ref else "port"
    line = {"impl": "reference", "metric": "frames/sec (ORB+match+local-BA) 640x480 mono", "value": fps,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "stages": ["match"], "frames_per_step": frames},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": kind,
                             "sample": "%d frames/step: xflann %s k=10 per frame pair, 1 thread (the reference runs "
                                       "xflann with threads=1)" % (frames, "HKMeans(32,0) build + 16-check search"
                                                                   if have_ref else "exact linear port")},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import numpy as np, torch
    import torch.distributed as dist
    import ucoslam_b200

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = ucoslam_b200.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    F = args.frames
    desc = synth_frames_descriptors(F, 1234 + rank)
    desc_pin = torch.from_numpy(desc).pin_memory()
    with torch.cuda.stream(stream):
        desc_dev = desc_pin.to("cuda", non_blocking=True)
        idx_dev = torch.empty((F, KPTS, K_NN), dtype=torch.int32, device="cuda")
        dist_dev = torch.empty((F, KPTS, K_NN), dtype=torch.int32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    idx_host = torch.empty((F, KPTS, K_NN), dtype=torch.int32).pin_memory()
    dist_host = torch.empty((F, KPTS, K_NN), dtype=torch.int32).pin_memory()
    ctx.sync()

    def step_device():
        for f in range(F):
            ctx.hamming_knn_dev(desc_dev[f + 1].data_ptr(), KPTS, desc_dev[f].data_ptr(), KPTS, K_NN,
                                ucoslam_b200.UCO_KNN_HEAP, idx_dev[f].data_ptr(), dist_dev[f].data_ptr())

    def step_host():  # the reference-facing call with HOST buffers: H2D + kernel + D2H per frame
        lib, h = ctx.lib, ctx.h
        for f in range(F):
            rc = lib.uco_b200_hamming_knn(h, desc_pin[f + 1].data_ptr(), KPTS, 32, desc_pin[f].data_ptr(), KPTS, 32,
                                          K_NN, 0, idx_host[f].data_ptr(), dist_host[f].data_ptr())
            if rc != 0:
                raise RuntimeError(lib.uco_b200_last_error(h))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps, warmup):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                fn()
                flush.zero_()
            barrier()
            evs = []
            n0 = ctx.launch_count()
            for _ in range(steps):
                flush.zero_()  # L2 flush between timed iterations (outside the event pair)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                fn()
                b.record(stream)
                evs.append((a, b))
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
            launches = ctx.launch_count() - n0
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    # e2e: wall clock around host-API steps (the call synchronises internally), max over ranks
    for _ in range(max(1, args.warmup)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clocks = sampler.summary() if sampler else None

    # roofline of the dominant kernel, timed live with CUDA events on the launching stream (same inputs, L2 flushed)
    with torch.cuda.stream(stream):
        per = []
        for f in range(min(F, 16)):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.hamming_knn_dev(desc_dev[f + 1].data_ptr(), KPTS, desc_dev[f].data_ptr(), KPTS, K_NN, 0,
                                idx_dev[f].data_ptr(), dist_dev[f].data_ptr())
            b.record(stream)
            per.append((a, b))
        barrier()
        k_ms = sum(a.elapsed_time(b) for a, b in per) / len(per)
    alg_bytes = 2 * KPTS * 32 + KPTS * K_NN * 8
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9

    if rank == 0:
        total_frames = F * world * args.steps
        line = {"metric": "frames/sec (ORB+match+local-BA) 640x480 mono", "value": total_frames / (ms_dev * 1e-3),
                "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "stages": ["match"], "frames_per_step_per_gpu": F,
                           "parallelism": "frames sharded over %d GPU(s), no collective" % world,
                           "l2": "flushed between timed steps (256 MB write)"},
                "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": "frames/s",
                        "h2d_bytes_per_step": F * 2 * KPTS * 32, "d2h_bytes_per_step": F * KPTS * K_NN * 8},
                "gpu_launches": launches,
                "clocks": clocks,
                "roofline": {"kernel": "hamming_knn_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                             "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                             "peak_source": "measured" if peaks else "fallback",
                             "kernel_us": k_ms * 1e3,
                             "note": "2000x2000 k-NN is L2-resident and popc/latency bound; see DESIGN.md"}}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(desc)
        print(json.dumps(line))
        sys.stdout.flush()
    barrier()
    if world > 1:
        dist.destroy_process_group()
    # the context (and its stream) outlives every torch object that references the stream; skip interpreter teardown
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def cpu_baseline(desc):
    """Bounded CPU sample of the same workload on this box's host cores (rank 0, N=1)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    have_ref = oracle_py.load_ref("libref_xflann.so") is not None
    n = min(len(desc) - 1, 8)
    t0 = time.perf_counter()
    for f in range(n):
        if have_ref:
            oracle_py.ref_xflann_knn(desc[f + 1], desc[f], K_NN, 1, 16, 0)
        else:
            oracle_py.hamming_knn(desc[f + 1], desc[f], K_NN, 0)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "reference" if have_ref else "port",
            "sample": "%d frame pairs, xflann %s, k=10" % (n, "HKMeans(32,0) build + 16-check search (reference setting)"
                                                            if have_ref else "exact linear port")}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
