#!/usr/bin/env python
"""bench.py — frames/s of the per-frame tracking hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F]

A step is one pass of the hot path over one batch of F synthetic frames of BASELINE config 2 (640x480 mono, 2000 ORB
keypoints/frame, 8 levels x1.2, local-BA window 10 KF): ORB extraction of every frame, Hamming k-NN (k=10) of every frame
against its predecessor, and one local bundle adjustment (10 free + 2 fixed keyframes, 2000 points, ~15k observations, the
reference's two-stage 5 + 10 LM iterations) per KF_EVERY = 8 frames.  Every stage runs through the C ABI of libucoslam_b200.so
(no CPU fallback: the library refuses to create a context without a CUDA device).
  value   : frames/s with the step's input frames already resident in HBM (CUDA events on the context stream, L2 flushed
            between timed steps, max over ranks)
  e2e     : the same work through the host-buffer C-ABI calls (pinned host frames in, keypoints/descriptors/matches out;
            H2D and D2H copies inside the timed region)
  roofline: the dominant kernel of the step, timed live with CUDA events on the launching stream
  cpu_baseline / --impl reference : the CPU path on this box's host cores (see DESIGN.md "Measurement")
Multi-GPU: frames shard across ranks (independent units, no data-path collective) -> weak scaling.
"""
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))

W, H, KPTS, K_NN = 640, 480, 2000, 10
METRIC = "frames/sec (ORB+match+local-BA) 640x480 mono"
WORKLOAD = "config2: 640x480 mono tracking, 2000 ORB kpts/frame, local-BA window=10 KF"
STAGES = ["orb_extract", "hamming_knn_match", "local_ba"]
KF_EVERY = 8          # one keyframe (= one local-BA call, mapmanager.cpp:4005) per KF_EVERY frames
BA_WINDOW = dict(n_poses=12, n_fixed=2, n_points=2000)   # 10 free KFs + 2 fixed observers, ~15k observations, nIters = 5
BA_ITERS = 5
N_MAPPERS = int(os.environ.get("UCO_BENCH_MAPPERS", "4"))          # mapper contexts alternating between steps
BA_CLUSTER = int(os.environ.get("UCO_BENCH_BA_CLUSTER", "4"))       # CTAs per BA cluster (0 = library default 8)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--res", default="640x480", choices=["640x480", "1280x720"],
                    help="640x480 / 2000 keypoints = BASELINE config 2 (the metric's configuration, default); 1280x720 / 4000 keypoints = the "
                         "mono part of config 3 (north_star asks for both stream sizes)")
    a = ap.parse_args()
    if a.res == "1280x720":
        global W, H, KPTS, METRIC, WORKLOAD
        W, H, KPTS = 1280, 720, 4000
        METRIC = "frames/sec (ORB+match+local-BA) 1280x720 mono"
        WORKLOAD = "config3 (mono part): 1280x720 tracking, 4000 ORB kpts/frame, local-BA window=10 KF"
    return a


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
            time.sleep(0.05)

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


def synth_clip(n_frames, seed):
    """Seeded synthetic clip: perspective views of a multi-octave block-noise texture along a smooth camera path
    (SURVEY.md 8(d)); FAST-dense, so every frame yields the full 2000 keypoints."""
    import numpy as np, cv2
    rng = np.random.default_rng(seed)
    size = 2048 if W <= 640 else 4096
    acc = np.zeros((size, size), np.float64)
    amp = 1.0
    for blk in (64, 32, 16, 8, 4):
        n = size // blk
        acc += amp * np.kron(rng.random((n, n)), np.ones((blk, blk)))
        amp *= 0.5
    acc -= acc.min()
    tex = (acc / acc.max() * 255).astype(np.uint8)
    out = np.empty((n_frames, H, W), np.uint8)
    for i in range(n_frames):
        t = i * 0.02
        c, s = np.cos(0.15 * np.sin(t)), np.sin(0.15 * np.sin(t))
        zoom = 1.6 + 0.2 * np.sin(0.7 * t)
        Hm = np.array([[c * zoom, -s * zoom, 300 + 120 * t], [s * zoom, c * zoom, 400 + 40 * np.sin(t)],
                       [1e-4 * np.sin(t), 5e-5, 1.0]])
        out[i] = cv2.warpPerspective(tex, Hm, (W, H), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                     borderMode=cv2.BORDER_REFLECT_101)
    return out


# ---------------------------------------------------------------------------------------------------------------------
def ba_windows(n, seed):
    from ucoslam_b200.synth import synth_ba_problem
    return [synth_ba_problem(seed + i, **BA_WINDOW) for i in range(n)]


def cpu_reference_step(frames, oracle_py, orb_oracle, have_xflann, windows=()):
    """The reference's CPU path for the same stages: ORB extraction of every frame (cv2-backed restatement of
    ORBextractor.cpp: the reference itself cannot be linked without OpenCV C++ headers) + FrameMatcher_Flann's
    xflann HKMeans(32,0) build + 16-check k-NN against the previous frame (the reference's own code)."""
    prev = None
    for f in frames:
        k, d = orb_oracle.extract(f, KPTS)
        if prev is not None and len(d) and len(prev):
            if have_xflann:
                oracle_py.ref_xflann_knn(d, prev, K_NN, 1, 16, 0)
            else:
                oracle_py.hamming_knn(d, prev, K_NN, 0)
        prev = d
    for pb in windows:  # the reference's own g2o + typesg2o.h (oracle/_ref) when it was built, else the C restatement
        if oracle_py.ref_ba_optimize(pb, BA_ITERS) is None:
            oracle_py.ba_optimize(pb, BA_ITERS)


def cpu_baseline_info(have_xflann, n):
    return {"cores": 1, "kind": "port",
            "sample": "%d frames + %d local-BA windows: ORB = Python/cv2-4.13 restatement of ORBextractor.cpp (blur/resize/FAST "
                      "native OpenCV, 1 thread; interpreter overhead included), match = %s, BA = the reference's g2o + "
                      "typesg2o.h compiled from its sources (oracle/_ref), 1 thread as the reference runs it" % (
                          n, max(1, n // KF_EVERY), "the reference's xflann HKMeans(32,0) build + 16-check search, 1 thread"
                          if have_xflann else "exact linear port")}


_REF = {}


def _ref_worker_init():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py, orb_oracle
    try:
        import cv2
        cv2.setNumThreads(1)          # one worker per host core: no nested thread pools
    except Exception:
        pass
    _REF["mods"] = (oracle_py, orb_oracle, oracle_py.load_ref("libref_xflann.so") is not None)


def _ref_worker_step(job):
    seed, n = job
    if "mods" not in _REF:
        _ref_worker_init()
    oracle_py, orb_oracle, have = _REF["mods"]
    key = ("data", seed, n)
    if key not in _REF:               # every worker tracks its own stream of frames (generated once, outside the timed steps)
        _REF[key] = (synth_clip(n, seed), ba_windows(max(1, n // KF_EVERY), 500 + seed))
    frames, windows = _REF[key]
    cpu_reference_step(frames, oracle_py, orb_oracle, have, windows)
    return n


def run_reference(args, rank, world):
    """The reference's CPU path on ALL host cores of the box: one worker process per core, each tracking its own stream of frames
    (the same sharding by independent streams the GPU arm uses); a step = every worker processes its bounded sample."""
    if rank != 0:
        return
    import multiprocessing as mp
    workers = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    n = min(args.frames, 8)
    jobs = [(1234 + 17 * w, n) for w in range(workers)]
    with mp.get_context("fork").Pool(workers, initializer=_ref_worker_init) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_ref_worker_step, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker_step, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    have = oracle_py.load_ref("libref_xflann.so") is not None
    fps = workers * n * args.steps / dt
    info = cpu_baseline_info(have, n)
    info.update({"value": fps, "unit": "frames/s", "cores": workers,
                 "sample": "%d worker processes (one per host core), each: %s" % (workers, info["sample"])})
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "stages": STAGES, "frames_per_step": n * workers, "host_workers": workers},
                      "cpu_baseline": info,
                      "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import numpy as np, torch
    import torch.distributed as dist
    import ucoslam_b200

    from ucoslam_b200 import shard
    torch.cuda.set_device(local_rank)
    shard.init("nccl", torch.device("cuda", local_rank))
    ctx = ucoslam_b200.Context(local_rank)      # tracker thread's context: ORB extraction + matching
    # mapper side (own streams): local bundle adjustment.  Two contexts alternate between steps, so the BA of step i (host
    # planner + H2D + cluster-resident kernel + D2H) overlaps the tracking of step i+1, as UcoSLAM's threaded mode lets the
    # mapper lag behind the tracker (mapmanager.cpp:1517); every BA finishes inside the timed region.
    ctx_bas = [ucoslam_b200.Context(local_rank) for _ in range(N_MAPPERS)]
    for c in ctx_bas:
        if BA_CLUSTER:
            c.ba_set_mode(0, BA_CLUSTER)
        if N_MAPPERS > 1:
            c.ba_set_host_threads(1)     # N_MAPPERS batches are already planned side by side: one planner thread per call
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    from concurrent.futures import ThreadPoolExecutor
    mapper = ThreadPoolExecutor(N_MAPPERS)      # UcoSLAM runs local BA in its mapper thread next to tracking (mapmanager.cpp)
    F = args.frames
    prm = ucoslam_b200.OrbParams(KPTS)
    clip = synth_clip(F, shard.unit_seed(1234, rank, 0))   # every rank tracks its own stream of frames (weak scaling)
    clip_pin = torch.from_numpy(clip).pin_memory()
    with torch.cuda.stream(stream):
        clip_dev = clip_pin.to("cuda", non_blocking=True)
        kps_dev = torch.zeros((F, KPTS, 28), dtype=torch.uint8, device="cuda")
        desc_dev = torch.zeros((F, KPTS, 32), dtype=torch.uint8, device="cuda")
        nout_dev = torch.zeros(F, dtype=torch.int32, device="cuda")
        idx_dev = torch.empty((F, KPTS, K_NN), dtype=torch.int32, device="cuda")
        dist_dev = torch.empty((F, KPTS, K_NN), dtype=torch.int32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    kps_host = np.zeros((F, KPTS), ucoslam_b200.KP_DTYPE)
    desc_host = torch.zeros((F, KPTS, 32), dtype=torch.uint8).pin_memory()
    nout_host = np.zeros(F, np.int32)
    idx_host = torch.empty((F, KPTS, K_NN), dtype=torch.int32).pin_memory()
    dist_host = torch.empty((F, KPTS, K_NN), dtype=torch.int32).pin_memory()
    img_ptrs = (ctypes_voidp_array(F))(*[clip_pin[i].data_ptr() for i in range(F)])
    VP = ctypes_voidp_array(F)
    q_ptrs = VP(*[desc_host[f].data_ptr() for f in range(F)])
    t_ptrs = VP(*[desc_host[f - 1].data_ptr() for f in range(F)])
    i_ptrs = VP(*[idx_host[f].data_ptr() for f in range(F)])
    d_ptrs = VP(*[dist_host[f].data_ptr() for f in range(F)])
    nq_h, nt_h = np.zeros(F, np.int32), np.zeros(F, np.int32)
    n_ba = max(1, F // KF_EVERY)
    windows = ba_windows(n_ba, shard.unit_seed(500, rank, 0))
    ba_packs = [c.ba_pack_batch(windows, BA_ITERS) for c in ctx_bas]   # one set of result buffers per mapper context
    ba_packed = ba_packs[0]
    ba_in_bytes = sum(sum(a.nbytes for a in keep.values()) for keep in ba_packed[2])
    ba_out_bytes = sum(sum(v.nbytes for v in o.values()) for o in ba_packed[3])
    ctx.sync()

    def ba_all(m=0):  # host-buffer C-ABI call (there is no device-resident variant: the window is assembled by the host mapper)
        ctx_bas[m].ba_solve_batch(None, BA_ITERS, packed=ba_packs[m])

    def orb_dev():
        ctx.orb_extract_batch_dev(clip_dev.data_ptr(), F, W, H, W, W * H, prm, kps_dev.data_ptr(), desc_dev.data_ptr(),
                                  nout_dev.data_ptr())

    def knn_dev(f):  # frame f against its predecessor (frame 0 against the last one of the batch)
        ctx.hamming_knn_dev(desc_dev[f].data_ptr(), KPTS, desc_dev[f - 1].data_ptr(), KPTS, K_NN,
                            ucoslam_b200.UCO_KNN_HEAP, idx_dev[f].data_ptr(), dist_dev[f].data_ptr())

    def knn_all_dev():
        # frames 1..F-1 against their predecessors in ONE launch (row counts read from the extractor's device-side n_out),
        # frame 0 against the last frame of the batch in a second one
        ctx.hamming_knn_batch_dev(F - 1, desc_dev[1].data_ptr(), KPTS * 32, KPTS, nout_dev[1:].data_ptr(),
                                  desc_dev[0].data_ptr(), KPTS * 32, KPTS, nout_dev.data_ptr(), K_NN,
                                  ucoslam_b200.UCO_KNN_HEAP, idx_dev[1].data_ptr(), dist_dev[1].data_ptr())
        knn_dev(0)

    def track_device():
        orb_dev()
        knn_all_dev()

    def step_device():  # one self-contained step: mapper (BA windows) and tracker (extract + match) side by side
        fut = mapper.submit(ba_all, 0)
        track_device()
        fut.result()

    def run_pipelined(track, n_steps, between=None):
        """n_steps steps; the BA windows of step i go to mapper i % N_MAPPERS and are waited for before that mapper is reused
        (at most N_MAPPERS BA batches in flight) and, for the last ones, before returning."""
        futs = []
        for i in range(n_steps):
            if len(futs) >= N_MAPPERS:
                futs.pop(0).result()
            futs.append(mapper.submit(ba_all, i % N_MAPPERS))
            track()
            if between:
                between()
        for f in futs:
            f.result()

    lib, h = ctx.lib, ctx.h
    import ctypes
    prm_p = ctypes.addressof(prm)

    def track_host():  # reference-facing calls with HOST buffers: H2D + kernels + D2H inside
        rc = lib.uco_b200_orb_extract_batch(h, ctypes.cast(img_ptrs, ctypes.c_void_p), F, W, H, W, prm_p,
                                            kps_host.ctypes.data, desc_host.data_ptr(), KPTS, nout_host.ctypes.data)
        if rc != 0:
            raise RuntimeError(lib.uco_b200_last_error(h))
        # every frame against its predecessor, host descriptor buffers in, host k-NN tables out: one batched call
        nq_h[:] = nout_host
        nt_h[:] = np.roll(nout_host, 1)
        rc = lib.uco_b200_hamming_knn_batch(h, F, ctypes.cast(q_ptrs, ctypes.c_void_p), nq_h.ctypes.data, 32,
                                            ctypes.cast(t_ptrs, ctypes.c_void_p), nt_h.ctypes.data, 32, K_NN, 0,
                                            ctypes.cast(i_ptrs, ctypes.c_void_p), ctypes.cast(d_ptrs, ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError(lib.uco_b200_last_error(h))

    def barrier():
        ctx.sync()
        for c in ctx_bas:
            c.sync()
        torch.cuda.synchronize()
        shard.barrier()

    def reduce_max(v):
        return shard.max_over_ranks(v, "cuda")

    def timed_events(fn, reps, flush_l2=True):
        """sum of per-repetition device durations (CUDA events on the context stream), L2 flushed between repetitions"""
        with torch.cuda.stream(stream):
            evs = []
            for _ in range(reps):
                if flush_l2:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                fn()
                b.record(stream)
                evs.append((a, b))
            barrier()
            return sum(a.elapsed_time(b) for a, b in evs)

    # warm-up (also builds the extractor plan and its buffers), sanity: every frame yields the full keypoint budget
    for m in range(N_MAPPERS):     # every mapper context allocates its arenas / pinned buffers on its first batch: do that here,
        ba_all(m)                  # whatever W is (with W < N_MAPPERS warm-up steps the last contexts would first run inside the timed region)
    with torch.cuda.stream(stream):
        run_pipelined(track_device, max(args.warmup, 3), between=flush.zero_)
    barrier()
    n_kp = nout_dev.cpu().numpy()
    assert (n_kp == KPTS).all(), "synthetic frames must give %d keypoints, got %s" % (KPTS, n_kp[:8])

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    count = lambda: ctx.launch_count() + sum(c.launch_count() for c in ctx_bas)
    n0 = count()
    # EXACTLY args.steps steps between two events on the tracker stream (the second one recorded after the last BA batch has
    # returned its results to the host), barrier + synchronize on both sides, L2 flushed between steps inside the region
    barrier()
    with torch.cuda.stream(stream):
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record(stream)
        run_pipelined(track_device, args.steps, between=flush.zero_)
        ev_b.record(stream)
    barrier()
    ms_dev = reduce_max(ev_a.elapsed_time(ev_b))
    launches = (count() - n0) // max(1, args.steps)

    run_pipelined(track_host, max(1, args.warmup))
    barrier()
    t0 = time.perf_counter()
    run_pipelined(track_host, args.steps)
    barrier()
    e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3)
    clocks = sampler.summary() if sampler else None

    # per-kernel device times of the same step (events at the stage boundaries inside the library)
    ctx.set_profiling(True)
    reps = 5
    acc = {}
    for _ in range(reps):
        with torch.cuda.stream(stream):
            flush.zero_()
            orb_dev()
        for k, v in ctx.orb_last_stage_ms().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    ctx.set_profiling(False)
    knn_ms = timed_events(knn_all_dev, reps) / reps
    ba_call_ms = timed_events(ba_all, reps) / reps          # the host-synchronous C-ABI call: planner + H2D + kernel + D2H
    ba_dev = []
    for _ in range(reps):                                   # the kernel alone: CUDA events around the launch on the mapper's stream
        ba_all()
        ba_dev.append(float(ba_packs[0][3][0]["device_ms"]))
    ba_ms = sum(ba_dev) / len(ba_dev)
    stage_ms = dict(acc)
    stage_ms["hamming_knn"] = knn_ms
    stage_ms["local_ba"] = ba_ms
    n_obs = sum(len(w["obs_pose"]) for w in windows)
    ba_trials = sum(int(o["trace"][:, 1].sum()) for o in ba_packed[3])
    pb = ctx.orb_plan_bytes()
    cand_bytes = 0  # candidate lists are small and L2 resident; not counted as algorithmic traffic
    alg = {  # ALGORITHMIC bytes per frame of each kernel (DESIGN.md "Measurement")
        "blur": 2 * W * H,                                                        # read frame, write level 0
        "resize": 2 * pb["pyramid_px"] - W * H - 179 * 134,                       # read levels 0..6, write levels 1..7
        "fast_cells": pb["pyramid_px"] + cand_bytes,                              # read every level once
        "select": 0,
        "orient_describe": KPTS * (28 + 32),                                      # write keypoints + descriptors
        "hamming_knn": 2 * KPTS * 32 + KPTS * K_NN * 8,
        # compulsory HBM traffic of a window: its inputs in, its results out (everything an LM trial touches, SURVEY.md 8(d)'s
        # 316 B per observation and trial, stays in L2: ncu shows ~1.7 MB of DRAM traffic per window); per frame
        "local_ba": (ba_in_bytes + ba_out_bytes) / F,
    }
    top = max(stage_ms, key=stage_ms.get)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    achieved = alg[top] * F / (stage_ms[top] * 1e-3) / 1e9
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture of the same step (profiles/r1_traffic.json)
    traffic = None
    tk = {"local_ba": "ba_cluster_kernel", "hamming_knn": "hamming_knn_kernel", "fast_cells": "fast_cells_kernel", "blur": "blur7_kernel",
          "select": "select_kernel", "orient_describe": "orient_describe_kernel"}.get(top)
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if tk and os.path.exists(tp) and F == 64 and W == 640:
        for name, v in json.load(open(tp))["kernels"].items():
            if tk in name:
                traffic = v["dram_bytes_per_launch"]
                break
    orb_total_ms = sum(acc.values())
    orb_alg = W * H + 2 * pb["pyramid_px"] + KPTS * 60                          # 2 328 264 B/frame
    if rank == 0:
        total_frames = F * world * args.steps
        line = {"metric": METRIC, "value": total_frames / (ms_dev * 1e-3), "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "stages": STAGES, "frames_per_step_per_gpu": F, "ba_windows_per_step_per_gpu": n_ba,
                           "ba_window": "12 KF (2 fixed), 2000 points, %d observations, nIters=5" % (n_obs // max(1, n_ba)),
                           "parallelism": "frames sharded over %d GPU(s), no collective" % world,
                           "l2": "flushed between timed steps (256 MB write, inside the timed region)",
                           "mapper": "%d BA contexts take turns (clusters of %d CTAs per window): the local BA of a step overlaps the tracking "
                                     "of the following steps (threaded mode: the mapper lags the tracker); all BA results are back on the "
                                     "host inside the timed region" % (N_MAPPERS, BA_CLUSTER or 8)},
                "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": "frames/s",
                        "h2d_bytes_per_step": F * (W * H + 2 * KPTS * 32) + ba_in_bytes,
                        "d2h_bytes_per_step": F * (KPTS * 60 + 4 + KPTS * K_NN * 8) + ba_out_bytes},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved / hbm_peak, "traffic": traffic,
                             "peak_source": "measured" if peaks else "fallback",
                             "kernel_ms_per_step": stage_ms[top], "algorithmic_bytes_per_frame": alg[top]},
                "stage_ms_per_step": stage_ms,
                "stage_note": "tracker stages (blur..hamming_knn) run on one stream, local_ba (ba_cluster_kernel, one launch per %d windows, "
                              "%d LM trials per window) on the mappers' streams in parallel; the host-synchronous C-ABI call around it "
                              "(planner + H2D + launch + D2H) takes %.2f ms" % (n_ba, ba_trials // max(1, n_ba), ba_call_ms),
                "orb_pipeline": {"ms_per_step": orb_total_ms, "algorithmic_bytes_per_frame": orb_alg,
                                 "achieved_gbs": orb_alg * F / (orb_total_ms * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": orb_alg * F / (orb_total_ms * 1e-3) / 1e9 / hbm_peak}}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(clip)
        print(json.dumps(line))
        sys.stdout.flush()
    barrier()
    # orderly teardown (the driver's exit hook records the loaded libraries, so the interpreter must exit normally): every torch
    # object that refers to the tracker context's stream goes first, then the mapper pool, then the contexts, then the process group
    mapper.shutdown(wait=True)
    del stream, clip_dev, kps_dev, desc_dev, nout_dev, idx_dev, dist_dev, flush, clip_pin, desc_host, idx_host, dist_host, ev_a, ev_b
    import gc
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()       # the caching allocators record events on the streams their blocks were used on: drop the blocks
    torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None
    for c in ctx_bas:
        c.close()
    ctx.close()
    shard.finalize()
    sys.stdout.flush()
    sys.stderr.flush()


def ctypes_voidp_array(n):
    import ctypes
    return ctypes.c_void_p * n


def cpu_baseline(clip):
    """Bounded CPU sample of the same workload on this box's host cores (rank 0, N=1)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py, orb_oracle
    have = oracle_py.load_ref("libref_xflann.so") is not None
    n = min(len(clip), 16)
    windows = ba_windows(max(1, n // KF_EVERY), 500)
    cpu_reference_step(clip[:2], oracle_py, orb_oracle, have)  # warm caches
    t0 = time.perf_counter()
    cpu_reference_step(clip[:n], oracle_py, orb_oracle, have, windows)
    dt = time.perf_counter() - t0
    info = cpu_baseline_info(have, n)
    info.update({"value": n / dt, "unit": "frames/s"})
    return info


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
