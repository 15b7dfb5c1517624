#!/usr/bin/env python
"""bench.py — frames/s of UcoSLAM's per-frame tracking hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F] [--res 640x480|1280x720]

A step is one pass of the hot path over F INDEPENDENT synthetic frames (streams / cameras) of BASELINE config 2 (640x480 mono,
2000 ORB keypoints/frame, 8 levels x1.2, local-BA window 10 KF), in the order the reference runs it (src/utils/system.cpp:6460-6960
per frame, src/utils/mapmanager.cpp per keyframe):
  per frame          ORB extraction -> kd-tree of the keypoints -> search by projection from the previous frame -> solvePnp ->
                     Map::matchFrameToMapPoints over the local map -> filter_ambiguous_query -> solvePnp
  per keyframe       (one frame in KF_EVERY = 8) fbow transform of its descriptors (computeBow), FrameMatcher (k-NN + filters)
                     against its 7 neighbouring frames (new-map-point creation), one local bundle adjustment (10 free + 2 fixed
                     keyframes, 2000 points, ~15k observations, the reference's two-stage 5 + 10 LM iterations)
Every stage runs through the C ABI of libucoslam_b200.so (no CPU fallback: the library refuses to create a context without a CUDA
device).
  value   : frames/s with the step's input frames already resident in HBM (CUDA events on the tracker stream, L2 flushed between
            timed steps, max over ranks)
  e2e     : the same work through the host-buffer C-ABI calls (pinned host frames in; keypoints, descriptors, poses, match lists,
            bags of words, BA results out; H2D and D2H copies inside the timed region)
  roofline: the kernel with the largest SM x time share of the step against the roof that bounds it; `roofline_stages` lists every
            stage the same way (INT-ALU issue rate for the extractor, popc rate for the k-NN, FP64 rate of the occupied SMs for BA)
  cpu_baseline / --impl reference : the same step on this box's host cores through the reference's own code where it compiles
            here (g2o, xflann, fbow, picoflann) and the cv2-backed restatement elsewhere (see DESIGN.md "Measurement")
Multi-GPU: frames shard across ranks (independent units, no data-path collective) -> weak scaling.  Under torchrun the line also
carries `collective_paths`: config 4 (row-sharded 10^6-descriptor map, NCCL all-gather) and config 5 (landmark-sharded global BA,
NCCL all-reduce) timed on all ranks.
"""
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))

W, H, KPTS, K_NN, FOCAL = 640, 480, 2000, 10, 525.0
METRIC = "frames/sec (ORB+match+local-BA) 640x480 mono"
WORKLOAD = "config2: 640x480 mono tracking, 2000 ORB kpts/frame, local-BA window=10 KF"
STAGES = ["orb_extract", "kdtree_build", "track_by_projection", "pose_only_lm", "local_map_match", "pose_only_lm",
          "bow_transform(per KF)", "frame_match_vs_7_neighbours(per KF)", "local_ba(per KF)"]
KF_EVERY = 8          # one keyframe (= one computeBow + new-point matching + local-BA call, mapmanager.cpp:4005) per KF_EVERY frames
BA_WINDOW = dict(n_poses=12, n_fixed=2, n_points=2000)   # 10 free KFs + 2 fixed observers, ~15k observations, nIters = 5
BA_ITERS = 5
MAX_DESC_DIST, PROJ_DIST_THR = 50.0, 15.0                          # ORBextractor.h:105, ucoslamtypes.cpp:49
N_MAPPERS = int(os.environ.get("UCO_BENCH_MAPPERS", "4"))          # mapper contexts alternating between steps
BA_CLUSTER = int(os.environ.get("UCO_BENCH_BA_CLUSTER", "4"))       # CTAs per BA cluster (0 = library default 8)
SM_COUNT, SM_GHZ = 148, 1.965
PEAK_ALU = 64 * SM_COUNT * SM_GHZ * 1e9      # alu-pipe thread-instructions/s (16 lanes/clk/SMSP, B300_MICROARCH.md "Pipe rates")
PEAK_POPC = 16 * SM_COUNT * SM_GHZ * 1e9     # popc32/s
PEAK_FP64_SM = 2 * 61.6 * SM_GHZ * 1e9       # DFMA flop/s per SM, MEASURED on this B200: 512 threads x 8 independent DFMA chains issue 61.6 DFMA/clk/SM
                                             # (profiles/r2_fp64_latency_microbench.txt: 66.5 cycles per 8 DFMA per thread); B200 keeps the full-rate FP64 pipe


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=128, help="frames per step per GPU")
    ap.add_argument("--profile-step", action="store_true", help="after the warm-up run ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra sections (720p, single-frame latency, ATE, collective paths)")
    ap.add_argument("--res", default="640x480", choices=["640x480", "1280x720"],
                    help="640x480 / 2000 keypoints = BASELINE config 2 (the metric's configuration, default); 1280x720 / 4000 keypoints = the "
                         "mono part of config 3 (north_star asks for both stream sizes)")
    a = ap.parse_args()
    if a.res == "1280x720":
        set_res_720p()
    return a


def set_res_720p():
    global W, H, KPTS, METRIC, WORKLOAD, FOCAL
    W, H, KPTS, FOCAL = 1280, 720, 4000, 1050.0
    METRIC = "frames/sec (ORB+match+local-BA) 1280x720 mono"
    WORKLOAD = "config3 (mono part): 1280x720 tracking, 4000 ORB kpts/frame, local-BA window=10 KF"


def config_common(n_ba):
    """keys shared by both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "stages": STAGES, "kf_every": KF_EVERY,
            "ba_window": "12 KF (2 fixed), 2000 points, ~15k observations, nIters=5", "ba_windows_per_%d_frames" % KF_EVERY: 1,
            "tracking_problem": "independent per frame: previous frame (2000 keypoints, each with a map point) + local map of ~4000 points + "
                                "pose prior = previous pose; thresholds maxDescDistance=50, projDistThr=15"}


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


def ba_windows(n, seed):
    from ucoslam_b200.synth import synth_ba_problem
    return [synth_ba_problem(seed + i, **BA_WINDOW) for i in range(n)]


def kf_groups(n_frames):
    """(keyframe, its neighbours) per group of KF_EVERY consecutive frames: the last frame of a group is the keyframe"""
    return [(g + KF_EVERY - 1, list(range(g, g + KF_EVERY - 1))) for g in range(0, n_frames - KF_EVERY + 1, KF_EVERY)]


def camera():
    from ucoslam_b200 import workload
    return workload.Camera(W, H, FOCAL)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the same step on the host cores
def cpu_step(imgs, scenes, windows, mods, voc, split):
    """one CPU pass over the frames: per frame extract + the tracker's sequence, per keyframe bag of words + frame matcher against
    its neighbours + local BA.  `split` accumulates seconds per stage."""
    oracle_py, orb_oracle, have_ref = mods
    from ucoslam_b200 import workload
    pnp = oracle_py.ref_pose_only if have_ref else oracle_py.pose_only
    feats = []
    for img, sc in zip(imgs, scenes):
        t0 = time.perf_counter()
        k, d = orb_oracle.extract(img, KPTS)
        t1 = time.perf_counter()
        oracle_py.track_frame(workload.with_current(sc, k, d), MAX_DESC_DIST, PROJ_DIST_THR, pnp=pnp)
        t2 = time.perf_counter()
        split["orb_extract"] += t1 - t0
        split["track_sequence"] += t2 - t1
        feats.append((k, d))
    for (kf, nb), pb in zip(kf_groups(len(imgs)), windows):
        t0 = time.perf_counter()
        if voc is not None:
            voc.transform(feats[kf][1], 3)
        t1 = time.perf_counter()
        r = oracle_py.ref_frame_match_multi(feats[kf][1], feats[kf][0], [feats[i][1] for i in nb], [feats[i][0] for i in nb],
                                            MAX_DESC_DIST * 2, 0.6, True, 1 << 30) if have_ref else None
        if r is None:
            for i in nb:
                oracle_py.frame_match(feats[i][1], feats[i][0], feats[kf][1], feats[kf][0], MAX_DESC_DIST * 2, 0.6, True, 1 << 30)
        t2 = time.perf_counter()
        if oracle_py.ref_ba_optimize(pb, BA_ITERS) is None:
            oracle_py.ba_optimize(pb, BA_ITERS)
        t3 = time.perf_counter()
        split["bow_transform"] += t1 - t0
        split["frame_match"] += t2 - t1
        split["local_ba"] += t3 - t2


def cpu_info(have_ref, n):
    return {"cores": 1, "kind": "port",
            "sample": "%d frames + %d keyframes: ORB = Python/cv2-4.13 restatement of ORBextractor.cpp (blur/resize/FAST native OpenCV, 1 "
                      "thread; interpreter overhead included - the reference's C++ extractor cannot be built here, expect it to be "
                      "several times cheaper than this arm's ORB share); tracker sequence = C/C++ restatements of system.cpp / map.cpp on "
                      "the reference's picoflann tree order with solvePnp by %s; per keyframe: fbow transform by %s, FrameMatcher = %s, "
                      "local BA = %s; 1 thread as the reference runs each of them" % (
                          n, max(1, n // KF_EVERY), "the reference's g2o + typesg2o.h" if have_ref else "the C restatement",
                          "the reference's fbow" if have_ref else "skipped (oracle/_ref absent)",
                          "the reference's xflann HKMeans(32,0) built once + 16-check search per neighbour + restated filters" if have_ref
                          else "exact linear port + restated filters",
                          "the reference's g2o compiled from its sources (oracle/_ref)" if have_ref else "the C restatement")}


_REF = {}


def _ref_mods():
    if "mods" not in _REF:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py, orb_oracle
        try:
            import cv2
            cv2.setNumThreads(1)          # one worker per host core: no nested thread pools
        except Exception:
            pass
        have = all(oracle_py.load_ref(n) is not None for n in ("libref_xflann.so", "libref_g2o.so", "libref_fbow.so"))
        _REF["mods"] = (oracle_py, orb_oracle, have)
    return _REF["mods"]


def _ref_data(seed, n):
    key = ("data", seed, n)
    if key not in _REF:               # every worker tracks its own streams (generated once, outside the timed steps)
        oracle_py, orb_oracle, have = _ref_mods()
        from ucoslam_b200 import workload
        imgs, scenes, _ = workload.make_problems(n, lambda im: orb_oracle.extract(im, KPTS), seed, camera())
        voc = None
        if have:
            if "vocb" not in _REF:
                _REF["vocb"] = workload.synth_vocabulary_full()
            voc = oracle_py.RefVocabulary(_REF["vocb"])
        _REF[key] = (imgs, scenes, ba_windows(max(1, n // KF_EVERY), 500 + seed), voc)
    return _REF[key]


def _ref_worker_step(job):
    seed, n, res = job
    if res == "1280x720" and W != 1280:
        set_res_720p()
    mods = _ref_mods()
    imgs, scenes, windows, voc = _ref_data(seed, n)
    split = dict.fromkeys(("orb_extract", "track_sequence", "bow_transform", "frame_match", "local_ba"), 0.0)
    cpu_step(imgs, scenes, windows, mods, voc, split)
    return split


def run_reference(args, rank, world):
    """The reference's CPU path on ALL host cores of the box: one worker process per core, each tracking its own streams
    (the same sharding by independent units the GPU arm uses); a step = every worker processes its bounded sample."""
    if rank != 0:
        return
    import multiprocessing as mp
    workers = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    n = KF_EVERY
    jobs = [(1234 + 17 * w, n, args.res) for w in range(workers)]
    with mp.get_context("fork").Pool(workers) as pool:
        for _ in range(max(1, min(args.warmup, 2))):
            pool.map(_ref_worker_step, jobs, chunksize=1)
        t0 = time.perf_counter()
        tot = {}
        for _ in range(args.steps):
            for sp in pool.map(_ref_worker_step, jobs, chunksize=1):
                for k, v in sp.items():
                    tot[k] = tot.get(k, 0.0) + v
        dt = time.perf_counter() - t0
    have = _ref_mods()[2]
    fps = workers * n * args.steps / dt
    info = cpu_info(have, n)
    ssum = sum(tot.values()) or 1.0
    info.update({"value": fps, "unit": "frames/s", "cores": workers,
                 "sample": "%d worker processes (one per host core), each: %s" % (workers, info["sample"]),
                 "stage_share": {k: v / ssum for k, v in tot.items()},
                 "stage_ms_per_frame_per_core": {k: v / (workers * n * args.steps) * 1e3 for k, v in tot.items()}})
    cfg = config_common(1)          # identical on both arms: the workload; how an arm runs it is in "arm"
    arm = {"frames_per_step_per_unit": n, "units": "%d host worker processes" % workers}
    print(json.dumps({"impl": "reference", "arm": arm, "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                      "config": cfg, "cpu_baseline": info,
                      "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(n):
    """Bounded CPU sample of the same workload on this box's host cores (rank 0, N=1), one thread."""
    mods = _ref_mods()
    imgs, scenes, windows, voc = _ref_data(1234, n)
    split = dict.fromkeys(("orb_extract", "track_sequence", "bow_transform", "frame_match", "local_ba"), 0.0)
    cpu_step(imgs[:2], scenes[:2], [], mods, voc, dict(split))  # warm caches
    t0 = time.perf_counter()
    cpu_step(imgs, scenes, windows, mods, voc, split)
    dt = time.perf_counter() - t0
    info = cpu_info(mods[2], n)
    info.update({"value": n / dt, "unit": "frames/s", "stage_ms_per_frame": {k: v / n * 1e3 for k, v in split.items()}})
    return info


# ---------------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank, world):
    """Multi-GPU runs on one host: keep a rank's threads (and with them its first-touched pinned buffers) on the NUMA node its GPU hangs off, so that
    the 72 MB of copies per step do not cross the socket link.  Returns a note for the bench line; does nothing when the topology is not exposed."""
    if world < 2 or os.environ.get("UCO_BENCH_NO_NUMA"):
        return "not bound (single rank)" if world < 2 else "not bound (UCO_BENCH_NO_NUMA)"
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "not bound (numa_node of %s is -1)" % bus
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "not bound (node %d has no allowed cpu)" % node
        os.sched_setaffinity(0, cpus)
        return "rank threads bound to NUMA node %d of GPU %s (%d cpus)" % (node, bus, len(cpus))
    except Exception as e:      # the benchmark must run wherever the topology files are missing
        return "not bound (%s)" % type(e).__name__


def run_b200(args, rank, world, local_rank):
    import ctypes, gc
    import numpy as np, torch
    import ucoslam_b200
    from ucoslam_b200 import shard, workload, chain
    from concurrent.futures import ThreadPoolExecutor

    numa_note = bind_to_gpu_numa_node(local_rank, world)
    torch.cuda.set_device(local_rank)
    shard.init("nccl", torch.device("cuda", local_rank))
    ctx = ucoslam_b200.Context(local_rank)      # tracker thread's context: extraction, tracking sequence, BoW, frame matcher
    # mapper side (own streams): local bundle adjustment.  N_MAPPERS contexts take turns, so the BA of step i (host planner + H2D +
    # cluster-resident kernel + D2H) overlaps the tracking of the following steps, as UcoSLAM's threaded mode lets the mapper lag
    # behind the tracker (mapmanager.cpp:1517); every BA finishes inside the timed region.
    ctx_bas = [ucoslam_b200.Context(local_rank) for _ in range(N_MAPPERS)]
    for c in ctx_bas:
        if BA_CLUSTER:
            c.ba_set_mode(0, BA_CLUSTER)
        if N_MAPPERS > 1:
            c.ba_set_host_threads(1)     # N_MAPPERS batches are already planned side by side: one planner thread per call
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    mapper = ThreadPoolExecutor(N_MAPPERS)
    lib, h = ctx.lib, ctx.h
    VP = lambda n: ctypes.c_void_p * n

    class Work:
        """everything one resolution needs: problems, device-resident state, buffers, the step functions"""

        def __init__(self, F, w, hgt, kpts, seed, c=None):
            c = c or ctx
            self.c, self.lib, self.hc = c, c.lib, c.h                      # the tracker context this work runs on
            self.F, self.w, self.h, self.kpts = F, w, hgt, kpts
            self.prm = ucoslam_b200.OrbParams(kpts)
            cam = workload.Camera(w, hgt, FOCAL if w == W else (1050.0 if w == 1280 else 525.0))
            ext = lambda im: ctx.orb_extract(im, self.prm)
            self.imgs, self.scenes, self.gt = workload.make_problems(F, ext, seed, cam)
            self.tprm = ucoslam_b200.TrackParams(self.scenes[0], MAX_DESC_DIST, PROJ_DIST_THR)
            pc = max(len(s["prev_octave"]) for s in self.scenes)
            mc = max(len(s["mp_id"]) for s in self.scenes)
            self.state = ucoslam_b200.TrackState(c, F, pc, mc)
            for f, s in enumerate(self.scenes):
                self.state.set_scene(f, s)
            self.prior = np.stack([np.asarray(s["pose44"], np.float32).reshape(16) for s in self.scenes])
            self.groups = kf_groups(F)
            self.mprm = ucoslam_b200.MatchParams(MAX_DESC_DIST * 2, 0.6, True, 1 << 30)     # mapmanager.cpp:9990
            self.clip_pin = torch.from_numpy(self.imgs).pin_memory()
            with torch.cuda.stream(stream):
                self.clip_dev = self.clip_pin.to("cuda", non_blocking=True)
                z = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
                self.kps_dev, self.desc_dev, self.nout_dev = z((F, kpts, 28), torch.uint8), z((F, kpts, 32), torch.uint8), z(F, torch.int32)
                self.prior_dev = torch.from_numpy(self.prior).to("cuda")
                self.m_dev, self.nm_dev, self.pose_dev = z((F, kpts, 16), torch.uint8), z(F, torch.int32), z((F, 16), torch.float32)
                self.good_dev, self.stat_dev, self.tbp_dev = z(F, torch.int32), z(F, torch.int32), z(F, torch.int32)
                nkf, npairs = len(self.groups), sum(len(g[1]) for g in self.groups)
                self.word_dev, self.wgt_dev, self.node_dev = z((nkf, kpts), torch.int32), z((nkf, kpts), torch.float32), z((nkf, kpts), torch.int32)
                self.fm_dev, self.fmn_dev = z((npairs, kpts, 16), torch.uint8), z(npairs, torch.int32)
            self.tout = ucoslam_b200.TrackOut(self.m_dev.data_ptr(), self.nm_dev.data_ptr(), self.pose_dev.data_ptr(), self.good_dev.data_ptr(),
                                              self.stat_dev.data_ptr(), self.tbp_dev.data_ptr(), None)
            # host-side buffers of the e2e path
            self.img_ptrs = VP(F)(*[self.clip_pin[i].data_ptr() for i in range(F)])
            self.kps_pin = torch.zeros((F, kpts, 28), dtype=torch.uint8).pin_memory()      # page-locked result buffers: direct D2H
            self.desc_pin = torch.zeros((F, kpts, 32), dtype=torch.uint8).pin_memory()
            self.kps_h = self.kps_pin.numpy().view(ucoslam_b200.KP_DTYPE).reshape(F, kpts)
            self.desc_h = self.desc_pin.numpy()
            self.nkp_h = np.zeros(F, np.int32)
            self.o_h = dict(matches=np.zeros((F, kpts), ucoslam_b200.MATCH_DTYPE), n_matches=np.zeros(F, np.int32), pose=np.zeros((F, 16), np.float32),
                            n_good=np.zeros(F, np.int32), status=np.zeros(F, np.int32), n_tbp=np.zeros(F, np.int32))
            self.tout_h = ucoslam_b200.TrackOut(*[self.o_h[k].ctypes.data for k in ("matches", "n_matches", "pose", "n_good", "status", "n_tbp")], None)
            self.kf_idx = np.array([g[0] for g in self.groups], np.int32)
            self.nb_ptr = np.zeros(nkf + 1, np.int32)
            self.nb_ptr[1:] = np.cumsum([len(g[1]) for g in self.groups])
            self.nb_idx = np.array([f for g in self.groups for f in g[1]], np.int32)
            self.bow_h = (np.zeros((nkf, kpts), np.uint32), np.zeros((nkf, kpts), np.float32), np.zeros((nkf, kpts), np.uint32))
            self.fm_h, self.fmn_h = np.zeros((npairs, kpts), ucoslam_b200.MATCH_DTYPE), np.zeros(npairs, np.int32)

        # -- device-resident step (tracker stream) --
        def orb_dev(self):
            self.c.orb_extract_batch_dev(self.clip_dev.data_ptr(), self.F, self.w, self.h, self.w, self.w * self.h, self.prm,
                                      self.kps_dev.data_ptr(), self.desc_dev.data_ptr(), self.nout_dev.data_ptr())

        def track_dev(self):
            rc = self.lib.uco_b200_track_state_step_dev(self.hc, self.state.h, self.kps_dev.data_ptr(), self.desc_dev.data_ptr(), self.nout_dev.data_ptr(),
                                                   self.kpts, self.prior_dev.data_ptr(), ctypes.addressof(self.tprm), ctypes.addressof(self.tout),
                                                   ucoslam_b200.UCO_TRACK_NO_SYNC)
            if rc != 0:
                raise RuntimeError(self.lib.uco_b200_last_error(self.hc))

        def keyframes_dev(self, with_voc=True, with_match=True):
            """per keyframe: bag of words + FrameMatcher against its neighbours, on the resident frames: three launches"""
            rc = self.lib.uco_b200_keyframes_batch_dev(self.hc, voc if with_voc else None, 3, self.kps_dev.data_ptr(), self.kpts, self.desc_dev.data_ptr(),
                                                  self.kpts * 32, self.nout_dev.data_ptr(), self.kpts, self.F, len(self.groups),
                                                  self.kf_idx.ctypes.data, (self.nb_ptr if with_match else np.zeros_like(self.nb_ptr)).ctypes.data,
                                                  self.nb_idx.ctypes.data, None, ctypes.addressof(self.mprm), self.word_dev.data_ptr(),
                                                  self.wgt_dev.data_ptr(), self.node_dev.data_ptr(), self.fm_dev.data_ptr(), self.fmn_dev.data_ptr())
            if rc != 0:
                raise RuntimeError(self.lib.uco_b200_last_error(self.hc))

        def step_dev(self):
            self.orb_dev()
            self.track_dev()
            self.keyframes_dev()

        # -- the same through host buffers --
        def step_host(self):
            rc = self.lib.uco_b200_track_frames(self.hc, self.state.h, ctypes.cast(self.img_ptrs, ctypes.c_void_p), self.w, self.h, self.w,
                                           ctypes.addressof(self.prm), ctypes.addressof(self.tprm), self.prior.ctypes.data, self.kps_h.ctypes.data,
                                           self.desc_h.ctypes.data, self.nkp_h.ctypes.data, ctypes.addressof(self.tout_h))
            if rc != 0:
                raise RuntimeError(self.lib.uco_b200_last_error(self.hc))
            rc = self.lib.uco_b200_keyframes_batch(self.hc, voc, 3, len(self.groups), self.kf_idx.ctypes.data, self.nb_ptr.ctypes.data, self.nb_idx.ctypes.data,
                                              None, ctypes.addressof(self.mprm), self.bow_h[0].ctypes.data, self.bow_h[1].ctypes.data,
                                              self.bow_h[2].ctypes.data, self.fm_h.ctypes.data, self.fmn_h.ctypes.data)
            if rc != 0:
                raise RuntimeError(self.lib.uco_b200_last_error(self.hc))

        def h2d_bytes(self):
            return self.F * (self.w * self.h + 64) + 8 * (len(self.groups) + len(self.nb_idx))

        def d2h_bytes(self):
            return self.F * (self.kpts * (28 + 32 + 16) + 64 + 24) + len(self.groups) * self.kpts * 12 + len(self.nb_idx) * (self.kpts * 16 + 4)

        def close(self):
            self.state.close()
            for k in list(self.__dict__):
                if k.endswith("_dev") or k.endswith("_pin") or k in ("kps_h", "desc_h"):
                    delattr(self, k)

    F = args.frames
    voc_bytes = workload.synth_vocabulary_full()
    voc = ctx.bow_load(voc_bytes)
    wk = Work(F, W, H, KPTS, shard.unit_seed(1234, rank, 0))   # every rank tracks its own streams (weak scaling)
    with torch.cuda.stream(stream):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    n_ba = max(1, F // KF_EVERY)
    windows = ba_windows(n_ba, shard.unit_seed(500, rank, 0))
    ba_packs = [c.ba_pack_batch(windows, BA_ITERS) for c in ctx_bas]   # one set of result buffers per mapper context
    ba_in_bytes = sum(sum(a.nbytes for a in keep.values()) for keep in ba_packs[0][2])
    ba_out_bytes = sum(sum(v.nbytes for v in o.values()) for o in ba_packs[0][3])
    ctx.sync()

    def ba_all(m=0):  # host-buffer C-ABI call (the window is assembled by the host mapper: there is no device-resident variant)
        ctx_bas[m].ba_solve_batch(None, BA_ITERS, packed=ba_packs[m])

    # end-to-end path: TWO tracker contexts take turns (a second camera stream's worth of state), so that the uploads / downloads of
    # one step overlap the kernels of the other; every step still uploads its own frames and downloads its own results
    ctx2 = ucoslam_b200.Context(local_rank)
    wk2 = Work(F, W, H, KPTS, shard.unit_seed(1234, rank, 0), ctx2)
    trackers = ThreadPoolExecutor(2)

    def run_e2e(n_steps):
        futs, tf, wks = [], [None, None], [wk, wk2]
        for i in range(n_steps):
            if len(futs) >= N_MAPPERS:
                futs.pop(0).result()
            futs.append(mapper.submit(ba_all, i % N_MAPPERS))
            if tf[i % 2] is not None:
                tf[i % 2].result()
            tf[i % 2] = trackers.submit(wks[i % 2].step_host)
        for f in tf + futs:
            if f is not None:
                f.result()

    stream2 = torch.cuda.ExternalStream(ctx2.stream, device=local_rank)

    def run_two_streams(n_steps):
        """device-resident steps alternate between the two tracker contexts' streams (the small per-frame kernels of one step overlap
        the wide ones of the other); the L2 flush follows every step on its own stream; BA batches as in run_pipelined"""
        futs, wks, sts = [], [wk, wk2], [stream, stream2]
        for i in range(n_steps):
            if len(futs) >= N_MAPPERS:
                futs.pop(0).result()
            futs.append(mapper.submit(ba_all, i % N_MAPPERS))
            with torch.cuda.stream(sts[i % 2]):
                wks[i % 2].step_dev()
                flush.zero_()
        for f in futs:
            f.result()

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

    def barrier():
        ctx.sync()
        for c in ctx_bas:
            c.sync()
        torch.cuda.synchronize()
        shard.barrier()

    def reduce_max(v):
        return shard.max_over_ranks(v, "cuda")

    def timed_events(fn, reps, flush_l2=True):
        """mean device duration of fn (CUDA events on the tracker stream), L2 flushed before every repetition"""
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
            return sum(a.elapsed_time(b) for a, b in evs) / reps

    evs = {}

    def finish():
        nonlocal stream, stream2, flush, wk, wk2
        # orderly teardown (the driver's exit hook records the loaded libraries, so the interpreter must exit normally): every torch
        # object that refers to the tracker context's stream goes first, then the mapper pool, then the contexts, then the process group
        mapper.shutdown(wait=True)
        trackers.shutdown(wait=True)
        wk2.close()
        wk.close()
        ctx.bow_free(voc)
        evs.clear()
        stream = stream2 = flush = wk = wk2 = None
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()       # the caching allocators record events on the streams their blocks were used on: drop the blocks
        if hasattr(torch._C, "_host_emptyCache"):
            torch._C._host_emptyCache()
        for c in ctx_bas:
            c.close()
        ctx2.close()
        ctx.close()
        shard.finalize()
        sys.stdout.flush()
        sys.stderr.flush()

    # warm-up (also builds the extractor plan and every workspace), sanity: every frame yields the full keypoint budget and tracks
    for m in range(N_MAPPERS):     # every mapper context allocates its arenas / pinned buffers on its first batch: do that here
        ba_all(m)
    with torch.cuda.stream(stream):
        run_pipelined(wk.step_dev, max(args.warmup, 3), between=flush.zero_)
    barrier()
    run_two_streams(max(args.warmup, 3))
    barrier()
    ctx2.sync()
    n_kp = wk.nout_dev.cpu().numpy()
    assert (n_kp == KPTS).all(), "synthetic frames must give %d keypoints, got %s" % (KPTS, n_kp[:8])
    good = wk.good_dev.cpu().numpy()
    assert (good > 300).all() and (wk.stat_dev.cpu().numpy() == 0).all(), "every synthetic frame must track: inliers %s" % good[:8]
    pose_err = float(np.abs(wk.pose_dev.cpu().numpy().reshape(F, 4, 4) - wk.gt.astype(np.float32)).max())

    if args.profile_step:      # ncu --profile-from-start off: one step of every stage (tracker stream + one BA batch), nothing else
        torch.cuda.cudart().cudaProfilerStart()
        with torch.cuda.stream(stream):
            wk.step_dev()
        ba_all(0)
        barrier()
        torch.cuda.cudart().cudaProfilerStop()
        if rank == 0:
            print(json.dumps({"profile_step": True, "frames": F, "ba_windows": n_ba}))
        finish()
        return
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    count = lambda: ctx.launch_count() + ctx2.launch_count() + sum(c.launch_count() for c in ctx_bas)
    n0 = count()
    # EXACTLY args.steps steps between two events on the tracker stream (the second one recorded after the last BA batch has
    # returned its results to the host), barrier + synchronize on both sides, L2 flushed between steps inside the region
    barrier()
    ev_a, ev_b, ev_j = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
    evs.update(a=ev_a, b=ev_b, j=ev_j)
    ev_a.record(stream)
    stream2.wait_event(ev_a)                 # the second tracker stream starts inside the region ...
    run_two_streams(args.steps)
    ev_j.record(stream2)
    stream.wait_event(ev_j)                  # ... and has finished before the closing event
    ev_b.record(stream)
    barrier()
    ctx2.sync()
    ms_dev = reduce_max(ev_a.elapsed_time(ev_b))
    del ev_a, ev_b, ev_j
    launches = (count() - n0) // max(1, args.steps)

    run_e2e(max(2, args.warmup))
    barrier()
    ctx2.sync()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    barrier()
    ctx2.sync()
    e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3)
    clocks = sampler.summary() if sampler else None
    assert np.array_equal(wk.o_h["n_good"], good), "host-buffer path and device-resident path disagree"
    assert args.steps < 2 or np.array_equal(wk2.o_h["n_good"], good), "the second tracker context disagrees"

    # ---- per-stage device times of the same step (events at the stage boundaries) ----
    reps = 5
    ctx.set_profiling(True)
    acc = {}
    for _ in range(reps):
        with torch.cuda.stream(stream):
            flush.zero_()
            wk.orb_dev()
        for k, v in ctx.orb_last_stage_ms().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    ctx.set_profiling(False)
    stage_ms = dict(acc)
    nodes_dev = torch.zeros((F, 2 * (KPTS // 5) + 2, 28), dtype=torch.uint8, device="cuda")
    leaf_dev = torch.zeros((F, KPTS), dtype=torch.int32, device="cuda")
    bbox_dev = torch.zeros((F, 4), dtype=torch.float64, device="cuda")
    nn_dev = torch.zeros(F, dtype=torch.int32, device="cuda")
    stage_ms["kdtree_build"] = timed_events(lambda: ctx._chk(lib.uco_b200_kdtree_build_batch_dev(
        h, F, wk.kps_dev.data_ptr(), KPTS, wk.nout_dev.data_ptr(), KPTS, nodes_dev.data_ptr(), nodes_dev.shape[1], leaf_dev.data_ptr(),
        bbox_dev.data_ptr(), nn_dev.data_ptr())), reps)
    stage_ms["track_sequence"] = timed_events(wk.track_dev, reps)          # kd-trees + tbp + solvePnp + local map + solvePnp
    stage_ms["bow_transform"] = timed_events(lambda: wk.keyframes_dev(True, False), reps)
    stage_ms["frame_match"] = timed_events(lambda: wk.keyframes_dev(False, True), reps)   # k-NN + filters, 7 pairs per keyframe, one launch each
    knn_pairs = len(wk.groups) * (KF_EVERY - 1)
    idx_dev = torch.empty((knn_pairs, KPTS, K_NN), dtype=torch.int32, device="cuda")
    dist_dev = torch.empty_like(idx_dev)

    def knn_only():      # the same number of 2000 x 2000 pairs in one launch (every frame against its successor)
        ctx.hamming_knn_batch_dev(knn_pairs, wk.desc_dev[0].data_ptr(), KPTS * 32, KPTS, wk.nout_dev.data_ptr(), wk.desc_dev[1].data_ptr(), KPTS * 32,
                                  KPTS, wk.nout_dev[1:].data_ptr(), K_NN, ucoslam_b200.UCO_KNN_HEAP, idx_dev.data_ptr(), dist_dev.data_ptr())
    stage_ms["hamming_knn"] = timed_events(knn_only, reps)                 # the k-NN kernel inside frame_match
    ba_call_ms = timed_events(ba_all, reps)                                # the host-synchronous C-ABI call: planner + H2D + kernel + D2H
    ba_dev = []
    for _ in range(reps):                                                  # the kernel alone: CUDA events around the launch on the mapper's stream
        ba_all()
        ba_dev.append(float(ba_packs[0][3][0]["device_ms"]))
    stage_ms["local_ba"] = sum(ba_dev) / len(ba_dev)
    del nodes_dev, leaf_dev, bbox_dev, nn_dev, idx_dev, dist_dev
    ba_trials = sum(int(o["trace"][:, 1].sum()) for o in ba_packs[0][3])
    n_obs = sum(len(w["obs_pose"]) for w in windows)
    pb = ctx.orb_plan_bytes()

    # ---- rooflines: every stage against the roof that bounds it; SM x time decides which one is "dominant" ----
    ba_sms = n_ba * (BA_CLUSTER or 8)
    px = pb["pyramid_px"]
    int_ops = {  # algorithmic integer operations per frame: blur 2 x 7 MAC separable, bicubic 2 x 4 taps + border,
        # FAST 16 ring compares + score for the ~3 % that pass, selection ~ per candidate, descriptor 512 gathers + 1024 f32 mul per keypoint
        "blur": 28 * W * H, "resize": 16 * (px - W * H), "fast_cells": 40 * px, "select": 64 * 4 * KPTS, "orient_describe": (709 * 2 + 2560) * KPTS}
    hbm_bytes = {"blur": 2 * W * H, "resize": 2 * px - W * H, "fast_cells": px, "select": 0, "orient_describe": KPTS * 60}
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tp) and F == 128 and W == 640:
        traffic = json.load(open(tp)).get("kernels", {})

    def traffic_of(kernel):
        for name, v in traffic.items():
            if kernel in name:
                return v.get("dram_bytes_per_launch")
        return None

    rl = {}
    for k in ("blur", "resize", "fast_cells", "select", "orient_describe"):
        ms = stage_ms.get(k, 0.0)
        if ms <= 0:
            continue
        ach = int_ops[k] * F / (ms * 1e-3)
        rl[k] = {"kernel": {"blur": "blur7_strip_kernel", "resize": "resize_cubic_strip_kernel x7", "fast_cells": "fast_cells_kernel", "select": "select_kernel",
                            "orient_describe": "orient_describe_kernel"}[k], "bound": "int_alu", "achieved": ach / 1e12, "peak": PEAK_ALU / 1e12,
                 "unit": "T int-op/s", "frac": ach / PEAK_ALU, "ms_per_step": ms, "sm_ms": ms * SM_COUNT,
                 "hbm_gbs": hbm_bytes[k] * F / (ms * 1e-3) / 1e9, "hbm_frac": hbm_bytes[k] * F / (ms * 1e-3) / 1e9 / hbm_peak,
                 "traffic": traffic_of({"blur": "blur7", "resize": "resize_cubic", "fast_cells": "fast_cells", "select": "select_kernel",
                                        "orient_describe": "orient_describe"}[k])}
    ms = stage_ms["hamming_knn"]
    popc = knn_pairs * KPTS * KPTS * 8.0 / (ms * 1e-3)
    rl["hamming_knn"] = {"kernel": "hamming_knn_lq_kernel", "bound": "int_popc", "achieved": popc / 1e12, "peak": PEAK_POPC / 1e12, "unit": "T popc32/s",
                         "frac": popc / PEAK_POPC, "ms_per_step": ms, "sm_ms": ms * SM_COUNT, "traffic": traffic_of("hamming_knn")}
    ms = stage_ms["local_ba"]
    # f64 flops of a window: per LM trial linearize ~400 / observation + Schur ~ (k(k+1)/2) x 216 per landmark (k = obs per landmark) + solve (6P)^3/3
    k_obs = n_obs / max(1, sum(len(w["points3"]) for w in windows))
    flop_trial = n_obs * 400.0 + sum(len(w["points3"]) for w in windows) * (k_obs * (k_obs + 1) / 2) * 216.0 + n_ba * (60.0 ** 3) / 3
    ba_flops = flop_trial * (ba_trials / max(1, n_ba)) / (ms * 1e-3)
    rl["local_ba"] = {"kernel": "ba_cluster_kernel", "bound": "fp64 (latency chain on %d of %d SMs)" % (ba_sms, SM_COUNT), "achieved": ba_flops / 1e12,
                      "peak": PEAK_FP64_SM * ba_sms / 1e12, "unit": "TFLOP/s f64", "frac": ba_flops / (PEAK_FP64_SM * ba_sms), "ms_per_launch": ms,
                      "ms_per_step": ms * ba_sms / SM_COUNT, "sm_ms": ms * ba_sms, "traffic": traffic_of("ba_cluster"),
                      "note": "runs on the mappers' streams beside the tracker: ms_per_step is its SM-share of a step"}
    for k in ("kdtree_build", "track_sequence", "bow_transform"):
        rl[k] = {"bound": "latency (L2-resident, a few hundred flops per unit)", "ms_per_step": stage_ms[k], "sm_ms": stage_ms[k] * SM_COUNT,
                 "achieved": None, "peak": None, "frac": None, "unit": None, "traffic": None}
    filt_ms = max(0.0, stage_ms["frame_match"] - stage_ms["hamming_knn"])
    rl["match_filters"] = {"bound": "latency", "ms_per_step": filt_ms, "sm_ms": filt_ms * SM_COUNT, "achieved": None, "peak": None, "frac": None,
                           "unit": None, "traffic": None}
    top = max((k for k in rl if rl[k]["frac"] is not None), key=lambda k: rl[k]["sm_ms"])
    roof = dict(rl[top])
    roof.update({"stage": top, "peak_source": "INT / popc: B300_MICROARCH.md pipe rates x 148 SMs x 1.965 GHz; FP64: DFMA rate measured on this B200 (profiles/r2_fp64_latency_microbench.txt); HBM " + ("measured" if peaks else "fallback"),
                 "kernel_ms_per_step": rl[top]["ms_per_step"], "dominant_by": "SM x time share of the step"})
    orb_total_ms = sum(acc.values())
    orb_alg = W * H + 2 * px + KPTS * 60                          # 2 328 264 B/frame

    # ---- extras (outside the timed region, bounded) ----
    extras = {}
    if not args.no_extras:
        try:
            extras.update(run_extras(args, rank, world, local_rank, ctx, stream, wk, Work, flush, barrier, reduce_max, run_pipelined,
                                     timed_events, voc))
        except Exception as e:      # an extra must never cost the headline line
            extras["error"] = repr(e)
    if rank == 0:
        total_frames = F * world * args.steps
        cfg = config_common(1)      # identical on both arms: the workload; how this arm runs it is in "arm"
        arm = ({"frames_per_step_per_unit": F, "units": "%d GPU(s)" % world, "parallelism": "frames sharded over %d GPU(s), no collective" % world,
                    "l2": "flushed between timed steps (256 MB write, inside the timed region)",
                    "streams": "two tracker contexts (streams) take alternate steps, the mappers run beside them",
                    "state": "previous frames + map blocks are device resident (uco_b200_track_state); a step's host inputs are its frames and pose priors",
                    "e2e": "two tracker contexts take turns through the host-buffer calls, so one step's copies overlap the other's kernels; every step uploads its frames and downloads its results",
                    "mapper": "%d BA contexts take turns (clusters of %d CTAs per window): the local BA of a step overlaps the tracking "
                              "of the following steps (threaded mode: the mapper lags the tracker); all BA results are back on the "
                              "host inside the timed region" % (N_MAPPERS, BA_CLUSTER or 8)})
        arm["frames_per_step_per_gpu"] = F
        arm["numa"] = numa_note
        line = {"arm": arm, "metric": METRIC, "value": total_frames / (ms_dev * 1e-3), "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": cfg,
                "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": "frames/s",
                        "h2d_bytes_per_step": wk.h2d_bytes() + ba_in_bytes, "d2h_bytes_per_step": wk.d2h_bytes() + ba_out_bytes},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": roof, "roofline_stages": rl,
                "stage_ms_per_step": stage_ms,
                "stage_note": "tracker stages run on one stream in the order extract (blur..orient_describe) -> track_sequence (contains "
                              "kdtree_build) -> bow_transform -> frame_match (contains hamming_knn); local_ba (ba_cluster_kernel, one launch "
                              "per %d windows, %d LM trials per window) runs on the mappers' streams in parallel; the host-synchronous C-ABI call "
                              "around it (planner + H2D + launch + D2H) takes %.2f ms" % (n_ba, ba_trials // max(1, n_ba), ba_call_ms),
                "tracking": {"min_inliers": int(good.min()), "mean_inliers": float(good.mean()), "mean_tbp_matches": float(wk.tbp_dev.cpu().numpy().mean()),
                             "mean_matches": float(wk.nm_dev.cpu().numpy().mean()), "max_abs_pose_entry_error_vs_ground_truth": pose_err},
                "orb_pipeline": {"ms_per_step": orb_total_ms, "algorithmic_bytes_per_frame": orb_alg,
                                 "achieved_gbs": orb_alg * F / (orb_total_ms * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": orb_alg * F / (orb_total_ms * 1e-3) / 1e9 / hbm_peak,
                                 "int_ops_per_frame": sum(int_ops.values()),
                                 "frac_of_int_alu_peak": sum(int_ops.values()) * F / (orb_total_ms * 1e-3) / PEAK_ALU}}
        line.update(extras)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(2 * KF_EVERY)
        print(json.dumps(line))
        sys.stdout.flush()
    barrier()
    finish()


def run_extras(args, rank, world, local_rank, ctx, stream, wk, Work, flush, barrier, reduce_max, run_pipelined, timed_events, voc):
    """single-frame latency, the 1280x720 stream, ATE of a sequential clip, and (under torchrun) the two collective paths"""
    import numpy as np, torch
    import ucoslam_b200
    from ucoslam_b200 import shard, chain, workload
    out = {}
    # -- single-frame latency through the host-buffer calls (the reference's caller issues one frame per call, frameextractor.cpp:3505)
    one = Work(1, W, H, KPTS, shard.unit_seed(77, rank, 0))
    lat, lat_orb = [], []
    kps1 = np.zeros(KPTS, ucoslam_b200.KP_DTYPE); d1 = np.zeros((KPTS, 32), np.uint8); n1 = np.zeros(1, np.int32)
    import ctypes
    for i in range(25):
        t0 = time.perf_counter()
        rc = ctx.lib.uco_b200_track_frames(ctx.h, one.state.h, ctypes.cast(one.img_ptrs, ctypes.c_void_p), W, H, W, ctypes.addressof(one.prm),
                                           ctypes.addressof(one.tprm), one.prior.ctypes.data, one.kps_h.ctypes.data, one.desc_h.ctypes.data,
                                           one.nkp_h.ctypes.data, ctypes.addressof(one.tout_h))
        t1 = time.perf_counter()
        ctx.lib.uco_b200_orb_extract(ctx.h, one.clip_pin[0].data_ptr(), W, H, W, ctypes.addressof(one.prm), kps1.ctypes.data, d1.ctypes.data,
                                     KPTS, n1.ctypes.data)
        t2 = time.perf_counter()
        if rc != 0:
            raise RuntimeError(ctx.lib.uco_b200_last_error(ctx.h))
        if i >= 5:
            lat.append((t1 - t0) * 1e3); lat_orb.append((t2 - t1) * 1e3)
    out["latency_ms_single_frame"] = {"extract_and_track_host_call_median": float(np.median(lat)), "orb_extract_host_call_median": float(np.median(lat_orb)),
                                      "note": "one 640x480 frame per call through host buffers (upload + kernels + download + sync), median of 20"}
    one.close()
    # -- ATE of a sequential clip through the CUDA path (BASELINE metric: "ATE vs ref"); the CPU arm's ATE over the same clip is
    # computed by tests/test_chain_ate_gpu.py and scripts/ate_check.py (identical trajectories: bit-exact matching, 1e-12 LM)
    n_ate = 40
    tex = chain.texture()
    gt = np.array([chain.gt_pose(i) for i in range(n_ate)])
    frames = [chain.render(tex, T) for T in gt]
    prm = ucoslam_b200.OrbParams(2000)
    poses, stats = chain.track_full(frames, gt[0], lambda im: ctx.orb_extract(im, prm), lambda sc: ctx.track_batch([sc])[0])
    out["ate"] = {"ate_m_cuda_path": chain.ate(poses, gt), "frames": n_ate, "min_inliers": int(min(s[2] for s in stats)),
                  "sequence": "extract -> search by projection -> solvePnp -> local-map search -> solvePnp, state carried from frame to frame",
                  "reference_path": "profiles/r2_ate_chain.json holds the CPU-oracle trajectory of the same clip (identical poses)"}
    # -- the 1280x720 / 4000-keypoint stream (north_star asks for both sizes), device-resident value only
    if W == 640:
        F2 = 32
        w2 = Work(F2, 1280, 720, 4000, shard.unit_seed(4321, rank, 0))
        with torch.cuda.stream(stream):
            for _ in range(3):
                w2.step_dev(); flush.zero_()
        barrier()
        ms = timed_events(w2.step_dev, 5)
        ms = reduce_max(ms)
        g2 = w2.good_dev.cpu().numpy()
        out["stream_1280x720"] = {"frames_per_s_device_resident": F2 * world / (ms * 1e-3), "ms_per_step": ms, "frames_per_step_per_gpu": F2,
                                  "kpts": 4000, "min_inliers": int(g2.min()),
                                  "note": "tracker stream only (extract + tracking sequence + BoW + frame matcher per keyframe), no BA overlap"}
        w2.close()
        del w2
    # -- collective paths under torchrun: config 4 all-gather, config 5 all-reduce
    if world > 1:
        comm = shard.make_comm(ctx, "cuda")
        coll = {}
        nt, nq, k = 1_000_000, 2000, 10
        rng = np.random.default_rng(1234)
        t_h = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        q_h = t_h[rng.integers(0, nt, nq)].copy()
        q_h[:, :2] ^= 0x5A
        t, q = torch.from_numpy(t_h).cuda(), torch.from_numpy(q_h).cuda()
        b, e = shard.shard_range(nt, rank, world)
        sh_i = torch.empty((nq, k), dtype=torch.int32, device="cuda"); sh_d = torch.empty_like(sh_i)
        fn = lambda: ctx.hamming_knn_sharded_dev(comm, q.data_ptr(), nq, t[b:].data_ptr(), e - b, b, k, sh_i.data_ptr(), sh_d.data_ptr())
        fn(); barrier()
        ms = reduce_max(timed_events(fn, 10, flush_l2=False))
        coll["config4_map_query"] = {"workload": "2000 queries x 10^6-row map, k=10, rows sharded over %d GPUs, NCCL all-gather of per-shard top-k + merge" % world,
                                     "ms": ms, "queries_per_s": nq / (ms * 1e-3), "popc32_per_s": nq * nt * 8.0 / (ms * 1e-3),
                                     "frac_of_aggregate_popc_peak": nq * nt * 8.0 / (ms * 1e-3) / (PEAK_POPC * world)}
        del t, q, sh_i, sh_d
        from ucoslam_b200.synth import synth_global_ba
        pb = synth_global_ba(42)
        for _ in range(2):
            o = ctx.ba_solve_sharded(pb, 5, comm=comm)
        barrier()
        dev = []
        for _ in range(3):
            o = ctx.ba_solve_sharded(pb, 5, comm=comm)
            dev.append(o["device_ms"])
        ms = reduce_max(float(np.mean(dev)))
        coll["config5_global_ba"] = {"workload": "global BA 500 KF / 50k landmarks / %d observations, landmarks sharded over %d GPUs, NCCL all-reduce of the packed "
                                                 "reduced Hessian per LM trial" % (len(pb["obs_pose"]), world), "ms_per_solve": ms,
                                     "lm_trials": int(o["trace"][:, 1].sum()), "allreduce_bytes_per_trial": float(o["profile"][3])}
        ctx.comm_destroy(comm)
        out["collective_paths"] = coll
    return out


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
