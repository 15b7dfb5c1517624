"""BASELINE config 4, first half: the fbow loop-closure / relocalisation BoW query over a large keyframe database
(KeyFrameDataBase::relocalizationCandidates steps 1-2, kfdb.cu).  One JSON line.

    python scripts/kfdb_bench.py [--kfs 20000] [--words 1800] [--ref]

Database: seeded keyframes over a 10^6-word vocabulary (the shipped orb.fbow has 971 k words) that revisit 200 places; a query is a
new view of one place.  `--ref` also times the reference's own keyframedatabase.cpp (oracle/_ref/libref_kfdb.so) on a bounded sample
of the same database (the first --ref-kfs keyframes) on one host core.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import numpy as np
import ucoslam_b200

ap = argparse.ArgumentParser()
ap.add_argument("--kfs", type=int, default=20000)
ap.add_argument("--words", type=int, default=1800)
ap.add_argument("--queries", type=int, default=50)
ap.add_argument("--ref", action="store_true")
ap.add_argument("--ref-kfs", type=int, default=2000)
args = ap.parse_args()

V = 1_000_000
rng = np.random.default_rng(4)
n_places = 200
places = [np.unique(rng.integers(0, V, args.words)) for _ in range(n_places)]


def view(p):
    base = places[p]
    keep = base[rng.random(len(base)) < 0.7]
    w = np.unique(np.concatenate([keep, rng.integers(0, V, int(args.words * 0.35))])).astype(np.uint32)
    return w, (rng.random(len(w)) * 0.004).astype(np.float32)


bows = [view(i % n_places) for i in range(args.kfs)]
ids = np.arange(args.kfs, dtype=np.uint32)
queries = [view(int(rng.integers(n_places))) for _ in range(args.queries)]
n_words = int(sum(len(b[0]) for b in bows))

ctx = ucoslam_b200.Context(0)
db = ucoslam_b200.KeyFrameDataBase(ctx)
t0 = time.perf_counter()
for a in range(0, args.kfs, 2000):
    db.add_batch(ids[a:a + 2000], bows[a:a + 2000])
add_ms = (time.perf_counter() - t0) * 1e3
ctx.set_profiling(True)
for q in queries[:5]:
    r = db.query(*q)
scan, score, wall, nres = [], [], [], []
for q in queries:
    t = time.perf_counter()
    r = db.query(*q)
    wall.append((time.perf_counter() - t) * 1e3)
    ms = db.last_ms()
    scan.append(float(ms[0])); score.append(float(ms[1])); nres.append(len(r["frame"]))
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = peaks.get("hbm_gbs")
peak_source = "measured" if hbm else "fallback (B200_PROFILING.md)"
hbm = float(hbm or 6650.0)
scan_ms = float(np.median(scan))
achieved = 4.0 * n_words / (scan_ms * 1e-3) / 1e9
line = {"workload": "config4 (BoW half): relocalizationCandidates votes + fBow::score over %d keyframes / %d stored words, query of ~%d words"
                    % (args.kfs, n_words, len(queries[0][0])),
        "n_gpus": 1, "keyframes": args.kfs, "stored_words": n_words, "queries": len(queries),
        "ms_per_query_scan_kernel": scan_ms, "ms_per_query_score_kernel": float(np.median(score)),
        "ms_per_query_host_call": float(np.median(wall)), "queries_per_s_host_call": 1e3 / float(np.median(wall)),
        "scored_frames_per_query": float(np.mean(nres)), "add_batch_ms_total": add_ms,
        "roofline": {"kernel": "kfdb_count_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "peak_source": peak_source, "algorithmic_bytes_per_query": 4 * n_words,
                     "traffic": None}}
if args.ref:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    n = min(args.ref_kfs, args.kfs)
    ref = oracle_py.RefKeyFrameDataBase(oracle_py.REF_VOC_PATH)
    for i in range(n):
        ref.add_bow(ids[i], *bows[i])
    sub = ucoslam_b200.KeyFrameDataBase(ctx)
    sub.add_batch(ids[:n], bows[:n])
    tq, same, tg = [], 0, []
    for q in queries[:20]:
        t = time.perf_counter()
        c = ref.query_bow(*q, True, 0.0)
        tq.append((time.perf_counter() - t) * 1e3)
        t = time.perf_counter()
        g = sub.relocalization_candidates(q[0], q[1], lambda f: [], True, 0.0)
        tg.append((time.perf_counter() - t) * 1e3)
        same += int(np.array_equal(c, g))
    line["cpu_reference"] = {"kind": "reference", "cores": 1, "sample": "the reference's keyframedatabase.cpp on the first %d keyframes, 20 queries" % n,
                             "ms_per_query": float(np.median(tq)), "gpu_ms_per_query_same_sample": float(np.median(tg)),
                             "identical_candidate_lists": "%d/20" % same}
print(json.dumps(line), flush=True)
