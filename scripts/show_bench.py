"""print the headline fields of a bench.py JSON line"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
import signal; signal.signal(signal.SIGPIPE, signal.SIG_DFL)
print("value %.0f  ms/step %.3f  e2e %.0f  launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches")))
print("stage_ms", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
print("roofline", d["roofline"]["stage"], round(d["roofline"]["frac"], 3), "| knn frac", round(d["roofline_stages"]["hamming_knn"]["frac"], 3))
for k in ("latency_ms_single_frame", "stream_1280x720", "ate", "error", "tracking", "collective_paths"):
    if k in d:
        print(k, d[k])
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("stage_ms_per_frame"))
