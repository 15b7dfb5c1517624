"""DRAM bytes per launch of every kernel of one bench step, from an ncu report of `bench.py --profile-step`:
    ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r2_step python bench.py --profile-step
    python scripts/traffic_json.py gpurun_out/r2_step.ncu-rep > profiles/r2_traffic.json
bench.py reads the file for its `roofline.traffic` fields."""
import csv, io, json, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ix = {n: i for i, n in enumerate(hdr)}
kern = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = re.sub(r"<unnamed>::", "", r[ix["Kernel Name"]]).split("(")[0].replace("void ", "")
    key = "%s %s" % (name, r[ix["Grid Size"]] if "Grid Size" in ix else "")
    def f(n):
        try:
            return float(r[ix[n]])
        except Exception:
            return 0.0
    unit_r, unit_w = rows[1][ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_write.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    b = f("dram__bytes_read.sum") * scale.get(unit_r, 1.0) + f("dram__bytes_write.sum") * scale.get(unit_w, 1.0)
    k = kern.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "ncu_duration_ns": 0.0})
    k["launches"] += 1
    k["dram_bytes"] += b
    k["ncu_duration_ns"] += f("gpu__time_duration.sum") * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(rows[1][ix["gpu__time_duration.sum"]], 1.0)
res = {"source": "ncu --set full --clock-control none over `bench.py --profile-step` (128 frames + 16 BA windows); per-launch averages",
       "kernels": {k: {"launches": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
                       "ncu_duration_s": v["ncu_duration_ns"] / v["launches"] * 1e-9} for k, v in kern.items()}}
print(json.dumps(res, indent=1))
