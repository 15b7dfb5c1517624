#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python scripts/ncu_summary.py launches gpurun_out/launches_r1e.csv > profiles/r1_launches.txt
    python scripts/ncu_summary.py full gpurun_out/prof_*.ncu-rep       > profiles/r1_ncu_full.txt
"""
import csv, subprocess, sys, io, collections, re

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    return name.split("(")[0]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        k = (short(r[4]), r[7], r[8])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# source: %s   launches: %d   total device time: %.3f ms" % (path, len(rows), tot / 1e6))
    print("%-70s %-14s %-16s %6s %12s %10s %7s" % ("kernel", "block", "grid", "count", "total_us", "avg_us", "share"))
    for (k, b, g), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %-14s %-16s %6d %12.1f %10.2f %6.1f%%" % (k[:70], b, g, n, t / 1e3, t / n / 1e3, 100 * t / tot))


def full(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print("== %s" % p)
        for r in rows[2:]:
            print("kernel: %s   grid %s block %s" % (short(r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")],
                                                     r[hdr.index("Block Size")]))
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print("    %-70s %14s %s" % (w, r[i], units[i]))
            if "--first" in sys.argv:
                break  # first captured launch; the others repeat it
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full([a for a in sys.argv[2:] if not a.startswith('--')])
