#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun visit left in gpurun_out/ into the committed summaries under profiles/.

    python scripts/summarize_ncu.py r01            # writes profiles/r01_*.txt|csv
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def launches(tag, name="launches.csv"):
    path = os.path.join(OUT, name)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v * 1000 if r[ui] in ("ms", "msecond") else v
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(PROF, tag + "_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: gpurun_out/%s ; command: see scripts/gpu_round.sh\n" % name)
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-70s n=%5d avg=%10.2f us total=%12.1f us share=%5.1f%%\n" % (k[:70], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    shutil.copy(path, os.path.join(PROF, tag + "_launches.csv"))


def full(tag, rep):
    path = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if not rows:
        return
    hdr = rows[0]
    with open(os.path.join(PROF, "%s_%s_full.txt" % (tag, rep)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source gpurun_out/%s.ncu-rep\n" % rep)
        ni = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
        if ni is not None:
            f.write("kernel: %s\n" % ", ".join(sorted(set(r[ni] for r in rows[2:]))))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write("%-75s %-14s %s\n" % (k, rows[1][i], " | ".join(r[i] for r in rows[2:])))
    with open(os.path.join(PROF, "%s_%s_raw.csv" % (tag, rep)), "w") as f:
        f.write(raw)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    for name in os.listdir(OUT):
        if name.endswith(".ncu-rep"):
            full(tag, name[:-8])
    b = os.path.join(OUT, "bench.json")
    if os.path.exists(b) and os.path.getsize(b):
        shutil.copy(b, os.path.join(PROF, tag + "_bench.json"))
        try:
            d = json.loads(open(b).read().strip().splitlines()[-1])
            print("bench:", d["value"], d["unit"], "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"))
        except Exception as e:
            print("bench.json unreadable:", e)


if __name__ == "__main__":
    main()
