#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an ncu report (read here, no GPU):
   python scripts/ncu_lines.py report.ncu-rep kernel_regex [top_n] [launch_index]"""
import csv, subprocess, sys, io, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, agg, launches = None, None, collections.OrderedDict(), 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        idx = {}
        for i, n in enumerate(hdr):
            idx.setdefault(n, i)
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    key = (fname, int(r[0]))
    s = int(r[idx["# Samples"]] or 0)
    st = {n: int(r[idx[n]] or 0) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
    ins = int(r[idx["Instructions Executed"]] or 0)
    if key not in agg:
        agg[key] = [r[1].strip(), 0, collections.Counter(), 0]
    agg[key][1] += s
    agg[key][2].update(st)
    agg[key][3] += ins
tot = sum(v[1] for v in agg.values()) or 1
allst = collections.Counter()
for v in agg.values():
    allst.update(v[2])
print("total samples", tot, "| stall mix:", ", ".join("%s %.0f%%" % (k[6:], 100.0 * n / tot) for k, n in allst.most_common(8)))
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-14s %4d %5.1f%% %-34s %s" % (f, ln, 100.0 * v[1] / tot, ",".join("%s:%d" % (k[6:], n) for k, n in v[2].most_common(3)), v[0][:100]))
