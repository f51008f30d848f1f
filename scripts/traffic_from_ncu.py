#!/usr/bin/env python
"""profiles/traffic.json entry for the headline step from an `ncu --set full` capture of the k_fir_* kernels
(gpurun_out/prof_fir_step.ncu-rep, see scripts/gpu_round.sh): DRAM bytes (read + write) per launch of each
kernel, times its launches per step as counted in the same capture window.

    python scripts/traffic_from_ncu.py [C F taps h [far [stagger]]]      # defaults: 256 4096 131072 1 12 1
"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", os.environ.get("NCU_REP", "prof_fir_step.ncu-rep"))
C, F, taps, h = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 4096, 131072, 1)))
far = int(sys.argv[5]) if len(sys.argv) >= 6 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
units = rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = collections.OrderedDict()
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("dspb200::", "")
    b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
    per.setdefault(name, []).append((b, float(r[ti])))
pipe = any(k.startswith("k_fir_pipe") for k in per)
stagger = int(sys.argv[6]) if len(sys.argv) >= 7 else 1
n_l0 = sum(len(v) for k, v in per.items() if k.startswith("k_fir_pipe" if pipe else "k_fir_level0"))   # one block kernel launch per step
out = {"kernels": {}, "window_steps": n_l0}
total = 0.0
for k, v in per.items():
    avg = sum(b for b, _ in v) / len(v)
    # launches per step by construction (the capture window cuts steps at both ends): the block kernel and the MAC
    # once, a batch tier once when staggered (one channel class per block) and every T-th block otherwise; plan-time
    # kernels (k_fir_fwd: the filter spectra) do not belong to a step
    if k.startswith("k_fir_fwd") or k.startswith("k_fir_inv"):
        lps = 0.0
    elif k.startswith("k_fir_mac_batch") and not stagger:
        lps = 1.0 / int(k.split("<")[1].split(",")[0])
    else:
        lps = 1.0
    out["kernels"][k] = {"dram_bytes_per_launch": avg, "launches_in_window": len(v), "launches_per_step": lps,
                         "avg_us_under_ncu": sum(t for _, t in v) / len(v)}
    total += avg * lps
out["dram_bytes_per_step"] = total
out["source"] = "ncu --set full --clock-control none -k regex:k_fir_ (gpurun_out/prof_fir_step.ncu-rep); dram__bytes_read.sum + dram__bytes_write.sum"
path = os.path.join(ROOT, "profiles", "traffic.json")
tr = json.load(open(path)) if os.path.exists(path) else {}
tr["step:C%d:F%d:taps%d:h%d:pipe%d:far%d" % (C, F, taps, h, 1 if pipe else 0, 0 if pipe else far)] = out
json.dump(tr, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
