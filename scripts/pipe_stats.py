#!/usr/bin/env python
"""Per-role cycle counters of the pipeline kernel (DSP_B200_FIR_PIPE_STATS=1): where do the team, the MAC warps and
the producer of each CTA spend the launch?   python scripts/pipe_stats.py [steps]"""
import os, sys
os.environ["DSP_B200_FIR_PIPE_STATS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import dsp_b200
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
C, F = 256, 4096
irs = bench.make_irs(bench.TAPS, C)
ch = dsp_b200.Chain(bench.FS, C).add_fir(irs, block_hint=F)
d_in = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(4)]
d_out = torch.empty((F, C), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for i in range(steps):
    ch.run_device(0, F, d_in[i % 4].data_ptr(), d_out.data_ptr(), st)
torch.cuda.synchronize()
ms, _ = bench.time_device(ch, d_in, d_out.data_ptr(), F, 40, 4, st)
s = ch.debug_read().reshape(-1, 8)[:148].astype(float)
print("plan", ch.describe()[0])
print("step us %.1f" % (ms / 40 * 1e3))
names = ["team_wait", "mac_wait", "prod_wait", "team_total", "mac_total", "prod_total", "stages", "items"]
for k, n in enumerate(names):
    col = s[:, k]
    print("%-11s mean %10.0f  min %10.0f  max %10.0f   (CTA 0: %.0f, CTA 120: %.0f)" % (n, col.mean(), col.min(), col.max(), col[0], col[120]))
print("kernel cycles (max over roles/CTAs): %.0f = %.1f us at 1.965 GHz" % (s[:, 3:6].max(), s[:, 3:6].max() / 1965.0))
ch.close()
