"""GPU-box helper: K1 (k_bq_cascade) device-resident time per block for a few shapes.
DSP_B200_BQ_CH=1|2|4 forces the channels-per-CTA variant."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dsp_b200
import bench

st = torch.cuda.current_stream().cuda_stream
res = {}
for (C, F, S) in [(256, 4096, 10), (1024, 4096, 10), (256, 4096, 1), (64, 4096, 10), (256, 512, 10)]:
    fs = 48000
    f = [31.25, 62.5, 125, 250, 500, 1000, 2000, 4000, 8000, 16000]; g = [-2, 1.5, -1, 2, -1.5, 1, -2, 1.5, -1, 2]
    coefs = np.array([dsp_b200.biquad_design(13, fs, f[i], 1.4, g[i]) for i in range(S)])
    ch = dsp_b200.Chain(fs, C).add_biquad(coefs)
    blocks = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(8)]
    d_out = torch.empty((F, C), dtype=torch.float64, device="cuda")
    for i in range(5):
        ch.run_device(0, F, blocks[i % 8].data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 200
    for i in range(n):
        ch.run_device(0, F, blocks[i % 8].data_ptr(), d_out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    res["C%d_F%d_S%d" % (C, F, S)] = {"us_per_block": round(us, 2), "Gsamples_per_s": round(C * F / us / 1e3, 2)}
    ch.close()
print(json.dumps({"BQ_CH": os.environ.get("DSP_B200_BQ_CH", "auto"), **res}))
