"""GPU-box helper for ncu: a few config-2 blocks (10-stage eq cascade, 256 ch, 4096-frame blocks) through K1."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dsp_b200
import bench
fs, C, F, S = 48000, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 4096, 10
st = torch.cuda.current_stream().cuda_stream
f = [31.25, 62.5, 125, 250, 500, 1000, 2000, 4000, 8000, 16000]; g = [-2, 1.5, -1, 2, -1.5, 1, -2, 1.5, -1, 2]
coefs = np.array([dsp_b200.biquad_design(13, fs, f[i], 1.4, g[i]) for i in range(S)])
ch = dsp_b200.Chain(fs, C).add_biquad(coefs)
blocks = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(2)]
d_out = torch.empty((F, C), dtype=torch.float64, device="cuda")
for i in range(6):
    ch.run_device(0, F, blocks[i % 2].data_ptr(), d_out.data_ptr(), st)
torch.cuda.synchronize()
ch.close()
