import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dsp_b200, bench
fs, C, F = 48000, 256, 4096
f = [31.25, 62.5, 125, 250, 500, 1000, 2000, 4000, 8000, 16000]; g = [-2, 1.5, -1, 2, -1.5, 1, -2, 1.5, -1, 2]
coefs = np.array([dsp_b200.biquad_design(13, fs, f[i], 1.4, g[i]) for i in range(10)])
ch = dsp_b200.Chain(fs, C).add_biquad(coefs)
d = torch.from_numpy(bench.make_block(F, C, 0)).cuda()
st = torch.cuda.current_stream().cuda_stream
for i in range(12):
    ch.run_device(0, F, d.data_ptr(), d.data_ptr(), st)
torch.cuda.synchronize()
