"""GPU-box helper for ncu: a few config-4 blocks (44100 -> 48000) through K3.  argv[1] = channels (default 1024)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dsp_b200
import bench
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
F = 4096
st = torch.cuda.current_stream().cuda_stream
ch = dsp_b200.Chain(44100, C).add_resample(48000)
blocks = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(2)]
d_out = torch.empty((ch.max_out_frames(F) + 8, C), dtype=torch.float64, device="cuda")
for i in range(6):
    ch.run_device(0, F, blocks[i % 2].data_ptr(), d_out.data_ptr(), st)
torch.cuda.synchronize()
ch.close()
