"""GPU-box probe: what the host call costs on its own.  A chain holding only `gain` (copy-bound) and the
headline fir chain, synchronous and with blocks in flight, for several slab counts."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dsp_b200
import bench

FS, C, F = 48000, 256, 4096
torch.cuda.init()
blocks = [bench.make_block(F, C, i) for i in range(4)]
pins = [dsp_b200.PinnedArray((F, C), write_combined=True) for _ in range(4)]
pins_nc = [dsp_b200.PinnedArray((F, C)) for _ in range(4)]
for p, q, b in zip(pins, pins_nc, blocks):
    p.array[:] = b; q.array[:] = b
outs = [dsp_b200.PinnedArray((F, C)) for _ in range(5)]

def run(chain, inputs, n=200, depth=0):
    for i in range(5):
        chain.run_raw(F, inputs[i % 4].ptr, outs[0].ptr)
    t0 = time.perf_counter()
    if depth == 0:
        for i in range(n):
            chain.run_raw(F, inputs[i % 4].ptr, outs[0].ptr)
    else:
        tk = []
        for i in range(n):
            _, t = chain.submit_raw(F, inputs[i % 4].ptr, outs[i % (depth + 1)].ptr)
            tk.append(t)
            if i >= depth:
                chain.wait(tk[i - depth])
        chain.sync()
    return (time.perf_counter() - t0) / n * 1e6

res = {}
for slabs in (1, 2, 4, 8, 16):
    g = dsp_b200.Chain(FS, C, slabs_per_device=slabs).add_gain(np.full(C, 0.5))
    res["gain_s%d" % slabs] = {"sync_wc": round(run(g, pins), 1), "sync_plain_pinned": round(run(g, pins_nc), 1),
                               "pipe3_wc": round(run(g, pins, depth=3), 1)}
    g.close()
print(json.dumps(res)); sys.stdout.flush()
irs = bench.make_irs(131072, C)
for slabs in (2, 4, 8):
    ch = dsp_b200.Chain(FS, C, slabs_per_device=slabs).add_fir(irs, block_hint=F)
    res = {"sync_wc": round(run(ch, pins), 1), "pipe1_wc": round(run(ch, pins, depth=1), 1), "pipe3_wc": round(run(ch, pins, depth=3), 1)}
    print(json.dumps({"fir_s%d" % slabs: res})); sys.stdout.flush()
    ch.close()
