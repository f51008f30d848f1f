"""GPU-box helper: how the host-call path (dspb200_chain_run_host) scales with the slab count, and raw PCIe copy times."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dsp_b200
import bench

C, F, TAPS = 256, 4096, 131072
irs = bench.make_irs(TAPS, C)
blocks = [bench.make_block(F, C, i) for i in range(4)]
res = {}
# raw copies
x = torch.empty((F, C), dtype=torch.float64).pin_memory()
d = torch.empty((F, C), dtype=torch.float64, device="cuda")
for name, fn in (("h2d_1d", lambda: d.copy_(x, non_blocking=True)), ("d2h_1d", lambda: x.copy_(d, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): fn()
    torch.cuda.synchronize(); res[name + "_us"] = (time.perf_counter() - t0) / 50 * 1e6
for slabs, wc in ((4, False), (4, True), (6, True), (8, True), (8, False)):
    ch = dsp_b200.Chain(48000, C, devices=[0], slabs_per_device=slabs).add_fir(irs, block_hint=F)
    pins = [dsp_b200.PinnedArray((F, C), write_combined=wc) for _ in range(4)]
    for p, b in zip(pins, blocks): p.array[:] = b
    pout = dsp_b200.PinnedArray((F, C))
    for i in range(6): ch.run_raw(F, pins[i % 4].ptr, pout.ptr)
    t0 = time.perf_counter()
    n = 200
    for i in range(n): ch.run_raw(F, pins[i % 4].ptr, pout.ptr)
    dt = (time.perf_counter() - t0) / n
    res["slabs_%d_wc%d_us" % (slabs, wc)] = dt * 1e6
    res["slabs_%d_wc%d_Msps" % (slabs, wc)] = C * F / dt / 1e6
    ch.close()
print(json.dumps(res, indent=1))
