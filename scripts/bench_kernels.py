"""GPU-box helper: device-resident throughput and roofline fractions of the other hot-path kernels
(BASELINE.json configs 2, 4, 5 in their single-GPU form).  Prints one JSON object; bench.py stays
the headline (config H).  Peaks: MEASURED_PEAKS.json hbm_gbs; FP64 peak measured here with an FMA loop
is not available, so the FP64 ceiling quoted is the datasheet 37 TFLOP/s class figure."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dsp_b200
import bench

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
st = torch.cuda.current_stream().cuda_stream
out = {}

def timed(chain, blocks, d_out, F, steps=200, warm=5, names=()):
    for i in range(warm):
        chain.run_device(0, F, blocks[i % len(blocks)].data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    # throughput: plain loop, CUDA events on the launching stream around all of it
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        chain.run_device(0, F, blocks[i % len(blocks)].data_ptr(), d_out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    # per-kernel durations: a second, instrumented pass
    for n in names: dsp_b200.profile_read(n)
    dsp_b200.profile_enable(True)
    for i in range(steps):
        chain.run_device(0, F, blocks[i % len(blocks)].data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    dsp_b200.profile_enable(False)
    prof = {n: dsp_b200.profile_read(n) for n in names}
    return e0.elapsed_time(e1) / steps, prof

# C2: 10-stage eq cascade, 256 ch, 48 kHz, 4096-frame blocks
fs, C, F = 48000, 256, 4096
f = [31.25, 62.5, 125, 250, 500, 1000, 2000, 4000, 8000, 16000]; g = [-2, 1.5, -1, 2, -1.5, 1, -2, 1.5, -1, 2]
coefs = np.array([dsp_b200.biquad_design(13, fs, f[i], 1.4, g[i]) for i in range(10)])
ch = dsp_b200.Chain(fs, C).add_biquad(coefs)
blocks = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(16)]   # 128 MB of distinct input > L2? no: 8 MB each
d_out = torch.empty((F, C), dtype=torch.float64, device="cuda")
ms, prof = timed(ch, blocks, d_out, F, names=("biquad",))
sps = C * F / (ms * 1e-3)
out["C2_biquad10_256ch"] = {"ms_per_block": ms, "Msamples_per_s": sps / 1e6, "bytes_per_sample": 16,
                           "hbm_GBs": sps * 16 / 1e9, "hbm_frac_of_measured": sps * 16 / 1e9 / peak,
                           "kernel_us": {k: v[0] / max(v[1], 1) * 1e3 for k, v in prof.items()}, "fp64_flops_per_sample_min": 90, "note": "one launch per block (k_bq_cascade)"}
ch.close()

# C4: resample 44100 -> 48000, 1024 ch (single GPU form)
fs, C, F = 44100, 1024, 4096
ch = dsp_b200.Chain(fs, C).add_resample(48000)
blocks = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(4)]
d_out = torch.empty((ch.max_out_frames(F) + 8, C), dtype=torch.float64, device="cuda")
ms, prof = timed(ch, blocks, d_out, F, steps=50, names=("resample",))
sps = C * F / (ms * 1e-3)
p = dsp_b200.resample_params(44100, 48000)
flop = 2.0 * p["in_len"] * p["n"] / p["d"]
out["C4_resample_1024ch"] = {"ms_per_block": ms, "Msamples_per_s_in": sps / 1e6, "bytes_per_sample": 8 * (1 + p["n"] / p["d"]),
                            "hbm_frac_of_measured": sps * 8 * (1 + p["n"] / p["d"]) / 1e9 / peak,
                            "fp64_flop_per_in_sample": flop, "fp64_TFLOPs": sps * flop / 1e12}
ch.close()

# C5 (one GPU's share): 8 eq + fir_p 65536 shared IR + resample, 256 ch
fs, C, F = 44100, 256, 4096
coefs = np.array([dsp_b200.biquad_design(13, fs, f[i], 1.4, g[i]) for i in range(8)])
ch = dsp_b200.Chain(fs, C).add_biquad(coefs).add_fir(bench.make_ir(65536, 0), block_hint=F).add_resample(48000)
blocks = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(8)]
d_out = torch.empty((ch.max_out_frames(F) + 8, C), dtype=torch.float64, device="cuda")
ms, prof = timed(ch, blocks, d_out, F, steps=100, names=("biquad", "fir_mac", "fir_level0", "resample"))
out["C5_chain_256ch_per_gpu"] = {"ms_per_block": ms, "Msamples_per_s_in": C * F / (ms * 1e-3) / 1e6,
                                 "kernel_ms_per_block": {k: v[0] / 100 for k, v in prof.items()}, "plan": ch.describe()}
ch.close()

# H with a shared (mono) IR
fs, C, F = 48000, 256, 4096
ch = dsp_b200.Chain(fs, C).add_fir(bench.make_ir(131072, 0), block_hint=F)
blocks = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(8)]
d_out = torch.empty((F, C), dtype=torch.float64, device="cuda")
ms, prof = timed(ch, blocks, d_out, F, steps=200, names=("fir_mac",))
out["H_shared_ir"] = {"ms_per_block": ms, "Msamples_per_s": C * F / (ms * 1e-3) / 1e6, "fir_mac_us": prof["fir_mac"][0] / max(prof["fir_mac"][1], 1) * 1e3}
ch.close()

# C3: fir_p 131072 taps, 64 ch, per-channel IR; and the headline shape at the CLI's default block (2048) and a large one (65536)
for tag, C, F, steps in (("C3_fir_p_64ch_block4096", 64, 4096, 200), ("H_block2048", 256, 2048, 300), ("H_block65536", 256, 65536, 12)):
    ch = dsp_b200.Chain(48000, C).add_fir(bench.make_irs(131072, C), block_hint=F)
    nb = 8 if F <= 4096 else 2
    blocks = [torch.from_numpy(bench.make_block(F, C, i)).cuda() for i in range(nb)]
    d_out = torch.empty((F, C), dtype=torch.float64, device="cuda")
    ms, prof = timed(ch, blocks, d_out, F, steps=steps, warm=3, names=("fir_mac",))
    out[tag] = {"ms_per_block": ms, "Msamples_per_s": C * F / (ms * 1e-3) / 1e6, "plan": ch.describe()}
    ch.close()
    del blocks, d_out
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
