#!/usr/bin/env python
"""bench.py -- headline benchmark of the effects-chain hot path on B200.

Metric (BASELINE.json): Msamples/s through a 256-channel x 131072-tap `fir_p` chain, 1 sample =
one double of one channel-frame at the chain input; weak scaling (every GPU owns 256 channels,
independent streams, no collective on the data path).  One "step" = one block of --block frames
(default 4096) x 256 channels through the chain.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0.  `value` = device-resident throughput (inputs in HBM, CUDA events
on the launching stream); `e2e` = the same blocks through the C-ABI host call
(dspb200_chain_run_host) from pinned host memory, copies inside the timed region; `roofline` =
the partition-MAC kernel (k_fir_mac) against the measured HBM peak; `cpu_baseline` = the compiled
reference (oracle/_ref) on this box's host cores over a bounded sample.
"""
import argparse
import math
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 48000
CHANNELS_PER_GPU = 256
TAPS = 131072
METRIC = "Msamples/s through 256-ch 128k-tap fir_p chain"
UNIT = "Msamples/s"


# ------------------------------------------------------------------------------------------------
# workload definition: seeded decaying-noise IR per channel, noise-like input blocks
# ------------------------------------------------------------------------------------------------
def park_miller(seed, count):
    """x -> 48271 x mod (2^31-1) (the reference's pm_rand1_r, util.h:127-148), vectorised."""
    M = np.uint64(0x7fffffff)
    pw = np.array([48271], dtype=np.uint64)
    while pw.shape[0] < count:
        pw = np.concatenate([pw, (pw * pw[-1]) % M])
    return ((pw[:count] * np.uint64(seed)) % M).astype(np.float64)


def make_ir(taps, channel):
    u = 2.0 * park_miller(1 + channel, taps) / 2147483647.0 - 1.0
    h = u * np.exp(-6.9 * np.arange(taps) / taps)
    return h * (0.5 / np.sum(np.abs(h)))


def make_irs(taps, channels, first_channel=0):
    return np.stack([make_ir(taps, first_channel + c) for c in range(channels)], axis=1)


def make_block(frames, channels, seed):
    u = park_miller(1000 + seed, frames * channels) / 2147483647.0 - 0.5
    return u.reshape(frames, channels)


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU legs: the compiled reference on host cores (the only place oracle/ is executed here)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    ir_path, ch, block, warm, steps, seconds = args
    sys.path.insert(0, ROOT)
    from oracle import ref          # checker / baseline only
    c = ref.RefChain("fir_p -t pcm -e double -c %d -r %d %s" % (ch, FS, ir_path), FS, ch)
    c.run_inplace(block, refill=True)
    for _ in range(warm):
        c.run_inplace(block)
    t0 = time.perf_counter()
    n = 0
    while (steps and n < steps) or (not steps and time.perf_counter() - t0 < seconds):
        c.run_inplace(block)
        n += 1
    dt = time.perf_counter() - t0
    c.close()
    return n, dt


def _cpu_trial(ir_path, procs, ch_per_proc, block, warm, steps, seconds):
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(ir_path, ch_per_proc, block, warm, steps, seconds)] * procs)
    # every process ran concurrently; job throughput = sum of the per-process rates
    rate = sum(n * block * ch_per_proc / dt for n, dt in res)
    return rate, sum(n for n, _ in res), max(dt for _, dt in res), max(n for n, _ in res)


def cpu_reference_run(block, warm, steps=0, seconds=10.0, ch_per_proc=2, procs=None):
    """The reference on the host cores: P processes, each a reference fir_p chain of `ch_per_proc` channels with
    its own 131072-tap IRs (channels are independent, so this is how the reference would use the box; each process
    also runs fir_p's own worker threads, fir_p.c:105-114).  P is tuned over {n/4, n/2, n} cores first and the best
    is kept, so the baseline is not handicapped by oversubscription."""
    from oracle import ref
    if not ref.available():
        return None
    ncpu = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="dspb200_bench_")
    ir_path = os.path.join(tmp, "ir.f64")
    make_irs(TAPS, ch_per_proc).astype("<f8").tofile(ir_path)
    try:
        if procs is None:
            cands = sorted(set(max(1, ncpu // d) for d in (4, 2, 1)))
        else:
            cands = [procs]
        best = None
        for p in cands:
            rate, _, _, _ = _cpu_trial(ir_path, p, ch_per_proc, block, 2, 0, 2.0)
            if best is None or rate > best[0]:
                best = (rate, p)
        procs = best[1]
        per_step = 1
        if steps:
            # fir_p defers most of its arithmetic to worker threads with a two-period deadline (fir_p.c:105-114,
            # 407): a handful of blocks is over before that work has been done once.  One "step" of this arm is
            # therefore a sample of `per_step` blocks per process, sized so that the K timed steps cover >= 5 s.
            blk_rate = best[0] / (procs * ch_per_proc * block)          # blocks per second and process
            per_step = max(1, int(math.ceil(5.0 * blk_rate / steps)))
        rate, blocks, wall, nmax = _cpu_trial(ir_path, procs, ch_per_proc, block, warm * per_step, steps * per_step, seconds)
    finally:
        os.remove(ir_path)
        os.rmdir(tmp)
    return {"value": rate / 1e6, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": "%d processes (of %d host cores; best of n/4, n/2, n) x %d ch x %d-frame blocks, %d blocks in %.1f s%s; "
                      "FFT backend = oracle/fftw3_shim.c (FFTW3 absent)" % (procs, ncpu, ch_per_proc, block, blocks, wall,
                                                                            (" (1 step = %d blocks per process)" % per_step) if steps else ""),
            "ms_per_step": wall / max(1, nmax // per_step) * 1e3, "steps_done": nmax // per_step, "channels": procs * ch_per_proc}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--block", type=int, default=4096)
    ap.add_argument("--channels", type=int, default=CHANNELS_PER_GPU, help="channels per GPU")
    ap.add_argument("--taps", type=int, default=TAPS)
    ap.add_argument("--shared-ir", action="store_true", help="one IR for all channels (-c 1) instead of one per channel")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-slabs", type=int, default=4)
    a = ap.parse_args()

    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) is sent to stderr, the line itself goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warm = max(1, a.steps), max(3, a.warmup)
    config = {"workload": "fir_p %d taps x %d ch per GPU, %s IR, %d-frame blocks, fs %d" %
                          (a.taps, a.channels, "shared" if a.shared_ir else "per-channel", a.block, FS),
              "block_frames": a.block, "channels_per_gpu": a.channels, "taps": a.taps,
              "parallelism": "channel-sharded x%d, no collective" % max(world, 1)}

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(a.block, warm, steps=steps)
        if r is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref/libdspref.so not built (needs /root/reference at build time)"})
            return 0
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
                "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import dsp_b200
    from dsp_b200.dist import Job
    if dsp_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    job = Job()                      # NCCL process group when WORLD_SIZE > 1: barrier + scalar reductions only
    barrier, reduce_max, reduce_sum = job.barrier, job.reduce_max, job.reduce_sum

    C, F = a.channels, a.block
    irs = make_ir(a.taps, 0)[:, None] if a.shared_ir else make_irs(a.taps, C, first_channel=rank * C)
    n_pool = 8
    blocks = [make_block(F, C, rank * 100 + i) for i in range(n_pool)]

    # ---------------- mode D: device-resident blocks, CUDA events on the launching stream ------------
    chain = dsp_b200.Chain(FS, C, devices=[local_rank]).add_fir(irs, block_hint=F)
    d_blocks = [torch.from_numpy(b).cuda() for b in blocks]
    d_out = torch.empty((F, C), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(warm):
        chain.run_device(0, F, d_blocks[i % n_pool].data_ptr(), d_out.data_ptr(), stream)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dsp_b200.profile_read("fir_mac")
    dsp_b200.profile_enable(True)
    launches0 = dsp_b200.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        chain.run_device(0, F, d_blocks[i % n_pool].data_ptr(), d_out.data_ptr(), stream)
    e1.record()
    barrier()
    launches = dsp_b200.kernel_launches() - launches0
    dsp_b200.profile_enable(False)
    ms = reduce_max(e0.elapsed_time(e1))
    total_samples = float(world) * C * F * steps
    value = total_samples / (ms * 1e-3) / 1e6
    checksum = float(d_out.abs().sum().item())

    # roofline of the dominant kernel = the MAC of the LAST partition level (it carries almost all taps).
    # Algorithmic bytes of one launch: per selected channel and bin, P FDL rows + (per-channel IR)
    # P filter rows read and 1 row written, 16 B each.
    plan = [op for op in chain.describe() if op.get("op") == "fir"][0]
    lvl = plan["levels"][-1]
    h = 0 if a.shared_ir else 1
    # tail_pf = partitions of the last level summed inside its fused FFT kernel (0: no tail machinery, the plain
    # three-kernel path; 1: an upper level carries the tail; 2: single level, the kernel sums partitions 0 and 1)
    pf = int(plan.get("tail_pf", 0))
    tb = int(plan.get("t_batch", 0))
    batch_bytes = 0.0
    if pf == 0:
        mac_parts = lvl["P"]
        mac_bytes = C * lvl["B"] * 16.0 * (mac_parts * (1 + h) + 1)
    elif tb:
        # time-batched tail: every period k_fir_mac streams T partitions plus the batched spectrum (read) and writes Y;
        # every T periods k_fir_mac_batch streams the other FDL rows and filter rows once for T outputs
        mac_parts = tb
        mac_bytes = C * lvl["B"] * 16.0 * (tb * (1 + h) + 2)
        batch_bytes = C * lvl["B"] * 16.0 * ((lvl["P"] - pf - 1) + (lvl["P"] - pf - tb) * h + tb)
    else:
        mac_parts = lvl["P"] - pf
        mac_bytes = C * lvl["B"] * 16.0 * (mac_parts * (1 + h) + 1)
    tail_per_sample = (mac_bytes + (batch_bytes / tb if tb else 0.0)) / (C * lvl["B"])
    step_bytes = tail_per_sample * C * F + sum(C * L["B"] * 16.0 * (min(L["P"], 2 if i == 0 else 1) * (1 + h)) * (F / L["B"])
                                               for i, L in enumerate(plan["levels"]))

    def step_budget():
        """algorithmic HBM bytes per input sample of every kernel of a step: total and per kernel"""
        n_lv = len(plan["levels"])
        parts = {"stash_unstash": 0.0, "fir_level0": 0.0, "fir_fwd_inv": 0.0, "fir_mac": mac_bytes / (C * lvl["B"]),
                 "fir_mac_batch": (batch_bytes / tb if tb else 0.0) / (C * lvl["B"])}
        direct = n_lv == 1 and pf != 0 or (n_lv == 1 and lvl["P"] <= 2)     # fused kernel reads/writes the interleaved block itself
        if not direct:
            parts["stash_unstash"] = 16.0 + (8.0 + 8.0 * (n_lv - 1) + 8.0)   # stash (read + write), unstash (y + pending sums + write)
        for i, L in enumerate(plan["levels"]):
            in_kernel = min(L["P"], 2 if i == 0 else 1)
            has_init = pf and i == n_lv - 1
            if pf == 0 and i == 0 and L["P"] > 2:
                parts["fir_fwd_inv"] += (8 + 16) + (16 + 8 + 8 + 8)           # separate forward and inverse transforms
                continue
            parts["fir_level0"] += 8 + 16 + 16 * (in_kernel - 1) + 16 * in_kernel * h + (16 if has_init else 0) + 8 + 8 + 8
        return sum(parts.values()), parts

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "step:C%d:F%d:taps%d:h%d" % (C, F, a.taps, h)
        traffic = tr.get(key, {}).get("dram_bytes_per_step")
    except Exception:
        pass
    roofline = None
    per_sample, parts = step_budget()
    step_s = ms / steps * 1e-3
    in_loop = {}
    for nme in ("fir_level0", "fir_mac", "fir_mac_batch", "fir_fwd", "fir_inv", "fir_mac_head", "fir_mac_bulk"):
        t_ms, n_l = dsp_b200.profile_read(nme)
        if n_l:
            in_loop[nme] = (t_ms / n_l * 1e3, n_l / float(steps))
    if step_s > 0:
        ach = per_sample * C * F / step_s / 1e9
        roofline = {"bound": "hbm",
                    "kernel": "one step = every kernel of the chain for one block (%s), on three concurrent streams" % ", ".join("k_" + k for k in in_loop),
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "algorithmic_bytes_per_step": per_sample * C * F, "algorithmic_bytes_per_sample": per_sample,
                    "bytes_per_sample_by_kernel": parts, "avg_step_us": step_s * 1e6,
                    "partition_levels": plan["levels"], "t_batch": tb, "tail_pf": pf,
                    "note": "achieved = algorithmic bytes of all kernels of a step / measured step time (CUDA events over the timed loop). "
                            "The kernels of a step run concurrently on three streams and time-slice the GPU, so a single kernel's "
                            "event-bracketed duration inside the loop is not its own speed; `kernels` gives each kernel timed alone "
                            "(side streams folded into one) next to its in-loop figure."}
    # isolated kernel durations: a short extra pass with the side streams folded into the caller's stream
    iso = {}
    dsp_b200.debug_serialize(True)
    for i in range(2 * 8):
        chain.run_device(0, F, d_blocks[i % n_pool].data_ptr(), d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    names = ("fir_mac", "fir_mac_batch", "fir_level0", "fir_inv", "fir_fwd", "fir_mac_bulk")
    for nme in names:
        dsp_b200.profile_read(nme)
    dsp_b200.profile_enable(True)
    for i in range(8 * 8):
        chain.run_device(0, F, d_blocks[i % n_pool].data_ptr(), d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    dsp_b200.profile_enable(False)
    dsp_b200.debug_serialize(False)
    for nme in names:
        t_ms, n_l = dsp_b200.profile_read(nme)
        if n_l:
            iso[nme] = t_ms / n_l * 1e3
    if roofline is not None:
        alg = {"fir_level0": parts["fir_level0"] * C * plan["levels"][0]["B"] if len(plan["levels"]) == 1 else None,
               "fir_mac": mac_bytes, "fir_mac_batch": batch_bytes if tb else None}
        kern = {}
        for nme in sorted(set(list(iso) + list(in_loop))):
            e = {"launches_per_step": in_loop.get(nme, (None, 0.0))[1], "in_step_event_us": in_loop.get(nme, (None, 0.0))[0],
                 "alone_us": iso.get(nme), "algorithmic_bytes_per_launch": alg.get(nme)}
            if e["alone_us"] and e["algorithmic_bytes_per_launch"]:
                e["alone_GBs"] = e["algorithmic_bytes_per_launch"] / (e["alone_us"] * 1e-6) / 1e9
                e["alone_frac"] = e["alone_GBs"] / peak
            kern["k_" + nme] = e
        roofline["kernels"] = kern
    chain.close()
    del chain

    # ---------------- mode A (e2e): the C-ABI host call on pinned host blocks ------------------------
    e2e = None
    if not a.no_e2e:
        ch2 = dsp_b200.Chain(FS, C, devices=[local_rank], slabs_per_device=a.e2e_slabs).add_fir(irs, block_hint=F)
        pins = [dsp_b200.PinnedArray((F, C), write_combined=True) for _ in range(n_pool)]   # input-only buffers
        for p, b in zip(pins, blocks):
            p.array[:] = b
        pout = dsp_b200.PinnedArray((F, C))
        for i in range(warm):
            ch2.run_raw(F, pins[i % n_pool].ptr, pout.ptr)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            ch2.run_raw(F, pins[i % n_pool].ptr, pout.ptr)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        dt = reduce_max(dt)
        e2e = {"value": total_samples / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": F * C * 8, "d2h_bytes_per_step": F * C * 8,
               "ms_per_step": dt / steps * 1e3, "api": "dspb200_chain_run_host, pinned host buffers (inputs write-combined), %d channel slabs" % a.e2e_slabs,
               "checksum": float(np.abs(pout.array).sum())}
        # the same blocks through submit/wait: up to `depth` blocks in flight, each with its own output buffer
        depth = 3
        pouts = [dsp_b200.PinnedArray((F, C)) for _ in range(depth + 1)]
        tickets, t_submit, acc = [], 0.0, 0.0
        for i in range(warm):
            ch2.run_raw(F, pins[i % n_pool].ptr, pouts[0].ptr)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            ts = time.perf_counter()
            _, t = ch2.submit_raw(F, pins[i % n_pool].ptr, pouts[i % (depth + 1)].ptr)
            t_submit += time.perf_counter() - ts
            tickets.append(t)
            if i >= depth:
                ch2.wait(tickets[i - depth])
                acc += float(pouts[(i - depth) % (depth + 1)].array[0, 0])     # the block is on the host now
        for t in tickets[-depth:]:
            ch2.wait(t)
        dtp = time.perf_counter() - t0
        barrier()
        dtp = reduce_max(dtp)
        e2e["pipelined"] = {"value": total_samples / dtp / 1e6, "unit": UNIT, "ms_per_step": dtp / steps * 1e3,
                            "host_submit_ms_per_step": t_submit / steps * 1e3, "blocks_in_flight": depth,
                            "api": "dspb200_chain_submit_host + dspb200_chain_wait (same copies, same kernels; the synchronous call above is what the drop-in effect->run() uses)",
                            "checksum": float(np.abs(pouts[(steps - 1) % (depth + 1)].array).sum())}
        ch2.close()

    clocks = sampler.stop()          # sampled across both timed loops (device-resident and host-call)
    launches_total = int(reduce_sum(float(launches)))

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        r = cpu_reference_run(F, 2, seconds=a.cpu_seconds)
        if r:
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        config["l2"] = "per-step working set %.2f GB of FDL+filter spectra streamed from HBM (> 126 MB L2); %d rotating input blocks" % (step_bytes / 1e9, n_pool)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches_total, "roofline": roofline, "cpu_baseline": cpu, "checksum": checksum}
        emit(line)
    job.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
