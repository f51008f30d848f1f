"""Regenerates tests/golden/*.npz from the compiled reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py

Each fixture stores the chain string, stream format, block size, the INPUT (float64) and the
reference's concatenated OUTPUT plus per-call frame counts, so the GPU box needs neither
/root/reference nor the compiled reference to check parity.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref, restate  # noqa: E402


def save(name, chain, fs, channels, block, x, files=None):
    c = ref.RefChain(chain, fs, channels, dir=HERE)
    y, counts = c.process(x, block)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), chain=chain, fs=fs, channels=channels, block=block,
                        x=x, y=y, counts=np.array(counts), out_fs=c.fs_out, effects=np.array(c.effect_names()))
    print(name, chain, x.shape, "->", y.shape, c.effect_names())


def main():
    rng = np.random.default_rng(20260924)
    # C1: plumbing
    x = ref.sgen("sine+1000S", 48000, 2, 1000)
    save("gain", "gain -6", 48000, 2, 256, x)
    # biquad cascade, selector on one stage
    x = ref.sgen("sine:freq=20-20k+1500S", 48000, 4, 1500) * 0.5
    x[:, 1] *= -0.7
    x[:, 3] = rng.standard_normal(1500) * 0.1
    save("biquad", "eq 31.25 1.4 -2 eq 1k 1.0 3 :0,2 lowshelf 200 0.7 4 : highpass 20 0.7 allpass_1 500", 48000, 4, 300, x)
    # fir_p, mono IR from a coefs: literal would be long; use a raw f64 file next to the fixtures
    h = restate.bench_ir(700)
    h.astype("<f8").tofile(os.path.join(HERE, "ir700.f64"))
    x = ref.sgen("sine:freq=20-20k+2000S", 48000, 3, 2000) * 0.9
    x[:, 2] = rng.standard_normal(2000) * 0.2
    save("fir_p", ":0,2 fir_p -t pcm -e double -c 1 -r 48000 ir700.f64", 48000, 3, 256, x)
    # per-channel IR (2 selected channels -> 2 filter columns)
    h2 = np.stack([restate.bench_ir(300, 1), restate.bench_ir(300, 2)], axis=1)
    h2.astype("<f8").tofile(os.path.join(HERE, "ir300x2.f64"))
    save("fir_p_2ch", ":0,2 fir_p -t pcm -e double -c 2 -r 48000 ir300x2.f64", 48000, 3, 1000, x)
    # fir (FFT path, latency len; the chain appends align) and the <=16-tap direct form
    save("fir", "fir -t pcm -e double -c 1 -r 48000 ir700.f64", 48000, 3, 512, x)
    save("fir_direct", "fir coefs:0.5,0.25,-0.125,0.0625,0.03", 48000, 3, 100, x)
    save("hilbert", "hilbert -p 255", 48000, 3, 333, x)
    # -a alignment (fir_util.c:187-205 -> channel_offsets -> the chain's align pass, effects_chain.c:744-864): the
    # filtered channels ask for a negative delay (peak index / offset from the end / centre tap), the chain turns the
    # difference into `align` effects on the OTHER channels.  Drop-in tier only (the align pass is reference code).
    save("fir_p_align", ":0,2 fir_p -a -t pcm -e double -c 1 -r 48000 ir700.f64", 48000, 3, 256, x)
    save("fir_align_end", ":1 fir -a-100S -t pcm -e double -c 1 -r 48000 ir700.f64", 48000, 3, 512, x)
    save("hilbert_c", ":0 hilbert -c 255", 48000, 3, 333, x)
    save("hilbert_pc", ":0,1 hilbert -p -c 255 :2 eq 1k 1.0 3", 48000, 3, 333, x)
    # reverse IIR (biquad -r, reverse_iir.c): the linear-phase LR4 crossover of examples/crossover_lr4_2kHz_riir_linphase
    # without its remix (4 channels in), two -r sections per band merged by the chain, plus a lone -r section
    x4 = np.concatenate([x, x[:, :1] * 0.5], axis=1)
    save("riir", ":0,1 highpass 2k bw2 highpass -r 2k bw2 : :2,3 lowpass 2k bw2 lowpass -r 2k bw2 : :1 eq -r 500 1.0 4", 48000, 4, 500, x4)
    # resample both ways, ragged blocks, drain included
    x = ref.sgen("sine:freq=1k+3000S", 44100, 2, 3000) * 0.8
    x[:, 1] = rng.standard_normal(3000) * 0.3
    save("resample_up", "resample 48000", 44100, 2, 500, x)
    save("resample_down", "resample 32000", 44100, 2, 777, x)
    save("resample_2x", "resample x2", 44100, 2, 64, x[:1500])
    # full chain (C5 in miniature)
    save("chain", "gain -3 eq 100 1.0 2 eq 5k 2.0 -3 fir_p -t pcm -e double -c 1 -r 44100 ir700.f64 resample 48000", 44100, 2, 512, x)


if __name__ == "__main__":
    main()
