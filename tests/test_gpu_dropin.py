"""GPU tier, drop-in level: the reference's OWN chain runtime -- effects_chain.c (parser,
optimizer/merge, align insertion, drain bookkeeping), effect.c, align.c, util.c, fir_util.c, sgen.c, all
unmodified -- linked against the shim's replacement objects (shim/*.c -> libdspb200.so), driven through
the same driver as the pure reference (oracle/ref_driver.c).  Same chain strings, same blocks: outputs
must agree to the tolerance and per-call frame counts must be identical.  Also runs the relinked CLI."""
import os
import subprocess

import numpy as np
import pytest

from conftest import rms

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
DROPIN = os.path.join(HERE, "dropin", "_build", "libdsp_dropin.so")
CLI_GPU = os.path.join(ROOT, "shim", "_build", "dsp_b200")
CLI_REF = os.path.join(ROOT, "oracle", "_ref", "dsp_ref")
RMS_TOL = 1e-10

NAMES = ["gain", "biquad", "fir_p", "fir_p_2ch", "fir", "fir_direct", "hilbert", "resample_up", "resample_down",
         "resample_2x", "chain"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def dropin(gpu_lib):
    if not os.path.exists(DROPIN):
        pytest.fail("tests/dropin/_build/libdsp_dropin.so missing (built by __graft_entry__.build() where /root/reference exists)")
    from oracle import ref
    return ref


@pytest.mark.parametrize("name", NAMES)
def test_dropin_golden(dropin, name):
    g = load(name)
    c = dropin.RefChain(str(g["chain"]), int(g["fs"]), int(g["channels"]), dir=GOLDEN, lib_path=DROPIN)
    y, counts = c.process(g["x"], int(g["block"]))
    assert c.fs_out == int(g["out_fs"])
    assert list(counts) == list(g["counts"])
    assert y.shape == g["y"].shape
    assert rms(y - g["y"]) <= RMS_TOL, rms(y - g["y"])
    c.close()


def test_dropin_merge_fuses_gpu_effects(dropin):
    """effects_chain_optimize() (effects_chain.c:605-641) + the shim's merge hook: runs of GPU effects become one."""
    c = dropin.RefChain("eq 100 1.0 2 eq 1k 1.0 -2 lowshelf 200 0.7 3 fir_p coefs:0.5,0.25,0.125,0.0625,0.03,0.01,0.005,"
                        "0.002,0.001,0.0005,0.0002,0.0001,0.00005,0.00002,0.00001,0.000005,0.000002,0.000001,0.0000005,"
                        "0.0000002,0.0000001,0.00000005,0.00000002,0.00000001,0.000000005,0.000000002,0.000000001,0.0000000005,"
                        "0.0000000002,0.0000000001,0.00000000005,0.00000000002,0.00000000001 eq 5k 2.0 1", 48000, 2, lib_path=DROPIN)
    assert len(c.effect_names()) == 1, c.effect_names()
    c.close()
    # `add` is not reorderable but adjacent merging keeps the signal order: still one device chain, same result
    spec = "eq 100 1.0 2 eq 1k 1.0 -2 add 0.001 eq 5k 2.0 1"
    c = dropin.RefChain(spec, 48000, 2, lib_path=DROPIN)
    assert len(c.effect_names()) == 1, c.effect_names()
    x = np.random.default_rng(0).standard_normal((3000, 2)) * 0.1
    y, _ = c.process(x, 700)
    c.close()
    if dropin.available():
        r = dropin.RefChain(spec, 48000, 2)
        want, _ = r.process(x, 700)
        assert rms(y - want) <= RMS_TOL
    # a reference CPU effect that is not reorderable splits the run: two GPU groups around it
    c = dropin.RefChain("eq 100 1.0 2 eq 1k 1.0 -2 noise -100 eq 5k 2.0 1", 48000, 2, lib_path=DROPIN)
    assert c.effect_names() == ["eq", "noise", "eq"], c.effect_names()
    c.close()
    # a reorderable CPU effect (delay.c) is hopped over, as the reference's own optimizer does for its effects
    c = dropin.RefChain("eq 100 1.0 2 delay 5S eq 5k 2.0 1", 48000, 2, lib_path=DROPIN)
    assert c.effect_names()[:2] == ["eq", "delay"] and "eq" not in c.effect_names()[1:], c.effect_names()
    c.close()


def test_dropin_matches_reference_on_c2(dropin, have_ref):
    """Config 2 shape: gain + 10 eq, 16 channels, 4096-frame blocks, reference sgen sweep."""
    if not have_ref:
        pytest.skip("compiled reference did not travel")
    chain = "gain -12 " + " ".join("eq %s 1.4 %s" % (f, g) for f, g in zip(
        ["31.25", "62.5", "125", "250", "500", "1k", "2k", "4k", "8k", "16k"], ["-2", "1.5", "-1", "2", "-1.5", "1", "-2", "1.5", "-1", "2"]))
    x = dropin.sgen("sine:freq=20-20k+20000S", 48000, 16, 20000)
    a = dropin.RefChain(chain, 48000, 16)
    b = dropin.RefChain(chain, 48000, 16, lib_path=DROPIN)
    assert b.effect_names() == ["gain"]                # gain + ten biquads: ONE device chain (one H2D/D2H per block)
    ya, ca = a.process(x, 4096)
    yb, cb = b.process(x, 4096)
    assert ca == cb
    assert rms(ya - yb) <= RMS_TOL, rms(ya - yb)


def test_dropin_cli(dropin, tmp_path, have_ref):
    """The relinked CLI (unmodified dsp.c) against the reference CLI, raw float64 out, bit layout identical."""
    if not (os.path.exists(CLI_GPU) and os.path.exists(CLI_REF)):
        pytest.skip("CLI builds did not travel")
    args = ["-q", "-b", "1024", "-t", "sgen", "-c", "4", "-r", "44100", "sine:freq=50-15k+30000S", "-o", "-t", "pcm", "-e", "double"]
    chain = ["gain", "-6", "eq", "200", "1.0", "3", ":0,1", "hilbert", "-p", "255", ":", "fir_p", "-t", "pcm", "-e", "double",
             "-c", "1", "-r", "44100", os.path.join(GOLDEN, "ir700.f64"), "resample", "48k"]
    outs = []
    for exe, name in ((CLI_REF, "ref.f64"), (CLI_GPU, "gpu.f64")):
        out = str(tmp_path / name)
        r = subprocess.run([exe] + args + [out] + chain, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(np.fromfile(out, dtype="<f8").reshape(-1, 4))
    assert outs[0].shape == outs[1].shape and outs[0].shape[0] > 30000
    assert rms(outs[0] - outs[1]) <= RMS_TOL, rms(outs[0] - outs[1])
