"""CPU tier of the drop-in boundary: the relinked chain runtime loads, CPU effects still work through
it, and a GPU effect fails its init() loudly (returns NULL -> chain build fails) without a device."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DROPIN = os.path.join(HERE, "dropin", "_build", "libdsp_dropin.so")


def test_dropin_library_without_gpu():
    if not os.path.exists(DROPIN):
        pytest.skip("drop-in harness not built (needs /root/reference at build time)")
    import dsp_b200
    from oracle import ref
    c = ref.RefChain("remix 1 0", 48000, 2, lib_path=DROPIN)      # reference remix.c, still CPU: swaps the channels
    x = np.stack([np.ones(16), np.zeros(16)], axis=1)
    y = c.run(x)
    assert np.array_equal(y[:, 0], x[:, 1]) and np.array_equal(y[:, 1], x[:, 0])
    c.close()
    if dsp_b200.device_count() < 1:
        with pytest.raises(ValueError):
            ref.RefChain("eq 1k 1.0 3", 48000, 2, lib_path=DROPIN)
        # "!" lets an effect fail without aborting the chain build (effects_chain.c:455-458)
        c = ref.RefChain("remix 1 0 ! eq 1k 1.0 3", 48000, 2, lib_path=DROPIN)
        assert c.effect_names() == ["remix"]


def test_shim_exports_reference_symbols():
    """The replacement objects export exactly the symbols the reference's biquad.o fir.o fir_p.o hilbert.o
    resample.o export."""
    import subprocess
    objdir = os.path.join(os.path.dirname(HERE), "shim", "_build", "obj")
    if not os.path.isdir(objdir):
        pytest.skip("shim objects not built")
    want = {"biquad.o": {"biquad_effect_init", "biquad_init", "biquad_reset", "biquad_init_using_type"},
            "gain.o": {"gain_effect_init"},
            "fir.o": {"fir_effect_init", "fir_effect_init_with_filter"},
            "fir_p.o": {"fir_p_effect_init", "fir_p_effect_init_with_filter"},
            "hilbert.o": {"hilbert_effect_init"}, "resample.o": {"resample_effect_init"}}
    for obj, syms in want.items():
        out = subprocess.run(["nm", "-g", "--defined-only", os.path.join(objdir, obj)], stdout=subprocess.PIPE, text=True).stdout
        have = {line.split()[-1] for line in out.splitlines() if " T " in line}
        assert syms <= have, (obj, syms - have)
