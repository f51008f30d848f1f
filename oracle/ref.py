"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).

ctypes binding of oracle/_ref/libdspref.so: the UNMODIFIED reference sources of bmc0/dsp
(compiled from /root/reference by oracle/Makefile) behind oracle/ref_driver.c.  This is the
real parity oracle; oracle/restate.py and oracle/port.c are restatements checked against it.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdspref.so")
CLI_PATH = os.path.join(_HERE, "_ref", "dsp_ref")

_libs = {}


def available():
    return os.path.exists(LIB_PATH)


def lib(path=None):
    """The reference driver library.  `path` selects another build with the same driver API
    (tests/dropin builds the reference chain runtime around the GPU effects)."""
    path = path or LIB_PATH
    if path not in _libs:
        if not os.path.exists(path):
            raise RuntimeError("%s missing: run `make -C oracle ref` where /root/reference exists" % path)
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.dspref_set_loglevel.argtypes = [C.c_int]
        L.dspref_chain_new.restype = C.c_void_p
        L.dspref_chain_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        for name in ("dspref_chain_out_fs", "dspref_chain_out_channels", "dspref_chain_n_effects"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        L.dspref_chain_effect_name.restype = C.c_char_p
        L.dspref_chain_effect_name.argtypes = [C.c_void_p, C.c_int]
        L.dspref_chain_max_out_frames.restype = C.c_long
        L.dspref_chain_max_out_frames.argtypes = [C.c_void_p, C.c_long]
        L.dspref_chain_buffer_len.restype = C.c_long
        L.dspref_chain_buffer_len.argtypes = [C.c_void_p, C.c_long]
        L.dspref_chain_delay.restype = C.c_double
        L.dspref_chain_delay.argtypes = [C.c_void_p]
        L.dspref_chain_drain_frames.restype = C.c_long
        L.dspref_chain_drain_frames.argtypes = [C.c_void_p]
        L.dspref_chain_run.restype = C.c_long
        L.dspref_chain_run.argtypes = [C.c_void_p, C.c_long, dp, dp]
        L.dspref_chain_run_inplace.restype = C.c_long
        L.dspref_chain_run_inplace.argtypes = [C.c_void_p, C.c_long, C.c_int]
        L.dspref_chain_drain.restype = C.c_long
        L.dspref_chain_drain.argtypes = [C.c_void_p, C.c_long, dp]
        L.dspref_chain_reset.argtypes = [C.c_void_p]
        L.dspref_chain_free.argtypes = [C.c_void_p]
        L.dspref_sgen.restype = C.c_long
        L.dspref_sgen.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_long, dp]
        L.dspref_biquad_design.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, dp]
        L.dspref_next_fast_fftw_len.restype = C.c_long
        L.dspref_next_fast_fftw_len.argtypes = [C.c_long]
        _libs[path] = L
    return _libs[path]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefChain:
    """A reference effects chain built from a chain string, e.g. "gain -6 eq 1k 1.0 3"."""

    def __init__(self, chain_str, fs, channels, dir=None, lib_path=None):
        self.L = lib(lib_path)
        self.h = self.L.dspref_chain_new(chain_str.encode(), fs, channels, dir.encode() if dir else None)
        if not self.h:
            raise ValueError("reference failed to build chain: %r" % chain_str)
        self.fs_in, self.channels_in = fs, channels
        self.fs_out = self.L.dspref_chain_out_fs(self.h)
        self.channels_out = self.L.dspref_chain_out_channels(self.h)

    def effect_names(self):
        return [self.L.dspref_chain_effect_name(self.h, i).decode() for i in range(self.L.dspref_chain_n_effects(self.h))]

    def max_out_frames(self, frames):
        return self.L.dspref_chain_max_out_frames(self.h, frames)

    def run(self, x):
        """x: [frames, channels] float64 -> [out_frames, channels_out]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        frames = x.shape[0]
        out = np.empty((max(self.max_out_frames(frames), 1), self.channels_out), dtype=np.float64)
        n = self.L.dspref_chain_run(self.h, frames, _dp(x), _dp(out))
        if n < 0:
            raise RuntimeError("dspref_chain_run failed")
        return out[:n].copy()

    def run_inplace(self, frames, refill=False):
        return self.L.dspref_chain_run_inplace(self.h, frames, int(refill))

    def drain(self, block):
        """Drain the chain as the CLI does (dsp.c:1038-1044): list of blocks until dry."""
        outs = []
        cap = max(self.max_out_frames(block), 1)
        while True:
            out = np.empty((cap, self.channels_out), dtype=np.float64)
            n = self.L.dspref_chain_drain(self.h, block, _dp(out))
            if n == -1:
                break
            if n < 0:
                raise RuntimeError("dspref_chain_drain failed")
            outs.append(out[:n].copy())
        return outs

    def process(self, x, block, drain=True):
        """Whole stream through the chain in `block`-frame calls; returns (concatenated output, per-call frame counts)."""
        outs, counts = [], []
        for i in range(0, x.shape[0], block):
            y = self.run(x[i:i + block])
            outs.append(y)
            counts.append(y.shape[0])
        if drain:
            for y in self.drain(block):
                outs.append(y)
                counts.append(y.shape[0])
        y = np.concatenate(outs, axis=0) if outs else np.zeros((0, self.channels_out))
        return y, counts

    def delay(self):
        return self.L.dspref_chain_delay(self.h)

    def reset(self):
        self.L.dspref_chain_reset(self.h)

    def close(self):
        if self.h:
            self.L.dspref_chain_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sgen(spec, fs, channels, frames):
    """The reference's own signal generator (sgen.c); spec e.g. "sine:freq=20-20k+10s"."""
    out = np.zeros((frames, channels), dtype=np.float64)
    n = lib().dspref_sgen(spec.encode(), fs, channels, frames, _dp(out))
    if n < 0:
        raise ValueError("sgen failed: %r" % spec)
    return out[:n]


def biquad_design(type_, fs, arg0, arg1=0.0, arg2=0.0, arg3=0.0, width_type=1):
    c = np.zeros(5)
    lib().dspref_biquad_design(type_, fs, arg0, arg1, arg2, arg3, width_type, _dp(c))
    return c


def next_fast_fftw_len(n):
    return lib().dspref_next_fast_fftw_len(n)
