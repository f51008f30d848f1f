"""ctypes binding of shim/_build/libdsp_b200_frontend.so: the reference's chain runtime (parser, optimizer, align
insertion, run_effects_chain -- reference objects, unmodified) with the GPU effects linked in place of
biquad.o gain.o fir.o fir_p.o hilbert.o resample.o, behind a library frontend (shim/frontend.c).

This is the drop-in as a user would embed it: chain strings in the reference's own grammar, host buffers,
`run_effects_chain()` per block.  bench.py's `e2e_dropin` numbers come from `DropinChain.time()`.
The library only exists where /root/reference was present at build time (it travels prebuilt to the GPU box).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "shim", "_build", "libdsp_b200_frontend.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("%s missing: run `make -C shim` where the reference sources exist" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.dspfront_set_loglevel.argtypes = [C.c_int]
        L.dspfront_chain_new.restype = C.c_void_p
        L.dspfront_chain_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.dspfront_chain_free.argtypes = [C.c_void_p]
        for n in ("dspfront_chain_out_fs", "dspfront_chain_out_channels", "dspfront_chain_n_effects"):
            getattr(L, n).restype = C.c_int
            getattr(L, n).argtypes = [C.c_void_p]
        L.dspfront_chain_effect_name.restype = C.c_char_p
        L.dspfront_chain_effect_name.argtypes = [C.c_void_p, C.c_int]
        L.dspfront_chain_max_out_frames.restype = C.c_long
        L.dspfront_chain_max_out_frames.argtypes = [C.c_void_p, C.c_long]
        L.dspfront_chain_run.restype = C.c_long
        L.dspfront_chain_run.argtypes = [C.c_void_p, C.c_long, dp, dp]
        L.dspfront_chain_time.restype = C.c_long
        L.dspfront_chain_time.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, dp, C.c_int, dp, dp]
        _lib = L
    return _lib


class DropinChain:
    """A reference effects chain (chain-string grammar) whose hot-path effects run on the GPU."""

    def __init__(self, chain_str, fs, channels, dir=None):
        self.L = lib()
        self.h = self.L.dspfront_chain_new(chain_str.encode(), int(fs), int(channels), dir.encode() if dir else None)
        if not self.h:
            raise ValueError("chain failed to build: %r" % chain_str)
        self.channels_in = int(channels)
        self.fs_out = self.L.dspfront_chain_out_fs(self.h)
        self.channels_out = self.L.dspfront_chain_out_channels(self.h)

    def effect_names(self):
        return [self.L.dspfront_chain_effect_name(self.h, i).decode() for i in range(self.L.dspfront_chain_n_effects(self.h))]

    def max_out_frames(self, frames):
        return self.L.dspfront_chain_max_out_frames(self.h, int(frames))

    def run(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty((max(self.max_out_frames(x.shape[0]), 1), self.channels_out), dtype=np.float64)
        dp = C.POINTER(C.c_double)
        n = self.L.dspfront_chain_run(self.h, x.shape[0], x.ctypes.data_as(dp), out.ctypes.data_as(dp))
        if n < 0:
            raise RuntimeError("dspfront_chain_run failed")
        return out[:n].copy()

    def time(self, pool, warm, blocks):
        """pool: [n_pool, frames, channels] host blocks.  -> (seconds summed over the `blocks` timed
        run_effects_chain() calls, frames of the last call, checksum of the last result)."""
        pool = np.ascontiguousarray(pool, dtype=np.float64)
        n_pool, frames, ch = pool.shape
        assert ch == self.channels_in
        sec, chk = C.c_double(), C.c_double()
        dp = C.POINTER(C.c_double)
        f = self.L.dspfront_chain_time(self.h, frames, int(warm), int(blocks), pool.ctypes.data_as(dp), n_pool, C.byref(sec), C.byref(chk))
        if f < 0:
            raise RuntimeError("dspfront_chain_time failed")
        return sec.value, f, chk.value

    def close(self):
        if getattr(self, "h", None):
            self.L.dspfront_chain_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
