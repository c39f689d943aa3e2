"""Deviates the way the reference draws them for Monte-Carlo trials (dispatcher, row a20).

sRnd::gauss (xt_base/miscutil/random.cc:51-72) is the polar Box-Muller method on the 31-bit
uniforms of sRnd::random (include/miscutil/random.h:76-79), which come from the BSD random(3)
TYPE_3 generator of xt_base/miscutil/randval.cc.  The stream is restated in the native library
(csrc/mcgen.cpp: b200_mc_gauss / b200_mc_uniform) and pinned bit for bit against the compiled
reference by tests/test_refrand.py; this module is its ctypes face.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def _lib():
    from .engine import load_library
    L = load_library()
    L.b200_mc_gauss.argtypes = [C.c_int, C.c_longlong, C.c_void_p]
    L.b200_mc_uniform.argtypes = [C.c_int, C.c_longlong, C.c_void_p]
    return L


def gauss_stream(seed: int, n: int) -> np.ndarray:
    """The first n values gauss() returns after seed(seed)."""
    out = np.empty(int(n), dtype=np.float64)
    rc = _lib().b200_mc_gauss(int(seed), int(n), out.ctypes.data)
    if rc:
        raise RuntimeError(f"b200_mc_gauss failed ({rc})")
    return out


def uniform_stream(seed: int, n: int) -> np.ndarray:
    """The first n values random() returns after seed(seed)."""
    out = np.empty(int(n), dtype=np.float64)
    rc = _lib().b200_mc_uniform(int(seed), int(n), out.ctypes.data)
    if rc:
        raise RuntimeError(f"b200_mc_uniform failed ({rc})")
    return out


class RefRand:
    """Sequential face of the same stream (one value per call)."""

    def __init__(self, seed: int = 1, block: int = 4096):
        self._seed, self._block = seed, block
        self._buf = gauss_stream(seed, block)
        self._i = 0

    def gauss(self) -> float:
        if self._i >= len(self._buf):
            self._buf = gauss_stream(self._seed, 2*len(self._buf))
        v = float(self._buf[self._i])
        self._i += 1
        return v
