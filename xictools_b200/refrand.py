"""Gaussian deviates the way the reference draws them for Monte-Carlo trials.

sRnd::gauss (xt_base/miscutil/random.cc:51-72): Box-Muller *polar* method on two uniforms
u = (r & 0x7fffffff)/2^31 (include/miscutil/random.h:76-79), returning v2*fac first and
caching v1*fac for the next call.

The 31-bit uniform stream itself comes, in the reference, from a FreeBSD random(3)
derivative (xt_base/miscutil/randval.cc).  Trial generation is SURVEY.md section 8(f) "next"
scope, so here the uniform integers come from a seeded numpy generator; parity tests always
feed the same generated values to both the engine and the reference.
"""
from __future__ import annotations

import math

import numpy as np


class RefRand:
    def __init__(self, seed: int = 1):
        self._rng = np.random.Generator(np.random.PCG64(seed))
        self._set = False
        self._val = 0.0

    def random(self) -> float:
        r = int(self._rng.integers(0, 2**32, dtype=np.uint64))
        return (r & 0x7FFFFFFF) / (0x7FFFFFFF + 1.0)

    def gauss(self) -> float:
        if not self._set:
            while True:
                v1 = 2.0 * self.random() - 1.0
                v2 = 2.0 * self.random() - 1.0
                r = v1 * v1 + v2 * v2
                if r < 1.0 and r > 0.0:
                    break
            fac = math.sqrt(-2.0 * math.log(r) / r)
            self._val = v1 * fac
            self._set = True
            return v2 * fac
        self._set = False
        return self._val
