"""xictools_b200: B200-native batched engine for WRspice's Newton inner loop.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C ABI), the host-side
deck description (deck.py) and the ctypes binding of the C ABI (engine.py).
"""
from .deck import Deck, ring_oscillator, mc_variation  # noqa: F401
