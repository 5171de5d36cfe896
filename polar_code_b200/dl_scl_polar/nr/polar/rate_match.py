"""Truncate / repeat rate matching and mean-combining de-rate-matching
(reference: dl_scl_polar/nr/polar/rate_match.py:8-39).  The decoder path fuses the de-rate-matching into the
LLR load of the decode kernel; this host form exists for callers that want the intermediate vector."""

from __future__ import annotations

import numpy as np


def rate_match_polar(bits: np.ndarray, E: int, mode: str = "puncture") -> np.ndarray:
    """First E entries of the codeword repeated as often as needed (rate_match.py:8-16)."""
    if bits.ndim != 1:
        raise ValueError("bits must be 1D")
    if E <= bits.size:
        return bits[:E]
    return np.resize(bits, E)


def derate_match_polar(bits_E: np.ndarray, N: int, mode: str = "puncture") -> np.ndarray:
    """E <= N: pad the unsent tail with -1.0; E > N: mean of the repeats per position (rate_match.py:19-39)."""
    if bits_E.ndim != 1:
        raise ValueError("bits_E must be 1D")
    E = bits_E.size
    out = np.full(N, -1.0, dtype=np.float64)
    if E <= N:
        out[:E] = bits_E
        return out
    total = np.zeros(N, dtype=np.float64)
    full = E // N
    for r in range(full):
        total += bits_E[r * N:(r + 1) * N]
    tail = E - full * N
    total[:tail] += bits_E[full * N:]
    count = np.full(N, full, dtype=np.int32)
    count[:tail] += 1
    return total / np.maximum(count, 1)


__all__ = ["rate_match_polar", "derate_match_polar"]
