"""32-row block (de)interleaver index maps (reference: dl_scl_polar/nr/polar/interleaver.py:10-37).

These are pure index permutations used by callers and tests; inside the engine the same maps are folded into
the gather tables of the LLR load / transmit kernels (csrc/polar_abi.cu pb200_set_rate_matching)."""

from __future__ import annotations

import numpy as np

_ROWS = 32


def _read_order(length: int) -> np.ndarray:
    """order[i] = (i % 32) * cols + i // 32 over the padded cols*32 grid (interleaver.py:20)."""
    cols = -(-length // _ROWS)
    return np.arange(cols * _ROWS, dtype=np.int32).reshape(_ROWS, cols).T.reshape(-1)


def subblock_interleave(bits: np.ndarray, mode: str = "default") -> np.ndarray:
    """Write row-wise into a 32 x cols grid padded with -1, read column-wise (interleaver.py:10-23)."""
    if bits.ndim != 1:
        raise ValueError("bits must be 1D")
    order = _read_order(bits.size)
    grid = np.full(order.size, -1, dtype=bits.dtype)
    grid[: bits.size] = bits
    return grid[order]


def subblock_deinterleave(bits: np.ndarray, original_len: int, mode: str = "default") -> np.ndarray:
    """Inverse gather; missing tail entries read as 0 (interleaver.py:26-37)."""
    if bits.ndim != 1:
        raise ValueError("bits must be 1D")
    order = _read_order(original_len)
    grid = np.zeros(order.size, dtype=bits.dtype)
    grid[: bits.size] = bits
    out = np.zeros(order.size, dtype=bits.dtype)
    out[order] = grid
    return out[:original_len]


__all__ = ["subblock_interleave", "subblock_deinterleave"]
