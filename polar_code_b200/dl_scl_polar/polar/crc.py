"""attach_crc / check_crc on the B200 engine (reference: dl_scl_polar/polar/crc.py)."""

from __future__ import annotations

import numpy as np

from polar_code_b200 import engine as _engine


def _poly_to_bits(poly: str) -> np.ndarray:
    """Hex string (leading 1 included) -> MSB-first coefficient bits (crc.py:10-16)."""
    if not poly:
        raise ValueError("CRC polynomial string must be non-empty")
    value = int(poly, 16)
    return np.array([int(c) for c in bin(value)[2:]] if value else [], dtype=np.int8)


def _degree(poly: str) -> int:
    degree = _poly_to_bits(poly).size - 1
    if degree <= 0:
        raise ValueError("Polynomial degree must be positive")
    return degree


def attach_crc(msg_bits: np.ndarray, poly: str) -> np.ndarray:
    """msg_bits followed by the CRC remainder of msg(x)*x^deg (crc.py:19-37)."""
    if msg_bits.ndim != 1:
        raise ValueError("msg_bits must be a 1D array")
    _degree(poly)
    bits = (msg_bits.astype(np.int8) & 1).astype(np.uint8).reshape(1, -1)
    return _engine.crc_attach(bits, poly).cpu().numpy()[0].astype(np.int8)


def check_crc(msg_with_crc: np.ndarray, poly: str) -> bool:
    """True when msg_with_crc(x) is divisible by the polynomial (crc.py:40-56)."""
    if msg_with_crc.ndim != 1:
        raise ValueError("msg_with_crc must be a 1D array")
    if msg_with_crc.size <= _degree(poly):
        raise ValueError("Message too short for the provided CRC polynomial")
    bits = (msg_with_crc.astype(np.int8) & 1).astype(np.uint8).reshape(1, -1)
    return bool(_engine.crc_check(bits, poly).cpu().numpy()[0])


__all__ = ["attach_crc", "check_crc"]
