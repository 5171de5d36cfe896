"""Defaults shared by the mirror modules (reference: dl_scl_polar/config.py:9-27)."""

from __future__ import annotations

import dataclasses
from typing import List


def _sizes() -> List[int]:
    return [1, 2, 4, 8]


def _sweep() -> List[float]:
    return [4.0, 6.5, 0.5]


@dataclasses.dataclass
class PolarConfig:
    """P(128,64) with the 24-bit CRC 0x1864CFB, list sizes 1/2/4/8, 8 flip retries, Eb/N0 4.0..6.5 dB."""

    N: int = 128
    K: int = 64
    crc_poly: str = "0x1864CFB"
    crc_bits: int = 24
    list_sizes: List[int] = dataclasses.field(default_factory=_sizes)
    retries: int = 8
    ebno_sweep: List[float] = dataclasses.field(default_factory=_sweep)
    seed: int = 0


DEFAULTS = PolarConfig()


def get_config() -> PolarConfig:
    """A fresh PolarConfig carrying the current DEFAULTS values."""
    return dataclasses.replace(DEFAULTS)
