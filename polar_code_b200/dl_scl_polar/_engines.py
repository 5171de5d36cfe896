"""Engine cache for the per-frame mirror API: one PolarEngine per (device, N, info_set, crc, E)."""

from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from polar_code_b200.engine import PolarEngine, require_cuda

_cache = {}


def check_power_of_two(n: int, what: str = "N must be a power of two") -> int:
    if n <= 0 or (n & (n - 1)) != 0:
        raise ValueError(what)
    return int(math.log2(n))


def engine_for(N: int, info_set, crc: Optional[str], E: int = 0) -> PolarEngine:
    require_cuda()
    a = np.ascontiguousarray(np.asarray(info_set), np.int32)
    key = (torch.cuda.current_device(), int(N), a.tobytes(), crc, int(E))
    eng = _cache.get(key)
    if eng is None:
        if len(_cache) > 64:
            _cache.clear()
        eng = PolarEngine(int(N), a, crc)
        if E:
            eng.set_rate_matching(int(E))
        _cache[key] = eng
    return eng


def llr_row(llr) -> np.ndarray:
    """The engine computes in fp32 (north star); a float64 input row is rounded once here."""
    return np.ascontiguousarray(np.asarray(llr, dtype=np.float64).astype(np.float32)).reshape(1, -1)
