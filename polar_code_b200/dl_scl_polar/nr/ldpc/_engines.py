"""One LdpcEngine per (device, H) for the per-call mirror API."""

from __future__ import annotations

import numpy as np
import torch

from polar_code_b200.engine import require_cuda
from polar_code_b200.ldpc import LdpcEngine

_cache = {}


def ldpc_engine_for(H: np.ndarray) -> LdpcEngine:
    require_cuda()
    Hm = np.ascontiguousarray(np.asarray(H) % 2, np.uint8)
    key = (torch.cuda.current_device(), Hm.shape, Hm.tobytes())
    eng = _cache.get(key)
    if eng is None:
        if len(_cache) > 16:
            _cache.clear()
        eng = _cache[key] = LdpcEngine(Hm)
    return eng
