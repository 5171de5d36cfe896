"""construct_info_set / encode / sc_decode on the B200 engine (reference: dl_scl_polar/polar/polar.py)."""

from __future__ import annotations

import functools

import numpy as np

from .. import config
from .._engines import check_power_of_two, engine_for, llr_row
from polar_code_b200.engine import construct_info_set as _construct


@functools.lru_cache(maxsize=None)
def construct_info_set(N: int, K: int, method: str = "gaussian", design_snr_db: float = 2.5) -> np.ndarray:
    """Sorted int32 indices of the K most reliable bit channels (polar.py:85-103).

    Like the reference the result is cached and the SAME array object is returned on every call."""
    check_power_of_two(int(N))
    if not (0 < K <= N):
        raise ValueError("K must satisfy 0 < K <= N")
    if method not in ("gaussian", "polarization"):
        raise ValueError(f"Unsupported construction method: {method}")
    return _construct(int(N), int(K), method, float(design_snr_db)).astype(np.int32)


def _polar_transform(u: np.ndarray) -> np.ndarray:
    """x = u F^{(x)n} in natural order (polar.py:17-29), evaluated by the encode kernel with A = all positions."""
    u = np.asarray(u)
    n = u.size
    eng = engine_for(n, np.arange(n, dtype=np.int32), None)
    return eng.encode((u.astype(np.int8) & 1).astype(np.uint8).reshape(1, -1)).cpu().numpy()[0].astype(u.dtype)


def encode(msg_bits: np.ndarray) -> np.ndarray:
    """Encode with the default P(N,K) of config.DEFAULTS (polar.py:106-119)."""
    cfg = config.DEFAULTS
    if msg_bits.ndim != 1:
        raise ValueError("msg_bits must be 1D")
    if msg_bits.size != cfg.K:
        raise ValueError(f"msg_bits must have length {cfg.K}")
    eng = engine_for(cfg.N, construct_info_set(cfg.N, cfg.K), None)
    bits = (msg_bits.astype(np.int8) & 1).astype(np.uint8).reshape(1, -1)
    return eng.encode(bits).cpu().numpy()[0].astype(np.int8)


def sc_decode(llr: np.ndarray, info_set: np.ndarray) -> np.ndarray:
    """Successive-cancellation decode; returns the estimated information bits (polar.py:130-168)."""
    if llr.ndim != 1:
        raise ValueError("llr must be 1D")
    check_power_of_two(llr.size)
    if info_set.ndim != 1:
        raise ValueError("info_set must be 1D")
    if np.any(info_set < 0) or np.any(info_set >= llr.size):
        raise ValueError("info_set indices out of range")
    eng = engine_for(llr.size, info_set, None)
    return eng.sc_decode(llr_row(llr)).cpu().numpy()[0].astype(np.int8)


__all__ = ["construct_info_set", "encode", "sc_decode"]
