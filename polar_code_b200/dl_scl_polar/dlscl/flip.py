"""Flip-retry controller on the B200 engine (reference: dl_scl_polar/dlscl/flip.py)."""

from __future__ import annotations

from typing import List, Optional

import numpy as np

from polar_code_b200 import engine as _engine
from .._engines import engine_for, llr_row
from ..polar.crc import check_crc
from ..polar.scl import decode_scl


def choose_flip_index(abs_l0: np.ndarray, beta: Optional[np.ndarray]) -> int:
    """argmin(abs_l0 @ beta), or argmin(abs_l0) without beta (flip.py:13-27)."""
    if abs_l0.ndim != 1:
        raise ValueError("abs_l0 must be a 1D array")
    if abs_l0.size == 0:
        raise ValueError("abs_l0 cannot be empty")
    if beta is not None and (beta.ndim != 2 or beta.shape[0] != beta.shape[1] or beta.shape[0] != abs_l0.size):
        raise ValueError("beta must be a square matrix matching abs_l0 length")
    row = np.asarray(abs_l0, np.float32).reshape(1, -1)
    return int(_engine.choose_flip_index(row, None if beta is None else np.asarray(beta, np.float32))[0].item())


def _force_vector(best_path_bits: np.ndarray, flip_index: int) -> np.ndarray:
    """Prefix of the reference path, the flipped bit, the rest free (flip.py:30-34)."""
    forced = np.full(best_path_bits.size, -1, dtype=np.int8)
    forced[:flip_index] = best_path_bits[:flip_index]
    forced[flip_index] = 1 - best_path_bits[flip_index]
    return forced


def retry_with_flip(
    llr_root: np.ndarray,
    info_set: np.ndarray,
    M: int,
    best_path_bits: np.ndarray,
    flip_index: int,
    crc: Optional[str] = None,
) -> dict:
    """One forced re-decode (flip.py:37-62)."""
    if best_path_bits.ndim != 1:
        raise ValueError("best_path_bits must be 1D")
    if flip_index < 0 or flip_index >= best_path_bits.size:
        raise IndexError("flip_index out of range")
    forced = _force_vector(best_path_bits, flip_index)
    result = decode_scl(llr_root, info_set, M, crc=crc, force_info_bits=forced)
    result["forced_info_bits"] = forced
    result["flip_index"] = flip_index
    return result


def decode_with_retries(
    llr_root: np.ndarray,
    info_set: np.ndarray,
    M: int,
    retries: int,
    *,
    crc: Optional[str] = None,
    beta: Optional[np.ndarray] = None,
) -> dict:
    """Baseline SCL plus up to ``retries`` beta-ranked flips (flip.py:65-141).

    The retries (replay of the reference path, |L0| @ beta ranking, forced decode) run fused on the GPU
    (csrc/polar_sweep.cuh); the per-attempt dictionaries of the reference's ``attempts`` list are then
    materialised with one decode_scl call per attempt along the flip indices the kernel chose."""
    if M <= 0:
        raise ValueError("List size M must be positive")
    K = info_set.size
    if beta is not None and (beta.ndim != 2 or beta.shape[0] != beta.shape[1] or beta.shape[0] != K):
        raise ValueError("beta must be a square matrix matching abs_l0 length")
    eng = engine_for(np.asarray(llr_root).size, info_set, crc)
    fused = eng.dlscl_decode(llr_row(llr_root), int(M), int(retries), beta=beta)
    tried = [int(i) for i in fused["tried"][0].cpu().numpy() if i >= 0]

    attempts: List[dict] = []
    baseline = decode_scl(llr_root, info_set, M, crc=crc)
    attempts.append({**baseline, "attempt_type": "baseline"})
    last = baseline
    for idx in tried:
        last = retry_with_flip(llr_root, info_set, M, last["best_path_bits"], flip_index=idx, crc=crc)
        attempts.append({**last, "attempt_type": "flip"})
    final = {**last}
    final["attempts"] = attempts
    final["tried_indices"] = [np.int64(i) for i in tried]
    bits = last.get("best_path_bits")
    final["success"] = bits is not None and (True if crc is None else check_crc(bits, crc))
    return final


__all__ = ["choose_flip_index", "retry_with_flip", "decode_with_retries"]
