"""decode_scl on the B200 engine (reference: dl_scl_polar/polar/scl.py:108-209)."""

from __future__ import annotations

from typing import Optional

import numpy as np

from .._engines import check_power_of_two, engine_for, llr_row


def decode_scl(
    llr: np.ndarray,
    info_set: np.ndarray,
    M: int,
    crc: Optional[str] = None,
    *,
    force_info_bits: Optional[np.ndarray] = None,
) -> dict:
    """SCL decode of one frame; same keys and ownership as the reference (scl.py:203-209).

    ``best_path_bits`` is the very array stored in ``candidates`` (scl.py:199-201)."""
    if M <= 0:
        raise ValueError("List size M must be positive")
    if info_set.ndim != 1:
        raise ValueError("info_set must be a 1D array")
    llr = np.asarray(llr)
    check_power_of_two(llr.size, "Channel LLR length must be a power of two")
    K = info_set.size
    force = None
    if force_info_bits is not None:
        if force_info_bits.ndim != 1:
            raise ValueError("force_info_bits must be 1D when provided")
        if force_info_bits.size != K:
            raise ValueError("force_info_bits length must match info_set")
        force = force_info_bits.astype(np.int8)
        if np.any((force < -1) | (force > 1)):
            raise ValueError("force_info_bits entries must be -1, 0, or 1")
        force = force.reshape(1, -1)
    eng = engine_for(llr.size, info_set, crc)
    out = eng.scl_decode(llr_row(llr), int(M), force=force,
                         want=("cand", "metrics", "info_llrs", "n_cand", "best_idx", "flags"))
    n = int(out["n_cand"][0].item())
    if n == 0:
        raise RuntimeError("All paths pruned during decoding")
    cand = out["cand"][0, :n].cpu().numpy().astype(np.int8)
    candidates = [cand[i].copy() for i in range(n)]
    metrics = [float(v) for v in out["metrics"][0, :n].cpu().numpy()]
    info_llrs = [row.astype(float) for row in out["info_llrs"][0, :n].cpu().numpy()]
    best = int(out["best_idx"][0].item())
    return {
        "candidates": candidates,
        "metrics": metrics,
        "best_path_bits": candidates[best],
        "info_llrs": info_llrs,
        "best_path_info_llrs": info_llrs[best],
    }


__all__ = ["decode_scl"]
