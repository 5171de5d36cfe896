"""Rate-matched polar encode / SCL decode on the B200 engine
(reference: dl_scl_polar/nr/polar/scl_nr.py:23-57)."""

from __future__ import annotations

from typing import Dict

import numpy as np

from ..._engines import engine_for, llr_row
from ...polar.crc import _degree


def encode_rate_matched(
    payload_bits: np.ndarray,
    crc_poly: str,
    N: int,
    E: int,
    info_set: np.ndarray,
    ilv_mode: str = "default",
) -> np.ndarray:
    """payload -> CRC -> polar encode -> sub-block interleave -> rate match, one fused kernel (scl_nr.py:23-35)."""
    if payload_bits.ndim != 1:
        raise ValueError("msg_bits must be a 1D array")
    if payload_bits.size + _degree(crc_poly) != np.asarray(info_set).size:
        raise ValueError("info_bits length must match info_set size")
    eng = engine_for(N, info_set, crc_poly, E=E)
    tx = eng.nr_encode((payload_bits.astype(np.int8) & 1).astype(np.uint8).reshape(1, -1), E)
    return tx.cpu().numpy()[0].astype(np.int8)


def decode_rate_matched_scl(
    llr_E: np.ndarray,
    crc_poly: str,
    N: int,
    E: int,
    info_set: np.ndarray,
    M: int,
    ilv_mode: str = "default",
) -> Dict[str, np.ndarray]:
    """De-rate-match + de-interleave fused into the LLR load, then SCL (scl_nr.py:38-57)."""
    llr_E = np.asarray(llr_E)
    if llr_E.size != E:
        raise ValueError("llr_E must have length E")
    eng = engine_for(N, info_set, crc_poly, E=E)
    out = eng.scl_decode(llr_row(llr_E), int(M), want=("best_bits", "crc_ok", "n_cand"))
    bits = out["best_bits"][0].cpu().numpy().astype(np.int8)
    return {"payload": bits[: len(info_set)], "crc_pass": bool(out["crc_ok"][0].item()), "best_path_bits": bits}


__all__ = ["encode_rate_matched", "decode_rate_matched_scl"]
