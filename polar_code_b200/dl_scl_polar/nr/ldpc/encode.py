"""Systematic LDPC encoder on the B200 engine (reference: dl_scl_polar/nr/ldpc/encode.py:52-66)."""

from __future__ import annotations

import numpy as np

from ._engines import ldpc_engine_for


def encode_ldpc(payload: np.ndarray, H: np.ndarray) -> np.ndarray:
    """payload [k] -> codeword [n] int8 = [payload | parity], parity solving H_par p = H_sys payload over GF(2)."""
    if payload.ndim != 1:
        raise ValueError("payload must be 1D")
    n = H.shape[1]
    if n <= payload.size:
        raise ValueError("Parity-check matrix too small for payload length")
    code, status = ldpc_engine_for(H).encode(payload.astype(np.uint8).reshape(1, -1) & 1, want_status=True)
    if int(status[0].item()):
        raise ValueError("Linear system over GF(2) has no solution")
    return code[0].cpu().numpy().astype(np.int8)


__all__ = ["encode_ldpc"]
