"""Layered normalised min-sum decoder on the B200 engine (reference: dl_scl_polar/nr/ldpc/decode_nms.py:8-40)."""

from __future__ import annotations

import numpy as np

from ._engines import ldpc_engine_for


def decode_ldpc_nms(llr: np.ndarray, H: np.ndarray, max_iter: int = 20, alpha: float = 0.8, early_stop: bool = True) -> dict:
    """Same keys as the reference: hard (int8 [n]), iters_used, parity_ok -- float64 on the GPU, bit-identical."""
    m, n = H.shape
    if llr.size != n:
        raise ValueError("llr length mismatch")
    out = ldpc_engine_for(H).decode(np.asarray(llr, np.float64).reshape(1, -1), max_iter=max_iter, alpha=alpha,
                                    early_stop=early_stop)
    return {"hard": out["hard"][0].cpu().numpy().astype(np.int8), "iters_used": int(out["iters_used"][0].item()),
            "parity_ok": bool(out["parity_ok"][0].item())}


__all__ = ["decode_ldpc_nms"]
