"""Puncture / repeat rate matching for the LDPC branch (reference: dl_scl_polar/nr/ldpc/rate_match.py:8-38).
The sweep and the batched decoder fuse the de-rate-matching into the LLR load; these per-vector forms go through
the same C-ABI kernels."""

from __future__ import annotations

import numpy as np
import torch

from polar_code_b200 import _lib as L
from polar_code_b200.engine import _ptr, _stream, require_cuda


def rate_match_ldpc(codeword: np.ndarray, E: int) -> np.ndarray:
    if codeword.ndim != 1:
        raise ValueError("codeword must be 1D")
    if E <= codeword.size:
        return codeword[:E]                      # a view, like the reference's slice
    require_cuda()
    c = torch.from_numpy(np.ascontiguousarray(codeword).astype(np.uint8)).cuda().reshape(1, -1)
    out = torch.empty((1, int(E)), dtype=torch.uint8, device=c.device)
    L.check(L.load().pb200_ldpc_rate_match_batch(_ptr(c), c.shape[1], int(E), _ptr(out), 1, _stream()))
    return out[0].cpu().numpy().astype(codeword.dtype)


def derate_match_ldpc(llr: np.ndarray, N: int) -> np.ndarray:
    if llr.ndim != 1:
        raise ValueError("llr must be 1D")
    require_cuda()
    x = torch.from_numpy(np.ascontiguousarray(llr, np.float64)).cuda().reshape(1, -1)
    out = torch.empty((1, int(N)), dtype=torch.float64, device=x.device)
    L.check(L.load().pb200_ldpc_derate_match_batch(_ptr(x), x.shape[1], int(N), _ptr(out), 1, _stream()))
    return out[0].cpu().numpy()


__all__ = ["rate_match_ldpc", "derate_match_ldpc"]
