"""Learnable symmetric flip metric beta (reference: dl_scl_polar/dlscl/beta.py:9-46).

beta = I + U + U^T where U is the strict upper triangle of one learnable matrix; `forward` scores
Q = |L0| @ beta for a vector or a batch.  Inference inside the engine only consumes the exported `.npy`
(csrc/polar_sweep.cuh dl_round_kernel); this module exists for the training side.
"""

from __future__ import annotations

import torch
from torch import nn


class SymmetricBeta(nn.Module):
    def __init__(self, dim: int, init_range: float = 0.2) -> None:
        if dim <= 0:
            raise ValueError("dim must be positive")
        super().__init__()
        self.dim, self.init_range = int(dim), float(init_range)
        w = (torch.rand(dim, dim) * 2.0 - 1.0) * self.init_range
        self.off_diag = nn.Parameter(w - torch.diag(torch.diagonal(w)))

    def clamp_diagonal(self) -> None:
        """Keep the (unused) diagonal of the raw parameter at zero."""
        with torch.no_grad():
            self.off_diag.diagonal().zero_()

    def beta_matrix(self) -> torch.Tensor:
        u = self.off_diag.triu(1)
        return u + u.T + torch.eye(self.dim, device=u.device, dtype=u.dtype)

    def forward(self, abs_l0: torch.Tensor) -> torch.Tensor:
        if abs_l0.dim() not in (1, 2):
            raise ValueError("abs_l0 must be 1D or 2D tensor")
        return abs_l0 @ self.beta_matrix()


__all__ = ["SymmetricBeta"]
