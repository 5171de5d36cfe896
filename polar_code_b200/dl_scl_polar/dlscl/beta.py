"""Learnable symmetric flip metric (training side).

Interface of the reference's `SymmetricBeta` (dl_scl_polar/dlscl/beta.py:9-46): parameter `off_diag` [dim,dim],
`beta_matrix()` = I + U + U^T with U the strict upper triangle of `off_diag`, `forward(x)` = x @ beta for a vector
or a batch, `clamp_diagonal()` zeroes the unused diagonal of the raw parameter.  The engine's retry rounds only
consume the exported float32 `.npy` (csrc/polar_sweep.cuh, dl_retry_kernel)."""

from __future__ import annotations

import torch


class SymmetricBeta(torch.nn.Module):
    def __init__(self, dim: int, init_range: float = 0.2) -> None:
        if dim <= 0:
            raise ValueError("dim must be positive")
        super().__init__()
        self.dim = int(dim)
        self.init_range = float(init_range)
        raw = torch.empty(self.dim, self.dim).uniform_(-self.init_range, self.init_range)
        raw.diagonal().zero_()
        self.off_diag = torch.nn.Parameter(raw)
        self.register_buffer("_eye", torch.eye(self.dim), persistent=False)

    @torch.no_grad()
    def clamp_diagonal(self) -> None:
        self.off_diag.diagonal().zero_()

    def beta_matrix(self) -> torch.Tensor:
        upper = self.off_diag.triu(diagonal=1)
        return self._eye.to(upper.dtype) + upper + upper.mT

    def forward(self, abs_l0: torch.Tensor) -> torch.Tensor:
        if abs_l0.dim() == 1 or abs_l0.dim() == 2:
            return abs_l0.matmul(self.beta_matrix())
        raise ValueError("abs_l0 must be 1D or 2D tensor")


__all__ = ["SymmetricBeta"]
