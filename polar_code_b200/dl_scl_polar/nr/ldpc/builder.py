"""Lifted parity-check matrix (reference: dl_scl_polar/nr/ldpc/builder.py:20-30)."""

from __future__ import annotations

import numpy as np

from polar_code_b200.ldpc import build_h_matrix as _abi_build_h
from .basegraphs import BaseGraph, load_base_graph


def build_h_matrix(base_graph: BaseGraph, Z: int) -> np.ndarray:
    """int8 [m*Z, n*Z]: block (r, c) is the identity shifted right by shifts[r, c] mod Z, or zero for -1."""
    known = load_base_graph(base_graph.bg) if getattr(base_graph, "bg", None) in (1, 2) else None
    if known is not None and np.array_equal(known.shifts, base_graph.shifts):
        return _abi_build_h(base_graph.bg, Z)
    # a caller-made base graph: same lifting rule, index arithmetic only
    H = np.zeros((base_graph.m * Z, base_graph.n * Z), np.int8)
    rows = np.arange(Z)
    for r in range(base_graph.m):
        for c in range(base_graph.n):
            s = int(base_graph.shifts[r, c])
            if s >= 0:
                H[r * Z + rows, c * Z + (rows + s % Z) % Z] = 1
    return H


__all__ = ["build_h_matrix"]
