"""Base-graph table (reference: dl_scl_polar/nr/ldpc/basegraphs.py:12-42).  Both ids name the same 3x6 demo
graph; the shift table itself lives in the C-ABI (pb200_ldpc_build_h) and is read back from a Z=4 lift, for which
every shift of the demo graph is its own residue."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from polar_code_b200.ldpc import build_h_matrix as _abi_build_h


@dataclass
class BaseGraph:
    name: str
    m: int          # rows of the base graph
    n: int          # columns of the base graph
    shifts: np.ndarray  # (m, n), -1 = zero block
    bg: int = 2     # id understood by the C-ABI


def _read_back(bg: int, name: str) -> BaseGraph:
    Z = 4
    H = _abi_build_h(bg, Z)
    m, n = H.shape[0] // Z, H.shape[1] // Z
    shifts = np.full((m, n), -1, np.int32)
    for r in range(m):
        for c in range(n):
            first_row = H[r * Z, c * Z:(c + 1) * Z]
            if first_row.any():
                shifts[r, c] = int(np.argmax(first_row))
    return BaseGraph(name=name, m=m, n=n, shifts=shifts, bg=bg)


_NAMES = {1: "BG_demo1", 2: "BG_demo2"}
_cache = {}


def load_base_graph(bg: int) -> BaseGraph:
    if bg not in _NAMES:
        raise ValueError(f"Unknown base graph: {bg}")
    if bg not in _cache:
        _cache[bg] = _read_back(bg, _NAMES[bg])
    return _cache[bg]


__all__ = ["BaseGraph", "load_base_graph"]
