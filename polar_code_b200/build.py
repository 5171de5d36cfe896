"""Build libpolar_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery)."""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_CSRC = _PKG / "csrc"
_LIB = _PKG / "libpolar_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-diag-suppress", "177",
]


def library_path() -> Path:
    return _LIB


def _sources():
    return sorted(list(_CSRC.glob("*.cu")) + list(_CSRC.glob("*.cuh")) + list(_CSRC.glob("*.inl")) +
                  [_PKG.parent / "include" / "polar_b200.h"])


def needs_build() -> bool:
    if not _LIB.exists():
        return True
    t = _LIB.stat().st_mtime
    return any(s.stat().st_mtime > t for s in _sources())


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/polar_abi.cu -> polar_code_b200/libpolar_b200.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return _LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libpolar_b200.so must be built on a box with the CUDA toolkit")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(_LIB), str(_CSRC / "polar_abi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return _LIB


if __name__ == "__main__":
    import sys
    build_library(force="--force" in sys.argv, verbose=True)
    print(_LIB)
