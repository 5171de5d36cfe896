"""Build libpolar_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

The kernel families are instantiated in separate translation units (csrc/k_*.cu) that compile in parallel;
csrc/polar_abi.cu holds the C-ABI and the small kernels.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_CSRC = _PKG / "csrc"
_OBJ = _CSRC / "obj"
_LIB = Path(os.environ["PB200_LIBRARY"]).resolve() if os.environ.get("PB200_LIBRARY") else _PKG / "libpolar_b200.so"
_EXTRA = os.environ.get("PB200_DEFINES", "").split()      # experiment knobs, e.g. -DPB_HSPLIT=6 (default build: none)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "177",
]


def library_path() -> Path:
    return _LIB


def _headers():
    return sorted(list(_CSRC.glob("*.cuh")) + list(_CSRC.glob("*.inl")) + list(_CSRC.glob("*.h")) +
                  [_PKG.parent / "include" / "polar_b200.h"])


def _units():
    return sorted(_CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not _LIB.exists():
        return True
    t = _LIB.stat().st_mtime
    return any(s.stat().st_mtime > t for s in _headers() + _units())


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu -> polar_code_b200/libpolar_b200.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return _LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libpolar_b200.so must be built on a box with the CUDA toolkit")
    obj_dir = _OBJ if not _EXTRA else _CSRC / ("obj_" + "_".join(x.strip("-").replace("=", "") for x in _EXTRA))
    obj_dir.mkdir(exist_ok=True)
    # *.inl files and the public header are included by the *_abi.cu units only; the kernel units depend on *.cuh / *.h
    inl_time = max(h.stat().st_mtime for h in _headers())
    hdr_time = max(h.stat().st_mtime for h in _headers() if h.suffix in (".cuh", ".h") and h.name != "polar_b200.h")

    def compile_one(src: Path) -> Path:
        obj = obj_dir / (src.stem + ".o")
        dep_time = inl_time if src.stem.endswith("_abi") else hdr_time
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, dep_time):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *_EXTRA, "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _units()))
    subprocess.check_call([nvcc, "-shared", "-o", str(_LIB), *[str(o) for o in objs]])
    return _LIB


if __name__ == "__main__":
    import sys
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(_LIB)
