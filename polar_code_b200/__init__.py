"""polar_code_b200 -- B200-native (sm_100a) frame-parallel polar decoding engine.

Layout:
  csrc/           hand-written CUDA kernels + the C-ABI (include/polar_b200.h) -> libpolar_b200.so
  _lib.py         ctypes binding of the C-ABI (fails loudly when the library or a GPU is missing)
  engine.py       batched engine object over torch device buffers
  dl_scl_polar/   drop-in mirror of the reference's Python interface (same module paths and signatures)
"""

from .build import build_library, library_path  # noqa: F401

__all__ = ["build_library", "library_path"]
__version__ = "0.1.0"
