"""CPU: the C-ABI library loads and exports every symbol include/polar_b200.h declares (no compute calls)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "polar_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from polar_code_b200 import _lib
    from polar_code_b200.build import build_library
    build_library()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/polar_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding and header disagree"
    assert lib.pb200_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to run (it must never fall back to a CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from polar_code_b200.engine import PolarEngine
    import numpy as np
    with pytest.raises(RuntimeError):
        PolarEngine(8, np.array([3, 5, 6, 7], np.int32), None)
    from polar_code_b200.ldpc import LdpcEngine, build_h_matrix
    with pytest.raises(RuntimeError):
        LdpcEngine(build_h_matrix(2, 2))               # building H is host code, running the code is not
    from polar_code_b200.dl_scl_polar.nr.ldpc import decode_ldpc_nms
    with pytest.raises(RuntimeError):
        decode_ldpc_nms(np.zeros(12), build_h_matrix(2, 2))


def test_host_side_construction_matches_oracle():
    """construct_info_set runs on the host inside the C-ABI (float64, polar.py:85-103)."""
    import numpy as np
    from oracle import oracle as O
    from polar_code_b200.engine import construct_info_set
    for N, K in [(128, 64), (128, 88), (16, 12), (256, 128), (512, 200), (2, 1), (64, 64)]:
        for method in ("gaussian", "polarization"):
            assert np.array_equal(construct_info_set(N, K, method), O.construct_info_set(N, K, method))
    with pytest.raises(ValueError):
        construct_info_set(100, 10)
    with pytest.raises(ValueError):
        construct_info_set(16, 17)
    with pytest.raises(ValueError):
        construct_info_set(16, 4, "bogus")
