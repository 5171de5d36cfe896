"""ctypes binding of the C-ABI declared in include/polar_b200.h.

The product path has no CPU fallback: if the shared library is missing this module raises at import
of the binding (``load()``), and ``pb200_create`` fails when no CUDA device is visible.
"""

from __future__ import annotations

import ctypes as C

from .build import library_path

OK, EINVAL, ECUDA, ERANGE, ENOSUP = 0, -1, -2, -3, -4
NCOUNTERS = 16
FLAG_NEAR_TIE, FLAG_RANK_TIE, FLAG_BAD_FORCE = 1, 2, 4

_vp = C.c_void_p
_i64 = C.c_int64


class SclOut(C.Structure):
    _fields_ = [(n, _vp) for n in ("cand", "metrics", "info_llrs", "n_cand", "best_idx", "best_bits", "best_words",
                                   "crc_ok", "flags")]


class DlOut(C.Structure):
    _fields_ = [(n, _vp) for n in ("best_bits", "best_words", "success", "n_attempts", "tried", "flags")]


class SweepCfg(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("retries", C.c_int), ("run_scl", C.c_int), ("k_payload", C.c_int), ("E", C.c_int),
        ("frame_error_mode", C.c_int), ("bit_error_span", C.c_int), ("include_uncoded", C.c_int),
        ("noise_var", C.c_double), ("noise_var_uncoded", C.c_double), ("seed", C.c_uint64),
        ("stream_id", C.c_uint32), ("frame_begin", _i64), ("n_frames", _i64),
    ]


class LdpcSweepCfg(C.Structure):
    _fields_ = [
        ("k_payload", C.c_int), ("k_crc", C.c_int), ("E", C.c_int), ("max_iter", C.c_int), ("early_stop", C.c_int),
        ("alpha", C.c_double), ("crc_poly", C.c_char_p), ("noise_var", C.c_double), ("seed", C.c_uint64),
        ("stream_id", C.c_uint32), ("frame_begin", _i64), ("n_frames", _i64),
    ]


# every symbol include/polar_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pb200_last_error": (C.c_char_p, []),
    "pb200_version": (C.c_int, []),
    "pb200_device_count": (C.c_int, []),
    "pb200_construct_info_set": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _vp]),
    "pb200_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, C.c_int, C.c_char_p]),
    "pb200_destroy": (None, [_vp]),
    "pb200_set_rate_matching": (C.c_int, [_vp, C.c_int]),
    "pb200_encode_batch": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "pb200_crc_attach_batch": (C.c_int, [C.c_char_p, _vp, _vp, _i64, C.c_int, _vp]),
    "pb200_crc_check_batch": (C.c_int, [C.c_char_p, _vp, _vp, _i64, C.c_int, _vp]),
    "pb200_nr_encode_batch": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, _vp]),
    "pb200_sc_decode_batch": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp]),
    "pb200_scl_decode_batch": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, C.c_int, C.POINTER(SclOut), _vp]),
    "pb200_dlscl_decode_batch": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(DlOut), _vp]),
    "pb200_choose_flip_index_batch": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, _vp]),
    "pb200_scl_decode_host": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp]),
    "pb200_scl_decode_host_f16": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp]),
    "pb200_sweep": (C.c_int, [_vp, C.POINTER(SweepCfg), _vp, _vp, _vp, _vp, _vp]),
    "pb200_channel_batch": (C.c_int, [_vp, C.POINTER(SweepCfg), _vp, _vp, _vp]),
    "pb200_debug_bin_stats": (C.c_int, [_vp, _vp]),
    "pb200_kernel_info": (C.c_int, [_vp, C.c_int] + [C.POINTER(C.c_int)] * 4),
    "pb200_ldpc_build_h": (C.c_int, [C.c_int, C.c_int, _vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pb200_ldpc_create": (C.c_int, [C.POINTER(_vp), C.c_int, _vp, C.c_int, C.c_int]),
    "pb200_ldpc_destroy": (None, [_vp]),
    "pb200_ldpc_parity_generator": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.POINTER(C.c_int)]),
    "pb200_ldpc_layers": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pb200_ldpc_encode_batch": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _i64, _vp]),
    "pb200_ldpc_rate_match_batch": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _i64, _vp]),
    "pb200_ldpc_derate_match_batch": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _i64, _vp]),
    "pb200_ldpc_decode_batch": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.c_double, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "pb200_ldpc_sweep": (C.c_int, [_vp, C.POINTER(LdpcSweepCfg), _vp, _vp, _vp, _vp]),
    "pb200_ldpc_channel_batch": (C.c_int, [_vp, C.POINTER(LdpcSweepCfg), _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load libpolar_b200.so; raises RuntimeError if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not path.exists():
            raise RuntimeError(
                f"{path} is missing: build it with `python -m polar_code_b200.build` "
                "(the polar_b200 engine has no CPU fallback)")
        lib = C.CDLL(str(path))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().pb200_last_error().decode()


def check(rc: int) -> None:
    """Map C-ABI status codes onto the exceptions the reference raises."""
    if rc == OK:
        return
    msg = last_error()
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ERANGE:
        raise IndexError(msg)
    if rc == ENOSUP:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
