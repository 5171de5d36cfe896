"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT the product).

Loads ``oracle/build/libpolar_oracle.so`` (built from ``polar_oracle.c`` and
``ldpc_oracle.c`` by ``oracle/Makefile``) and exposes NumPy-level helpers that mirror the reference's
functions.  Only ``tests/``, ``bench.py``'s CPU-baseline legs and
``__graft_entry__.smoke()`` may import this module, and only as the checker.

It also restates, with NumPy's own PCG64 ``Generator``, the per-frame channel
loops of the reference sweep CLIs so that the exact LLR stream of a published
CSV can be regenerated on a box where ``/root/reference`` does not exist:

* ``fer_sweep_frames``  -> ``dl_scl_polar/eval/run_fer_sweep.py:60-121``
* ``ber_sweep_frames``  -> ``dl_scl_polar/eval/run_ber_sweep.py:112-142,228-291``
* ``ldpc_ber_point``    -> the same loop for ``--scheme nr_ldpc`` (:134-136,258-271)
"""

from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "build" / "libpolar_oracle.so"
_lib = None

_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class _SclInfo(C.Structure):
    _fields_ = [("n_cand", C.c_int), ("best_idx", C.c_int), ("min_rel_gap", C.c_double), ("min_rel_gap_prune", C.c_double)]


class _DlInfo(C.Structure):
    _fields_ = [
        ("success", C.c_int), ("n_attempts", C.c_int), ("n_tried", C.c_int), ("n_cand", C.c_int),
        ("best_idx", C.c_int), ("min_rel_gap", C.c_double), ("min_rank_gap", C.c_double),
    ]


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc); building the checker is not using it."""
    newest = max((_HERE / f).stat().st_mtime for f in ("polar_oracle.c", "ldpc_oracle.c", "Makefile"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < newest:
        subprocess.check_call(["make", "-s", "-C", str(_HERE)])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.po_num_threads.restype = C.c_int
    return _lib


def _ptr(a, typ):
    return a.ctypes.data_as(C.POINTER(typ)) if a is not None else None


def _crc(crc):
    return crc.encode() if crc is not None else None


# ----------------------------------------------------------------------------
# polar/polar.py
# ----------------------------------------------------------------------------

def construct_info_set(N: int, K: int, method: str = "gaussian", design_snr_db: float = 2.5) -> np.ndarray:
    out = np.zeros(K, np.int32)
    m = {"gaussian": 0, "polarization": 1}[method]
    rc = lib().po_construct_info_set(C.c_int(N), C.c_int(K), C.c_int(m), C.c_double(design_snr_db), _ptr(out, C.c_int32))
    if rc != 0:
        raise ValueError("bad (N, K)")
    return out


def polar_transform(u: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(u, np.int8).copy()
    lib().po_polar_transform(_ptr(x, C.c_int8), C.c_int(x.size))
    return x


def encode(msg: np.ndarray, info_set: np.ndarray, N: int) -> np.ndarray:
    msg = np.ascontiguousarray(msg, np.int8)
    info_set = np.ascontiguousarray(info_set, np.int32)
    x = np.zeros(N, np.int8)
    rc = lib().po_encode(_ptr(msg, C.c_int8), C.c_int(msg.size), _ptr(info_set, C.c_int32), C.c_int(N), _ptr(x, C.c_int8))
    if rc != 0:
        raise ValueError("bad N")
    return x


def sc_decode(llr: np.ndarray, info_set: np.ndarray) -> np.ndarray:
    llr = np.ascontiguousarray(llr, np.float64)
    info_set = np.ascontiguousarray(info_set, np.int32)
    out = np.zeros(info_set.size, np.int8)
    rc = lib().po_sc_decode(_ptr(llr, C.c_double), C.c_int(llr.size), _ptr(info_set, C.c_int32), C.c_int(info_set.size), _ptr(out, C.c_int8))
    if rc != 0:
        raise ValueError("sc_decode failed")
    return out


# ----------------------------------------------------------------------------
# polar/crc.py
# ----------------------------------------------------------------------------

def attach_crc(msg: np.ndarray, poly: str) -> np.ndarray:
    msg = np.ascontiguousarray(msg, np.int8)
    out = np.zeros(msg.size + 64, np.int8)
    deg = lib().po_crc_attach(_ptr(msg, C.c_int8), C.c_int(msg.size), poly.encode(), _ptr(out, C.c_int8))
    if deg < 0:
        raise ValueError("bad polynomial")
    return out[: msg.size + deg].copy()


def check_crc(msg_crc: np.ndarray, poly: str) -> bool:
    msg_crc = np.ascontiguousarray(msg_crc, np.int8)
    rc = lib().po_crc_check(_ptr(msg_crc, C.c_int8), C.c_int(msg_crc.size), poly.encode())
    if rc < 0:
        raise ValueError("message too short / bad polynomial")
    return bool(rc)


# ----------------------------------------------------------------------------
# polar/scl.py, dlscl/flip.py, nr/polar (batched)
# ----------------------------------------------------------------------------

def scl_decode_batch(llr, info_set, M, crc=None, force=None, want_info_llrs=True, nthreads=0):
    """decode_scl over llr[B,N]; returns dict of arrays (cand[B,M,K] ... min_gap[B])."""
    llr = np.ascontiguousarray(np.atleast_2d(llr), np.float64)
    B, N = llr.shape
    info_set = np.ascontiguousarray(info_set, np.int32)
    K = info_set.size
    cand = np.zeros((B, M, K), np.int8)
    metrics = np.full((B, M), np.inf, np.float64)
    ill = np.zeros((B, M, K), np.float64) if want_info_llrs else None
    n_cand = np.zeros(B, np.int32)
    best = np.zeros(B, np.int32)
    gap = np.zeros(B, np.float64)
    if force is not None:
        force = np.ascontiguousarray(np.atleast_2d(force), np.int8)
        assert force.shape == (B, K)
    rc = lib().po_scl_decode_batch(
        _ptr(llr, C.c_double), C.c_int(B), C.c_int(N), _ptr(info_set, C.c_int32), C.c_int(K), C.c_int(M),
        _crc(crc), _ptr(force, C.c_int8), _ptr(cand, C.c_int8), _ptr(metrics, C.c_double),
        _ptr(ill, C.c_double), _ptr(n_cand, C.c_int32), _ptr(best, C.c_int32), _ptr(gap, C.c_double),
        C.c_int(nthreads))
    if rc == -2:
        raise ValueError("force_info_bits entries must be -1, 0, or 1")
    if rc != 0:
        raise ValueError("scl decode failed")
    best_bits = cand[np.arange(B), best]
    return {"cand": cand, "metrics": metrics, "info_llrs": ill, "n_cand": n_cand, "best_idx": best,
            "best_bits": best_bits, "min_gap": gap}


def sc_decode_batch(llr, info_set, nthreads=0):
    llr = np.ascontiguousarray(np.atleast_2d(llr), np.float64)
    B, N = llr.shape
    info_set = np.ascontiguousarray(info_set, np.int32)
    out = np.zeros((B, info_set.size), np.int8)
    rc = lib().po_sc_decode_batch(_ptr(llr, C.c_double), C.c_int(B), C.c_int(N), _ptr(info_set, C.c_int32),
                                  C.c_int(info_set.size), _ptr(out, C.c_int8), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("sc decode failed")
    return out


def dlscl_decode_batch(llr, info_set, M, retries, crc=None, beta=None, nthreads=0):
    """decode_with_retries over llr[B,N]."""
    llr = np.ascontiguousarray(np.atleast_2d(llr), np.float64)
    B, N = llr.shape
    info_set = np.ascontiguousarray(info_set, np.int32)
    K = info_set.size
    R = max(retries, 1)
    if beta is not None:
        beta = np.ascontiguousarray(beta, np.float32)
        assert beta.shape == (K, K)
    bits = np.zeros((B, K), np.int8)
    success = np.zeros(B, np.int32)
    n_att = np.zeros(B, np.int32)
    tried = np.full((B, R), -1, np.int32)
    n_tried = np.zeros(B, np.int32)
    gap = np.zeros(B, np.float64)
    rgap = np.zeros(B, np.float64)
    rc = lib().po_dlscl_decode_batch(
        _ptr(llr, C.c_double), C.c_int(B), C.c_int(N), _ptr(info_set, C.c_int32), C.c_int(K), C.c_int(M),
        C.c_int(retries), _crc(crc), _ptr(beta, C.c_float), _ptr(bits, C.c_int8), _ptr(success, C.c_int32),
        _ptr(n_att, C.c_int32), _ptr(tried, C.c_int32), _ptr(n_tried, C.c_int32), _ptr(gap, C.c_double),
        _ptr(rgap, C.c_double), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("dl-scl decode failed")
    return {"best_bits": bits, "success": success.astype(bool), "n_attempts": n_att, "tried": tried,
            "n_tried": n_tried, "min_gap": gap, "min_rank_gap": rgap}


def nr_decode_batch(llr_E, crc, N, info_set, M, nthreads=0):
    llr_E = np.ascontiguousarray(np.atleast_2d(llr_E), np.float64)
    B, E = llr_E.shape
    info_set = np.ascontiguousarray(info_set, np.int32)
    K = info_set.size
    bits = np.zeros((B, K), np.int8)
    ok = np.zeros(B, np.int32)
    gap = np.zeros(B, np.float64)
    rc = lib().po_nr_decode_batch(_ptr(llr_E, C.c_double), C.c_int(B), C.c_int(E), crc.encode(), C.c_int(N),
                                  _ptr(info_set, C.c_int32), C.c_int(K), C.c_int(M), _ptr(bits, C.c_int8),
                                  _ptr(ok, C.c_int32), _ptr(gap, C.c_double), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("nr decode failed")
    return {"best_bits": bits, "crc_pass": ok.astype(bool), "min_gap": gap}


def choose_flip_index(abs_l0, beta=None) -> int:
    abs_l0 = np.ascontiguousarray(abs_l0, np.float64)
    if beta is not None:
        beta = np.ascontiguousarray(beta, np.float32)
    return int(lib().po_choose_flip_index(_ptr(abs_l0, C.c_double), C.c_int(abs_l0.size), _ptr(beta, C.c_float)))


# ----------------------------------------------------------------------------
# nr/polar element-wise helpers
# ----------------------------------------------------------------------------

def subblock_interleave(v):
    v = np.ascontiguousarray(v, np.float64)
    out = np.zeros(((v.size + 31) // 32) * 32, np.float64)
    n = lib().po_subblock_interleave(_ptr(v, C.c_double), C.c_int(v.size), _ptr(out, C.c_double))
    return out[:n]


def subblock_deinterleave(v, original_len):
    v = np.ascontiguousarray(v, np.float64)
    out = np.zeros(original_len, np.float64)
    lib().po_subblock_deinterleave(_ptr(v, C.c_double), C.c_int(v.size), C.c_int(original_len), _ptr(out, C.c_double))
    return out


def derate_match(v, N):
    v = np.ascontiguousarray(v, np.float64)
    out = np.zeros(N, np.float64)
    lib().po_derate_match(_ptr(v, C.c_double), C.c_int(v.size), C.c_int(N), _ptr(out, C.c_double))
    return out


def rate_match(bits, E):
    """nr/polar/rate_match.py:8-16"""
    N = bits.size
    if E <= N:
        return bits[:E]
    return np.tile(bits, (E + N - 1) // N)[:E]


# ----------------------------------------------------------------------------
# Channel loops of the reference CLIs, restated with NumPy's PCG64 Generator.
# ----------------------------------------------------------------------------

def fer_sweep_frames(snr_db: float, frames: int, seed: int = 0, N: int = 128, K: int = 64, crc_bits: int = 24,
                     crc_poly: str = "0x1864CFB", include_uncoded: bool = False):
    """run_fer_sweep.py:60-121 for one SNR point.

    Returns msgs[frames,K] int8, llr[frames,N] float64 and, if include_uncoded,
    the per-frame uncoded bit-error counts (the uncoded draws sit between coded
    frames in the same stream, :111-121, so they must be consumed either way).
    """
    info_set = construct_info_set(N, K)
    payload_bits = K - crc_bits
    rng = np.random.default_rng(seed + int(snr_db * 10))          # :61
    ebno = 10 ** (snr_db / 10.0)                                 # :62
    nv = 1.0 / (2.0 * (K / N) * ebno)                            # :63-64
    sg = math.sqrt(nv)
    nvu = 1.0 / (2.0 * ebno)                                     # :66
    sgu = math.sqrt(nvu)
    msgs = np.zeros((frames, K), np.int8)
    llrs = np.zeros((frames, N), np.float64)
    unc = np.zeros(frames, np.int64)
    for f in range(frames):
        payload = rng.integers(0, 2, size=payload_bits, dtype=np.int8)   # :80
        msg = attach_crc(payload, crc_poly)                               # :81
        code = encode(msg, info_set, N)                                   # :82
        symbols = 1.0 - 2.0 * code                                        # :24-25
        noise = rng.normal(0.0, sg, size=symbols.shape)                   # :85
        llrs[f] = 2.0 * (symbols + noise) / nv                            # :86-87
        msgs[f] = msg
        if include_uncoded:                                               # :111-121
            us = 1.0 - 2.0 * payload
            nu = rng.normal(0.0, sgu, size=us.shape)
            lu = 2.0 * (us + nu) / nvu
            unc[f] = int(np.count_nonzero((lu < 0).astype(np.int8) != payload))
    return msgs, llrs, unc


def noise_var_ber(EbN0_dB: float, payload_bits: int, coded_bits: int) -> float:
    """run_ber_sweep.py:105-109"""
    ebno = 10 ** (EbN0_dB / 10.0)
    return 1.0 / (2.0 * (ebno * (payload_bits / coded_bits)))


def ber_frame(rng, scheme: str, K_payload: int, K_crc: int, crc_poly: str, N: int, E: int, info_set, noise_var: float):
    """One iteration of run_ber_sweep.py:127-142 up to the LLR; returns payload, llr."""
    payload = rng.integers(0, 2, size=K_payload, dtype=np.int8)
    msg = payload if K_crc == 0 and scheme != "nr_polar_scl" else attach_crc(payload, crc_poly)
    code = encode(msg, info_set, N)
    if scheme == "nr_polar_scl":                     # scl_nr.py:31-35
        ilv = subblock_interleave(code.astype(np.float64))
        tx = rate_match(ilv, E)
    else:
        tx = code.astype(np.float64)
    symbols = 1.0 - 2.0 * tx
    noise = rng.normal(0.0, math.sqrt(noise_var), size=symbols.shape)
    return payload, 2.0 * (symbols + noise) / noise_var


# ----------------------------------------------------------------------------
# nr/ldpc (toy NR-LDPC family)
# ----------------------------------------------------------------------------

def ldpc_build_h(bg: int, Z: int) -> np.ndarray:
    """basegraphs.py:39-42 + builder.py:20-30"""
    if Z <= 0:
        raise ValueError("Z must be positive")
    H = np.zeros((3 * Z, 6 * Z), np.int8)
    m, n = C.c_int(), C.c_int()
    if lib().po_ldpc_build_h(C.c_int(bg), C.c_int(Z), _ptr(H, C.c_int8), C.byref(m), C.byref(n)) != 0:
        raise ValueError(f"Unknown base graph: {bg}")
    return H


def ldpc_encode(payload: np.ndarray, H: np.ndarray) -> np.ndarray:
    """encode.py:52-66"""
    payload = np.ascontiguousarray(payload, np.int8)
    H = np.ascontiguousarray(H, np.int8)
    m, n = H.shape
    out = np.zeros(n, np.int8)
    rc = lib().po_ldpc_encode(_ptr(payload, C.c_int8), C.c_int(payload.size), _ptr(H, C.c_int8), C.c_int(m), C.c_int(n),
                              _ptr(out, C.c_int8))
    if rc == -1:
        raise ValueError("Parity-check matrix too small for payload length")
    if rc != 0:
        raise ValueError("Linear system over GF(2) has no solution")
    return out


def ldpc_rate_match(codeword: np.ndarray, E: int) -> np.ndarray:
    """rate_match.py:8-15"""
    N = codeword.size
    if E <= N:
        return codeword[:E]
    return np.tile(codeword, (E + N - 1) // N)[:E]


def ldpc_derate_match(llr: np.ndarray, N: int) -> np.ndarray:
    """rate_match.py:18-38"""
    llr = np.ascontiguousarray(llr, np.float64)
    out = np.zeros(N, np.float64)
    lib().po_ldpc_derate(_ptr(llr, C.c_double), C.c_int(llr.size), C.c_int(N), _ptr(out, C.c_double))
    return out


def ldpc_decode_batch(llr, H, max_iter: int = 20, alpha: float = 0.8, early_stop: bool = True):
    """decode_nms.py:8-40 over llr[B,n] -> hard[B,n] int8, iters_used[B], parity_ok[B]."""
    llr = np.ascontiguousarray(np.atleast_2d(llr), np.float64)
    H = np.ascontiguousarray(H, np.int8)
    m, n = H.shape
    B = llr.shape[0]
    if llr.shape[1] != n:
        raise ValueError("llr length mismatch")
    hard = np.zeros((B, n), np.int8)
    iters = np.zeros(B, np.int32)
    ok = np.zeros(B, np.int32)
    lib().po_ldpc_decode_batch(_ptr(llr, C.c_double), C.c_int(B), _ptr(H, C.c_int8), C.c_int(m), C.c_int(n),
                               C.c_int(max_iter), C.c_double(alpha), C.c_int(int(early_stop)), _ptr(hard, C.c_int8),
                               _ptr(iters, C.c_int32), _ptr(ok, C.c_int32))
    return {"hard": hard, "iters_used": iters, "parity_ok": ok.astype(bool)}


def ldpc_ber_point(rng, EbN0_dB: float, *, K_payload: int, K_crc: int, crc_poly: str, H: np.ndarray, E: int,
                   max_iter: int, alpha: float, err_cap: int, bits_cap: float):
    """One Eb/N0 point of run_ber_sweep.py:112-181 for --scheme nr_ldpc (encoder/decoder of :258-271), consuming
    `rng` exactly like the reference.  Returns (bits_total, bit_errors, frame_errors, frames, work_sum)."""
    n = H.shape[1]
    k = n - H.shape[0]
    nv = noise_var_ber(EbN0_dB, K_payload, E)
    sg = math.sqrt(nv)
    bits_total = bit_errors = frame_errors = frames = 0
    work = 0.0
    while bit_errors < err_cap and bits_total < bits_cap:
        payload = rng.integers(0, 2, size=K_payload, dtype=np.int8)
        message = payload if K_crc == 0 else attach_crc(payload, crc_poly)
        tx = ldpc_rate_match(ldpc_encode(message[:k], H), E)
        symbols = 1.0 - 2.0 * tx
        noise = rng.normal(0.0, sg, size=symbols.shape)
        llr = 2.0 * (symbols + noise) / nv
        res = ldpc_decode_batch(ldpc_derate_match(llr, n), H, max_iter, alpha)
        cand = res["hard"][0][: K_payload + K_crc]
        be = int(np.count_nonzero(payload != cand[:K_payload]))
        bits_total += K_payload
        bit_errors += be
        frame_errors += int(be > 0)
        frames += 1
        work += float(res["iters_used"][0])
    return bits_total, bit_errors, frame_errors, frames, work
