"""Batched engine object over the C-ABI: PyTorch only owns device memory and streams.

One ``PolarEngine`` = one ``pb200_engine`` handle = one (device, N, info_set, CRC polynomial).
Every method takes/returns torch CUDA tensors (or NumPy arrays, which are staged through torch) and
enqueues exactly the kernels of ``csrc/``; there is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L

_METHODS = {"gaussian": 0, "polarization": 1}


def construct_info_set(N: int, K: int, method: str = "gaussian", design_snr_db: float = 2.5) -> np.ndarray:
    """polar/polar.py:85-103 (host float64 math inside the C-ABI)."""
    lib = L.load()
    if method not in _METHODS:
        raise ValueError(f"Unsupported construction method: {method}")
    out = np.zeros(max(int(K), 1), np.int32)
    L.check(lib.pb200_construct_info_set(int(N), int(K), _METHODS[method], float(design_snr_db), out.ctypes.data))
    return out[:K]


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _u16(t: Optional[torch.Tensor], name: str, n: int) -> Optional[torch.Tensor]:
    """Per-frame counter buffers are u16[n_frames] on the device side (exact counts); torch buffers are int16/uint16."""
    if t is None:
        return None
    if t.element_size() != 2 or t.numel() < n or not t.is_contiguous() or not t.is_cuda:
        raise ValueError(f"{name} must be a contiguous 16-bit CUDA tensor with at least n_frames elements")
    return t


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("polar_code_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


class PolarEngine:
    def __init__(self, N: int, info_set, crc_poly: Optional[str] = None, device: int | None = None):
        require_cuda()
        self.lib = L.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        self.N = int(N)
        self.info_set = np.ascontiguousarray(np.asarray(info_set), np.int32)
        if self.info_set.ndim != 1:
            raise ValueError("info_set must be a 1D array")
        self.K = int(self.info_set.size)
        self.crc_poly = crc_poly
        self.E = 0
        h = C.c_void_p()
        L.check(self.lib.pb200_create(C.byref(h), self.device, self.N, self.info_set.ctypes.data, self.K,
                                      crc_poly.encode() if crc_poly is not None else None))
        self._h = h
        self.xw = max(self.N // 32, 1)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _dev(self, a, dtype) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.dev, dtype=dtype)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)).astype(
                {torch.float32: np.float32, torch.uint8: np.uint8, torch.int8: np.int8}[dtype], copy=False)).to(self.dev)
        return t.contiguous()

    def in_len(self) -> int:
        return self.E if self.E else self.N

    def set_rate_matching(self, E: int) -> None:
        """Fuse nr/polar/rate_match.py:19-39 + interleaver.py:26-37 into the LLR load (E=0: off)."""
        L.check(self.lib.pb200_set_rate_matching(self._h, int(E)))
        self.E = int(E)

    def kernel_info(self, M: int) -> dict:
        v = [C.c_int() for _ in range(4)]
        L.check(self.lib.pb200_kernel_info(self._h, int(M), *[C.byref(x) for x in v]))
        return {"warps_per_cta": v[0].value, "ctas_per_sm": v[1].value, "smem_bytes": v[2].value, "regs": v[3].value}

    # ------------------------------------------------------------------ encoder / CRC
    def encode(self, msg) -> torch.Tensor:
        """polar.py:106-119 batched: msg[B,K] -> code[B,N] (uint8)."""
        m = self._dev(msg, torch.uint8)
        if m.ndim != 2 or m.shape[1] != self.K:
            raise ValueError(f"msg_bits must have length {self.K}")
        out = torch.empty((m.shape[0], self.N), dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_encode_batch(self._h, _ptr(m), _ptr(out), m.shape[0], _stream()))
        return out

    def nr_encode(self, payload, E: int) -> torch.Tensor:
        """scl_nr.py:23-35 batched: payload[B,Kp] -> tx[B,E] int8 (requires set_rate_matching(E))."""
        p = self._dev(payload, torch.uint8)
        out = torch.empty((p.shape[0], int(E)), dtype=torch.int8, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_nr_encode_batch(self._h, _ptr(p), _ptr(out), p.shape[0], int(E), _stream()))
        return out

    # ------------------------------------------------------------------ decoders
    def sc_decode(self, llr) -> torch.Tensor:
        """polar.py:130-168 batched: llr[B,in_len] f32 -> bits[B,K] uint8."""
        x = self._dev(llr, torch.float32)
        B = x.shape[0]
        out = torch.empty((B, self.K), dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_sc_decode_batch(self._h, _ptr(x), B, x.shape[1], _ptr(out), _stream()))
        return out

    def scl_decode(self, llr, M: int, force=None, want=("cand", "metrics", "n_cand", "best_idx", "best_bits",
                                                         "crc_ok", "flags")) -> dict:
        """scl.py:108-209 batched.  `want` selects the outputs to materialise."""
        x = self._dev(llr, torch.float32)
        if x.ndim != 2:
            raise ValueError("llr must be [B, N]")
        B, K, M = x.shape[0], self.K, int(M)
        if M <= 0:
            raise ValueError("List size M must be positive")
        f = None
        if force is not None:
            f = self._dev(force, torch.int8)
            if f.shape != (B, K):
                raise ValueError("force_info_bits length must match info_set")
        o = {}
        mk = lambda shape, dt, fill=None: (torch.empty(shape, dtype=dt, device=self.dev) if fill is None
                                           else torch.full(shape, fill, dtype=dt, device=self.dev))
        if "cand" in want: o["cand"] = mk((B, M, K), torch.uint8, 0)
        if "metrics" in want: o["metrics"] = mk((B, M), torch.float64, float("inf"))
        if "info_llrs" in want: o["info_llrs"] = mk((B, M, K), torch.float32, 0.0)
        if "n_cand" in want: o["n_cand"] = mk((B,), torch.int32)
        if "best_idx" in want: o["best_idx"] = mk((B,), torch.int32)
        if "best_bits" in want: o["best_bits"] = mk((B, K), torch.uint8)
        if "best_words" in want: o["best_words"] = mk((B, self.xw), torch.int32)
        if "crc_ok" in want: o["crc_ok"] = mk((B,), torch.uint8)
        if "flags" in want: o["flags"] = mk((B,), torch.int32)
        so = L.SclOut(**{k: (o[k].data_ptr() if k in o else None) for k, _ in L.SclOut._fields_})
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_scl_decode_batch(self._h, _ptr(x), B, x.shape[1], _ptr(f), M, C.byref(so), _stream()))
        return o

    def dlscl_decode(self, llr, M: int, retries: int, beta=None) -> dict:
        """flip.py:65-141 batched; returns the last attempt per frame."""
        x = self._dev(llr, torch.float32)
        B, K = x.shape[0], self.K
        R = max(int(retries), 1)
        b = None
        if beta is not None:
            b = self._dev(beta, torch.float32)
            if b.ndim != 2 or b.shape[0] != b.shape[1] or b.shape[0] != K:
                raise ValueError("beta must be a square matrix matching abs_l0 length")
        o = {
            "best_bits": torch.empty((B, K), dtype=torch.uint8, device=self.dev),
            "best_words": torch.empty((B, self.xw), dtype=torch.int32, device=self.dev),
            "success": torch.empty((B,), dtype=torch.uint8, device=self.dev),
            "n_attempts": torch.empty((B,), dtype=torch.int32, device=self.dev),
            "tried": torch.full((B, R), -1, dtype=torch.int32, device=self.dev),
            "flags": torch.empty((B,), dtype=torch.int32, device=self.dev),
        }
        do = L.DlOut(**{k: o[k].data_ptr() for k, _ in L.DlOut._fields_})
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_dlscl_decode_batch(self._h, _ptr(x), B, x.shape[1], int(M), int(retries), _ptr(b),
                                                      C.byref(do), _stream()))
        return o

    def scl_decode_host(self, llr_host: torch.Tensor, M: int, best_bits: torch.Tensor, crc_ok: torch.Tensor,
                        flags: torch.Tensor) -> None:
        """Host buffers in / out (pinned for full speed); chunked copy/compute overlap inside the library.
        `llr_host` is float32 [B, in_len], or float16 (optional ingest format: rows are widened exactly on load)."""
        B = llr_host.shape[0]
        if llr_host.dtype not in (torch.float32, torch.float16) or llr_host.is_cuda or not llr_host.is_contiguous():
            raise ValueError("llr_host must be a contiguous float32 / float16 host tensor")
        fn = self.lib.pb200_scl_decode_host if llr_host.dtype == torch.float32 else self.lib.pb200_scl_decode_host_f16
        L.check(fn(self._h, C.c_void_p(llr_host.data_ptr()), B, llr_host.shape[1], int(M),
                                               C.c_void_p(best_bits.data_ptr()), C.c_void_p(crc_ok.data_ptr()),
                                               C.c_void_p(flags.data_ptr())))

    # ------------------------------------------------------------------ Monte-Carlo
    def _cfg(self, **kw) -> L.SweepCfg:
        c = L.SweepCfg()
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def sweep(self, counters: torch.Tensor, *, M: int, noise_var: float, n_frames: int, frame_begin: int = 0,
              seed: int = 0, stream_id: int = 0, retries: int = -1, run_scl: bool = True, k_payload: int | None = None,
              frame_error_mode: int = 0, bit_error_span: int | None = None, include_uncoded: bool = False,
              noise_var_uncoded: float = 1.0, beta=None, frame_bit_errors: Optional[torch.Tensor] = None,
              frame_work: Optional[torch.Tensor] = None) -> None:
        """Fused channel + decode + counters (run_fer_sweep.py:60-121 / run_ber_sweep.py:112-181).
        `counters` is an int64[16] CUDA tensor that is ADDED to; `frame_bit_errors` / `frame_work` are optional 16-bit
        per-frame outputs (bit errors of the last decoder run, attempts-1) for the adaptive stop."""
        b = self._dev(beta, torch.float32) if beta is not None else None
        frame_bit_errors = _u16(frame_bit_errors, "frame_bit_errors", int(n_frames))
        frame_work = _u16(frame_work, "frame_work", int(n_frames))
        kp = self.K if k_payload is None else int(k_payload)
        cfg = self._cfg(M=int(M), retries=int(retries), run_scl=int(run_scl), k_payload=kp, E=int(self.E),
                        frame_error_mode=int(frame_error_mode),
                        bit_error_span=int(self.K if bit_error_span is None else bit_error_span),
                        include_uncoded=int(include_uncoded), noise_var=float(noise_var),
                        noise_var_uncoded=float(noise_var_uncoded), seed=int(seed) & (2**64 - 1), stream_id=int(stream_id),
                        frame_begin=int(frame_begin), n_frames=int(n_frames))
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_sweep(self._h, C.byref(cfg), _ptr(b), _ptr(counters), _ptr(frame_bit_errors),
                                         _ptr(frame_work), _stream()))

    def channel(self, *, noise_var: float, n_frames: int, frame_begin: int = 0, seed: int = 0, stream_id: int = 0,
                k_payload: int | None = None, want_msg: bool = True):
        """Channel only (same Philox stream as `sweep`): returns msg[B,K] uint8 (or None) and llr[B,in_len] f32."""
        kp = self.K if k_payload is None else int(k_payload)
        cfg = self._cfg(M=1, retries=-1, run_scl=1, k_payload=kp, E=int(self.E), frame_error_mode=0,
                        bit_error_span=self.K, include_uncoded=0, noise_var=float(noise_var), noise_var_uncoded=1.0,
                        seed=int(seed) & (2**64 - 1), stream_id=int(stream_id), frame_begin=int(frame_begin),
                        n_frames=int(n_frames))
        msg = torch.empty((n_frames, self.K), dtype=torch.uint8, device=self.dev) if want_msg else None
        llr = torch.empty((n_frames, self.in_len()), dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_channel_batch(self._h, C.byref(cfg), _ptr(msg), _ptr(llr), _stream()))
        return msg, llr


# ---------------------------------------------------------------------------- code-independent batched helpers
def crc_attach(msg, poly: str, device=None) -> torch.Tensor:
    """crc.py:19-37 batched: msg[B,L] -> out[B,L+deg] uint8."""
    require_cuda()
    lib = L.load()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    m = (msg if isinstance(msg, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(msg, np.uint8))).to(dev, torch.uint8).contiguous()
    if not poly:
        raise ValueError("CRC polynomial string must be non-empty")
    deg = int(poly, 16).bit_length() - 1
    if deg <= 0:
        raise ValueError("Polynomial degree must be positive")
    out = torch.empty((m.shape[0], m.shape[1] + deg), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.pb200_crc_attach_batch(poly.encode(), _ptr(m), _ptr(out), m.shape[0], m.shape[1], _stream()))
    return out


def crc_check(msg, poly: str, device=None) -> torch.Tensor:
    """crc.py:40-56 batched: msg[B,L] -> ok[B] uint8."""
    require_cuda()
    lib = L.load()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    m = (msg if isinstance(msg, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(msg, np.uint8))).to(dev, torch.uint8).contiguous()
    out = torch.empty((m.shape[0],), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.pb200_crc_check_batch(poly.encode(), _ptr(m), _ptr(out), m.shape[0], m.shape[1], _stream()))
    return out


def choose_flip_index(abs_l0, beta=None, device=None) -> torch.Tensor:
    """flip.py:13-27 batched over rows."""
    require_cuda()
    lib = L.load()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    a = torch.as_tensor(np.asarray(abs_l0, np.float32) if not isinstance(abs_l0, torch.Tensor) else abs_l0).to(dev, torch.float32).contiguous()
    b = None
    if beta is not None:
        b = torch.as_tensor(np.asarray(beta, np.float32) if not isinstance(beta, torch.Tensor) else beta).to(dev, torch.float32).contiguous()
    out = torch.empty((a.shape[0],), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.pb200_choose_flip_index_batch(_ptr(a), _ptr(b), _ptr(out), a.shape[0], a.shape[1], _stream()))
    return out
