"""Batched engine object for the toy NR-LDPC family over the C-ABI (include/polar_b200.h, "NR LDPC").

One ``LdpcEngine`` = one ``pb200_ldpc`` handle = one (device, parity-check matrix).  All arithmetic is float64
on the GPU with the reference's operation order (dl_scl_polar/nr/ldpc/decode_nms.py:8-40), so results are
bit-identical to the reference's on the same LLRs.  There is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from .engine import _ptr, _stream, _u16, require_cuda


def build_h_matrix(bg: int, Z: int) -> np.ndarray:
    """basegraphs.py:39-42 + builder.py:20-30 (host side of the C-ABI): int8 [3Z, 6Z]."""
    lib = L.load()
    m, n = C.c_int(), C.c_int()
    L.check(lib.pb200_ldpc_build_h(int(bg), int(Z), None, C.byref(m), C.byref(n)))
    H = np.zeros((m.value, n.value), np.uint8)
    L.check(lib.pb200_ldpc_build_h(int(bg), int(Z), H.ctypes.data, C.byref(m), C.byref(n)))
    return H.astype(np.int8)


def parity_generator(H, k: int):
    """Host side of encode.py:52-66: (G, C) with parity = G @ payload (mod 2) and "no solution" iff C @ payload != 0.
    G: uint8 [n-k, k], C: uint8 [n_check, k]."""
    lib = L.load()
    Hu = np.ascontiguousarray(np.asarray(H) % 2, np.uint8)
    m, n = Hu.shape
    kw = max(1, (int(k) + 31) // 32)
    Gw = np.zeros((max(n - int(k), 0), kw), np.uint32)
    Cw = np.zeros((m, kw), np.uint32)
    nc = C.c_int()
    L.check(lib.pb200_ldpc_parity_generator(Hu.ctypes.data, m, n, int(k), Gw.ctypes.data, Cw.ctypes.data, C.byref(nc)))
    unpack = lambda W: ((W[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(W.shape[0], kw * 32)[:, :int(k)].astype(np.uint8)
    return unpack(Gw), unpack(Cw[: nc.value])


def layers(H):
    """Host side of the launch planning: (layer_ptr, lanes_per_frame) of the group-per-frame kernels."""
    lib = L.load()
    Hu = np.ascontiguousarray(np.asarray(H) % 2, np.uint8)
    m, n = Hu.shape
    lp = np.zeros(m + 1, np.int32)
    nl, g = C.c_int(), C.c_int()
    L.check(lib.pb200_ldpc_layers(Hu.ctypes.data, m, n, lp.ctypes.data, C.byref(nl), C.byref(g)))
    return lp[: nl.value + 1].copy(), g.value


class LdpcEngine:
    def __init__(self, H, device: int | None = None):
        require_cuda()
        self.lib = L.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        Hm = np.asarray(H)
        if Hm.ndim != 2:
            raise ValueError("H must be a 2D matrix")
        self.H = np.ascontiguousarray(Hm % 2 if Hm.dtype.kind in "iu" else Hm, np.uint8)
        self.m, self.n = (int(v) for v in self.H.shape)
        self.k = self.n - self.m
        h = C.c_void_p()
        L.check(self.lib.pb200_ldpc_create(C.byref(h), self.device, self.H.ctypes.data, self.m, self.n))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pb200_ldpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, a, dtype) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            return a.to(device=self.dev, dtype=dtype).contiguous()
        npdt = {torch.float64: np.float64, torch.uint8: np.uint8}[dtype]
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a)).astype(npdt, copy=False)).to(self.dev).contiguous()

    # ------------------------------------------------------------------ encoder / rate matching
    def encode(self, payload, want_status: bool = False):
        """encode.py:52-66 batched: payload[B,k] -> code[B,n] uint8 (any k < n)."""
        p = self._dev(payload, torch.uint8)
        if p.ndim != 2:
            raise ValueError("payload must be [B, k]")
        B, k = p.shape
        out = torch.empty((B, self.n), dtype=torch.uint8, device=self.dev)
        status = torch.zeros((B,), dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_ldpc_encode_batch(self._h, _ptr(p), int(k), _ptr(out), _ptr(status), B, _stream()))
        return (out, status) if want_status else out

    def rate_match(self, code, E: int) -> torch.Tensor:
        """rate_match.py:8-15 batched."""
        c = self._dev(code, torch.uint8)
        out = torch.empty((c.shape[0], int(E)), dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_ldpc_rate_match_batch(_ptr(c), c.shape[1], int(E), _ptr(out), c.shape[0], _stream()))
        return out

    def derate_match(self, llr, N: int | None = None) -> torch.Tensor:
        """rate_match.py:18-38 batched: llr[B,E] f64 -> [B,N] f64."""
        x = self._dev(llr, torch.float64)
        N = self.n if N is None else int(N)
        out = torch.empty((x.shape[0], N), dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_ldpc_derate_match_batch(_ptr(x), x.shape[1], N, _ptr(out), x.shape[0], _stream()))
        return out

    # ------------------------------------------------------------------ decoder
    def decode(self, llr, max_iter: int = 20, alpha: float = 0.8, early_stop: bool = True, want_posterior: bool = False) -> dict:
        """decode_nms.py:8-40 batched: llr[B,n] (or [B,E]: de-rate-matching fused) float64."""
        x = self._dev(llr, torch.float64)
        if x.ndim != 2:
            raise ValueError("llr must be [B, n]")
        B = x.shape[0]
        o = {"hard": torch.empty((B, self.n), dtype=torch.uint8, device=self.dev),
             "iters_used": torch.empty((B,), dtype=torch.int32, device=self.dev),
             "parity_ok": torch.empty((B,), dtype=torch.uint8, device=self.dev)}
        if want_posterior:
            o["posterior"] = torch.empty((B, self.n), dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_ldpc_decode_batch(self._h, _ptr(x), B, x.shape[1], int(max_iter), float(alpha),
                                                     int(bool(early_stop)), _ptr(o["hard"]), _ptr(o.get("posterior")),
                                                     _ptr(o["iters_used"]), _ptr(o["parity_ok"]), _stream()))
        return o

    # ------------------------------------------------------------------ Monte-Carlo
    def _cfg(self, *, k_payload, k_crc, E, max_iter, alpha, early_stop, crc_poly, noise_var, seed, stream_id, frame_begin,
             n_frames) -> L.LdpcSweepCfg:
        c = L.LdpcSweepCfg()
        c.k_payload, c.k_crc, c.E, c.max_iter, c.early_stop = int(k_payload), int(k_crc), int(E), int(max_iter), int(early_stop)
        c.alpha = float(alpha)
        c.crc_poly = crc_poly.encode() if crc_poly else None
        c.noise_var = float(noise_var)
        c.seed, c.stream_id = int(seed) & (2**64 - 1), int(stream_id)
        c.frame_begin, c.n_frames = int(frame_begin), int(n_frames)
        return c

    def configure_sweep(self, *, k_crc: int, E: int, max_iter: int = 20, alpha: float = 0.8, crc_poly: Optional[str] = None):
        """Fix the scheme parameters so that `sweep` has the call signature montecarlo.ber_point uses."""
        self._sw = dict(k_crc=int(k_crc), E=int(E), max_iter=int(max_iter), alpha=float(alpha), crc_poly=crc_poly)

    def sweep(self, counters: torch.Tensor, *, noise_var: float, n_frames: int, frame_begin: int = 0, seed: int = 0,
              stream_id: int = 0, k_payload: int | None = None, frame_bit_errors: Optional[torch.Tensor] = None,
              frame_work: Optional[torch.Tensor] = None, **_ignored) -> None:
        """Fused channel + NMS decode + counters (run_ber_sweep.py:112-181, scheme nr_ldpc); `counters` is ADDED to."""
        sw = self._sw
        frame_bit_errors = _u16(frame_bit_errors, "frame_bit_errors", int(n_frames))
        frame_work = _u16(frame_work, "frame_work", int(n_frames))
        kp = self.k - sw["k_crc"] if k_payload is None else int(k_payload)
        cfg = self._cfg(k_payload=kp, early_stop=1, noise_var=noise_var, seed=seed, stream_id=stream_id,
                        frame_begin=frame_begin, n_frames=n_frames, **sw)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_ldpc_sweep(self._h, C.byref(cfg), _ptr(counters), _ptr(frame_bit_errors),
                                              _ptr(frame_work), _stream()))

    def channel(self, *, noise_var: float, n_frames: int, frame_begin: int = 0, seed: int = 0, stream_id: int = 0,
                k_payload: int | None = None):
        """Channel only (same Philox stream as `sweep`): payload[B,kp] uint8, llr[B,E] float64."""
        sw = self._sw
        kp = self.k - sw["k_crc"] if k_payload is None else int(k_payload)
        cfg = self._cfg(k_payload=kp, early_stop=1, noise_var=noise_var, seed=seed, stream_id=stream_id,
                        frame_begin=frame_begin, n_frames=n_frames, **sw)
        payload = torch.empty((n_frames, kp), dtype=torch.uint8, device=self.dev)
        llr = torch.empty((n_frames, sw["E"]), dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            L.check(self.lib.pb200_ldpc_channel_batch(self._h, C.byref(cfg), _ptr(payload), _ptr(llr), _stream()))
        return payload, llr
