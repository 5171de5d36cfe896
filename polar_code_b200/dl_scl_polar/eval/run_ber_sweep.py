"""Unified payload-BER sweep CLI (polar_scl, dl_scl, nr_polar_scl, nr_ldpc) on the B200 engine.

Same flags, row keys and CSV columns as the reference (dl_scl_polar/eval/run_ber_sweep.py:197-317).  The adaptive
loop `while bit_errors < err_cap and bits_total < bits_cap` (:127) is reproduced exactly in global frame order on
batched GPU chunks (polar_code_b200/montecarlo.py adaptive_cut).  `--scheme nr_ldpc` (:258-271) runs the fused
float64 LDPC sweep kernel of polar_code_b200/csrc/ldpc_kernels.cuh through the same adaptive loop.
"""

from __future__ import annotations

import argparse
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterable, List, Optional

import numpy as np

from .. import config as global_config
from ..utils.seeding import seed_all
from ..polar.polar import construct_info_set
from .._engines import engine_for
from ._plot import semilogy_plot
from polar_code_b200 import montecarlo as mc

HEADER = ["scheme", "code", "N_or_E", "K_payload", "K_crc", "rate", "params", "EbN0_dB", "bits_total", "bit_errors",
          "ber", "fer", "avg_work"]


@dataclass
class SimulationStats:
    """Running totals of one Eb/N0 point (run_ber_sweep.py:36-62)."""
    bits_total: int = 0
    bit_errors: int = 0
    frame_errors: int = 0
    work_sum: float = 0.0
    frames: int = 0

    def update(self, bit_err: int, work: float, frame_error: bool, payload_len: int) -> None:
        self.bits_total += payload_len
        self.bit_errors += bit_err
        self.work_sum += work
        self.frames += 1
        self.frame_errors += int(bool(frame_error))

    def row(self) -> Dict[str, float]:
        nan = float("nan")
        return {"bits_total": self.bits_total, "bit_errors": self.bit_errors,
                "ber": self.bit_errors / self.bits_total if self.bits_total > 0 else nan,
                "fer": self.frame_errors / self.frames if self.frames > 0 else nan,
                "avg_work": self.work_sum / self.frames if self.frames > 0 else 0.0}


def _payload_bit_errors(payload: np.ndarray, candidate: Optional[np.ndarray], K_payload: int) -> int:
    """Errors on the first K_payload bits only; a missing candidate costs all of them (run_ber_sweep.py:77-82)."""
    if candidate is None:
        return int(K_payload)
    if candidate.size < K_payload:
        raise ValueError("Candidate bits shorter than payload")
    return int(np.count_nonzero(payload != candidate[:K_payload]))


def _noise_params(EbN0_dB: float, payload_bits: int, coded_bits: int) -> float:
    return mc.ber_noise_var(EbN0_dB, payload_bits, coded_bits)


_FLAGS = [
    ("--scheme", dict(required=True, choices=["polar_scl", "dl_scl", "nr_polar_scl", "nr_ldpc"], help="Coding scheme")),
    ("--K_payload", dict(type=int, required=True, help="Payload bits per frame")),
    ("--K_crc", dict(type=int, required=True, help="CRC bits per frame")),
    ("--E", dict(type=int, required=True, help="Coded bits transmitted")),
    ("--N", dict(type=int, help="Polar length before rate match (defaults to E)")),
    ("--crc_poly", dict(type=str, default=global_config.DEFAULTS.crc_poly)),
    ("--M", dict(type=int, default=4, help="List size for polar decoders")),
    ("--retries", dict(type=int, default=8, help="Retries for DL-SCL")),
    ("--beta", dict(type=str, help="Path to beta matrix (DL-SCL)")),
    ("--ilv_mode", dict(type=str, default="default")),
    ("--bg", dict(type=int, default=2, help="LDPC base graph")),
    ("--Z", dict(type=int, default=2, help="LDPC lifting size")),
    ("--max_iter", dict(type=int, default=20)),
    ("--alpha", dict(type=float, default=0.8)),
    ("--EbN0_lo", dict(type=float, required=True)),
    ("--EbN0_hi", dict(type=float, required=True)),
    ("--EbN0_step", dict(type=float, default=0.5)),
    ("--bits_cap", dict(type=float, default=1e7)),
    ("--err_cap", dict(type=int, default=1000)),
    ("--seed", dict(type=int, default=0)),
    ("--out", dict(type=str, required=True, help="CSV output path")),
    ("--plot", dict(type=str, help="Optional plot path")),
]


def parse_args(argv: Optional[Iterable[str]] = None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description="BER/FER sweep across schemes")
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)
    args = parser.parse_args(list(argv) if argv is not None else None)
    if args.scheme == "dl_scl" and not args.beta:
        raise ValueError("--beta is required for dl_scl scheme")
    return args


def run(args: argparse.Namespace) -> List[Dict[str, float]]:
    """Rows of the sweep, one per Eb/N0 point (run_ber_sweep.py:228-293)."""
    mc.maybe_init_distributed()
    seed_all(args.seed)
    N = args.N if args.N is not None else args.E
    K_total = args.K_payload + args.K_crc
    beta, retries = None, -1
    if args.scheme == "nr_ldpc":
        from ..nr.ldpc import build_h_matrix, load_base_graph
        from ..nr.ldpc._engines import ldpc_engine_for
        H = build_h_matrix(load_base_graph(args.bg), args.Z)
        if H.shape[1] - H.shape[0] != K_total:
            raise ValueError("LDPC payload+CRC size mismatch with base graph")
        eng = ldpc_engine_for(H)
        eng.configure_sweep(k_crc=args.K_crc, E=args.E, max_iter=args.max_iter, alpha=args.alpha, crc_poly=args.crc_poly)
        params_label = f"bg={args.bg},Z={args.Z},iter={args.max_iter},alpha={args.alpha}"
        return _sweep_rows(args, eng, params_label, -1, None)
    info_set = construct_info_set(N, K_total)
    if args.scheme == "polar_scl":
        # like the reference (run_ber_sweep.py:240-245): the N-bit mother code is sent; E only enters the noise variance
        # (rate = K_payload / E) and the N_or_E / rate columns, for polar_scl and dl_scl alike
        eng, params_label = engine_for(N, info_set, args.crc_poly), f"M={args.M}"
    elif args.scheme == "dl_scl":
        beta = np.load(args.beta)
        if beta.ndim != 2 or beta.shape[0] != beta.shape[1] or beta.shape[0] != K_total:
            raise ValueError("beta must be a square matrix matching abs_l0 length")   # the reference fails at its first retry
        eng, params_label, retries = engine_for(N, info_set, args.crc_poly), f"M={args.M},retries={args.retries}", args.retries
    else:
        eng, params_label = engine_for(N, info_set, args.crc_poly, E=args.E), f"M={args.M},ilv={args.ilv_mode}"
    return _sweep_rows(args, eng, params_label, retries, beta)


def _sweep_rows(args: argparse.Namespace, eng, params_label: str, retries: int, beta) -> List[Dict[str, float]]:
    """The Eb/N0 loop of run_ber_sweep.py:275-291 over an engine that offers `sweep` (polar or LDPC)."""
    rows: List[Dict[str, float]] = []
    grid = np.arange(args.EbN0_lo, args.EbN0_hi + 1e-12, args.EbN0_step)
    for point, EbN0_dB in enumerate(grid):
        st = mc.ber_point(eng, M=args.M, ebn0_db=float(EbN0_dB), payload_len=args.K_payload, coded_len=args.E,
                          seed=args.seed, stream_id=point, err_cap=args.err_cap, bits_cap=args.bits_cap,
                          retries=retries, beta=beta)
        stats = SimulationStats(bits_total=st.frames * args.K_payload, bit_errors=st.bit_errors, frame_errors=st.frame_errors,
                                work_sum=float(st.work_sum) if args.scheme in ("dl_scl", "nr_ldpc") else 0.0, frames=st.frames)
        row = stats.row()
        row.update({"scheme": args.scheme, "code": args.scheme, "N_or_E": args.E, "K_payload": args.K_payload,
                    "K_crc": args.K_crc, "rate": args.K_payload / args.E, "params": params_label, "EbN0_dB": float(EbN0_dB)})
        rows.append(row)
    return rows


def write_csv(rows: List[Dict[str, float]], path: Path) -> None:
    """13 columns, values written with str() (run_ber_sweep.py:296-317)."""
    if not rows:
        return
    body = [",".join(HEADER)] + [",".join(str(row[col]) for col in HEADER) for row in rows]
    Path(path).write_text("\n".join(body) + "\n")


def plot_rows(rows: List[Dict[str, float]], path: Path) -> None:
    if not rows:
        return
    rs = sorted(rows, key=lambda r: r["EbN0_dB"])
    snrs = [r["EbN0_dB"] for r in rs]
    semilogy_plot(Path(path), [("BER", snrs, [r["ber"] for r in rs]), ("FER", snrs, [r["fer"] for r in rs])],
                  "Eb/N0 (dB)", "Error Rate")


def main(argv: Optional[Iterable[str]] = None) -> None:
    args = parse_args(argv)
    started_here = mc.world()[1] == 1
    rows = run(args)
    is_writer = mc.world()[0] == 0
    mc.shutdown_distributed(started_here)
    if not is_writer:
        return
    out_path = Path(args.out)
    out_path.parent.mkdir(parents=True, exist_ok=True)
    write_csv(rows, out_path)
    if args.plot:
        plot_rows(rows, Path(args.plot))


if __name__ == "__main__":
    main()
