"""FER/BER sweep CLI, SCL vs DL-SCL (+ uncoded), on the B200 engine.

Same flags, printed lines, CSV name/columns/formatting and plot file as the reference
(dl_scl_polar/eval/run_fer_sweep.py:41-217).  The per-frame Python loop (:79-121) is replaced by ONE fused
launch per SNR point: Philox payload -> CRC -> encode -> BPSK/AWGN -> LLR -> SCL(M) -> DL-SCL retry rounds ->
int64 counters.  Under torchrun the frames of each point are sharded over the ranks and the counter block is
all-reduced over NCCL; rank 0 prints and writes.  Random numbers come from Philox4x32-10 keyed by
(--seed, int(snr_db*10)) instead of NumPy's PCG64, so rows agree with the reference statistically (inside its
binomial confidence interval), not draw for draw.
"""

from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

from .. import config
from ..utils.seeding import seed_all
from ..polar.polar import construct_info_set
from ..polar.scl import decode_scl
from ..dlscl.flip import decode_with_retries
from .._engines import engine_for
from ._plot import semilogy_plot
from polar_code_b200 import montecarlo as mc


def _bpsk(bits: np.ndarray) -> np.ndarray:
    return 1.0 - 2.0 * bits


def simulate_frame(llr, info_set, M, crc_poly, retries, beta) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """Per-frame pair (plain SCL, DL-SCL) as in run_fer_sweep.py:28-38."""
    return (decode_scl(llr, info_set, M, crc=crc_poly),
            decode_with_retries(llr, info_set, M, retries, crc=crc_poly, beta=beta))


def _snr_grid(args) -> np.ndarray:
    if args.snr_step > 0:
        return np.arange(args.snr_lo, args.snr_hi + 1e-9, args.snr_step)
    return np.array([args.snr_lo])


def sweep_rows(args: argparse.Namespace) -> List[Dict[str, float]]:
    """Counters -> the reference's row dictionaries (run_fer_sweep.py:123-148)."""
    cfg = config.get_config()
    seed_all(args.seed)
    info_set = construct_info_set(cfg.N, cfg.K)
    payload_bits = cfg.K - cfg.crc_bits
    beta = np.load(args.beta) if args.beta else None
    if beta is not None and beta.shape != (cfg.K, cfg.K):
        raise ValueError("beta must be a square matrix matching abs_l0 length")
    t_start = time.perf_counter()
    eng = engine_for(cfg.N, info_set, cfg.crc_poly)
    rank, _ = mc.world()
    t_engine = time.perf_counter() - t_start
    t_warm = mc.warm_up(eng, M=args.M, retries=args.retries, beta=beta, k_payload=payload_bits)
    if rank == 0:                                 # start-up is reported apart from the per-point steady state
        print(f"[b200] set-up: engine {t_engine:.3f} s + first launch {t_warm:.3f} s (kernel attributes, scratch, queues); "
              f"the points below are steady state", file=sys.stderr)
    rows: List[Dict[str, float]] = []
    for snr_db in _snr_grid(args):
        t0 = time.perf_counter()
        c = mc.fer_point(eng, M=args.M, snr_db=float(snr_db), frames=args.frames, seed=args.seed, retries=args.retries,
                         beta=beta, include_uncoded=args.include_uncoded, k_payload=payload_bits)
        dt = time.perf_counter() - t0            # includes the counter all-reduce and the device->host read
        frames = int(c[0])
        if rank == 0:                             # timing goes to stderr: stdout keeps the reference's lines only
            print(f"[b200] {snr_db:.2f} dB: {frames} frames in {dt:.3f} s = {frames / max(dt, 1e-9):.3e} frames/s "
                  f"over {mc.world()[1]} GPU(s)", file=sys.stderr)
        bits_coded = frames * cfg.K
        nan = float("nan")
        row = {"snr_db": snr_db,
               "fer_scl": c[1] / frames if frames else nan, "fer_dl": c[3] / frames if frames else nan,
               "ber_scl": c[2] / bits_coded if bits_coded else nan, "ber_dl": c[4] / bits_coded if bits_coded else nan,
               "frames": frames, "near_tie_frames": int(c[8]), "avg_retries": c[7] / frames if frames else nan}
        if args.include_uncoded:
            bits_unc = frames * payload_bits
            row["fer_uncoded"] = c[5] / frames if frames else nan
            row["ber_uncoded"] = c[6] / bits_unc if bits_unc else nan
        if rank == 0:
            head = f"SNR={snr_db:.2f} dB -> "
            if args.include_uncoded:
                head += f"Uncoded FER={row['fer_uncoded']:.3e}, BER={row['ber_uncoded']:.3e}; "
            print(head + f"SCL FER={row['fer_scl']:.3e}, BER={row['ber_scl']:.3e}; "
                         f"DL FER={row['fer_dl']:.3e}, BER={row['ber_dl']:.3e}")
        rows.append(row)
    return rows


def write_outputs(args: argparse.Namespace, rows: List[Dict[str, float]]) -> None:
    """results/fer_M{M}.csv (columns and number formats of run_fer_sweep.py:150-173) and plots/fer_M{M}.png."""
    out_dir = Path(args.out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    csv_path = out_dir / f"fer_M{args.M}.csv"
    cols = ["fer_uncoded", "ber_uncoded"] if args.include_uncoded else []
    cols += ["fer_scl", "ber_scl", "fer_dl", "ber_dl"]
    lines = [",".join(["snr_db"] + cols)]
    for row in rows:
        lines.append(",".join([f"{row['snr_db']:.3f}"] + [f"{row[c]:.6e}" for c in cols]))
    csv_path.write_text("\n".join(lines) + "\n")
    print(f"Saved FER table to {csv_path}")
    plot_path = Path(args.plot_dir) / f"fer_M{args.M}.png"
    snrs = [row["snr_db"] for row in rows]
    series = [("Uncoded", snrs, [r["fer_uncoded"] for r in rows])] if args.include_uncoded else []
    series += [("SCL", snrs, [r["fer_scl"] for r in rows]), ("DL-SCL", snrs, [r["fer_dl"] for r in rows])]
    semilogy_plot(plot_path, series, "Eb/N0 (dB)", "Frame Error Rate")
    print(f"Saved FER plot to {plot_path}")


def run_sweep(args: argparse.Namespace) -> None:
    started_here = mc.world()[1] == 1
    mc.maybe_init_distributed()
    rows = sweep_rows(args)
    if mc.world()[0] == 0:
        write_outputs(args, rows)
    mc.shutdown_distributed(started_here)


_FLAGS = [
    ("--M", dict(type=int, required=True, help="List size")),
    ("--frames", dict(type=int, default=10000, help="Frames per SNR point")),
    ("--snr_lo", dict(type=float, default=4.0)),
    ("--snr_hi", dict(type=float, default=6.5)),
    ("--snr_step", dict(type=float, default=0.5)),
    ("--retries", dict(type=int, default=8)),
    ("--beta", dict(type=str, help="Path to trained beta matrix (.npy)")),
    ("--seed", dict(type=int, default=0)),
    ("--out_dir", dict(type=str, default="results")),
    ("--plot_dir", dict(type=str, default="plots")),
    ("--include_uncoded", dict(action="store_true", help="Also simulate an uncoded BPSK baseline")),
]


def build_argparser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Run FER sweep for DL-SCL")
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)
    return parser


def main(argv: Optional[List[str]] = None) -> None:
    run_sweep(build_argparser().parse_args(argv))


if __name__ == "__main__":
    main()
