"""Dataset generation for beta training, batched on the B200 engine.

Same CLI, `.npz` schema (`abs_l0` float32 [n,K], `flip_idx` int32 [n], `meta` JSON) and labelling rule as the
reference (dl_scl_polar/train/make_dataset.py:24-121): all-zero payload, SCL(M) baseline; for every CRC failure
the 8 smallest-|L0| positions are tried in order with retry_with_flip and the first flip that yields the
transmitted word (and passes the CRC) becomes the label.  Instead of one frame at a time, the baseline decode runs
over chunks of frames and each of the (at most 8) flip attempts is ONE batched forced decode over the frames
that are still unlabelled.  Noise comes from Philox (statistical, not draw-for-draw, parity with the reference).
"""

from __future__ import annotations

import argparse
import json
from pathlib import Path
from typing import List, Optional

import numpy as np
import torch

from .. import config
from ..utils.seeding import seed_all
from ..polar.polar import construct_info_set
from .._engines import engine_for
from polar_code_b200 import montecarlo as mc

MAX_ATTEMPTS = 8       # make_dataset.py:69


def label_failures(eng, llr: torch.Tensor, M: int, info: torch.Tensor):
    """For a chunk of LLR rows: indices of baseline CRC failures, their |L0| rows and oracle flip labels (-1: none)."""
    K = eng.K
    base = eng.scl_decode(llr, M, want=("best_bits", "best_idx", "crc_ok", "info_llrs"))
    fail = torch.nonzero(base["crc_ok"] == 0).flatten()
    if fail.numel() == 0:
        return fail, torch.empty((0, K), dtype=torch.float32, device=eng.dev), torch.empty(0, dtype=torch.int64, device=eng.dev)
    best_bits = base["best_bits"][fail]                                            # [F,K]
    ill = base["info_llrs"][fail, base["best_idx"][fail].long()]                   # best_path_info_llrs, [F,K]
    abs_l0 = ill.abs()
    order = torch.argsort(abs_l0, dim=1)[:, :min(MAX_ATTEMPTS, K)]                 # :67-69
    label = torch.full((fail.numel(),), -1, dtype=torch.int64, device=eng.dev)
    cols = torch.arange(K, device=eng.dev).unsqueeze(0)
    for a in range(order.shape[1]):
        todo = torch.nonzero(label < 0).flatten()
        if todo.numel() == 0:
            break
        idx = order[todo, a].unsqueeze(1)                                          # flip position per frame
        bb = best_bits[todo].to(torch.int8)
        force = torch.where(cols < idx, bb, torch.full_like(bb, -1))               # _force_vector, flip.py:30-34
        force = torch.where(cols == idx, 1 - bb, force)
        out = eng.scl_decode(llr[fail[todo]], M, force=force, want=("best_bits", "crc_ok"))
        ok = (out["crc_ok"] != 0) & (out["best_bits"] == info.unsqueeze(0)).all(dim=1)   # :82
        label[todo[ok]] = idx.flatten()[ok]
    return fail, abs_l0, label


def _geometry(args: argparse.Namespace):
    """(N, K_total, crc_poly, crc_bits): the reference is fixed to config.DEFAULTS (make_dataset.py:25-29); the optional
    --N / --K_total / --crc_poly flags (absent from a reference-style Namespace) let a beta be trained for any
    geometry, e.g. the 88 x 88 one run_ber_sweep --scheme dl_scl --K_payload 64 --K_crc 24 needs."""
    cfg = config.get_config()
    N = getattr(args, "N", None) or cfg.N
    K = getattr(args, "K_total", None) or cfg.K
    poly = getattr(args, "crc_poly", None) or cfg.crc_poly
    crc_bits = int(poly, 16).bit_length() - 1
    if not 0 < crc_bits < K <= N:
        raise ValueError("need 0 < CRC degree < K_total <= N")
    return int(N), int(K), poly, crc_bits


def generate_samples(args: argparse.Namespace) -> None:
    N, K, crc_poly, crc_bits = _geometry(args)
    seed_all(args.seed)
    info_set = construct_info_set(N, K)
    eng = engine_for(N, info_set, crc_poly)
    nv = mc.fer_noise_var(args.snr_db, K, N)
    info = torch.zeros(K, dtype=torch.uint8, device=eng.dev)                       # attach_crc(0...0) = 0...0 (:31-33)

    rows: List[torch.Tensor] = []
    labels: List[torch.Tensor] = []
    failures = 0
    chunk = 1 << 18
    for begin in range(0, args.frames, chunk):
        n = min(chunk, args.frames - begin)
        msg, llr = eng.channel(noise_var=nv, n_frames=n, frame_begin=begin, seed=args.seed, stream_id=0,
                               k_payload=K - crc_bits)
        # the channel kernel draws a random payload; flipping the LLR signs by its codeword gives the all-zero
        # codeword's observation under the same noise (the BPSK/AWGN channel is symmetric)
        llr = llr * (1.0 - 2.0 * eng.encode(msg).to(torch.float32))
        fail, abs_l0, label = label_failures(eng, llr, args.M, info)
        good = label >= 0
        failures += int((~good).sum().item())
        rows.append(abs_l0[good])
        labels.append(label[good])
    abs_array = torch.cat(rows).cpu().numpy().astype(np.float32) if rows else np.zeros((0, K), np.float32)
    label_array = torch.cat(labels).cpu().numpy().astype(np.int32) if labels else np.zeros(0, np.int32)
    if label_array.size == 0:
        raise RuntimeError("No samples collected; consider increasing frames or SNR")
    meta = {"M": args.M, "EbN0_dB": args.snr_db, "seed": args.seed, "frames": args.frames, "crc_poly": crc_poly,
            "crc_bits": crc_bits, "samples": int(label_array.size), "failures": int(failures)}
    if (N, K) != (config.DEFAULTS.N, config.DEFAULTS.K):      # the reference's schema is kept for its own geometry
        meta.update({"N": N, "K_total": K})
    out_path = Path(args.out)
    out_dir = out_path.parent if out_path.parent != Path("") else Path(".")
    out_dir.mkdir(parents=True, exist_ok=True)
    shard = out_dir / f"{out_path.name}_part0.npz"
    np.savez_compressed(shard, abs_l0=abs_array, flip_idx=label_array, meta=json.dumps(meta))
    print(f"Saved {label_array.size} samples to {shard}")


def build_argparser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Generate DL-SCL flip dataset")
    parser.add_argument("--M", type=int, required=True, help="SCL list size")
    parser.add_argument("--snr_db", type=float, default=5.0, help="AWGN Eb/N0 in dB")
    parser.add_argument("--frames", type=int, default=100000, help="Number of frames to simulate")
    parser.add_argument("--seed", type=int, default=0, help="RNG seed")
    parser.add_argument("--out", type=str, required=True, help="Output prefix for dataset shards")
    parser.add_argument("--N", type=int, default=None, help="Code length (default: config.N)")
    parser.add_argument("--K_total", type=int, default=None, help="Information bits incl. CRC (default: config.K)")
    parser.add_argument("--crc_poly", type=str, default=None, help="CRC polynomial, hex with leading 1 (default: config.crc_poly)")
    return parser


def main(argv: Optional[List[str]] = None) -> None:
    generate_samples(build_argparser().parse_args(argv))


if __name__ == "__main__":
    main()
