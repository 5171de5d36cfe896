"""Train the symmetric beta flip metric (reference: dl_scl_polar/train/train_beta.py:60-151), device-resident.

Same flags, log file (`logs/train_M{M}.csv`: epoch,train_loss,train_acc,val_loss,val_acc) and checkpoint
(`checkpoints/beta_M{M}.npy`, float32 [K,K]) as the reference; objective = cross entropy of -Q = -|L0| @ beta
against the oracle flip index plus lambda_l2 * mean(off_diag^2), RMSprop.  The whole dataset lives on the device
and minibatches are index slices of it (datasets from the GPU make_dataset are 100-1000x the reference's).
"""

from __future__ import annotations

import argparse
import csv
from glob import glob
from pathlib import Path
from typing import Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from ..utils.seeding import seed_all
from ..dlscl.beta import SymmetricBeta


def _load_dataset(patterns: Iterable[str]) -> Tuple[np.ndarray, np.ndarray]:
    files: List[str] = []
    for pat in patterns:
        hits = sorted(glob(pat))
        files += hits if hits else ([pat] if Path(pat).exists() else [])
    if not files:
        raise FileNotFoundError("No dataset shards found for the provided --data patterns")
    parts = [np.load(f) for f in files]
    return (np.concatenate([p["abs_l0"] for p in parts]).astype(np.float32),
            np.concatenate([p["flip_idx"] for p in parts]).astype(np.int64))


def _epoch(model, x, y, batch, lambda_l2, opt=None, gen=None):
    """One pass over (x, y); returns (mean loss, accuracy).  Trains when `opt` is given."""
    n = x.shape[0]
    if n == 0:
        return float("nan"), float("nan")
    order = torch.randperm(n, device=x.device, generator=gen) if opt is not None else torch.arange(n, device=x.device)
    loss_sum = torch.zeros((), device=x.device)
    hits = torch.zeros((), device=x.device)
    for lo in range(0, n, batch):
        idx = order[lo:lo + batch]
        with torch.set_grad_enabled(opt is not None):
            logits = -model(x[idx])
            loss = F.cross_entropy(logits, y[idx])
            if opt is not None and lambda_l2 > 0:
                loss = loss + lambda_l2 * model.off_diag.pow(2).sum() / (model.dim * model.dim)
        if opt is not None:
            opt.zero_grad(set_to_none=True)
            loss.backward()
            model.clamp_diagonal()
            opt.step()
        loss_sum += loss.detach() * idx.numel()
        hits += (logits.argmax(dim=1) == y[idx]).sum()
    return float(loss_sum / n), float(hits / n)


def train_beta(args: argparse.Namespace) -> None:
    seed_all(args.seed)
    abs_l0, labels = _load_dataset(args.data)
    device = torch.device("cuda" if (torch.cuda.is_available() and not args.cpu) else "cpu")
    perm = np.random.default_rng(args.seed).permutation(abs_l0.shape[0])
    split = int(abs_l0.shape[0] * (1.0 - args.val_frac))
    to_dev = lambda a: torch.from_numpy(a).to(device)
    xt, yt = to_dev(abs_l0[perm[:split]]), to_dev(labels[perm[:split]])
    xv, yv = to_dev(abs_l0[perm[split:]]), to_dev(labels[perm[split:]])

    model = SymmetricBeta(abs_l0.shape[1]).to(device)
    opt = torch.optim.RMSprop(model.parameters(), lr=args.lr)
    gen = torch.Generator(device=device)
    gen.manual_seed(args.seed)

    Path(args.log_dir).mkdir(parents=True, exist_ok=True)
    Path(args.checkpoint_dir).mkdir(parents=True, exist_ok=True)
    ckpt_path = Path(args.checkpoint_dir) / f"beta_M{args.M}.npy"
    best_val, best_beta = float("inf"), None
    with (Path(args.log_dir) / f"train_M{args.M}.csv").open("w", newline="") as f:
        log = csv.writer(f)
        log.writerow(["epoch", "train_loss", "train_acc", "val_loss", "val_acc"])
        for epoch in range(1, args.epochs + 1):
            model.train()
            tr_loss, tr_acc = _epoch(model, xt, yt, args.batch, args.lambda_l2, opt, gen)
            model.eval()
            va_loss, va_acc = _epoch(model, xv, yv, args.batch, 0.0)
            log.writerow([epoch, tr_loss, tr_acc, va_loss, va_acc])
            f.flush()
            if xv.shape[0] > 0 and va_loss < best_val:
                best_val, best_beta = va_loss, model.beta_matrix().detach().cpu().numpy()
    if best_beta is None:
        best_beta = model.beta_matrix().detach().cpu().numpy()
    np.save(ckpt_path, best_beta.astype(np.float32))
    print(f"Saved β checkpoint to {ckpt_path}")


_FLAGS = [("--M", dict(type=int, required=True, help="SCL list size")),
          ("--data", dict(nargs="+", required=True, help="Glob(s) to dataset shards")),
          ("--epochs", dict(type=int, default=8)), ("--lr", dict(type=float, default=1e-4)),
          ("--batch", dict(type=int, default=128)), ("--lambda_l2", dict(type=float, default=0.25)),
          ("--seed", dict(type=int, default=0)), ("--val_frac", dict(type=float, default=0.1)),
          ("--checkpoint_dir", dict(type=str, default="checkpoints")), ("--log_dir", dict(type=str, default="logs")),
          ("--cpu", dict(action="store_true", help="Force CPU even if CUDA is available"))]


def build_argparser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Train symmetric β for DL-SCL")
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)
    return parser


def main(argv: Optional[List[str]] = None) -> None:
    train_beta(build_argparser().parse_args(argv))


if __name__ == "__main__":
    main()
