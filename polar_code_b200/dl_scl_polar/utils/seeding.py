"""Seeding helper with the reference's import side effects (dl_scl_polar/utils/seeding.py:8,18,21-31)."""

from __future__ import annotations

import os

os.environ.setdefault("OMP_NUM_THREADS", "1")      # seeding.py:8

import random  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_num_threads(1)                            # seeding.py:18


def seed_all(seed: int, deterministic_torch: bool = False) -> None:
    """Seed Python, NumPy (legacy global state) and torch, CUDA included."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    for fn in (random.seed, np.random.seed, torch.manual_seed):
        fn(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic_torch:
        torch.use_deterministic_algorithms(True)


__all__ = ["seed_all"]
