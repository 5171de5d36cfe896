"""`seed_all` for the mirror package.

Importing this module has the same process-wide side effects as the reference's
(dl_scl_polar/utils/seeding.py:8,18): OpenMP defaults to one thread and torch's intra-op pool is pinned to one
thread.  They are kept because callers of the reference may rely on them; the GPU engine itself does not care."""

from __future__ import annotations

import os
import random

os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_num_threads(1)


def seed_all(seed: int, deterministic_torch: bool = False) -> None:
    """One seed for every generator a sweep can touch: hash seed, `random`, NumPy's legacy global state, torch
    (CPU and all CUDA devices).  The Philox streams of the GPU sweeps take the same seed through `--seed`."""
    os.environ["PYTHONHASHSEED"] = f"{seed}"
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.use_deterministic_algorithms(True) if deterministic_torch else None


__all__ = ["seed_all"]
