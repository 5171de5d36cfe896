"""Top-level alias so that ``import dl_scl_polar...`` / ``python -m dl_scl_polar.eval.run_fer_sweep`` resolve to the
B200 mirror in ``polar_code_b200/dl_scl_polar`` (same module paths as the reference package)."""

from pathlib import Path as _Path

import polar_code_b200.dl_scl_polar as _impl

__path__ = [str(_Path(_impl.__file__).resolve().parent)]

from polar_code_b200.dl_scl_polar import config  # noqa: E402,F401

__all__ = ["config"]
