"""Default geometry of the mirror package (reference values: dl_scl_polar/config.py:9-27).

`PolarConfig` is a plain dataclass with the reference's field names; `DEFAULTS` is the module-level instance that
`polar.encode` reads, `get_config()` hands out a copy."""

from __future__ import annotations

import copy
import dataclasses

#  field        type         default / factory                meaning
_SPEC = (
    ("N",          int,   128,                              "code length (power of two)"),
    ("K",          int,   64,                               "information bits, CRC included"),
    ("crc_poly",   str,   "0x1864CFB",                      "CRC polynomial, hex with the leading 1 (24-bit CRC)"),
    ("crc_bits",   int,   24,                               "CRC length"),
    ("list_sizes", list,  lambda: [1, 2, 4, 8],             "SCL list sizes of the experiments"),
    ("retries",    int,   8,                                "DL-SCL flip retries"),
    ("ebno_sweep", list,  lambda: [4.0, 6.5, 0.5],          "Eb/N0 grid: lo, hi, step (dB)"),
    ("seed",       int,   0,                                "base RNG seed"),
)

PolarConfig = dataclasses.make_dataclass(
    "PolarConfig",
    [(name, typ, dataclasses.field(default_factory=dflt) if callable(dflt) else dataclasses.field(default=dflt))
     for name, typ, dflt, _ in _SPEC],
)
PolarConfig.__doc__ = "; ".join(f"{name}: {doc}" for name, _, _, doc in _SPEC)
PolarConfig.__module__ = __name__

DEFAULTS = PolarConfig()


def get_config():
    """Copy of DEFAULTS; assigning to its fields leaves the module-level defaults untouched."""
    return copy.copy(DEFAULTS)
