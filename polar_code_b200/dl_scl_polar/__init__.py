"""Drop-in mirror of the reference package ``dl_scl_polar`` (heimrih/polar_code) on the B200 engine.

Same module paths, function names, argument meaning, return structures and exceptions as the reference for
the decode hot path; every call forwards to the CUDA kernels behind ``include/polar_b200.h``.  There is no
CPU implementation here: without ``libpolar_b200.so`` and a CUDA device these functions raise.

    dl_scl_polar.polar.polar   construct_info_set, encode, sc_decode          (polar/polar.py)
    dl_scl_polar.polar.crc     attach_crc, check_crc                          (polar/crc.py)
    dl_scl_polar.polar.scl     decode_scl                                     (polar/scl.py)
    dl_scl_polar.dlscl.flip    choose_flip_index, retry_with_flip, decode_with_retries  (dlscl/flip.py)
    dl_scl_polar.nr.polar      the six NR helpers                             (nr/polar/__init__.py)
    dl_scl_polar.eval.run_fer_sweep / run_ber_sweep   CLIs, same flags and CSV columns
"""

from . import config

__all__ = ["config"]
