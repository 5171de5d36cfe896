"""NR polar helpers of the mirror package, served by the B200 engine.

Public surface = the reference's (dl_scl_polar/nr/polar/__init__.py:3-14): two (de)interleaver index maps, two
rate-matching helpers and the fused rate-matched encode / SCL decode."""

from . import interleaver as _ilv, rate_match as _rm, scl_nr as _nr

_EXPORTS = {
    _ilv: ("subblock_interleave", "subblock_deinterleave"),
    _rm: ("rate_match_polar", "derate_match_polar"),
    _nr: ("encode_rate_matched", "decode_rate_matched_scl"),
}
__all__ = []
for _mod, _names in _EXPORTS.items():
    for _name in _names:
        globals()[_name] = getattr(_mod, _name)
        __all__.append(_name)
del _mod, _names, _name
