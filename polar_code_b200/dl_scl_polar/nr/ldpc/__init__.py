"""NR LDPC helpers of the mirror package, served by the B200 engine (reference surface:
dl_scl_polar/nr/ldpc/__init__.py:3-16)."""

from .basegraphs import BaseGraph, load_base_graph
from .builder import build_h_matrix
from .encode import encode_ldpc
from .rate_match import rate_match_ldpc, derate_match_ldpc
from .decode_nms import decode_ldpc_nms

__all__ = ["load_base_graph", "build_h_matrix", "encode_ldpc", "rate_match_ldpc", "derate_match_ldpc", "decode_ldpc_nms"]
