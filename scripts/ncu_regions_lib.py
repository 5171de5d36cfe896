"""Source-line -> code region of the list decoder (shared by ncu_regions.py and sass_static.py)."""
import csv, sys, re, collections
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1] / "polar_code_b200" / "csrc"
def find(fn, pat):
    for i, l in enumerate((ROOT / fn).read_text().splitlines(), 1):
        if pat in l: return i
    raise SystemExit(f"marker not found: {pat}")
D, C = "polar_decode.cuh", "polar_core.cuh"
bounds = {D: [(1, "misc"), (find(D, "void pair_llr("), "pair_llr dispatch"), (find(D, "void set_bit_odd("), "set_bit dispatch"),
              (find(D, "auto phase = "), "phase prologue + L"), (find(D, "const float tail = softplus_tail"), "metric (m0/m1, doubles)"),
              (find(D, "if constexpr (MP == 1) {"), "MP1 pick"), (find(D, "if (!is_info) {"), "frozen"),
              (find(D, "// Keys: IEEE"), "keys + rank"), (find(D, "// scl.py:173-174: the sorted"), "publish + tie + src"),
              (find(D, "// metric of the candidate"), "state move (metric, shuffles)"), (find(D, "if (!odd) p.bw[0]"), "bit store + sync"),
              (find(D, "if constexpr (!FORCED && !TRACE && MP == 1)"), "phase loop"), (find(D, "// final list order"), "final rank"),
              (find(D, "void trace_walk("), "trace/replay")],
          C: [(1, "misc"), (find(C, "float f_op("), "f/g ops"), (find(C, "float softplus_tail("), "softplus"), (find(C, "struct WarpMem"), "misc"),
              (find(C, "void stage_channel_rows("), "channel staging"), (find(C, "BitsCfg"), "bits: ascend/store"), (find(C, "void transform_words("), "transform"),
              (find(C, "struct Tree {"), "tree load/store/loop")]}
def region(f, l):
    if f not in bounds: return f
    name = "misc"
    for lo, n in bounds[f]:
        if l >= lo: name = n
    return (f.split("_")[1].split(".")[0] + ": " + name)
