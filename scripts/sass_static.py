"""Static instruction counts of one kernel by source region / line (no GPU needed).
usage: python scripts/sass_static.py "<MP>,<LOGMAX>,<FORCED>,<METRIC>" [--lines lo hi] [--kernel decode|sweep|sweeptrace|retry]
Compiles a single-kernel translation unit with -lineinfo, disassembles it with nvdisasm -g and aggregates the SASS
instructions per source line.  The list kernels hold TWO copies of the phase body (even / odd phase of a pair), so a
line of the info-phase path that shows n instructions costs about n/2 per information phase."""
import collections, re, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "scripts"))
args = sys.argv[1:]
tp = args[0] if args and not args[0].startswith("--") else "4,7,false,true"
kind = args[args.index("--kernel") + 1] if "--kernel" in args else "decode"
inst = {"decode": f"template __global__ void pb::decode_kernel<{tp}>(const Code, const Tables, const DecodeArgs);",
        "sweep": f"template __global__ void pb::sweep_kernel<{tp}>(const Code, const Tables, const SweepArgs);",
        "sweeptrace": f"template __global__ void pb::sweep_kernel<{tp}>(const Code, const Tables, const SweepArgs);",
        "retry": f"template __global__ void pb::dl_retry_kernel<{tp}>(const Code, const Tables, const SweepArgs);"}[kind]
d = Path(tempfile.mkdtemp())
(d / "one.cu").write_text(f'#include "{ROOT}/polar_code_b200/csrc/polar_sweep.cuh"\nusing namespace pb;\n{inst}\n')
r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-diag-suppress", "177",
                    "-Xptxas", "-v", *[a for a in args if a.startswith("-D")], "-cubin", "-o", str(d / "one.cubin"), str(d / "one.cu")], capture_output=True, text=True)
if r.returncode: raise SystemExit(r.stderr)
print([l for l in r.stderr.splitlines() if "registers" in l or "spill" in l])
sass = subprocess.run(["nvdisasm", "-g", "-c", str(d / "one.cubin")], capture_output=True, text=True).stdout
cur = None; per = collections.Counter(); ops = collections.defaultdict(list); total = 0
want_fn = {"decode": "decode_kernel", "sweep": "sweep_kernel", "sweeptrace": "sweep_kernel", "retry": "dl_retry_kernel"}[kind]
in_fn = False
for l in sass.splitlines():
    m = re.match(r"//-+ \.text\.(\S+)", l)
    if m: in_fn = want_fn in m.group(1); cur = None; continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m and cur:
        per[cur] += 1; total += 1; ops[cur].append(m.group(1).strip())
print("static SASS instructions:", total)
if "--lines" in args:
    i = args.index("--lines"); fn, lo, hi = args[i + 1], int(args[i + 2]), int(args[i + 3])
    src = (ROOT / "polar_code_b200" / "csrc" / fn).read_text().splitlines()
    for ln in range(lo, hi + 1):
        if per[(fn, ln)]:
            print(f"--- {fn}:{ln} [{per[(fn, ln)]}]  {src[ln - 1].strip()[:110]}")
            if "--ops" in args:
                for o in ops[(fn, ln)]: print("        ", o)
else:
    import importlib.util
    from ncu_regions_lib import region
    agg = collections.Counter()
    for (f, ln), n in per.items(): agg[region(f, ln)] += n
    for k, n in agg.most_common(): print(f"{n:6d}  {k}")
