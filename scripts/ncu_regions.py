"""Instruction count / stall samples / smem wavefronts of an ncu source-page csv by code REGION of the list decoder.
usage: ncu -i x.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python scripts/ncu_regions.py src.csv <groups>
Regions are found from marker lines of the CURRENT sources (the capture must be of the current build)."""
import csv, sys, collections
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from ncu_regions_lib import region
rows = list(csv.reader(open(sys.argv[1])))
groups = float(sys.argv[2]) if len(sys.argv) > 2 else 131072.0
agg = collections.defaultdict(lambda: [0, 0, 0]); cur = hdr = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]; hdr = None
    elif r and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        g = lambda n: int(r[hdr.index(n)] or 0)
        a = agg[region(cur, int(r[0]))]
        a[0] += g("Instructions Executed"); a[1] += g("Warp Stall Sampling (All Samples)"); a[2] += g("L1 Wavefronts Shared")
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp-inst/group {ti / groups:.0f}   stall samples {ts}")
print(f"{'inst/group':>10} {'inst%':>6} {'samp%':>6} {'smem wf/group':>13}  region")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{a[0] / groups:10.0f} {100 * a[0] / ti:6.1f} {100 * a[1] / ts:6.1f} {a[2] / groups:13.0f}  {k}")
