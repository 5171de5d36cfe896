"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line.
usage: ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python scripts/ncu_lines.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
agg, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]; hdr = None
    elif r and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        g = lambda name: int(r[hdr.index(name)] or 0)
        agg.append(dict(file=cur, line=int(r[0]), src=r[1].strip()[:100], inst=g("Instructions Executed"), samp=g("Warp Stall Sampling (All Samples)"),
                        wait=g("stall_wait"), ssb=g("stall_short_sb"), lsb=g("stall_long_sb"), br=g("stall_branch_resolving"),
                        noi=g("stall_no_inst"), math=g("stall_math"), wf=g("L1 Wavefronts Shared"), wfi=g("L1 Wavefronts Shared Ideal")))
ti = sum(a["inst"] for a in agg); ts = sum(a["samp"] for a in agg)
print(f"total warp-inst {ti}  samples {ts}  smem wavefronts {sum(a['wf'] for a in agg)} (ideal {sum(a['wfi'] for a in agg)})")
print(" inst%  samp%   wait   ssb   lsb    br  noin  math    wf/ideal  where")
for a in sorted(agg, key=lambda a: -a["samp"])[:top]:
    print(f"{100*a['inst']/ti:5.1f} {100*a['samp']/ts:6.1f}  {a['wait']:5d} {a['ssb']:5d} {a['lsb']:5d} {a['br']:5d} {a['noi']:5d} {a['math']:5d}  {a['wf']:>9d}/{a['wfi']:<9d} {a['file']}:{a['line']}  {a['src']}")
