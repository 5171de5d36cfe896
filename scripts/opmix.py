"""Opcode mix + IMAD/MOV attribution from an ncu source-page csv."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
agg = collections.Counter(); byline = collections.Counter(); hdr=None; cur=None; curfile=None
want = sys.argv[2] if len(sys.argv) > 2 else None
for r in rows:
    if r and r[0]=="File Path": curfile=r[1].split('/')[-1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)==len(hdr):
        if r[0].isdigit() and r[2]=="-": cur=(curfile,int(r[0]),r[1].strip()[:80]); continue
        if r[0]=="" and r[2] not in ("","...","-"):
            sass=r[3].strip()
            try: n=int(r[hdr.index("Instructions Executed")] or 0)
            except: continue
            m=re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
            if m:
                op=m.group(2)
                key = op if op.startswith("IMAD.MOV") else op.split('.')[0]
                agg[key]+=n
                if want and op.startswith(want): byline[cur]+=n
tot=sum(agg.values()); print("total", tot)
print("  ".join(f"{k}:{100*v/tot:.1f}%" for k,v in agg.most_common(28)))
if want:
    t=sum(byline.values())
    for k,v in byline.most_common(14): print(f"{100*v/t:5.1f}% {k[0]}:{k[1]} {k[2]}")
