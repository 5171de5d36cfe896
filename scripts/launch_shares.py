"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: python scripts/launch_shares.py launches.csv"""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
t, n = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) != len(h) or r[0] == "ID": continue
    name = re.sub(r"\(.*", "", r[ik])[:100]
    ms = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    t[name] += ms; n[name] += 1
tot = sum(t.values())
print("launch list of: python bench.py --steps 2 --warmup 3 --frames 1048576 --e2e-frames 262144 --cpu-sample 60000 "
      "(ncu --metrics gpu__time_duration.sum --clock-control none -c 400)")
print(f"total device time of the listed launches: {tot:.1f} ms (cold-cache, serialised: compare shares)")
for k, v in t.most_common():
    print(f"{100 * v / tot:5.1f}%  {v:8.3f} ms  {n[k]:4d}x {k}")
