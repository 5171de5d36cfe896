"""Print the main numbers of a bench.py JSON line (file argument)."""
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g  e2e %.4g  ms/step %.2f  issue_frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline_issue"]["frac"]))
        print("kernel", d["kernel"], "clocks", d["clocks"])
        print("extras", {k: float("%.4g" % v) for k, v in d.get("extras", {}).items()})
        print("config5", d.get("config5_dlscl_M8"))
        print("cpu", d.get("cpu_baseline"))
        print("check", d["check"])
