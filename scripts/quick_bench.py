"""Short bench summary used while iterating on kernels (run on the GPU box)."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--cpu-sample", "60000"] + sys.argv[1:], capture_output=True, text=True)
for l in out.stdout.splitlines():
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g  e2e %.4g  ms/step %.2f  issue_frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline_issue"]["frac"]))
        print("kernel", d["kernel"], "clocks", d["clocks"])
        print("extras", {k: float("%.4g" % v) for k, v in d.get("extras", {}).items()})
        print("check", d["check"])
if out.returncode: print(out.stdout[-2000:], out.stderr[-3000:])
