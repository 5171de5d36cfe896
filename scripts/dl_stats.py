"""Binned DL-SCL retry kernel: throughput and scheduler statistics (run on the GPU box).
usage: python scripts/dl_stats.py [M] [snr] [frames]     env: PB200_DL_TARGET=<multiple of the resident frames kept in flight>"""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from polar_code_b200.engine import PolarEngine, construct_info_set
args = [x for x in sys.argv[1:] if not x.startswith("--")]
M = int(args[0]) if len(args) > 0 else 4
snr = float(args[1]) if len(args) > 1 else 4.0
B = int(args[2]) if len(args) > 2 else 1 << 20
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
nv = 1.0 / (2 * 0.5 * 10 ** (snr / 10))
g = np.load("tests/golden/scl_p128.npz")
c = torch.zeros(16, dtype=torch.int64, device="cuda")
beta = torch.as_tensor(g[f"beta_M{M}"], device="cuda")
fn = lambda: eng.sweep(c, M=M, noise_var=nv, n_frames=B, seed=1, stream_id=3, k_payload=40, retries=8, beta=beta)
ts = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
st = (C.c_uint * 8)()
eng.lib.pb200_debug_bin_stats(eng._h, st)
s = list(st)
print(f"M={M} snr={snr} B={B}: best {min(ts[2:]):.3f} ms median {np.median(ts[2:]):.3f} ms -> {B / np.median(ts[2:]) * 1e3:.4g} frames/s | waits {s[0]} lost {s[1]} batches {s[2]} decodes {s[3]} "
      f"(fill {s[3] / max(s[2], 1):.2f}/{32 // max(M, 1) if M in (1, 2, 4, 8) else '?'}) mean start phase {s[4] / max(s[2], 1):.1f} mixed {s[5]} max-sched {s[6] * 1.024e-3:.2f} ms cta-life {s[7] * 1.024e-3:.2f} ms | all {[round(t, 1) for t in ts]}")
if "--trace" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3): fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
    import collections
    agg = collections.defaultdict(list)
    for e in evs: agg[e.name[:60]].append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])): print(f"  {k:60s} n={len(v):3d} mean {np.mean(v):9.1f} us")
