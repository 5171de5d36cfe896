"""Profiling target: a few launches of one hot kernel on the headline geometry (run under ncu on the GPU box).
usage: python scripts/prof_decode.py [decode|sweep|dl] [M] [snr] [frames]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from polar_code_b200.engine import PolarEngine, construct_info_set
what = sys.argv[1] if len(sys.argv) > 1 else "decode"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4
snr = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
nv = 1.0 / (2 * 0.5 * 10 ** (snr / 10))
g = np.load("tests/golden/scl_p128.npz")
c = torch.zeros(16, dtype=torch.int64, device="cuda")
if what == "decode":
    _, llr = eng.channel(noise_var=nv, n_frames=B, seed=2026, stream_id=0, k_payload=40)
    fn = lambda: eng.scl_decode(llr, M, want=("best_bits", "crc_ok", "flags"))
elif what == "sweep":
    fn = lambda: eng.sweep(c, M=M, noise_var=nv, n_frames=B, seed=1, stream_id=2, k_payload=40)
else:
    beta = torch.as_tensor(g[f"beta_M{M}"], device="cuda")
    fn = lambda: eng.sweep(c, M=M, noise_var=nv, n_frames=B, seed=1, stream_id=3, k_payload=40, retries=8, beta=beta)
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{what} M={M} snr={snr}: {e0.elapsed_time(e1):.3f} ms  {B / e0.elapsed_time(e1) * 1e3:.4g} frames/s", flush=True)
