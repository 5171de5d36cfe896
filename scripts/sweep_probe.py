"""Run the fused sweep / DL-SCL rounds a few times (profiling target)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from polar_code_b200.engine import PolarEngine, construct_info_set
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
snr = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
retries = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
nv = 1.0 / (2 * 0.5 * 10 ** (snr / 10))
c = torch.zeros(16, dtype=torch.int64, device="cuda")
for it in range(3):
    c.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.sweep(c, M=4, noise_var=nv, n_frames=n, seed=1, stream_id=3, k_payload=40, retries=retries)
    e1.record(); torch.cuda.synchronize()
    print("sweep snr %.1f retries %d: %.3f ms  %.4g frames/s  counters %s" % (snr, retries, e0.elapsed_time(e1), n / e0.elapsed_time(e1) * 1e3, c.cpu().numpy()[:9]))
