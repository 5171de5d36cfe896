"""Steady-state fused sweep / DL-SCL throughput for (M, snr, retries) combinations (first call excluded)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from polar_code_b200.engine import PolarEngine, construct_info_set
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
g = np.load("tests/golden/scl_p128.npz")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
for M, snr, retries, beta in [(4, 7.0, -1, None), (4, 7.0, 8, None), (8, 7.0, -1, None), (8, 7.0, 8, None), (8, 5.0, -1, None), (8, 5.0, 8, None), (8, 5.0, 8, g["beta_M8"]), (4, 5.0, 8, g["beta_M4"]), (4, 5.0, 8, None), (8, 4.0, 8, g["beta_M8"])]:
    nv = 1.0 / (2 * 0.5 * 10 ** (snr / 10))
    c = torch.zeros(16, dtype=torch.int64, device="cuda")
    b = None if beta is None else torch.as_tensor(beta, device="cuda")
    eng.sweep(c, M=M, noise_var=nv, n_frames=n, seed=1, stream_id=3, k_payload=40, retries=retries, beta=b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(3):
        eng.sweep(c, M=M, noise_var=nv, n_frames=n, seed=1, stream_id=3, k_payload=40, retries=retries, beta=b, frame_begin=(it + 1) * n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    cc = c.cpu().numpy()
    print(f"M={M} snr={snr} retries={retries} beta={'yes' if beta is not None else 'no'}: {n / ms * 1e3:.3e} frames/s  scl_fer={cc[1] / cc[0]:.4f} dl_fer={cc[3] / cc[0]:.4f} retries/frame={cc[7] / cc[0]:.3f}")
