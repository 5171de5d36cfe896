"""Throughput probe of the fused LDPC sweep kernel (scheme nr_ldpc) on one GPU: frames/s per (Z, E, Eb/N0)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from polar_code_b200.ldpc import LdpcEngine, build_h_matrix
from polar_code_b200.montecarlo import ber_noise_var

for Z, E, kcrc, poly, snr in [(2, 12, 0, None, 5.0), (8, 48, 0, None, 3.0), (32, 384, 24, "0x1864CFB", 1.0), (32, 192, 24, "0x1864CFB", 3.0)]:
    eng = LdpcEngine(build_h_matrix(2, Z))
    eng.configure_sweep(k_crc=kcrc, E=E, max_iter=20, alpha=0.8, crc_poly=poly)
    kp = eng.k - kcrc
    nv = ber_noise_var(snr, kp, E)
    B = 1 << 22 if Z <= 8 else 1 << 20
    c = torch.zeros(16, dtype=torch.int64, device="cuda")
    eng.sweep(c, noise_var=nv, n_frames=B, seed=1)
    torch.cuda.synchronize()
    c.zero_()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(3):
        eng.sweep(c, noise_var=nv, n_frames=B, seed=1, frame_begin=i * B)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    cc = c.cpu().numpy()
    print(f"Z={Z} n={eng.n} E={E} EbN0={snr}: {B / ms * 1e3:.3e} frames/s  fer={cc[1] / cc[0]:.4f} avg_iters={cc[7] / cc[0]:.2f}")
