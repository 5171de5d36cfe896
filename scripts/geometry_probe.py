"""Throughput of non-headline geometries on one GPU: NR rate-matched sweep (config 4), N = 256 / 512 list decoding."""
import sys
sys.path.insert(0, ".")
import torch
from polar_code_b200.engine import PolarEngine, construct_info_set
from polar_code_b200.montecarlo import ber_noise_var

def timed(fn, n_frames, reps=3):
    fn(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize()
    return n_frames * reps / (t0.elapsed_time(t1) * 1e-3)

CRC = "0x1864CFB"
# config 4: A(128, 88), E = 256, M = 4
eng = PolarEngine(128, construct_info_set(128, 88), CRC)
eng.set_rate_matching(256)
c = torch.zeros(16, dtype=torch.int64, device="cuda")
B = 1 << 20
for snr in (1.0, 4.0):
    nv = ber_noise_var(snr, 64, 256)
    r = timed(lambda: eng.sweep(c, M=4, noise_var=nv, n_frames=B, seed=1, stream_id=1, k_payload=64, frame_error_mode=1, bit_error_span=64), B)
    print(f"NR sweep N=128 K=88 E=256 M=4 @{snr} dB: {r:.3e} frames/s")
eng.set_rate_matching(0)
r = timed(lambda: eng.sweep(c, M=4, noise_var=0.5, n_frames=B, seed=1, stream_id=1, k_payload=64), B)
print(f"plain sweep N=128 K=88 M=4: {r:.3e} frames/s")
for N, K in ((256, 128), (512, 256), (64, 32), (32, 16)):
    e = PolarEngine(N, construct_info_set(N, K), CRC if K > 24 else None)
    Bn = (1 << 27) // N // 4
    _, llr = e.channel(noise_var=0.6, n_frames=Bn, seed=2, k_payload=K - 24 if K > 24 else K)
    for M in (1, 4, 8):
        r = timed(lambda: e.scl_decode(llr, M, want=("best_bits", "crc_ok", "flags")), Bn)
        print(f"decode N={N} K={K} M={M}: {r:.3e} frames/s ({r * N / 1e9:.1f} G coded bits/s)")
