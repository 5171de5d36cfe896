"""Tiny end-to-end run of every kernel family on ragged batch sizes (launch-failure / fault smoke; also the script to put
under compute-sanitizer where that tool is available)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from polar_code_b200.engine import PolarEngine, construct_info_set
from polar_code_b200.ldpc import LdpcEngine, build_h_matrix
g = np.load("tests/golden/scl_p128.npz")
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
nv = 1.0 / (2 * 0.5 * 10 ** 0.4)
c = torch.zeros(16, dtype=torch.int64, device="cuda")
for M, beta in ((4, g["beta_M4"]), (8, None), (1, None)):
    eng.sweep(c, M=M, noise_var=nv, n_frames=3001, seed=1, stream_id=3, k_payload=40, retries=8, beta=beta)
    eng.sweep(c, M=M, noise_var=nv, n_frames=2999, seed=1, stream_id=3, k_payload=40, retries=-1, include_uncoded=True, noise_var_uncoded=0.3)
msg, llr = eng.channel(noise_var=nv, n_frames=1501, seed=2, k_payload=40)
for M in (1, 2, 4, 8):
    eng.scl_decode(llr, M, want=("cand", "metrics", "info_llrs", "n_cand", "best_idx", "best_bits", "crc_ok", "flags"))
    eng.dlscl_decode(llr, M, 4, beta=g["beta_M4"])
eng.sc_decode(llr)
e2 = PolarEngine(128, construct_info_set(128, 88), "0x1864CFB"); e2.set_rate_matching(256)
e2.sweep(c, M=4, noise_var=0.8, n_frames=2001, seed=1, k_payload=64, frame_error_mode=1, bit_error_span=64)
e3 = PolarEngine(512, construct_info_set(512, 256), "0x1864CFB")
e3.sweep(c, M=4, noise_var=0.6, n_frames=1001, seed=1, k_payload=232, retries=3)
for Z, E, kc, poly in ((2, 12, 0, None), (8, 70, 4, "0x17"), (32, 384, 24, "0x1864CFB"), (200, 1200, 0, None)):
    le = LdpcEngine(build_h_matrix(2, Z)); le.configure_sweep(k_crc=kc, E=E, max_iter=8, alpha=0.8, crc_poly=poly)
    le.sweep(c, noise_var=0.5, n_frames=777, seed=3)
    p, l = le.channel(noise_var=0.5, n_frames=333, seed=3)
    le.decode(l, max_iter=8)
    le.encode(p if kc == 0 else torch.zeros((5, le.k), dtype=torch.uint8))
torch.cuda.synchronize()
print("memcheck run finished", c.cpu().numpy()[:4])
