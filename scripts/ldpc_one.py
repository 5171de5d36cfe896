import sys
sys.path.insert(0, ".")
import torch
from polar_code_b200.ldpc import LdpcEngine, build_h_matrix
from polar_code_b200.montecarlo import ber_noise_var
eng = LdpcEngine(build_h_matrix(2, 32))
eng.configure_sweep(k_crc=24, E=384, max_iter=20, alpha=0.8, crc_poly="0x1864CFB")
c = torch.zeros(16, dtype=torch.int64, device="cuda")
for i in range(3):
    eng.sweep(c, noise_var=ber_noise_var(1.0, 72, 384), n_frames=1 << 20, seed=1, frame_begin=i << 20)
torch.cuda.synchronize()
print(c.cpu().numpy()[:8])
