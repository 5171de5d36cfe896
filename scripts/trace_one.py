import sys
sys.path.insert(0, ".")
import torch
from polar_code_b200.engine import PolarEngine, construct_info_set
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
c = torch.zeros(16, dtype=torch.int64, device="cuda")
nv = 1.0 / (2 * 0.5 * 10 ** 0.7)
for i in range(3):
    eng.sweep(c, M=4, noise_var=nv, n_frames=1 << 20, seed=1, stream_id=3, k_payload=40, retries=8, frame_begin=i << 20)
torch.cuda.synchronize()
print(c.cpu().numpy()[:9])
