"""Why is M=1 slow behind an M=2 launch?  (run on the GPU box)"""
import sys, ctypes
sys.path.insert(0, ".")
import torch
from polar_code_b200.engine import PolarEngine, construct_info_set
rt = ctypes.CDLL("libcudart.so.12")
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
B = 1 << 21
_, llr = eng.channel(noise_var=0.3, n_frames=B, seed=1, stream_id=0, k_payload=40)
def t(M, reps=5):
    fn = lambda: eng.scl_decode(llr, M, want=("best_bits", "crc_ok", "flags"))
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return B / (e0.elapsed_time(e1) / reps) * 1e3
seq = [int(x) for x in sys.argv[1].split(",")]
for M in seq:
    if M == 0:
        torch.cuda.synchronize(); print("reset ->", rt.cudaCtxResetPersistingL2Cache()); continue
    if M == -1:
        lim = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(lim), 6); print("persisting limit", lim.value); continue
    print(f"M={M}: {t(M):.4g}", flush=True)
