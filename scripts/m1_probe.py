"""M=1 / SC throughput after a list-kernel launch took the L2 set-aside (run on the GPU box)."""
import sys
sys.path.insert(0, ".")
import torch
from polar_code_b200.engine import PolarEngine, construct_info_set
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
B = 1 << 21
_, llr = eng.channel(noise_var=0.3, n_frames=B, seed=1, stream_id=0, k_payload=40)
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return B / (e0.elapsed_time(e1) / reps) * 1e3
print("fresh process: M1 %.4g  SC %.4g" % (t(lambda: eng.scl_decode(llr, 1, want=("best_bits", "crc_ok", "flags"))), t(lambda: eng.sc_decode(llr))))
print("M4 %.4g" % t(lambda: eng.scl_decode(llr, 4, want=("best_bits", "crc_ok", "flags"))))
print("after M4 (set-aside taken): M1 %.4g  SC %.4g" % (t(lambda: eng.scl_decode(llr, 1, want=("best_bits", "crc_ok", "flags"))), t(lambda: eng.sc_decode(llr))))
print("M4 again %.4g" % t(lambda: eng.scl_decode(llr, 4, want=("best_bits", "crc_ok", "flags"))))
