"""DL-SCL sweep legs in bench.py order (beta / no beta, M = 4 / 8): per-call times, to spot host-side stalls between launches."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from polar_code_b200.engine import PolarEngine, construct_info_set
eng = PolarEngine(128, construct_info_set(128, 64), "0x1864CFB")
g = np.load("tests/golden/scl_p128.npz")
nv = lambda s: 1.0 / (2 * 0.5 * 10 ** (s / 10))
B = 1 << 21
c = torch.zeros(16, dtype=torch.int64, device="cuda")
def run(M, snr, beta, tag):
    b = None if beta is None else torch.as_tensor(beta, device="cuda")
    ts = []
    for _ in range(6):
        c.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.sweep(c, M=M, noise_var=nv(snr), n_frames=B, seed=1, stream_id=3, k_payload=40, retries=8, beta=b); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 2))
    cc = c.cpu().numpy()
    print(tag, M, snr, "beta" if beta is not None else "nobeta", ts, "retries/frame %.3f queued %.4f" % (cc[7] / cc[0], cc[1] / cc[0]), flush=True)
run(4, 5.0, None, "first")
run(4, 5.0, g["beta_M4"], "")
run(4, 4.0, None, "")
run(4, 4.0, g["beta_M4"], "")
run(8, 4.0, g["beta_M8"], "")
run(8, 5.0, g["beta_M8"], "")
run(4, 5.0, None, "after M8")
run(4, 6.0, None, "")
