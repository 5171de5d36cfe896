"""A/B probe for kernel variants (run on the GPU box): PB200_LIBRARY=<variant.so> python scripts/ab_probe.py [tag] [--parity N]
Prints steady-state frames/s of the hot kernels on the headline geometry and, with --parity, mismatch counts vs the oracle."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from polar_code_b200.engine import PolarEngine, construct_info_set

tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "variant"
parity = int(sys.argv[sys.argv.index("--parity") + 1]) if "--parity" in sys.argv else 0
CRC = "0x1864CFB"
A = construct_info_set(128, 64)
eng = PolarEngine(128, A, CRC)
g = np.load("tests/golden/scl_p128.npz")
nv = lambda snr: 1.0 / (2 * 0.5 * 10 ** (snr / 10))
B = 1 << 21


def timeit(fn, reps=5):
    """median of `reps` individually timed calls after three warm-up calls"""
    fn(); fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


parts = [eng.channel(noise_var=nv(s), n_frames=B // 6 + 1, frame_begin=i * B, seed=2026, stream_id=i, k_payload=40)[1] for i, s in enumerate([4.0, 4.5, 5.0, 5.5, 6.0, 6.5])]
llr = torch.cat(parts)[:B].contiguous()
res = {}
for M in (4, 8, 2, 1):
    ms = timeit(lambda: eng.scl_decode(llr, M, want=("best_bits", "crc_ok", "flags")))
    res[f"scl_M{M}"] = B / ms * 1e3
c = torch.zeros(16, dtype=torch.int64, device="cuda")
ms = timeit(lambda: eng.sweep(c, M=4, noise_var=nv(5.0), n_frames=B, seed=1, stream_id=2, k_payload=40), 3)
res["sweep_M4_5dB"] = B / ms * 1e3
for M, snr in ((4, 4.0), (4, 5.0), (8, 5.0), (8, 4.0)):
    beta = torch.as_tensor(g[f"beta_M{M}"], device="cuda")
    ms = timeit(lambda: eng.sweep(c, M=M, noise_var=nv(snr), n_frames=B, seed=1, stream_id=3, k_payload=40, retries=8, beta=beta), 7)
    res[f"dl_M{M}_r8_beta_{snr}dB"] = B / ms * 1e3
ms = timeit(lambda: eng.sweep(c, M=4, noise_var=nv(4.0), n_frames=B, seed=1, stream_id=3, k_payload=40, retries=8), 7)
res["dl_M4_r8_nobeta_4.0dB"] = B / ms * 1e3
print(tag, " ".join(f"{k}={v:.4g}" for k, v in res.items()), flush=True)

if parity:
    from oracle import oracle as O
    for M, snr in [(2, 3.0), (4, 2.0), (4, 4.0), (8, 3.0), (8, 5.0)]:
        rng = np.random.default_rng(5 + M)
        payload = rng.integers(0, 2, (parity, 40), dtype=np.int8)
        msgs = np.array([O.attach_crc(p, CRC) for p in payload[:512]])
        codes = np.tile(np.array([O.encode(m, A, 128) for m in msgs]), ((parity + 511) // 512, 1))[:parity]
        x = (2.0 * (1.0 - 2.0 * codes + rng.normal(0, np.sqrt(nv(snr)), codes.shape)) / nv(snr)).astype(np.float32)
        ref = O.scl_decode_batch(x.astype(np.float64), A, M, crc=CRC, want_info_llrs=False)
        out = eng.scl_decode(x, M)
        cand = out["cand"].cpu().numpy().astype(np.int8); fl = out["flags"].cpu().numpy()
        diff = ~((cand == ref["cand"]).all(axis=(1, 2)) & (out["best_idx"].cpu().numpy() == ref["best_idx"]))
        m = out["metrics"].cpu().numpy(); fin = np.isfinite(ref["metrics"]) & ~diff[:, None]
        rel = np.abs(m[fin] - ref["metrics"][fin]) / np.maximum(np.abs(ref["metrics"][fin]), 1e-30)
        print(f"  parity M={M} snr={snr} B={parity}: mismatched={diff.sum()} unflagged={(diff & ((fl & 1) == 0)).sum()} flagged={(fl & 1).sum()} metric_rel={rel.max():.2e}", flush=True)
    for M, snr in [(4, 4.0), (8, 4.5)]:
        rng = np.random.default_rng(50 + M)
        Bd = max(parity // 4, 1000)
        payload = rng.integers(0, 2, (512, 40), dtype=np.int8)
        msgs = np.array([O.attach_crc(p, CRC) for p in payload])
        codes = np.tile(np.array([O.encode(m, A, 128) for m in msgs]), ((Bd + 511) // 512, 1))[:Bd]
        x = (2.0 * (1.0 - 2.0 * codes + rng.normal(0, np.sqrt(nv(snr)), codes.shape)) / nv(snr)).astype(np.float32)
        beta = g[f"beta_M{M}"]
        ref = O.dlscl_decode_batch(x.astype(np.float64), A, M, 8, crc=CRC, beta=beta)
        out = eng.dlscl_decode(x, M, 8, beta=beta)
        same = (out["best_bits"].cpu().numpy().astype(np.int8) == ref["best_bits"]).all(axis=1) & (out["n_attempts"].cpu().numpy() == ref["n_attempts"])
        fl = out["flags"].cpu().numpy()
        print(f"  parity DL M={M} snr={snr} B={Bd}: retried={(ref['n_attempts'] > 1).sum()} mismatched={(~same).sum()} unflagged={((~same) & ((fl & 3) == 0)).sum()} flagged={((fl & 3) != 0).sum()}", flush=True)
