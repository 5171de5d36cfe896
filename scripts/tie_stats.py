"""GPU vs oracle mismatch / near-tie statistics (run on the GPU box)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import oracle as O
from polar_code_b200.engine import PolarEngine
CRC = "0x1864CFB"
A = O.construct_info_set(128, 64)
eng = PolarEngine(128, A, CRC)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for M, snr in [(1, 3.0), (2, 3.0), (4, 2.0), (4, 4.0), (4, 5.0), (8, 3.0), (8, 5.0)]:
    rng = np.random.default_rng(5 + M)
    nv = 1.0 / (2.0 * 0.5 * 10 ** (snr / 10))
    payload = rng.integers(0, 2, (B, 40), dtype=np.int8)
    msgs = np.array([O.attach_crc(p, CRC) for p in payload])
    codes = np.array([O.encode(m, A, 128) for m in msgs])
    llr = (2.0 * (1.0 - 2.0 * codes + rng.normal(0, np.sqrt(nv), codes.shape)) / nv).astype(np.float32)
    t = time.time()
    ref = O.scl_decode_batch(llr.astype(np.float64), A, M, crc=CRC, want_info_llrs=False)
    tc = time.time() - t
    out = eng.scl_decode(llr, M)
    cand = out["cand"].cpu().numpy().astype(np.int8); flags = out["flags"].cpu().numpy()
    diff = ~((cand == ref["cand"]).all(axis=(1, 2)) & (out["best_idx"].cpu().numpy() == ref["best_idx"]))
    bestdiff = ~(out["best_bits"].cpu().numpy().astype(np.int8) == ref["best_bits"]).all(axis=1)
    m = out["metrics"].cpu().numpy(); fin = np.isfinite(ref["metrics"]) & ~diff[:, None]
    rel = np.abs(m[fin] - ref["metrics"][fin]) / np.maximum(np.abs(ref["metrics"][fin]), 1e-30)
    print(f"M={M} snr={snr} B={B}: mismatched frames={diff.sum()} (best word differs {bestdiff.sum()}), unflagged mismatches={(diff & ((flags&1)==0)).sum()}, "
          f"flagged={(flags&1).sum()}, oracle gap<1e-6: {(ref['min_gap']<1e-6).sum()}, gap<4e-6: {(ref['min_gap']<4e-6).sum()}, "
          f"max metric rel err={rel.max():.2e}, oracle {B/tc:.0f} frames/s on {O.lib().po_num_threads()} threads", flush=True)

# DL-SCL (decode_with_retries, shipped-style beta = identity-like ranking when none given): last attempt vs oracle
g128 = np.load("tests/golden/scl_p128.npz")
for M, snr, R, shipped in [(4, 4.0, 8, False), (8, 4.5, 8, False), (4, 4.0, 8, True), (8, 4.0, 8, True), (4, 3.0, 8, True)]:
    rng = np.random.default_rng(50 + M + (7 if shipped else 0))
    Bd = max(B // 4, 1000)
    nv = 1.0 / (2.0 * 0.5 * 10 ** (snr / 10))
    payload = rng.integers(0, 2, (Bd, 40), dtype=np.int8)
    msgs = np.array([O.attach_crc(p, CRC) for p in payload])
    codes = np.array([O.encode(m, A, 128) for m in msgs])
    llr = (2.0 * (1.0 - 2.0 * codes + rng.normal(0, np.sqrt(nv), codes.shape)) / nv).astype(np.float32)
    beta = g128[f"beta_M{M}"] if shipped else (np.eye(64) + 0.05 * rng.standard_normal((64, 64))).astype(np.float32)
    ref = O.dlscl_decode_batch(llr.astype(np.float64), A, M, R, crc=CRC, beta=beta)
    out = eng.dlscl_decode(llr, M, R, beta=beta)
    same = (out["best_bits"].cpu().numpy().astype(np.int8) == ref["best_bits"]).all(axis=1)
    same &= out["n_attempts"].cpu().numpy() == ref["n_attempts"]
    fl = out["flags"].cpu().numpy()
    print(f"DL-SCL M={M} snr={snr} retries={R} beta={'shipped' if shipped else 'random'} B={Bd}: frames with retries={(ref['n_attempts'] > 1).sum()}, mismatched={(~same).sum()}, "
          f"unflagged mismatches={((~same) & ((fl & 3) == 0)).sum()}, flagged={((fl & 3) != 0).sum()}", flush=True)
