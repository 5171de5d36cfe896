"""GPU: genericity the reference's own tests demand of a drop-in: any power-of-two N, any 0<K<=N, any
CRC polynomial, any list size 1..8 -- each checked frame-by-frame against the float64 oracle on seeded AWGN frames."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    # N, K, poly, snr_db, frames
    (2, 1, None, 3.0, 300), (2, 2, None, 3.0, 300), (4, 2, None, 2.0, 500), (4, 3, "0x3", 2.0, 500),
    (8, 5, "0x7", 2.0, 600), (16, 9, "0x17", 2.0, 800), (32, 17, "0x1D5", 2.0, 800), (64, 33, "0x107", 2.0, 800),
    (128, 41, "0x1864CFB", 1.0, 800), (128, 127, "0x1864CFB", 8.0, 400), (256, 131, "0x11021", 2.0, 500),
    (512, 256, "0x1864CFB", 2.5, 300), (512, 77, "0x104C11DB7", 0.0, 200),
]


def _frames(rng, N, K, poly, snr, B):
    A = O.construct_info_set(N, K)
    deg = (int(poly, 16).bit_length() - 1) if poly else 0
    msgs = []
    for _ in range(B):
        p = rng.integers(0, 2, K - deg, dtype=np.int8)
        msgs.append(O.attach_crc(p, poly) if poly else p)
    msgs = np.array(msgs, np.int8)
    codes = np.array([O.encode(m, A, N) for m in msgs])
    nv = 1.0 / (2.0 * (K / N) * 10 ** (snr / 10))
    llr = (2.0 * (1.0 - 2.0 * codes + rng.normal(0, np.sqrt(nv), codes.shape)) / nv).astype(np.float32)
    return A, msgs, llr


@pytest.mark.parametrize("N,K,poly,snr,B", CASES)
def test_all_list_sizes_match_oracle(N, K, poly, snr, B):
    from polar_code_b200.engine import PolarEngine
    rng = np.random.default_rng(N * 1000 + K)
    A, msgs, llr = _frames(rng, N, K, poly, snr, B)
    eng = PolarEngine(N, A, poly)
    assert np.array_equal(eng.encode(msgs.astype(np.uint8)).cpu().numpy().astype(np.int8),
                          np.array([O.encode(m, A, N) for m in msgs]))
    assert np.array_equal(eng.sc_decode(llr).cpu().numpy().astype(np.int8), O.sc_decode_batch(llr.astype(np.float64), A))
    for M in (1, 2, 3, 4, 5, 6, 7, 8):
        ref = O.scl_decode_batch(llr.astype(np.float64), A, M, crc=poly)
        out = eng.scl_decode(llr, M, want=("cand", "metrics", "n_cand", "best_idx", "best_bits", "crc_ok", "flags", "info_llrs"))
        flags = out["flags"].cpu().numpy()
        cand = out["cand"].cpu().numpy().astype(np.int8)
        same = ((cand == ref["cand"]).all(axis=(1, 2)) & (out["best_idx"].cpu().numpy() == ref["best_idx"])
                & (out["n_cand"].cpu().numpy() == ref["n_cand"]))
        assert not (~same & ((flags & 1) == 0)).any(), f"N={N} K={K} M={M}: unflagged mismatch"
        assert (~same).sum() <= max(2, B // 100)
        m = out["metrics"].cpu().numpy()
        fin = np.isfinite(ref["metrics"]) & same[:, None]
        np.testing.assert_allclose(m[fin], ref["metrics"][fin], rtol=1e-4, atol=1e-6)
        il = out["info_llrs"].cpu().numpy()
        ok = np.isfinite(ref["metrics"])[:, :, None] & same[:, None, None] & np.ones_like(il, bool)
        np.testing.assert_allclose(il[ok], ref["info_llrs"][ok], rtol=1e-4, atol=5e-4)
        if poly:
            bb = out["best_bits"].cpu().numpy().astype(np.int8)
            okc = out["crc_ok"].cpu().numpy().astype(bool)
            for b in range(0, B, 37):
                assert okc[b] == O.check_crc(bb[b], poly)


@pytest.mark.parametrize("N,K,poly,M,retries", [(16, 12, "0x17", 2, 4), (64, 40, "0x107", 4, 8), (256, 120, "0x1864CFB", 4, 3),
                                                  (128, 88, "0x1864CFB", 8, 8), (128, 64, "0x1864CFB", 3, 5), (512, 200, "0x1864CFB", 2, 2)])
def test_dlscl_generic_matches_oracle(N, K, poly, M, retries):
    from polar_code_b200.engine import PolarEngine
    rng = np.random.default_rng(N + K + M)
    A, msgs, llr = _frames(rng, N, K, poly, 1.5, 400)
    beta = (np.eye(K) + 0.1 * rng.normal(size=(K, K))).astype(np.float32)
    beta = ((beta + beta.T) / 2).astype(np.float32)
    eng = PolarEngine(N, A, poly)
    for b in (None, beta):
        ref = O.dlscl_decode_batch(llr.astype(np.float64), A, M, retries, crc=poly, beta=b)
        out = eng.dlscl_decode(llr, M, retries, beta=b)
        flags = out["flags"].cpu().numpy()
        same = ((out["best_bits"].cpu().numpy().astype(np.int8) == ref["best_bits"]).all(axis=1)
                & (out["success"].cpu().numpy().astype(bool) == ref["success"])
                & (out["n_attempts"].cpu().numpy() == ref["n_attempts"])
                & (out["tried"].cpu().numpy() == ref["tried"]).all(axis=1))
        assert not (~same & ((flags & 3) == 0)).any(), f"unflagged DL-SCL mismatch N={N} M={M}"
        assert (~same).sum() <= 6
        assert ref["n_attempts"].max() > 1


def test_dlscl_through_rate_matched_input():
    """DL-SCL on E-long rate-matched LLRs (de-rate-match + de-interleave fused into the retry kernel's load) equals the
    oracle's decode_with_retries on the oracle's de-rate-matched vector."""
    from polar_code_b200.engine import PolarEngine
    rng = np.random.default_rng(99)
    N, K, E, M, poly = 128, 88, 256, 4, "0x1864CFB"
    A = O.construct_info_set(N, K)
    eng = PolarEngine(N, A, poly)
    eng.set_rate_matching(E)
    nv = O.noise_var_ber(3.5, 64, E)
    rows, internal = [], []
    for _ in range(600):
        _, l = O.ber_frame(rng, "nr_polar_scl", 64, 24, poly, N, E, A, nv)
        l32 = l.astype(np.float32)
        rows.append(l32)
        internal.append(O.subblock_deinterleave(O.derate_match(l32.astype(np.float64), N), N))
    rows = np.array(rows)
    ref = O.dlscl_decode_batch(np.array(internal), A, M, 8, crc=poly)
    out = eng.dlscl_decode(rows, M, 8)
    flags = out["flags"].cpu().numpy()
    same = ((out["best_bits"].cpu().numpy().astype(np.int8) == ref["best_bits"]).all(axis=1)
            & (out["success"].cpu().numpy().astype(bool) == ref["success"])
            & (out["n_attempts"].cpu().numpy() == ref["n_attempts"])
            & (out["tried"].cpu().numpy() == ref["tried"]).all(axis=1))
    assert not (~same & ((flags & 3) == 0)).any()
    assert (~same).sum() <= 6 and ref["n_attempts"].max() > 2


@pytest.mark.parametrize("N,K,poly", [(128, 64, "0x1864CFB"), (64, 33, "0x107"), (32, 17, None)])
def test_zero_and_equal_magnitude_llrs(N, K, poly):
    """polar.py:122-123: f = sign(a) sign(b) min(|a|,|b|) with np.sign(0) = 0; the kernels' one-instruction f
    (min.xorsign.abs) returns a signed zero there.  Rows full of exact zeros, +-equal magnitudes and coarse integer LLRs
    (many exact metric ties): SC decisions must be identical, list decisions identical outside flagged frames, and the
    cooperative phase 0 (N = 64, 128) sees the same rows."""
    from polar_code_b200.engine import PolarEngine
    rng = np.random.default_rng(N + 7)
    A = O.construct_info_set(N, K)
    B = 600
    llr = rng.integers(-3, 4, (B, N)).astype(np.float32)              # values in {-3..3}: ~14 % exact zeros, many ties
    llr[:100] *= rng.integers(0, 2, (100, N)).astype(np.float32)       # half of the entries zeroed
    llr[100:150] = 0.0                                                 # all-zero rows
    llr[150:200] = np.where(rng.integers(0, 2, (50, N)) > 0, 2.5, -2.5).astype(np.float32)   # equal magnitudes everywhere
    eng = PolarEngine(N, A, poly)
    assert np.array_equal(eng.sc_decode(llr).cpu().numpy().astype(np.int8), O.sc_decode_batch(llr.astype(np.float64), A))
    for M in (1, 2, 4, 8):
        ref = O.scl_decode_batch(llr.astype(np.float64), A, M, crc=poly)
        out = eng.scl_decode(llr, M, want=("cand", "metrics", "n_cand", "best_idx", "flags"))
        flags = out["flags"].cpu().numpy()
        same = ((out["cand"].cpu().numpy().astype(np.int8) == ref["cand"]).all(axis=(1, 2))
                & (out["best_idx"].cpu().numpy() == ref["best_idx"]) & (out["n_cand"].cpu().numpy() == ref["n_cand"]))
        assert not (~same & ((flags & 1) == 0)).any(), f"N={N} M={M}: unflagged mismatch on tied / zero LLRs"
        fin = np.isfinite(ref["metrics"]) & same[:, None]
        np.testing.assert_allclose(out["metrics"].cpu().numpy()[fin], ref["metrics"][fin], rtol=1e-4, atol=1e-6)
