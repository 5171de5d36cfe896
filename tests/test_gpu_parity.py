"""GPU parity: CUDA engine (through the C-ABI) vs golden vectors produced by the reference and vs the oracle.

Bar: bit-exact decisions on identical (float32-representable) LLRs.  The one allowed exception
(BASELINE.json north_star) is a frame whose competing path metrics tie within ~1e-6 relative; the engine
flags those (PB200_FLAG_NEAR_TIE) and every mismatch must be a flagged frame.  Metrics / LLRs: fp32 LLR
arithmetic + fp64 metric accumulation vs the reference's float64, tolerance 1e-4 relative (north_star).
"""
import numpy as np
import pytest
import torch

from conftest import CRC24
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # north_star: fp32 LLRs / metrics within 1e-4 relative of the float64 reference


@pytest.fixture(scope="module")
def eng128(g128):
    from polar_code_b200.engine import PolarEngine
    return PolarEngine(128, g128["info_set"], CRC24)


def _np(t):
    return t.cpu().numpy()


def _check_scl(out, gold, tag, B, allow_flagged=True):
    flags = _np(out["flags"])
    cand, n_cand = _np(out["cand"]).astype(np.int8), _np(out["n_cand"])
    bad = [b for b in range(B) if not (n_cand[b] == gold[tag + "_n_cand"][b]
                                       and np.array_equal(cand[b], gold[tag + "_cand"][b])
                                       and _np(out["best_idx"])[b] == gold[tag + "_best"][b])]
    for b in bad:
        assert allow_flagged and (flags[b] & 1), f"{tag}: frame {b} differs from the reference without a near-tie flag"
    ok = np.array([b not in bad for b in range(B)])
    m = _np(out["metrics"])
    gm = gold[tag + "_metrics"][:B]
    fin = np.isfinite(gm) & ok[:, None]
    assert np.array_equal(np.isfinite(m)[ok], np.isfinite(gm)[ok])
    np.testing.assert_allclose(m[fin], gm[fin], rtol=RTOL, atol=1e-6)
    return len(bad)


@pytest.mark.parametrize("M", [1, 2, 4, 8])
def test_scl_golden(eng128, g128, M):
    B = g128["llr"].shape[0]
    out = eng128.scl_decode(g128["llr"], M, want=("cand", "metrics", "n_cand", "best_idx", "best_bits", "crc_ok",
                                                    "flags", "info_llrs", "best_words"))
    tag = f"scl_M{M}"
    nbad = _check_scl(out, g128, tag, B)
    assert nbad <= 2  # frame 0 (+-1e6) and frame 1 (all-zero LLRs) are saturated / fully tied by construction
    # best_bits == cand[best]; crc_ok == check_crc(best_bits)
    bb = _np(out["best_bits"]).astype(np.int8)
    for b in range(B):
        assert np.array_equal(bb[b], _np(out["cand"])[b, _np(out["best_idx"])[b]])
        assert bool(_np(out["crc_ok"])[b]) == O.check_crc(bb[b], CRC24)
    # info_llrs of every candidate (scl.py:159,167), fp32 vs float64
    il = _np(out["info_llrs"])
    for b in range(2, B):
        if np.array_equal(_np(out["cand"])[b].astype(np.int8), g128[tag + "_cand"][b]):
            nc = g128[tag + "_n_cand"][b]
            np.testing.assert_allclose(il[b, :nc], g128[tag + "_info_llrs"][b, :nc], rtol=RTOL, atol=2e-4)


def test_sc_golden(eng128, g128):
    out = _np(eng128.sc_decode(g128["llr"])).astype(np.int8)
    # frame 1 is all-zero LLRs (every decision is a tie: llr<0 is False -> all zeros in both)
    assert np.array_equal(out, g128["sc_bits"])


def test_scl_no_crc_M3(g128):
    from polar_code_b200.engine import PolarEngine
    e = PolarEngine(128, g128["info_set"], None)
    out = e.scl_decode(g128["llr"][:16], 3)
    _check_scl(out, g128, "scl_M3_nocrc", 16)
    assert not _np(out["best_idx"]).any()


def test_scl_forced(eng128, g128):
    out = eng128.scl_decode(g128["llr"], 4, force=g128["force"])
    _check_scl(out, g128, "scl_M4_forced", g128["llr"].shape[0])
    bad = g128["force"].copy()
    bad[5, 7] = 3
    out = eng128.scl_decode(g128["llr"], 4, force=bad)
    assert _np(out["flags"])[5] & 4 and not (_np(out["flags"])[4] & 4)


@pytest.mark.parametrize("tag,poly", [("n16", "0x17"), ("n8", None), ("n32", "0x1D5"), ("n256", CRC24)])
def test_toy_sizes(gtoy, tag, poly):
    from polar_code_b200.engine import PolarEngine
    A = gtoy[tag + "_info_set"]
    llr = gtoy[tag + "_llr"]
    e = PolarEngine(llr.shape[1], A, poly)
    for M in (1, 2, 4):
        out = e.scl_decode(llr, M)
        _check_scl(out, gtoy, f"{tag}_M{M}", llr.shape[0], allow_flagged=True)


@pytest.mark.parametrize("tag,M,retries,beta", [("dl_M1", 1, 8, "beta_M1"), ("dl_M2", 2, 8, "beta_M2"),
                                               ("dl_M4", 4, 8, "beta_M4"), ("dl_M8", 8, 8, "beta_M8"),
                                               ("dl_M2_nobeta_r4", 2, 4, None), ("dl_M4_r0", 4, 0, None)])
def test_dlscl_golden(eng128, g128, tag, M, retries, beta):
    n = g128[tag + "_bits"].shape[0]
    out = eng128.dlscl_decode(g128["llr"][:n], M, retries, beta=None if beta is None else g128[beta])
    flags = _np(out["flags"])
    for b in range(n):
        same = (np.array_equal(_np(out["best_bits"])[b].astype(np.int8), g128[tag + "_bits"][b])
                and bool(_np(out["success"])[b]) == bool(g128[tag + "_success"][b])
                and _np(out["n_attempts"])[b] == g128[tag + "_n_attempts"][b]
                and np.array_equal(_np(out["tried"])[b], g128[tag + "_tried"][b]))
        assert same or (flags[b] & 3), f"{tag}: frame {b} differs without a tie flag"


def test_encode_crc(eng128, g128):
    from polar_code_b200 import engine as E
    code = _np(eng128.encode(g128["msgs"].astype(np.uint8))).astype(np.int8)
    assert np.array_equal(code, g128["codes"])
    att = _np(E.crc_attach(g128["msgs"][:, :40].astype(np.uint8), CRC24)).astype(np.int8)
    assert np.array_equal(att, g128["msgs"])
    ok = _np(E.crc_check(g128["msgs"].astype(np.uint8), CRC24))
    assert ok.all()
    bad = g128["msgs"].astype(np.uint8).copy()
    bad[:, 13] ^= 1
    assert not _np(E.crc_check(bad, CRC24)).any()
    assert np.array_equal(_np(E.crc_attach(np.array([[1] + [0] * 39], np.uint8), CRC24))[0].astype(np.int8), g128["crc_kat"])
    with pytest.raises(ValueError):
        E.crc_check(np.zeros((1, 24), np.uint8), CRC24)


def test_beta_ranking(g128):
    from polar_code_b200 import engine as E
    ib = _np(E.choose_flip_index(g128["rank_abs_l0"], g128["beta_M4"]))
    i0 = _np(E.choose_flip_index(g128["rank_abs_l0"], None))
    assert np.array_equal(ib, g128["rank_idx_beta"])
    assert np.array_equal(i0, g128["rank_idx_none"])


@pytest.mark.parametrize("E", [256, 96, 128, 300])
def test_nr_chain(gnr, E):
    from polar_code_b200.engine import PolarEngine
    e = PolarEngine(128, gnr["info_set"], CRC24)
    e.set_rate_matching(E)
    out = e.scl_decode(gnr[f"E{E}_llr"], 4)
    flags = _np(out["flags"])
    bb = _np(out["best_bits"]).astype(np.int8)
    for b in range(bb.shape[0]):
        assert np.array_equal(bb[b], gnr[f"E{E}_bits"][b]) or (flags[b] & 1)
        assert bool(_np(out["crc_ok"])[b]) == bool(gnr[f"E{E}_crc_pass"][b]) or (flags[b] & 1)
    tx = _np(e.nr_encode(gnr[f"E{E}_payload"].astype(np.uint8), E))
    for b in range(tx.shape[0]):
        ref = O.rate_match(O.subblock_interleave(O.encode(O.attach_crc(gnr[f"E{E}_payload"][b], CRC24), gnr["info_set"], 128).astype(np.float64)), E)
        assert np.array_equal(tx[b].astype(np.float64), ref)


@pytest.mark.parametrize("M,snr", [(1, 2.0), (4, 2.0), (4, 4.0), (8, 3.0), (2, 1.0)])
def test_scl_vs_oracle_random(eng128, g128, M, snr):
    """4096 seeded frames per case: decisions identical to the float64 oracle except flagged near-ties."""
    rng = np.random.default_rng(1000 + M)
    A = g128["info_set"]
    B = 4096
    nv = 1.0 / (2.0 * 0.5 * 10 ** (snr / 10))
    payload = rng.integers(0, 2, (B, 40), dtype=np.int8)
    msgs = np.array([O.attach_crc(p, CRC24) for p in payload])
    codes = np.array([O.encode(m, A, 128) for m in msgs])
    llr = (2.0 * (1.0 - 2.0 * codes + rng.normal(0, np.sqrt(nv), codes.shape)) / nv).astype(np.float32)
    ref = O.scl_decode_batch(llr.astype(np.float64), A, M, crc=CRC24, want_info_llrs=False)
    out = eng128.scl_decode(llr, M)
    cand = _np(out["cand"]).astype(np.int8)
    flags = _np(out["flags"])
    diff = np.array([not (np.array_equal(cand[b], ref["cand"][b]) and _np(out["best_idx"])[b] == ref["best_idx"][b])
                     for b in range(B)])
    assert not (diff & ((flags & 1) == 0)).any(), "unflagged mismatch vs oracle"
    assert diff.sum() <= 8, f"{diff.sum()} mismatching frames (all flagged)"
    # every frame the oracle sees as a near tie (<1e-6 relative) must carry the flag
    assert ((flags & 1) != 0)[ref["min_gap"] < 1e-6].all()
    assert (flags & 1).mean() <= 0.03
    m = _np(out["metrics"])
    good = ~diff
    fin = np.isfinite(ref["metrics"]) & good[:, None]
    np.testing.assert_allclose(m[fin], ref["metrics"][fin], rtol=RTOL, atol=1e-6)


def test_dlscl_vs_oracle_random(eng128, g128):
    rng = np.random.default_rng(77)
    A = g128["info_set"]
    B = 2048
    nv = 1.0 / (2.0 * 0.5 * 10 ** 0.3)
    payload = rng.integers(0, 2, (B, 40), dtype=np.int8)
    msgs = np.array([O.attach_crc(p, CRC24) for p in payload])
    codes = np.array([O.encode(m, A, 128) for m in msgs])
    llr = (2.0 * (1.0 - 2.0 * codes + rng.normal(0, np.sqrt(nv), codes.shape)) / nv).astype(np.float32)
    ref = O.dlscl_decode_batch(llr.astype(np.float64), A, 4, 8, crc=CRC24, beta=g128["beta_M4"])
    out = eng128.dlscl_decode(llr, 4, 8, beta=g128["beta_M4"])
    flags = _np(out["flags"])
    same = ((_np(out["best_bits"]).astype(np.int8) == ref["best_bits"]).all(axis=1)
            & (_np(out["success"]).astype(bool) == ref["success"])
            & (_np(out["n_attempts"]) == ref["n_attempts"])
            & (_np(out["tried"]) == ref["tried"]).all(axis=1))
    assert not (~same & ((flags & 3) == 0)).any(), "unflagged DL-SCL mismatch vs oracle"
    assert (~same).sum() <= 8
    assert ref["n_attempts"].mean() > 1.5   # the case really exercises the retry rounds


def test_ragged_and_empty(eng128, g128):
    """Batch sizes that do not fill a warp / the grid, and B = 0."""
    llr = np.tile(g128["llr"][2:10], (5, 1))
    full = eng128.scl_decode(llr, 4)
    for B in (1, 3, 7, 9, 33):
        part = eng128.scl_decode(llr[:B], 4)
        assert torch.equal(part["cand"], full["cand"][:B]) and torch.equal(part["best_idx"], full["best_idx"][:B])
    out = eng128.scl_decode(np.zeros((0, 128), np.float32), 4)
    assert out["cand"].shape[0] == 0


def test_error_mapping(g128):
    from polar_code_b200.engine import PolarEngine
    e = PolarEngine(128, g128["info_set"], CRC24)
    with pytest.raises(ValueError):
        e.scl_decode(g128["llr"], 0)
    with pytest.raises(NotImplementedError):
        e.scl_decode(g128["llr"], 9)
    with pytest.raises(ValueError):
        e.scl_decode(g128["llr"][:, :64], 4)
    with pytest.raises(ValueError):
        PolarEngine(100, g128["info_set"], CRC24)
    with pytest.raises(ValueError):
        PolarEngine(128, g128["info_set"][:20], CRC24)  # message too short for CRC-24
