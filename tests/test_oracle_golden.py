"""CPU: pin the oracle (oracle/polar_oracle.c) against vectors produced by the reference itself
(tests/golden/*.npz, written by oracle/gen_golden.py) and against the published results/*.csv."""
import numpy as np
import pytest

from conftest import CRC24
from oracle import oracle as O

A64 = [7, 11, 13, 15, 19, 21, 23, 27, 29, 30, 31, 35, 39, 43, 45, 46, 47, 51, 53, 54, 55, 57, 58, 59, 61, 62, 63,
       71, 75, 77, 78, 79, 83, 85, 86, 87, 89, 90, 91, 93, 94, 95, 99, 101, 102, 103, 105, 106, 107, 109, 110,
       111, 113, 115, 117, 118, 119, 121, 122, 123, 124, 125, 126, 127]


def test_info_sets(g128):
    # SURVEY 8 golden list and polar/polar.py:85-103 outputs
    assert list(O.construct_info_set(128, 64)) == A64
    assert np.array_equal(O.construct_info_set(128, 64), g128["info_set"])
    assert np.array_equal(O.construct_info_set(128, 88), g128["info_set_88"])
    assert np.array_equal(O.construct_info_set(128, 64, "polarization"), g128["info_set_pw"])
    with pytest.raises(ValueError):
        O.construct_info_set(100, 10)
    with pytest.raises(ValueError):
        O.construct_info_set(16, 0)


def test_crc_kat(g128):
    msg = np.array([1] + [0] * 39, np.int8)
    out = O.attach_crc(msg, CRC24)
    assert "".join(map(str, out[40:])) == "110110000101001111100011"
    assert np.array_equal(out, g128["crc_kat"])
    assert O.check_crc(out, CRC24)
    bad = out.copy()
    bad[17] ^= 1
    assert not O.check_crc(bad, CRC24)
    with pytest.raises(ValueError):
        O.check_crc(np.zeros(24, np.int8), CRC24)


def test_encode(g128):
    A = g128["info_set"]
    for m, c in zip(g128["msgs"], g128["codes"]):
        assert np.array_equal(O.encode(m, A, 128), c)
        assert np.array_equal(O.attach_crc(m[:40], CRC24), m)


def test_sc(g128):
    out = O.sc_decode_batch(g128["llr"].astype(np.float64), g128["info_set"])
    assert np.array_equal(out, g128["sc_bits"])


@pytest.mark.parametrize("M", [1, 2, 4, 8])
def test_scl_exact(g128, M):
    o = O.scl_decode_batch(g128["llr"].astype(np.float64), g128["info_set"], M, crc=CRC24)
    t = f"scl_M{M}"
    assert np.array_equal(o["n_cand"], g128[t + "_n_cand"])
    assert np.array_equal(o["cand"], g128[t + "_cand"])
    assert np.array_equal(o["metrics"], g128[t + "_metrics"])       # bit-identical float64
    assert np.array_equal(o["info_llrs"], g128[t + "_info_llrs"])
    assert np.array_equal(o["best_idx"], g128[t + "_best"])


def test_scl_nocrc_and_forced(g128):
    A = g128["info_set"]
    llr = g128["llr"].astype(np.float64)
    o = O.scl_decode_batch(llr[:16], A, 3, crc=None)
    assert np.array_equal(o["cand"], g128["scl_M3_nocrc_cand"])
    assert np.array_equal(o["best_idx"], g128["scl_M3_nocrc_best"]) and not o["best_idx"].any()
    o = O.scl_decode_batch(llr, A, 4, crc=CRC24, force=g128["force"])
    assert np.array_equal(o["cand"], g128["scl_M4_forced_cand"])
    assert np.array_equal(o["metrics"], g128["scl_M4_forced_metrics"])
    assert np.array_equal(o["n_cand"], g128["scl_M4_forced_n_cand"])
    bad = g128["force"].copy()
    bad[0, 0] = 2
    with pytest.raises(ValueError):
        O.scl_decode_batch(llr, A, 4, crc=CRC24, force=bad)


@pytest.mark.parametrize("tag,M,retries,beta", [("dl_M1", 1, 8, "beta_M1"), ("dl_M2", 2, 8, "beta_M2"),
                                               ("dl_M4", 4, 8, "beta_M4"), ("dl_M8", 8, 8, "beta_M8"),
                                               ("dl_M2_nobeta_r4", 2, 4, None), ("dl_M4_r0", 4, 0, None)])
def test_dlscl_exact(g128, tag, M, retries, beta):
    n = g128[tag + "_bits"].shape[0]
    o = O.dlscl_decode_batch(g128["llr"][:n].astype(np.float64), g128["info_set"], M, retries, crc=CRC24,
                             beta=None if beta is None else g128[beta])
    assert np.array_equal(o["best_bits"], g128[tag + "_bits"])
    assert np.array_equal(o["success"], g128[tag + "_success"].astype(bool))
    assert np.array_equal(o["n_attempts"], g128[tag + "_n_attempts"])
    assert np.array_equal(o["tried"], g128[tag + "_tried"])


def test_beta_ranking(g128):
    for a, ib, i0 in zip(g128["rank_abs_l0"], g128["rank_idx_beta"], g128["rank_idx_none"]):
        assert O.choose_flip_index(a.astype(np.float64), g128["beta_M4"]) == ib
        assert O.choose_flip_index(a.astype(np.float64), None) == i0


@pytest.mark.parametrize("tag,poly", [("n16", "0x17"), ("n8", None), ("n32", "0x1D5"), ("n256", CRC24)])
def test_toy_sizes(gtoy, tag, poly):
    A = gtoy[tag + "_info_set"]
    llr = gtoy[tag + "_llr"].astype(np.float64)
    assert np.array_equal(O.construct_info_set(llr.shape[1], A.size), A)
    for M in (1, 2, 4):
        o = O.scl_decode_batch(llr, A, M, crc=poly)
        assert np.array_equal(o["cand"], gtoy[f"{tag}_M{M}_cand"])
        assert np.array_equal(o["metrics"], gtoy[f"{tag}_M{M}_metrics"])
        assert np.array_equal(o["best_idx"], gtoy[f"{tag}_M{M}_best"])


def test_nr_chain(gnr):
    assert np.array_equal(O.subblock_interleave(np.arange(40.0)), gnr["ilv40"])
    assert np.array_equal(O.subblock_deinterleave(gnr["ilv40"], 40), gnr["deilv40"])
    assert np.array_equal(gnr["deilv40"], np.arange(40.0))
    assert np.array_equal(O.subblock_interleave(np.arange(128.0)), gnr["ilv128"])
    A = gnr["info_set"]
    for E in (256, 96, 128, 300):
        llr = gnr[f"E{E}_llr"].astype(np.float64)
        internal = np.array([O.subblock_deinterleave(O.derate_match(l, 128), 128) for l in llr])
        assert np.array_equal(internal, gnr[f"E{E}_internal"])
        o = O.nr_decode_batch(llr, CRC24, 128, A, 4)
        assert np.array_equal(o["best_bits"], gnr[f"E{E}_bits"])
        assert np.array_equal(o["crc_pass"], gnr[f"E{E}_crc_pass"].astype(bool))
        tx = O.rate_match(O.subblock_interleave(np.arange(128.0)), E)
        assert np.array_equal(tx.astype(np.int64), gnr[f"E{E}_tx"][0])


def _fer_row(M, snr, frames, beta):
    A = O.construct_info_set(128, 64)
    msgs, llrs, unc = O.fer_sweep_frames(snr, frames, seed=0, include_uncoded=True)
    s = O.scl_decode_batch(llrs, A, M, crc=CRC24, want_info_llrs=False)
    d = O.dlscl_decode_batch(llrs, A, M, 8, crc=CRC24, beta=beta)
    fe = lambda bits: sum(not O.check_crc(b, CRC24) for b in bits)  # run_fer_sweep.py:91-94: CRC failure
    tb = frames * 64
    vals = [f"{snr:.3f}", f"{(unc > 0).sum() / frames:.6e}", f"{unc.sum() / (frames * 40):.6e}",
            f"{fe(s['best_bits']) / frames:.6e}", f"{int((s['best_bits'] != msgs).sum()) / tb:.6e}",
            f"{fe(d['best_bits']) / frames:.6e}", f"{int((d['best_bits'] != msgs).sum()) / tb:.6e}"]
    return ",".join(vals)


@pytest.mark.parametrize("name", ["fer_M4", "fer_M8", "fer_M1"])
def test_published_csv_reproduced(g128, published, name):
    """results/fer_M{1,4,8}.csv byte-for-byte: PCG64 channel restatement + oracle SCL + DL-SCL."""
    rec = published["recipe"][name]
    rows = published["rows"][name]
    assert rows[0] == "snr_db,fer_uncoded,ber_uncoded,fer_scl,ber_scl,fer_dl,ber_dl"
    beta = g128[f"beta_M{rec['M']}"]
    for snr, want in zip(rec["snr"], rows[1:]):
        assert _fer_row(rec["M"], snr, rec["frames"], beta) == want
