"""CPU: pin the LDPC oracle (oracle/ldpc_oracle.c) against vectors produced by the reference's nr/ldpc package
(tests/golden/ldpc.npz, written by `python oracle/gen_golden.py ldpc`) -- float64 outputs bit for bit."""
import json

import numpy as np
import pytest

from oracle import oracle as O

CASES = ["z2", "z4", "z8e61", "z8e130", "z32e384", "z4it0", "z4it3"]


def test_h_matrices(gldpc):
    for bg, Z in [(1, 2), (2, 2), (2, 4), (1, 8), (2, 32)]:
        assert np.array_equal(O.ldpc_build_h(bg, Z), gldpc[f"H_bg{bg}_Z{Z}"])
    with pytest.raises(ValueError):
        O.ldpc_build_h(3, 2)                       # basegraphs.py:40-41


@pytest.mark.parametrize("tag", CASES)
def test_encode_derate_decode(gldpc, tag):
    Z, E, max_iter = (int(v) for v in gldpc[f"{tag}_cfg"])
    alpha = float(gldpc[f"{tag}_alpha"][0])
    H = O.ldpc_build_h(2, Z)
    n = H.shape[1]
    for p, c in zip(gldpc[f"{tag}_payload"], gldpc[f"{tag}_code"]):
        cw = O.ldpc_encode(p, H)
        assert np.array_equal(cw, c)
        assert not ((H.astype(np.int64) @ cw) % 2).any()
    der = np.array([O.ldpc_derate_match(l, n) for l in gldpc[f"{tag}_llr"]])
    assert np.array_equal(der, gldpc[f"{tag}_derated"])                 # bit-identical float64
    r = O.ldpc_decode_batch(der, H, max_iter, alpha)
    assert np.array_equal(r["hard"], gldpc[f"{tag}_hard"])
    assert np.array_equal(r["iters_used"], gldpc[f"{tag}_iters"])
    assert np.array_equal(r["parity_ok"], gldpc[f"{tag}_ok"].astype(bool))


def test_no_early_stop(gldpc):
    H = O.ldpc_build_h(2, 4)
    r = O.ldpc_decode_batch(gldpc["z4_derated"], H, 6, 0.8, early_stop=False)
    assert np.array_equal(r["hard"], gldpc["z4_noearly_hard"])
    assert np.array_equal(r["iters_used"], gldpc["z4_noearly_iters"])
    assert np.array_equal(r["parity_ok"], gldpc["z4_noearly_ok"].astype(bool))


def test_encode_errors():
    H = O.ldpc_build_h(2, 2)
    with pytest.raises(ValueError):
        O.ldpc_encode(np.zeros(12, np.int8), H)    # encode.py:57-58
    Hbad = np.zeros((2, 4), np.int8)
    Hbad[0, 0] = Hbad[1, 1] = 1                    # H_par = 0: any non-zero syndrome has no solution (encode.py:35-37)
    with pytest.raises(ValueError):
        O.ldpc_encode(np.array([1, 0], np.int8), Hbad)
    assert np.array_equal(O.ldpc_encode(np.array([0, 0], np.int8), Hbad), [0, 0, 0, 0])


def test_rate_match_shapes():
    c = np.arange(12, dtype=np.int8)
    assert np.array_equal(O.ldpc_rate_match(c, 7), c[:7])
    assert np.array_equal(O.ldpc_rate_match(c, 17), np.concatenate([c, c[:5]]))
    d = O.ldpc_derate_match(np.arange(5, dtype=np.float64), 12)
    assert np.array_equal(d, [0, 1, 2, 3, 4] + [0] * 7)


def test_reference_cli_rows_reproduced(gldpc):
    """The reference's own run_ber_sweep --scheme nr_ldpc rows (PCG64 channel) from the oracle, value for value."""
    cli = json.loads(bytes(gldpc["cli_json"]).decode())
    for tag, rec in cli.items():
        a = dict(zip(rec["argv"][0::2], rec["argv"][1::2]))
        kp, kc, E, Z, bg = int(a["--K_payload"]), int(a["--K_crc"]), int(a["--E"]), int(a["--Z"]), int(a["--bg"])
        H = O.ldpc_build_h(bg, Z)
        rng = np.random.default_rng(int(a["--seed"]))
        grid = np.arange(float(a["--EbN0_lo"]), float(a["--EbN0_hi"]) + 1e-12, float(a["--EbN0_step"]))
        assert len(grid) == len(rec["rows"])
        for snr, row in zip(grid, rec["rows"]):
            bits, be, fe, frames, work = O.ldpc_ber_point(
                rng, float(snr), K_payload=kp, K_crc=kc, crc_poly=a["--crc_poly"], H=H, E=E,
                max_iter=int(a.get("--max_iter", 20)), alpha=float(a.get("--alpha", 0.8)),
                err_cap=int(a["--err_cap"]), bits_cap=float(a["--bits_cap"]))
            assert bits == row["bits_total"] and be == row["bit_errors"]
            assert fe / frames == row["fer"] and work / frames == row["avg_work"]


# ---- host logic of the product's LDPC path (C-ABI host entry points, no GPU) ------------------------------------
def test_parity_generator_matches_the_reference_encoder(gldpc):
    """parity = G * payload reproduces encode_ldpc (oracle, pinned to the reference) for every payload tried,
    including unit vectors, truncated payload lengths and a rank-deficient parity part."""
    from polar_code_b200.ldpc import parity_generator
    rng = np.random.default_rng(9)
    cases = [(O.ldpc_build_h(2, Z), None) for Z in (2, 4, 8, 32)]
    Hr = (rng.random((10, 24)) < 0.25).astype(np.int8)
    Hr[:, 14:] |= np.eye(10, dtype=np.int8)
    cases += [(Hr, None), (Hr, 9), (O.ldpc_build_h(1, 8), 17)]
    for H, k in cases:
        n = H.shape[1]
        k = n - H.shape[0] if k is None else k
        G, Cc = parity_generator(H, k)
        assert G.shape == (n - k, k) and Cc.shape[0] == 0
        payloads = np.concatenate([np.eye(k, dtype=np.int8), rng.integers(0, 2, (40, k), dtype=np.int8), np.zeros((1, k), np.int8)])
        for p in payloads:
            ref = O.ldpc_encode(p, H)
            assert np.array_equal((G.astype(np.int64) @ p) % 2, ref[k:])
    # rank-deficient parity part: the consistency rows flag exactly the payloads for which the reference raises
    Hd = (rng.random((8, 20)) < 0.3).astype(np.int8)
    Hd[5] = Hd[2]
    Hd[5, :12] ^= np.array([1, 0, 1] * 4, np.int8)          # same parity part as row 2, different systematic part
    G, Cc = parity_generator(Hd, 12)
    assert Cc.shape[0] >= 1
    seen = set()
    for p in rng.integers(0, 2, (200, 12), dtype=np.int8):
        bad = bool(((Cc.astype(np.int64) @ p) % 2).any())
        try:
            ref = O.ldpc_encode(p, Hd)
            assert not bad and np.array_equal((G.astype(np.int64) @ p) % 2, ref[12:])
        except ValueError:
            assert bad
        seen.add(bad)
    assert seen == {True, False}
    with pytest.raises(ValueError):
        parity_generator(O.ldpc_build_h(2, 2), 12)


def test_layers_are_column_disjoint_and_cover_all_rows():
    from polar_code_b200.ldpc import layers
    rng = np.random.default_rng(4)
    for H, want_lanes in [(O.ldpc_build_h(2, 2), 0), (O.ldpc_build_h(2, 4), 4), (O.ldpc_build_h(1, 8), 8), (O.ldpc_build_h(2, 32), 32),
                          (O.ldpc_build_h(2, 50), 32), ((rng.random((12, 30)) < 0.2).astype(np.int8), None)]:
        lp, lanes = layers(H)
        assert lp[0] == 0 and lp[-1] == H.shape[0] and np.all(np.diff(lp) > 0) and np.diff(lp).max() <= 32
        for a, b in zip(lp[:-1], lp[1:]):
            assert H[a:b].astype(np.int64).sum(axis=0).max() <= 1          # rows of a layer share no column
            if b < H.shape[0] and b - a < 32:
                assert (H[a:b].sum(axis=0) * H[b]).any()                   # greedy: the next row really clashes
        if want_lanes is not None:
            assert lanes == want_lanes
        else:
            assert lanes in (0, 4, 8, 16, 32) and (lanes == 0) == (np.diff(lp).max() < 4)
