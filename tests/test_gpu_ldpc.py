"""GPU: the toy NR-LDPC family through the C-ABI vs the float64 oracle and the vectors the
reference itself produced (tests/golden/ldpc.npz).  Everything is float64 -> results must be bit-identical."""
import json

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

CASES = ["z2", "z4", "z8e61", "z8e130", "z32e384", "z4it0", "z4it3"]


@pytest.fixture(scope="module")
def engines():
    from polar_code_b200.ldpc import LdpcEngine, build_h_matrix
    cache = {}

    def get(Z, bg=2):
        if (bg, Z) not in cache:
            cache[(bg, Z)] = LdpcEngine(build_h_matrix(bg, Z))
        return cache[(bg, Z)]
    return get


def test_build_h_matches_golden(gldpc):
    from polar_code_b200.ldpc import build_h_matrix
    for bg, Z in [(1, 2), (2, 2), (2, 4), (1, 8), (2, 32)]:
        assert np.array_equal(build_h_matrix(bg, Z), gldpc[f"H_bg{bg}_Z{Z}"])
    with pytest.raises(ValueError):
        build_h_matrix(3, 2)


@pytest.mark.parametrize("tag", CASES)
def test_golden_encode_derate_decode(gldpc, engines, tag):
    Z, E, max_iter = (int(v) for v in gldpc[f"{tag}_cfg"])
    alpha = float(gldpc[f"{tag}_alpha"][0])
    eng = engines(Z)
    code = eng.encode(gldpc[f"{tag}_payload"]).cpu().numpy()
    assert np.array_equal(code, gldpc[f"{tag}_code"])
    tx = eng.rate_match(code, E).cpu().numpy()
    assert np.array_equal(tx, np.array([O.ldpc_rate_match(c, E) for c in gldpc[f"{tag}_code"]]))
    der = eng.derate_match(gldpc[f"{tag}_llr"]).cpu().numpy()
    assert np.array_equal(der, gldpc[f"{tag}_derated"])                       # float64, bit for bit
    for src in (gldpc[f"{tag}_derated"], gldpc[f"{tag}_llr"]):                 # plain and fused de-rate-matching
        r = eng.decode(src, max_iter=max_iter, alpha=alpha)
        assert np.array_equal(r["hard"].cpu().numpy(), gldpc[f"{tag}_hard"])
        assert np.array_equal(r["iters_used"].cpu().numpy(), gldpc[f"{tag}_iters"])
        assert np.array_equal(r["parity_ok"].cpu().numpy(), gldpc[f"{tag}_ok"])


def test_no_early_stop(gldpc, engines):
    r = engines(4).decode(gldpc["z4_derated"], max_iter=6, alpha=0.8, early_stop=False)
    assert np.array_equal(r["hard"].cpu().numpy(), gldpc["z4_noearly_hard"])
    assert np.array_equal(r["iters_used"].cpu().numpy(), gldpc["z4_noearly_iters"])
    assert np.array_equal(r["parity_ok"].cpu().numpy(), gldpc["z4_noearly_ok"])


@pytest.mark.parametrize("Z,E,snr_db", [(2, 12, 2.0), (4, 30, 1.0), (16, 96, 1.5), (32, 384, -1.0), (64, 500, 1.0), (128, 768, 2.0), (200, 1200, 2.0)])
def test_decode_vs_oracle_random(engines, Z, E, snr_db):
    """Seeded noisy frames (float64 LLRs): hard decisions, iteration counts and posteriors identical to the oracle.
    Z <= 64 keeps the per-thread state in shared memory; Z = 128 and 200 take the global-scratch path."""
    eng = engines(Z)
    H = O.ldpc_build_h(2, Z)
    n, k = H.shape[1], H.shape[1] - H.shape[0]
    rng = np.random.default_rng(100 + Z)
    B = 600 if Z <= 32 else 80
    payload = rng.integers(0, 2, size=(B, k), dtype=np.int8)
    code = eng.encode(payload).cpu().numpy().astype(np.int8)
    assert np.array_equal(code[:8], np.array([O.ldpc_encode(p, H) for p in payload[:8]]))
    assert not ((H.astype(np.int64) @ code.T.astype(np.int64)) % 2).any()
    nv = 1.0 / (2.0 * 10 ** (snr_db / 10.0) * k / E)
    tx = np.array([O.ldpc_rate_match(c, E) for c in code])
    llr = 2.0 * (1.0 - 2.0 * tx + rng.normal(0.0, np.sqrt(nv), size=tx.shape)) / nv
    ref = O.ldpc_decode_batch(np.array([O.ldpc_derate_match(l, n) for l in llr]), H, 20, 0.8)
    got = eng.decode(llr, max_iter=20, alpha=0.8, want_posterior=True)
    assert np.array_equal(got["hard"].cpu().numpy(), ref["hard"])
    assert np.array_equal(got["iters_used"].cpu().numpy(), ref["iters_used"])
    assert np.array_equal(got["parity_ok"].cpu().numpy().astype(bool), ref["parity_ok"])
    assert np.array_equal(got["posterior"].cpu().numpy()[:, 0] < 0, ref["hard"][:, 0].astype(bool))
    if Z <= 4:
        assert len(set(ref["iters_used"].tolist())) > 1      # small lifts converge: early exits are exercised


def test_generic_h_and_no_solution(engines):
    """A caller-supplied H (not a lifted base graph), incl. the reference's 'no solution' error path."""
    from polar_code_b200.ldpc import LdpcEngine
    rng = np.random.default_rng(5)
    H = (rng.random((10, 24)) < 0.25).astype(np.int8)
    H[:, 14:] |= np.eye(10, dtype=np.int8)                 # keep H_par full rank
    H[3, :] = 0                                             # an empty check row (decode_nms.py:27-28)
    H[3, 14 + 3] = 0
    eng = LdpcEngine(H)
    payload = rng.integers(0, 2, size=(50, 14), dtype=np.int8)
    code, status = eng.encode(payload, want_status=True)
    code, status = code.cpu().numpy().astype(np.int8), status.cpu().numpy()
    for p, c, s in zip(payload, code, status):
        try:
            ref = O.ldpc_encode(p, H)
            assert s == 0 and np.array_equal(c, ref)
        except ValueError:
            assert s == 1
    llr = 4.0 * (1.0 - 2.0 * code) + rng.normal(0, 2.0, size=code.shape)
    ref = O.ldpc_decode_batch(llr, H, 15, 0.7)
    got = eng.decode(llr, max_iter=15, alpha=0.7)
    assert np.array_equal(got["hard"].cpu().numpy(), ref["hard"])
    assert np.array_equal(got["iters_used"].cpu().numpy(), ref["iters_used"])
    Hbad = np.zeros((2, 4), np.int8)
    Hbad[0, 0] = Hbad[1, 1] = 1
    _, st = LdpcEngine(Hbad).encode(np.array([[1, 0], [0, 0]], np.uint8), want_status=True)
    assert st.cpu().tolist() == [1, 0]


@pytest.mark.parametrize("Z,kcrc,poly,E", [(2, 0, None, 12), (8, 4, "0x17", 70), (8, 0, None, 30), (32, 24, "0x1864CFB", 384)])
def test_sweep_equals_channel_plus_oracle(engines, Z, kcrc, poly, E):
    """Fused sweep counters == oracle decode of the very LLRs the Philox channel emits (payload errors, iterations),
    and the result does not depend on how the frame range is split."""
    eng = engines(Z)
    H = O.ldpc_build_h(2, Z)
    n, k = H.shape[1], H.shape[1] - H.shape[0]
    kp = k - kcrc
    eng.configure_sweep(k_crc=kcrc, E=E, max_iter=20, alpha=0.8, crc_poly=poly)
    nv = 1.0 / (2.0 * 10 ** 0.25 * kp / E)
    B = 4000
    payload, llr = eng.channel(noise_var=nv, n_frames=B, seed=11, stream_id=3)
    payload, llr = payload.cpu().numpy().astype(np.int8), llr.cpu().numpy()
    # the channel really carries the encoded message: at vanishing noise the sign pattern IS the rate-matched codeword
    p0, l0 = eng.channel(noise_var=1e-4, n_frames=64, seed=11, stream_id=3)
    assert np.array_equal(p0.cpu().numpy().astype(np.int8), payload[:64])
    msg = np.array([O.attach_crc(p, poly)[:k] if kcrc else p for p in payload[:64]])
    tx = np.array([O.ldpc_rate_match(O.ldpc_encode(m_, H), E) for m_ in msg])
    assert np.array_equal(l0.cpu().numpy() < 0, tx == 1)
    ref = O.ldpc_decode_batch(np.array([O.ldpc_derate_match(l, n) for l in llr]), H, 20, 0.8)
    be = (ref["hard"][:, :kp] != payload).sum(axis=1)
    counters = torch.zeros(16, dtype=torch.int64, device="cuda")
    fbe = torch.zeros(B, dtype=torch.int16, device="cuda")
    fw = torch.zeros(B, dtype=torch.int16, device="cuda")
    eng.sweep(counters, noise_var=nv, n_frames=B, seed=11, stream_id=3, frame_bit_errors=fbe, frame_work=fw)
    c = counters.cpu().numpy()
    assert c[0] == B and c[1] == int((be > 0).sum()) and c[2] == int(be.sum()) and c[7] == int(ref["iters_used"].sum())
    assert np.array_equal(fbe.cpu().numpy(), be) and np.array_equal(fw.cpu().numpy(), ref["iters_used"])
    assert c[1] > 0                                           # the operating point produces errors
    split = torch.zeros(16, dtype=torch.int64, device="cuda")
    for b0, nb in [(0, 1), (1, 1234), (1235, B - 1235)]:
        eng.sweep(split, noise_var=nv, n_frames=nb, frame_begin=b0, seed=11, stream_id=3)
    assert torch.equal(split, counters)


def test_mirror_functions_and_cli(gldpc, tmp_path):
    """Reference-shaped calls (tests/test_nr_ldpc.py, tests/test_ber_eval.py::test_nr_ldpc_args_parsed) on the mirror."""
    from dl_scl_polar.nr.ldpc import (load_base_graph, build_h_matrix, encode_ldpc, rate_match_ldpc, derate_match_ldpc,
                                      decode_ldpc_nms)
    from dl_scl_polar.eval import run_ber_sweep
    bg = load_base_graph(2)
    H = build_h_matrix(bg, 4)
    k = H.shape[1] - H.shape[0]
    rng = np.random.default_rng(1)
    payload = rng.integers(0, 2, size=k, dtype=np.int8)
    cw = encode_ldpc(payload, H)
    assert cw.dtype == np.int8 and not ((H @ cw) % 2).any()
    rm = rate_match_ldpc(cw, cw.size + 5)
    der = derate_match_ldpc(rm.astype(np.float64), cw.size)
    assert der.size == cw.size
    llr = 2.0 * (1.0 - 2.0 * cw + rng.normal(0.0, 0.05, size=cw.shape)) / 0.05 ** 2
    r = decode_ldpc_nms(llr, H, max_iter=10, alpha=0.9)
    assert r["parity_ok"] and np.array_equal(r["hard"][:k], payload)
    one = gldpc["z4_derated"][0]
    g = decode_ldpc_nms(one, H)
    assert np.array_equal(g["hard"], gldpc["z4_hard"][0]) and g["iters_used"] == int(gldpc["z4_iters"][0])
    with pytest.raises(ValueError):
        decode_ldpc_nms(one[:-1], H)
    with pytest.raises(ValueError):
        encode_ldpc(np.zeros(H.shape[1], np.int8), H)
    args = run_ber_sweep.parse_args(["--scheme", "nr_ldpc", "--K_payload", "6", "--K_crc", "0", "--E", "12", "--bg", "2",
                                     "--Z", "2", "--EbN0_lo", "5.0", "--EbN0_hi", "5.0", "--bits_cap", "64", "--err_cap", "2",
                                     "--out", str(tmp_path / "x.csv"), "--crc_poly", "0x1"])
    rows = run_ber_sweep.run(args)
    assert rows[0]["scheme"] == "nr_ldpc" and rows[0]["params"] == "bg=2,Z=2,iter=20,alpha=0.8"
    assert rows[0]["bits_total"] <= 66 and rows[0]["bits_total"] % 6 == 0
    bad = run_ber_sweep.parse_args(["--scheme", "nr_ldpc", "--K_payload", "64", "--K_crc", "24", "--E", "384", "--bg", "2",
                                    "--Z", "32", "--EbN0_lo", "1.0", "--EbN0_hi", "1.0", "--out", str(tmp_path / "y.csv")])
    with pytest.raises(ValueError):
        run_ber_sweep.run(bad)                                # README.md:113-118 as written: 88 != 96 (reference raises too)


def test_cli_statistics_vs_reference_rows(gldpc, tmp_path):
    """The mirror CLI (Philox channel) against the reference's own CLI rows (PCG64 channel): BER/FER within the
    binomial 4-sigma band at a sample size where that is meaningful."""
    from dl_scl_polar.eval import run_ber_sweep
    cli = json.loads(bytes(gldpc["cli_json"]).decode())["cli_b"]
    argv = list(cli["argv"])
    argv[argv.index("--bits_cap") + 1] = "4000000"
    argv[argv.index("--err_cap") + 1] = "100000000"
    argv[argv.index("--out") + 1] = str(tmp_path / "b.csv")
    rows = run_ber_sweep.run(run_ber_sweep.parse_args(argv))
    # oracle on the reference's PCG64 stream with a larger sample than the golden rows hold
    a = dict(zip(argv[0::2], argv[1::2]))
    H = O.ldpc_build_h(int(a["--bg"]), int(a["--Z"]))
    rng = np.random.default_rng(int(a["--seed"]))
    for row, snr in zip(rows, (1.0, 2.0)):
        bits, be, fe, frames, work = O.ldpc_ber_point(rng, snr, K_payload=20, K_crc=4, crc_poly="0x17", H=H, E=70,
                                                      max_iter=12, alpha=0.75, err_cap=10 ** 9, bits_cap=400000)
        p_ref, n_ref = fe / frames, frames
        p, n_ = row["fer"], row["bits_total"] // 20
        sd = np.sqrt(p_ref * (1 - p_ref) * (1 / n_ref + 1 / n_))
        assert abs(p - p_ref) < 4 * sd + 1e-9, (snr, p, p_ref)
        assert abs(row["avg_work"] - work / frames) < 0.1 * (work / frames)
