"""GPU: the drop-in `dl_scl_polar` mirror, exercised the way the reference's own tests exercise the original
(reference tests/test_polar_basics.py, test_scl_crc.py, test_flip_logic.py, test_nr_polar.py, test_ber_eval.py,
test_cli_end2end.py) plus per-call differential checks against the oracle."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _mods():
    from dl_scl_polar import config
    from dl_scl_polar.utils.seeding import seed_all
    from dl_scl_polar.polar.polar import construct_info_set, encode, sc_decode
    from dl_scl_polar.polar.crc import attach_crc, check_crc
    from dl_scl_polar.polar.scl import decode_scl
    from dl_scl_polar.dlscl import flip
    return config, seed_all, construct_info_set, encode, sc_decode, attach_crc, check_crc, decode_scl, flip


def _awgn_llr(code, ebno_db, seed, rate=0.5):
    from dl_scl_polar.utils.seeding import seed_all
    seed_all(seed)
    nv = 1 / (2 * rate * 10 ** (ebno_db / 10))
    return 2 * ((1.0 - 2.0 * code) + np.random.normal(0.0, np.sqrt(nv), size=code.shape)) / nv


def test_encode_shape_and_noiseless_sc_roundtrip():
    config, seed_all, cis, encode, sc_decode, attach_crc, check_crc, decode_scl, flip = _mods()
    cfg = config.DEFAULTS
    A = cis(cfg.N, cfg.K)
    seed_all(0)
    msg = np.random.randint(0, 2, size=cfg.K, dtype=np.int8)
    code = encode(msg)
    assert code.shape == (cfg.N,) and code.dtype == np.int8
    assert np.array_equal(code, O.encode(msg, A, cfg.N))
    llr = (1.0 - 2.0 * code) * 1e6
    assert np.array_equal(sc_decode(llr, A), msg)
    noisy = _awgn_llr(code, 6.0, 123)
    assert np.array_equal(sc_decode(noisy, A), O.sc_decode(noisy.astype(np.float32).astype(np.float64), A))


def test_crc_attach_check_and_single_bit_corruption():
    config, seed_all, cis, encode, sc_decode, attach_crc, check_crc, decode_scl, flip = _mods()
    poly = config.DEFAULTS.crc_poly
    seed_all(1)
    payload = np.random.randint(0, 2, size=40, dtype=np.int8)
    word = attach_crc(payload, poly)
    assert word.size == 64 and np.array_equal(word, O.attach_crc(payload, poly)) and check_crc(word, poly)
    for pos in (0, 17, 63):
        bad = word.copy(); bad[pos] ^= 1
        assert not check_crc(bad, poly)
    assert np.array_equal(attach_crc(np.array([1, 0, 1, 1, 0, 0, 1, 0], np.int8), "0x17"),
                          O.attach_crc(np.array([1, 0, 1, 1, 0, 0, 1, 0], np.int8), "0x17"))


def test_scl_recovers_frames_sc_loses():
    """Within a few hundred seeded trials at 2 dB there is a frame where SC fails the CRC but SCL M=4 returns the
    exact word (reference tests/test_scl_crc.py:37-70)."""
    config, seed_all, cis, encode, sc_decode, attach_crc, check_crc, decode_scl, flip = _mods()
    cfg = config.DEFAULTS
    A = cis(cfg.N, cfg.K)
    found = False
    for t in range(300):
        seed_all(500 + t)
        info = attach_crc(np.random.randint(0, 2, size=40, dtype=np.int8), cfg.crc_poly)
        llr = _awgn_llr(encode(info), 2.0, 900 + t)
        if not check_crc(sc_decode(llr, A), cfg.crc_poly):
            r = decode_scl(llr, A, M=4, crc=cfg.crc_poly)
            if np.array_equal(r["best_path_bits"], info):
                found = True
                break
    assert found


def test_decode_scl_dict_matches_oracle():
    config, seed_all, cis, encode, sc_decode, attach_crc, check_crc, decode_scl, flip = _mods()
    cfg = config.DEFAULTS
    A = cis(cfg.N, cfg.K)
    seed_all(3)
    info = attach_crc(np.random.randint(0, 2, size=40, dtype=np.int8), cfg.crc_poly)
    llr = _awgn_llr(encode(info), 3.0, 4).astype(np.float32).astype(np.float64)
    for M in (1, 2, 4, 8):
        r = decode_scl(llr, A, M, crc=cfg.crc_poly)
        o = O.scl_decode_batch(llr, A, M, crc=cfg.crc_poly)
        n = int(o["n_cand"][0])
        assert set(r) == {"candidates", "metrics", "best_path_bits", "info_llrs", "best_path_info_llrs"}
        assert len(r["candidates"]) == n and all(c.dtype == np.int8 for c in r["candidates"])
        assert np.array_equal(np.array(r["candidates"]), o["cand"][0, :n])
        np.testing.assert_allclose(r["metrics"], o["metrics"][0, :n], rtol=1e-4)
        np.testing.assert_allclose(np.array(r["info_llrs"]), o["info_llrs"][0, :n], rtol=1e-4, atol=2e-4)
        assert any(r["best_path_bits"] is c for c in r["candidates"])     # aliasing as in scl.py:199-201
        assert np.array_equal(r["best_path_bits"], o["best_bits"][0])
    r = decode_scl(llr, A, 4, crc=None)
    assert r["best_path_bits"] is r["candidates"][0]


def test_flip_logic():
    config, seed_all, cis, encode, sc_decode, attach_crc, check_crc, decode_scl, flip = _mods()
    cfg = config.DEFAULTS
    A = cis(cfg.N, cfg.K)
    assert flip.choose_flip_index(np.array([0.8, 0.3, 1.5, 0.2]), beta=None) == 3
    seed_all(0)
    info = attach_crc(np.random.randint(0, 2, size=40, dtype=np.int8), cfg.crc_poly)
    llr = _awgn_llr(encode(info), 6.0, 10)
    base = decode_scl(llr, A, M=4, crc=cfg.crc_poly)
    best = base["best_path_bits"]
    assert check_crc(best, cfg.crc_poly)
    res = flip.retry_with_flip(llr, A, M=4, best_path_bits=best, flip_index=5, crc=cfg.crc_poly)
    forced = res["forced_info_bits"]
    assert forced[5] == 1 - best[5] and np.array_equal(forced[:5], best[:5]) and res["flip_index"] == 5
    for cand in res["candidates"]:
        assert np.array_equal(cand[:5], best[:5]) and cand[5] == forced[5]
    # retries=0 is the baseline
    llr2 = _awgn_llr(encode(info), 2.0, 200)
    r0 = flip.decode_with_retries(llr2, A, M=4, retries=0, crc=cfg.crc_poly)
    assert np.array_equal(r0["best_path_bits"], decode_scl(llr2, A, M=4, crc=cfg.crc_poly)["best_path_bits"])
    assert r0["tried_indices"] == [] and r0["attempts"][0]["attempt_type"] == "baseline"


def test_retries_recover_a_crc_failure_and_match_oracle():
    config, seed_all, cis, encode, sc_decode, attach_crc, check_crc, decode_scl, flip = _mods()
    cfg = config.DEFAULTS
    A = cis(cfg.N, cfg.K)
    recovered = False
    for t in range(1, 201):
        seed_all(1000 + t)
        info = attach_crc(np.random.randint(0, 2, size=40, dtype=np.int8), cfg.crc_poly)
        llr = _awgn_llr(encode(info), 1.0, 2000 + t).astype(np.float32).astype(np.float64)
        base = decode_scl(llr, A, M=2, crc=cfg.crc_poly)
        if check_crc(base["best_path_bits"], cfg.crc_poly):
            continue
        res = flip.decode_with_retries(llr, A, M=2, retries=4, crc=cfg.crc_poly)       # beta=None: |L0| ranking
        o = O.dlscl_decode_batch(llr, A, 2, 4, crc=cfg.crc_poly)
        if o["min_gap"][0] > 1e-5 and o["min_rank_gap"][0] > 1e-5:
            assert [int(i) for i in res["tried_indices"]] == [int(i) for i in o["tried"][0] if i >= 0]
            assert res["success"] == bool(o["success"][0]) and len(res["attempts"]) == o["n_attempts"][0]
            assert np.array_equal(res["best_path_bits"], o["best_bits"][0])
        assert all(a["attempt_type"] == "flip" for a in res["attempts"][1:])
        if res["success"] and np.array_equal(res["best_path_bits"], info):
            recovered = True
            break
    assert recovered


def test_nr_polar_roundtrip():
    from dl_scl_polar.nr.polar import encode_rate_matched, decode_rate_matched_scl
    from dl_scl_polar.polar.polar import construct_info_set
    poly = "0x1864CFB"
    A = construct_info_set(128, 88)
    rng = np.random.default_rng(5)
    payload = rng.integers(0, 2, 64, dtype=np.int8)
    for E in (128, 256):
        tx = encode_rate_matched(payload, poly, 128, E, A)
        assert tx.size == E
        r = decode_rate_matched_scl((1.0 - 2.0 * tx) * 20.0, poly, 128, E, A, 4)
        assert r["crc_pass"] and np.array_equal(r["payload"][:64], payload) and np.array_equal(r["best_path_bits"][:64], payload)
        y = 1.0 - 2.0 * tx + rng.normal(0, 0.3, E)
        r = decode_rate_matched_scl(2 * y / 0.09, poly, 128, E, A, 4)
        assert r["crc_pass"]


def test_ber_sweep_small_configs(tmp_path):
    """reference tests/test_ber_eval.py:19-89 geometries: N=16, K=8+4, poly 0x17, M=2."""
    from dl_scl_polar.eval import run_ber_sweep as R
    common = ["--K_payload", "8", "--K_crc", "4", "--E", "16", "--crc_poly", "0x17", "--M", "2", "--EbN0_step", "0.5",
              "--bits_cap", "64", "--err_cap", "2"]
    rows = R.run(R.parse_args(["--scheme", "polar_scl", "--EbN0_lo", "6.0", "--EbN0_hi", "6.0", "--out", str(tmp_path / "a.csv")] + common))
    assert len(rows) == 1
    row = rows[0]
    assert row["scheme"] == "polar_scl" and row["K_payload"] == 8 and row["rate"] == pytest.approx(0.5)
    assert row["bits_total"] > 0 and row["ber"] >= 0.0 and row["avg_work"] == 0.0
    assert row["bits_total"] <= 64 and row["bits_total"] % 8 == 0
    rows = R.run(R.parse_args(["--scheme", "nr_polar_scl", "--N", "16", "--EbN0_lo", "5.0", "--EbN0_hi", "5.0",
                               "--out", str(tmp_path / "b.csv")] + common))
    assert rows[0]["scheme"] == "nr_polar_scl"
    R.main(["--scheme", "polar_scl", "--EbN0_lo", "4.0", "--EbN0_hi", "5.0", "--out", str(tmp_path / "c.csv"),
            "--plot", str(tmp_path / "c.png")] + common)
    txt = (tmp_path / "c.csv").read_text().splitlines()
    assert txt[0] == ",".join(R.HEADER) and len(txt) == 4 and (tmp_path / "c.png").exists()
    rows = R.run(R.parse_args(["--scheme", "nr_ldpc", "--K_payload", "6", "--K_crc", "0", "--E", "12", "--EbN0_lo", "5",
                               "--EbN0_hi", "5", "--bits_cap", "600", "--out", str(tmp_path / "d.csv")]))
    assert rows[0]["scheme"] == "nr_ldpc" and rows[0]["bits_total"] % 6 == 0 and rows[0]["avg_work"] >= 1.0


def test_ber_sweep_stop_rule_is_the_sequential_loop(tmp_path):
    """bits_total / bit_errors of the batched run equal the frame-by-frame loop over the same Philox frames."""
    import torch
    from dl_scl_polar.eval import run_ber_sweep as R
    from dl_scl_polar._engines import engine_for
    from dl_scl_polar.polar.polar import construct_info_set
    from polar_code_b200 import montecarlo as mc
    args = R.parse_args(["--scheme", "polar_scl", "--K_payload", "64", "--K_crc", "24", "--E", "128", "--M", "4",
                         "--EbN0_lo", "2.0", "--EbN0_hi", "2.0", "--bits_cap", "4e6", "--err_cap", "300", "--out", str(tmp_path / "x.csv")])
    row = R.run(args)[0]
    eng = engine_for(128, construct_info_set(128, 88), "0x1864CFB")
    n = 1 << 16
    err = torch.zeros(n, dtype=torch.int16, device=eng.dev)
    cnt = torch.zeros(16, dtype=torch.int64, device=eng.dev)
    eng.sweep(cnt, M=4, noise_var=mc.ber_noise_var(2.0, 64, 128), n_frames=n, seed=0, stream_id=0, k_payload=64,
              frame_error_mode=1, bit_error_span=64, frame_bit_errors=err)
    e = err.cpu().numpy().astype(np.int64)
    tot = frames = 0
    while tot < 300 and frames * 64 < 4e6:
        tot += e[frames]; frames += 1
    assert (row["bits_total"], row["bit_errors"]) == (frames * 64, tot)
    assert row["fer"] == pytest.approx((e[:frames] > 0).sum() / frames)


def test_fer_cli_end_to_end(tmp_path):
    """CSV header exactly as the reference (tests/test_cli_end2end.py:60-88); files are produced."""
    from dl_scl_polar.eval import run_fer_sweep as F
    beta = np.eye(64, dtype=np.float32)
    np.save(tmp_path / "beta.npy", beta)
    F.main(["--M", "2", "--frames", "2000", "--snr_lo", "4.0", "--snr_hi", "5.0", "--snr_step", "1.0", "--retries", "2",
            "--beta", str(tmp_path / "beta.npy"), "--seed", "3", "--out_dir", str(tmp_path / "r"), "--plot_dir", str(tmp_path / "p")])
    lines = (tmp_path / "r" / "fer_M2.csv").read_text().splitlines()
    assert lines[0] == "snr_db,fer_scl,ber_scl,fer_dl,ber_dl" and len(lines) == 3 and lines[1].startswith("4.000,")
    assert (tmp_path / "p" / "fer_M2.png").exists()
    F.main(["--M", "2", "--frames", "500", "--snr_lo", "5.0", "--snr_step", "0", "--include_uncoded",
            "--out_dir", str(tmp_path / "r2"), "--plot_dir", str(tmp_path / "p2")])
    lines = (tmp_path / "r2" / "fer_M2.csv").read_text().splitlines()
    assert lines[0] == "snr_db,fer_uncoded,ber_uncoded,fer_scl,ber_scl,fer_dl,ber_dl" and len(lines) == 2
    vals = [float(v) for v in lines[1].split(",")]
    assert vals[0] == 5.0 and 0.1 < vals[1] < 0.35 and vals[5] <= vals[3]     # uncoded FER ~0.22; DL-SCL never worse than SCL on CRC-FER


def test_make_dataset_labels_match_oracle(tmp_path):
    """batched dataset generation; labels / |L0| rows equal the frame-by-frame rule
    (reference train/make_dataset.py:47-91) evaluated with the float64 oracle on the same LLRs."""
    import json
    import torch
    from dl_scl_polar.train import make_dataset as D
    from dl_scl_polar._engines import engine_for
    from dl_scl_polar.polar.polar import construct_info_set
    from polar_code_b200 import montecarlo as mc
    A = construct_info_set(128, 64)
    eng = engine_for(128, A, "0x1864CFB")
    n, M = 20000, 2
    msg, llr = eng.channel(noise_var=mc.fer_noise_var(4.0, 64, 128), n_frames=n, seed=5, stream_id=0, k_payload=40)
    llr = llr * (1.0 - 2.0 * eng.encode(msg).to(torch.float32))
    fail, abs_l0, label = D.label_failures(eng, llr, M, torch.zeros(64, dtype=torch.uint8, device=eng.dev))
    fail, abs_l0, label = fail.cpu().numpy(), abs_l0.cpu().numpy(), label.cpu().numpy()
    l64 = llr.cpu().numpy().astype(np.float64)
    base = O.scl_decode_batch(l64, A, M, crc="0x1864CFB")
    ofail = np.array([b for b in range(n) if not O.check_crc(base["best_bits"][b], "0x1864CFB")])
    gaps_ok = base["min_gap"] > 1e-5
    assert set(ofail[gaps_ok[ofail]]) <= set(fail) and len(fail) > 200
    checked = 0
    for row, b in enumerate(fail[:400]):
        if not gaps_ok[b] or b not in set(ofail):
            continue
        ol0 = np.abs(base["info_llrs"][b, base["best_idx"][b]]).astype(np.float32)
        np.testing.assert_allclose(abs_l0[row], ol0, rtol=1e-4, atol=2e-4)
        srt = np.sort(ol0)
        if np.min(np.diff(srt[:9])) < 1e-4:
            continue                                        # |L0| order itself is a near tie
        want = -1
        for idx in np.argsort(ol0)[:8]:
            force = np.full(64, -1, np.int8); force[:idx] = base["best_bits"][b][:idx]; force[idx] = 1 - base["best_bits"][b][idx]
            r = O.scl_decode_batch(l64[b], A, M, crc="0x1864CFB", force=force)
            if r["min_gap"][0] < 1e-5:
                want = None
                break
            if O.check_crc(r["best_bits"][0], "0x1864CFB") and not r["best_bits"][0].any():
                want = int(idx)
                break
        if want is not None:
            assert label[row] == want, (b, label[row], want)
            checked += 1
    assert checked > 100
    D.main(["--M", "2", "--snr_db", "4.0", "--frames", "30000", "--seed", "1", "--out", str(tmp_path / "train_M2")])
    z = np.load(tmp_path / "train_M2_part0.npz")
    meta = json.loads(str(z["meta"]))
    assert z["abs_l0"].dtype == np.float32 and z["abs_l0"].shape[1] == 64 and z["flip_idx"].dtype == np.int32
    assert z["abs_l0"].shape[0] == z["flip_idx"].shape[0] == meta["samples"] > 0
    assert set(meta) == {"M", "EbN0_dB", "seed", "frames", "crc_poly", "crc_bits", "samples", "failures"}
    assert ((z["flip_idx"] >= 0) & (z["flip_idx"] < 64)).all()
    # the label is always among the 8 smallest |L0| positions of its row
    rank = (z["abs_l0"] < z["abs_l0"][np.arange(len(z["flip_idx"])), z["flip_idx"]][:, None]).sum(axis=1)
    assert (rank < 8).all()


def test_dataset_train_sweep_loop(tmp_path):
    """generate_samples -> train_beta -> run_sweep end to end (reference tests/test_cli_end2end.py:11-58)."""
    from dl_scl_polar.train import make_dataset as D, train_beta as T
    from dl_scl_polar.eval import run_fer_sweep as F
    D.main(["--M", "2", "--snr_db", "4.0", "--frames", "40000", "--seed", "0", "--out", str(tmp_path / "data" / "train_M2")])
    T.main(["--M", "2", "--data", str(tmp_path / "data" / "train_M2_part*.npz"), "--epochs", "2", "--batch", "256",
            "--checkpoint_dir", str(tmp_path / "ck"), "--log_dir", str(tmp_path / "lg")])
    beta = np.load(tmp_path / "ck" / "beta_M2.npy")
    assert beta.shape == (64, 64) and np.allclose(beta, beta.T) and np.allclose(np.diag(beta), 1.0)
    F.main(["--M", "2", "--frames", "20000", "--snr_lo", "4.0", "--snr_hi", "4.5", "--retries", "8", "--beta",
            str(tmp_path / "ck" / "beta_M2.npy"), "--out_dir", str(tmp_path / "res"), "--plot_dir", str(tmp_path / "plt")])
    lines = (tmp_path / "res" / "fer_M2.csv").read_text().splitlines()
    assert lines[0] == "snr_db,fer_scl,ber_scl,fer_dl,ber_dl" and len(lines) == 3
    for ln in lines[1:]:
        v = [float(x) for x in ln.split(",")]
        assert 0 < v[3] <= v[1] < 0.5                         # DL-SCL CRC-FER never above SCL's
    assert (tmp_path / "plt" / "fer_M2.png").exists()
