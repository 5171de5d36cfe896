"""GPU: `run_ber_sweep --scheme dl_scl` (reference eval/run_ber_sweep.py:184-190,228-293) with a K_total = 88 beta.

The reference's README command for this scheme fails with the shipped 64 x 64 beta (matmul 88 vs 64);
the scheme itself is generic, so it is tested here with an 88 x 88 matrix: (1) the CLI rows equal a frame-by-frame
evaluation of the reference's loop over the SAME Philox frames decoded by the batched DL-SCL entry point, (2) that
batched decode equals the float64 oracle on those frames (flagged near-ties excepted), (3) an 88 x 88 beta can be
produced on-device: make_dataset --N 128 --K_total 88 -> train_beta -> run_ber_sweep --scheme dl_scl end to end.
"""
import json

import numpy as np
import pytest
import torch

from conftest import CRC24
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _beta88(seed=3):
    rng = np.random.default_rng(seed)
    b = 0.05 * rng.standard_normal((88, 88))
    return (np.eye(88) + (b + b.T) / 2).astype(np.float32)


def _sequential(err, work, payload_len, err_cap, bits_cap):
    """run_ber_sweep.py:127 + SimulationStats :36-62"""
    bits = errs = fe = frames = 0
    wsum = 0.0
    while errs < err_cap and bits < bits_cap:
        bits += payload_len; errs += int(err[frames]); fe += int(err[frames] > 0); wsum += float(work[frames]); frames += 1
    return dict(bits_total=bits, bit_errors=errs, ber=errs / bits, fer=fe / frames, avg_work=wsum / frames)


@pytest.mark.parametrize("err_cap,bits_cap", [(10**9, 64 * 3000), (400, 1e7)])
def test_run_ber_sweep_dl_scl_k88(tmp_path, err_cap, bits_cap):
    from dl_scl_polar.eval import run_ber_sweep as R
    from dl_scl_polar._engines import engine_for
    from dl_scl_polar.polar.polar import construct_info_set
    from polar_code_b200 import montecarlo as mc
    beta = _beta88()
    np.save(tmp_path / "beta88.npy", beta)
    M, retries, seed = 4, 4, 5
    args = R.parse_args(["--scheme", "dl_scl", "--K_payload", "64", "--K_crc", "24", "--E", "128", "--M", str(M),
                         "--retries", str(retries), "--beta", str(tmp_path / "beta88.npy"), "--EbN0_lo", "2.5", "--EbN0_hi", "3.0",
                         "--EbN0_step", "0.5", "--bits_cap", str(bits_cap), "--err_cap", str(err_cap), "--seed", str(seed),
                         "--out", str(tmp_path / "dl.csv")])
    rows = R.run(args)
    assert [r["EbN0_dB"] for r in rows] == [2.5, 3.0]
    A = construct_info_set(128, 88)
    eng = engine_for(128, A, CRC24)
    n = 1 << 15
    for point, row in enumerate(rows):
        assert row["scheme"] == "dl_scl" and row["params"] == f"M={M},retries={retries}" and row["N_or_E"] == 128
        nv = mc.ber_noise_var(row["EbN0_dB"], 64, 128)
        msg, llr = eng.channel(noise_var=nv, n_frames=n, seed=seed, stream_id=point, k_payload=64)
        out = eng.dlscl_decode(llr, M, retries, beta=beta)
        be = (out["best_bits"][:, :64] != msg[:, :64]).sum(dim=1).cpu().numpy()
        work = out["n_attempts"].cpu().numpy() - 1
        want = _sequential(be, work, 64, err_cap, bits_cap)
        # (1) the fused sweep behind the CLI == the batched API on the same frames, cut by the sequential rule
        assert (row["bits_total"], row["bit_errors"]) == (want["bits_total"], want["bit_errors"])
        assert row["ber"] == pytest.approx(want["ber"]) and row["fer"] == pytest.approx(want["fer"])
        assert row["avg_work"] == pytest.approx(want["avg_work"]) and row["avg_work"] > 0
        # (2) the batched API == the float64 oracle (decode_with_retries, flip.py:65-141) on those frames
        take = want["bits_total"] // 64
        ref = O.dlscl_decode_batch(llr[:take].cpu().numpy().astype(np.float64), A, M, retries, crc=CRC24, beta=beta)
        same = (out["best_bits"][:take].cpu().numpy().astype(np.int8) == ref["best_bits"]).all(axis=1)
        same &= out["n_attempts"][:take].cpu().numpy() == ref["n_attempts"]
        flagged = (out["flags"][:take].cpu().numpy() & 3) != 0
        assert not (~same & ~flagged).any() and (~same).sum() <= 2
    # the mismatched-size beta is refused up front (the reference dies at its first retry with the same ValueError type)
    np.save(tmp_path / "beta64.npy", np.eye(64, dtype=np.float32))
    with pytest.raises(ValueError):
        R.run(R.parse_args(["--scheme", "dl_scl", "--K_payload", "64", "--K_crc", "24", "--E", "128", "--M", "4", "--beta",
                            str(tmp_path / "beta64.npy"), "--EbN0_lo", "3", "--EbN0_hi", "3", "--out", str(tmp_path / "x.csv")]))


def test_dataset_train_sweep_k88(tmp_path):
    """an 88 x 88 beta trained on-device and used by run_ber_sweep --scheme dl_scl."""
    from dl_scl_polar.train import make_dataset as D, train_beta as T
    from dl_scl_polar.eval import run_ber_sweep as R
    D.main(["--M", "2", "--snr_db", "3.0", "--frames", "60000", "--seed", "0", "--N", "128", "--K_total", "88",
            "--out", str(tmp_path / "data" / "train88")])
    z = np.load(tmp_path / "data" / "train88_part0.npz")
    meta = json.loads(str(z["meta"]))
    assert z["abs_l0"].shape[1] == 88 and z["abs_l0"].shape[0] == meta["samples"] > 100
    assert meta["N"] == 128 and meta["K_total"] == 88 and ((z["flip_idx"] >= 0) & (z["flip_idx"] < 88)).all()
    T.main(["--M", "2", "--data", str(tmp_path / "data" / "train88_part*.npz"), "--epochs", "2", "--batch", "256",
            "--checkpoint_dir", str(tmp_path / "ck"), "--log_dir", str(tmp_path / "lg")])
    beta = np.load(tmp_path / "ck" / "beta_M2.npy")
    assert beta.shape == (88, 88) and beta.dtype == np.float32 and np.allclose(beta, beta.T) and np.allclose(np.diag(beta), 1.0)
    common = ["--K_payload", "64", "--K_crc", "24", "--E", "128", "--M", "2", "--EbN0_lo", "3.0", "--EbN0_hi", "3.0",
              "--bits_cap", str(64 * 40000), "--err_cap", str(10**9), "--seed", "1"]
    dl = R.run(R.parse_args(["--scheme", "dl_scl", "--retries", "8", "--beta", str(tmp_path / "ck" / "beta_M2.npy"),
                             "--out", str(tmp_path / "dl.csv")] + common))[0]
    scl = R.run(R.parse_args(["--scheme", "polar_scl", "--out", str(tmp_path / "scl.csv")] + common))[0]
    assert dl["bits_total"] == scl["bits_total"] == 64 * 40000
    assert 0 < dl["fer"] < scl["fer"] and dl["avg_work"] > 0 and scl["avg_work"] == 0.0     # retries recover frames
    R.write_csv([scl, dl], tmp_path / "both.csv")
    assert (tmp_path / "both.csv").read_text().splitlines()[0] == ",".join(R.HEADER)
