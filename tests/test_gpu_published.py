"""GPU: the reference's PUBLISHED result files reproduced through the CUDA decoders, and the large parity campaign.

results/fer_M{1,4,8}.csv of the reference (tests/golden/published.json) are byte-reproducible from its PCG64 channel
(oracle.fer_sweep_frames restates run_fer_sweep.py:60-121).  Here the very LLRs of those runs are fed to the CUDA
SCL / DL-SCL decoders (fp32 rows) and the CSV row is rebuilt from the GPU decisions: it must equal the published row
byte for byte; a frame may differ from the float64 oracle only if the kernel flagged it (near-tie / rank-tie).

The campaign (round 1: scripts/tie_stats.py, a builder-run text file) is a test now: 200 000 Philox frames per case.
"""
import numpy as np
import pytest
import torch

from conftest import CRC24
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from polar_code_b200.engine import PolarEngine, construct_info_set
    return PolarEngine(128, construct_info_set(128, 64), CRC24)


def _nv(snr):
    return 1.0 / (2.0 * 0.5 * 10 ** (snr / 10))


@pytest.mark.parametrize("name", ["fer_M4", "fer_M8", "fer_M1"])
def test_published_csv_through_cuda(eng, g128, published, name):
    rec, rows = published["recipe"][name], published["rows"][name]
    M, frames = rec["M"], rec["frames"]
    A = O.construct_info_set(128, 64)
    beta = g128[f"beta_M{M}"]
    for snr, want in zip(rec["snr"], rows[1:]):
        msgs, llrs, unc = O.fer_sweep_frames(snr, frames, seed=0, include_uncoded=True)
        x = llrs.astype(np.float32)
        s = eng.scl_decode(x, M, want=("best_bits", "crc_ok", "flags"))
        d = eng.dlscl_decode(x, M, 8, beta=beta)
        sb, db = s["best_bits"].cpu().numpy().astype(np.int8), d["best_bits"].cpu().numpy().astype(np.int8)
        # run_fer_sweep.py:91-94,100-103: frame error = CRC failure of the returned word; bit errors over all K bits
        s_fe = int((s["crc_ok"] == 0).sum().item())
        d_fe = int((d["success"] == 0).sum().item())
        tb = frames * 64
        row = ",".join([f"{snr:.3f}", f"{(unc > 0).sum() / frames:.6e}", f"{unc.sum() / (frames * 40):.6e}",
                        f"{s_fe / frames:.6e}", f"{int((sb != msgs).sum()) / tb:.6e}",
                        f"{d_fe / frames:.6e}", f"{int((db != msgs).sum()) / tb:.6e}"])
        # frame-level parity with the float64 oracle on the reference's own (float64) LLRs
        so = O.scl_decode_batch(llrs, A, M, crc=CRC24, want_info_llrs=False)
        do = O.dlscl_decode_batch(llrs, A, M, 8, crc=CRC24, beta=beta)
        s_diff = ~(sb == so["best_bits"]).all(axis=1)
        d_diff = ~((db == do["best_bits"]).all(axis=1) & (d["n_attempts"].cpu().numpy() == do["n_attempts"]))
        s_flag = (s["flags"].cpu().numpy() & 1) != 0
        d_flag = (d["flags"].cpu().numpy() & 3) != 0
        assert not (s_diff & ~s_flag).any() and not (d_diff & ~d_flag).any(), "unflagged frame differs from the oracle"
        assert s_diff.sum() + d_diff.sum() <= 2
        if not s_diff.any() and not d_diff.any():
            assert row == want, f"{name} @ {snr} dB: CUDA row differs from the published row"
        # crc_ok / success are the CRC check of the returned word
        assert s_fe == sum(not O.check_crc(b, CRC24) for b in sb) and d_fe == sum(not O.check_crc(b, CRC24) for b in db)


CAMPAIGN = [(1, 3.0), (2, 3.0), (4, 2.0), (4, 4.0), (4, 5.0), (8, 3.0), (8, 5.0)]


@pytest.mark.parametrize("M,snr", CAMPAIGN)
def test_scl_parity_campaign(eng, M, snr):
    """200 000 frames per case: all M candidates, their order, the best index; metrics within the north-star tolerance."""
    B = 200_000
    A = O.construct_info_set(128, 64)
    _, llr = eng.channel(noise_var=_nv(snr), n_frames=B, seed=77 + M, stream_id=int(snr * 10), k_payload=40)
    ref = O.scl_decode_batch(llr.cpu().numpy().astype(np.float64), A, M, crc=CRC24, want_info_llrs=False)
    out = eng.scl_decode(llr, M, want=("cand", "metrics", "best_idx", "best_bits", "flags", "n_cand"))
    cand = out["cand"].cpu().numpy().astype(np.int8)
    flagged = (out["flags"].cpu().numpy() & 1) != 0
    diff = ~((cand == ref["cand"]).all(axis=(1, 2)) & (out["best_idx"].cpu().numpy() == ref["best_idx"]))
    assert not (diff & ~flagged).any(), f"{int((diff & ~flagged).sum())} unflagged frames differ from the float64 oracle"
    assert diff.sum() <= 20 and flagged.mean() < 0.02
    m = out["metrics"].cpu().numpy()
    fin = np.isfinite(ref["metrics"]) & ~diff[:, None]
    rel = np.abs(m[fin] - ref["metrics"][fin]) / np.maximum(np.abs(ref["metrics"][fin]), 1e-30)
    assert rel.max() < 1e-4          # BASELINE.json north_star: fp32 path metrics within 1e-4 relative of float64


@pytest.mark.parametrize("M,snr", [(4, 4.0), (8, 4.5)])
def test_dlscl_parity_campaign(eng, g128, M, snr):
    """50 000 frames, 8 retries with the SHIPPED beta: returned word and number of attempts."""
    B = 50_000
    A = O.construct_info_set(128, 64)
    beta = g128[f"beta_M{M}"]
    _, llr = eng.channel(noise_var=_nv(snr), n_frames=B, seed=177 + M, stream_id=int(snr * 10), k_payload=40)
    ref = O.dlscl_decode_batch(llr.cpu().numpy().astype(np.float64), A, M, 8, crc=CRC24, beta=beta)
    out = eng.dlscl_decode(llr, M, 8, beta=beta)
    same = (out["best_bits"].cpu().numpy().astype(np.int8) == ref["best_bits"]).all(axis=1)
    same &= out["n_attempts"].cpu().numpy() == ref["n_attempts"]
    flagged = (out["flags"].cpu().numpy() & 3) != 0
    assert (ref["n_attempts"] > 1).sum() > 500
    assert not (~same & ~flagged).any(), f"{int((~same & ~flagged).sum())} unflagged DL-SCL frames differ from the oracle"
    assert (~same).sum() <= 10
