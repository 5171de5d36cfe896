"""GPU: the binned, prefix-skipping DL-SCL retry kernel (csrc/polar_sweep.cuh dl_bin_kernel) against the
frame-per-group retry kernel, which re-decodes every attempt from phase 0 as dlscl/flip.py:37-62 literally does.

A retry that jump-starts at the flipped phase must give the very result of the full re-decode: same word, same number
of attempts, same flip sequence.  The only licence is the near-tie window (metrics restart at the jump, so WHICH frames
are flagged PB200_FLAG_NEAR_TIE may differ): a frame may differ only if one of the two kernels flagged it.
The oracle comparisons of the binned kernel itself are in test_gpu_parity.py / test_gpu_generic.py (it is the default).
"""
import os

import numpy as np
import pytest
import torch

from conftest import CRC24

pytestmark = pytest.mark.gpu


def _nv(snr_db, rate=0.5):
    return 1.0 / (2.0 * rate * 10 ** (snr_db / 10.0))


def _engine(binned: bool):
    from polar_code_b200.engine import PolarEngine, construct_info_set
    old = os.environ.get("PB200_DL_BINNED")
    os.environ["PB200_DL_BINNED"] = "1" if binned else "0"
    try:
        eng = PolarEngine(128, construct_info_set(128, 64), CRC24)
        # the mode is latched by the engine's first DL-SCL launch
        _, llr = eng.channel(noise_var=_nv(3.0), n_frames=64, seed=1, stream_id=9, k_payload=40)
        eng.dlscl_decode(llr, 4, 1)
    finally:
        if old is None:
            os.environ.pop("PB200_DL_BINNED", None)
        else:
            os.environ["PB200_DL_BINNED"] = old
    return eng


@pytest.fixture(scope="module")
def engines():
    return _engine(True), _engine(False)


def _np(out):
    return {k: v.cpu().numpy() for k, v in out.items()}


def _same(a, b):
    return ((a["best_bits"] == b["best_bits"]).all(axis=1) & (a["success"] == b["success"])
            & (a["n_attempts"] == b["n_attempts"]) & (a["tried"] == b["tried"]).all(axis=1))


@pytest.mark.parametrize("M,retries,use_beta,snr", [(4, 8, True, 3.5), (8, 8, True, 4.0), (4, 8, False, 4.0), (2, 5, True, 3.0),
                                                    (1, 8, False, 4.5), (4, 64, True, 2.0)])
def test_binned_equals_full_redecode(engines, g128, M, retries, use_beta, snr):
    eb, ef = engines
    n = 30000
    _, llr = eb.channel(noise_var=_nv(snr), n_frames=n, seed=21 + M, stream_id=5, k_payload=40)
    beta = g128[f"beta_M{M}"] if (use_beta and f"beta_M{M}" in g128) else (g128["beta_M4"] if use_beta else None)
    a, b = _np(eb.dlscl_decode(llr, M, retries, beta=beta)), _np(ef.dlscl_decode(llr, M, retries, beta=beta))
    same = _same(a, b)
    flagged = ((a["flags"] | b["flags"]) & 3) != 0
    assert not (~same & ~flagged).any(), "binned and full re-decode differ on a frame neither flagged"
    assert (~same).sum() <= 3
    assert (a["n_attempts"] > 1).sum() > 100          # the case really retries
    # the two kernels agree on the rank-tie flag exactly (same scores); the near-tie window is relative to the metric, which
    # restarts at the jump, so the set of near-tie-flagged frames differs a little (it stays a fraction of a percent)
    assert ((a["flags"] ^ b["flags"]) & 2).sum() <= (~same).sum()
    if retries <= 8:
        assert ((a["flags"] ^ b["flags"]) & 1).sum() <= n // 100


def test_result_does_not_depend_on_batching(engines, g128):
    """Frames migrate between warps and share a warp with whatever the rings hold, so two runs batch differently; the
    metric restarts at each frame's OWN start phase, so every output -- flags included -- must still be identical."""
    eb, _ = engines
    n = 50000
    _, llr = eb.channel(noise_var=_nv(4.0), n_frames=n, seed=77, stream_id=6, k_payload=40)
    beta = g128["beta_M4"]
    a = _np(eb.dlscl_decode(llr, 4, 8, beta=beta))
    perm = torch.randperm(n, device=llr.device)
    b = _np(eb.dlscl_decode(llr[perm].contiguous(), 4, 8, beta=beta))      # same frames, other queue order
    p = perm.cpu().numpy()
    for k in ("best_bits", "success", "n_attempts", "tried", "flags"):
        assert np.array_equal(a[k][p], b[k]), k
    c = _np(eb.dlscl_decode(llr, 4, 8, beta=beta))
    for k in ("best_bits", "success", "n_attempts", "tried", "flags"):
        assert np.array_equal(a[k], c[k]), k


def test_admission_by_replay_equals_traced_baseline(engines, g128, monkeypatch):
    """PB200_DL_REPLAY=0 (baseline pass traces every frame) and the default (queued frames get their first |L0| row from
    an attempt-0 replay inside the retry kernel) are the same decode: identical outputs and identical sweep counters."""
    eb, _ = engines
    n = 40000
    beta = g128["beta_M4"]
    _, llr = eb.channel(noise_var=_nv(4.0), n_frames=n, seed=5, stream_id=7, k_payload=40)
    outs, cnts = [], []
    for replay in ("1", "0"):
        monkeypatch.setenv("PB200_DL_REPLAY", replay)
        outs.append(_np(eb.dlscl_decode(llr, 4, 8, beta=beta)))
        c = torch.zeros(16, dtype=torch.int64, device="cuda")
        eb.sweep(c, M=4, noise_var=_nv(4.0), n_frames=n, seed=5, stream_id=7, k_payload=40, retries=8, beta=torch.as_tensor(beta, device="cuda"))
        cnts.append(c.cpu().numpy())
    for k in ("best_bits", "success", "n_attempts", "tried", "flags"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    assert np.array_equal(cnts[0], cnts[1])


def test_sweep_counters_equal_full_redecode(engines, g128):
    """Fused Monte-Carlo sweep (Philox channel + SCL + DL-SCL + counters): binned and frame-per-group kernels count the
    same frame errors, bit errors and retries on the same Philox frames."""
    eb, ef = engines
    beta = torch.as_tensor(g128["beta_M4"], device="cuda")
    res = []
    for eng in (eb, ef):
        c = torch.zeros(16, dtype=torch.int64, device="cuda")
        eng.sweep(c, M=4, noise_var=_nv(4.0), n_frames=200000, seed=11, stream_id=3, k_payload=40, retries=8, beta=beta)
        res.append(c.cpu().numpy())
    a, b = res
    assert a[0] == b[0] == 200000 and a[1] == b[1] and a[2] == b[2]           # frames, SCL frame / bit errors
    assert abs(int(a[3]) - int(b[3])) <= 2 and abs(int(a[7]) - int(b[7])) <= 16   # DL frame errors, retries (near-tie frames)


def test_scheduler_statistics(engines, g128, monkeypatch):
    """pb200_debug_bin_stats: every retry decode is counted once (plus one attempt-0 replay per queued frame when the
    baseline pass ran without the trace), batches are (almost) full, decodes start late."""
    import ctypes as C
    eb, _ = engines
    beta = torch.as_tensor(g128["beta_M4"], device="cuda")
    for replay in ("1", "0"):
        monkeypatch.setenv("PB200_DL_REPLAY", replay)      # (unpinned, the engine picks the mode from its last failure fraction)
        c = torch.zeros(16, dtype=torch.int64, device="cuda")
        eb.sweep(c, M=4, noise_var=_nv(4.0), n_frames=400000, seed=2, stream_id=3, k_payload=40, retries=8, beta=beta)
        st = (C.c_uint * 8)()
        assert eb.lib.pb200_debug_bin_stats(eb._h, st) == 0
        cc = c.cpu().numpy()
        queued = int(cc[1])                                  # frames whose baseline decode failed the CRC
        assert st[3] == int(cc[7]) + (queued if replay == "1" else 0)
        assert st[3] / st[2] > 6.5                           # frames per batch (of 8)
        assert st[4] / st[2] > 40                            # mean start phase of a batch


def test_adaptive_admission_is_invisible(engines, g128, monkeypatch):
    """Unpinned, the engine switches the admission mode on the failure fraction of its previous DL-SCL piece (> 12 %: traced
    baseline, else replay).  Whatever it picks, the counters are those of the pinned modes."""
    eb, _ = engines
    monkeypatch.delenv("PB200_DL_REPLAY", raising=False)
    beta = torch.as_tensor(g128["beta_M4"], device="cuda")
    res = {}
    for snr in (3.0, 6.0, 3.0, 6.0, 6.0):                    # alternating: every call sees the other regime's history
        c = torch.zeros(16, dtype=torch.int64, device="cuda")
        eb.sweep(c, M=4, noise_var=_nv(snr), n_frames=100000, seed=9, stream_id=3, k_payload=40, retries=8, beta=beta)
        torch.cuda.synchronize()
        res.setdefault(snr, []).append(c.cpu().numpy())
    for snr, rs in res.items():
        for r in rs[1:]:
            assert np.array_equal(rs[0], r), snr
