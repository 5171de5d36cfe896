"""GPU: Monte-Carlo path -- fused sweep vs its parts, independence from sharding, statistical parity with the
reference's published rows, and size-independent properties at Monte-Carlo batch sizes."""
import math

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


def _counters(eng, **kw):
    c = torch.zeros(16, dtype=torch.int64, device=eng.dev)
    eng.sweep(c, **kw)
    return c.cpu().numpy()


def test_channel_is_a_valid_awgn_channel(eng):
    msg, llr = eng.channel(noise_var=_nv(5.0), n_frames=20000, seed=11, stream_id=50, k_payload=40)
    m = msg.cpu().numpy().astype(np.int8)
    # payload bits are fair, CRC bits are the CRC of the payload, LLR = 2(s+n)/sigma^2
    assert abs(m[:, :40].mean() - 0.5) < 0.01
    for b in range(0, 200, 7):
        assert np.array_equal(O.attach_crc(m[b, :40], CRC24), m[b])
    code = eng.encode(msg).cpu().numpy().astype(np.float64)
    noise = llr.cpu().numpy().astype(np.float64) * _nv(5.0) / 2.0 - (1.0 - 2.0 * code)
    z = noise / math.sqrt(_nv(5.0))
    assert abs(z.mean()) < 3e-3 and abs(z.var() - 1.0) < 5e-3
    assert abs((z ** 3).mean()) < 1e-2 and abs((z ** 4).mean() - 3.0) < 3e-2
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.03
    # same (seed, stream, frame) -> same numbers whatever the launch split; different stream -> different numbers
    _, part = eng.channel(noise_var=_nv(5.0), n_frames=100, frame_begin=7000, seed=11, stream_id=50, k_payload=40)
    assert torch.equal(part, llr[7000:7100])
    _, other = eng.channel(noise_var=_nv(5.0), n_frames=100, frame_begin=7000, seed=11, stream_id=51, k_payload=40)
    assert not torch.equal(other, llr[7000:7100])


@pytest.mark.parametrize("M,retries,beta", [(4, -1, False), (4, 8, True), (1, 8, False), (8, 3, True), (2, 0, False)])
def test_fused_sweep_equals_channel_plus_decode(eng, g128, M, retries, beta):
    n = 6000
    nv = _nv(4.0)
    b = g128[f"beta_M{M}"] if beta else None
    c = _counters(eng, M=M, noise_var=nv, n_frames=n, seed=5, stream_id=40, k_payload=40, retries=retries, beta=b,
                  include_uncoded=True, noise_var_uncoded=1.0 / (2 * 10 ** 0.4))
    msg, llr = eng.channel(noise_var=nv, n_frames=n, seed=5, stream_id=40, k_payload=40)
    scl = eng.scl_decode(llr, M)
    assert c[0] == n
    assert c[1] == int((scl["crc_ok"] == 0).sum())                           # frame error = CRC failure (run_fer_sweep.py:91-94)
    assert c[2] == int((scl["best_bits"] != msg).sum())                      # bit errors over all K bits (:98)
    if retries < 0:
        assert c[8] == int((scl["flags"] & 1).ne(0).sum())
    if retries >= 0:
        dl = eng.dlscl_decode(llr, M, retries, beta=b)
        assert c[8] == int(((scl["flags"] | dl["flags"]) & 1).ne(0).sum())   # ties seen in any attempt of the frame
        assert c[3] == int((dl["success"] == 0).sum())
        assert c[4] == int((dl["best_bits"] != msg).sum())
        assert c[7] == int((dl["n_attempts"] - 1).sum())
        assert c[3] <= c[1]
    assert 0.2 < c[5] / n < 0.45 and c[6] >= c[5]                             # uncoded 40-bit frames at 4 dB


def test_result_does_not_depend_on_sharding(eng, g128):
    kw = dict(M=4, noise_var=_nv(4.5), seed=9, stream_id=45, k_payload=40, retries=8, beta=g128["beta_M4"])
    whole = _counters(eng, n_frames=10000, **kw)
    parts = sum(_counters(eng, n_frames=c, frame_begin=b, **kw) for b, c in [(0, 3333), (3333, 1), (3334, 6666)])
    assert np.array_equal(whole, parts)
    from polar_code_b200 import montecarlo as mc
    a = mc.fer_point(eng, M=4, snr_db=4.5, frames=10000, seed=9, retries=8, beta=g128["beta_M4"], k_payload=40)
    assert np.array_equal(a, whole)


def test_statistical_parity_with_published_rows(eng, g128, published):
    """FER/BER at 5.0 dB (M=4) from 400k Philox frames lies inside the 95% binomial interval of the reference's
    2000-frame row results/fer_M4.csv (and far inside it around the GPU estimate itself)."""
    row = [float(v) for v in published["rows"]["fer_M4"][1].split(",")]
    n_ref, n = 2000, 400_000
    c = _counters(eng, M=4, noise_var=_nv(5.0), n_frames=n, seed=0, stream_id=50, k_payload=40, retries=8,
                  beta=g128["beta_M4"], include_uncoded=True, noise_var_uncoded=1.0 / (2 * 10 ** 0.5))
    est = {"fer_unc": c[5] / n, "fer_scl": c[1] / n, "fer_dl": c[3] / n}
    ref = {"fer_unc": row[1], "fer_scl": row[3], "fer_dl": row[5]}
    for k in est:
        p = est[k]
        half = 1.96 * math.sqrt(p * (1 - p) / n_ref) + 1.96 * math.sqrt(p * (1 - p) / n)
        assert abs(ref[k] - p) <= half, (k, ref[k], p, half)
    # uncoded BER against theory Q(sqrt(2 Eb/N0))
    q = 0.5 * math.erfc(math.sqrt(10 ** 0.5))
    assert abs(c[6] / (n * 40) - q) < 4 * math.sqrt(q / (n * 40))


def test_noiseless_roundtrip_at_scale(eng):
    """encode -> (almost) noiseless channel -> decode returns every word, 1M frames, all list sizes."""
    n = 1 << 20
    msg, llr = eng.channel(noise_var=1e-3, n_frames=n, seed=1, stream_id=0, k_payload=40)
    for M in (1, 4, 8):
        out = eng.scl_decode(llr, M, want=("best_bits", "crc_ok", "flags"))
        assert torch.equal(out["best_bits"], msg) and bool(out["crc_ok"].all())
    assert torch.equal(eng.sc_decode(llr), msg)


def test_encoder_linearity_and_involution(eng):
    g = torch.Generator(device="cpu").manual_seed(0)
    a = torch.randint(0, 2, (4096, 64), generator=g, dtype=torch.uint8)
    b = torch.randint(0, 2, (4096, 64), generator=g, dtype=torch.uint8)
    xa, xb, xab = eng.encode(a), eng.encode(b), eng.encode(a ^ b)
    assert torch.equal(xa ^ xb, xab)
    from polar_code_b200.engine import PolarEngine
    full = PolarEngine(128, np.arange(128, dtype=np.int32), None)          # K = N: the bare transform
    u = torch.randint(0, 2, (4096, 128), generator=g, dtype=torch.uint8)
    assert torch.equal(full.encode(full.encode(u)).cpu(), u)                 # F^{(x)n} is an involution


def test_nr_sweep_and_ber_point(eng):
    from polar_code_b200.engine import PolarEngine, construct_info_set
    from polar_code_b200 import montecarlo as mc
    e = PolarEngine(128, construct_info_set(128, 88), CRC24)
    e.set_rate_matching(256)
    nv = mc.ber_noise_var(3.0, 64, 256)
    n = 4000
    err = torch.zeros(n, dtype=torch.int16, device=e.dev)
    c = torch.zeros(16, dtype=torch.int64, device=e.dev)
    e.sweep(c, M=4, noise_var=nv, n_frames=n, seed=2, stream_id=4, k_payload=64, frame_error_mode=1, bit_error_span=64,
            frame_bit_errors=err)
    msg, llr = e.channel(noise_var=nv, n_frames=n, seed=2, stream_id=4, k_payload=64)
    assert llr.shape == (n, 256)
    out = e.scl_decode(llr, 4)
    be = (out["best_bits"][:, :64] != msg[:, :64]).sum(dim=1)
    assert torch.equal(be.to(torch.int16), err)
    c = c.cpu().numpy()
    assert c[2] == int(be.sum()) and c[1] == int((be > 0).sum())
    # the oracle decodes the very same rate-matched LLRs to the same words (except flagged ties)
    ref = O.nr_decode_batch(llr[:256].cpu().numpy().astype(np.float64), CRC24, 128, construct_info_set(128, 88), 4)
    same = (out["best_bits"][:256].cpu().numpy().astype(np.int8) == ref["best_bits"]).all(axis=1)
    assert ((~same) & ((out["flags"][:256].cpu().numpy() & 1) == 0)).sum() == 0
    st = mc.ber_point(e, M=4, ebn0_db=3.0, payload_len=64, coded_len=256, seed=2, stream_id=4, err_cap=50, bits_cap=1e6)
    ee = err.cpu().numpy().astype(np.int64)
    frames = int(np.argmax(np.cumsum(ee) >= 50)) + 1
    assert (st.frames, st.bit_errors) == (frames, int(ee[:frames].sum()))


def test_full_size_sweep_properties(eng):
    """BASELINE configs[1] size: 1e7 frames at one SNR point.  Properties that do not need the oracle: the counters of
    three unequal shards add up to the single-launch counters; frame errors <= frames; bit errors are consistent with
    frame errors; SC (M=1) is never better than SCL M=4 in CRC-failure rate on the same frames."""
    n = 10_000_000
    kw = dict(noise_var=_nv(5.5), seed=123, stream_id=55, k_payload=40)
    whole = _counters(eng, M=4, n_frames=n, **kw)
    parts = sum(_counters(eng, M=4, n_frames=c, frame_begin=b, **kw) for b, c in [(0, 1), (1, 3_999_999), (4_000_000, 6_000_000)])
    assert np.array_equal(whole, parts)
    assert whole[0] == n and 0 < whole[1] < n and whole[2] >= whole[1] - whole[9]
    sc = _counters(eng, M=1, n_frames=n, **kw)
    assert sc[1] >= whole[1]
    fer = whole[1] / n
    assert 0.005 < fer < 0.03            # SCL M=4 at 5.5 dB (reference fer_M4-like operating point)


def test_decode_calls_on_different_streams_do_not_interfere(eng):
    """Two decode launches of ONE engine enqueued on two streams (they may overlap on the device) give the same
    results as serial execution -- each stream owns its scratch."""
    msg, llr = eng.channel(noise_var=_nv(4.0), n_frames=1 << 19, seed=3, stream_id=1, k_payload=40)
    ref = eng.scl_decode(llr, 4, want=("best_bits", "crc_ok"))
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    half = llr.shape[0] // 2
    for _ in range(3):
        with torch.cuda.stream(s1):
            a = eng.scl_decode(llr[:half], 4, want=("best_bits", "crc_ok"))
        with torch.cuda.stream(s2):
            b = eng.scl_decode(llr[half:], 4, want=("best_bits", "crc_ok"))
        torch.cuda.synchronize()
        assert torch.equal(a["best_bits"], ref["best_bits"][:half]) and torch.equal(b["best_bits"], ref["best_bits"][half:])


def test_host_buffer_path_matches_device_path(eng):
    """pb200_scl_decode_host (chunked, three internal streams) == device-resident decode, ragged batch size."""
    n = (1 << 19) + 12345
    msg, llr = eng.channel(noise_var=_nv(4.5), n_frames=n, seed=8, stream_id=2, k_payload=40)
    ref = eng.scl_decode(llr, 4, want=("best_bits", "crc_ok", "flags"))
    h_llr = torch.empty((n, 128), dtype=torch.float32, pin_memory=True)
    h_llr.copy_(llr)
    h_bits = torch.empty((n, 64), dtype=torch.uint8, pin_memory=True)
    h_ok = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
    h_fl = torch.empty((n,), dtype=torch.int32, pin_memory=True)
    eng.scl_decode_host(h_llr, 4, h_bits, h_ok, h_fl)
    assert torch.equal(h_bits, ref["best_bits"].cpu()) and torch.equal(h_ok, ref["crc_ok"].cpu())
    assert torch.equal(h_fl, ref["flags"].cpu())


@pytest.mark.parametrize("M", [4, 1])
def test_host_buffer_f16_ingest_is_exact_widening(eng, M):
    """pb200_scl_decode_host_f16: binary16 rows are widened to fp32 on load -- the result equals the fp32 host path on
    the widened values bit for bit (ragged batch, > 2 chunks).  Quantising is the caller's choice, outside scl.py's contract."""
    n = (1 << 17) + 4321
    _, llr = eng.channel(noise_var=_nv(4.5), n_frames=n, seed=9, stream_id=2, k_payload=40)
    h16 = torch.empty((n, 128), dtype=torch.float16, pin_memory=True)
    h16.copy_(llr.half())
    ref = eng.scl_decode(h16.to("cuda").float(), M, want=("best_bits", "crc_ok", "flags"))
    h_bits = torch.empty((n, 64), dtype=torch.uint8, pin_memory=True)
    h_ok = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
    h_fl = torch.empty((n,), dtype=torch.int32, pin_memory=True)
    eng.scl_decode_host(h16, M, h_bits, h_ok, h_fl)
    assert torch.equal(h_bits, ref["best_bits"].cpu()) and torch.equal(h_ok, ref["crc_ok"].cpu())
    assert torch.equal(h_fl, ref["flags"].cpu())
    with pytest.raises(ValueError):
        eng.scl_decode_host(h16.double(), M, h_bits, h_ok, h_fl)


@pytest.mark.parametrize("M,snr", [(4, 4.5), (1, 5.0), (8, 4.0)])
def test_fer_matches_oracle_on_reference_channel(eng, g128, M, snr):
    """FER / BER of the GPU sweep (Philox channel, 1e7 frames) against the float64 oracle fed by the reference's own
    PCG64 channel loop (run_fer_sweep.py:60-121 restated in oracle.fer_sweep_frames, 2e5 frames): SCL and DL-SCL
    frame-error rates and the SCL bit-error rate agree within 4 binomial sigmas of the smaller sample."""
    n_ref, n = 200_000, 10_000_000
    msgs, llrs, _ = O.fer_sweep_frames(snr, n_ref, seed=12345)
    A = g128["info_set"]
    s = O.scl_decode_batch(llrs, A, M, crc=CRC24, want_info_llrs=False)
    d = O.dlscl_decode_batch(llrs, A, M, 8, crc=CRC24, beta=None)
    ok_scl = np.array([O.check_crc(b, CRC24) for b in s["best_bits"]])
    ref = {"fer_scl": 1 - ok_scl.mean(), "fer_dl": 1 - d["success"].mean(),
           "ber_scl": (s["best_bits"] != msgs).mean(), "work": (d["n_attempts"] - 1).mean()}
    c = _counters(eng, M=M, noise_var=_nv(snr), n_frames=n, seed=777, stream_id=int(snr * 10), k_payload=40, retries=8)
    est = {"fer_scl": c[1] / n, "fer_dl": c[3] / n, "ber_scl": c[2] / (n * 64), "work": c[7] / n}
    for k in ("fer_scl", "fer_dl"):
        p = est[k]
        sigma = math.sqrt(p * (1 - p) / n_ref)
        assert abs(ref[k] - p) < 4 * sigma, (k, ref[k], p, sigma)
    # bit errors come in bursts (a wrong frame has ~10 wrong bits): scale sigma by the frame-level dispersion
    assert abs(ref["ber_scl"] - est["ber_scl"]) < 4 * est["ber_scl"] / math.sqrt(max(est["fer_scl"] * n_ref, 1))
    assert abs(ref["work"] - est["work"]) < 0.05 * max(est["work"], 1e-3) + 4e-3


def test_frame_bit_errors_exact_above_255():
    """N = 512, K = 488 + 24 (rate 1): at a hopeless operating point every second payload bit is wrong, i.e. 244 +- 11
    errors per frame -- many frames carry > 255; the u16 per-frame counters (and hence run_ber_sweep's adaptive stop)
    must report them exactly (round 1 clamped at 255)."""
    from polar_code_b200.engine import PolarEngine, construct_info_set
    from polar_code_b200 import montecarlo as mc
    kp = 488
    e = PolarEngine(512, construct_info_set(512, 512), CRC24)
    nv = mc.ber_noise_var(-20.0, kp, 512)
    n = 2048
    err = torch.zeros(n, dtype=torch.int16, device=e.dev)
    c = torch.zeros(16, dtype=torch.int64, device=e.dev)
    e.sweep(c, M=2, noise_var=nv, n_frames=n, seed=5, stream_id=1, k_payload=kp, frame_error_mode=1, bit_error_span=kp,
            frame_bit_errors=err)
    msg, llr = e.channel(noise_var=nv, n_frames=n, seed=5, stream_id=1, k_payload=kp)
    out = e.scl_decode(llr, 2)
    be = (out["best_bits"][:, :kp] != msg[:, :kp]).sum(dim=1)
    assert int((be > 255).sum()) > 20, "operating point too benign for this test"
    assert torch.equal(be.to(torch.int64), err.to(torch.int64) & 0xFFFF)
    assert int(c[2]) == int(be.sum())
    with pytest.raises(ValueError):
        e.sweep(c, M=2, noise_var=nv, n_frames=n, k_payload=kp, frame_bit_errors=torch.zeros(n, dtype=torch.uint8, device=e.dev))
