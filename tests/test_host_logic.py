"""CPU: host-side logic -- rank sharding, the sequential stopping rule on batched chunks (1 rank and a
world_size-2 gloo run), CSV writers, NR index helpers, argument parsers, the PNG fallback."""
import math
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from polar_code_b200 import montecarlo as mc


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000, 12345):
        for w in (1, 2, 3, 4, 8):
            spans = [mc.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _sequential(err, work, payload_len, err_cap, bits_cap):
    """The reference loop (run_ber_sweep.py:127, SimulationStats :36-62)."""
    st = dict(bits_total=0, bit_errors=0, frame_errors=0, work=0, frames=0)
    i = 0
    while st["bit_errors"] < err_cap and st["bits_total"] < bits_cap:
        st["bits_total"] += payload_len
        st["bit_errors"] += int(err[i]); st["work"] += int(work[i]); st["frames"] += 1
        st["frame_errors"] += int(err[i] > 0)
        i += 1
    return st


def _chunked(err, work, payload_len, err_cap, bits_cap, first_chunk, rank=0, world=1):
    state = mc.CutState()
    max_frames = int(math.ceil(bits_cap / payload_len))
    begin, chunk = 0, min(first_chunk, max_frames)
    while not state.done and begin < max_frames:
        chunk = min(chunk, max_frames - begin)
        lb, ln = mc.shard_range(chunk, rank, world)
        e = torch.from_numpy(err[begin + lb: begin + lb + ln].astype(np.int16))
        w = torch.from_numpy(work[begin + lb: begin + lb + ln].astype(np.int16))
        state = mc.adaptive_cut(state, e, w, begin + lb, begin, chunk, payload_len, err_cap, bits_cap)
        begin += chunk
        chunk *= 2
    return state


CASES = [(64, 50, 1e4, 7), (64, 1000, 1e7, 64), (8, 2, 64, 3), (64, 10**9, 6400, 16), (40, 1, 1e6, 5), (64, 37, 1e5, 1000)]


@pytest.mark.parametrize("payload_len,err_cap,bits_cap,first_chunk", CASES)
def test_adaptive_cut_equals_sequential_loop(payload_len, err_cap, bits_cap, first_chunk):
    rng = np.random.default_rng(payload_len + err_cap)
    n = int(math.ceil(bits_cap / payload_len)) + 8
    err = (rng.random(n) < 0.03) * rng.integers(1, 9, n)
    work = rng.integers(0, 9, n)
    want = _sequential(err, work, payload_len, err_cap, bits_cap)
    got = _chunked(err, work, payload_len, err_cap, bits_cap, first_chunk)
    assert (got.frames, got.bit_errors, got.frame_errors, got.work_sum) == \
        (want["frames"], want["bit_errors"], want["frame_errors"], want["work"])


def test_adaptive_cut_counts_above_255_exactly():
    """K_payload = 300 (N = 512): a badly failed frame carries more than 255 payload errors; the per-frame counters
    are 16-bit so the prefix sums of the stopping rule stay exact (round-1 clamp at 255 undercounted)."""
    rng = np.random.default_rng(7)
    n = 4000
    err = (rng.random(n) < 0.05) * rng.integers(200, 301, n)
    work = rng.integers(0, 400, n)
    assert err.max() > 255 and work.max() > 255
    want = _sequential(err, work, 300, 20000, 1e9)
    got = _chunked(err, work, 300, 20000, 1e9, 64)
    assert (got.frames, got.bit_errors, got.frame_errors, got.work_sum) == \
        (want["frames"], want["bit_errors"], want["frame_errors"], want["work"])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for payload_len, err_cap, bits_cap, first_chunk in CASES:
        rng = np.random.default_rng(payload_len + err_cap)
        n = int(math.ceil(bits_cap / payload_len)) + 8
        err = (rng.random(n) < 0.03) * rng.integers(1, 9, n)
        work = rng.integers(0, 9, n)
        st = _chunked(err, work, payload_len, err_cap, bits_cap, first_chunk, rank, world)
        out.append((st.frames, st.bit_errors, st.frame_errors, st.work_sum))
    # counter block all-reduce + sharded counting == whole
    begin, count = mc.shard_range(1001, rank, world)
    c = torch.zeros(mc.NCOUNTERS, dtype=torch.int64)
    c[0] = count; c[1] = sum(1 for f in range(begin, begin + count) if f % 7 == 0)
    mc.reduce_counters(c)
    out.append(tuple(int(v) for v in c[:2]))
    q.put((rank, out))
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(timeout=60)
    assert res[0] == res[1]
    for (payload_len, err_cap, bits_cap, first_chunk), got in zip(CASES, res[0]):
        rng = np.random.default_rng(payload_len + err_cap)
        n = int(math.ceil(bits_cap / payload_len)) + 8
        err = (rng.random(n) < 0.03) * rng.integers(1, 9, n)
        work = rng.integers(0, 9, n)
        want = _sequential(err, work, payload_len, err_cap, bits_cap)
        assert got == (want["frames"], want["bit_errors"], want["frame_errors"], want["work"])
    assert res[0][-1] == (1001, len([f for f in range(1001) if f % 7 == 0]))


def test_noise_variances():
    assert mc.fer_noise_var(5.0, 64, 128) == pytest.approx(1.0 / (2 * 0.5 * 10 ** 0.5))
    assert mc.ber_noise_var(3.0, 64, 256) == pytest.approx(O.noise_var_ber(3.0, 64, 256))


def test_nr_index_helpers_match_oracle():
    from dl_scl_polar.nr.polar import (subblock_interleave, subblock_deinterleave, rate_match_polar, derate_match_polar)
    rng = np.random.default_rng(3)
    for n in (16, 40, 128, 100):
        v = rng.normal(size=n)
        assert np.array_equal(subblock_interleave(v), O.subblock_interleave(v))
        assert np.array_equal(subblock_deinterleave(subblock_interleave(v), n), v)
    x = rng.normal(size=300)
    for E in (16, 96, 128, 256, 300):
        assert np.array_equal(derate_match_polar(x[:E], 128), O.derate_match(x[:E], 128))
        assert np.array_equal(rate_match_polar(np.arange(128), E), O.rate_match(np.arange(128), E))
    assert rate_match_polar(np.arange(128), 64).size == 64 and rate_match_polar(np.arange(128), 200).size == 200
    with pytest.raises(ValueError):
        subblock_interleave(np.zeros((2, 2)))


def test_ber_cli_helpers(tmp_path):
    from dl_scl_polar.eval import run_ber_sweep as R
    payload = np.array([0, 1, 1, 0], dtype=np.int8)
    cand = np.concatenate([payload, np.array([1, 0, 0, 1], dtype=np.int8)])
    cand[-1] ^= 1
    assert R._payload_bit_errors(payload, cand, 4) == 0            # CRC-only error is not a payload error
    assert R._payload_bit_errors(payload, None, 4) == 4
    with pytest.raises(ValueError):
        R._payload_bit_errors(payload, cand[:2], 4)
    st = R.SimulationStats()
    st.update(2, 1.0, True, 8); st.update(0, 3.0, False, 8)
    assert st.row() == {"bits_total": 16, "bit_errors": 2, "ber": 0.125, "fer": 0.5, "avg_work": 2.0}
    with pytest.raises(ValueError):
        R.parse_args(["--scheme", "dl_scl", "--K_payload", "64", "--K_crc", "24", "--E", "128", "--EbN0_lo", "1",
                      "--EbN0_hi", "2", "--out", "x.csv"])
    rows = [dict(zip(R.HEADER, ["polar_scl", "polar_scl", 16, 8, 4, 0.5, "M=2", 6.0, 64, 0, 0.0, 0.0, 0.0]))]
    R.write_csv(rows, tmp_path / "o.csv")
    txt = (tmp_path / "o.csv").read_text().splitlines()
    assert txt[0] == "scheme,code,N_or_E,K_payload,K_crc,rate,params,EbN0_dB,bits_total,bit_errors,ber,fer,avg_work"
    assert txt[1] == "polar_scl,polar_scl,16,8,4,0.5,M=2,6.0,64,0,0.0,0.0,0.0"


def test_fer_cli_csv_format(tmp_path):
    from dl_scl_polar.eval import run_fer_sweep as F
    args = F.build_argparser().parse_args(["--M", "4", "--out_dir", str(tmp_path / "r"), "--plot_dir", str(tmp_path / "p"),
                                           "--include_uncoded"])
    assert (args.frames, args.snr_lo, args.snr_hi, args.snr_step, args.retries, args.seed) == (10000, 4.0, 6.5, 0.5, 8, 0)
    rows = [{"snr_db": 5.0, "fer_uncoded": 0.218, "ber_uncoded": 0.0062, "fer_scl": 0.0455, "ber_scl": 0.007132813,
             "fer_dl": 0.0355, "ber_dl": 0.01257813}]
    F.write_outputs(args, rows)
    txt = (tmp_path / "r" / "fer_M4.csv").read_text().splitlines()
    assert txt[0] == "snr_db,fer_uncoded,ber_uncoded,fer_scl,ber_scl,fer_dl,ber_dl"
    assert txt[1] == "5.000,2.180000e-01,6.200000e-03,4.550000e-02,7.132813e-03,3.550000e-02,1.257813e-02"  # results/fer_M4.csv:2
    png = (tmp_path / "p" / "fer_M4.png").read_bytes()
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    assert list(F._snr_grid(args)) == pytest.approx([4.0, 4.5, 5.0, 5.5, 6.0, 6.5])


def test_mirror_validation_without_gpu():
    """Argument errors are raised before any device work, with the reference's exception types."""
    from dl_scl_polar.polar.polar import construct_info_set, encode, sc_decode
    from dl_scl_polar.polar.crc import _poly_to_bits, attach_crc, check_crc
    from dl_scl_polar.polar.scl import decode_scl
    from dl_scl_polar.dlscl.flip import choose_flip_index, retry_with_flip, _force_vector
    A = construct_info_set(128, 64)
    assert A is construct_info_set(128, 64) and A.dtype == np.int32 and A.size == 64
    assert "".join(map(str, _poly_to_bits("0x1864CFB"))) == "1100001100100110011111011"
    with pytest.raises(ValueError):
        construct_info_set(100, 50)
    with pytest.raises(ValueError):
        encode(np.zeros((2, 32), np.int8))
    with pytest.raises(ValueError):
        encode(np.zeros(63, np.int8))
    with pytest.raises(ValueError):
        sc_decode(np.zeros(100), A)
    with pytest.raises(ValueError):
        sc_decode(np.zeros(64), A)             # info_set out of range for N=64
    with pytest.raises(ValueError):
        attach_crc(np.zeros((2, 2), np.int8), "0x17")
    with pytest.raises(ValueError):
        attach_crc(np.zeros(4, np.int8), "")
    with pytest.raises(ValueError):
        check_crc(np.zeros(4, np.int8), "0x17")
    with pytest.raises(ValueError):
        decode_scl(np.zeros(128), A, 0)
    with pytest.raises(ValueError):
        decode_scl(np.zeros(128), A, 4, force_info_bits=np.full(64, 2, np.int8))
    with pytest.raises(ValueError):
        decode_scl(np.zeros(128), A, 4, force_info_bits=np.zeros(10, np.int8))
    with pytest.raises(ValueError):
        choose_flip_index(np.array([]), None)
    with pytest.raises(ValueError):
        choose_flip_index(np.ones(4), np.ones((3, 3)))
    with pytest.raises(IndexError):
        retry_with_flip(np.zeros(128), A, 4, np.zeros(64, np.int8), 64)
    f = _force_vector(np.array([1, 0, 1, 1], np.int8), 2)
    assert list(f) == [1, 0, 0, -1]


def test_ldpc_h_matrix_host_side(gldpc):
    """pb200_ldpc_build_h is host code of the C-ABI: no GPU needed (basegraphs.py + builder.py)."""
    from polar_code_b200.ldpc import build_h_matrix
    from polar_code_b200.dl_scl_polar.nr.ldpc import load_base_graph
    from polar_code_b200.dl_scl_polar.nr.ldpc.builder import build_h_matrix as mirror_build
    for bg, Z in [(1, 2), (2, 2), (2, 4), (1, 8), (2, 32)]:
        assert np.array_equal(build_h_matrix(bg, Z), gldpc[f"H_bg{bg}_Z{Z}"])
        assert np.array_equal(mirror_build(load_base_graph(bg), Z), gldpc[f"H_bg{bg}_Z{Z}"])
    g = load_base_graph(2)
    assert (g.m, g.n) == (3, 6) and g.shifts[1, 2] == 3 and g.shifts[0, 4] == -1
    with pytest.raises(ValueError):
        load_base_graph(7)
    with pytest.raises(ValueError):
        build_h_matrix(2, 0)


def test_numa_binding_is_best_effort():
    """bind_to_gpu_numa must never raise: no CUDA device / no sysfs PCI tree -> None, affinity untouched."""
    from polar_code_b200.montecarlo import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa(0) is None or isinstance(bind_to_gpu_numa(0), str)
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before
