"""Generate tests/golden/*.npz by running the REFERENCE itself (test infrastructure).

Run in the build container only (``/root/reference`` is not present on the GPU
box):  ``python oracle/gen_golden.py``.  The reference's hot-path modules need
only NumPy, so they are imported straight from ``/root/reference``; nothing is
copied.  All LLR inputs are float32-representable (stored as float32) so the
same vectors drive the float64 reference/oracle and the fp32 CUDA engine.

Fixtures written:
  scl_p128.npz   P(128,64)+CRC-24: decode_scl (M=1,2,4,8), forced decode,
                 sc_decode, decode_with_retries (beta / |L0| ranking), encode, CRC
  scl_toy.npz    N=16,K=12 poly 0x17 (tests/test_ber_eval.py geometry), N=8, N=32
  nr_p128.npz    NR chain E=256 and E=96 -> N=128, K=88, decode_rate_matched_scl
  published.json rows of results/fer_M{1,4,8}.csv + the recipe that reproduces them
  ldpc.npz       nr/ldpc: H matrices, encode_ldpc, derate_match_ldpc, decode_ldpc_nms outputs and rows of the
                 reference's own run_ber_sweep --scheme nr_ldpc (``python oracle/gen_golden.py ldpc`` writes only this)
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)

from dl_scl_polar.polar.polar import construct_info_set, sc_decode, _polar_transform  # noqa: E402
from dl_scl_polar.polar.crc import attach_crc, check_crc  # noqa: E402
from dl_scl_polar.polar.scl import decode_scl  # noqa: E402
from dl_scl_polar.dlscl.flip import decode_with_retries, choose_flip_index  # noqa: E402
from dl_scl_polar.nr.polar import (  # noqa: E402
    subblock_interleave, subblock_deinterleave, rate_match_polar, derate_match_polar,
    encode_rate_matched, decode_rate_matched_scl,
)

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"
CRC24 = "0x1864CFB"


def _frames(rng, n_frames, N, info_set, k_payload, poly, snr_db, rate_bits):
    """Random payload -> CRC -> encode -> BPSK/AWGN -> float32 LLRs."""
    msgs, llrs = [], []
    nv = 1.0 / (2.0 * (rate_bits / N) * 10 ** (snr_db / 10.0))
    for _ in range(n_frames):
        payload = rng.integers(0, 2, size=k_payload, dtype=np.int8)
        msg = attach_crc(payload, poly) if poly else payload
        u = np.zeros(N, np.int8)
        u[info_set] = msg
        x = _polar_transform(u)
        y = 1.0 - 2.0 * x + rng.normal(0.0, np.sqrt(nv), size=N)
        llrs.append((2.0 * y / nv).astype(np.float32))
        msgs.append(msg)
    return np.array(msgs, np.int8), np.array(llrs, np.float32)


def _pack_scl(out, tag, llr32, info_set, M, crc, force=None):
    B, K = llr32.shape[0], info_set.size
    cand = np.zeros((B, M, K), np.int8)
    met = np.full((B, M), np.inf)
    ill = np.zeros((B, M, K))
    ncand = np.zeros(B, np.int32)
    best = np.zeros(B, np.int32)
    for b in range(B):
        r = decode_scl(llr32[b].astype(np.float64), info_set, M, crc=crc,
                       force_info_bits=None if force is None else force[b])
        nc = len(r["candidates"])
        ncand[b] = nc
        cand[b, :nc] = np.array(r["candidates"])
        met[b, :nc] = r["metrics"]
        ill[b, :nc] = np.array(r["info_llrs"])
        best[b] = next(i for i, c in enumerate(r["candidates"]) if c is r["best_path_bits"])
    out[f"{tag}_cand"] = cand
    out[f"{tag}_metrics"] = met
    out[f"{tag}_info_llrs"] = ill
    out[f"{tag}_n_cand"] = ncand
    out[f"{tag}_best"] = best


def _pack_dl(out, tag, llr32, info_set, M, retries, crc, beta):
    B, K = llr32.shape[0], info_set.size
    bits = np.zeros((B, K), np.int8)
    succ = np.zeros(B, np.int8)
    natt = np.zeros(B, np.int32)
    tried = np.full((B, max(retries, 1)), -1, np.int32)
    for b in range(B):
        r = decode_with_retries(llr32[b].astype(np.float64), info_set, M, retries, crc=crc, beta=beta)
        bits[b] = r["best_path_bits"]
        succ[b] = r["success"]
        natt[b] = len(r["attempts"])
        t = [int(i) for i in r["tried_indices"]]
        tried[b, : len(t)] = t
    out[f"{tag}_bits"] = bits
    out[f"{tag}_success"] = succ
    out[f"{tag}_n_attempts"] = natt
    out[f"{tag}_tried"] = tried


def gen_p128():
    rng = np.random.default_rng(20261018)
    A = construct_info_set(128, 64)
    out = {"info_set": A, "info_set_88": construct_info_set(128, 88),
           "info_set_pw": construct_info_set(128, 64, "polarization")}
    m1, l1 = _frames(rng, 24, 128, A, 40, CRC24, 3.0, 64)
    m2, l2 = _frames(rng, 24, 128, A, 40, CRC24, 4.5, 64)
    msgs = np.concatenate([m1, m2])
    llr = np.concatenate([l1, l2])
    # two saturated / degenerate frames: noiseless +-1e6 and all-zero LLRs
    u = np.zeros(128, np.int8)
    u[A] = msgs[0]
    llr[0] = ((1.0 - 2.0 * _polar_transform(u)) * 1e6).astype(np.float32)
    llr[1] = 0.0
    out["msgs"] = msgs
    out["llr"] = llr
    codes = []
    for m in msgs:
        u = np.zeros(128, np.int8)
        u[A] = m
        codes.append(_polar_transform(u))
    out["codes"] = np.array(codes, np.int8)
    out["crc_kat"] = attach_crc(np.array([1] + [0] * 39, np.int8), CRC24)
    out["sc_bits"] = np.array([sc_decode(l.astype(np.float64), A) for l in llr], np.int8)
    for M in (1, 2, 4, 8):
        _pack_scl(out, f"scl_M{M}", llr, A, M, CRC24)
    _pack_scl(out, "scl_M3_nocrc", llr[:16], A, 3, None)
    # forced decode: prefix of the M=4 best path, flip at a frame-dependent index
    force = np.full((llr.shape[0], 64), -1, np.int8)
    for b in range(llr.shape[0]):
        i = int(rng.integers(0, 64))
        bb = out["scl_M4_cand"][b, out["scl_M4_best"][b]]
        force[b, :i] = bb[:i]
        force[b, i] = 1 - bb[i]
    force[3, 40] = 1  # a forced bit after free bits (scattered force pattern)
    out["force"] = force
    _pack_scl(out, "scl_M4_forced", llr, A, 4, CRC24, force)
    for M in (1, 2, 4, 8):
        beta = np.load(f"{REF}/checkpoints/beta_M{M}.npy")
        out[f"beta_M{M}"] = beta
        _pack_dl(out, f"dl_M{M}", llr, A, M, 8, CRC24, beta)
    _pack_dl(out, "dl_M2_nobeta_r4", llr, A, 2, 4, CRC24, None)
    _pack_dl(out, "dl_M4_r0", llr[:8], A, 4, 0, CRC24, None)
    # beta ranking on the shipped dataset rows (data/train_M4...npz)
    d = np.load(f"{REF}/data/train_M4_snr5_seed0_part0.npz")
    al = d["abs_l0"][:64].astype(np.float32)
    out["rank_abs_l0"] = al
    out["rank_idx_beta"] = np.array([choose_flip_index(a.astype(np.float64), out["beta_M4"]) for a in al], np.int32)
    out["rank_idx_none"] = np.array([choose_flip_index(a.astype(np.float64), None) for a in al], np.int32)
    np.savez_compressed(OUT / "scl_p128.npz", **out)


def gen_toy():
    rng = np.random.default_rng(7)
    out = {}
    for (N, K, kp, poly, snr, tag) in [(16, 12, 8, "0x17", 3.0, "n16"), (8, 4, 4, None, 1.0, "n8"),
                                       (32, 20, 12, "0x1D5", 2.0, "n32"), (256, 128, 104, CRC24, 2.5, "n256")]:
        A = construct_info_set(N, K)
        msgs, llr = _frames(rng, 12, N, A, kp, poly, snr, K)
        out[f"{tag}_info_set"] = A
        out[f"{tag}_msgs"] = msgs
        out[f"{tag}_llr"] = llr
        for M in (1, 2, 4):
            _pack_scl(out, f"{tag}_M{M}", llr, A, M, poly)
    np.savez_compressed(OUT / "scl_toy.npz", **out)


def gen_nr():
    rng = np.random.default_rng(11)
    A = construct_info_set(128, 88)
    out = {"info_set": A}
    v = np.arange(40, dtype=np.float64)
    out["ilv40"] = subblock_interleave(v)
    out["deilv40"] = subblock_deinterleave(out["ilv40"], 40)
    out["ilv128"] = subblock_interleave(np.arange(128, dtype=np.float64))
    for E in (256, 96, 128, 300):
        nv = 1.0 / (2.0 * 10 ** 0.3 * 64 / E)
        pays, llrs, bits, ok = [], [], [], []
        for _ in range(12):
            payload = rng.integers(0, 2, size=64, dtype=np.int8)
            tx = encode_rate_matched(payload, CRC24, 128, E, A)
            y = 1.0 - 2.0 * tx.astype(np.float64) + rng.normal(0.0, np.sqrt(nv), size=E)
            l32 = (2.0 * y / nv).astype(np.float32)
            r = decode_rate_matched_scl(l32.astype(np.float64), CRC24, 128, E, A, 4)
            pays.append(payload); llrs.append(l32); bits.append(r["best_path_bits"]); ok.append(r["crc_pass"])
        out[f"E{E}_payload"] = np.array(pays, np.int8)
        out[f"E{E}_llr"] = np.array(llrs, np.float32)
        out[f"E{E}_bits"] = np.array(bits, np.int8)
        out[f"E{E}_crc_pass"] = np.array(ok, np.int8)
        out[f"E{E}_internal"] = np.array(
            [subblock_deinterleave(derate_match_polar(l.astype(np.float64), 128), 128) for l in llrs])
        out[f"E{E}_tx"] = np.array([rate_match_polar(subblock_interleave(np.arange(128)), E)], np.int64)
    np.savez_compressed(OUT / "nr_p128.npz", **out)


def gen_published():
    rows = {}
    for M in (1, 4, 8):
        rows[f"fer_M{M}"] = Path(f"{REF}/results/fer_M{M}.csv").read_text().strip().splitlines()
    recipe = {
        "fer_M4": {"M": 4, "frames": 2000, "snr": [5.0]},
        "fer_M8": {"M": 8, "frames": 2000, "snr": [5.0]},
        "fer_M1": {"M": 1, "frames": 3000, "snr": [4.5, 5.0, 5.5, 6.0]},
        "common": {"seed": 0, "retries": 8, "include_uncoded": True, "beta": "checkpoints/beta_M{M}.npy"},
    }
    (OUT / "published.json").write_text(json.dumps({"rows": rows, "recipe": recipe}, indent=1))


def gen_ldpc():
    """nr/ldpc fixtures from the reference."""
    import types
    from dl_scl_polar.nr.ldpc import (load_base_graph, build_h_matrix, encode_ldpc, rate_match_ldpc,
                                      derate_match_ldpc, decode_ldpc_nms)
    rng = np.random.default_rng(23)
    out = {}
    for bg, Z in [(1, 2), (2, 2), (2, 4), (1, 8), (2, 32)]:
        out[f"H_bg{bg}_Z{Z}"] = build_h_matrix(load_base_graph(bg), Z)
    # encoder + decoder on noisy frames; LLRs are float32-representable doubles
    for Z, E, snr_db, max_iter, alpha, tag in [(2, 12, 3.0, 20, 0.8, "z2"), (4, 24, 2.0, 20, 0.8, "z4"),
                                               (8, 61, 1.0, 10, 0.9, "z8e61"), (8, 130, -1.0, 20, 0.8, "z8e130"),
                                               (32, 384, -1.5, 20, 0.8, "z32e384"), (4, 24, 1.0, 0, 0.8, "z4it0"),
                                               (4, 24, 0.0, 3, 0.75, "z4it3")]:
        H = build_h_matrix(load_base_graph(2), Z)
        n = H.shape[1]
        k = n - H.shape[0]
        nv = 1.0 / (2.0 * 10 ** (snr_db / 10.0) * k / E)
        pays, codes, llrs, ders, hards, its, oks = [], [], [], [], [], [], []
        for _ in range(24):
            payload = rng.integers(0, 2, size=k, dtype=np.int8)
            cw = encode_ldpc(payload, H)
            tx = rate_match_ldpc(cw, E)
            y = 1.0 - 2.0 * tx.astype(np.float64) + rng.normal(0.0, np.sqrt(nv), size=E)
            l = (2.0 * y / nv).astype(np.float32).astype(np.float64)
            d = derate_match_ldpc(l, n)
            r = decode_ldpc_nms(d, H, max_iter=max_iter, alpha=alpha)
            pays.append(payload); codes.append(cw); llrs.append(l); ders.append(d)
            hards.append(r["hard"]); its.append(r["iters_used"]); oks.append(r["parity_ok"])
        out[f"{tag}_cfg"] = np.array([Z, E, max_iter], np.int64)
        out[f"{tag}_alpha"] = np.array([alpha])
        out[f"{tag}_payload"] = np.array(pays, np.int8)
        out[f"{tag}_code"] = np.array(codes, np.int8)
        out[f"{tag}_llr"] = np.array(llrs, np.float64)
        out[f"{tag}_derated"] = np.array(ders, np.float64)
        out[f"{tag}_hard"] = np.array(hards, np.int8)
        out[f"{tag}_iters"] = np.array(its, np.int32)
        out[f"{tag}_ok"] = np.array(oks, np.int8)
    # no early stop
    H = build_h_matrix(load_base_graph(2), 4)
    r = [decode_ldpc_nms(l, H, max_iter=6, alpha=0.8, early_stop=False) for l in out["z4_derated"]]
    out["z4_noearly_hard"] = np.array([x["hard"] for x in r], np.int8)
    out["z4_noearly_iters"] = np.array([x["iters_used"] for x in r], np.int32)
    out["z4_noearly_ok"] = np.array([x["parity_ok"] for x in r], np.int8)
    # the reference's own CLI loop (PCG64 stream): rows of run_ber_sweep.run for two small nr_ldpc set-ups
    mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("figure", "semilogy", "xlabel", "ylabel", "grid", "legend", "tight_layout", "savefig", "close"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl); sys.modules.setdefault("matplotlib.pyplot", plt)
    from dl_scl_polar.eval import run_ber_sweep
    cli = {
        "cli_a": ["--scheme", "nr_ldpc", "--K_payload", "6", "--K_crc", "0", "--E", "12", "--bg", "2", "--Z", "2",
                  "--EbN0_lo", "2.0", "--EbN0_hi", "4.0", "--EbN0_step", "1.0", "--bits_cap", "6000", "--err_cap", "60",
                  "--out", "/tmp/_ldpc_a.csv", "--crc_poly", "0x1", "--seed", "3"],
        "cli_b": ["--scheme", "nr_ldpc", "--K_payload", "20", "--K_crc", "4", "--E", "70", "--bg", "1", "--Z", "8",
                  "--EbN0_lo", "1.0", "--EbN0_hi", "2.0", "--EbN0_step", "1.0", "--bits_cap", "8000", "--err_cap", "50",
                  "--max_iter", "12", "--alpha", "0.75", "--out", "/tmp/_ldpc_b.csv", "--crc_poly", "0x17", "--seed", "5"],
    }
    cli_rows = {}
    for tag, argv in cli.items():
        rows = run_ber_sweep.run(run_ber_sweep.parse_args(argv))
        cli_rows[tag] = {"argv": argv, "rows": [{k: (v if isinstance(v, str) else float(v)) for k, v in r.items()} for r in rows]}
    out["cli_json"] = np.frombuffer(json.dumps(cli_rows).encode(), np.uint8)
    np.savez_compressed(OUT / "ldpc.npz", **out)


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    if sys.argv[1:] == ["ldpc"]:
        gen_ldpc(); print("ldpc done")
        sys.exit(0)
    gen_p128(); print("p128 done")
    gen_toy(); print("toy done")
    gen_nr(); print("nr done")
    gen_published()
    gen_ldpc(); print("ldpc done")
    print("golden vectors written to", OUT)
