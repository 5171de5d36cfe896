#!/usr/bin/env python
"""bench.py -- headline benchmark: SCL M=4, P(128,64)+CRC-24 decoded frames/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (decode_scl, list size 4, CRC-24 selection) over one batch of synthetic
AWGN frames spread evenly over the reference sweep's SNR grid 4.0 .. 6.5 dB (config[0] of BASELINE.json, run
at a Monte-Carlo batch size).  `value` times the decode kernel with the LLRs already resident in HBM; `e2e`
times the same decode through the host-buffer C-ABI call (pinned host LLRs in, decisions out, copies inside
the timed region).  The reference arm (--impl reference) times the CPU path (the pinned C oracle, all host
threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

# rank 0 prints exactly ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner
# with plain stdio), so the real stdout is kept on a private descriptor and fd 1 is pointed at stderr for everything else.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CRC24 = "0x1864CFB"
N, K, KP, M = 128, 64, 40, 4
SNR_GRID = [4.0, 4.5, 5.0, 5.5, 6.0, 6.5]          # run_fer_sweep defaults (config.py:19, run_fer_sweep.py:198-200)
METRIC = "SCL M=4 N=128 decoded frames/s"
UNIT = "frames/s"
# algorithmic work per frame (lazy SC schedule, golden info set)
ELEM_OPS = {1: 1088, 2: 2023, 4: 3843, 8: 7439}
HBM_BYTES_PER_FRAME = 4 * N + K + 1 + 4            # LLR row in, best_bits + crc_ok + flags out (u8 decisions)


def noise_var(snr_db: float) -> float:
    return 1.0 / (2.0 * (K / N) * 10 ** (snr_db / 10.0))   # run_fer_sweep.py:62-64


def ncu_constants() -> dict:
    """Per-frame constants of the dominant kernel measured with ncu and tracked under profiles/ (written by
    scripts/ncu_summary.py --json from the capture of the SAME build; the roofline lines name the file they come from)."""
    p = ROOT / "profiles" / "r02_constants.json"
    if p.exists():
        d = json.loads(p.read_text())
        k = d.get("decode_kernel_M4", {})
        if {"warp_instr_per_frame", "issue_slots_busy_pct", "dram_bytes_per_frame"} <= set(k):
            return {**k, "file": "profiles/r02_constants.json"}
    return {"warp_instr_per_frame": None, "issue_slots_busy_pct": None, "dram_bytes_per_frame": None, "file": None}


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "sm_max_mhz": float(d.get("sm_max_mhz", 1965.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons streamed at 50 ms during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self._t = None

    def _reader(self):
        for line in self.proc.stdout:
            cells = [c.strip() for c in line.split(",")]
            if len(cells) > 8:
                self.rows.append((time.perf_counter(), cells))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu),
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._reader, daemon=True)
            self._t.start()
        except Exception:
            self.proc = None

    def mark(self) -> float:
        return time.perf_counter()

    def stop(self, t0: float = 0.0, t1: float = float("inf")) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self._t:
            self._t.join(timeout=2)
        inside = [c for (t, c) in self.rows if t0 <= t <= t1] or [c for (_, c) in self.rows]
        num = lambda v: float(v) if v.replace(".", "", 1).isdigit() else None
        sm = [num(c[1]) for c in inside if num(c[1]) is not None]
        mx = [num(c[2]) for c in inside if num(c[2]) is not None]
        pw = [num(c[3]) for c in inside if num(c[3]) is not None]
        reasons = set()
        for c in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(sample_frames: int, threads: int = 0) -> dict:
    """Oracle port (oracle/polar_oracle.c, fp64, pthreads over frames) on a bounded sample of the workload."""
    from oracle import oracle as O
    A = O.construct_info_set(N, K)
    rng = np.random.default_rng(12345)
    per = sample_frames // len(SNR_GRID)
    llrs = []
    for s in SNR_GRID:
        nv = noise_var(s)
        payload = rng.integers(0, 2, (per, KP), dtype=np.int8)
        msgs = np.array([O.attach_crc(p, CRC24) for p in payload[:256]])
        reps = (per + 255) // 256
        codes = np.tile(np.array([O.encode(m, A, N) for m in msgs]), (reps, 1))[:per]
        llrs.append(2.0 * (1.0 - 2.0 * codes + rng.normal(0, np.sqrt(nv), codes.shape)) / nv)
    llr = np.ascontiguousarray(np.concatenate(llrs))
    nthreads = threads or O.lib().po_num_threads()
    O.scl_decode_batch(llr[:2048], A, M, crc=CRC24, want_info_llrs=False, nthreads=nthreads)   # warm
    t = time.perf_counter()
    O.scl_decode_batch(llr, A, M, crc=CRC24, want_info_llrs=False, nthreads=nthreads)
    dt = time.perf_counter() - t
    return {"value": llr.shape[0] / dt, "unit": UNIT, "cores": int(nthreads), "kind": "port",
            "sample": f"{llr.shape[0]} frames (SNR grid 4.0-6.5 dB), decode_scl M=4 + CRC-24, float64 C oracle, {dt:.2f} s wall"}


def run_reference(args) -> None:
    """Reference arm: the CPU path (pinned C oracle, all host threads) on bounded samples of the same workload.
    Under torchrun only rank 0 works; the other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    # keep the whole run within a few minutes whatever K is: ~2.4 M frames of CPU work in total (about 10 s on 16 cores)
    sample = int(min(args.cpu_sample, max(60_000, 2_400_000 // steps)))
    for _ in range(min(warm, 2)):
        cpu_baseline(max(sample // 4, 12_000))
    vals = [cpu_baseline(sample) for _ in range(steps)]
    v = float(np.mean([x["value"] for x in vals]))
    cb = dict(vals[-1])
    cb["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(sample, args.gpus, cpu=True),
        "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def workload_config(frames_per_step: int, gpus: int, cpu: bool = False) -> dict:
    return {"workload": "decode_scl M=4 + CRC-24A selection on P(128,64) (40 payload + 24 CRC), AWGN BPSK LLRs, "
                        "frames spread evenly over Eb/N0 4.0-6.5 dB step 0.5 (BASELINE configs[0] geometry at Monte-Carlo batch size)",
            "N": N, "K": K, "M": M, "crc": CRC24, "frames_per_step_per_gpu": frames_per_step,
            "inputs": "CPU-resident float64 LLRs" if cpu else "fp32 LLR rows resident in HBM (larger than L2; no flush needed)",
            "parallelism": f"frames sharded over {gpus} GPU(s), no data-path collective; int64 counter all-reduce per step"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1 << 22, help="frames per step per GPU (device-resident leg)")
    ap.add_argument("--e2e-frames", type=int, default=1 << 21, help="frames per step per GPU (host-buffer leg)")
    ap.add_argument("--cpu-sample", type=int, default=1_200_000, help="frames of the CPU baseline sample")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from polar_code_b200.engine import PolarEngine, construct_info_set

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the polar_b200 engine has no CPU fallback")
    torch.cuda.set_device(local)
    from polar_code_b200.montecarlo import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local) if world > 1 else None     # host buffers of a rank live on its GPU's socket
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    Ksteps = max(args.steps, 1)
    B = args.frames

    A = construct_info_set(N, K)
    eng = PolarEngine(N, A, CRC24, device=local)
    # synthetic input: Philox AWGN channel on the device (csrc/polar_sweep.cuh), one SNR point per slice
    per = B // len(SNR_GRID)
    parts, msgs = [], []
    for i, s in enumerate(SNR_GRID):
        nfr = per if i < len(SNR_GRID) - 1 else B - per * (len(SNR_GRID) - 1)
        m, l = eng.channel(noise_var=noise_var(s), n_frames=nfr, frame_begin=rank * B + i * per, seed=2026, stream_id=i, k_payload=KP)
        parts.append(l)
        msgs.append(m)
    llr = torch.cat(parts)
    msg = torch.cat(msgs)
    del parts, msgs
    best_bits = torch.empty((B, K), dtype=torch.uint8, device=dev)
    crc_ok = torch.empty((B,), dtype=torch.uint8, device=dev)
    flags = torch.empty((B,), dtype=torch.int32, device=dev)
    counters = torch.zeros(16, dtype=torch.int64, device=dev)

    import ctypes as C
    from polar_code_b200 import _lib as L
    so = L.SclOut(best_bits=best_bits.data_ptr(), crc_ok=crc_ok.data_ptr(), flags=flags.data_ptr())
    launches = {"n": 0}

    def step():
        L.check(eng.lib.pb200_scl_decode_batch(eng._h, C.c_void_p(llr.data_ptr()), B, N, None, M, C.byref(so),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        launches["n"] += 1
        if world > 1:
            # per-step error counters reduced over NVLink; frames themselves never move
            counters[0] = B
            counters[1] = B - crc_ok.sum()
            dist.all_reduce(counters)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(W):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 3.0:   # first nvidia-smi line before the timed region
            time.sleep(0.02)
    launches["n"] = 0
    t_begin = sampler.mark()
    ms = timed(step, Ksteps)
    t_end = sampler.mark()
    n_launch = launches["n"]
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else {}
    value = world * B * Ksteps / (ms * 1e-3)

    # correctness of what was timed: decisions vs the transmitted words
    fer = float((~(best_bits == msg).all(dim=1)).float().mean().item())
    crc_fail = float(1.0 - crc_ok.float().mean().item())
    tie_frac = float((flags & 1).ne(0).float().mean().item())

    # ---- e2e: host buffers through the C-ABI (pinned), copies inside the timed region --------------------
    Be = min(args.e2e_frames, B)
    h_llr = torch.empty((Be, N), dtype=torch.float32, pin_memory=True)
    h_llr.copy_(llr[:Be])
    h_bits = torch.empty((Be, K), dtype=torch.uint8, pin_memory=True)
    h_ok = torch.empty((Be,), dtype=torch.uint8, pin_memory=True)
    h_fl = torch.empty((Be,), dtype=torch.int32, pin_memory=True)

    def e2e_step():
        eng.scl_decode_host(h_llr, M, h_bits, h_ok, h_fl)

    for _ in range(W):
        e2e_step()
    # pb200_scl_decode_host is synchronous and runs on the library's own copy/compute streams, so the region is
    # timed on the host clock between two full device synchronisations (events on torch's stream would not see it)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ksteps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * Be * Ksteps / (e2e_ms * 1e-3)
    assert torch.equal(h_bits.to(dev), best_bits[:Be]), "host-buffer path and device path disagree"

    # ---- roofline -------------------------------------------------------------------------------------
    pk = peaks()
    nc = ncu_constants()              # ncu constants of this build's decode_kernel<4,7> (profiles/r02_constants.json)
    ms_kernel = ms / Ksteps            # one decode_kernel launch per step, timed with CUDA events on its stream
    ach_gbs = HBM_BYTES_PER_FRAME * B / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": ach_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_gbs / pk["hbm_gbs"],
                "traffic": (nc["dram_bytes_per_frame"] * B) if nc["dram_bytes_per_frame"] else None,
                "traffic_source": f"ncu dram__bytes per frame of a 1 Mi-frame launch x B ({nc['file']})", "peak_source": pk["source"],
                "algorithmic_bytes_per_frame": HBM_BYTES_PER_FRAME,
                "note": "decode_kernel<4,7> is SM-issue bound, not HBM bound (SURVEY 8(d)): see roofline_issue and profiles/"}
    lane_ops = ELEM_OPS[M] * B / (ms_kernel * 1e-3)
    sm_clock = (clocks.get("sm_mhz") or pk["sm_max_mhz"]) * 1e6
    props = torch.cuda.get_device_properties(dev)
    peak_max = props.multi_processor_count * 128 * pk["sm_max_mhz"] * 1e6
    wipf = nc["warp_instr_per_frame"]
    roofline_issue = {"bound": "issue", "achieved": lane_ops, "unit": "element-ops/s (W_fg + W_pm = %d per frame)" % ELEM_OPS[M],
                      "peak": peak_max, "frac": lane_ops / peak_max,
                      "peak_at_measured_clock": props.multi_processor_count * 128 * sm_clock,
                      "frac_at_measured_clock": lane_ops / (props.multi_processor_count * 128 * sm_clock),
                      "ncu_issue_slots_busy_pct": nc["issue_slots_busy_pct"], "ncu_warp_instructions_per_frame": wipf,
                      "ncu_l1_data_pipe_wavefronts_pct": nc.get("l1tex_data_pipe_wavefronts_pct"),
                      "ncu_smem_wavefronts_per_frame": nc.get("smem_wavefronts_per_frame"),
                      "ncu_source": nc["file"],
                      "issue_slot_frac_live": (wipf * B / (ms_kernel * 1e-3) / (props.multi_processor_count * 4 * sm_clock)) if wipf else None,
                      "note": "SURVEY 8(d) definition (algorithmic element-ops / lane-op peak); issue_slot_frac_live = ncu "
                              "warp-instructions per frame (profiles/) x this run's frames/s / (SMs x 4 schedulers x clock) -- "
                              "the gap between the two is per-phase list management, not idle hardware; the co-limiter is the L1 / "
                              "shared-memory data pipe (ncu_l1_data_pipe_wavefronts_pct of its peak in the same capture)"}

    # ---- e2e ceiling: bare pinned H2D copies of the same buffers in the same chunks on all ranks at once -------
    chunk = 1 << 16
    d_sink = torch.empty((chunk, N), dtype=torch.float32, device=dev)
    side = torch.cuda.Stream(device=dev)

    def h2d_only():
        with torch.cuda.stream(side):
            for lo in range(0, Be, chunk):
                n = min(chunk, Be - lo)
                d_sink[:n].copy_(h_llr[lo:lo + n], non_blocking=True)

    h2d_only(); side.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(3, Ksteps // 4)):
        h2d_only()
    side.synchronize()
    h2d_ms = (time.perf_counter() - t0) * 1e3 / max(3, Ksteps // 4)
    barrier()
    if world > 1:
        t = torch.tensor([h2d_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d_ms = float(t.item())
    h2d_gbs_per_gpu = Be * N * 4 / (h2d_ms * 1e-3) / 1e9
    e2e_ceiling = {"h2d_ceiling_gbs_per_gpu": h2d_gbs_per_gpu, "h2d_ceiling_gbs_total": h2d_gbs_per_gpu * world,
                   "h2d_ceiling_frames_per_s": world * Be / (h2d_ms * 1e-3),
                   "how": "concurrent cudaMemcpyAsync of the same pinned LLR buffer in 2^16-frame chunks on every rank, nothing else running"}

    extras = {}
    if not args.no_extras and rank == 0 and world == 1:      # informational legs only on the single-GPU run
        extras = extra_legs(eng, llr, msg, dev, B)
    # BASELINE config 5 at every N: DL-SCL M=8, 8 retries, shipped beta_M8, frames sharded over the ranks, counters
    # all-reduced (weak scaling: fixed frames per GPU), timed as the max over ranks
    scale_leg = None
    if not args.no_extras:
        scale_leg = dlscl_scale_leg(eng, dev, rank, world, barrier, min(B, 1 << 21))

    cb = cpu_baseline(args.cpu_sample) if (rank == 0 and world == 1) else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": Ksteps, "warmup": W,
            "ms_per_step": ms / Ksteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(B, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": Be * N * 4, "d2h_bytes_per_step": Be * (K + 1 + 4),
                    "rank0_cpu_affinity": numa_cpus,
                    "frames_per_step_per_gpu": Be, "ms_per_step": e2e_ms / Ksteps,
                    "call": "pb200_scl_decode_host (pinned host LLRs -> best_bits, crc_ok, flags on the host)", **e2e_ceiling},
            "gpu_launches": n_launch, "clocks": clocks, "roofline": roofline, "roofline_issue": roofline_issue,
            "kernel": eng.kernel_info(M),
            "check": {"fer_vs_sent": fer, "crc_fail_rate": crc_fail, "near_tie_flag_rate": tie_frac},
        }
        if cb:
            line["cpu_baseline"] = cb
        if extras:
            line["extras"] = extras
        if scale_leg:
            line["config5_dlscl_M8"] = scale_leg
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _golden_beta(M: int, dev):
    import torch
    g = np.load(ROOT / "tests" / "golden" / "scl_p128.npz")       # holds the reference's checkpoints/beta_M{M}.npy
    return torch.as_tensor(g[f"beta_M{M}"], device=dev)


def dlscl_scale_leg(eng, dev, rank, world, barrier, frames_per_gpu) -> dict:
    """BASELINE configs[4]: fused Philox channel + SCL M=8 + DL-SCL (8 retries, shipped beta_M8) + counters at 5.0 dB."""
    import torch
    import torch.distributed as dist
    beta = _golden_beta(8, dev)
    counters = torch.zeros(16, dtype=torch.int64, device=dev)
    nv = noise_var(5.0)

    def run(rep):
        eng.sweep(counters, M=8, noise_var=nv, n_frames=frames_per_gpu, frame_begin=(rep * world + rank) * frames_per_gpu, seed=5,
                  stream_id=50, k_payload=KP, retries=8, beta=beta)
        if world > 1:
            dist.all_reduce(counters)

    run(0); run(1)
    counters.zero_()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for r in range(reps):
        run(2 + r)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    c = counters.cpu().numpy()
    return {"frames_per_s": world * frames_per_gpu / (ms * 1e-3), "frames_per_step_per_gpu": frames_per_gpu, "snr_db": 5.0,
            "beta": "checkpoints/beta_M8.npy (tests/golden/scl_p128.npz)", "retries": 8, "ms_per_step": ms,
            "note": "counters are all-reduced every step, so with N ranks they hold N x the frames of the timed steps"}


def extra_legs(eng, llr, msg, dev, B) -> dict:
    """Other BASELINE configs, timed the same way on rank 0 (informational; not the headline)."""
    import torch
    from polar_code_b200.engine import PolarEngine, construct_info_set
    from polar_code_b200 import montecarlo as mc
    out = {}

    def time_it(fn, reps=3):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    Bx = min(B, 1 << 21)
    x = llr[:Bx]
    for m in (1, 8):                                                      # configs[1]
        ms = time_it(lambda: eng.scl_decode(x, m, want=("best_bits", "crc_ok", "flags")))
        out[f"scl_M{m}_frames_per_s"] = Bx / (ms * 1e-3)
    ms = time_it(lambda: eng.sc_decode(x))
    out["sc_frames_per_s"] = Bx / (ms * 1e-3)
    # decode_scl's FULL output set (scl.py:203-209): all M candidates, metrics and info_llrs
    xs = x[: 1 << 19]
    ms = time_it(lambda: eng.scl_decode(xs, M, want=("cand", "metrics", "info_llrs", "n_cand", "best_idx", "best_bits", "crc_ok", "flags")))
    out["scl_M4_full_outputs_frames_per_s"] = xs.shape[0] / (ms * 1e-3)
    # fused Monte-Carlo sweep (Philox channel + SCL M=4 + counters), 5.0 dB
    counters = torch.zeros(16, dtype=torch.int64, device=dev)
    ms = time_it(lambda: eng.sweep(counters, M=M, noise_var=noise_var(5.0), n_frames=Bx, seed=1, stream_id=2, k_payload=KP))
    out["fused_sweep_M4_frames_per_s"] = Bx / (ms * 1e-3)
    # configs[2] / configs[4]: DL-SCL, 8 retries, with the reference's shipped beta (and, for comparison, |L0| ranking)
    for m, snrs in ((4, (4.0, 5.0)), (8, (4.0, 5.0))):
        beta = _golden_beta(m, dev)
        for snr in snrs:
            counters.zero_()
            ms = time_it(lambda: eng.sweep(counters, M=m, noise_var=noise_var(snr), n_frames=Bx, seed=1, stream_id=3, k_payload=KP,
                                           retries=8, beta=beta))
            out[f"fused_dlscl_M{m}_r8_beta_{snr}dB_frames_per_s"] = Bx / (ms * 1e-3)
    for snr in (4.0, 5.0):
        counters.zero_()
        ms = time_it(lambda: eng.sweep(counters, M=M, noise_var=noise_var(snr), n_frames=Bx, seed=1, stream_id=3, k_payload=KP, retries=8))
        out[f"fused_dlscl_M4_r8_nobeta_{snr}dB_frames_per_s"] = Bx / (ms * 1e-3)
    # configs[3]: NR rate-matched sweep, A(128,88) = 64 payload + CRC-24, E = 256, SCL M=4, Eb/N0 3.0 dB
    nr = PolarEngine(N, construct_info_set(N, 88), CRC24, device=dev.index)
    nr.set_rate_matching(256)
    counters.zero_()
    ms = time_it(lambda: nr.sweep(counters, M=M, noise_var=mc.ber_noise_var(3.0, 64, 256), n_frames=Bx, seed=1, stream_id=4, k_payload=64,
                                  frame_error_mode=1, bit_error_span=64))
    out["fused_nr_E256_K88_M4_frames_per_s"] = Bx / (ms * 1e-3)
    # OPTIONAL binary16 ingest of the host-buffer path (pb200_scl_decode_host_f16: half the host->device bytes, rows widened
    # exactly on load).  Not the headline: quantising LLRs to binary16 is the caller's decision (include/polar_b200.h).
    Bh = min(B, 1 << 20)
    h16 = torch.empty((Bh, N), dtype=torch.float16, pin_memory=True)
    h16.copy_(llr[:Bh].half())
    h_bits = torch.empty((Bh, K), dtype=torch.uint8, pin_memory=True)
    h_ok = torch.empty((Bh,), dtype=torch.uint8, pin_memory=True)
    h_fl = torch.empty((Bh,), dtype=torch.int32, pin_memory=True)
    for _ in range(2):
        eng.scl_decode_host(h16, M, h_bits, h_ok, h_fl)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        eng.scl_decode_host(h16, M, h_bits, h_ok, h_fl)
    torch.cuda.synchronize()
    out["e2e_f16_ingest_frames_per_s"] = 5 * Bh / (time.perf_counter() - t0)
    return out


if __name__ == "__main__":
    main()
