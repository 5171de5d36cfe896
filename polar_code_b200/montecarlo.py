"""Monte-Carlo drivers on top of PolarEngine.sweep: rank sharding, counter all-reduce, adaptive stop.

Frames of one SNR point are numbered globally (0, 1, 2, ...); frame f always draws the same Philox numbers,
so every result below is independent of the world size and of the chunking.  The only cross-rank traffic is
the int64 counter block: one all-reduce per SNR point (NCCL over NVLink on GPUs; gloo in the
CPU tests of this host logic).
"""

from __future__ import annotations

import math
import time
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

COUNTER_NAMES = ("frames", "scl_frame_errors", "scl_bit_errors", "dl_frame_errors", "dl_bit_errors",
                 "uncoded_frame_errors", "uncoded_bit_errors", "dl_attempts_minus_1", "near_tie_frames",
                 "scl_undetected", "dl_undetected", "rank_tie_frames")
NCOUNTERS = 16


def world() -> Tuple[int, int]:
    """(rank, world_size) of the initialised process group, (0, 1) otherwise."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def bind_to_gpu_numa(device_index: int) -> Optional[str]:
    """Pin this process to the CPUs next to its GPU (sysfs `local_cpulist` of the GPU's PCI function), so that the
    pinned host buffers it allocates afterwards are first-touched on that NUMA node and every rank's host<->device
    copies stay on its own socket.  Returns the cpulist applied, or None when sysfs has nothing to say
    (single-socket box, container without the PCI tree)."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        text = open(path).read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return text
    except Exception:
        return None


def maybe_init_distributed() -> Tuple[int, int]:
    """Join the torchrun rendezvous if this process was launched by it (one process per GPU, NCCL)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            bind_to_gpu_numa(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    return world()


def shutdown_distributed(started_here: bool) -> None:
    """Leave the process group again if maybe_init_distributed() created it."""
    if started_here and dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def shard_range(n_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, begin+count) slice of range(n_frames) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_frames), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def reduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """Sum the int64 counter block over ranks, in place (the decode kernels accumulate straight into it)."""
    if world()[1] > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def fer_noise_var(snr_db: float, K: int, N: int) -> float:
    """run_fer_sweep.py:62-64 -- the rate counts the CRC bits as information."""
    return 1.0 / (2.0 * (K / N) * 10 ** (snr_db / 10.0))


def ber_noise_var(ebn0_db: float, payload_bits: int, coded_bits: int) -> float:
    """run_ber_sweep.py:105-109 -- the CRC bits are overhead."""
    return 1.0 / (2.0 * (10 ** (ebn0_db / 10.0)) * (payload_bits / coded_bits))


def warm_up(engine, *, M: int, retries: int, beta=None, k_payload: Optional[int] = None) -> float:
    """First-launch set-up, paid once per process and configuration BEFORE the first timed point: kernel attribute /
    occupancy queries, scratch and queue allocation, the L2 set-aside, module load of the kernels.  Runs one tiny
    sweep (64 frames of a private Philox stream, counters discarded) and returns the seconds it took, so the CLIs can
    report start-up separately from the steady-state frames/s (config 1 spends most of its
    5 s wall time here)."""
    t0 = time.perf_counter()
    scratch = torch.zeros(NCOUNTERS, dtype=torch.int64, device=engine.dev)
    kp = engine.K if k_payload is None else k_payload
    engine.sweep(scratch, M=M, noise_var=1.0, n_frames=64, frame_begin=0, seed=0x5EED, stream_id=0xFFFF, retries=retries,
                 run_scl=True, k_payload=kp, frame_error_mode=0, bit_error_span=engine.K, beta=beta)
    torch.cuda.synchronize(engine.dev)
    return time.perf_counter() - t0


def fer_point(engine, *, M: int, snr_db: float, frames: int, seed: int, retries: int, beta=None,
              include_uncoded: bool = False, k_payload: Optional[int] = None, stream_id: Optional[int] = None) -> np.ndarray:
    """One SNR point of run_fer_sweep.py:60-121 -> summed int64 counters (see COUNTER_NAMES)."""
    rank, ws = world()
    begin, count = shard_range(frames, rank, ws)
    counters = torch.zeros(NCOUNTERS, dtype=torch.int64, device=engine.dev)
    kp = engine.K if k_payload is None else k_payload
    sid = int(round(snr_db * 10)) if stream_id is None else stream_id
    if count > 0:
        engine.sweep(counters, M=M, noise_var=fer_noise_var(snr_db, engine.K, engine.N), n_frames=count, frame_begin=begin,
                     seed=seed, stream_id=sid, retries=retries, run_scl=True, k_payload=kp, frame_error_mode=0,
                     bit_error_span=engine.K, include_uncoded=include_uncoded,
                     noise_var_uncoded=1.0 / (2.0 * 10 ** (snr_db / 10.0)), beta=beta)
    reduce_counters(counters)
    return counters.cpu().numpy()


@dataclass
class CutState:
    """Running totals of the adaptive loop of run_ber_sweep.py:127 (SimulationStats :36-62)."""
    frames: int = 0
    bit_errors: int = 0
    frame_errors: int = 0
    work_sum: int = 0
    done: bool = False


def adaptive_cut(state: CutState, local_err: torch.Tensor, local_work: torch.Tensor, local_begin: int, chunk_begin: int,
                 chunk_frames: int, payload_len: int, err_cap: int, bits_cap: float) -> CutState:
    """Fold one chunk of per-frame results into `state`, cutting at the first global frame index where the
    sequential loop `while bit_errors < err_cap and bits_total < bits_cap` would have stopped.

    local_err / local_work: this rank's per-frame bit errors and attempts-1 for the contiguous frames
    [local_begin, local_begin + len) of the chunk [chunk_begin, chunk_begin + chunk_frames).  Works on CPU
    tensors with gloo and on CUDA tensors with NCCL."""
    dev = local_err.device
    rank, ws = world()
    err = local_err.to(torch.int64) & 0xFFFF          # the kernels write u16 (torch buffers are int16)
    csum = torch.cumsum(err, 0)
    local_total = csum[-1:].clone() if err.numel() else torch.zeros(1, dtype=torch.int64, device=dev)
    if ws > 1:
        totals = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ws)]
        dist.all_gather(totals, local_total)
        before = int(sum(int(t.item()) for t in totals[:rank]))
    else:
        before = 0
    # frames allowed by bits_cap: the loop runs while frames*payload_len < bits_cap
    max_frames = int(math.ceil(bits_cap / payload_len))
    big = chunk_begin + chunk_frames          # "no cut in this chunk"
    cut_last = big                            # global index of the LAST frame that is counted
    if err.numel():
        hit = (state.bit_errors + before + csum) >= err_cap
        if bool(hit.any()):
            cut_last = local_begin + int(torch.nonzero(hit)[0].item())
    if max_frames - 1 < big:
        cut_last = min(cut_last, max_frames - 1)
    t = torch.tensor([cut_last], dtype=torch.int64, device=dev)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    cut_last = int(t.item())
    stop = cut_last < big
    n_take = max(0, min(err.numel(), cut_last + 1 - local_begin))   # frames of this rank that count
    part = torch.zeros(4, dtype=torch.int64, device=dev)
    if n_take > 0:
        e = err[:n_take]
        part[0] = n_take
        part[1] = e.sum()
        part[2] = (e > 0).sum()
        part[3] = (local_work[:n_take].to(torch.int64) & 0xFFFF).sum()
    if ws > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
    p = part.cpu().numpy()
    return CutState(frames=state.frames + int(p[0]), bit_errors=state.bit_errors + int(p[1]),
                    frame_errors=state.frame_errors + int(p[2]), work_sum=state.work_sum + int(p[3]), done=stop)


def ber_point(engine, *, M: int, ebn0_db: float, payload_len: int, coded_len: int, seed: int, stream_id: int,
              err_cap: int, bits_cap: float, retries: int = -1, beta=None, first_chunk: int = 1 << 14,
              max_chunk: int = 1 << 22) -> CutState:
    """One Eb/N0 point of run_ber_sweep.py:112-181 with the sequential stopping rule reproduced exactly
    (in global frame order) on batched GPU chunks."""
    state = CutState()
    if not (state.bit_errors < err_cap and 0 < bits_cap):
        return state
    nv = ber_noise_var(ebn0_db, payload_len, coded_len)
    max_frames = int(math.ceil(bits_cap / payload_len))
    rank, ws = world()
    chunk_begin, chunk = 0, max(1, min(first_chunk, max_frames))
    while not state.done and chunk_begin < max_frames:
        chunk = min(chunk, max_frames - chunk_begin)
        lb, ln = shard_range(chunk, rank, ws)
        err = torch.zeros(ln, dtype=torch.int16, device=engine.dev)     # u16 on the device side: exact per-frame counts
        work = torch.zeros(ln, dtype=torch.int16, device=engine.dev)
        counters = torch.zeros(NCOUNTERS, dtype=torch.int64, device=engine.dev)
        if ln > 0:
            engine.sweep(counters, M=M, noise_var=nv, n_frames=ln, frame_begin=chunk_begin + lb, seed=seed,
                         stream_id=stream_id, retries=retries, run_scl=(retries < 0), k_payload=payload_len,
                         frame_error_mode=1, bit_error_span=payload_len, beta=beta, frame_bit_errors=err, frame_work=work)
        state = adaptive_cut(state, err, work, chunk_begin + lb, chunk_begin, chunk, payload_len, err_cap, bits_cap)
        chunk_begin += chunk
        chunk = min(chunk * 2, max_chunk)
    state.done = True
    return state
