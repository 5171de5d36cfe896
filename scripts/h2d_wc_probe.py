"""Bare pinned H2D bandwidth: default pinned vs write-combined pinned host memory (one process per GPU under torchrun).
    python scripts/h2d_wc_probe.py            |  torchrun --nproc-per-node 8 scripts/h2d_wc_probe.py"""
import ctypes, glob, os, sys, time
import torch
rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = ctypes.CDLL(cands[0] if cands else "libcudart.so")
BYTES = 1 << 30; CHUNK = 1 << 25
dst = torch.empty(CHUNK, dtype=torch.uint8, device="cuda")
def bench(flags, tag):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(BYTES), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 1, BYTES)
    st = torch.cuda.Stream()
    def run():
        for off in range(0, BYTES, CHUNK):
            assert rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(p.value + off), ctypes.c_size_t(CHUNK), 1, ctypes.c_void_p(st.cuda_stream)) == 0
    run(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): run()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    gbs = 5 * BYTES / dt / 1e9
    if world > 1:
        t = torch.tensor([gbs], device="cuda"); dist.all_reduce(t); tot = float(t.item())
    else: tot = gbs
    if rank == 0: print(f"{tag}: {gbs:.1f} GB/s on rank 0, {tot:.1f} GB/s total over {world} rank(s)", flush=True)
    rt.cudaFreeHost(p)
bench(0, "pinned default      ")
bench(4, "pinned write-combined")
bench(0, "pinned default (2)  ")
bench(4, "pinned write-combined (2)")
