"""How fast can this box take frames off its GPUs?  Pinned D2H copies of one 4K RGBA8 frame (33 MB), every rank
alone and then all ranks at once (torchrun --nproc-per-node N tools/d2h_probe.py).  Explains the e2e numbers of
bench.py at N > 1: the copy-out is bound by the host side of PCIe, not by the kernels."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
nbytes = 3840 * 2160 * 4
src = torch.zeros(nbytes, dtype=torch.uint8, device=f"cuda:{local}")
dst = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(3)]


def run(reps: int) -> float:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        dst[i % 3].copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


run(10)
for r in range(world):           # one rank at a time
    if world > 1:
        dist.barrier()
    if r == rank:
        print(f"rank {rank} alone: {run(200):.1f} GB/s", flush=True)
if world > 1:
    dist.barrier()
    g = run(200)
    t = torch.tensor([g], device=f"cuda:{local}")
    dist.all_reduce(t)
    if rank == 0:
        print(f"all {world} ranks at once: {t.item():.1f} GB/s total ({t.item() / world:.1f} per GPU)", flush=True)
    dist.destroy_process_group()
