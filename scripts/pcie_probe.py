"""How fast can N GPUs of one box pull one host buffer together?  (sizing the map replication of nbvgpu_commit)
   torchrun --nproc-per-node N scripts/pcie_probe.py
Each rank maps the same POSIX shared-memory file, page-locks its mapping and copies (a) the whole buffer, (b) its 1/N slice to its
GPU, all ranks at once; prints per-rank and aggregate GB/s, and the NCCL all-gather time that would complete case (b)."""
import mmap
import os
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 2 << 30
path = "/dev/shm/nbv_pcie_probe"
if rank == 0:
    with open(path, "wb") as f:
        f.truncate(N)
if world > 1:
    dist.barrier()
fd = os.open(path, os.O_RDWR)
mm = mmap.mmap(fd, N)
a = np.frombuffer(mm, dtype=np.uint8)
if rank == 0:
    a[:] = 1
rc = int(torch.cuda.cudart().cudaHostRegister(a.ctypes.data, N, 1))   # portable
h = torch.from_numpy(a)
d = torch.empty(N, dtype=torch.uint8, device=dev)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        best = min(best, time.perf_counter() - t)
    return best


def whole():
    d.copy_(h, non_blocking=True)


S = N // world


def part():
    d[rank * S:(rank + 1) * S].copy_(h[rank * S:(rank + 1) * S], non_blocking=True)


def gather():
    if world > 1:
        dist.all_gather_into_tensor(d, d[rank * S:(rank + 1) * S].clone() if False else d[rank * S:(rank + 1) * S])


t_whole, t_part = timed(whole), timed(part)
t_ag = timed(gather) if world > 1 else 0.0
vals = torch.tensor([t_whole, t_part, t_ag], dtype=torch.float64, device=dev)
if world > 1:
    allv = torch.empty((world, 3), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allv, vals)
else:
    allv = vals[None]
if rank == 0:
    tw, tp, ta = allv[:, 0].max().item(), allv[:, 1].max().item(), allv[:, 2].max().item()
    print(f"register rc {rc}; {world} GPU(s), {N / 2**30:.0f} GiB in one shared-memory buffer")
    print(f"every GPU copies the whole buffer : {tw * 1e3:8.1f} ms  = {N / tw / 1e9:6.1f} GB/s per GPU, {world * N / tw / 1e9:6.1f} GB/s aggregate")
    print(f"every GPU copies its 1/{world} slice    : {tp * 1e3:8.1f} ms  = {N / tp / 1e9:6.1f} GB/s aggregate")
    print(f"NCCL all-gather of the slices     : {ta * 1e3:8.1f} ms  -> slice pull + all-gather {1e3 * (tp + ta):.1f} ms against {tw * 1e3:.1f} ms (one link + broadcast)")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    os.unlink(path)
