"""Which shared-memory sizes can N ranks page-lock at once?  torchrun --nproc-per-node N scripts/shm_register_probe.py
Per size: rank 0 creates the /dev/shm file (optionally touching every page first), every rank maps it and calls
cudaHostRegister(portable); prints each rank's return code."""
import mmap
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rt = torch.cuda.cudart()
for gib, populate, stagger in [(2, False, False), (4, False, False), (7, False, False), (7, True, False), (7, True, True), (7, False, True)]:
    n = int(gib * (1 << 30))
    path = f"/dev/shm/nbv_reg_probe_{gib}_{int(populate)}_{int(stagger)}"
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(n)
        if populate:
            fd0 = os.open(path, os.O_RDWR)
            mm0 = mmap.mmap(fd0, n)
            np.frombuffer(mm0, dtype=np.uint8)[:] = 0
    dist.barrier()
    fd = os.open(path, os.O_RDWR)
    mm = mmap.mmap(fd, n)
    a = np.frombuffer(mm, dtype=np.uint8)
    rc = -1
    t0 = time.perf_counter()
    for turn in range(world if stagger else 1):
        if not stagger or turn == rank:
            rc = int(rt.cudaHostRegister(a.ctypes.data, n, 1))
        if stagger:
            dist.barrier()
    dt = time.perf_counter() - t0
    res = torch.tensor([rc, int(dt * 1e3)], dtype=torch.int64, device=dev)
    allr = torch.empty((world, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allr, res)
    if rank == 0:
        print(f"{gib} GiB populate={populate} one-rank-at-a-time={stagger}: return codes {allr[:, 0].tolist()} ms {allr[:, 1].tolist()}", flush=True)
    if rc == 0:
        rt.cudaHostUnregister(a.ctypes.data)
    dist.barrier()
    del a
    mm.close()
    os.close(fd)
    if rank == 0:
        os.unlink(path)
    dist.barrier()
dist.destroy_process_group()
