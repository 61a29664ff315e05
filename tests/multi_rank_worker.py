"""One rank of a one-process-per-GPU context (nbvgpu_create_rank), driven by tests/test_gpu_multi.py:
rank 0 stages the map, every rank joins the commit (NCCL broadcast inside the library), every rank evaluates the
whole batch through the host-buffer entry point (its shard runs on its GPU, the best node of the whole batch comes
back through the shared block) and its shard again from device memory.  Results go to <out>.<rank>.npz."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("NBV_WORKER_TIMEOUT", "300")), exit=True)   # a rank that waits for a peer for ever says where, and goes away
    rank, world, idfile, out, n_cand = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5])
    import torch

    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, synth
    torch.cuda.set_device(rank)
    if rank == 0:
        uid = NbvGpu.unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.01)
        uid = open(idfile, "rb").read()
    g = NbvGpu.rank(rank, rank, world, uid)
    info = g.group_info()
    assert info == {"rank": rank, "world": world, "local_gpus": 1}, info
    m = synth.make_map("cfg1")       # every rank can build it (seeded); only rank 0 uploads it
    cfg = synth.CONFIGS["cfg2"]
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 64)
    le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 64)
    if rank == 0:
        half = len(m.keys) // 2
        g.layer_update(lt, m.keys[:half], m.tsdf[:half])
        g.layer_update(le, m.keys, m.esdf)
        g.commit()                      # first commit: half the TSDF, all the ESDF
        g.layer_update(lt, m.keys[half:], m.tsdf[half:])
        g.layer_remove(lt, m.keys[:3])  # ... then the rest, with a removal and its re-insertion in between
        g.commit()
        g.layer_update(lt, m.keys[:3], m.tsdf[:3])
        g.commit()
    else:
        g.commit(); g.commit(); g.commit()
    assert g.layer_num_blocks(lt) == len(m.keys) and g.layer_num_blocks(le) == len(m.keys)
    g.set_camera(m.camera())
    c = synth.make_candidates(cfg, n_cand, None)     # unfiltered: some edges collide
    pyaw = np.zeros(n_cand)
    r = g.evaluate_candidates(lt, le, cfg.robot_radius, c.positions, c.segments, pyaw, c.distance, 1.0, 0.5, 0.0)
    per = (n_cand + world - 1) // world
    lo, hi = min(n_cand, rank * per), min(n_cand, rank * per + per)
    # the same shard from device memory
    dev = torch.device("cuda", rank)
    d_pos = torch.from_numpy(c.positions[lo:hi].copy()).to(dev)
    d_seg = torch.from_numpy(c.segments[lo:hi].copy()).to(dev)
    d_py = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
    d_di = torch.from_numpy(c.distance[lo:hi].copy()).to(dev)
    d_free = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
    d_gain = torch.empty(hi - lo, dtype=torch.float64, device=dev)
    d_yaw = torch.empty(hi - lo, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    b2 = None
    for _ in range(3):       # repeated: the slots alternate
        b2 = g.evaluate_shard_device(lt, le, cfg.robot_radius, hi - lo, lo, d_pos, d_seg, d_py, d_di, 1.0, 0.5, 0.0, d_free, d_gain, d_yaw)
    g.synchronize()
    b3 = g.gain_eval_best(lt, c.positions, pyaw, c.distance, r["free"] if world == 1 else None, 1.0, 0.5, 0.0)
    # the whole map again, this time out of memory every rank has mapped (nbvgpu_shared_alloc: collective): each GPU pulls its
    # slice of every chunk over its own PCIe link, an all-gather completes the chunk (forced for these small chunks)
    buf = g.shared_alloc(m.tsdf.nbytes + m.esdf.nbytes)
    g.layer_clear(lt)
    g.layer_clear(le)
    if rank == 0:
        t = buf[:m.tsdf.nbytes].view(m.tsdf.dtype).reshape(m.tsdf.shape)
        e = buf[m.tsdf.nbytes:].view(m.esdf.dtype).reshape(m.esdf.shape)
        t[...] = m.tsdf
        e[...] = m.esdf
        g.layer_update_pinned(lt, m.keys, t)
        g.layer_update_pinned(le, m.keys, e)
    g.commit()
    assert g.layer_num_blocks(lt) == len(m.keys) and g.layer_num_blocks(le) == len(m.keys)
    rp = g.evaluate_candidates(lt, le, cfg.robot_radius, c.positions, c.segments, pyaw, c.distance, 1.0, 0.5, 0.0)
    pulled = g.pulled_chunks()
    g.shared_free(buf)
    np.savez(out + f".{rank}.npz", lo=lo, hi=hi, gain=r["gain"][lo:hi], yaw=r["yaw_deg"][lo:hi], free=r["free"][lo:hi],
             best=np.array([r["index"], r["score"], r["best_gain"], r["best_yaw"]]),
             best_dev=np.array([b2["index"], b2["score"], b2["gain"], b2["yaw"]]),
             best_nomask=np.array([b3["index"], b3["score"], b3["best_gain"], b3["best_yaw"]]),
             d_gain=d_gain.cpu().numpy(), d_yaw=d_yaw.cpu().numpy(), d_free=d_free.cpu().numpy(),
             gain_pulled=rp["gain"][lo:hi], free_pulled=rp["free"][lo:hi], pulled_chunks=pulled)
    g.close()


if __name__ == "__main__":
    main()
