"""N > 1 host logic on CPU: world_size-2 gloo processes shard the candidates, evaluate their shard
(with the oracle standing in for the GPU evaluator), broadcast changed blocks and gather the per-rank
best records; the result must equal the single-process result, including the tie-break."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _score(gain, yaw, parent_yaw, distance, v_max=1.0, dyaw_max=0.5):
    import math
    dy = yaw - parent_yaw
    if dy > math.pi:
        dy -= 2.0 * math.pi
    if dy < -math.pi:
        dy += 2.0 * math.pi
    return gain / max(dy / dyaw_max, distance / v_max)


def _local_best(gains, yaws, parent_yaw, distance, free, zero_gain=0.0):
    best, bi = zero_gain, -1
    for i in range(len(gains)):
        if not free[i]:
            continue
        s = _score(gains[i], yaws[i], parent_yaw[i], distance[i])
        if s > best:
            best, bi = s, i
    if bi < 0:
        return torch.tensor([-1.0, zero_gain, 0.0, 0.0], dtype=torch.float64)
    return torch.tensor([float(bi), best, gains[bi], yaws[bi]], dtype=torch.float64)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nbv_exploration_planner_b200 import dist as nd
    from nbv_exploration_planner_b200 import synth
    from oracle.binding import Oracle

    cfg = synth.CONFIGS["tiny"]
    m = synth.make_map("tiny") if rank == 0 else None
    nvox = cfg.vps ** 3
    dev = torch.device("cpu")
    # map replication: rank 0 broadcasts keys + payload, every rank applies them to its replica
    n_t, dk, dp = nd.broadcast_blocks(m.keys if rank == 0 else None, m.tsdf if rank == 0 else None, 2 * nvox, dev)
    n_e, dk2, dp2 = nd.broadcast_blocks(m.keys if rank == 0 else None, m.esdf if rank == 0 else None, nvox, dev)
    o = Oracle()
    lt = o.layer_create(0, cfg.voxel_size, cfg.vps)
    le = o.layer_create(1, cfg.voxel_size, cfg.vps)
    o.layer_update(lt, dk.numpy(), dp.numpy())
    o.layer_update(le, dk2.numpy(), dp2.numpy())
    o.set_camera(synth.camera_for(cfg).as_dict())
    # identical candidate list on every rank (seeded), contiguous shards
    c = synth.make_candidates(cfg, 23, None)
    rng = np.random.default_rng(4)
    parent_yaw = rng.uniform(-3, 3, 23)
    lo, hi = nd.shard_bounds(23, world, rank)
    r = o.gain_eval(lt, c.positions[lo:hi])
    free = o.check_segments(le, cfg.robot_radius, c.segments[lo:hi])["free"]
    rec = _local_best(r["gain"], r["yaw"], parent_yaw[lo:hi], c.distance[lo:hi], free)
    best = nd.gather_best(rec, lo)
    # a second round with an artificial exact tie across ranks: lowest global index must win
    tie = torch.tensor([0.0, 5.0, 1.0, 0.5], dtype=torch.float64)
    best_tie = nd.gather_best(tie, lo)
    none = nd.gather_best(torch.tensor([-1.0, 0.0, 0.0, 0.0], dtype=torch.float64), lo)
    # raw 32-byte nbvgpu_best records (the form the GPU path gathers): same winner
    raw = torch.zeros(32, dtype=torch.uint8)
    raw.view(torch.int64)[0] = int(rec[0])
    raw.view(torch.float64)[1:] = rec[1:]
    best_raw = nd.gather_best_raw(raw, 23)
    assert best_raw == best, (best_raw, best)
    q.put((rank, lo, hi, n_t, n_e, best, best_tie, none, o.num_blocks(lt)))
    dist.destroy_process_group()


def test_world_size_2_matches_single_process(oracle_mod):
    from nbv_exploration_planner_b200 import dist as nd
    from nbv_exploration_planner_b200 import synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    m = synth.make_map("tiny")
    cfg = m.cfg
    o = oracle_mod.Oracle()
    lt = o.layer_create(0, cfg.voxel_size, cfg.vps)
    le = o.layer_create(1, cfg.voxel_size, cfg.vps)
    o.layer_update(lt, m.keys, m.tsdf)
    o.layer_update(le, m.keys, m.esdf)
    o.set_camera(m.camera().as_dict())
    c = synth.make_candidates(cfg, 23, None)
    parent_yaw = np.random.default_rng(4).uniform(-3, 3, 23)
    r = o.gain_eval(lt, c.positions)
    free = o.check_segments(le, cfg.robot_radius, c.segments)["free"]
    want = _local_best(r["gain"], r["yaw"], parent_yaw, c.distance, free)
    assert [x[1:3] for x in res] == [(0, 12), (12, 23)]
    for rank, lo, hi, n_t, n_e, best, best_tie, none, nblocks in res:
        assert n_t == len(m.keys) and n_e == len(m.keys) and nblocks == len(m.keys)
        assert best["index"] == int(want[0]) and best["score"] == float(want[1])
        assert best["gain"] == float(want[2]) and best["yaw"] == float(want[3])
        assert best_tie["index"] == 0          # tie -> lowest global index (rank 0's candidate 0)
        assert none["index"] == -1
    assert int(want[0]) >= 0


def _worker_sharded(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nbv_exploration_planner_b200 import dist as nd
    from nbv_exploration_planner_b200 import synth
    from oracle.binding import Oracle
    m = synth.make_map("tiny")          # every rank holds a replica (the broadcast itself is covered above)
    cfg = m.cfg
    o = Oracle()
    lt = o.layer_create(0, cfg.voxel_size, cfg.vps)
    le = o.layer_create(1, cfg.voxel_size, cfg.vps)
    o.layer_update(lt, m.keys, m.tsdf)
    o.layer_update_esdf_observed(le, m.keys, m.esdf, (m.tsdf[..., 1] > 0).astype(np.uint8))
    o.set_camera(m.camera().as_dict())
    pts = synth.make_candidates(cfg, 11, None).positions      # 11 over 2 ranks: uneven shards (6 + 5)
    lo, hi = nd.shard_bounds(len(pts), world, rank)
    f = o.frontier_potential(lt, pts[lo:hi], 6.0)
    g = o.distance_and_gradient(le, pts[lo:hi])
    gain = nd.gather_sharded(torch.from_numpy(f["gain"]), len(pts))
    grad = nd.gather_sharded(torch.from_numpy(g["gradient"]), len(pts))          # a [n][3] result
    ok = nd.gather_sharded(torch.from_numpy(g["ok"]), len(pts))
    q.put((rank, gain.numpy(), grad.numpy(), ok.numpy()))
    dist.destroy_process_group()


def test_sharded_vertices_and_queries_gather_in_batch_order(oracle_mod):
    """The widened rows shard like the candidates: world_size 2 over gloo, uneven shards, results back in batch order
    on every rank and equal to the single-process evaluation (the oracle stands in for the GPU evaluator)."""
    from nbv_exploration_planner_b200 import synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = synth.make_map("tiny")
    o = oracle_mod.Oracle()
    lt = o.layer_create(0, m.voxel_size, m.vps)
    le = o.layer_create(1, m.voxel_size, m.vps)
    o.layer_update(lt, m.keys, m.tsdf)
    o.layer_update_esdf_observed(le, m.keys, m.esdf, (m.tsdf[..., 1] > 0).astype(np.uint8))
    o.set_camera(m.camera().as_dict())
    pts = synth.make_candidates(m.cfg, 11, None).positions
    want_f = o.frontier_potential(lt, pts, 6.0)["gain"]
    want_g = o.distance_and_gradient(le, pts)
    for rank, gain, grad, ok in res:
        assert np.array_equal(gain, want_f) and np.array_equal(grad, want_g["gradient"]) and np.array_equal(ok, want_g["ok"])
    assert want_f.min() > 0 and len(set(want_f.tolist())) > 3 and np.ptp(want_g["gradient"], 0).max() > 0.1   # not a trivial gather


def _worker_pull(rank, world, port, q, payload_bytes):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nbv_exploration_planner_b200 import dist as nd
    # the "shared arena": every rank can read the same bytes (here: generates them), each copies only its slice
    src = torch.from_numpy(np.random.default_rng(5).integers(0, 256, payload_bytes, dtype=np.uint8))
    lo, hi, sl = nd.pull_slice(payload_bytes, world, rank)
    stage = torch.zeros(world * sl, dtype=torch.uint8)
    stage[rank * sl:rank * sl + (hi - lo)] = src[lo:hi]                     # the rank's own "PCIe pull"
    parts = [torch.empty(sl, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(parts, stage[rank * sl:(rank + 1) * sl].clone())        # the all-gather that completes the chunk
    q.put((rank, bool(torch.equal(torch.cat(parts)[:payload_bytes], src)), lo, hi, sl))
    dist.destroy_process_group()


@pytest.mark.parametrize("payload_bytes", [1, 255, 4096 * 12 * 3 + 17, 1 << 20])
def test_pulled_chunk_slices_reassemble_the_payload(payload_bytes):
    """A pulled commit chunk (every rank copies its slice, an all-gather completes it): world_size 2 over gloo rebuilds the
    payload byte for byte with the library's slicing rule, including payloads smaller than a slice and ragged tails."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_pull, args=(r, world, port, q, payload_bytes)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, *_ in res)
    assert res[0][2] == 0 and res[-1][3] == payload_bytes and res[0][4] % 256 == 0


def test_pull_slices_cover_the_payload():
    from nbv_exploration_planner_b200.dist import pull_slice
    for n in (0, 1, 256, 257, 7043678208, (64 << 20) + 3):
        for w in (1, 2, 4, 8):
            spans = [pull_slice(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] or (a[1] == n and b[0] == n) for a, b in zip(spans, spans[1:]))
            assert all(hi - lo <= sl and sl % 256 == 0 and sl * w >= n for lo, hi, sl in spans)


def test_shard_bounds_cover_everything():
    from nbv_exploration_planner_b200.dist import shard_bounds
    for n in (0, 1, 7, 1000, 10001):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= (n + w - 1) // w


def test_single_process_helpers_without_process_group():
    from nbv_exploration_planner_b200 import dist as nd
    keys = np.array([[1, 2, 3], [-4, 5, -6]], np.int32)
    pay = np.arange(2 * 8, dtype=np.float32).reshape(2, 8)
    n, dk, dp = nd.broadcast_blocks(keys, pay, 8, torch.device("cpu"))
    assert n == 2 and np.array_equal(dk.numpy(), keys) and np.array_equal(dp.numpy(), pay)
    b = nd.gather_best(torch.tensor([3.0, 2.5, 1.0, 0.1], dtype=torch.float64), 100)
    assert b["index"] == 103 and b["score"] == 2.5
    raw = torch.zeros(32, dtype=torch.uint8)
    raw.view(torch.int64)[0] = 7
    raw.view(torch.float64)[1:] = torch.tensor([1.5, 2.5, 3.5], dtype=torch.float64)
    f = nd.best_record_as_f64(raw)
    assert f.tolist() == [7.0, 1.5, 2.5, 3.5]
