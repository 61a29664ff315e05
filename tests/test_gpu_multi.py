"""Multi-GPU contexts behind the C ABI: N-GPU == 1-GPU, bit for bit.

  * nbvgpu_create_multi with ONE device and nbvgpu_create_rank with world = 1 run everywhere (the whole group machinery --
    chunked commit, shard bookkeeping, best records through the shared block -- without a second GPU);
  * with >= 2 GPUs: one process driving all of them (Python binding and the C++ shim), and one process per GPU
    (tests/multi_rank_worker.py), against a plain single-GPU context and the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_gpu, load_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _load(g, m, chunks=1):
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 64)
    le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 64)
    step = (len(m.keys) + chunks - 1) // chunks
    for i in range(0, len(m.keys), step):
        g.layer_update(lt, m.keys[i:i + step], m.tsdf[i:i + step])
        g.layer_update(le, m.keys[i:i + step], m.esdf[i:i + step])
    g.commit()
    g.set_camera(m.camera())
    return lt, le


def _reference_results(m, cfg, c):
    """Plain single-GPU context: the two-call form, verdicts first."""
    g, lt, le = load_gpu(m, m.camera())
    free = g.check_segments(le, cfg.robot_radius, c.segments)
    r = g.gain_eval_best(lt, c.positions, np.zeros(len(c.positions)), c.distance, free, 1.0, 0.5, 0.0)
    g.close()
    return free, r


@pytest.mark.parametrize("devices", [[0], "all"])
def test_one_process_all_gpus_equals_single_gpu(oracle_mod, cfg1_map, devices, monkeypatch):
    from nbv_exploration_planner_b200 import NbvGpu, synth
    if devices == "all":
        if _n_gpus() < 2:
            pytest.skip("needs >= 2 GPUs")
        devices = list(range(_n_gpus()))
    monkeypatch.setenv("NBVGPU_CHUNK_MB", "1")     # 1 MiB chunks: the commit runs its pipelined multi-chunk path (243 blocks = 11 MB)
    m, cfg = cfg1_map, synth.CONFIGS["cfg2"]
    c = synth.make_candidates(cfg, 1000, None)     # unfiltered samples: part of the edges collide
    free1, r1 = _reference_results(m, cfg, c)
    assert 0 < free1.sum() < len(free1)
    g = NbvGpu.multi(devices)
    assert g.group_info() == {"rank": 0, "world": len(devices), "local_gpus": len(devices)}
    lt, le = _load(g, m, chunks=3)
    assert g.layer_num_blocks(lt) == len(m.keys)
    r = g.evaluate_candidates(lt, le, cfg.robot_radius, c.positions, c.segments, np.zeros(1000), c.distance, 1.0, 0.5, 0.0)
    assert np.array_equal(r["free"], free1)
    assert np.array_equal(r["gain"], r1["gain"]) and np.array_equal(r["yaw_deg"], r1["yaw_deg"])
    assert (r["index"], r["score"], r["best_gain"], r["best_yaw"]) == (r1["index"], r1["score"], r1["best_gain"], r1["best_yaw"])
    assert np.all(r["gain"][free1 == 0] == 0)      # rrt.cpp:210-216: no gain behind a colliding edge
    # best record only (no per-candidate outputs): the no-copy path
    from nbv_exploration_planner_b200 import _native as N
    import ctypes as C
    best = N.Best()
    pos, seg, py, di = (np.ascontiguousarray(a, np.float64) for a in (c.positions, c.segments, np.zeros(1000), c.distance))
    for _ in range(3):
        rc = g._L.nbvgpu_evaluate_candidates(g._h, lt, le, float(cfg.robot_radius), 1000, pos.ctypes.data, seg.ctypes.data, py.ctypes.data,
                                             di.ctypes.data, 1.0, 0.5, 0.0, None, C.byref(best), None, None)
        assert rc == 0 and (best.index, best.score) == (r1["index"], r1["score"])
    # the sharded batch entry points without a best record
    rg = g.gain_eval(lt, c.positions, per_yaw=True)
    o, lt_o, le_o = load_oracle(oracle_mod, m, m.camera())
    sel = np.arange(0, 1000, 25)
    ro = o.gain_eval(lt_o, c.positions[sel], per_yaw=True, threads=16)
    for k in ("per_yaw", "gain", "yaw_deg", "steps"):
        assert np.array_equal(rg[k][sel], ro[k]), k
    big_seg = np.tile(c.segments, (5, 1))
    assert np.array_equal(g.check_segments(le, cfg.robot_radius, big_seg), np.tile(free1, 5))
    # n = 1 shims on the same handle
    st = np.array([*c.positions[7], 0.0])
    assert g.optimized_gain(lt, st) == rg["gain"][7] and g.check_motion(le, cfg.robot_radius, c.segments[7, :3], c.segments[7, 3:]) == bool(free1[7])
    # incremental commit: remove + overwrite, replicas stay identical to a single-GPU mirror updated the same way
    g.layer_remove(lt, m.keys[:40])
    t2 = m.tsdf[40:80].copy()
    t2[..., 1] = 0.0
    g.layer_update(lt, m.keys[40:80], t2)
    g.commit()
    g1, lt1, le1 = load_gpu(m, m.camera())
    g1.layer_remove(lt1, m.keys[:40]); g1.layer_update(lt1, m.keys[40:80], t2); g1.commit()
    assert g.layer_num_blocks(lt) == g1.layer_num_blocks(lt1) == len(m.keys) - 40
    ra = g.gain_eval(lt, c.positions, per_yaw=True)
    rb = g1.gain_eval(lt1, c.positions, per_yaw=True)
    assert np.array_equal(ra["per_yaw"], rb["per_yaw"]) and np.array_equal(ra["gain"], rb["gain"])
    st = g.stats()
    assert st["voxel_steps"] > 0 and st["blocks_uploaded"] >= len(devices) * (2 * len(m.keys) + 40)
    g1.close()
    g.close()


def test_pinned_updates_equal_copied_updates(cfg1_map):
    """nbvgpu_layer_update_pinned (no staging copy: the commit reads the caller's page-locked memory) gives the same mirror."""
    import torch
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, NbvGpuError
    m = cfg1_map
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
    le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 8)
    pt = torch.from_numpy(m.tsdf).pin_memory()
    pe = torch.from_numpy(m.esdf).pin_memory()
    g.layer_update_pinned(lt, m.keys, pt)
    g.layer_update_pinned(le, m.keys, pe)
    g.commit()
    with pytest.raises(NbvGpuError):
        g.layer_update_pinned(lt, m.keys, m.tsdf)          # pageable memory is refused
    g.set_camera(m.camera())
    g1, lt1, le1 = load_gpu(m, m.camera())
    from nbv_exploration_planner_b200 import synth
    c = synth.make_candidates(m.cfg, 64, None)
    a, b = g.gain_eval(lt, c.positions, per_yaw=True), g1.gain_eval(lt1, c.positions, per_yaw=True)
    assert np.array_equal(a["per_yaw"], b["per_yaw"]) and np.array_equal(a["gain"], b["gain"])
    assert np.array_equal(g.check_segments(le, 0.3, c.segments), g1.check_segments(le1, 0.3, c.segments))
    g.close(); g1.close()


@pytest.mark.parametrize("devices", [[0], "all"])
def test_shared_arena_commit_pulls_slices(cfg1_map, devices, monkeypatch):
    """nbvgpu_shared_alloc + nbvgpu_layer_update_pinned: with >= 2 GPUs every GPU copies its slice of each chunk and an
    all-gather completes it (forced here for 1 MiB chunks); the mirrors answer like a plain single-GPU context's."""
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, NbvGpuError, synth
    if devices == "all":
        if _n_gpus() < 2:
            pytest.skip("needs >= 2 GPUs")
        devices = list(range(_n_gpus()))
    monkeypatch.setenv("NBVGPU_CHUNK_MB", "1")
    monkeypatch.setenv("NBVGPU_PULL_MIN_MB", "0")
    m, cfg = cfg1_map, synth.CONFIGS["cfg2"]
    g = NbvGpu.multi(devices)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 64)
    le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 64)
    buf = g.shared_alloc(m.tsdf.nbytes + m.esdf.nbytes + 512)
    t = buf[:m.tsdf.nbytes].view(m.tsdf.dtype).reshape(m.tsdf.shape)
    e = buf[m.tsdf.nbytes + 256:m.tsdf.nbytes + 256 + m.esdf.nbytes].view(m.esdf.dtype).reshape(m.esdf.shape)
    t[...] = m.tsdf
    e[...] = m.esdf
    half = len(m.keys) // 2
    g.layer_update_pinned(lt, m.keys[:half], t[:half])      # several pieces per chunk, and a piece that is not shared memory
    g.layer_update_pinned(lt, m.keys[half:], t[half:])
    g.layer_update_pinned(le, m.keys, e)
    g.layer_update(le, m.keys[:2], m.esdf[:2])              # (copied into the library's own staging: that chunk is broadcast)
    with pytest.raises(NbvGpuError):
        g.shared_free(buf)                                  # staged blocks still point into it
    g.commit()
    assert g.layer_num_blocks(lt) == len(m.keys) and g.layer_num_blocks(le) == len(m.keys)
    assert (g.pulled_chunks() > 0) == (len(devices) > 1)
    g.shared_free(buf)
    g.set_camera(m.camera())
    c = synth.make_candidates(cfg, 600, None)
    free1, r1 = _reference_results(m, cfg, c)
    r = g.evaluate_candidates(lt, le, cfg.robot_radius, c.positions, c.segments, np.zeros(600), c.distance, 1.0, 0.5, 0.0)
    assert np.array_equal(r["free"], free1) and np.array_equal(r["gain"], r1["gain"]) and np.array_equal(r["yaw_deg"], r1["yaw_deg"])
    assert (r["index"], r["score"]) == (r1["index"], r1["score"])
    g.close()


def _run_ranks(tmp_path, world, n_cand):
    idfile, out = str(tmp_path / "uid.bin"), str(tmp_path / "res")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_rank_worker.py"), str(r), str(world), idfile, out, str(n_cand)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    try:
        logs = [p.communicate(timeout=420)[0] for p in procs]
    except subprocess.TimeoutExpired:
        for p in procs:
            p.kill()
        raise AssertionError("a rank hung:\n" + "\n".join((p.communicate()[0] or "")[-2000:] for p in procs))
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    return [np.load(out + f".{r}.npz") for r in range(world)]


@pytest.mark.parametrize("world", [1, "all"])
def test_one_process_per_gpu_equals_single_gpu(cfg1_map, tmp_path, world, monkeypatch):
    from nbv_exploration_planner_b200 import synth
    monkeypatch.setenv("NBVGPU_PULL_MIN_MB", "0")   # the workers' last commit comes out of shared memory: pulled whatever its size
    if world == "all":
        if _n_gpus() < 2:
            pytest.skip("needs >= 2 GPUs")
        world = _n_gpus()
    m, cfg = cfg1_map, synth.CONFIGS["cfg2"]
    n_cand = 1001                                   # uneven shards
    c = synth.make_candidates(cfg, n_cand, None)
    free1, r1 = _reference_results(m, cfg, c)
    res = _run_ranks(tmp_path, world, n_cand)
    want_best = np.array([r1["index"], r1["score"], r1["best_gain"], r1["best_yaw"]])
    for r, z in enumerate(res):
        lo, hi = int(z["lo"]), int(z["hi"])
        assert np.array_equal(z["free"], free1[lo:hi]) and np.array_equal(z["d_free"], free1[lo:hi])
        assert np.array_equal(z["gain"], r1["gain"][lo:hi]) and np.array_equal(z["d_gain"], r1["gain"][lo:hi])
        assert np.array_equal(z["yaw"], r1["yaw_deg"][lo:hi]) and np.array_equal(z["d_yaw"], r1["yaw_deg"][lo:hi])
        assert np.array_equal(z["best"], want_best) and np.array_equal(z["best_dev"], want_best)   # every rank knows the winner
        assert np.array_equal(z["gain_pulled"], r1["gain"][lo:hi]) and np.array_equal(z["free_pulled"], free1[lo:hi])
        assert (int(z["pulled_chunks"]) > 0) == (world > 1)
    assert sum(int(z["hi"]) - int(z["lo"]) for z in res) == n_cand


def test_cpp_shim_drives_all_gpus(oracle_mod, cfg1_map, tmp_path):
    """tests/cpp/shim_multi_gpu.cc: one C++ process, include/nbvgpu_shim.hpp, every visible GPU."""
    from cpp_dropin_fixture import write_case
    from nbv_exploration_planner_b200 import _native as N, synth
    m, cfg = cfg1_map, synth.CONFIGS["cfg2"]
    cam = m.camera()
    c = synth.make_candidates(cfg, 600, None)
    N.build()
    exe = str(tmp_path / "shim_multi_gpu")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "cpp", "shim_multi_gpu.cc"), "-o", exe, "-L", N.CSRC, "-lnbvgpu",
                           "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{N.CSRC}"])
    inp, outp = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    write_case(inp, m, cam, c, 2)
    r = subprocess.run([exe, inp, outp, "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(outp, np.float64)
    rows, tail = raw[:-8].reshape(-1, 3), raw[-8:]
    assert int(tail[7]) == _n_gpus()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    fo = o.check_segments(le_o, cfg.robot_radius, c.segments, threads=8)["free"]
    ro = o.gain_eval(lt_o, c.positions, threads=16)
    assert np.array_equal(rows[:, 2].astype(np.uint8), fo)
    gain = np.where(fo == 1, ro["gain"], 0.0)
    yaw = np.where(fo == 1, ro["yaw_deg"], 0)
    assert np.array_equal(rows[:, 0], gain) and np.array_equal(rows[:, 1], yaw)
    dist = 1.0 + 0.001 * np.arange(600)
    score = np.where(fo == 1, gain / np.maximum((yaw * np.pi / 180.0) / 0.5, dist / 1.0), -1.0)
    assert int(tail[0]) == int(np.argmax(score)) and tail[1] == score.max()
    assert tail[4] == ro["gain"][0] and tail[5] == ro["yaw_deg"][0] * np.pi / 180.0 and bool(tail[6]) == bool(fo[0])
    print(r.stdout.strip())
