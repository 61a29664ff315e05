"""CPU suite: the oracle reproduces the committed golden vectors (guards the checker itself), and the
synthetic generators are deterministic."""
import os

import numpy as np

from conftest import load_oracle
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_rays_golden():
    g = np.load(os.path.join(GOLD, "rays_golden.npz"))
    off = 0
    for s, e, n in zip(g["starts"], g["ends"], g["lengths"]):
        idx = ob.cast_ray(s, e)
        assert len(idx) == n and np.array_equal(idx, g["indices"][off:off + n])
        off += n


def test_tiny_golden(oracle_mod, tiny_map):
    g = np.load(os.path.join(GOLD, "tiny_golden.npz"))
    assert tiny_map.sha256() == str(g["map_sha256"])
    o, lt, le = load_oracle(oracle_mod, tiny_map, tiny_map.camera())
    r = o.gain_eval(lt, g["positions"], per_yaw=True, threads=4)
    assert np.array_equal(r["per_yaw"], g["per_yaw"]) and np.array_equal(r["yaw_deg"], g["yaw_deg"])
    assert np.array_equal(r["gain"], g["gain"]) and np.array_equal(r["steps"], g["steps"])
    f = o.check_segments(le, tiny_map.cfg.robot_radius, g["segments"])
    assert np.array_equal(f["free"], g["seg_free"])


def test_cfg1_golden_subset(oracle_mod, cfg1_map):
    g = np.load(os.path.join(GOLD, "cfg1_golden.npz"))
    assert cfg1_map.sha256() == str(g["map_sha256"])
    o, lt, le = load_oracle(oracle_mod, cfg1_map, cfg1_map.camera())
    sel = np.arange(0, 100, 9)
    r = o.gain_eval(lt, g["positions"][sel], per_yaw=True, threads=8)
    assert np.array_equal(r["per_yaw"], g["per_yaw"][sel]) and np.array_equal(r["gain"], g["gain"][sel])
    f = o.check_segments(le, cfg1_map.cfg.robot_radius, g["segments"])
    assert np.array_equal(f["free"], g["seg_free"]) and np.array_equal(f["steps"], g["seg_steps"])


def test_frontier_golden(oracle_mod, tiny_map):
    """History::bfs on the tiny map: the oracle reproduces the committed vectors, whose gains are also the outputs
    of the reference's own bfs (recorded when oracle/_ref was built at generation time)."""
    g = np.load(os.path.join(GOLD, "frontier_golden.npz"))
    assert tiny_map.sha256() == str(g["map_sha256"])
    o, lt, _ = load_oracle(oracle_mod, tiny_map, tiny_map.camera())
    r = o.frontier_potential(lt, g["points"], float(g["max_distance"]), threads=4)
    for k in ("gain", "frontier_voxels", "visited", "cells_checksum"):
        assert np.array_equal(r[k], g[k]), k
    assert len(g["reference_gain"]) == len(g["points"]) and np.array_equal(g["reference_gain"], g["gain"])


def test_interp_golden(oracle_mod, tiny_map):
    """ESDF interpolated distance / gradient on the tiny map: the oracle reproduces the committed vectors, which are also
    the outputs of the reference's own interpolator (recorded when oracle/_ref was built at generation time)."""
    g = np.load(os.path.join(GOLD, "interp_golden.npz"))
    assert tiny_map.sha256() == str(g["map_sha256"])
    o = oracle_mod.Oracle()
    le = o.layer_create(1, tiny_map.voxel_size, tiny_map.vps)
    o.layer_update_esdf_observed(le, tiny_map.keys, tiny_map.esdf, (tiny_map.tsdf[..., 1] > 0).astype(np.uint8))
    r = o.distance_and_gradient(le, g["positions"], threads=4)
    for k in ("ok", "distance", "gradient"):
        assert np.array_equal(r[k], g[k]), k
        assert np.array_equal(g["reference_" + k], g[k]), k
    assert 100 < int(g["ok"].sum()) < len(g["ok"]) - 100


def test_threaded_oracle_equals_single_thread(oracle_mod, tiny_map):
    o, lt, _ = load_oracle(oracle_mod, tiny_map, tiny_map.camera())
    pos = np.load(os.path.join(GOLD, "tiny_golden.npz"))["positions"]
    a = o.gain_eval(lt, pos, per_yaw=True, threads=1)
    b = o.gain_eval(lt, pos, per_yaw=True, threads=8)
    for k in ("per_yaw", "gain", "yaw_deg", "steps", "counts"):
        assert np.array_equal(a[k], b[k])


def test_generators_are_deterministic(tiny_map):
    from nbv_exploration_planner_b200 import synth
    m2 = synth.make_map("tiny")
    assert m2.sha256() == tiny_map.sha256()
    c1 = synth.make_candidates(tiny_map.cfg, 16, None)
    c2 = synth.make_candidates(tiny_map.cfg, 16, None)
    assert np.array_equal(c1.positions, c2.positions) and np.array_equal(c1.segments, c2.segments)
    # tree rule: every node within extension_range of its parent, inside the bounds
    lo, hi = np.array(tiny_map.cfg.bounds_min), np.array(tiny_map.cfg.bounds_max)
    d = np.linalg.norm(c1.positions - c1.segments[:, :3], axis=1)
    assert np.all(d <= tiny_map.cfg.extension_range + 1e-12)
    assert np.all(c1.positions >= lo - 1.0) and np.all(c1.positions <= hi + 1.0)


def test_nearest_parent_matches_reference_kdtree():
    """The generator's brute-force nearest parent equals the reference's own kdtree.c (built from
    /root/reference/kdtree/src/kdtree.c into oracle/_ref), which RrtTree::iterate uses (rrt.cpp:190-197)."""
    import ctypes as C
    import pytest
    path = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "libkdtree_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (reference absent)")
    kd = C.CDLL(path)
    kd.kd_create.restype = C.c_void_p
    kd.kd_create.argtypes = [C.c_int]
    kd.kd_insert3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
    kd.kd_nearest3.restype = C.c_void_p
    kd.kd_nearest3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    kd.kd_res_item_data.restype = C.c_void_p
    kd.kd_res_item_data.argtypes = [C.c_void_p]
    kd.kd_res_free.argtypes = [C.c_void_p]
    kd.kd_free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 20, (300, 3))
    t = kd.kd_create(3)
    for i, p in enumerate(pts):
        kd.kd_insert3(t, p[0], p[1], p[2], C.c_void_p(i + 1))
    for q in rng.uniform(0, 20, (200, 3)):
        res = kd.kd_nearest3(t, q[0], q[1], q[2])
        got = kd.kd_res_item_data(res) - 1
        kd.kd_res_free(res)
        assert got == int(((pts - q) ** 2).sum(1).argmin())
    kd.kd_free(t)
