"""Pins the oracle against the only known answers the reference holds for this path: the index-math
KATs of voxblox/voxblox/test/test_tsdf_map.cc, plus hand-computed micro-cases of the RayCaster
(voxblox/src/integrator/integrator_utils.cc:111-179) including its degenerate inf/NaN behaviour
(SURVEY.md App. A.5), which no reference test covers."""
import numpy as np
import pytest

from oracle import binding as ob


# ---- test_tsdf_map.cc: voxel 0.1, 8 voxels per side -> block size 0.8 ------------------------------
VS, VPS = np.float32(0.1), 8
BS = np.float32(VS * VPS)
BS_INV = np.float32(1.0) / BS
VS_INV = np.float32(1.0) / VS


def block_of(p):
    return ob.grid_index_from_point(np.asarray(p, np.float32), BS_INV).tolist()


def test_block_index_from_coordinates_kats():
    # test_tsdf_map.cc:36-136 (IndexLookups): origin / within-block points of blocks 0, 1, -1
    assert block_of((0, 0, 0)) == [0, 0, 0]
    assert block_of((0, VS, 0)) == [0, 0, 0]
    assert block_of((BS, BS, BS)) == [1, 1, 1]
    assert block_of((BS, BS + VS, BS)) == [1, 1, 1]
    assert block_of((-BS, -BS, -BS)) == [-1, -1, -1]
    assert block_of((-BS, -BS + VS, -BS)) == [-1, -1, -1]
    # test_tsdf_map.cc:24-34 (BlockAllocation): (0,0.15,0) and (0,0.13,0) share a block, (-10,13.5,20) does not
    assert block_of((0, 0.15, 0)) == block_of((0, 0.13, 0))
    assert block_of((-10.0, 13.5, 20.0)) != block_of((0, 0.15, 0))


def test_linear_index_kats():
    # test_tsdf_map.cc:146 -> voxel (0,1,0) has linear index 8; :305 -> voxel (3,6,5) has 371 (vps = 8)
    assert ob.linear_index(8, 0, 1, 0) == 8
    assert ob.linear_index(8, 3, 6, 5) == 371
    # voxel index from coordinates relative to the block origin (block.h:67-72)
    assert ob.grid_index_from_point(np.array([0, 1 * VS, 0], np.float32), VS_INV).tolist() == [0, 1, 0]
    assert ob.grid_index_from_point(np.array([3 * VS, 6 * VS, 5 * VS], np.float32), VS_INV).tolist() == [3, 6, 5]


def test_last_voxel_of_negative_block():
    # test_tsdf_map.cc:249-264: a point just below the origin lies in block (-1,-1,-1), last voxel
    g = np.array([-1, -1, -1], np.int64)
    b, l = ob.block_and_local(g, 8)
    assert b.tolist() == [-1, -1, -1] and l.tolist() == [7, 7, 7]
    assert ob.linear_index(8, *l) == 8 ** 3 - 1


def test_global_block_local_round_trip():
    # test_tsdf_map.cc:325-360 round-trips origin <-> index over a cube of blocks; same idea on the
    # integer path the gain uses (common.h:205-243)
    rng = np.random.default_rng(0)
    for vps in (8, 16):
        for g in rng.integers(-3000, 3000, size=(200, 3)):
            b, l = ob.block_and_local(g.astype(np.int64), vps)
            assert np.all(l >= 0) and np.all(l < vps)
            assert np.array_equal(b.astype(np.int64) * vps + l, g)
            assert np.array_equal(b, np.floor_divide(g, vps))


def test_any_index_hash():
    # block_hash.h:27-30: x + y*17191 + z*17191^2 truncated to unsigned int
    assert ob.block_hash(0, 0, 0) == 0
    assert ob.block_hash(1, 0, 0) == 1
    assert ob.block_hash(0, 1, 0) == 17191
    assert ob.block_hash(0, 0, 1) == 17191 ** 2
    assert ob.block_hash(-1, 0, 0) == 2 ** 32 - 1
    assert ob.block_hash(3, -2, 5) == (3 - 2 * 17191 + 5 * 17191 ** 2) % 2 ** 32


# ---- RayCaster micro-cases ------------------------------------------------------------------------
def test_cast_ray_axis_aligned_positive():
    idx = ob.cast_ray([0.5, 0.5, 0.5], [3.5, 0.5, 0.5])
    # y and z components are exactly zero: t = (0 - 0.5)/0 = -inf is picked once (y), then NaN; z
    # likewise -> the start voxel is repeated and the ray ends short (len + 1 = 4 emissions)
    assert len(idx) == 4
    assert idx[0].tolist() == [0, 0, 0]


def test_cast_ray_generic_diagonal():
    idx = ob.cast_ray([0.2, 0.3, 0.1], [2.7, 1.6, 0.4])
    # len = |2-0| + |1-0| + |0-0| = 3 -> 4 emissions, 6-connected steps, ends in the end voxel
    assert len(idx) == 4
    assert idx[0].tolist() == [0, 0, 0] and idx[-1].tolist() == [2, 1, 0]
    assert np.all(np.abs(np.diff(idx, axis=0)).sum(1) == 1)


def test_cast_ray_emission_count_is_manhattan_plus_one():
    rng = np.random.default_rng(1)
    for _ in range(200):
        s = rng.uniform(-20, 20, 3).astype(np.float32)
        e = rng.uniform(-20, 20, 3).astype(np.float32)
        idx = ob.cast_ray(s, e)
        si = np.floor(s + np.float32(1e-6)).astype(np.int64)
        ei = np.floor(e + np.float32(1e-6)).astype(np.int64)
        assert len(idx) == np.abs(ei - si).sum() + 1
        assert idx[0].tolist() == si.tolist()
        assert np.all(np.abs(np.diff(idx, axis=0)).sum(1) == 1)


def test_cast_ray_zero_x_component_sticks_to_start_voxel():
    # r_x == 0 with shift > 0 -> t_x = -inf then NaN at index 0, which
    # minCoeff always selects: the start voxel is emitted len + 1 times.
    idx = ob.cast_ray([0.5, 0.25, 0.25], [0.5, 4.25, 0.75])
    assert len(idx) == 5
    assert all(i.tolist() == [0, 0, 0] for i in idx)


def test_cast_ray_zero_y_component_loses_one_step():
    idx = ob.cast_ray([0.25, 0.5, 0.25], [4.75, 0.5, 1.25])
    # len = 4 + 0 + 1 = 5 -> 6 emissions; the y NaN costs one step: the start voxel repeats once
    assert len(idx) == 6
    assert idx[0].tolist() == [0, 0, 0] and idx[1].tolist() == [0, 0, 0]
    assert idx[-1].tolist() != [4, 0, 1]


def test_cast_ray_same_voxel():
    idx = ob.cast_ray([1.2, 1.3, 1.4], [1.6, 1.7, 1.8])
    assert len(idx) == 1 and idx[0].tolist() == [1, 1, 1]


def test_cast_ray_negative_direction():
    idx = ob.cast_ray([2.5, 2.5, 2.5], [-0.5, 1.5, 2.2])
    assert idx[0].tolist() == [2, 2, 2] and idx[-1].tolist() == [-1, 1, 2]
    assert len(idx) == 3 + 1 + 0 + 1


# ---- single-voxel worlds: status rule + frustum --------------------------------------------------
def _one_block_layer(weight, distance, vs=0.2, vps=16):
    o = ob.Oracle()
    lt = o.layer_create(0, vs, vps)
    data = np.zeros((1, vps ** 3, 2), np.float32)
    data[..., 0] = distance
    data[..., 1] = weight
    o.layer_update(lt, np.array([[0, 0, 0]], np.int32), data)
    return o, lt


def test_voxel_status_rule():
    # voxblox_manager.cpp:23-29: weight <= 1e-1 -> unknown (0.1f = 0.100000001 is NOT unknown)
    for w, d, want in ((0.0, 1.0, 0), (0.05, -1.0, 0), (np.float32(0.1), 1.0, 1), (np.float32(0.1), 0.0, 2),
                       (1.0, -0.3, 2), (1.0, 0.3, 1), (np.nextafter(np.float32(0.1), np.float32(0)), 1.0, 0)):
        o, lt = _one_block_layer(w, d)
        assert o.voxel_status(lt, [1, 2, 3]) == want
        assert o.voxel_status(lt, [16, 0, 0]) == 0  # unallocated block -> unknown


def test_gain_counts_free_block_only():
    """A camera inside one fully free block, box = that block: every counted voxel is free, nothing
    is unknown inside the block, and voxels outside the exploration box are not counted at all."""
    from nbv_exploration_planner_b200.planner import CameraParams
    vs = 0.2
    o, lt = _one_block_layer(1.0, 1.0, vs)
    cam = CameraParams(hfov_deg=90, vfov_deg=60, near_m=0.1, far_m=1.0, bbx_min=(0, 0, 0), bbx_max=(3.0, 3.0, 3.0))
    o.set_camera(cam.as_dict())
    r = o.gain_eval(lt, np.array([[1.6, 1.6, 1.6]]), per_yaw=True)
    assert r["per_yaw"][0, :, 0].sum() == 0 and r["per_yaw"][0, :, 2].sum() == 0
    assert r["per_yaw"][0, :, 1].sum() > 0
    assert r["gain"][0] == 0.0 and r["yaw_deg"][0] == 0
    # shrink the exploration box: fewer voxels pass the in-view test
    cam2 = CameraParams(hfov_deg=90, vfov_deg=60, near_m=0.1, far_m=1.0, bbx_min=(1.0, 1.0, 1.0), bbx_max=(2.0, 2.0, 2.0))
    o.set_camera(cam2.as_dict())
    r2 = o.gain_eval(lt, np.array([[1.6, 1.6, 1.6]]), per_yaw=True)
    assert 0 < r2["per_yaw"][0, :, 1].sum() < r["per_yaw"][0, :, 1].sum()


def test_gain_unknown_world_prefers_any_yaw_and_scales_with_cube():
    """Empty map: every in-view voxel is unknown, gain = window sum * vs^3 > 0, best yaw is the
    first 5-degree window with the maximal sum."""
    from nbv_exploration_planner_b200.planner import CameraParams
    o = ob.Oracle()
    lt = o.layer_create(0, 0.2, 16)
    cam = CameraParams(hfov_deg=90, vfov_deg=60, near_m=0.1, far_m=2.0, bbx_min=(-50, -50, -50), bbx_max=(50, 50, 50))
    o.set_camera(cam.as_dict())
    r = o.gain_eval(lt, np.array([[0.3, 0.1, 0.7]]), per_yaw=True)
    py = r["per_yaw"][0].astype(np.float64)
    assert py[:, 1].sum() == 0 and py[:, 2].sum() == 0
    vg = py[:, 0]
    best, best_yaw = 0.0, 0
    for i in range(-180, 180, 5):
        w = sum(vg[(i + 180 + d) % 360] for d in range(-45, 46))
        if w > best:
            best, best_yaw = w, i
    assert r["yaw_deg"][0] == best_yaw
    assert r["gain"][0] == pytest.approx(best * float(np.float32(0.2)) ** 3, rel=1e-12)


def test_occupied_voxel_stops_ray():
    """Fully occupied block around the camera: each ray counts exactly one occupied voxel (its first
    in-view voxel) and stops."""
    from nbv_exploration_planner_b200.planner import CameraParams
    o, lt = _one_block_layer(1.0, -0.1, 0.2)
    cam = CameraParams(hfov_deg=90, vfov_deg=60, near_m=0.1, far_m=1.0, bbx_min=(0, 0, 0), bbx_max=(3.2, 3.2, 3.2))
    o.set_camera(cam.as_dict())
    r = o.gain_eval(lt, np.array([[1.6, 1.6, 1.6]]), per_yaw=True)
    py = r["per_yaw"][0]
    assert py[:, 0].sum() == 0 and py[:, 1].sum() == 0
    assert py[:, 2].sum() <= r["rays"][0]
    assert py[:, 2].sum() > 0.5 * r["rays"][0]


def test_check_motion_rules():
    o = ob.Oracle()
    le = o.layer_create(1, 0.2, 16)
    d = np.full((1, 4096), 1.0, np.float32)
    o.layer_update(le, np.array([[0, 0, 0]], np.int32), d)
    seg_in = np.array([[0.5, 0.5, 0.5, 2.5, 2.0, 1.0]])
    assert o.check_segments(le, 0.4, seg_in)["free"][0] == 1
    assert o.check_segments(le, 1.0, seg_in)["free"][0] == 0      # radius >= distance collides (>=)
    assert o.check_segments(le, 0.9999, seg_in)["free"][0] == 1
    seg_out = np.array([[0.5, 0.5, 0.5, 4.0, 0.5, 0.5]])          # leaves the allocated block -> collision
    assert o.check_segments(le, 0.1, seg_out)["free"][0] == 0
