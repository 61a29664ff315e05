"""Pins the oracle restatement against the REFERENCE'S OWN SOURCE: oracle/_ref/libnbv_ref.so is compiled
from /root/reference's integrator_utils.cc, camera_model.cpp, the gain loops of rrt.cpp and the status /
collision members of voxblox_manager.{h,cpp}, with the real Voxblox core headers, against stand-ins for the
third-party headers only (oracle/ref_shim/).  Every comparison below is bit-exact."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from oracle import ref_binding as rb

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref/libnbv_ref.so not built (reference absent)")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def reflib():
    return rb.RefLib()


def test_cast_ray_matches_reference_raycaster(reflib):
    """voxblox::castRay / RayCaster (integrator_utils.cc:111-179) incl. the degenerate inf/NaN rays."""
    g = np.load(os.path.join(GOLD, "rays_golden.npz"))
    off = 0
    for s, e, n in zip(g["starts"], g["ends"], g["lengths"]):
        r = reflib.cast_ray(s, e)
        assert len(r) == n and np.array_equal(r, g["indices"][off:off + n])
        off += n
    rng = np.random.default_rng(3)
    for k in range(3000):
        s = rng.uniform(-60, 60, 3).astype(np.float32)
        e = (s + rng.uniform(-40, 40, 3)).astype(np.float32)
        if k % 7 == 0:
            e[rng.integers(0, 3)] = s[rng.integers(0, 3)] if k % 14 else e[0]   # zero / odd components
        if k % 11 == 0:
            ax = rng.integers(0, 3)
            e[ax] = s[ax]
        if k % 13 == 0:
            s = np.round(s)                                                      # starts on voxel corners
        assert np.array_equal(ob.cast_ray(s, e), reflib.cast_ray(s, e)), (s, e)


def test_index_math_and_hash_match_reference_headers(reflib):
    """voxblox/core/common.h:151-243 and block_hash.h:20-31 as compiled from the reference."""
    rng = np.random.default_rng(4)
    for vps in (8, 16, 32):
        for g in rng.integers(-100000, 100000, (300, 3)):
            bo, lo = ob.block_and_local(g.astype(np.int64), vps)
            br, lr = reflib.block_and_local(g.astype(np.int64), vps)
            assert np.array_equal(bo, br) and np.array_equal(lo, lr)
    for p in rng.uniform(-50, 50, (500, 3)).astype(np.float32):
        for inv in (np.float32(1 / 0.8), np.float32(10.0), np.float32(1 / 2.4)):
            assert np.array_equal(ob.grid_index_from_point(p, inv), reflib.grid_index_from_point(p, inv))
    for x, y, z in rng.integers(-5000, 5000, (500, 3)):
        assert ob.block_hash(int(x), int(y), int(z)) == reflib.block_hash(int(x), int(y), int(z))


def _both(m, cam, robot_radius):
    o = ob.Oracle()
    lt = o.layer_create(0, m.voxel_size, m.vps)
    le = o.layer_create(1, m.voxel_size, m.vps)
    o.layer_update(lt, m.keys, m.tsdf)
    o.layer_update(le, m.keys, m.esdf)
    o.set_camera(cam.as_dict())
    r = rb.Reference(m.voxel_size, m.vps)
    r.tsdf_update(m.keys, m.tsdf)
    r.esdf_update(m.keys, m.esdf)
    r.set_camera(cam.as_dict(), robot_radius)
    return o, lt, le, r


def test_camera_model_frames_match_reference(tiny_map):
    """CameraModel::setBodyPose / calculateBoundingPlanes / getFarPlanePoints (camera_model.cpp) for every
    yaw of several body positions: planes, camera position and far-plane points, bit for bit."""
    from nbv_exploration_planner_b200.planner import CameraParams
    cams = [tiny_map.camera(), CameraParams.mount(-20.0, (0.05, 0.02, -0.1), hfov_deg=69.0, vfov_deg=42.0, near_m=0.3, far_m=3.0,
                                                 bbx_min=(-4, -1, 0), bbx_max=(10, 14, 3)),
            CameraParams(hfov_deg=120, vfov_deg=90, near_m=0.1, far_m=7.5, bbx_min=(-50, -50, -5), bbx_max=(50, 50, 5))]
    rng = np.random.default_rng(5)
    for cam in cams:
        o, lt, le, r = _both(tiny_map, cam, 0.4)
        for pos in np.concatenate([rng.uniform(-30, 30, (6, 3)), [[0.0, 0.0, 0.0], [4.0, 4.0, 1.2]]]):
            for yaw in list(range(-180, 180, 7)) + [-90, 0, 90, 179]:
                fo, fr = o.yaw_frame(lt, pos, yaw), r.yaw_frame(pos, yaw)
                assert np.array_equal(fo["planes"], fr["planes"]) and np.array_equal(fo["cam"], fr["cam"])
                assert np.array_equal(fo["pp"], fr["pp"])


@pytest.mark.parametrize("sum_all", [0, 1])
def test_gain_matches_reference_rrt_loops(tiny_map, sum_all):
    """RrtTree::optimized_gain (rrt.cpp:351-479) / RrtTree::gain (:481-579), verbatim: gain and state[3]."""
    from nbv_exploration_planner_b200 import synth
    cam = tiny_map.camera(sum_all_yaws=sum_all)
    o, lt, le, r = _both(tiny_map, cam, tiny_map.cfg.robot_radius)
    c = synth.make_candidates(tiny_map.cfg, 16, None)
    pos = np.concatenate([c.positions, [[4.0, 4.0, 1.2], [3.2, 3.2, 1.6], [0.0, 0.0, 0.0], [8.0, 8.0, 3.0]]])
    ro, rr = o.gain_eval(lt, pos, threads=8), r.gain_eval(pos)
    assert np.array_equal(ro["gain"], rr["gain"]) and np.array_equal(ro["yaw"], rr["yaw"])
    assert ro["gain"].max() > 0


def test_gain_matches_reference_other_geometry(tiny_map):
    """The shipped yaml's camera (d435 69x42 deg, 3 m range) on a 0.3 m / 16^3 layer, non-default gains."""
    from nbv_exploration_planner_b200 import synth
    cfg = synth.SynthConfig("pin", (-6, -5, -1), (6, 7, 3), 0.3, 3.0, 10, "pillars", 77, 78, (0.5, 0.3, 0.9), obs_radius=4.0,
                            traj_points=5, traj_step=2.0, vicinity=3.0, hfov_deg=69.0, vfov_deg=42.0, pitch_deg=0.0)
    m = synth.make_map(cfg)
    cam = m.camera(gain_free=0.25, gain_occupied=0.5, gain_unmapped=1.0)   # exactly representable weights
    o, lt, le, r = _both(m, cam, 0.4)
    c = synth.make_candidates(cfg, 10, None)
    ro, rr = o.gain_eval(lt, c.positions, threads=8), r.gain_eval(c.positions)
    assert np.array_equal(ro["gain"], rr["gain"]) and np.array_equal(ro["yaw"], rr["yaw"])


@pytest.mark.parametrize("which", ["cfg3-like", "cfg4-like"])
def test_gain_matches_reference_at_fine_voxels(which):
    """The geometries BASELINE configs 3 and 4 live on -- 0.10 m voxels / 5 m range (58 ray rows, ~1.2 M steps per
    viewpoint) and 0.05 m voxels / 8 m range (185 rows, ~14 M steps) -- on small maps: the oracle equals the
    reference's own optimized_gain (gain and state[3]) and checkMotion, bit for bit."""
    from nbv_exploration_planner_b200 import synth
    if which == "cfg3-like":
        cfg = synth.SynthConfig("cfg3small", (0, 0, 0), (20, 20, 10), 0.10, 5.0, 8, "office", 1203, 2203, (7.5, 7.5, 1.5), obs_radius=6.0,
                                traj_points=6, traj_step=4.0, vicinity=4.0)
        n = 8
    else:
        cfg = synth.SynthConfig("cfg4small", (0, 0, 0), (16, 16, 6), 0.05, 8.0, 6, "pillars", 1204, 2204, (8.0, 8.0, 2.0), obs_radius=7.0,
                                traj_points=3, traj_step=3.0, vicinity=2.0)
        n = 4
    m = synth.make_map(cfg)
    o, lt, le, r = _both(m, m.camera(), cfg.robot_radius)
    c = synth.make_candidates(cfg, n, None)
    ro, rr = o.gain_eval(lt, c.positions, threads=8), r.gain_eval(c.positions)
    assert np.array_equal(ro["gain"], rr["gain"]) and np.array_equal(ro["yaw"], rr["yaw"])
    assert ro["gain"].max() > 0 and ro["steps"].min() > 1.0e6
    assert np.array_equal(o.check_segments(le, cfg.robot_radius, c.segments)["free"], r.check_segments(c.segments))


@pytest.mark.parametrize("sum_all", [0, 1])
@pytest.mark.parametrize("state", ["unknown", "free", "occupied"])
def test_rays_stuck_at_yaw_plus_minus_90_match_reference(state, sum_all):
    """At the integer yaws +-90 deg the ray has no x component; the reference's RayCaster then emits the voxel
    under the camera len + 1 times (integrator_utils.cc:119-123,165-179: t = -inf or 0/0, dt = 0/0, minCoeff
    picks that axis for ever).  The reference only returns gain and yaw, so the repeated voxel is pinned
    through totals: exactly representable weights make an unknown / free / occupied start voxel show up in the
    gain of both RrtTree::gain (all-yaw sum) and RrtTree::optimized_gain (best window)."""
    from nbv_exploration_planner_b200.planner import CameraParams
    vs, vps = 0.2, 16
    cam = CameraParams(hfov_deg=90, vfov_deg=60, near_m=0.01, far_m=2.0, bbx_min=(-4, -4, -4), bbx_max=(8, 8, 8),
                       gain_free=0.25, gain_occupied=0.5, gain_unmapped=1.0, sum_all_yaws=sum_all)  # identity T_C_B
    keys = np.array([[0, 0, 0]], np.int32)
    tsdf = np.zeros((1, vps ** 3, 2), np.float32)
    lin = 5 + 5 * vps + 5 * vps * vps
    if state == "free":
        tsdf[0, lin] = (0.4, 1.0)
    elif state == "occupied":
        tsdf[0, lin] = (-0.05, 1.0)
    # camera near the +y face of voxel (5,5,5): its min corner is in view looking along -y, behind the camera along +y
    pos = np.array([[(5 + 0.02) * vs, (5 + 0.9) * vs, (5 + 0.03) * vs], [(5 + 0.5) * vs, (5 + 0.5) * vs, (5 + 0.5) * vs]])
    o = ob.Oracle()
    lt = o.layer_create(0, vs, vps)
    o.layer_update(lt, keys, tsdf)
    o.set_camera(cam.as_dict())
    r = rb.Reference(vs, vps)
    r.tsdf_update(keys, tsdf)
    r.set_camera(cam.as_dict(), 0.4)
    ro, rr = o.gain_eval(lt, pos, per_yaw=True, threads=2), r.gain_eval(pos)
    assert np.array_equal(ro["gain"], rr["gain"]) and np.array_equal(ro["yaw"], rr["yaw"])
    # and the oracle's own per-yaw view of it: the voxel is counted once per step of every ray of the column
    u_max = int(ro["rays"][0]) // 360
    want = {"unknown": 0, "free": 1, "occupied": 2}[state]
    col = ro["per_yaw"][0, 90]
    assert col[want] >= u_max and col.sum() == col[want] and tuple(ro["per_yaw"][0, 270]) == (0, 0, 0)


def test_check_motion_and_voxel_status_match_reference(tiny_map):
    """HighResManager::checkMotion (voxblox_manager.h:89-108, .cpp:121-129) and getVoxelStatusByIndex (:14-33)."""
    from nbv_exploration_planner_b200 import synth
    cfg = tiny_map.cfg
    for radius in (0.0, 0.2, cfg.robot_radius, 1.0):
        o, lt, le, r = _both(tiny_map, tiny_map.camera(), radius)
        c = synth.make_candidates(cfg, 200, None)
        rng = np.random.default_rng(6)
        seg = np.concatenate([c.segments, np.concatenate([rng.uniform(-1, 9, (200, 2)), rng.uniform(-0.5, 3.5, (200, 1)),
                                                          rng.uniform(-1, 9, (200, 2)), rng.uniform(-0.5, 3.5, (200, 1))], 1)])
        assert np.array_equal(o.check_segments(le, radius, seg)["free"], r.check_segments(seg))
    rng = np.random.default_rng(7)
    for g in rng.integers(-10, 60, (400, 3)):
        b, l = ob.block_and_local(g.astype(np.int64), 16)
        assert o.voxel_status(lt, g) == r.voxel_status(b, l)


def test_cfg1_golden_subset_against_reference(cfg1_map):
    """BASELINE config 1 candidates: the committed golden gains / yaws / edge verdicts equal the reference's."""
    g = np.load(os.path.join(GOLD, "cfg1_golden.npz"))
    r = rb.Reference(cfg1_map.voxel_size, cfg1_map.vps)
    r.tsdf_update(cfg1_map.keys, cfg1_map.tsdf)
    r.esdf_update(cfg1_map.keys, cfg1_map.esdf)
    r.set_camera(cfg1_map.camera().as_dict(), cfg1_map.cfg.robot_radius)
    sel = np.arange(0, 100, 6)
    rr = r.gain_eval(g["positions"][sel])
    assert np.array_equal(rr["gain"], g["gain"][sel])
    assert np.array_equal(rr["yaw"], g["yaw_deg"][sel] * np.pi / 180.0)
    assert np.array_equal(r.check_segments(g["segments"]), g["seg_free"])


def test_voxel_status_by_coordinates_matches_reference(tiny_map):
    """LowResManager::getVoxelStatus(position) (voxblox_manager.cpp:73-90): float cast, block by coordinates,
    truncated voxel index -- on random points, points on voxel / block faces and points outside the map."""
    cam = tiny_map.camera()
    o, lt, le, r = _both(tiny_map, cam, 0.4)
    vs = float(tiny_map.voxel_size)
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.uniform(-2, 10, (3000, 3)),
                          rng.integers(-10, 50, (1500, 3)) * vs,                      # exactly on voxel faces (as doubles)
                          rng.integers(-2, 4, (300, 3)) * (16 * vs) + rng.choice([-1e-7, 0.0, 1e-7, 1e-6], (300, 3)),  # block faces
                          rng.integers(-10, 50, (1500, 3)) * np.float64(np.float32(vs))])  # accumulated-float-step lattice
    seen = set()
    for q in pts:
        a, b = o.voxel_status_at(lt, q), r.voxel_status_at(q)
        assert a == b, q
        seen.add(a)
    assert seen == {0, 1, 2}


@pytest.mark.parametrize("which", ["tiny", "quarter"])
def test_frontier_bfs_matches_reference_history(tiny_map, which):
    """History::bfs (history.cpp:96-137), verbatim: cells keyed by std::to_string of accumulated doubles, FIFO
    order, unknown neighbours counted per encounter, 6.0 m radius, exploration bounds.  Vertices in free,
    unknown and occupied space, near and outside the bounds; `quarter` uses a voxel size of exactly 0.25 m,
    where 24 steps reach 6.0 m exactly and the `<= 6.0` test sits on accumulated roundings."""
    from nbv_exploration_planner_b200 import synth
    if which == "tiny":
        m, cfg = tiny_map, tiny_map.cfg
    else:
        cfg = synth.SynthConfig("quarter", (-3, -2, 0), (9, 10, 3), 0.25, 3.0, 8, "pillars", 91, 92, (3.0, 4.0, 1.0), obs_radius=4.0,
                                traj_points=4, traj_step=2.5, vicinity=3.0)
        m = synth.make_map(cfg)
    cam = m.camera()
    o, lt, le, r = _both(m, cam, 0.4)
    c = synth.make_candidates(cfg, 6, None)
    lo, hi = np.array(cfg.bounds_min, float), np.array(cfg.bounds_max, float)
    extra = np.array([lo + 0.01, hi - 0.01, lo - 0.5, [lo[0], 0.5 * (lo[1] + hi[1]), 1.0],      # corners, outside, on a face
                      0.5 * (lo + hi), [cfg.root[0], cfg.root[1], cfg.root[2]], [1.0, 2.0, 0.5], [3.0, 4.0, 1.0]])
    pts = np.concatenate([c.positions, extra])
    ro, rr = o.frontier_potential(lt, pts, 6.0, threads=8), r.frontier_potential(pts)
    assert np.array_equal(ro["gain"], rr), (ro["gain"], rr)
    assert ro["frontier_voxels"].max() > 100 and ro["visited"].max() > 1000   # the fill really spreads


def test_frontier_bfs_negative_zero_coordinates_match_reference():
    """std::to_string(-0.0) == "-0.000000" while -0.0 + 0.0 == +0.0, and -1e-9 prints "-0.000000" for ever: start
    points with such coordinates make the reference's string keys split a cell.  Same map and points as
    tests/test_gpu_frontier.py::test_negative_zero_and_tiny_negative_coordinates (the reference hard-codes 6.0 m;
    the blocks only reach 3.2 m, so the fill is bounded by the unknown space around them)."""
    from nbv_exploration_planner_b200.planner import CameraParams
    vs, vps = 0.2, 16
    keys = np.array([[bx, by, bz] for bx in (-1, 0) for by in (-1, 0) for bz in (-1, 0)], np.int32)
    tsdf = np.zeros((len(keys), vps ** 3, 2), np.float32)
    tsdf[..., 0] = 0.4
    tsdf[..., 1] = 1.0
    rng = np.random.default_rng(2)
    tsdf[rng.random(tsdf.shape[:2]) < 0.03, 1] = 0.0
    tsdf[rng.random(tsdf.shape[:2]) < 0.03, 0] = -0.1
    tsdf[3, 15, 1] = 0.0        # the -x neighbour of the cell at the origin is unknown,
    tsdf[7, 0, :] = (0.4, 1.0)  # the cell at the origin is free
    cam = CameraParams(hfov_deg=90, vfov_deg=60, near_m=0.1, far_m=2.0, bbx_min=(-2.5, -2.5, -2.5), bbx_max=(2.5, 2.5, 2.5))
    o = ob.Oracle()
    lt = o.layer_create(0, vs, vps)
    o.layer_update(lt, keys, tsdf)
    o.set_camera(cam.as_dict())
    r = rb.Reference(vs, vps)
    r.tsdf_update(keys, tsdf)
    r.set_camera(cam.as_dict(), 0.4)
    pts = np.array([[-0.0, 0.1, 0.1], [0.0, 0.1, 0.1], [0.1, -0.0, -0.0], [-1e-9, 0.1, 0.1], [1e-9, -1e-9, 0.1], [-0.0, -0.0, -0.0],
                    [0.0078125, 0.0234375, 0.1], [-0.0078125, 0.3, -0.0234375]])
    ro, rr = o.frontier_potential(lt, pts, 6.0, threads=8), r.frontier_potential(pts)
    assert np.array_equal(ro["gain"], rr), (ro["gain"], rr)
    assert ro["visited"].min() > 1000
    # observable in the reference's own result: the -0.0 start counts the unknown neighbour of the origin cell twice
    assert rr[0] != rr[1] and ro["frontier_voxels"][0] == ro["frontier_voxels"][1] + 1


def _esdf_observed(m):
    """EsdfVoxel::observed for a synthetic map: the voxels the TSDF observed (weight > 0), as the ESDF integrator sets it."""
    return (m.tsdf[..., 1] > 0).astype(np.uint8)


def _interp_queries(m, rng, n=4000):
    vs = float(m.voxel_size)
    lo, hi = np.array(m.cfg.bounds_min, float), np.array(m.cfg.bounds_max, float)
    rand = rng.uniform(lo - 0.5, hi + 0.5, (n, 3))
    centres = (rng.integers(0, 40, (n // 4, 3)) + 0.5) * np.float64(np.float32(vs))        # voxel centres: exact interpolation
    faces = rng.integers(0, 40, (n // 4, 3)) * np.float64(np.float32(vs))                   # voxel faces / block corners
    near = centres[: n // 8] + rng.choice([-1e-6, 1e-6, -1e-4, 1e-4], (n // 8, 3))          # either side of a centre
    return np.concatenate([rand, centres, faces, near, [[0.0, 0.0, 0.0], [-0.0, 1.0, 1.0], lo, hi]])


@pytest.mark.parametrize("which", ["tiny", "quarter"])
def test_esdf_distance_and_gradient_match_reference_interpolator(tiny_map, which):
    """EsdfMap::getDistanceAndGradientAtPosition (esdf_map.cc:22-52) -> Interpolator (interpolator_inl.h), the
    reference's own lines compiled against the Eigen stand-in: success flag, distance and gradient -- the partial
    gradient a failing call leaves behind included -- identical to the oracle's restatement on random queries, voxel
    centres, voxel / block faces, points either side of a centre and points outside the map.  (The stand-in sums the
    8-term products left to right; against a real Eigen build the values are only claimed within float tolerance.)"""
    from nbv_exploration_planner_b200 import synth
    if which == "tiny":
        m = tiny_map
    else:
        m = synth.make_map(synth.SynthConfig("quarter", (-3, -2, 0), (9, 10, 3), 0.25, 3.0, 8, "pillars", 91, 92, (3.0, 4.0, 1.0), obs_radius=4.0,
                                             traj_points=4, traj_step=2.5, vicinity=3.0))
    obs = _esdf_observed(m)
    o = ob.Oracle()
    le = o.layer_create(1, m.voxel_size, m.vps)
    o.layer_update_esdf_observed(le, m.keys, m.esdf, obs)
    r = rb.Reference(m.voxel_size, m.vps)
    r.esdf_update_observed(m.keys, m.esdf, obs)
    q = _interp_queries(m, np.random.default_rng(17))
    ro, rr = o.distance_and_gradient(le, q, threads=8), r.distance_and_gradient(q)
    assert np.array_equal(ro["ok"], rr["ok"])
    assert np.array_equal(ro["distance"], rr["distance"]) and np.array_equal(ro["gradient"], rr["gradient"])
    assert 0.1 < ro["ok"].mean() < 0.95 and np.abs(ro["gradient"][ro["ok"] == 1]).max() > 0.5   # both outcomes, real gradients



@pytest.mark.parametrize("which", ["tiny", "cfg1"])
def test_sequential_tree_expansion_matches_reference_iterate(tiny_map, cfg1_map, which):
    """RrtTree::iterate (rrt.cpp:163-264), the reference's own lines with its own kd-tree (kdtree.c), called until a stream of
    uniform triples is used up -- only its clock-seeded RNG is replaced by that stream.  The oracle's sequential restatement
    (nbvo_expand_tree_seq) must create the same nodes: which triples are rejected / collide / create a node, and for every
    created node position, parent (among the given nodes AND the nodes created earlier in the stream), num_unmapped_,
    state_[3], distance_, gain_; and the same bestNode_."""
    m = tiny_map if which == "tiny" else cfg1_map
    cfg = m.cfg
    o, lt, le, r = _both(m, m.camera(), cfg.robot_radius)
    P = {"root": cfg.root, "vicinity": cfg.vicinity, "extension_range": cfg.extension_range, "overshoot": cfg.overshoot,
         "robot_radius": cfg.robot_radius, "v_max": 1.0, "dyaw_max": 0.5, "zero_gain": 0.0}
    rng = np.random.default_rng(31)
    nodes = np.array([[*cfg.root, 0.3], [cfg.root[0] + 0.5, cfg.root[1] - 0.2, cfg.root[2], -1.0]])
    nd = np.array([0.0, 0.55])
    u = rng.random((240 if which == "tiny" else 96, 3))
    a = o.expand_tree(lt, le, P, nodes, nd, u, threads=8, sequential=True)
    b = r.iterate_stream(P, nodes, nd, u)
    assert np.array_equal(a["status"], b["status"]) and a["best_index"] == b["best_index"] and a["best_index"] >= 0
    made = a["status"] == 2
    for k in ("pos", "parent", "gain", "yaw", "distance", "score"):
        assert np.array_equal(a[k][made], b[k][made]), k
    assert made.sum() >= 20 and (a["status"] == 1).sum() >= 1 and (a["status"] == 0).sum() >= 1
    assert (a["parent"][made] >= len(nodes)).sum() >= 10      # parents created earlier in the same stream
    # and the independent mode differs: the intra-batch dependence is real
    ind = o.expand_tree(lt, le, P, nodes, nd, u, threads=8, sequential=False)
    assert not np.array_equal(ind["parent"][made], a["parent"][made])
