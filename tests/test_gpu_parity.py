"""GPU parity tests proper: the CUDA path behind the C ABI vs the CPU oracle on identical seeded
inputs.  Bar (BASELINE.json north_star): per-viewpoint / per-yaw unknown, free and occupied counts
bit-exact, same best yaw, same best viewpoint, floating-point gain within 1e-6 relative."""
import math

import numpy as np
import pytest

from conftest import load_gpu, load_oracle

pytestmark = pytest.mark.gpu

GAIN_RTOL = 1e-6  # north_star tolerance for the floating-point gain


def _compare_gain(o, lt_o, g, lt_g, pos, expect_exact_gain=True):
    ro = o.gain_eval(lt_o, pos, per_yaw=True, threads=8)
    rg = g.gain_eval(lt_g, pos, per_yaw=True, steps=True)
    assert np.array_equal(rg["per_yaw"], ro["per_yaw"]), "per-yaw unknown/free/occupied counts differ"
    assert np.array_equal(rg["steps"], ro["steps"]), "executed voxel-steps differ"
    assert np.array_equal(rg["yaw_deg"], ro["yaw_deg"]), "best yaw differs"
    assert np.array_equal(rg["counts"], ro["counts"])
    if expect_exact_gain:
        assert np.array_equal(rg["gain"], ro["gain"])
    else:
        np.testing.assert_allclose(rg["gain"], ro["gain"], rtol=GAIN_RTOL, atol=0)
    return ro, rg


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_tiny_map_all_kernel_variants(oracle_mod, tiny_map, variant):
    from nbv_exploration_planner_b200 import synth
    cam = tiny_map.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, tiny_map, cam)
    g, lt_g, le_g = load_gpu(tiny_map, cam, variant)
    c = synth.make_candidates(tiny_map.cfg, 16, lambda s: o.check_segments(le_o, tiny_map.cfg.robot_radius, s)["free"])
    _compare_gain(o, lt_o, g, lt_g, c.positions)
    if variant in (2, 5):
        assert g.classifier_mismatches() == 0
    g.close()


@pytest.mark.parametrize("variant", [0, 1, 2, 4, 5])
def test_cfg1_100_candidates(oracle_mod, cfg1_map, variant):
    """BASELINE config 1: 20x20x5 m, 0.15 m voxels, 90x60 deg, 5 m range, 100 candidates."""
    from nbv_exploration_planner_b200 import synth
    cam = cfg1_map.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, cfg1_map, cam)
    g, lt_g, le_g = load_gpu(cfg1_map, cam, variant)
    cfg = cfg1_map.cfg
    c = synth.make_candidates(cfg, 100, lambda s: o.check_segments(le_o, cfg.robot_radius, s)["free"])
    ro, rg = _compare_gain(o, lt_o, g, lt_g, c.positions)
    assert ro["rays"][0] == 360 * 39 and ro["steps"].sum() > 50e6
    if variant in (2, 5):
        assert g.classifier_mismatches() == 0
    st = g.stats()
    assert st["voxel_steps"] == int(ro["steps"].sum()) and st["rays"] == int(ro["rays"].sum())
    g.close()


def test_cfg1_golden_fixture(cfg1_map):
    """CUDA path vs the committed golden vectors (oracle output, tests/golden/make_golden.py)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1_golden.npz"))
    assert cfg1_map.sha256() == str(gold["map_sha256"])
    g, lt, le = load_gpu(cfg1_map, cfg1_map.camera())
    r = g.gain_eval(lt, gold["positions"], per_yaw=True)
    assert np.array_equal(r["per_yaw"], gold["per_yaw"])
    assert np.array_equal(r["yaw_deg"], gold["yaw_deg"])
    assert np.array_equal(r["gain"], gold["gain"])
    free = g.check_segments(le, cfg1_map.cfg.robot_radius, gold["segments"])
    assert np.array_equal(free, gold["seg_free"])
    g.close()


@pytest.mark.parametrize("hfov,vfov,far,pitch,vs,vps", [
    (69.0, 42.0, 3.0, 0.0, 0.3, 16),     # shipped yaml: d435, low-res 0.3 m, range 3 (resource/exploration_indoor.yaml)
    (120.0, 90.0, 2.5, -20.0, 0.2, 8),   # wide fov, looking up, 8 voxels per side (test_tsdf_map.cc block size)
    (45.0, 30.0, 6.0, 35.0, 0.25, 32),   # narrow fov, steep pitch, 32 voxels per side
    (90.0, 60.0, 4.0, 15.0, 0.1, 16),
])
def test_camera_and_layer_parameter_sweep(oracle_mod, hfov, vfov, far, pitch, vs, vps):
    from nbv_exploration_planner_b200 import synth
    cfg = synth.SynthConfig("sweep", (-6, -5, -1), (6, 7, 3), vs, far, 12, "pillars", 77, 78, (0.5, 0.3, 0.9), obs_radius=4.0,
                            traj_points=5, traj_step=2.0, vicinity=3.0, hfov_deg=hfov, vfov_deg=vfov, pitch_deg=pitch, vps=vps)
    m = synth.make_map(cfg)   # negative coordinates: blocks with negative indices
    cam = m.camera()
    o, lt_o, _ = load_oracle(oracle_mod, m, cam)
    for variant in (0, 2, 5):
        g, lt_g, _ = load_gpu(m, cam, variant)
        c = synth.make_candidates(cfg, 12, None)
        _compare_gain(o, lt_o, g, lt_g, c.positions)
        if variant in (2, 5):
            assert g.classifier_mismatches() == 0
        g.close()


def test_axis_aligned_and_grid_aligned_positions(oracle_mod, tiny_map):
    """Positions on voxel corners / block corners and with a zero extrinsic translation: the
    degenerate DDA (zero ray components at yaw 0/+-90/180) and the plane-boundary cases."""
    from nbv_exploration_planner_b200.planner import CameraParams
    vs = float(tiny_map.voxel_size)
    cam = CameraParams(hfov_deg=90, vfov_deg=60, near_m=0.1, far_m=3.0, bbx_min=(0, 0, 0), bbx_max=(8, 8, 3))  # identity T_C_B
    o, lt_o, _ = load_oracle(oracle_mod, tiny_map, cam)
    g, lt_g, _ = load_gpu(tiny_map, cam, 2)
    pos = np.array([[4.0, 4.0, 1.2], [16 * vs, 16 * vs, 8 * vs], [3.2, 3.2, 1.6], [4.0 + 0.5 * vs, 4.0 + 0.5 * vs, 1.0 + 0.5 * vs],
                    [0.0, 0.0, 0.0], [8.0, 8.0, 3.0], [1e-9, 4.0, 1.0]])
    _compare_gain(o, lt_o, g, lt_g, pos)
    assert g.classifier_mismatches() == 0
    g.close()


@pytest.mark.parametrize("variant", [0, 1, 2, 4])
@pytest.mark.parametrize("state", ["unknown", "free", "occupied"])
def test_rays_stuck_at_yaw_plus_minus_90(oracle_mod, state, variant):
    """Integer yaws +-90 deg give rays with no x component: the reference's RayCaster divides by that zero
    (t = -inf or 0/0, dt = 0/0), minCoeff then picks the axis for ever and the SAME voxel -- the one under the
    camera -- is emitted len + 1 times (integrator_utils.cc:119-123,165-179).  The kernel evaluates that in
    closed form.  Here that voxel is unknown / free / occupied; its min corner lies inside the frustum for
    yaw -90 (camera near the +y face of the voxel, looking along -y) and behind the camera for yaw +90."""
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu
    from nbv_exploration_planner_b200.planner import CameraParams
    vs, vps = 0.2, 16
    cam = CameraParams(hfov_deg=90, vfov_deg=60, near_m=0.01, far_m=2.0, bbx_min=(-4, -4, -4), bbx_max=(8, 8, 8))  # identity T_C_B
    keys = np.array([[0, 0, 0]], np.int32)
    tsdf = np.zeros((1, vps ** 3, 2), np.float32)  # weight 0: unknown
    lin = 5 + 5 * vps + 5 * vps * vps
    if state == "free":
        tsdf[0, lin] = (0.4, 1.0)
    elif state == "occupied":
        tsdf[0, lin] = (-0.05, 1.0)
    pos = np.array([[(5 + 0.02) * vs, (5 + 0.9) * vs, (5 + 0.03) * vs]])
    o = oracle_mod.Oracle()
    lt_o = o.layer_create(0, vs, vps)
    o.layer_update(lt_o, keys, tsdf)
    o.set_camera(cam.as_dict())
    g = NbvGpu(0)
    lt_g = g.layer_create(LAYER_TSDF, vs, vps, 8)
    g.layer_update(lt_g, keys, tsdf)
    g.commit()
    g.set_camera(cam)
    g.set_kernel_variant(variant)
    ro, rg = _compare_gain(o, lt_o, g, lt_g, pos)
    if variant == 2:
        assert g.classifier_mismatches() == 0
    # the signature of a stuck ray, so that the test cannot pass without exercising it
    looking_minus_y, looking_plus_y = ro["per_yaw"][0, 90], ro["per_yaw"][0, 270]
    assert tuple(looking_plus_y) == (0, 0, 0)            # the voxel's corner is behind the camera: skipped every time
    u_max = int(ro["rays"][0]) // 360
    if state == "unknown":
        assert looking_minus_y[0] > u_max and looking_minus_y[1] == 0 and looking_minus_y[2] == 0
    elif state == "free":
        assert looking_minus_y[0] == 0 and looking_minus_y[1] > u_max and looking_minus_y[2] == 0
    else:
        assert tuple(looking_minus_y) == (0, 0, u_max)   # every ray of the column stops on its first voxel
    g.close()


@pytest.mark.parametrize("variant", [0, 2, 3])
def test_very_long_range_uses_global_probes(oracle_mod, tiny_map, variant):
    """A reach too large for the shared-memory slot grid (60 m range on 0.2 m voxels: 346 ray rows,
    ~125 k rays per viewpoint) takes the global-probe path (MODE kProbe); rays leave the map and the
    exploration box long before they end."""
    cam = tiny_map.camera()
    cam.far_m = 60.0
    o, lt_o, _ = load_oracle(oracle_mod, tiny_map, cam)
    g, lt_g, _ = load_gpu(tiny_map, cam, variant)
    pos = np.array([[4.0, 4.0, 1.2], [2.3, 5.1, 0.9]])
    ro, rg = _compare_gain(o, lt_o, g, lt_g, pos)
    assert ro["rays"][0] == 360 * 347 or ro["rays"][0] > 100000
    if variant == 2:
        assert g.classifier_mismatches() == 0
    g.close()


def test_sum_all_yaws_mode_is_reference_gain(oracle_mod, tiny_map):
    """camera.sum_all_yaws = 1 == RrtTree::gain (rrt.cpp:481-579): total over every yaw, yaw = 0."""
    from nbv_exploration_planner_b200 import synth
    cam = tiny_map.camera(sum_all_yaws=1)
    o, lt_o, _ = load_oracle(oracle_mod, tiny_map, cam)
    g, lt_g, _ = load_gpu(tiny_map, cam)
    c = synth.make_candidates(tiny_map.cfg, 8, None)
    ro, rg = _compare_gain(o, lt_o, g, lt_g, c.positions)
    assert np.all(rg["yaw_deg"] == 0)
    assert np.array_equal(rg["counts"], rg["per_yaw"].astype(np.uint64).sum(1))
    g.close()


def test_non_default_gain_weights_within_tolerance(oracle_mod, tiny_map):
    from nbv_exploration_planner_b200 import synth
    cam = tiny_map.camera(gain_free=0.1, gain_occupied=0.37, gain_unmapped=1.3)
    o, lt_o, _ = load_oracle(oracle_mod, tiny_map, cam)
    g, lt_g, _ = load_gpu(tiny_map, cam)
    c = synth.make_candidates(tiny_map.cfg, 8, None)
    _compare_gain(o, lt_o, g, lt_g, c.positions, expect_exact_gain=False)
    g.close()


def test_empty_map_everything_unknown(oracle_mod):
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu
    from nbv_exploration_planner_b200.planner import CameraParams
    cam = CameraParams.mount(15.0, hfov_deg=90, vfov_deg=60, near_m=0.1, far_m=2.0, bbx_min=(-5, -5, -5), bbx_max=(5, 5, 5))
    o = oracle_mod.Oracle()
    lt_o = o.layer_create(0, 0.15, 16)
    o.set_camera(cam.as_dict())
    g = NbvGpu(0)
    lt_g = g.layer_create(LAYER_TSDF, 0.15, 16, 8)
    g.set_camera(cam)
    pos = np.array([[0.1, 0.2, 0.3], [-1.0, 2.0, 0.0]])
    ro, rg = _compare_gain(o, lt_o, g, lt_g, pos)
    assert rg["per_yaw"][..., 1].sum() == 0 and rg["per_yaw"][..., 2].sum() == 0 and rg["gain"].min() > 0
    g.close()


def test_segment_collision_parity(oracle_mod, cfg1_map):
    """BASELINE config 2: 1,000 tree edges, robot radius = half diagonal of 0.5x0.5x0.3 m, overshoot 0.5 m."""
    from nbv_exploration_planner_b200 import synth
    cfg = synth.CONFIGS["cfg2"]
    cam = cfg1_map.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, cfg1_map, cam)
    g, lt_g, le_g = load_gpu(cfg1_map, cam)
    c = synth.make_candidates(cfg, 1000, None)      # unfiltered: free and colliding edges
    fo = o.check_segments(le_o, cfg.robot_radius, c.segments)
    fg = g.check_segments(le_g, cfg.robot_radius, c.segments)
    assert np.array_equal(fg, fo["free"])
    assert 0 < fo["free"].sum() < 1000
    assert g.stats()["collision_steps"] == int(fo["steps"].sum())
    # long random segments through unknown space and obstacles, different radii
    rng = np.random.default_rng(5)
    seg = np.concatenate([rng.uniform(0, 20, (500, 2)), rng.uniform(0.2, 4.8, (500, 1)),
                          rng.uniform(0, 20, (500, 2)), rng.uniform(0.2, 4.8, (500, 1))], 1)
    for radius in (0.0, 0.2, 0.3841, 1.0, 2.5):
        assert np.array_equal(g.check_segments(le_g, radius, seg), o.check_segments(le_o, radius, seg)["free"])
    # n = 1 shim == HighResManager::checkMotion
    for i in range(5):
        assert g.check_motion(le_g, cfg.robot_radius, c.segments[i, :3], c.segments[i, 3:]) == bool(fo["free"][i])
    g.close()


def test_best_viewpoint_selection(oracle_mod, cfg1_map):
    """Same best viewpoint as the reference's sequential scoring (rrt.cpp:220-246) on oracle gains."""
    from nbv_exploration_planner_b200 import synth
    cfg = cfg1_map.cfg
    cam = cfg1_map.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, cfg1_map, cam)
    g, lt_g, le_g = load_gpu(cfg1_map, cam)
    c = synth.make_candidates(cfg, 100, None)
    free = o.check_segments(le_o, cfg.robot_radius, c.segments)["free"]
    ro = o.gain_eval(lt_o, c.positions, threads=8)
    rng = np.random.default_rng(3)
    parent_yaw = rng.uniform(-math.pi, math.pi, 100)
    v_max, dyaw_max, zero_gain = 1.0, 0.5, 0.0
    best, best_i = zero_gain, -1
    for i in range(100):                      # sequential restatement of the caller
        if not free[i]:
            continue
        dy = ro["yaw"][i] - parent_yaw[i]
        if dy > math.pi:
            dy -= 2.0 * math.pi
        if dy < -math.pi:
            dy += 2.0 * math.pi
        t = max(dy / dyaw_max, c.distance[i] / v_max)
        s = ro["gain"][i] / t
        if s > best:
            best, best_i = s, i
    r = g.gain_eval_best(lt_g, c.positions, parent_yaw, c.distance, free, v_max, dyaw_max, zero_gain)
    assert r["index"] == best_i
    assert best_i >= 0
    assert r["score"] == pytest.approx(best, rel=GAIN_RTOL)
    assert r["best_gain"] == ro["gain"][best_i] and r["best_yaw"] == ro["yaw"][best_i]
    # rrt.cpp:210-216: the gain is only evaluated behind a free edge
    assert np.array_equal(r["gain"], np.where(free == 1, ro["gain"], 0.0)) and np.array_equal(r["yaw_deg"], np.where(free == 1, ro["yaw_deg"], 0))
    # nothing beats a huge zero_gain -> index -1
    r2 = g.gain_eval_best(lt_g, c.positions, parent_yaw, c.distance, free, v_max, dyaw_max, 1e30)
    assert r2["index"] == -1
    g.close()


def test_reference_shaped_wrappers(oracle_mod, tiny_map):
    """RrtGain.optimized_gain / gain and HighResManager.checkMotion behave like the reference members."""
    from nbv_exploration_planner_b200 import HighResManager, LowResManager, NbvGpu, RrtGain
    cam = tiny_map.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, tiny_map, cam)
    gpu = NbvGpu(0)
    low = LowResManager(gpu, tiny_map.voxel_size, tiny_map.vps, 64)
    assert low.isTsdfEmpty()
    low.update(tiny_map.keys, tiny_map.tsdf)
    assert not low.isTsdfEmpty() and low.getVoxelsPerSide() == 16
    high = HighResManager(gpu, tiny_map.voxel_size, tiny_map.cfg.robot_radius, tiny_map.vps, 64)
    high.update(tiny_map.keys, tiny_map.esdf)
    tree = RrtGain(gpu, low, cam)
    state = np.array([4.0, 4.1, 1.3, 123.0])
    gval = tree.optimized_gain(state)
    ro = o.gain_eval(lt_o, state[:3])
    assert gval == ro["gain"][0] and state[3] == ro["yaw"][0]
    o.set_camera({**cam.as_dict(), "sum_all_yaws": 1})
    state2 = np.array([4.0, 4.1, 1.3, 123.0])
    g2 = tree.gain(state2)
    ro2 = o.gain_eval(lt_o, state2[:3])
    assert g2 == ro2["gain"][0] and state2[3] == 0.0
    seg = np.array([4.0, 4.0, 1.2, 5.0, 4.5, 1.4])
    assert high.checkMotion(seg[:3], seg[3:]) == bool(o.check_segments(le_o, tiny_map.cfg.robot_radius, seg)["free"][0])
    # getVoxelStatusByIndex through the mirror
    for b, v in (((1, 1, 0), (3, 4, 5)), ((0, 0, 0), (0, 0, 0)), ((50, 0, 0), (1, 1, 1))):
        gidx = np.array(b, np.int64) * 16 + np.array(v, np.int64)
        assert low.getVoxelStatusByIndex(b, v) == o.voxel_status(lt_o, gidx)
    gpu.close()


def test_error_behaviour(tiny_map):
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, NbvGpuError
    from nbv_exploration_planner_b200 import _native as N
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, 0.2, 16, 4)
    with pytest.raises(NbvGpuError) as e:                      # camera not set
        g.gain_eval(lt, np.zeros((1, 3)))
    assert e.value.code == N.ERR_STATE
    g.set_camera(tiny_map.camera())
    for bad in (np.nan, np.inf, 1e9):
        with pytest.raises(NbvGpuError) as e:                  # non-finite / out of range position
            g.gain_eval(lt, np.array([[bad, 0.0, 0.0]]))
        assert e.value.code == N.ERR_RANGE
    le = g.layer_create(LAYER_ESDF, 0.2, 16, 4)
    with pytest.raises(NbvGpuError) as e:                      # wrong layer kind
        g.gain_eval(le, np.zeros((1, 3)))
    assert e.value.code == N.ERR_STATE
    with pytest.raises(NbvGpuError) as e:
        g.check_segments(lt, 0.3, np.zeros((1, 6)))
    assert e.value.code == N.ERR_STATE
    with pytest.raises(NbvGpuError) as e:                      # tile pool exhausted
        g.layer_update(lt, tiny_map.keys, tiny_map.tsdf)
        g.commit()
    assert e.value.code == N.ERR_CAPACITY
    with pytest.raises(NbvGpuError):
        g.layer_create(LAYER_TSDF, 0.2, 12, 4)                 # voxels_per_side not a power of two (layer.h:38 CHECK)
    bad_cam = tiny_map.camera()
    bad_cam.q_CB = (2.0, 0.0, 0.0, 0.0)
    with pytest.raises(NbvGpuError):
        g.set_camera(bad_cam)
    r = g.gain_eval(lt, np.zeros((0, 3)))                      # empty batch is fine
    assert len(r["gain"]) == 0
    assert len(g.check_segments(le, 0.3, np.zeros((0, 6)))) == 0
    g.close()
