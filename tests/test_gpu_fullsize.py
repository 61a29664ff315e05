"""Parity at BASELINE.json's full sizes through properties that do not need the (slow) oracle on every
candidate: (i) two independent decision paths on the GPU -- the plain FP64 transcription kernel
(variant 1) and the default kernel -- agree on every candidate; (ii) a random subset agrees bit-exactly
with the oracle; (iii) the cross-checking kernel variants report zero mismatches; (iv) an incremental
replay (changed-block uploads, removals) tracks an oracle map updated the same way."""
import numpy as np
import pytest

from conftest import load_gpu, load_oracle

pytestmark = pytest.mark.gpu


def _subset_vs_oracle(o, lt_o, g, lt_g, pos, k, seed, base=None):
    """`base`: the result of the FULL batch (the launch shape the bench times): its rows for the subset are compared with
    the oracle as they are, and so is a re-run of the subset alone (a smaller batch, i.e. another launch shape)."""
    sel = np.random.default_rng(seed).choice(len(pos), size=min(k, len(pos)), replace=False)
    ro = o.gain_eval(lt_o, pos[sel], per_yaw=True, threads=16)
    rg = g.gain_eval(lt_g, pos[sel], per_yaw=True)
    assert np.array_equal(rg["per_yaw"], ro["per_yaw"]) and np.array_equal(rg["steps"], ro["steps"])
    assert np.array_equal(rg["yaw_deg"], ro["yaw_deg"]) and np.array_equal(rg["gain"], ro["gain"])
    if base is not None:
        for key in ("per_yaw", "steps", "yaw_deg", "gain"):
            assert np.array_equal(base[key][sel], ro[key]), key


def test_cfg2_1000_candidates(oracle_mod, cfg1_map):
    """BASELINE config 2 (the bench workload): 1,000 candidates + their edges."""
    from nbv_exploration_planner_b200 import synth
    cfg = synth.CONFIGS["cfg2"]
    cam = cfg1_map.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, cfg1_map, cam)
    g, lt_g, le_g = load_gpu(cfg1_map, cam)
    c = synth.make_candidates(cfg, 1000, lambda s: g.check_segments(le_g, cfg.robot_radius, s))
    assert np.array_equal(g.check_segments(le_g, cfg.robot_radius, c.segments),
                          o.check_segments(le_o, cfg.robot_radius, c.segments, threads=16)["free"])
    base = g.gain_eval(lt_g, c.positions, per_yaw=True)
    assert g.wide_cta_launches() >= 1              # the launch the bench times: 90 yaws per 512-thread CTA, dense volume
    for variant in (1, 2, 4):
        g.set_kernel_variant(variant)
        r = g.gain_eval(lt_g, c.positions, per_yaw=True)
        for k in ("per_yaw", "gain", "yaw_deg", "steps", "counts"):
            assert np.array_equal(r[k], base[k]), (variant, k)
    assert g.classifier_mismatches() == 0
    g.set_kernel_variant(0)
    _subset_vs_oracle(o, lt_o, g, lt_g, c.positions, 96, 1, base)
    # best viewpoint: device scoring == host replay of rrt.cpp:220-246 on the same gains
    best = g.gain_eval_best(lt_g, c.positions, np.zeros(1000), c.distance, None, 1.0, 0.5, 0.0)
    score = base["gain"] / np.maximum(base["yaw"] / 0.5, c.distance / 1.0)
    assert best["index"] == int(np.argmax(score)) and score[best["index"]] > 0
    g.close()


def test_cfg3_office_map_subset(oracle_mod):
    """BASELINE config 3 geometry (50x50x10 m office-like, 0.10 m voxels): 2,000 candidates on the GPU,
    variants agree on 400, 200 rows of the full batch checked against the oracle."""
    from nbv_exploration_planner_b200 import synth
    m = synth.make_map("cfg3")
    cfg = m.cfg
    cam = m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam)
    c = synth.make_candidates(cfg, 2000, lambda s: g.check_segments(le_g, cfg.robot_radius, s))
    base = g.gain_eval(lt_g, c.positions, per_yaw=True)
    assert g.wide_cta_launches() >= 1              # config 3's launch: half-SM dense volume in 512-thread CTAs
    assert base["steps"].mean() > 1.0e6            # ~1.65 M steps per viewpoint without occlusion
    for variant in (1, 5):
        g.set_kernel_variant(variant)
        r = g.gain_eval(lt_g, c.positions[:400], per_yaw=True)
        for k in ("per_yaw", "gain", "yaw_deg", "steps"):
            assert np.array_equal(r[k], base[k][:400]), (variant, k)
    assert g.classifier_mismatches() == 0
    g.set_kernel_variant(0)
    _subset_vs_oracle(o, lt_o, g, lt_g, c.positions, 200, 2, base)
    seg_sel = np.arange(0, 2000, 4)
    assert np.array_equal(g.check_segments(le_g, cfg.robot_radius, c.segments[seg_sel]),
                          o.check_segments(le_o, cfg.robot_radius, c.segments[seg_sel], threads=16)["free"])
    g.close()


def test_cfg4_full_size_map(oracle_mod):
    """BASELINE config 4 at full size: the 100x100x20 m outdoor map at 0.05 m voxels (143 k blocks, 4.7 GB of TSDF tiles in
    the mirror), 8 m range.  400 candidates in one batch (the slot-grid sweep: the volume of a yaw chunk is megabytes);
    the rows of the four best and eight more candidates of that batch equal the oracle's, and the cross-checking variant
    reports no mismatch on a 24-candidate batch."""
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu, synth
    m = synth.make_map("cfg4")
    cfg = m.cfg
    cam = m.camera()
    assert len(m.keys) > 100000
    o = oracle_mod.Oracle()
    lt_o = o.layer_create(0, m.voxel_size, m.vps)
    o.layer_update(lt_o, m.keys, m.tsdf)
    o.set_camera(cam.as_dict())
    g = NbvGpu(0)
    lt_g = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 64)
    for i in range(0, len(m.keys), 8192):          # bounded pinned staging
        g.layer_update(lt_g, m.keys[i:i + 8192], m.tsdf[i:i + 8192])
        g.commit()
    assert g.layer_num_blocks(lt_g) == len(m.keys)
    g.set_camera(cam)
    pos = synth.make_candidates(cfg, 400, None).positions
    base = g.gain_eval(lt_g, pos, per_yaw=True)
    assert base["steps"].mean() > 1.0e7 and base["gain"].max() > 0
    sel = np.unique(np.concatenate([np.argsort(base["gain"])[-4:], np.random.default_rng(4).choice(400, 8, replace=False)]))
    ro = o.gain_eval(lt_o, pos[sel], per_yaw=True, threads=16)
    for key in ("per_yaw", "steps", "yaw_deg", "gain"):
        assert np.array_equal(base[key][sel], ro[key]), key
    g.set_kernel_variant(5)
    r5 = g.gain_eval(lt_g, pos[:24], per_yaw=True)
    for key in ("per_yaw", "steps", "yaw_deg", "gain"):
        assert np.array_equal(r5[key], base[key][:24]), key
    assert g.classifier_mismatches() == 0
    g.close()


def test_long_range_fine_voxels(oracle_mod):
    """BASELINE config 4 ray geometry (0.05 m voxels, 8 m range: 66,600 rays, ~250 steps per ray) on a
    small map: exercises the slot-grid path (volume too large for shared memory)."""
    from nbv_exploration_planner_b200 import synth
    cfg = synth.SynthConfig("cfg4small", (0, 0, 0), (16, 16, 6), 0.05, 8.0, 6, "pillars", 1204, 2204, (8.0, 8.0, 2.0),
                            obs_radius=7.0, traj_points=3, traj_step=3.0, vicinity=2.0)
    m = synth.make_map(cfg)
    cam = m.camera()
    o, lt_o, _ = load_oracle(oracle_mod, m, cam)
    c = synth.make_candidates(cfg, 6, None)
    for variant in (0, 2):
        g, lt_g, _ = load_gpu(m, cam, variant)
        ro = o.gain_eval(lt_o, c.positions, per_yaw=True, threads=16)
        rg = g.gain_eval(lt_g, c.positions, per_yaw=True)
        assert ro["rays"][0] == 360 * 185
        assert np.array_equal(rg["per_yaw"], ro["per_yaw"]) and np.array_equal(rg["steps"], ro["steps"])
        assert np.array_equal(rg["gain"], ro["gain"]) and np.array_equal(rg["yaw_deg"], ro["yaw_deg"])
        if variant == 2:
            assert g.classifier_mismatches() == 0
        g.close()


def test_slot_grid_limited_to_the_rays_z_extent(oracle_mod, monkeypatch):
    """Config 4's geometry on the slot-grid sweep: the grid spans only the block layers the rays reach (14 of 25), which is
    what lets four CTAs share an SM.  Same results as with the full cube of blocks (NBVGPU_CUBE_GRID) and as the oracle, also
    for candidates next to the floor and the ceiling of the map and for a camera pitched up."""
    from nbv_exploration_planner_b200 import synth
    for pitch in (15.0, -25.0):
        cfg = synth.SynthConfig("cfg4grid", (0, 0, 0), (16, 16, 6), 0.05, 8.0, 8, "pillars", 1205, 2205, (8.0, 8.0, 2.0),
                                obs_radius=7.0, traj_points=3, traj_step=3.0, vicinity=2.0, pitch_deg=pitch)
        m = synth.make_map(cfg)
        cam = m.camera()
        c = synth.make_candidates(cfg, 8, None)
        pos = c.positions.copy()
        pos[0, 2], pos[1, 2] = 0.12, 5.9
        o, lt_o, _ = load_oracle(oracle_mod, m, cam)
        ro = o.gain_eval(lt_o, pos, per_yaw=True, threads=16)
        res = {}
        for cube in (False, True):
            if cube:
                monkeypatch.setenv("NBVGPU_CUBE_GRID", "1")
            else:
                monkeypatch.delenv("NBVGPU_CUBE_GRID", raising=False)
            g, lt_g, _ = load_gpu(m, cam, 5)        # slot-grid sweep with every shortcut cross-checked
            res[cube] = g.gain_eval(lt_g, pos, per_yaw=True)
            assert g.classifier_mismatches() == 0
            g.close()
        monkeypatch.delenv("NBVGPU_CUBE_GRID", raising=False)
        for k in ("per_yaw", "steps", "gain", "yaw_deg"):
            assert np.array_equal(res[False][k], res[True][k]), k
            assert np.array_equal(res[False][k], ro[k]), k


def test_incremental_replay_tracks_oracle(oracle_mod):
    """BASELINE config 5 in miniature: per step upload only the changed blocks (sphere around the next
    trajectory point), drop some stale blocks, evaluate; the GPU mirror must track an oracle layer that
    received the same updates."""
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, synth
    cfg = synth.SynthConfig("replay", (0, 0, 0), (12, 12, 4), 0.2, 3.0, 8, "pillars", 1300, 2300, (6.0, 6.0, 1.5),
                            obs_radius=3.0, traj_points=10, traj_step=1.2, vicinity=2.0)
    rng = np.random.Generator(np.random.MT19937(cfg.seed_map))
    objects = synth._world_objects(cfg, rng)
    traj = synth._trajectory(cfg, rng)
    cam = synth.camera_for(cfg)
    o = oracle_mod.Oracle()
    lt_o, le_o = o.layer_create(0, cfg.voxel_size, 16), o.layer_create(1, cfg.voxel_size, 16)
    o.set_camera(cam.as_dict())
    g = NbvGpu(0)
    lt_g, le_g = g.layer_create(LAYER_TSDF, cfg.voxel_size, 16, 512), g.layer_create(LAYER_ESDF, cfg.voxel_size, 16, 512)
    g.set_camera(cam)
    off = np.random.default_rng(0).uniform(-1.5, 1.5, (8, 3)) * np.array([1, 1, 0.3])
    known = set()
    for k in range(len(traj)):
        keys = synth.blocks_touching(cfg.voxel_size, 16, traj[k:k + 1], cfg.obs_radius, cfg.bounds_min, cfg.bounds_max)
        tsdf, esdf = synth.generate_blocks(cfg, keys, objects, traj[:k + 1])
        o.layer_update(lt_o, keys, tsdf); o.layer_update(le_o, keys, esdf)
        g.layer_update(lt_g, keys, tsdf); g.layer_update(le_g, keys, esdf)
        known.update(map(tuple, keys))
        if k % 3 == 2:   # forget a few far-away blocks (Layer::removeBlock)
            far = np.array([b for b in known if np.linalg.norm((np.array(b) + 0.5) * 3.2 - traj[k]) > 5.0][:5], np.int32).reshape(-1, 3)
            if len(far):
                o.layer_remove(lt_o, far); g.layer_remove(lt_g, far)
                known.difference_update(map(tuple, far))
        g.commit()
        assert g.layer_num_blocks(lt_g) == o.num_blocks(lt_o) == len(known)
        pos = np.clip(traj[k] + off, [0.3, 0.3, 0.3], [11.7, 11.7, 3.7])
        ro = o.gain_eval(lt_o, pos, per_yaw=True, threads=8)
        rg = g.gain_eval(lt_g, pos, per_yaw=True)
        assert np.array_equal(rg["per_yaw"], ro["per_yaw"]) and np.array_equal(rg["gain"], ro["gain"])
        seg = np.concatenate([np.broadcast_to(traj[k], pos.shape), pos], 1)
        assert np.array_equal(g.check_segments(le_g, cfg.robot_radius, seg), o.check_segments(le_o, cfg.robot_radius, seg)["free"])
    g.close()
