"""The z-clipped dense shared-memory volume of the ray sweep: for fine voxels the frustum reaches far below / above the
exploration box, the unclipped volume does not fit the shared-memory budget, and the kernel keeps only the rows of the box slab
(+- 3 voxels) plus the rows the rays start in.  Results must be identical to the oracle, the cross-checking variant must report
no mismatch, and the clipped path must really have been taken."""
import numpy as np
import pytest

from conftest import load_gpu, load_oracle

pytestmark = pytest.mark.gpu


def _fine_map():
    from nbv_exploration_planner_b200 import synth
    # 0.10 m voxels, 5 m range: the unclipped volume of a yaw chunk needs more than the 55 KB budget; the box is 3 m high
    cfg = synth.SynthConfig("fine", (0, 0, 0), (14, 12, 3), 0.10, 5.0, 10, "office", 3101, 3102, (7.0, 6.0, 1.4), obs_radius=5.0,
                            traj_points=5, traj_step=2.5, vicinity=4.0)
    return synth.make_map(cfg)


@pytest.mark.parametrize("variant", [0, 2])
def test_fine_voxels_take_the_clipped_volume(oracle_mod, variant, monkeypatch):
    from nbv_exploration_planner_b200 import synth
    m = _fine_map()
    cam = m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam, variant)
    c = synth.make_candidates(m.cfg, 10, None)
    lo, hi = np.array(m.cfg.bounds_min, float), np.array(m.cfg.bounds_max, float)
    # candidates at different heights, one just under the floor and one just above the ceiling of the box (their rays start
    # outside the slab), one far outside
    extra = np.array([[7.0, 6.0, 0.05], [7.0, 6.0, 2.95], [7.0, 6.0, -0.4], [7.0, 6.0, 3.3], [3.0, 3.0, 1.0], [7.0, 6.0, 9.0]])
    pos = np.concatenate([c.positions, extra])
    ro = o.gain_eval(lt_o, pos, per_yaw=True, threads=8)
    rg = g.gain_eval(lt_g, pos, per_yaw=True, steps=True)
    assert g.dense_clipped_launches() >= 1, "the clipped dense path was not taken"
    for k in ("per_yaw", "steps", "yaw_deg", "gain"):
        assert np.array_equal(rg[k], ro[k]), k
    if variant == 2:
        assert g.classifier_mismatches() == 0
    # the same through the slot-grid kernel (clipping disabled): identical again
    monkeypatch.setenv("NBVGPU_NO_ZCLIP", "1")
    before = g.dense_clipped_launches()
    rn = g.gain_eval(lt_g, pos, per_yaw=True, steps=True)
    assert g.dense_clipped_launches() == before
    for k in ("per_yaw", "steps", "yaw_deg", "gain"):
        assert np.array_equal(rn[k], ro[k]), k
    g.close()


def test_expand_tree_with_clipped_volume(oracle_mod):
    """nbvgpu_expand_tree keeps its working set in the context's scratch buffer while the sweep runs: the z-range read-back of the
    clipped path must not touch it (it has a buffer of its own)."""
    from test_gpu_expand import _check, _tree
    m = _fine_map()
    cfg, cam = m.cfg, m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam)
    rng = np.random.default_rng(8)
    state, dist = _tree(cfg, o, le_o, lt_o, 16, rng)
    params = {"root": cfg.root, "vicinity": cfg.vicinity, "extension_range": cfg.extension_range, "overshoot": cfg.overshoot,
              "robot_radius": cfg.robot_radius, "v_max": 1.0, "dyaw_max": 0.75, "zero_gain": 0.0}
    u = rng.random((80, 3))
    ro = o.expand_tree(lt_o, le_o, params, state, dist, u, threads=16)
    rg = g.expand_tree(lt_g, le_g, params, state, dist, u)
    assert g.dense_clipped_launches() >= 1, "the clipped dense path was not taken"
    _check(rg, ro)
    assert (ro["status"] == 2).sum() > 5
    g.close()


@pytest.mark.parametrize("variant", [0, 2])
def test_tall_box_takes_the_half_sm_volume(oracle_mod, variant):
    """Fine voxels under a TALL box (config 3's situation): nothing to clip in z, the volume of a yaw chunk needs more than
    a quarter of an SM's shared memory and runs as 512-thread CTAs, two per SM.  Identical to the oracle; NBVGPU_NO_WIDE_CTA
    (the slot-grid sweep on the same input) gives the same numbers."""
    from nbv_exploration_planner_b200 import synth
    cfg = synth.SynthConfig("fine-tall", (0, 0, 0), (14, 12, 8), 0.10, 5.0, 12, "office", 3201, 3202, (7.0, 6.0, 4.0), obs_radius=5.0,
                            traj_points=5, traj_step=2.5, vicinity=4.0)
    m = synth.make_map(cfg)
    cam = m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam, variant)
    c = synth.make_candidates(m.cfg, 12, None)
    extra = np.array([[7.0, 6.0, 0.3], [7.0, 6.0, 7.8], [7.0, 6.0, -0.4], [3.0, 3.0, 4.0], [13.9, 11.9, 4.0], [7.0, 6.0, 12.0]])
    pos = np.concatenate([c.positions, extra])
    ro = o.gain_eval(lt_o, pos, per_yaw=True, threads=8)
    rg = g.gain_eval(lt_g, pos, per_yaw=True, steps=True)           # 18 candidates: small batch, not fused (wide CTAs)
    big = np.tile(pos, (8, 1))                                        # 144 candidates: large batch
    rb = g.gain_eval(lt_g, big, per_yaw=True, steps=True)
    assert g.wide_cta_launches() >= 2 and g.dense_clipped_launches() == 0 and g.fused_launches() == 0
    for k in ("per_yaw", "steps", "yaw_deg", "gain"):
        assert np.array_equal(rg[k], ro[k]), k
        assert np.array_equal(rb[k], np.concatenate([ro[k]] * 8)), k
    st = np.array([*pos[0], 0.0])
    assert g.optimized_gain(lt_g, st) == ro["gain"][0] and st[3] == ro["yaw_deg"][0] * np.pi / 180.0
    if variant == 2:
        assert g.classifier_mismatches() == 0
    g.close()
