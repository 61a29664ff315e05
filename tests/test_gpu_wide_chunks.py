"""Yaw chunks wider than 40 per CTA: chosen by the library for large batches on geometries with few ray rows per yaw (the
reference's shipped indoor set-up: 69x42 deg, 3 m, 0.3 m voxels -> 8 rows), and forced through NBVGPU_YC for any geometry.
Counts, steps, yaw and gain must equal the oracle's for every width."""
import numpy as np
import pytest

from conftest import load_gpu, load_oracle

pytestmark = pytest.mark.gpu


def _shipped_like():
    from nbv_exploration_planner_b200 import synth
    cfg = synth.SynthConfig("shipped-indoor", (-4, -1, 0), (10, 14, 3), 0.3, 3.0, 1000, "pillars", 4001, 4002, (3.0, 6.0, 1.0), obs_radius=5.0,
                            traj_points=8, traj_step=3.0, vicinity=6.0, hfov_deg=69.0, vfov_deg=42.0, pitch_deg=0.0, near_m=0.1)
    return synth.make_map(cfg)


@pytest.mark.parametrize("variant", [0, 2])
def test_large_batch_on_few_rows_geometry(oracle_mod, variant):
    from nbv_exploration_planner_b200 import synth
    m = _shipped_like()
    cam = m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    g, lt_g, le_g = load_gpu(m, cam, variant)
    pos = synth.make_candidates(m.cfg, 1000, None).positions
    ro = o.gain_eval(lt_o, pos, per_yaw=True, threads=16)
    rg = g.gain_eval(lt_g, pos, per_yaw=True, steps=True)
    for k in ("per_yaw", "steps", "yaw_deg", "gain"):
        assert np.array_equal(rg[k], ro[k]), k
    assert len(np.unique(ro["yaw_deg"])) > 10 and ro["gain"].max() > 0
    if variant == 2:
        assert g.classifier_mismatches() == 0
    g.close()


@pytest.mark.parametrize("yc", [120, 90, 72, 60, 45, 36])
def test_forced_widths_on_tiny_map(oracle_mod, tiny_map, yc, monkeypatch):
    from nbv_exploration_planner_b200 import synth
    cam = tiny_map.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, tiny_map, cam)
    pos = synth.make_candidates(tiny_map.cfg, 24, None).positions
    ro = o.gain_eval(lt_o, pos, per_yaw=True, threads=8)
    monkeypatch.setenv("NBVGPU_YC", str(yc))
    for variant in (0, 2, 4):
        g, lt_g, le_g = load_gpu(tiny_map, cam, variant)
        rg = g.gain_eval(lt_g, pos, per_yaw=True, steps=True)
        for k in ("per_yaw", "steps", "yaw_deg", "gain"):
            assert np.array_equal(rg[k], ro[k]), (k, variant)
        if variant == 2:
            assert g.classifier_mismatches() == 0
        g.close()
