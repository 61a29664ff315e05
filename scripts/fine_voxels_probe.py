"""Fine voxels under a flat exploration box (14x12x3 m office world, 0.10 m voxels, 5 m range, 500 candidates): which ray
sweep the library picks and what each alternative costs -- the z-clipped dense volume (256-thread CTAs), the half-SM
dense volume (512-thread CTAs), the slot-grid sweep.  Results are identical by construction (tests/test_gpu_zclip.py).

  python scripts/fine_voxels_probe.py
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(env):
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu, synth
    for k in ("NBVGPU_NO_ZCLIP", "NBVGPU_NO_WIDE_CTA"):
        os.environ.pop(k, None)
    for k in env:
        os.environ[k] = "1"
    cfg = synth.SynthConfig("fine", (0, 0, 0), (14, 12, 3), 0.10, 5.0, 500, "office", 3101, 3102, (7.0, 6.0, 1.4), obs_radius=5.0,
                            traj_points=5, traj_step=2.5, vicinity=4.0)
    m = synth.make_map(cfg)
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
    g.layer_update(lt, m.keys, m.tsdf)
    g.commit()
    g.set_camera(m.camera())
    pos = synth.make_candidates(cfg, 500, None).positions
    for _ in range(3):
        r = g.gain_eval(lt, pos, steps=True)
    g.set_profiling(True)
    g.profile_collect()
    t0 = time.perf_counter()
    for _ in range(10):
        r = g.gain_eval(lt, pos, steps=True)
    dt = (time.perf_counter() - t0) / 10
    kms, kn = g.profile_collect()
    out = {"env": sorted(env), "ray_sweep_ms": kms / kn, "voxel_steps_per_s": float(r["steps"].sum()) / (kms / kn * 1e-3),
           "viewpoints_per_s_host_buffers": len(pos) / dt, "dense_clipped_launches": g.dense_clipped_launches(),
           "wide_cta_launches": g.wide_cta_launches(), "gain_checksum": float(r["gain"].sum())}
    g.close()
    return out


if __name__ == "__main__":
    for env in ([], ["NBVGPU_NO_ZCLIP"], ["NBVGPU_NO_ZCLIP", "NBVGPU_NO_WIDE_CTA"]):
        print(json.dumps(run(env)))
