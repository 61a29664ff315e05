"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): map upload, remove, gain (dense, grid and
probe paths), collision, best selection, and the small-batch paths (fused window scan, one warp per segment, eight
lanes per interpolation query, mapped in-place I/O, completion ticket)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, synth  # noqa: E402

m = synth.make_map("tiny")
cfg = m.cfg
g = NbvGpu(0)
lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, 64)
le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, 64)
g.layer_update(lt, m.keys, m.tsdf)
g.layer_update(le, m.keys, m.esdf)
g.commit()
g.set_camera(m.camera())
c = synth.make_candidates(cfg, 4, None)
for variant in (0, 1, 2, 3, 4, 5):
    g.set_kernel_variant(variant)
    r = g.gain_eval(lt, c.positions, per_yaw=True)
free = g.check_segments(le, cfg.robot_radius, c.segments)
best = g.gain_eval_best(lt, c.positions, np.zeros(4), c.distance, free, 1.0, 0.5, 0.0)
g.layer_remove(lt, m.keys[:3])
g.commit()
g.set_kernel_variant(0)
r2 = g.gain_eval(lt, c.positions)
# long range on a coarse grid -> probe path (no shared-memory grid)
cam = m.camera()
cam.far_m = 60.0
g.set_camera(cam)
r3 = g.gain_eval(lt, c.positions[:1])
# small-batch paths on the default camera: single calls and batches below 64, dense sweep with the fused finish
g.set_camera(m.camera())
c2 = synth.make_candidates(cfg, 12, None)
for variant in (0, 2):
    g.set_kernel_variant(variant)
    for i in range(3):
        st = np.array([*c2.positions[i], 0.0])
        g.optimized_gain(lt, st)
    g.gain_eval(lt, c2.positions, steps=False)
    g.gain_eval(lt, c2.positions[:5], steps=True)
assert g.fused_launches() >= 8, g.fused_launches()
for i in range(3):
    g.check_motion(le, cfg.robot_radius, c2.segments[i, :3], c2.segments[i, 3:])
g.check_segments(le, 0.05, np.concatenate([c2.segments[:6, :3], c2.segments[6:12, :3]], axis=1))   # long segments
q = np.random.default_rng(3).uniform(cfg.bounds_min, cfg.bounds_max, (70, 3))
g.esdf_distance_gradient(le, q[:1])
g.esdf_distance_gradient(le, q[:9])
g.esdf_distance_gradient(le, q)
print("sanitize case ok", r["gain"], free, best["index"], r2["gain"], r3["gain"], g.classifier_mismatches())
g.close()
