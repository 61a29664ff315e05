"""One data point for the geometry of the reference's shipped indoor configuration (nbveplanner/resource/exploration_indoor.yaml,
launch/voxblox_exploration_indoor.launch): d435 camera 69 x 42 deg, gain range 3.0 m, low-resolution TSDF 0.3 m voxels in 16^3 blocks,
exploration box (-4,-1,0)..(10,14,3) -- on a synthetic pillars world of that size.  Prints one JSON line: viewpoints/s for a batch of
candidates through the host-buffer C ABI and with buffers in HBM, and the reference's own optimized_gain on a bounded sample."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    import torch
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu, synth
    cfg = synth.SynthConfig("shipped-indoor", (-4, -1, 0), (10, 14, 3), 0.3, 3.0, n, "pillars", 4001, 4002, (3.0, 6.0, 1.0), obs_radius=5.0,
                            traj_points=8, traj_step=3.0, vicinity=6.0, hfov_deg=69.0, vfov_deg=42.0, pitch_deg=0.0, near_m=0.1)
    m = synth.make_map(cfg)
    cam = m.camera()
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
    g.layer_update(lt, m.keys, m.tsdf)
    g.commit()
    g.set_camera(cam)
    pos = synth.make_candidates(cfg, n, None).positions
    for _ in range(3):
        r = g.gain_eval(lt, pos, steps=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        r = g.gain_eval(lt, pos, steps=True)
    dt = (time.perf_counter() - t0) / 20
    out = {"metric": "candidate viewpoints/s, shipped indoor geometry", "value": n / dt, "unit": "viewpoints/s", "ms_per_step": dt * 1e3,
           "timing": "host wall clock around nbvgpu_gain_eval (host buffers)", "n_candidates": n,
           "voxel_steps_per_viewpoint": float(r["steps"].mean()), "voxel_steps_per_s": float(r["steps"].sum()) / dt,
           "config": {"workload": "69x42 deg, 3.0 m range, 0.3 m voxels, 16^3 blocks, box (-4,-1,0)..(10,14,3), pillars world", "map_blocks": int(len(m.keys))}}
    t1 = time.perf_counter()
    for i in range(100):
        g.gain_eval(lt, pos[i:i + 1])
    out["single_candidate_latency_us"] = (time.perf_counter() - t1) / 100 * 1e6
    from oracle import ref_binding as rb
    if rb.available():
        ref = rb.Reference(m.voxel_size, m.vps)
        ref.tsdf_update(m.keys, m.tsdf)
        ref.set_camera(cam.as_dict(), 0.4)
        k = min(n, 200)
        t1 = time.perf_counter()
        rr = ref.gain_eval(pos[:k])
        t_ref = time.perf_counter() - t1
        out["cpu_baseline"] = {"kind": "reference", "value": k / t_ref, "unit": "viewpoints/s", "cores": 1, "sample": f"first {k} candidates, 1 thread",
                               "gpu_equals_reference_on_sample": bool(np.array_equal(rr["gain"], r["gain"][:k]))}
    print(json.dumps(out))
    g.close()


if __name__ == "__main__":
    main()
