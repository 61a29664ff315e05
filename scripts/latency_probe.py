"""Single-call latency of the drop-in entry points (what the reference's sequential iterate() sees):
n = 1 nbvgpu_gain_eval and nbvgpu_check_segments through host buffers, medians over many calls, with
the ray-sweep kernel's own time (CUDA events) beside them.  NBVGPU_YC=<yaws per CTA> is honoured by
the library; `--yc 0,4,2,1` runs the probe once per value (0 = the library's own choice).

  python scripts/latency_probe.py [--yc 0,4,2] [--calls 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def shipped_cfg(synth, n):
    return synth.SynthConfig("shipped-indoor", (-4, -1, 0), (10, 14, 3), 0.3, 3.0, n, "pillars", 4001, 4002, (3.0, 6.0, 1.0), obs_radius=5.0,
                             traj_points=8, traj_step=3.0, vicinity=6.0, hfov_deg=69.0, vfov_deg=42.0, pitch_deg=0.0, near_m=0.1)


def probe(name, cfg_or_name, calls):
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, synth
    m = synth.make_map(cfg_or_name)
    cfg = m.cfg
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
    le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 8)
    g.layer_update(lt, m.keys, m.tsdf)
    g.layer_update(le, m.keys, m.esdf)
    g.commit()
    g.set_camera(m.camera())
    c = synth.make_candidates(cfg, 64, None)
    pos, seg = c.positions, c.segments
    for i in range(16):
        g.gain_eval(lt, pos[i:i + 1])
        g.check_segments(le, cfg.robot_radius, seg[i:i + 1])
    tg, ts, to, tk = [], [], [], []
    for i in range(calls):   # nbvgpu_gain_eval with counts and steps (the Python wrapper's default)
        p = pos[i % 64:i % 64 + 1]
        t0 = time.perf_counter()
        g.gain_eval(lt, p)
        tg.append(time.perf_counter() - t0)
    st = np.zeros(4)
    for i in range(16):
        st[:3] = pos[i]
        g.optimized_gain(lt, st)
    for i in range(calls):   # nbvgpu_optimized_gain: the drop-in call (gain + yaw only)
        st[:3] = pos[i % 64]
        t0 = time.perf_counter()
        g.optimized_gain(lt, st)
        to.append(time.perf_counter() - t0)
    g.set_profiling(True)
    g.profile_collect()
    for i in range(calls):   # the same with CUDA events around the sweep (they cost a few us themselves)
        st[:3] = pos[i % 64]
        g.optimized_gain(lt, st)
    kms, kn = g.profile_collect()
    g.set_profiling(False)
    for i in range(calls):
        s = seg[i % 64:i % 64 + 1]
        t0 = time.perf_counter()
        g.check_segments(le, cfg.robot_radius, s)
        ts.append(time.perf_counter() - t0)
    for i in range(calls):
        s = seg[i % 64]
        t0 = time.perf_counter()
        g.check_motion(le, cfg.robot_radius, s[:3], s[3:])
        tk.append(time.perf_counter() - t0)
    med = lambda x: float(np.median(x) * 1e6)
    out = {"geometry": name, "env": {k: v for k, v in os.environ.items() if k.startswith("NBVGPU_")},
           "optimized_gain_us": med(to), "optimized_gain_p10_us": float(np.percentile(to, 10) * 1e6),
           "check_motion_us": med(tk), "gain_eval_n1_with_counts_and_steps_us": med(tg), "check_segments_n1_us": med(ts),
           "ray_sweep_kernel_us": kms / max(1, kn) * 1e3, "fused_launches": g.fused_launches(), "calls": calls}
    g.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--yc", default="0")
    ap.add_argument("--env", default="", help="comma-separated NBVGPU_* switches to set to 1, e.g. NBVGPU_NO_FUSE")
    ap.add_argument("--calls", type=int, default=200)
    ap.add_argument("--cpp", action="store_true", help="also time the C++ drop-in loop (tests/cpp/shim_iterate_loop.cc): no Python in the loop")
    a = ap.parse_args()
    from nbv_exploration_planner_b200 import synth
    for e in filter(None, a.env.split(",")):
        os.environ[e] = "1"
    for yc in [int(x) for x in a.yc.split(",")]:
        if yc:
            os.environ["NBVGPU_YC"] = str(yc)
        else:
            os.environ.pop("NBVGPU_YC", None)
        print(json.dumps(probe("cfg2", "cfg2", a.calls)))
        print(json.dumps(probe("shipped", shipped_cfg(synth, 64), a.calls)))
        if a.cpp:
            import tempfile
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
            from cpp_dropin_fixture import build_loop_exe, run_case
            with tempfile.TemporaryDirectory() as tmp:
                exe = build_loop_exe(tmp)
                for name, cfg in (("cfg2", "cfg2"), ("shipped", shipped_cfg(synth, 64))):
                    m = synth.make_map(cfg)
                    c = synth.make_candidates(m.cfg, 64, None)
                    _, lat = run_case(exe, tmp, m, m.camera(), c, 4)
                    print(json.dumps({"geometry": name, "caller": "C++ (include/nbvgpu_shim.hpp, tests/cpp/shim_iterate_loop.cc)",
                                      "env": {k: v for k, v in os.environ.items() if k.startswith("NBVGPU_")}, **lat}))


if __name__ == "__main__":
    main()
