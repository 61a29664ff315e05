"""Randomised GPU-vs-oracle parity sweep for the gain / collision path: random cameras (FOVs, pitch, mount offset,
near / far range), voxel sizes, worlds and exploration boxes; per-yaw counts, steps, yaw and gain must be identical
for the default kernel and the cross-checking variant (whose mismatch counter must stay 0), edge verdicts identical.
Test infrastructure (it uses the oracle as the checker); not collected by pytest (the oracle side takes a while):  python tests/fuzz/fuzz_parity.py [--cases 40] [--seed 1]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--candidates", type=int, default=6)
    args = ap.parse_args()
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, synth
    from nbv_exploration_planner_b200.planner import CameraParams
    from oracle.binding import Oracle
    rng = np.random.default_rng(args.seed)
    bad = 0
    t0 = time.time()
    for case in range(args.cases):
        vs = float(rng.choice([0.1, 0.15, 0.2, 0.25, 0.3, 0.4]))
        far = float(rng.uniform(1.5, 5.0))
        hfov = float(rng.choice([58.0, 69.0, 90.0, 110.0, 120.0]))
        vfov = float(rng.choice([42.0, 45.0, 60.0, 75.0]))
        pitch = float(rng.uniform(-25, 30))
        near = float(rng.choice([0.05, 0.1, 0.3, 0.5]))
        world = str(rng.choice(["pillars", "office"]))
        sx, sy, sz = rng.uniform(6, 14), rng.uniform(6, 14), rng.uniform(2.5, 5)
        lo = (float(rng.uniform(-3, 1)), float(rng.uniform(-3, 1)), float(rng.uniform(-0.5, 0.2)))
        hi = (lo[0] + float(sx), lo[1] + float(sy), lo[2] + float(sz))
        root = (0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), lo[2] + 1.2)
        cfg = synth.SynthConfig(f"fuzz{case}", lo, hi, vs, far, args.candidates, world, 7000 + case, 8000 + case, root,
                                obs_radius=float(rng.uniform(3, 6)), traj_points=int(rng.integers(3, 8)), traj_step=float(rng.uniform(1.5, 3)),
                                vicinity=float(rng.uniform(2, 5)), hfov_deg=hfov, vfov_deg=vfov, pitch_deg=pitch, near_m=near)
        m = synth.make_map(cfg)
        t_BC = tuple(float(v) for v in rng.uniform(-0.15, 0.15, 3)) if rng.random() < 0.7 else (0.0, 0.0, 0.0)
        # the exploration box is sometimes tighter than the map, sometimes looser
        shrink = float(rng.uniform(-1.0, 1.5))
        cam = CameraParams.mount(pitch, t_BC, hfov_deg=hfov, vfov_deg=vfov, near_m=near, far_m=far,
                                 bbx_min=tuple(v + shrink for v in lo), bbx_max=tuple(v - shrink for v in hi))
        o = Oracle()
        lo_t, lo_e = o.layer_create(0, m.voxel_size, m.vps), o.layer_create(1, m.voxel_size, m.vps)
        o.layer_update(lo_t, m.keys, m.tsdf)
        o.layer_update(lo_e, m.keys, m.esdf)
        o.set_camera(cam.as_dict())
        c = synth.make_candidates(cfg, args.candidates, None)
        pos = np.concatenate([c.positions, [[lo[0] - 0.3, root[1], root[2]], [root[0], root[1], hi[2] + 0.2]]])   # two outside the map
        ro = o.gain_eval(lo_t, pos, per_yaw=True, threads=os.cpu_count())
        fo = o.check_segments(lo_e, cfg.robot_radius, c.segments)
        msg = []
        for variant in (0, 2, 4):
            g = NbvGpu(0)
            lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
            le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 8)
            g.layer_update(lt, m.keys, m.tsdf)
            g.layer_update(le, m.keys, m.esdf)
            g.commit()
            g.set_camera(cam)
            g.set_kernel_variant(variant)
            rg = g.gain_eval(lt, pos, per_yaw=True, steps=True)
            fg = g.check_segments(le, cfg.robot_radius, c.segments)
            for k in ("per_yaw", "steps", "yaw_deg", "gain"):
                if not np.array_equal(rg[k], ro[k]):
                    msg.append(f"variant {variant}: {k} differs")
            if not np.array_equal(fg, fo["free"]):
                msg.append(f"variant {variant}: edge verdicts differ")
            if variant in (0, 2):
                # the single-item paths (mapped in-place I/O, fused window scan, completion ticket, one warp per segment)
                rs = g.gain_eval(lt, pos, steps=False)
                if not (np.array_equal(rs["gain"], ro["gain"]) and np.array_equal(rs["yaw_deg"], ro["yaw_deg"]) and np.array_equal(rs["counts"], ro["counts"])):
                    msg.append(f"variant {variant}: small-batch gain differs")
                for i in range(len(pos)):
                    st = np.array([*pos[i], 0.0])
                    if g.optimized_gain(lt, st) != ro["gain"][i] or st[3] != ro["yaw_deg"][i] * np.pi / 180.0:
                        msg.append(f"variant {variant}: optimized_gain({i}) differs")
                for i in range(len(c.segments)):
                    if g.check_motion(le, cfg.robot_radius, c.segments[i, :3], c.segments[i, 3:]) != bool(fo["free"][i]):
                        msg.append(f"variant {variant}: check_motion({i}) differs")
            if variant == 0:
                # one call per planning step against the two calls and the oracle (no gain behind a colliding edge: rrt.cpp:210-216)
                nc = len(c.positions)
                py = np.linspace(-3.0, 3.0, nc)
                two = g.gain_eval_best(lt, c.positions, py, c.distance, fg, 1.0, 0.5, 0.0)
                one = g.evaluate_candidates(lt, le, cfg.robot_radius, c.positions, c.segments, py, c.distance, 1.0, 0.5, 0.0)
                fr = fo["free"] == 1
                if not (np.array_equal(one["free"], fo["free"]) and np.array_equal(one["gain"], np.where(fr, ro["gain"][:nc], 0.0)) and
                        np.array_equal(one["yaw_deg"], np.where(fr, ro["yaw_deg"][:nc], 0)) and
                        all(one[k] == two[k] for k in ("index", "score", "best_gain", "best_yaw"))):
                    msg.append("evaluate_candidates differs from the two calls / the oracle")
            if variant == 2 and g.classifier_mismatches() != 0:
                msg.append(f"variant 2: {g.classifier_mismatches()} cross-check mismatches")
            g.close()
        tag = "OK " if not msg else "BAD"
        bad += bool(msg)
        print(f"{tag} case {case}: vs {vs} far {far:.2f} fov {hfov}x{vfov} pitch {pitch:.1f} near {near} {world} t_BC {t_BC} shrink {shrink:.2f} "
              f"steps {int(ro['steps'].sum())} {'; '.join(msg)}", flush=True)
    print(f"{args.cases - bad}/{args.cases} cases identical, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
