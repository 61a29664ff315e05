"""Randomised GPU-vs-oracle sweep for the widened rows: the history-graph frontier potential (History::bfs) and the ESDF
interpolated distance / gradient, on random worlds, voxel sizes, exploration boxes, radii and start points (free, unknown,
occupied, outside the box, on voxel faces, with -0.0 coordinates).  Frontier: count, gain, visited-set size and the per-cell
coordinate checksum identical; interpolation: flag, distance, gradient identical.  Test infrastructure (it uses the oracle as the checker); not collected by pytest:
python tests/fuzz/fuzz_widened.py [--cases 30] [--seed 1]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=30)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, PACK_VOXBLOX, NbvGpu, synth
    from nbv_exploration_planner_b200.planner import CameraParams
    from oracle.binding import Oracle
    from voxblox_fixture import esdf_words
    rng = np.random.default_rng(args.seed)
    bad = 0
    redo = 0
    t0 = time.time()
    for case in range(args.cases):
        vs = float(rng.choice([0.1, 0.15, 0.2, 0.25, 0.3, 0.4, 0.5]))
        world = str(rng.choice(["pillars", "office"]))
        sx, sy, sz = rng.uniform(6, 12), rng.uniform(6, 12), rng.uniform(2.5, 4)
        lo = (float(rng.uniform(-3, 1)), float(rng.uniform(-3, 1)), float(rng.uniform(-0.5, 0.2)))
        hi = (lo[0] + float(sx), lo[1] + float(sy), lo[2] + float(sz))
        root = (0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), lo[2] + 1.2)
        cfg = synth.SynthConfig(f"fz{case}", lo, hi, vs, 3.0, 6, world, 9000 + case, 9500 + case, root, obs_radius=float(rng.uniform(3, 5)),
                                traj_points=int(rng.integers(3, 7)), traj_step=float(rng.uniform(1.5, 3)), vicinity=float(rng.uniform(2, 4)))
        m = synth.make_map(cfg)
        shrink = float(rng.uniform(-0.5, 1.0))
        cam = CameraParams.mount(10.0, (0.1, 0.0, 0.0), hfov_deg=90, vfov_deg=60, near_m=0.1, far_m=3.0,
                                 bbx_min=tuple(v + shrink for v in lo), bbx_max=tuple(v - shrink for v in hi))
        obs = (m.tsdf[..., 1] > 0).astype(np.uint8)
        o = Oracle()
        lt_o, le_o = o.layer_create(0, m.voxel_size, m.vps), o.layer_create(1, m.voxel_size, m.vps)
        o.layer_update(lt_o, m.keys, m.tsdf)
        o.layer_update_esdf_observed(le_o, m.keys, m.esdf, obs)
        o.set_camera(cam.as_dict())
        g = NbvGpu(0)
        lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
        le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 8)
        g.layer_update(lt, m.keys, m.tsdf)
        g.layer_update(le, m.keys, esdf_words(m.esdf, flags=obs.astype(np.uint32)), PACK_VOXBLOX)
        g.commit()
        g.set_camera(cam)
        msg = []
        # ---- frontier potential: a few vertices, radius small enough for the oracle's string-keyed fill to stay quick
        radius = float(rng.choice([1.0, 2.0, 3.0, 6.0 if vs >= 0.25 else 2.5]))
        c = synth.make_candidates(cfg, 4, None).positions
        vsd = float(np.float32(vs))
        pts = np.concatenate([c, [np.array(lo) + shrink + 0.01, [root[0], root[1], root[2]], np.round(np.array(root) / vsd) * vsd,
                                  [-0.0, root[1], root[2]], np.array(lo) - 1.0]])
        ro, rg = o.frontier_potential(lt_o, pts, radius, threads=os.cpu_count()), g.frontier_potential(lt, pts, radius)
        for k in ("frontier_voxels", "gain", "visited", "cells_checksum"):
            if not np.array_equal(rg[k], ro[k]):
                msg.append(f"frontier {k} differs")
        redo += g.frontier_hash_redo()
        # ---- interpolation
        q = np.concatenate([rng.uniform(np.array(lo) - 0.5, np.array(hi) + 0.5, (4000, 3)), (rng.integers(-10, 60, (1000, 3)) + 0.5) * vsd,
                            rng.integers(-10, 60, (1000, 3)) * vsd])
        io, ig = o.distance_and_gradient(le_o, q, threads=os.cpu_count()), g.esdf_distance_gradient(le, q)
        for k in ("ok", "distance", "gradient"):
            if not np.array_equal(ig[k], io[k]):
                msg.append(f"interp {k} differs")
        g.close()
        bad += bool(msg)
        print(f"{'OK ' if not msg else 'BAD'} case {case}: vs {vs} {world} radius {radius} shrink {shrink:.2f} visited {ro['visited'].tolist()} "
              f"interp ok {int(io['ok'].sum())}/{len(q)} {'; '.join(msg)}", flush=True)
    print(f"{args.cases - bad}/{args.cases} cases identical, {redo} vertices took the hash-set redo, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
