"""Measures the history-graph frontier potential (History::bfs) on one GPU: vertices/s through
the host-buffer C ABI (H2D of the vertex positions, kernel, D2H of the results inside the timed region), with the
reference's own History::bfs (oracle/_ref, one thread: the reference runs it on its planner thread) and the
oracle port timed on a bounded sample of the same vertices.  Prints one JSON line.
usage: python scripts/bench_frontier.py [--config cfg1] [--vertices 148] [--steps 5] [--cpu-sample 3]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg1")
    ap.add_argument("--vertices", type=int, default=148)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=3)
    ap.add_argument("--max-distance", type=float, default=6.0)
    args = ap.parse_args()
    import torch
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu, synth
    m = synth.make_map(args.config)
    cam = m.camera()
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, len(m.keys) + 8)
    g.layer_update(lt, m.keys, m.tsdf)
    g.commit()
    g.set_camera(cam)
    pts = synth.make_candidates(m.cfg, args.vertices, None).positions
    for _ in range(args.warmup):
        r = g.frontier_potential(lt, pts, args.max_distance)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = g.frontier_potential(lt, pts, args.max_distance)
    dt = (time.perf_counter() - t0) / args.steps
    one = []
    for i in range(min(8, len(pts))):
        t1 = time.perf_counter()
        g.frontier_potential(lt, pts[i:i + 1], args.max_distance)
        one.append(time.perf_counter() - t1)
    out = {"metric": "history-graph vertices/s (frontier potential, History::bfs)", "value": len(pts) / dt, "unit": "vertices/s",
           "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "timing": "host wall clock around the C-ABI call (host buffers in, host buffers out)",
           "config": {"workload": f"{args.config} map, {len(pts)} vertices per call, max_distance {args.max_distance} m, "
                                  f"voxel {m.voxel_size} m", "map_blocks": int(len(m.keys))},
           "visited_cells_per_vertex": float(r["visited"].mean()), "cells_per_s": float(r["visited"].sum()) / dt,
           "single_vertex_latency_ms": float(np.median(one) * 1e3)}
    # CPU side on a bounded sample: the reference's own code (1 thread) and the oracle port
    k = min(args.cpu_sample, len(pts))
    if k > 0:
        from oracle import binding as ob, ref_binding as rb
        o = ob.Oracle()
        lo = o.layer_create(0, m.voxel_size, m.vps)
        o.layer_update(lo, m.keys, m.tsdf)
        o.set_camera(cam.as_dict())
        t1 = time.perf_counter()
        ro = o.frontier_potential(lo, pts[:k], args.max_distance, threads=1)
        t_or = (time.perf_counter() - t1) / k
        ok = bool(np.array_equal(ro["gain"], r["gain"][:k]) and np.array_equal(ro["cells_checksum"], r["cells_checksum"][:k]))
        cpu = {"oracle_port_ms_per_vertex": t_or * 1e3, "sample": f"first {k} vertices, 1 thread", "gpu_equals_oracle_on_sample": ok}
        if rb.available() and args.max_distance == 6.0:
            ref = rb.Reference(m.voxel_size, m.vps)
            ref.tsdf_update(m.keys, m.tsdf)
            ref.set_camera(cam.as_dict(), 0.4)
            t1 = time.perf_counter()
            rr = ref.frontier_potential(pts[:k])
            t_ref = (time.perf_counter() - t1) / k
            cpu.update({"kind": "reference", "reference_ms_per_vertex": t_ref * 1e3, "value": 1.0 / t_ref, "unit": "vertices/s", "cores": 1,
                        "gpu_equals_reference_on_sample": bool(np.array_equal(rr, r["gain"][:k]))})
        else:
            cpu.update({"kind": "port", "value": 1.0 / t_or, "unit": "vertices/s", "cores": 1})
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    g.close()


if __name__ == "__main__":
    main()
