"""Measures the batched ESDF interpolated distance / gradient (EsdfMap::getDistanceAndGradientAtPosition) on
one GPU: queries/s through the host-buffer C ABI (H2D of the positions, kernel, D2H of flag / distance / gradient inside the
timed region), with the reference's own interpolator (oracle/_ref, one thread: History::refineVertexPosition runs on the planner
thread) and the oracle port timed on a bounded sample of the same queries.  Prints one JSON line.
usage: python scripts/bench_interp.py [--config cfg1] [--queries 1000000] [--steps 10] [--cpu-sample 200000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg1")
    ap.add_argument("--queries", type=int, default=1000000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=200000)
    args = ap.parse_args()
    import torch
    from nbv_exploration_planner_b200 import LAYER_ESDF, PACK_VOXBLOX, NbvGpu, synth
    from voxblox_fixture import esdf_words
    m = synth.make_map(args.config)
    obs = (m.tsdf[..., 1] > 0).astype(np.uint8)
    g = NbvGpu(0)
    le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 8)
    g.layer_update(le, m.keys, esdf_words(m.esdf, flags=obs.astype(np.uint32)), PACK_VOXBLOX)
    g.commit()
    rng = np.random.default_rng(5)
    lo, hi = np.array(m.cfg.bounds_min, float), np.array(m.cfg.bounds_max, float)
    q = rng.uniform(lo, hi, (args.queries, 3))
    for _ in range(args.warmup):
        r = g.esdf_distance_gradient(le, q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = g.esdf_distance_gradient(le, q)
    dt = (time.perf_counter() - t0) / args.steps
    t1 = time.perf_counter()
    for i in range(200):
        g.esdf_distance_gradient(le, q[i:i + 1])
    one = (time.perf_counter() - t1) / 200
    # inputs and outputs resident in HBM: nbvgpu_esdf_distance_gradient_device on an explicit stream, CUDA events
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        d_q = torch.from_numpy(q).cuda()
        d_ok = torch.empty(len(q), dtype=torch.uint8, device="cuda")
        d_dist = torch.empty(len(q), dtype=torch.float64, device="cuda")
        d_grad = torch.empty((len(q), 3), dtype=torch.float64, device="cuda")
        for _ in range(args.warmup):
            g.esdf_distance_gradient_device(le, len(q), d_q, d_ok, d_dist, d_grad)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            g.esdf_distance_gradient_device(le, len(q), d_q, d_ok, d_dist, d_grad)
        e1.record()
        stream.synchronize()
        dev_ms = e0.elapsed_time(e1) / args.steps
        same_dev = bool(np.array_equal(d_ok.cpu().numpy(), r["ok"]) and np.array_equal(d_dist.cpu().numpy(), r["distance"]))
    g.set_stream(None)
    out = {"metric": "ESDF interpolated distance+gradient queries/s (EsdfMap::getDistanceAndGradientAtPosition)", "value": len(q) / dt,
           "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "timing": "host wall clock around the C-ABI call (host buffers in, host buffers out)",
           "config": {"workload": f"{args.config} map, {len(q)} uniform random positions inside the map bounds per call, voxel {m.voxel_size} m",
                      "map_blocks": int(len(m.keys))},
           "success_fraction": float(r["ok"].mean()), "single_query_latency_us": one * 1e6,
           "device_resident": {"value": len(q) / (dev_ms * 1e-3), "unit": "queries/s", "ms_per_step": dev_ms,
                               "timing": "CUDA events on the launching stream, buffers in HBM (nbvgpu_esdf_distance_gradient_device)",
                               "equals_host_path": same_dev},
           "h2d_bytes_per_step": len(q) * 24, "d2h_bytes_per_step": len(q) * 33}
    k = min(args.cpu_sample, len(q))
    if k > 0:
        from oracle import binding as ob, ref_binding as rb
        o = ob.Oracle()
        lo_e = o.layer_create(1, m.voxel_size, m.vps)
        o.layer_update_esdf_observed(lo_e, m.keys, m.esdf, obs)
        t1 = time.perf_counter()
        ro = o.distance_and_gradient(lo_e, q[:k], threads=1)
        t_or = time.perf_counter() - t1
        same = bool(np.array_equal(ro["ok"], r["ok"][:k]) and np.array_equal(ro["distance"], r["distance"][:k]) and
                    np.array_equal(ro["gradient"], r["gradient"][:k]))
        cpu = {"oracle_port_queries_per_s": k / t_or, "sample": f"first {k} queries, 1 thread", "gpu_equals_oracle_on_sample": same}
        if rb.available():
            ref = rb.Reference(m.voxel_size, m.vps)
            ref.esdf_update_observed(m.keys, m.esdf, obs)
            t1 = time.perf_counter()
            rr = ref.distance_and_gradient(q[:k])
            t_ref = time.perf_counter() - t1
            cpu.update({"kind": "reference", "value": k / t_ref, "unit": "queries/s", "cores": 1,
                        "note": "the reference's interpolator lines compiled against the Eigen stand-in (scalar products)",
                        "gpu_equals_reference_on_sample": bool(np.array_equal(rr["ok"], r["ok"][:k]) and np.array_equal(rr["distance"], r["distance"][:k])
                                                               and np.array_equal(rr["gradient"], r["gradient"][:k]))})
        else:
            cpu.update({"kind": "port", "value": k / t_or, "unit": "queries/s", "cores": 1})
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    g.close()


if __name__ == "__main__":
    main()
