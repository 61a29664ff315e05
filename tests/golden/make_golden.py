"""Generates the committed golden vectors under tests/golden/.

The reference has no build that works here (catkin / ROS) and holds no fixtures for this path (SURVEY.md section
8c).  The vectors are outputs of the CPU oracle (oracle/nbv_oracle.cc) on seeded synthetic inputs -- the oracle
itself is pinned bit for bit against the reference's own hot-path sources compiled into oracle/_ref
(tests/test_ref_pinning.py) -- and, where the reference's own code returns the quantity, outputs of that code
(`frontier_golden.npz`: `reference_gain` comes from the reference's History::bfs; `interp_golden.npz`: `reference_*`
come from the reference's interpolator, compiled against the Eigen stand-in -- see its summation-order caveat).  They guard the oracle against
silent regressions and give the GPU tests fixtures that need neither the oracle nor /root/reference at run time.
Re-run:  python tests/golden/make_golden.py [frontier | interp]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from nbv_exploration_planner_b200 import synth  # noqa: E402
from oracle.binding import Oracle, cast_ray  # noqa: E402


def golden_cfg(name, n, out):
    m = synth.make_map(name)
    cfg = m.cfg
    o = Oracle()
    lt = o.layer_create(0, m.voxel_size, m.vps)
    le = o.layer_create(1, m.voxel_size, m.vps)
    o.layer_update(lt, m.keys, m.tsdf)
    o.layer_update(le, m.keys, m.esdf)
    o.set_camera(m.camera().as_dict())
    c = synth.make_candidates(cfg, n, lambda s: o.check_segments(le, cfg.robot_radius, s)["free"])
    r = o.gain_eval(lt, c.positions, per_yaw=True, threads=os.cpu_count())
    # unfiltered edges: a mix of free and colliding segments
    cu = synth.make_candidates(cfg, 4 * n, None, seed=cfg.seed_cand + 500)
    fr = o.check_segments(le, cfg.robot_radius, cu.segments)
    np.savez_compressed(out, map_sha256=m.sha256(), positions=c.positions, per_yaw=r["per_yaw"], yaw_deg=r["yaw_deg"],
                        gain=r["gain"], steps=r["steps"], counts=r["counts"], segments=cu.segments, seg_free=fr["free"],
                        seg_steps=fr["steps"], cand_segments=c.segments, cand_distance=c.distance)
    print(name, "map sha", m.sha256()[:16], "steps", int(r["steps"].sum()), "free edges", int(fr["free"].sum()), "/", len(fr["free"]))


def golden_rays(out):
    rng = np.random.default_rng(42)
    starts = rng.uniform(-30, 30, (64, 3)).astype(np.float32)
    ends = (starts + rng.uniform(-12, 12, (64, 3))).astype(np.float32)
    # degenerate components
    ends[0:8, 0] = starts[0:8, 0]
    ends[8:16, 1] = starts[8:16, 1]
    ends[16:24, 2] = starts[16:24, 2]
    ends[24:28] = starts[24:28]
    starts[28:32] = np.round(starts[28:32])
    seqs = [cast_ray(s, e) for s, e in zip(starts, ends)]
    np.savez_compressed(out, starts=starts, ends=ends, lengths=np.array([len(s) for s in seqs]),
                        indices=np.concatenate(seqs, 0))


def golden_voxblox_file(out):
    """A 3-block TSDF + ESDF .voxblox file encoded by the protobuf library (tests/voxblox_fixture.py)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import voxblox_fixture as vf
    m = synth.make_map("tiny")
    data = vf.voxblox_file_bytes([("tsdf", m.voxel_size, m.vps, m.keys[:3], vf.tsdf_words(m.tsdf[:3])),
                                  ("esdf", m.voxel_size, m.vps, m.keys[:3], vf.esdf_words(m.esdf[:3]))])
    open(out, "wb").write(data)
    print("voxblox fixture", len(data), "bytes")


def golden_frontier(out):
    """History::bfs (history.cpp:96-137) on the tiny map: vertices in free / unknown / occupied space, on and outside
    the bounds; oracle outputs plus the gains of the reference's own bfs (oracle/_ref) when it is built."""
    from oracle import ref_binding as rb
    m = synth.make_map("tiny")
    cfg, cam = m.cfg, m.camera()
    o = Oracle()
    lt = o.layer_create(0, m.voxel_size, m.vps)
    o.layer_update(lt, m.keys, m.tsdf)
    o.set_camera(cam.as_dict())
    lo, hi = np.array(cfg.bounds_min, float), np.array(cfg.bounds_max, float)
    special = np.array([lo + 0.01, hi - 0.01, lo - 0.5, [lo[0], 0.5 * (lo[1] + hi[1]), 1.0], 0.5 * (lo + hi),
                        [cfg.root[0], cfg.root[1], cfg.root[2]], [1.0, 2.0, 0.5], [3.0, 4.0, 1.0], hi, lo, [-0.0, 4.0, 1.0]])
    pts = np.concatenate([synth.make_candidates(cfg, 12, None).positions, special])
    r = o.frontier_potential(lt, pts, 6.0, threads=os.cpu_count())
    ref_gain = np.zeros(0)
    if rb.available():
        ref = rb.Reference(m.voxel_size, m.vps)
        ref.tsdf_update(m.keys, m.tsdf)
        ref.set_camera(cam.as_dict(), 0.4)
        ref_gain = ref.frontier_potential(pts)
        assert np.array_equal(ref_gain, r["gain"]), "oracle and reference disagree: do not write a golden file"
    np.savez_compressed(out, map_sha256=m.sha256(), points=pts, max_distance=6.0, gain=r["gain"], frontier_voxels=r["frontier_voxels"],
                        visited=r["visited"], cells_checksum=r["cells_checksum"], reference_gain=ref_gain)
    print("frontier golden:", len(pts), "vertices, frontier voxels", r["frontier_voxels"].tolist(), "reference gains recorded:", len(ref_gain) > 0)


def golden_interp(out):
    """EsdfMap::getDistanceAndGradientAtPosition (esdf_map.cc:22-52) on the tiny map with observed = "the TSDF saw this
    voxel": random queries, voxel centres, faces; oracle outputs plus the reference interpolator's own (oracle/_ref)."""
    from oracle import ref_binding as rb
    m = synth.make_map("tiny")
    obs = (m.tsdf[..., 1] > 0).astype(np.uint8)
    o = Oracle()
    le = o.layer_create(1, m.voxel_size, m.vps)
    o.layer_update_esdf_observed(le, m.keys, m.esdf, obs)
    rng = np.random.default_rng(31)
    vs = float(np.float32(m.voxel_size))
    lo, hi = np.array(m.cfg.bounds_min, float), np.array(m.cfg.bounds_max, float)
    q = np.concatenate([rng.uniform(lo - 0.3, hi + 0.3, (600, 3)), (rng.integers(0, 40, (150, 3)) + 0.5) * vs, rng.integers(0, 40, (150, 3)) * vs,
                        [[0.0, 0.0, 0.0], lo, hi, [100.0, 0.0, 0.0]]])
    r = o.distance_and_gradient(le, q, threads=os.cpu_count())
    ref_ok, ref_d, ref_g = np.zeros(0, np.uint8), np.zeros(0), np.zeros((0, 3))
    if rb.available():
        ref = rb.Reference(m.voxel_size, m.vps)
        ref.esdf_update_observed(m.keys, m.esdf, obs)
        rr = ref.distance_and_gradient(q)
        assert np.array_equal(rr["ok"], r["ok"]) and np.array_equal(rr["distance"], r["distance"]) and np.array_equal(rr["gradient"], r["gradient"]), \
            "oracle and reference disagree: do not write a golden file"
        ref_ok, ref_d, ref_g = rr["ok"], rr["distance"], rr["gradient"]
    np.savez_compressed(out, map_sha256=m.sha256(), positions=q, ok=r["ok"], distance=r["distance"], gradient=r["gradient"],
                        reference_ok=ref_ok, reference_distance=ref_d, reference_gradient=ref_g)
    print("interp golden:", len(q), "queries,", int(r["ok"].sum()), "succeed; reference outputs recorded:", len(ref_ok) > 0)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "frontier":
        golden_frontier(os.path.join(HERE, "frontier_golden.npz"))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "interp":
        golden_interp(os.path.join(HERE, "interp_golden.npz"))
        sys.exit(0)
    golden_cfg("cfg1", 100, os.path.join(HERE, "cfg1_golden.npz"))
    golden_cfg("tiny", 16, os.path.join(HERE, "tiny_golden.npz"))
    golden_rays(os.path.join(HERE, "rays_golden.npz"))
    golden_voxblox_file(os.path.join(HERE, "tiny_tsdf_esdf.voxblox"))
    golden_frontier(os.path.join(HERE, "frontier_golden.npz"))
    golden_interp(os.path.join(HERE, "interp_golden.npz"))
