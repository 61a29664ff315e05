"""Generates the committed golden vectors under tests/golden/ from the CPU oracle.

The reference itself cannot be built or imported here, so the vectors are
outputs of the oracle restatement (oracle/nbv_oracle.cc) on seeded synthetic inputs; they pin the
oracle against silent regressions and give the GPU tests a fixture that does not need the oracle
at run time.  Re-run:  python tests/golden/make_golden.py
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


if __name__ == "__main__":
    golden_cfg("cfg1", 100, os.path.join(HERE, "cfg1_golden.npz"))
    golden_cfg("tiny", 16, os.path.join(HERE, "tiny_golden.npz"))
    golden_rays(os.path.join(HERE, "rays_golden.npz"))
    golden_voxblox_file(os.path.join(HERE, "tiny_tsdf_esdf.voxblox"))
