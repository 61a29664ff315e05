"""Helpers around tests/cpp/shim_iterate_loop.cc (the C++ drop-in loop): build it, write its input file, run it.
Used by tests/test_gpu_cpp_dropin.py and scripts/latency_probe.py --cpp."""
import json
import os
import struct
import subprocess

import numpy as np

from voxblox_fixture import esdf_words, tsdf_words

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_loop_exe(out_dir):
    from nbv_exploration_planner_b200 import _native as N
    N.build()
    exe = os.path.join(str(out_dir), "shim_iterate_loop")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_iterate_loop.cc"), "-o", exe,
                           "-L", N.CSRC, "-lnbvgpu", f"-Wl,-rpath,{N.CSRC}"])
    return exe


def write_case(path, m, cam, c, reps):
    d = cam.as_dict()
    camv = [d["hfov_deg"], d["vfov_deg"], d["near_m"], d["far_m"], *d["q_CB"], *d["t_CB"], *d["bbx_min"], *d["bbx_max"],
            d["gain_free"], d["gain_occupied"], d["gain_unmapped"], float(m.cfg.robot_radius)]
    assert len(camv) == 21
    samples = np.concatenate([c.positions, c.segments], axis=1).astype(np.float64)
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", m.vps, len(m.keys), len(samples), reps))
        f.write(struct.pack("<f", float(m.voxel_size)))
        f.write(np.asarray(camv, np.float64).tobytes())
        f.write(np.ascontiguousarray(m.keys, np.int32).tobytes())
        f.write(tsdf_words(m.tsdf).tobytes())
        f.write(esdf_words(m.esdf).tobytes())
        f.write(samples.tobytes())


def run_case(exe, tmp, m, cam, c, reps):
    inp, outp = os.path.join(str(tmp), "case.bin"), os.path.join(str(tmp), "out.bin")
    write_case(inp, m, cam, c, reps)
    r = subprocess.run([exe, inp, outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = np.fromfile(outp, np.float64).reshape(-1, 3)
    return out, json.loads(r.stdout.strip().splitlines()[-1])
