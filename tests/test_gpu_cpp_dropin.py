"""The drop-in from C++: tests/cpp/shim_iterate_loop.cc drives include/nbvgpu_shim.hpp the way RrtTree::iterate drives the
reference (one checkMotion + one optimized_gain per sample, sequentially) on a map uploaded in the reference's own block
serialisation; gains, yaws and collision flags must equal the oracle's bit for bit.  The per-call latencies it prints are
what a C++ planner sees (no Python in the loop)."""
import numpy as np
import pytest

from conftest import load_oracle
from cpp_dropin_fixture import build_loop_exe, run_case

pytestmark = pytest.mark.gpu


def test_cpp_iterate_loop_equals_oracle(oracle_mod, tiny_map, tmp_path):
    from nbv_exploration_planner_b200 import synth
    m = tiny_map
    cam = m.camera()
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    c = synth.make_candidates(m.cfg, 120, None)
    out, lat = run_case(build_loop_exe(tmp_path), tmp_path, m, cam, c, 2)
    ro = o.gain_eval(lt_o, c.positions, threads=8)
    fo = o.check_segments(le_o, m.cfg.robot_radius, c.segments)["free"]
    assert np.array_equal(out[:, 0], ro["gain"])
    assert np.array_equal(out[:, 1], ro["yaw_deg"] * np.pi / 180.0)
    assert np.array_equal(out[:, 2].astype(np.uint8), fo)
    assert 0 < int(fo.sum()) < len(fo) and len(np.unique(ro["yaw_deg"])) > 3
    print("C++ drop-in latency:", lat)
