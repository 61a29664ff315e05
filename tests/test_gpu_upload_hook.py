"""The drop-in's upload hook against the reference's own Voxblox classes: oracle/_ref/shim_upload_hook is
tests/cpp/shim_upload_hook.cc compiled with -DNBVGPU_SHIM_WITH_VOXBLOX against /root/reference's voxblox/core/{layer,block}.h
and src/core/block.cc (oracle/Makefile, target `ref`; prebuilt where the reference is present, it travels to the GPU box).
GpuMap::uploadUpdatedBlocks must send exactly the blocks whose Update::kMap bit is set (Layer::getAllUpdatedBlocks,
layer.h:194-203) in the reference's serialisation (Block::serializeToIntegers, block.cc:159-234), clear the bit, and leave
the mirror answering like the oracle on a layer updated the same way."""
import os
import subprocess

import numpy as np
import pytest

from conftest import load_oracle
from cpp_dropin_fixture import write_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "shim_upload_hook")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/shim_upload_hook not built (reference absent at build time)")
def test_upload_hook_with_reference_layer_classes(oracle_mod, tiny_map, tmp_path):
    from nbv_exploration_planner_b200 import synth
    m = tiny_map
    cam = m.camera()
    c = synth.make_candidates(m.cfg, 80, None)
    inp, outp = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    write_case(inp, m, cam, c, 1)
    r = subprocess.run([EXE, inp, outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(outp, np.float64)
    n = len(c.positions)
    phase = raw[:6 * n].reshape(2, n, 3)
    marked1, up1, marked2, up2 = raw[6 * n:]
    assert marked1 == 0 and marked2 == 0                   # Update::kMap cleared on every uploaded block
    assert up1 == 2 * len(m.keys) and up2 == 3             # everything once, then only the three changed blocks
    o, lt_o, le_o = load_oracle(oracle_mod, m, cam)
    fo = o.check_segments(le_o, m.cfg.robot_radius, c.segments)["free"]
    for ph in range(2):
        if ph == 1:
            t2 = m.tsdf[:3].copy()
            t2[..., 1] = 0.0
            o.layer_update(lt_o, m.keys[:3], t2)
        ro = o.gain_eval(lt_o, c.positions, threads=8)
        assert np.array_equal(phase[ph, :, 2].astype(np.uint8), fo)
        assert np.array_equal(phase[ph, :, 0], np.where(fo == 1, ro["gain"], 0.0))
        assert np.array_equal(phase[ph, :, 1], np.where(fo == 1, ro["yaw_deg"], 0))
    assert not np.array_equal(phase[0, :, 0], phase[1, :, 0])   # the wiped blocks are visible in the gains
