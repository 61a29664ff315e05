import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- shared fixtures ---------------------------------------------------------------------------
@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import binding
    binding.build()
    return binding


def load_oracle(binding, m, cam):
    """Oracle context holding the TSDF (+ESDF) layers of SynthMap `m` and camera `cam`."""
    o = binding.Oracle()
    lt = o.layer_create(0, m.voxel_size, m.vps)
    o.layer_update(lt, m.keys, m.tsdf)
    le = None
    if m.esdf is not None:
        le = o.layer_create(1, m.voxel_size, m.vps)
        o.layer_update(le, m.keys, m.esdf)
    o.set_camera(cam.as_dict())
    return o, lt, le


def load_gpu(m, cam, variant=0, max_blocks=None):
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu
    g = NbvGpu(0)
    nb = max_blocks or max(16, len(m.keys) + 8)
    lt = g.layer_create(LAYER_TSDF, m.voxel_size, m.vps, nb)
    g.layer_update(lt, m.keys, m.tsdf)
    le = None
    if m.esdf is not None:
        le = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, nb)
        g.layer_update(le, m.keys, m.esdf)
    g.commit()
    g.set_camera(cam)
    g.set_kernel_variant(variant)
    return g, lt, le


@pytest.fixture(scope="session")
def tiny_map():
    from nbv_exploration_planner_b200 import synth
    return synth.make_map("tiny")


@pytest.fixture(scope="session")
def cfg1_map():
    from nbv_exploration_planner_b200 import synth
    return synth.make_map("cfg1")
