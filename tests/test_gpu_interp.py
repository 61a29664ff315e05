"""GPU parity for the ESDF interpolated distance / gradient: nbvgpu_esdf_distance_gradient vs the
oracle's restatement of EsdfMap::getDistanceAndGradientAtPosition (esdf_map.cc:22-52, interpolator_inl.h), which is
pinned against the reference's own interpolator lines in tests/test_ref_pinning.py.  Bar: the success flag exact; the
values identical to the oracle (both sum the two 8-term products left to right) -- against a build of the reference
with real Eigen the values are only claimed within float tolerance, because that summation order belongs to Eigen --
and exactly the voxel's own distance at voxel centres, where seven of the eight weights are exactly zero."""
import numpy as np
import pytest

from conftest import load_oracle
from voxblox_fixture import esdf_words

pytestmark = pytest.mark.gpu


def _observed(m):
    return (m.tsdf[..., 1] > 0).astype(np.uint8)   # what the ESDF integrator marks observed: the voxels the TSDF saw


def _setup(oracle_mod, m, packing="voxblox"):
    from nbv_exploration_planner_b200 import LAYER_ESDF, PACK_VOXBLOX, NbvGpu
    obs = _observed(m)
    o = oracle_mod.Oracle()
    le_o = o.layer_create(1, m.voxel_size, m.vps)
    g = NbvGpu(0)
    le_g = g.layer_create(LAYER_ESDF, m.voxel_size, m.vps, len(m.keys) + 8)
    if packing == "voxblox":        # the reference's own serialisation carries EsdfVoxel::observed in bit 0 of the flags byte
        o.layer_update_esdf_observed(le_o, m.keys, m.esdf, obs)
        g.layer_update(le_g, m.keys, esdf_words(m.esdf, flags=obs.astype(np.uint32)), PACK_VOXBLOX)
    else:                           # planar convenience packing: observed unless the distance is NaN
        esdf = np.where(obs > 0, m.esdf, np.float32("nan")).astype(np.float32)
        o.layer_update(le_o, m.keys, esdf)
        g.layer_update(le_g, m.keys, esdf)
    g.commit()
    return o, le_o, g, le_g


def _queries(m, rng, n=6000):
    vs = float(np.float32(m.voxel_size))
    lo, hi = np.array(m.cfg.bounds_min, float), np.array(m.cfg.bounds_max, float)
    rand = rng.uniform(lo - 0.5, hi + 0.5, (n, 3))
    centres = (rng.integers(0, 40, (n // 4, 3)) + 0.5) * vs
    faces = rng.integers(0, 40, (n // 4, 3)) * vs
    near = centres[: n // 8] + rng.choice([-1e-6, 1e-6, -1e-4, 1e-4], (n // 8, 3))
    return np.concatenate([rand, centres, faces, near, [[0.0, 0.0, 0.0], [-0.0, 1.0, 1.0], lo, hi]])


def _reference_centres(vi, vs, vps):
    vsf = np.float32(vs)
    block_size = np.float32(vps) * vsf                                                           # layer.h:39, float
    origin = (vi // vps).astype(np.float32) * block_size                                          # common.h:196-201
    centre = (((vi % vps).astype(np.float32).astype(np.float64) + 0.5) * np.float64(vsf)).astype(np.float32)   # :188-193
    return (origin + centre).astype(np.float64)                                                   # float add


def _compare(o, le_o, g, le_g, q):
    ro, rg = o.distance_and_gradient(le_o, q, threads=8), g.esdf_distance_gradient(le_g, q)
    assert np.array_equal(rg["ok"], ro["ok"])
    assert np.array_equal(rg["distance"], ro["distance"])
    assert np.array_equal(rg["gradient"], ro["gradient"])    # the partial sums a failing gradient leaves behind included
    return ro, rg


@pytest.mark.parametrize("packing", ["voxblox", "planar"])
def test_tiny_map_random_and_structured_queries(oracle_mod, tiny_map, packing):
    o, le_o, g, le_g = _setup(oracle_mod, tiny_map, packing)
    ro, rg = _compare(o, le_o, g, le_g, _queries(tiny_map, np.random.default_rng(23)))
    assert 0.1 < ro["ok"].mean() < 0.95 and np.abs(ro["gradient"][ro["ok"] == 1]).max() > 0.5
    g.close()


def test_against_committed_golden_vectors(tiny_map):
    """No oracle at run time: tests/golden/interp_golden.npz (oracle outputs == the reference interpolator's own,
    recorded when the file was generated)."""
    import os
    from nbv_exploration_planner_b200 import LAYER_ESDF, PACK_VOXBLOX, NbvGpu
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interp_golden.npz"))
    assert tiny_map.sha256() == str(gold["map_sha256"])
    g = NbvGpu(0)
    le = g.layer_create(LAYER_ESDF, tiny_map.voxel_size, tiny_map.vps, len(tiny_map.keys) + 8)
    g.layer_update(le, tiny_map.keys, esdf_words(tiny_map.esdf, flags=_observed(tiny_map).astype(np.uint32)), PACK_VOXBLOX)
    g.commit()
    r = g.esdf_distance_gradient(le, gold["positions"])
    for k in ("ok", "distance", "gradient"):
        assert np.array_equal(r[k], gold[k]), k
        assert np.array_equal(r[k], gold["reference_" + k]), k
    g.close()


def test_cfg1_map(oracle_mod, cfg1_map):
    o, le_o, g, le_g = _setup(oracle_mod, cfg1_map)
    ro, _ = _compare(o, le_o, g, le_g, _queries(cfg1_map, np.random.default_rng(29), 20000))
    assert ro["ok"].sum() > 2000
    g.close()


def test_linear_field_and_voxel_centres(oracle_mod):
    """d = x + y/2 - z/4 on fully observed blocks: trilinear interpolation reproduces a linear field (to float
    rounding) with gradient (1, 1/2, -1/4); at a voxel centre the result is that voxel's own value, bit for bit."""
    from nbv_exploration_planner_b200 import LAYER_ESDF, PACK_VOXBLOX, NbvGpu
    vs, vps = 0.2, 16
    keys = np.array([[bx, by, bz] for bx in (0, 1) for by in (0, 1) for bz in (0, 1)], np.int32)
    idx = np.arange(vps ** 3)
    lx, ly, lz = idx % vps, (idx // vps) % vps, idx // (vps * vps)
    esdf = np.zeros((len(keys), vps ** 3), np.float32)
    for i, k in enumerate(keys):
        cx, cy, cz = (k[0] * vps + lx + 0.5) * vs, (k[1] * vps + ly + 0.5) * vs, (k[2] * vps + lz + 0.5) * vs
        esdf[i] = (cx + 0.5 * cy - 0.25 * cz).astype(np.float32)
    g = NbvGpu(0)
    le = g.layer_create(LAYER_ESDF, vs, vps, 16)
    g.layer_update(le, keys, esdf_words(esdf), PACK_VOXBLOX)
    g.commit()
    rng = np.random.default_rng(3)
    q = rng.uniform(0.5, 5.9, (2000, 3))
    r = g.esdf_distance_gradient(le, q)
    assert r["ok"].all()
    np.testing.assert_allclose(r["distance"], q[:, 0] + 0.5 * q[:, 1] - 0.25 * q[:, 2], rtol=0, atol=5e-6)
    np.testing.assert_allclose(r["gradient"], np.tile([1.0, 0.5, -0.25], (len(q), 1)), rtol=0, atol=2e-5)
    # voxel centres with the reference's own expression (block.h:90-92, common.h:188-201): float(block) * block_size
    # + float((float(voxel) + 0.5) * voxel_size), the sum in float -> offset exactly 0 -> the voxel's own value
    vi = rng.integers(2, 28, (500, 3))
    qc = _reference_centres(vi, vs, vps)
    rc = g.esdf_distance_gradient(le, qc)
    own = np.array([esdf[(b[0] * 2 + b[1]) * 2 + b[2], l[0] + vps * (l[1] + vps * l[2])] for b, l in zip(vi // vps, vi % vps)], np.float64)
    assert rc["ok"].all() and np.array_equal(rc["distance"], own)
    g.close()


def test_device_pointer_entry_point(oracle_mod, tiny_map):
    """nbvgpu_esdf_distance_gradient_device: torch tensors in HBM, stream-ordered; same results as the host entry point,
    and non-finite / out-of-range positions (which the host entry point rejects) fail with zero outputs."""
    import torch
    o, le_o, g, le_g = _setup(oracle_mod, tiny_map)
    q = _queries(tiny_map, np.random.default_rng(41), 3000)
    want = g.esdf_distance_gradient(le_g, q)
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        bad = np.array([[float("nan"), 0.0, 0.0], [float("inf"), 1.0, 1.0], [1e9, 0.0, 0.0]])
        d_q = torch.from_numpy(np.concatenate([q, bad])).cuda()
        n = d_q.shape[0]
        d_ok = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
        d_dist = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
        d_grad = torch.full((n, 3), -1.0, dtype=torch.float64, device="cuda")
        g.esdf_distance_gradient_device(le_g, n, d_q, d_ok, d_dist, d_grad)
        stream.synchronize()
    ok, dist, grad = d_ok.cpu().numpy(), d_dist.cpu().numpy(), d_grad.cpu().numpy()
    assert np.array_equal(ok[:-3], want["ok"]) and np.array_equal(dist[:-3], want["distance"]) and np.array_equal(grad[:-3], want["gradient"])
    assert not ok[-3:].any() and not dist[-3:].any() and not grad[-3:].any()
    g.close()


def test_failures_and_errors(oracle_mod, tiny_map):
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpu, NbvGpuError
    o, le_o, g, le_g = _setup(oracle_mod, tiny_map)
    far = np.array([[100.0, 100.0, 100.0], [-50.0, 0.0, 0.0]])        # no block there: ok = 0, outputs zero
    r = g.esdf_distance_gradient(le_g, far)
    assert not r["ok"].any() and not r["distance"].any() and not r["gradient"].any()
    assert g.esdf_distance_gradient(le_g, np.zeros((0, 3)))["ok"].shape == (0,)
    lt = g.layer_create(LAYER_TSDF, 0.2, 16, 8)
    with pytest.raises(NbvGpuError):
        g.esdf_distance_gradient(lt, [[0.0, 0.0, 0.0]])
    with pytest.raises(NbvGpuError):
        g.esdf_distance_gradient(le_g, [[float("nan"), 0.0, 0.0]])
    g.close()


def test_small_batches_eight_lanes_per_query(oracle_mod, tiny_map):
    """Up to 64 queries per call run on eight lanes each (distance, six gradient samples, the gradient's block probe) with
    mapped in-place I/O; one query per call also uses the completion ticket.  Same results as the one-thread kernel and
    the oracle, including failing queries with their partial gradient sums, and out-of-range device-resident inputs."""
    import torch
    o, le_o, g, le_g = _setup(oracle_mod, tiny_map)
    q = _queries(tiny_map, np.random.default_rng(59), 2000)
    ro = o.distance_and_gradient(le_o, q, threads=8)
    big = g.esdf_distance_gradient(le_g, q)
    for k in ("ok", "distance", "gradient"):
        assert np.array_equal(big[k], ro[k]), k
    part = np.flatnonzero((ro["ok"] == 0) & (np.abs(ro["gradient"]).sum(axis=1) > 0))
    assert len(part) > 0, "no query with a partially summed failing gradient in the sample"
    at = 0
    for n in (1, 1, 2, 7, 8, 9, 31, 64, 1, 33):
        r = g.esdf_distance_gradient(le_g, q[at:at + n])
        for k in ("ok", "distance", "gradient"):
            assert np.array_equal(r[k], ro[k][at:at + n]), (k, n, at)
        at += n
    for i in list(part[:20]) + list(np.flatnonzero(ro["ok"] == 1)[:20]):
        r = g.esdf_distance_gradient(le_g, q[i:i + 1])
        assert r["ok"][0] == ro["ok"][i] and r["distance"][0] == ro["distance"][i] and np.array_equal(r["gradient"][0], ro["gradient"][i])
    d_q = torch.from_numpy(np.concatenate([q[:9], [[float("nan"), 0.0, 0.0], [1e9, 0.0, 0.0]]])).cuda()
    d_ok = torch.full((11,), 7, dtype=torch.uint8, device="cuda")
    d_dist = torch.full((11,), -1.0, dtype=torch.float64, device="cuda")
    d_grad = torch.full((11, 3), -1.0, dtype=torch.float64, device="cuda")
    g.esdf_distance_gradient_device(le_g, 11, d_q, d_ok, d_dist, d_grad)
    g.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(d_ok.cpu().numpy()[:9], ro["ok"][:9]) and np.array_equal(d_dist.cpu().numpy()[:9], ro["distance"][:9])
    assert np.array_equal(d_grad.cpu().numpy()[:9], ro["gradient"][:9])
    assert not d_ok.cpu().numpy()[9:].any() and not d_dist.cpu().numpy()[9:].any() and not d_grad.cpu().numpy()[9:].any()
    g.close()
