"""GPU mirror of the Voxblox block hash: every voxel written reads back equal, for both packings,
with overwrite / remove / clear, against the oracle's layer (core/layer.h, block.h semantics)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_layer(rng, n_blocks, vps, lo=-40, hi=40):
    keys = set()
    while len(keys) < n_blocks:
        keys.add(tuple(rng.integers(lo, hi, 3)))
    keys = np.array(sorted(keys), np.int32)
    tsdf = np.empty((n_blocks, vps ** 3, 2), np.float32)
    tsdf[..., 0] = rng.uniform(-0.6, 0.6, (n_blocks, vps ** 3))
    tsdf[..., 1] = rng.choice(np.array([0.0, 0.05, 0.1, np.float32(0.1), 0.2, 1.0, 50.0], np.float32), (n_blocks, vps ** 3))
    return keys, tsdf


def _all_voxels(keys, vps):
    l = np.arange(vps)
    z, y, x = np.meshgrid(l, l, l, indexing="ij")              # linear index = x + vps*(y + vps*z)
    local = np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.int64)
    return (keys.astype(np.int64)[:, None, :] * vps + local[None]).reshape(-1, 3)


@pytest.mark.parametrize("vps", [8, 16])
def test_every_voxel_reads_back(oracle_mod, vps):
    from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, PACK_VOXBLOX, NbvGpu
    rng = np.random.default_rng(vps)
    keys, tsdf = _random_layer(rng, 300, vps)
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, 0.1, vps, 512)
    g.layer_update(lt, keys, tsdf)
    g.commit()
    assert g.layer_num_blocks(lt) == 300
    gi = _all_voxels(keys, vps)
    vals, found, status = g.layer_read_voxels(lt, gi)
    assert found.all()
    assert np.array_equal(vals.reshape(tsdf.shape), tsdf)
    # status tile == getVoxelStatusByIndex rule
    w, d = tsdf[..., 1].ravel().astype(np.float64), tsdf[..., 0].ravel().astype(np.float64)
    want = np.where(w <= 1e-1, 0, np.where(d <= 0.0, 2, 1))
    assert np.array_equal(status, want)
    o = oracle_mod.Oracle()
    lo = o.layer_create(0, 0.1, vps)
    o.layer_update(lo, keys, tsdf)
    for i in rng.integers(0, len(gi), 200):
        assert o.voxel_status(lo, gi[i]) == status[i]
    # unallocated neighbours
    far = gi[:64] + np.array([1000 * vps, 0, 0])
    _, f2, s2 = g.layer_read_voxels(lt, far)
    assert not f2.any() and not s2.any()

    # Voxblox packing (block.cc:159-183): distance bits, weight bits, rgba
    packed = np.empty((300, vps ** 3, 3), np.uint32)
    packed[..., 0] = tsdf[..., 0].view(np.uint32)
    packed[..., 1] = tsdf[..., 1].view(np.uint32)
    packed[..., 2] = rng.integers(0, 2 ** 32, (300, vps ** 3), dtype=np.uint64).astype(np.uint32)
    lt2 = g.layer_create(LAYER_TSDF, 0.1, vps, 512)
    g.layer_update(lt2, keys, packed, PACK_VOXBLOX)
    g.commit()
    v2, f2, s2 = g.layer_read_voxels(lt2, gi)
    assert f2.all() and np.array_equal(v2, vals) and np.array_equal(s2, status)

    # ESDF, planar and Voxblox packing (block.cc:203-234: distance bits, parent|flags)
    esdf = rng.uniform(-2, 2, (300, vps ** 3)).astype(np.float32)
    le = g.layer_create(LAYER_ESDF, 0.1, vps, 512)
    g.layer_update(le, keys, esdf)
    pk = np.empty((300, vps ** 3, 2), np.uint32)
    pk[..., 0] = esdf.view(np.uint32)
    pk[..., 1] = rng.integers(0, 2 ** 32, (300, vps ** 3), dtype=np.uint64).astype(np.uint32)
    le2 = g.layer_create(LAYER_ESDF, 0.1, vps, 512)
    g.layer_update(le2, keys, pk, PACK_VOXBLOX)
    g.commit()
    for layer in (le, le2):
        ve, fe, _ = g.layer_read_voxels(layer, gi)
        assert fe.all() and np.array_equal(ve[:, 0].reshape(esdf.shape), esdf)
    g.close()


def test_overwrite_remove_clear_and_reuse():
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu
    rng = np.random.default_rng(9)
    vps = 16
    keys, tsdf = _random_layer(rng, 200, vps)
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, 0.15, vps, 256)
    g.layer_update(lt, keys[:150], tsdf[:150])
    g.commit()
    # incremental: overwrite 50 blocks, add 50 new ones, duplicate key inside one batch (last wins)
    new = tsdf.copy()
    new[100:150] = -new[100:150]
    g.layer_update(lt, keys[100:200], new[100:200])
    g.layer_update(lt, keys[100:101], tsdf[0:1])               # same key again before commit
    new[100] = tsdf[0]
    g.commit()
    assert g.layer_num_blocks(lt) == 200
    gi = _all_voxels(keys, vps)
    vals, found, _ = g.layer_read_voxels(lt, gi)
    assert found.all() and np.array_equal(vals.reshape(new.shape), new)
    # remove half (Layer::removeBlock), the rest is untouched
    g.layer_remove(lt, keys[::2])
    g.commit()
    assert g.layer_num_blocks(lt) == 100
    vals, found, _ = g.layer_read_voxels(lt, gi)
    f = found.reshape(200, -1)
    assert not f[::2].any() and f[1::2].all()
    assert np.array_equal(vals.reshape(new.shape)[1::2], new[1::2])
    # churn: re-insert / remove repeatedly (tombstones + slot reuse + rehash), pool must not leak
    for it in range(12):
        g.layer_update(lt, keys[::2], new[::2])
        g.commit()
        assert g.layer_num_blocks(lt) == 200
        g.layer_remove(lt, keys[::2])
        g.commit()
        assert g.layer_num_blocks(lt) == 100
    # update staged then removed before commit -> no-op; remove then update -> present
    g.layer_update(lt, keys[0:1], new[0:1]); g.layer_remove(lt, keys[0:1])
    g.layer_remove(lt, keys[1:2]); g.layer_update(lt, keys[1:2], new[0:1])
    g.commit()
    v, f, _ = g.layer_read_voxels(lt, _all_voxels(keys[0:2], vps))
    f = f.reshape(2, -1)
    assert not f[0].any() and f[1].all() and np.array_equal(v.reshape(2, -1, 2)[1], new[0])
    # clear (Layer::removeAllBlocks)
    g.layer_clear(lt)
    assert g.layer_num_blocks(lt) == 0
    _, found, _ = g.layer_read_voxels(lt, gi[:4096])
    assert not found.any()
    g.layer_update(lt, keys, tsdf)
    g.commit()
    assert g.layer_num_blocks(lt) == 200
    g.close()


def test_device_resident_update_matches_host_update():
    """nbvgpu_layer_update_device (the path fed by the NCCL map broadcast) == host-staged update."""
    import torch
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu
    rng = np.random.default_rng(11)
    keys, tsdf = _random_layer(rng, 64, 16)
    g = NbvGpu(0)
    la = g.layer_create(LAYER_TSDF, 0.15, 16, 128)
    lb = g.layer_create(LAYER_TSDF, 0.15, 16, 128)
    g.layer_update(la, keys, tsdf)
    g.commit()
    dk = torch.from_numpy(keys).cuda()
    dp = torch.from_numpy(tsdf).cuda()
    torch.cuda.synchronize()
    g.layer_update_device(lb, 64, dk, dp)
    gi = _all_voxels(keys, 16)
    va, fa, sa = g.layer_read_voxels(la, gi)
    vb, fb, sb = g.layer_read_voxels(lb, gi)
    assert fa.all() and fb.all() and np.array_equal(va, vb) and np.array_equal(sa, sb)
    g.close()


def test_removing_a_block_twice_in_one_commit_frees_it_once():
    """Layer::removeBlock on an already removed block is a no-op in the reference (layer.h:171-173), so one commit may
    carry the same key twice (and keys that were never there): the block count drops once per block, every tile slot
    stays unique, and a full refill of the pool still fits."""
    from nbv_exploration_planner_b200 import LAYER_TSDF, NbvGpu
    rng = np.random.default_rng(21)
    keys, tsdf = _random_layer(rng, 64, 8)
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, 0.1, 8, 64)          # pool of exactly 64 tiles
    g.layer_update(lt, keys, tsdf)
    g.commit()
    dup = np.concatenate([keys[:20], keys[:20], keys[5:15], [[999, 999, 999]]]).astype(np.int32)
    g.layer_remove(lt, dup)
    g.commit()
    assert g.layer_num_blocks(lt) == 44
    # refill: 20 new blocks must each get a tile of their own (a doubly freed slot would be handed out twice)
    new_keys = (keys[:20] + np.array([200, 0, 0])).astype(np.int32)
    new_tsdf = np.empty((20, 8 ** 3, 2), np.float32)
    new_tsdf[..., 0] = np.arange(20, dtype=np.float32)[:, None] + 0.5
    new_tsdf[..., 1] = 1.0
    g.layer_update(lt, new_keys, new_tsdf)
    g.commit()
    assert g.layer_num_blocks(lt) == 64
    vals, found, _ = g.layer_read_voxels(lt, _all_voxels(np.concatenate([keys[20:], new_keys]), 8))
    assert found.all()
    assert np.array_equal(vals.reshape(64, 8 ** 3, 2), np.concatenate([tsdf[20:], new_tsdf]))
    g.close()
