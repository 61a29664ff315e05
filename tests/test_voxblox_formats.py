"""ingest of the reference's own map formats.  CPU part: the hand-written reader parses
what the real protobuf library wrote (framing of voxblox/src/utils/protobuf_utils.cc, messages of
proto/voxblox/*.proto) and what ROS 1 would put on the wire for voxblox_msgs/Layer; GPU part: a map loaded
through those formats reads back identical and gives the same gains as the planar upload."""
import os

import numpy as np
import pytest

import voxblox_fixture as vf
from nbv_exploration_planner_b200 import LAYER_ESDF, LAYER_TSDF, NbvGpuError
from nbv_exploration_planner_b200.planner import parse_blocks

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _layers(m):
    return [("tsdf", m.voxel_size, m.vps, m.keys, vf.tsdf_words(m.tsdf)),
            ("esdf", m.voxel_size, m.vps, m.keys, vf.esdf_words(m.esdf))]


def test_voxblox_file_round_trip(tiny_map):
    data = vf.voxblox_file_bytes(_layers(tiny_map))            # TSDF followed by ESDF, as EsdfServer::saveMap writes
    keys, words, _ = parse_blocks(data, LAYER_TSDF, tiny_map.voxel_size, 16)
    assert np.array_equal(keys, tiny_map.keys) and np.array_equal(words, vf.tsdf_words(tiny_map.tsdf))
    keys, words, _ = parse_blocks(data, LAYER_ESDF, tiny_map.voxel_size, 16)      # second layer: the first is skipped
    assert np.array_equal(keys, tiny_map.keys) and np.array_equal(words, vf.esdf_words(tiny_map.esdf))


def test_committed_voxblox_fixture(tiny_map):
    """A small .voxblox file written by the protobuf library and committed (tests/golden/make_golden.py)."""
    data = open(os.path.join(GOLD, "tiny_tsdf_esdf.voxblox"), "rb").read()
    keys, words, _ = parse_blocks(data, LAYER_TSDF, tiny_map.voxel_size, 16)
    assert np.array_equal(keys, tiny_map.keys[:3]) and np.array_equal(words, vf.tsdf_words(tiny_map.tsdf[:3]))


def test_incompatible_and_malformed_files(tiny_map):
    data = vf.voxblox_file_bytes(_layers(tiny_map)[:1])
    for kind, vs, vps in ((LAYER_ESDF, tiny_map.voxel_size, 16),     # no esdf layer in the file
                          (LAYER_TSDF, 0.25, 16),                    # voxel size differs (Layer::isCompatible)
                          (LAYER_TSDF, tiny_map.voxel_size, 8)):     # voxels per side differ
        with pytest.raises(NbvGpuError):
            parse_blocks(data, kind, vs, vps)
    for bad in (b"", data[:100], data[:-7], b"\x00"):               # empty / truncated / zero message count
        with pytest.raises(NbvGpuError):
            parse_blocks(bad, LAYER_TSDF, tiny_map.voxel_size, 16)


def test_negative_block_indices_and_packed_encoding():
    """Block index = round(origin / block_size) for negative origins (common.h:180-185); a writer using the
    packed encoding for voxel_data (legal protobuf) is accepted too."""
    rng = np.random.default_rng(0)
    keys = np.array([[-3, 2, -1], [0, -7, 5], [-1, -1, -1]], np.int32)
    esdf = rng.uniform(-2, 2, (3, 512)).astype(np.float32)
    data = vf.voxblox_file_bytes([("esdf", 0.1, 8, keys, vf.esdf_words(esdf))])
    k, w, _ = parse_blocks(data, LAYER_ESDF, 0.1, 8)
    assert np.array_equal(k, keys) and np.array_equal(w, vf.esdf_words(esdf))
    # hand-built packed variant of one block message
    words = vf.esdf_words(esdf)[0]
    packed = b"".join(vf._varint(int(x)) for x in words)
    import struct
    blk = b"\x08\x08" + b"\x11" + struct.pack("<d", float(np.float32(0.1))) + b"\x19" + struct.pack("<d", float(np.float32(-3) * np.float32(0.8))) + \
        b"\x21" + struct.pack("<d", float(np.float32(2) * np.float32(0.8))) + b"\x29" + struct.pack("<d", float(np.float32(-1) * np.float32(0.8))) + \
        b"\x30\x01" + b"\x3a" + vf._varint(len(packed)) + packed
    lay = vf.voxblox_file_bytes([("esdf", 0.1, 8, keys[:0], [])])[1:]          # layer header without the count
    data2 = vf._varint(2) + lay + vf._varint(len(blk)) + blk
    k2, w2, _ = parse_blocks(data2, LAYER_ESDF, 0.1, 8)
    assert np.array_equal(k2, keys[:1]) and np.array_equal(w2[0], words)


def test_layer_msg_round_trip(tiny_map):
    w = vf.tsdf_words(tiny_map.tsdf)
    for action in (0, 2):
        msg = vf.layer_msg_bytes("tsdf", tiny_map.voxel_size, 16, tiny_map.keys, w, action)
        keys, words, act = parse_blocks(msg, LAYER_TSDF, tiny_map.voxel_size, 16, fmt="msg")
        assert act == action and np.array_equal(keys, tiny_map.keys) and np.array_equal(words, w)
    with pytest.raises(NbvGpuError):       # wrong layer type (conversions_inl.h:55-57)
        parse_blocks(vf.layer_msg_bytes("esdf", tiny_map.voxel_size, 16, tiny_map.keys, w), LAYER_TSDF, tiny_map.voxel_size, 16, fmt="msg")
    with pytest.raises(NbvGpuError):       # truncated
        parse_blocks(msg[:-5], LAYER_TSDF, tiny_map.voxel_size, 16, fmt="msg")


@pytest.mark.gpu
def test_gpu_map_from_reference_formats(oracle_mod, tiny_map, tmp_path):
    from conftest import load_gpu, load_oracle
    from nbv_exploration_planner_b200 import NbvGpu, synth
    cam = tiny_map.camera()
    path = tmp_path / "map.voxblox"
    path.write_bytes(vf.voxblox_file_bytes(_layers(tiny_map)))
    g = NbvGpu(0)
    lt = g.layer_create(LAYER_TSDF, tiny_map.voxel_size, 16, 64)
    le = g.layer_create(LAYER_ESDF, tiny_map.voxel_size, 16, 64)
    assert g.layer_load_voxblox_file(lt, str(path)) == len(tiny_map.keys)
    assert g.layer_load_voxblox_file(le, str(path)) == len(tiny_map.keys)
    g.set_camera(cam)
    ref, lt_r, le_r = load_gpu(tiny_map, cam)
    c = synth.make_candidates(tiny_map.cfg, 8, None)
    a, b = g.gain_eval(lt, c.positions, per_yaw=True), ref.gain_eval(lt_r, c.positions, per_yaw=True)
    assert np.array_equal(a["per_yaw"], b["per_yaw"]) and np.array_equal(a["gain"], b["gain"])
    assert np.array_equal(g.check_segments(le, 0.38, c.segments), ref.check_segments(le_r, 0.38, c.segments))
    o, lt_o, _ = load_oracle(oracle_mod, tiny_map, cam)
    assert np.array_equal(a["per_yaw"], o.gain_eval(lt_o, c.positions, per_yaw=True)["per_yaw"])
    # incremental Layer message (ACTION_UPDATE) flips three blocks to unknown; ACTION_RESET replaces the map
    unk = np.zeros((3, 4096, 2), np.float32)
    assert g.layer_apply_layer_msg(lt, vf.layer_msg_bytes("tsdf", tiny_map.voxel_size, 16, tiny_map.keys[:3], vf.tsdf_words(unk), 0)) == 3
    ref.layer_update(lt_r, tiny_map.keys[:3], unk); ref.commit()
    assert np.array_equal(g.gain_eval(lt, c.positions)["gain"], ref.gain_eval(lt_r, c.positions)["gain"])
    assert g.layer_num_blocks(lt) == len(tiny_map.keys)
    g.layer_apply_layer_msg(lt, vf.layer_msg_bytes("tsdf", tiny_map.voxel_size, 16, tiny_map.keys[:2], vf.tsdf_words(tiny_map.tsdf[:2]), 2))
    assert g.layer_num_blocks(lt) == 2
    with pytest.raises(NbvGpuError):       # ACTION_MERGE is refused
        g.layer_apply_layer_msg(lt, vf.layer_msg_bytes("tsdf", tiny_map.voxel_size, 16, tiny_map.keys[:2], vf.tsdf_words(tiny_map.tsdf[:2]), 1))
    g.close(); ref.close()
