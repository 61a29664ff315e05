"""Writers for the reference's map formats, used only to build test inputs.

`.voxblox` files are encoded with the real protobuf library (python `google.protobuf`, message types
built from descriptors that transcribe voxblox/voxblox/proto/voxblox/{Layer,Block}.proto) and framed as
voxblox::utils::writeProtoMsgCountToStream / writeProtoMsgToStream do (protobuf_utils.cc:29-90): an
independent encoder for the hand-written reader in csrc/voxblox_io.cc.  voxblox_msgs/Layer messages
are written in ROS 1 wire serialisation with `struct`."""
import struct

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory


def _messages():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "voxblox_fixture.proto"
    fd.package = "voxblox"
    fd.syntax = "proto2"
    T = descriptor_pb2.FieldDescriptorProto
    blk = fd.message_type.add()
    blk.name = "BlockProto"
    for name, num, typ, label in (("voxels_per_side", 1, T.TYPE_INT32, T.LABEL_OPTIONAL), ("voxel_size", 2, T.TYPE_DOUBLE, T.LABEL_OPTIONAL),
                                  ("origin_x", 3, T.TYPE_DOUBLE, T.LABEL_OPTIONAL), ("origin_y", 4, T.TYPE_DOUBLE, T.LABEL_OPTIONAL),
                                  ("origin_z", 5, T.TYPE_DOUBLE, T.LABEL_OPTIONAL), ("has_data", 6, T.TYPE_BOOL, T.LABEL_OPTIONAL),
                                  ("voxel_data", 7, T.TYPE_UINT32, T.LABEL_REPEATED)):
        f = blk.field.add()
        f.name, f.number, f.type, f.label = name, num, typ, label
    lay = fd.message_type.add()
    lay.name = "LayerProto"
    for name, num, typ in (("voxel_size", 1, T.TYPE_DOUBLE), ("voxels_per_side", 2, T.TYPE_UINT32), ("type", 3, T.TYPE_STRING)):
        f = lay.field.add()
        f.name, f.number, f.type, f.label = name, num, typ, T.LABEL_OPTIONAL
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = message_factory.GetMessageClass
    return get(pool.FindMessageTypeByName("voxblox.LayerProto")), get(pool.FindMessageTypeByName("voxblox.BlockProto"))


def _varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def tsdf_words(tsdf: np.ndarray, rgba=None) -> np.ndarray:
    """Block::serializeToIntegers for TSDF (block.cc:159-183): distance bits, weight bits, rgba."""
    n, nvox, _ = tsdf.shape
    w = np.empty((n, nvox, 3), np.uint32)
    w[..., 0] = tsdf[..., 0].view(np.uint32)
    w[..., 1] = tsdf[..., 1].view(np.uint32)
    w[..., 2] = 0x7F7F7FFF if rgba is None else rgba
    return w.reshape(n, nvox * 3)


def esdf_words(esdf: np.ndarray, flags=None) -> np.ndarray:
    """Block::serializeToIntegers for ESDF (block.cc:203-234): distance bits, parent << 8 | flags."""
    n, nvox = esdf.shape
    w = np.empty((n, nvox, 2), np.uint32)
    w[..., 0] = esdf.view(np.uint32)
    w[..., 1] = 1 if flags is None else flags
    return w.reshape(n, nvox * 2)


def voxblox_file_bytes(layers) -> bytes:
    """layers: list of (type "tsdf"/"esdf", voxel_size float32, vps, keys int32[n][3], words uint32[n][W]).
    One file may hold several layers back to back (EsdfServer::saveMap appends the ESDF to the TSDF)."""
    LayerProto, BlockProto = _messages()
    out = bytearray()
    for typ, vs, vps, keys, words in layers:
        vs = float(np.float32(vs))
        out += _varint(1 + len(keys))                       # writeProtoMsgCountToStream (layer + blocks)
        lp = LayerProto(voxel_size=vs, voxels_per_side=vps, type=typ).SerializeToString()
        out += _varint(len(lp)) + lp                        # writeProtoMsgToStream
        block_size = np.float32(np.float32(vs) * np.float32(vps))
        for k, w in zip(keys, words):
            origin = (k.astype(np.float32) * block_size).astype(np.float64)   # getOriginPointFromGridIndex (float)
            bp = BlockProto(voxels_per_side=vps, voxel_size=vs, origin_x=origin[0], origin_y=origin[1], origin_z=origin[2],
                            has_data=True)
            bp.voxel_data.extend(int(x) for x in w)
            b = bp.SerializeToString()
            out += _varint(len(b)) + b
    return bytes(out)


def layer_msg_bytes(typ: str, voxel_size, vps: int, keys, words, action: int = 0) -> bytes:
    """ROS 1 serialisation of voxblox_msgs/Layer (msg/Layer.msg, msg/Block.msg)."""
    out = bytearray()
    out += struct.pack("<dI", float(np.float32(voxel_size)), vps)
    t = typ.encode()
    out += struct.pack("<I", len(t)) + t
    out += struct.pack("<B", action)
    out += struct.pack("<I", len(keys))
    for k, w in zip(keys, words):
        out += struct.pack("<iii", int(k[0]), int(k[1]), int(k[2]))
        out += struct.pack("<I", len(w)) + np.ascontiguousarray(w, "<u4").tobytes()
    return bytes(out)
