// voxblox_io.cc -- see voxblox_io.h.  A ~150-line protobuf wire-format reader (varint / 64-bit /
// length-delimited / 32-bit fields) is all the two messages need; there is no protobuf library in
// this image.
#include "voxblox_io.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

namespace nbvgpu_io {
namespace {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool eof() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) { ok = false; return 0; }
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  uint64_t fixed64() {
    if (end - p < 8) { ok = false; return 0; }
    uint64_t v;
    memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  uint32_t fixed32() {
    if (end - p < 4) { ok = false; return 0; }
    uint32_t v;
    memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  Reader sub(size_t n) {
    if ((size_t)(end - p) < n) { ok = false; return Reader{p, p}; }
    Reader r{p, p + n};
    p += n;
    return r;
  }
  void skip(int wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: fixed64(); break;
      case 2: { const uint64_t n = varint(); sub((size_t)n); break; }
      case 5: fixed32(); break;
      default: ok = false;
    }
  }
};

inline double as_double(uint64_t bits) { double d; memcpy(&d, &bits, 8); return d; }

struct LayerHeader { double voxel_size = 0; uint32_t vps = 0; std::string type; };

bool read_layer_proto(Reader r, LayerHeader* h) {  // proto/voxblox/Layer.proto
  while (!r.eof() && r.ok) {
    const uint64_t tag = r.varint();
    const int field = (int)(tag >> 3), wire = (int)(tag & 7);
    if (field == 1 && wire == 1) h->voxel_size = as_double(r.fixed64());
    else if (field == 2 && wire == 0) h->vps = (uint32_t)r.varint();
    else if (field == 3 && wire == 2) { const uint64_t n = r.varint(); Reader s = r.sub((size_t)n); if (r.ok) h->type.assign((const char*)s.p, (size_t)n); }
    else r.skip(wire);
  }
  return r.ok;
}

struct BlockHeader { int32_t vps = 0; double voxel_size = 0, origin[3] = {0, 0, 0}; };

bool read_block_proto(Reader r, BlockHeader* h, std::vector<uint32_t>* words) {  // proto/voxblox/Block.proto
  while (!r.eof() && r.ok) {
    const uint64_t tag = r.varint();
    const int field = (int)(tag >> 3), wire = (int)(tag & 7);
    if (field == 1 && wire == 0) h->vps = (int32_t)r.varint();
    else if (field == 2 && wire == 1) h->voxel_size = as_double(r.fixed64());
    else if (field >= 3 && field <= 5 && wire == 1) h->origin[field - 3] = as_double(r.fixed64());
    else if (field == 7 && wire == 0) words->push_back((uint32_t)r.varint());  // proto2 default: unpacked
    else if (field == 7 && wire == 2) {                                       // packed encoding is legal too
      const uint64_t n = r.varint();
      Reader s = r.sub((size_t)n);
      while (!s.eof() && s.ok) words->push_back((uint32_t)s.varint());
      if (!s.ok) r.ok = false;
    } else r.skip(wire);
  }
  return r.ok;
}

const char* type_name(int kind) { return kind == 0 ? "tsdf" : "esdf"; }  // voxblox/core/voxel.h:52-53
size_t words_per_voxel(int kind) { return kind == 0 ? 3 : 2; }           // block.cc:159-234

}  // namespace

bool parse_voxblox_bytes(const uint8_t* buf, size_t len, int kind, float voxel_size, int vps, ParsedBlocks* out) {
  Reader r{buf, buf + len};
  const size_t per_block = (size_t)vps * vps * vps * words_per_voxel(kind);
  // Layer members as the reference computes them (layer.h:34-44): float block size and its inverse.
  const float block_size = voxel_size * (float)vps;
  const float block_size_inv = (float)(1.0 / block_size);
  bool found = false;
  // LoadBlocksFromFile with multiple_layer_support (layer_io_inl.h:36-86): a file may hold several
  // layers back to back; incompatible ones are read and skipped.
  while (!r.eof() && !found) {
    const uint32_t num_protos = (uint32_t)r.varint();
    if (!r.ok) { out->error = "could not read number of messages"; return false; }
    if (num_protos == 0) { out->error = "empty protobuf file"; return false; }
    const uint64_t hsize = r.varint();
    LayerHeader lh;
    if (!r.ok || hsize == 0 || !read_layer_proto(r.sub((size_t)hsize), &lh) || !r.ok) {
      out->error = "could not read layer protobuf message";
      return false;
    }
    // Layer::isCompatible(LayerProto) (layer_inl.h:237-260)
    const bool compatible = std::fabs(lh.voxel_size - (double)voxel_size) < std::numeric_limits<float>::epsilon() &&
                            lh.vps == (uint32_t)vps && lh.type == type_name(kind);
    for (uint32_t b = 0; b + 1 < num_protos; ++b) {
      const uint64_t bsize = r.varint();
      if (!r.ok || bsize == 0) { out->error = "could not read block protobuf message"; return false; }
      Reader br = r.sub((size_t)bsize);
      if (!r.ok) { out->error = "truncated block message"; return false; }
      if (!compatible) continue;
      BlockHeader bh;
      const size_t w0 = out->words.size();
      if (!read_block_proto(br, &bh, &out->words)) { out->error = "could not parse block message"; return false; }
      // Layer::isCompatible(BlockProto) (layer_inl.h:263-270) and deserializeFromIntegers' size CHECK
      if (!(std::fabs(bh.voxel_size - (double)voxel_size) < std::numeric_limits<float>::epsilon()) || bh.vps != vps ||
          out->words.size() - w0 != per_block) {
        out->error = "block message incompatible with the layer";
        return false;
      }
      // Layer::addBlockFromProto: index = round(origin * block_size_inv) in float (common.h:180-185)
      for (int k = 0; k < 3; ++k) out->keys.push_back((int32_t)std::round((float)bh.origin[k] * block_size_inv));
      out->n_blocks++;
    }
    found = compatible;
  }
  if (!found) { out->error = "no compatible layer in the file"; return false; }
  return true;
}

bool parse_voxblox_file(const char* path, int kind, float voxel_size, int vps, ParsedBlocks* out) {
  FILE* f = fopen(path, "rb");
  if (!f) { out->error = std::string("could not open ") + path; return false; }
  std::vector<uint8_t> buf;
  uint8_t chunk[1 << 16];
  size_t n;
  while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
  fclose(f);
  return parse_voxblox_bytes(buf.data(), buf.size(), kind, voxel_size, vps, out);
}

// ROS 1 serialisation of voxblox_msgs/Layer: float64 voxel_size, uint32 voxels_per_side, string layer_type
// (uint32 length + bytes), uint8 action, Block[] (uint32 count; each int32 x, y, z, uint32[] data).
bool parse_layer_msg(const uint8_t* buf, size_t len, int kind, float voxel_size, int vps, ParsedBlocks* out) {
  Reader r{buf, buf + len};
  const double msg_vs = as_double(r.fixed64());
  const uint32_t msg_vps = r.fixed32();
  const uint32_t slen = r.fixed32();
  Reader s = r.sub(slen);
  if (!r.ok) { out->error = "truncated Layer message header"; return false; }
  const std::string type((const char*)s.p, slen);
  if (r.end - r.p < 1) { out->error = "truncated Layer message"; return false; }
  out->action = *r.p++;
  // deserializeMsgToLayer (conversions_inl.h:50-66): type and sizes must match
  if (type != type_name(kind)) { out->error = "layer type mismatch"; return false; }
  if (msg_vps != (uint32_t)vps || std::fabs(msg_vs - (double)voxel_size) > 1e-5) { out->error = "sizes don't match"; return false; }
  const size_t per_block = (size_t)vps * vps * vps * words_per_voxel(kind);
  const uint32_t nb = r.fixed32();
  for (uint32_t b = 0; b < nb && r.ok; ++b) {
    for (int k = 0; k < 3; ++k) out->keys.push_back((int32_t)r.fixed32());
    const uint32_t nw = r.fixed32();
    if (!r.ok || nw != per_block || (size_t)(r.end - r.p) < (size_t)nw * 4) { out->error = "bad block data size"; return false; }
    const size_t w0 = out->words.size();
    out->words.resize(w0 + nw);
    memcpy(out->words.data() + w0, r.p, (size_t)nw * 4);
    r.p += (size_t)nw * 4;
    out->n_blocks++;
  }
  if (!r.ok) { out->error = "truncated Layer message"; return false; }
  return true;
}

}  // namespace nbvgpu_io
