// voxblox_io.h -- readers for the reference's own map formats, host only.
//   * .voxblox files: varint message count, LayerProto, BlockProto... (voxblox/io/layer_io_inl.h:14-127,
//     voxblox/src/utils/protobuf_utils.cc:10-90, proto/voxblox/{Layer,Block}.proto)
//   * voxblox_msgs/Layer messages in ROS 1 wire serialisation (voxblox_msgs/msg/{Layer,Block}.msg,
//     voxblox_ros/conversions_inl.h:43-115)
// Both yield BlockIndex keys + the Block::serializeToIntegers words (NBVGPU_PACK_VOXBLOX payload).
#ifndef NBVGPU_VOXBLOX_IO_H_
#define NBVGPU_VOXBLOX_IO_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace nbvgpu_io {

struct ParsedBlocks {
  std::vector<int32_t> keys;    // [n][3] BlockIndex
  std::vector<uint32_t> words;  // [n][words_per_block]
  size_t n_blocks = 0;
  int action = 0;               // Layer.msg action (0 update, 1 merge, 2 reset); 0 for files
  std::string error;
};

// kind: 0 TSDF ("tsdf", 3 words per voxel), 1 ESDF ("esdf", 2 words per voxel).
// Returns false (with out->error) on malformed input or when no compatible layer is found.
bool parse_voxblox_file(const char* path, int kind, float voxel_size, int vps, ParsedBlocks* out);
bool parse_voxblox_bytes(const uint8_t* buf, size_t len, int kind, float voxel_size, int vps, ParsedBlocks* out);
bool parse_layer_msg(const uint8_t* buf, size_t len, int kind, float voxel_size, int vps, ParsedBlocks* out);

}  // namespace nbvgpu_io
#endif
