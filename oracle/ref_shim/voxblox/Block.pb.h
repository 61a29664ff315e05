// Stand-in for the protoc-generated voxblox/Block.pb.h (proto/voxblox/Block.proto), TEST INFRASTRUCTURE.
#ifndef NBV_REF_SHIM_BLOCK_PB_
#define NBV_REF_SHIM_BLOCK_PB_
#include <cstdint>
#include <google/protobuf/message.h>
#include <vector>
namespace voxblox {
class BlockProto : public google::protobuf::Message {
 public:
  int32_t voxels_per_side() const { return vps_; }
  double voxel_size() const { return voxel_size_; }
  double origin_x() const { return o_[0]; }
  double origin_y() const { return o_[1]; }
  double origin_z() const { return o_[2]; }
  bool has_data() const { return has_data_; }
  const std::vector<uint32_t>& voxel_data() const { return data_; }
  int voxel_data_size() const { return (int)data_.size(); }
  void set_voxels_per_side(int32_t v) { vps_ = v; }
  void set_voxel_size(double v) { voxel_size_ = v; }
  void set_origin_x(double v) { o_[0] = v; }
  void set_origin_y(double v) { o_[1] = v; }
  void set_origin_z(double v) { o_[2] = v; }
  void set_has_data(bool v) { has_data_ = v; }
  void add_voxel_data(uint32_t v) { data_.push_back(v); }
 private:
  int32_t vps_ = 0;
  double voxel_size_ = 0, o_[3] = {0, 0, 0};
  bool has_data_ = false;
  std::vector<uint32_t> data_;
};
}  // namespace voxblox
#endif
