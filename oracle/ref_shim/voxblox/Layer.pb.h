// Stand-in for the protoc-generated voxblox/Layer.pb.h (proto/voxblox/Layer.proto), TEST INFRASTRUCTURE.
#ifndef NBV_REF_SHIM_LAYER_PB_
#define NBV_REF_SHIM_LAYER_PB_
#include <cstdint>
#include <google/protobuf/message.h>
#include <string>
namespace voxblox {
class LayerProto : public google::protobuf::Message {
 public:
  double voxel_size() const { return voxel_size_; }
  uint32_t voxels_per_side() const { return vps_; }
  const std::string& type() const { return type_; }
  void set_voxel_size(double v) { voxel_size_ = v; }
  void set_voxels_per_side(uint32_t v) { vps_ = v; }
  void set_type(const std::string& t) { type_ = t; }
 private:
  double voxel_size_ = 0;
  uint32_t vps_ = 0;
  std::string type_;
};
}  // namespace voxblox
#endif
