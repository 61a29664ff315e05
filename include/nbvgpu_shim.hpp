// nbvgpu_shim.hpp -- header-only C++ shims with the reference's own signatures, over the C ABI.
//
// Drop these next to the reference's classes: the planner keeps calling
//   double RrtTree::optimized_gain(Pose& state)            nbveplanner/src/rrt.cpp:351
//   double RrtTree::gain(Pose& state)                      nbveplanner/src/rrt.cpp:481
//   bool   HighResManager::checkMotion<T>(start, end)      nbveplanner/include/nbveplanner/voxblox_manager.h:89
//   double History::bfs(const Eigen::Vector3d& point)      nbveplanner/src/history.cpp:103
//   bool   HighResManager::getDistanceAndGradientAtPosition(pos, &distance, &grad)   nbveplanner/src/voxblox_manager.cpp:143
// and the bodies become one-line calls into libnbvgpu.so.  No Eigen / ROS / Voxblox headers are
// needed here: `Pose` is anything indexable with [0..3] (Eigen::Vector4d), `Point` anything with
// x(), y(), z() (Eigen::Vector3d), `Layer`/`Block` anything with the Voxblox accessors used below.
//
// Error behaviour follows the reference's convention (glog CHECK aborts + bool returns): a failing
// ABI call throws std::runtime_error from the shim, which a maintainer may turn into LOG(FATAL).
#ifndef NBVGPU_SHIM_HPP_
#define NBVGPU_SHIM_HPP_

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "nbvgpu.h"

namespace nbvgpu_shim {

inline void check(nbvgpu_ctx* ctx, int rc, const char* what) {
  if (rc != NBVGPU_OK) throw std::runtime_error(std::string(what) + ": " + nbvgpu_last_error(ctx) + " (" + std::to_string(rc) + ")");
}

/// Owns the context and the two layers the planner reads: the low-resolution TSDF (gain) and the
/// high-resolution ESDF (collision), nbvep.cpp:21-27.
class GpuMap {
 public:
  GpuMap(int device, float lowres_voxel_size, int lowres_vps, size_t lowres_max_blocks, float hires_voxel_size, int hires_vps,
         size_t hires_max_blocks) {
    check(nullptr, nbvgpu_create(device, &ctx_), "nbvgpu_create");
    check(ctx_, nbvgpu_layer_create(ctx_, NBVGPU_LAYER_TSDF, lowres_voxel_size, lowres_vps, lowres_max_blocks, &tsdf_), "layer_create");
    check(ctx_, nbvgpu_layer_create(ctx_, NBVGPU_LAYER_ESDF, hires_voxel_size, hires_vps, hires_max_blocks, &esdf_), "layer_create");
  }
  /// The same on several GPUs of this box (nbvgpu_create_multi): the planner keeps ONE handle; batches are sharded over the
  /// GPUs, every upload is replicated by an NCCL broadcast inside nbvgpu_commit.  n = 1 calls run on the first GPU.
  GpuMap(const std::vector<int>& devices, float lowres_voxel_size, int lowres_vps, size_t lowres_max_blocks, float hires_voxel_size,
         int hires_vps, size_t hires_max_blocks) {
    check(nullptr, nbvgpu_create_multi((int)devices.size(), devices.data(), &ctx_), "nbvgpu_create_multi");
    check(ctx_, nbvgpu_layer_create(ctx_, NBVGPU_LAYER_TSDF, lowres_voxel_size, lowres_vps, lowres_max_blocks, &tsdf_), "layer_create");
    check(ctx_, nbvgpu_layer_create(ctx_, NBVGPU_LAYER_ESDF, hires_voxel_size, hires_vps, hires_max_blocks, &esdf_), "layer_create");
  }
  ~GpuMap() { nbvgpu_destroy(ctx_); }
  GpuMap(const GpuMap&) = delete;
  GpuMap& operator=(const GpuMap&) = delete;

  nbvgpu_ctx* ctx() const { return ctx_; }
  int tsdf_layer() const { return tsdf_; }
  int esdf_layer() const { return esdf_; }

#ifdef NBVGPU_SHIM_WITH_VOXBLOX
  /// Upload hook for the L0 -> L1 boundary (after TsdfServer::integratePointcloud, tsdf_server.cc:414,
  /// and EsdfServer::updateEsdf, esdf_server.cc:196): send every block whose Update::kMap bit is set
  /// (Layer::getAllUpdatedBlocks, layer.h:194-203) in the reference's own serialisation
  /// (Block::serializeToIntegers, block.cc:159-234), clear the bit, commit.  Needs the Voxblox
  /// headers, so it is only compiled inside the reference tree.
  /// The blocks are serialised straight into memory every GPU of the context can read (nbvgpu_shared_alloc) and staged from there
  /// without another copy (nbvgpu_layer_update_pinned): a single GPU copies them once, several GPUs each pull their slice of a
  /// large upload over their own PCIe link before an all-gather over NVLink completes it (nbvgpu.h).
  template <class VoxelType>
  void uploadUpdatedBlocks(voxblox::Layer<VoxelType>* layer, int gpu_layer) {
    voxblox::BlockIndexList updated;
    layer->getAllUpdatedBlocks(voxblox::Update::kMap, &updated);
    std::vector<int32_t> keys;
    std::vector<uint32_t> one;
    uint32_t* dst = nullptr;
    size_t per = 0, n = 0;
    for (const voxblox::BlockIndex& bi : updated) {
      typename voxblox::Block<VoxelType>::Ptr block = layer->getBlockPtrByIndex(bi);
      if (!block) continue;
      block->serializeToIntegers(&one);
      if (!dst) {  // every block of a layer serialises to the same number of words
        per = one.size();
        dst = stagingWords(per * updated.size());
      }
      if (one.size() != per) throw std::runtime_error("nbvgpu shim: blocks of one layer serialise to different sizes");
      std::memcpy(dst + n * per, one.data(), per * sizeof(uint32_t));
      keys.insert(keys.end(), {bi.x(), bi.y(), bi.z()});
      ++n;
      block->updated().reset(voxblox::Update::kMap);
    }
    if (n) {
      check(ctx_, nbvgpu_layer_update_pinned(ctx_, gpu_layer, n, keys.data(), dst, NBVGPU_PACK_VOXBLOX), "layer_update_pinned");
      check(ctx_, nbvgpu_commit(ctx_), "commit");
    }
  }
#endif
  /// Raw variant for callers that already hold serialised blocks (e.g. a voxblox_msgs/Layer message).
  void uploadSerialized(int gpu_layer, size_t n_blocks, const int32_t* keys, const uint32_t* data) {
    check(ctx_, nbvgpu_layer_update(ctx_, gpu_layer, n_blocks, keys, data, NBVGPU_PACK_VOXBLOX), "layer_update");
    check(ctx_, nbvgpu_commit(ctx_), "commit");
  }
  void clear() {  // VoxbloxManager::clear, voxblox_manager.cpp:92-95,150-154
    check(ctx_, nbvgpu_layer_clear(ctx_, tsdf_), "layer_clear");
    check(ctx_, nbvgpu_layer_clear(ctx_, esdf_), "layer_clear");
  }

 private:
  /// Staging for the upload hook: one shared allocation, grown when an upload needs more (the commit has returned by then).
  uint32_t* stagingWords(size_t words) {
    const size_t need = words * sizeof(uint32_t);
    if (need > staging_bytes_) {
      if (staging_) check(ctx_, nbvgpu_shared_free(ctx_, staging_), "shared_free");
      staging_ = nullptr;
      staging_bytes_ = need + need / 2 + (1u << 20);
      check(ctx_, nbvgpu_shared_alloc(ctx_, staging_bytes_, &staging_), "shared_alloc");
    }
    return static_cast<uint32_t*>(staging_);
  }
  nbvgpu_ctx* ctx_ = nullptr;
  int tsdf_ = -1, esdf_ = -1;
  void* staging_ = nullptr;
  size_t staging_bytes_ = 0;
};

/// What nbvePlanner::setUpCamera (nbvep.cpp:40-61) configures, as plain numbers.
inline void setUpCamera(GpuMap& map, double hfov_deg, double vfov_deg, double sensor_min_range, double gain_range,
                        const double q_CB_wxyz[4], const double t_CB[3], const double bbx_min[3], const double bbx_max[3],
                        double gain_free, double gain_occupied, double gain_unmapped) {
  nbvgpu_camera c;
  std::memset(&c, 0, sizeof(c));
  c.hfov_deg = hfov_deg; c.vfov_deg = vfov_deg; c.near_m = sensor_min_range; c.far_m = gain_range;
  for (int k = 0; k < 4; ++k) c.q_CB[k] = q_CB_wxyz[k];
  for (int k = 0; k < 3; ++k) { c.t_CB[k] = t_CB[k]; c.bbx_min[k] = bbx_min[k]; c.bbx_max[k] = bbx_max[k]; }
  c.gain_free = gain_free; c.gain_occupied = gain_occupied; c.gain_unmapped = gain_unmapped;
  c.sum_all_yaws = (hfov_deg == 360.0) ? 1 : 0;  // CameraModel::hasHorizontalLimit(), rrt.cpp:215-219
  check(map.ctx(), nbvgpu_set_camera(map.ctx(), &c), "set_camera");
}

/// double RrtTree::optimized_gain(Pose& state) / double RrtTree::gain(Pose& state): returns the gain and
/// writes the chosen yaw into state[3] (0 for gain()), exactly as rrt.cpp:477-478 / :577-578.
template <class Pose>
inline double optimized_gain(GpuMap& map, Pose& state) {
  double s[4] = {state[0], state[1], state[2], state[3]};
  double g = 0.0;
  check(map.ctx(), nbvgpu_optimized_gain(map.ctx(), map.tsdf_layer(), s, &g), "optimized_gain");
  state[3] = s[3];
  return g;
}

/// bool HighResManager::checkMotion<T>(const T& start, const T& end): true = free.
template <class T>
inline bool checkMotion(GpuMap& map, double robot_radius, const T& start, const T& end) {
  const double a[3] = {start.x(), start.y(), start.z()}, b[3] = {end.x(), end.y(), end.z()};
  int free_seg = 0;
  check(map.ctx(), nbvgpu_check_motion(map.ctx(), map.esdf_layer(), robot_radius, a, b, &free_seg), "check_motion");
  return free_seg != 0;
}

/// The body of RrtTree::iterate after the sampler (rrt.cpp:205-246) for a whole batch of samples in one call: checkMotion on
/// every edge, optimized_gain behind the free ones, node score against the parent's yaw / distance, first strict maximum
/// over zero_gain.  pos[n][3], seg[n][6] (parent xyz, end point with overshoot), parent_yaw[n], distance[n]; outputs
/// edge_free[n], gain[n], yaw_deg[n] (any may be null) and the best node.
inline nbvgpu_best evaluateCandidates(GpuMap& map, double robot_radius, size_t n, const double* pos, const double* seg,
                                      const double* parent_yaw, const double* distance, double v_max, double dyaw_max,
                                      double zero_gain, uint8_t* edge_free, double* gain, int32_t* yaw_deg) {
  nbvgpu_best best;
  check(map.ctx(), nbvgpu_evaluate_candidates(map.ctx(), map.tsdf_layer(), map.esdf_layer(), robot_radius, n, pos, seg, parent_yaw,
                                              distance, v_max, dyaw_max, zero_gain, edge_free, &best, gain, yaw_deg),
        "evaluate_candidates");
  return best;
}

/// bool HighResManager::getDistanceAndGradientAtPosition(const Point& pos, double* distance, Eigen::Vector3d* grad)
/// (voxblox_manager.cpp:143-148 -> EsdfMap, esdf_map.cc:22-52): the outputs are written in every case, as there.
template <class P, class G>
inline bool getDistanceAndGradientAtPosition(GpuMap& map, const P& pos, double* distance, G* grad) {
  const double p[3] = {pos.x(), pos.y(), pos.z()};
  uint8_t ok = 0;
  double g[3] = {0.0, 0.0, 0.0};
  check(map.ctx(), nbvgpu_esdf_distance_gradient(map.ctx(), map.esdf_layer(), 1, p, &ok, distance, g), "esdf_distance_gradient");
  (*grad)[0] = g[0];
  (*grad)[1] = g[1];
  (*grad)[2] = g[2];
  return ok != 0;
}

/// double History::bfs(const Eigen::Vector3d& point) (history.cpp:103-137): the frontier potential of one vertex,
/// frontierVoxels * pow(voxel_size, 3.0) on the low-resolution TSDF layer, 6.0 m radius as in the reference.
template <class T>
inline double bfs(GpuMap& map, const T& point) {
  const double p[3] = {point.x(), point.y(), point.z()};
  double gain = 0.0;
  check(map.ctx(), nbvgpu_frontier_potential(map.ctx(), map.tsdf_layer(), 1, p, 6.0, &gain, nullptr, nullptr, nullptr), "frontier_potential");
  return gain;
}

/// The batched form History::recalculatePotential wants when it sweeps the active vertices (history.cpp:196-222):
/// potential_gain[i] for points[3 i .. 3 i + 2], one call.
inline void bfs_batch(GpuMap& map, size_t n, const double* points_xyz, double* potential_gain) {
  check(map.ctx(), nbvgpu_frontier_potential(map.ctx(), map.tsdf_layer(), n, points_xyz, 6.0, potential_gain, nullptr, nullptr, nullptr),
        "frontier_potential");
}

}  // namespace nbvgpu_shim
#endif  // NBVGPU_SHIM_HPP_
