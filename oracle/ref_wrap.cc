// ref_wrap.cc -- C API over the REFERENCE'S OWN hot-path sources, compiled where they lie under
// /root/reference into oracle/_ref/libnbv_ref.so (recipe: oracle/Makefile target `ref`).
// TEST INFRASTRUCTURE: used only to pin the oracle restatement (tests/test_ref_pinning.py).
//
// What is the reference's and what is not:
//   * verbatim reference translation units, compiled as files: voxblox/src/integrator/integrator_utils.cc
//     (RayCaster, castRay) and nbveplanner/src/camera_model.cpp (frustum planes, isPointInView);
//   * verbatim reference headers: voxblox/core/{common,block,block_inl,layer,layer_inl,block_hash,voxel,
//     color}.h, voxblox/integrator/integrator_utils.h, voxblox/utils/timing.h, nbveplanner/{camera_model,
//     common}.h;
//   * function bodies that live in files which also include ROS headers are cut out BY LINE RANGE at
//     build time (sed into oracle/_ref/extract/, deleted after the compile; the Makefile checks the first
//     line of each range) and #included below inside class skeletons that only declare the members
//     those bodies use: RrtTree::optimized_gain (rrt.cpp:351-479), RrtTree::gain (:481-579),
//     VoxbloxManager::getVoxelStatusByIndex (voxblox_manager.cpp:14-33),
//     HighResManager::checkCollisionWithRobotAtVoxel (:121-129), getResolution / getVoxelsPerSide
//     (voxblox_manager.h:35-37), HighResManager::checkMotion<T> (voxblox_manager.h:89-108),
//     LowResManager::getVoxelStatus (voxblox_manager.cpp:73-90), convertToString + History::bfs
//     (history.cpp:96-137), EsdfMap::getDistanceAndGradientAtPosition (voxblox/src/core/esdf_map.cc:22-52),
//     RrtTree::iterate (rrt.cpp:163-264) with the reference's own kd-tree (kdtree/src/kdtree.c, compiled as a file) --
//     its clock-seeded std::mt19937 / std::uniform_real_distribution are replaced, by two #defines around the
//     #include of that range, with a generator that replays the caller's stream of uniforms;
//   * verbatim as well: voxblox/interpolator/interpolator{,_inl}.h (trilinear interpolation, central-difference
//     gradient) and voxblox/src/utils/evaluation_utils.cc (isObservedVoxel);
//   * NOT the reference's: the third-party headers it needs (Eigen, minkindr, glog, protoc output) --
//     stand-ins under oracle/ref_shim/ restating their published semantics -- and this C glue.
// Nothing from /root/reference is copied into the repository.
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <queue>
#include <random>
#include <unordered_map>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

#include "kdtree.h"  // the reference's own kd-tree (kdtree/include/kdtree/kdtree.h)
#include "nbveplanner/camera_model.h"
#include "voxblox/core/layer.h"
#include "voxblox/integrator/integrator_utils.h"
#include "voxblox/interpolator/interpolator.h"

#define SQ(x) ((x) * (x))            // nbveplanner/include/nbveplanner/rrt.h:22
#define CUBE(x) ((x) * (x) * (x))    // nbveplanner/include/nbveplanner/rrt.h:23

namespace voxblox {
// esdf_map.h pulls in Eigen::Ref / pybind conveniences; the two members the planner calls are cut out of esdf_map.cc
// and compiled inside this skeleton, which only declares what they use.
class EsdfMap {
 public:
  explicit EsdfMap(const Layer<EsdfVoxel>* layer) : interpolator_(layer) {}
  bool getDistanceAndGradientAtPosition(const Eigen::Vector3d& position, double* distance, Eigen::Vector3d* gradient) const;
  bool getDistanceAndGradientAtPosition(const Eigen::Vector3d& position, bool interpolate, double* distance,
                                        Eigen::Vector3d* gradient) const;
  Interpolator<EsdfVoxel> interpolator_;
};
#include "extract/esdf_map_cc_22_52.inc"
}  // namespace voxblox

namespace nbveplanner {

enum VoxelStatus { kUnknown, kOccupied, kFree };  // voxblox_manager.h:18

struct Params {  // the members of nbveplanner/params.h that the extracted bodies read
  CameraModel camera_model_;
  double camera_hfov_ = 90.0;
  double gain_free_ = 0.0, gain_occupied_ = 0.0, gain_unmapped_ = 1.0;
  double robot_radius_ = 0.5;
  Point bbx_min_, bbx_max_;  // params.h: the exploration bounds History::bfs tests against
  double extension_range_ = 1.0, dist_overshoot_ = 0.5, v_max_ = 1.0, dyaw_max_ = 0.5, zero_gain_ = 0.0;  // RrtTree::iterate
  bool log_ = false;
};

class VoxbloxManager {
 public:
  VoxelStatus getVoxelStatusByIndex(const voxblox::BlockIndex& block_idx, const voxblox::VoxelIndex& voxel_idx) const;
#include "extract/voxblox_manager_h_35_37.inc"
  Params* params_ = nullptr;
  voxblox::Layer<voxblox::TsdfVoxel>* tsdf_layer_ = nullptr;
  double voxel_size_ = 0.0;
};

class HighResManager : public VoxbloxManager {
 public:
  bool checkCollisionWithRobotAtVoxel(const voxblox::GlobalIndex& global_index) const;
#include "extract/voxblox_manager_h_89_108.inc"
  voxblox::Layer<voxblox::EsdfVoxel>* esdf_layer_ = nullptr;
};

class LowResManager : public VoxbloxManager {
 public:
  VoxelStatus getVoxelStatus(const Point& position) const;
};

class History {
 public:
  double bfs(const Eigen::Vector3d& point);
  LowResManager* manager_lowres_ = nullptr;
  Params* params_ = nullptr;
};

class Node {  // nbveplanner/include/nbveplanner/tree.h:17-32, constructor tree.cpp:5-12 (tree.h itself includes ROS headers)
 public:
  Pose state_;
  Node* parent_ = nullptr;
  std::vector<Node*> children_;
  double num_unmapped_ = 0.0;
  double time_to_reach_ = 0.0;
  double gain_ = 0.0;
  double distance_ = DBL_MAX;
  int id_ = -1;
};

class RrtTree {
 public:
  double optimized_gain(Pose& state);
  double gain(Pose& state);
  void iterate();
  void publishNode(Node* node) { last_published_ = node; }  // rrt.cpp publishes an RViz marker here; the wrapper takes note
  VoxbloxManager* manager_lowres_ = nullptr;
  HighResManager* manager_ = nullptr;
  Params* params_ = nullptr;
  // TreeBase / RrtTree members iterate() touches (tree.h:34-47, rrt.h)
  int counter_ = 0;
  double bestGain_ = 0.0;
  Node* bestNode_ = nullptr;
  Node* rootNode_ = nullptr;
  kdtree* kdTree_ = nullptr;
  double root_vicinity_ = 0.0;
  int g_ID_ = 0;
  std::ofstream file_tree_;
  Node* last_published_ = nullptr;
};

#include "extract/voxblox_manager_cpp_14_33.inc"
#include "extract/voxblox_manager_cpp_121_129.inc"
#include "extract/rrt_cpp_351_479.inc"
#include "extract/rrt_cpp_481_579.inc"
}  // namespace nbveplanner

// RrtTree::iterate draws from a clock-seeded std::mt19937 through std::uniform_real_distribution (rrt.cpp:172-178).  For a
// reproducible comparison the two names are redirected, for the lines of iterate() only, to a generator that hands out the
// caller's uniforms in order and throws when the stream is exhausted.
struct nbv_stream_end {};
struct nbv_uniform_stream {
  const double* u = nullptr;
  size_t n = 0, next = 0;
};
static nbv_uniform_stream g_stream;
namespace std {
struct nbv_stream_rng {
  template <class T>
  explicit nbv_stream_rng(T) {}
};
template <class T>
struct nbv_stream_dist {
  nbv_stream_dist(T, T) {}
  T operator()(nbv_stream_rng&) {
    if (g_stream.next >= g_stream.n) throw nbv_stream_end{};
    return g_stream.u[g_stream.next++];
  }
};
}  // namespace std
namespace nbveplanner {
#define mt19937 nbv_stream_rng
#define uniform_real_distribution nbv_stream_dist
#include "extract/rrt_cpp_163_264.inc"
#undef mt19937
#undef uniform_real_distribution
#include "extract/voxblox_manager_cpp_73_90.inc"
#include "extract/history_cpp_96_137.inc"

}  // namespace nbveplanner

namespace {
struct RefCtx {
  std::unique_ptr<voxblox::Layer<voxblox::TsdfVoxel>> tsdf;
  std::unique_ptr<voxblox::Layer<voxblox::EsdfVoxel>> esdf;
  std::unique_ptr<voxblox::Layer<voxblox::TsdfVoxel>> hires_tsdf;
  nbveplanner::Params params;
  nbveplanner::LowResManager low;
  nbveplanner::History history;
  std::unique_ptr<voxblox::EsdfMap> esdf_map;
  nbveplanner::HighResManager high;
  nbveplanner::RrtTree tree;
};
}  // namespace

extern "C" {

void* nbvr_create(float tsdf_voxel_size, int tsdf_vps, float esdf_voxel_size, int esdf_vps) {
  RefCtx* c = new RefCtx();
  c->tsdf.reset(new voxblox::Layer<voxblox::TsdfVoxel>(tsdf_voxel_size, tsdf_vps));
  c->esdf.reset(new voxblox::Layer<voxblox::EsdfVoxel>(esdf_voxel_size, esdf_vps));
  c->low.params_ = &c->params;
  c->low.tsdf_layer_ = c->tsdf.get();
  c->low.voxel_size_ = c->tsdf->voxel_size();
  c->high.params_ = &c->params;
  c->high.esdf_layer_ = c->esdf.get();
  // HighResManager::checkMotion scales by tsdf_layer_->voxel_size() of the HIGH-RES server's TSDF layer,
  // which has the ESDF's voxel size: an (empty) TSDF layer object of that size stands in for it.
  c->hires_tsdf.reset(new voxblox::Layer<voxblox::TsdfVoxel>(esdf_voxel_size, esdf_vps));
  c->high.tsdf_layer_ = c->hires_tsdf.get();
  c->tree.manager_lowres_ = &c->low;
  c->tree.manager_ = &c->high;
  c->tree.params_ = &c->params;
  c->esdf_map.reset(new voxblox::EsdfMap(c->esdf.get()));
  c->history.manager_lowres_ = &c->low;
  c->history.params_ = &c->params;
  return c;
}
void nbvr_destroy(void* c_) {
  delete static_cast<RefCtx*>(c_);
}

// planar payloads as in the oracle: TSDF [n][nvox][2] = (distance, weight), ESDF [n][nvox] = distance
void nbvr_tsdf_update(void* c_, size_t n, const int32_t* keys, const float* data) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  const size_t nvox = c->tsdf->voxels_per_side() * c->tsdf->voxels_per_side() * c->tsdf->voxels_per_side();
  for (size_t i = 0; i < n; ++i) {
    auto block = c->tsdf->allocateBlockPtrByIndex(voxblox::BlockIndex(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]));
    for (size_t v = 0; v < nvox; ++v) {
      voxblox::TsdfVoxel& vox = block->getVoxelByLinearIndex(v);
      vox.distance = data[(i * nvox + v) * 2];
      vox.weight = data[(i * nvox + v) * 2 + 1];
    }
  }
}
void nbvr_esdf_update(void* c_, size_t n, const int32_t* keys, const float* data) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  const size_t nvox = c->esdf->voxels_per_side() * c->esdf->voxels_per_side() * c->esdf->voxels_per_side();
  for (size_t i = 0; i < n; ++i) {
    auto block = c->esdf->allocateBlockPtrByIndex(voxblox::BlockIndex(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]));
    for (size_t v = 0; v < nvox; ++v) block->getVoxelByLinearIndex(v).distance = data[i * nvox + v];
  }
}
// ESDF blocks with explicit observed flags (EsdfVoxel::observed, voxel.h:21): observed[n][nvox] bytes
void nbvr_esdf_update_observed(void* c_, size_t n, const int32_t* keys, const float* data, const uint8_t* observed) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  const size_t nvox = c->esdf->voxels_per_side() * c->esdf->voxels_per_side() * c->esdf->voxels_per_side();
  for (size_t i = 0; i < n; ++i) {
    auto block = c->esdf->allocateBlockPtrByIndex(voxblox::BlockIndex(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]));
    for (size_t v = 0; v < nvox; ++v) {
      voxblox::EsdfVoxel& vox = block->getVoxelByLinearIndex(v);
      vox.distance = data[i * nvox + v];
      vox.observed = observed[i * nvox + v] != 0;
    }
  }
}
void nbvr_tsdf_remove(void* c_, size_t n, const int32_t* keys) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  for (size_t i = 0; i < n; ++i) c->tsdf->removeBlock(voxblox::BlockIndex(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]));
}

// nbvePlanner::setUpCamera (nbvep.cpp:40-61) with the TF lookup replaced by the given T_C_B
void nbvr_set_camera(void* c_, double hfov, double vfov, double near_m, double far_m, const double* q_CB_wxyz, const double* t_CB,
                     const double* bbx_min, const double* bbx_max, double g_free, double g_occ, double g_unmapped, double robot_radius) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  nbveplanner::Params& p = c->params;
  p.camera_hfov_ = hfov;
  p.gain_free_ = g_free; p.gain_occupied_ = g_occ; p.gain_unmapped_ = g_unmapped;
  p.robot_radius_ = robot_radius;
  p.camera_model_.setIntrinsicsFromFoV(hfov, vfov, near_m, far_m);
  nbveplanner::Point lo(bbx_min[0], bbx_min[1], bbx_min[2]), hi(bbx_max[0], bbx_max[1], bbx_max[2]);
  p.camera_model_.setBoundingBox(lo, hi);
  p.bbx_min_ = lo;  // the same Params::bbx_min_/bbx_max_ that nbvep.cpp:45 hands to the camera model
  p.bbx_max_ = hi;
  const nbveplanner::Transformation T_C_B(nbveplanner::Point(t_CB[0], t_CB[1], t_CB[2]),
                                          Eigen::Quaterniond(q_CB_wxyz[0], q_CB_wxyz[1], q_CB_wxyz[2], q_CB_wxyz[3]));
  p.camera_model_.setExtrinsics(T_C_B);
}

// RrtTree::optimized_gain / gain exactly as the planner calls them (rrt.cpp:215-219)
void nbvr_gain_eval(void* c_, size_t n, const double* pos, int sum_all, double* gain, double* yaw) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  for (size_t i = 0; i < n; ++i) {
    nbveplanner::Pose state(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 123.0);
    gain[i] = sum_all ? c->tree.gain(state) : c->tree.optimized_gain(state);
    yaw[i] = state[3];
  }
}
// RrtTree::iterate (rrt.cpp:163-264), called again and again on a tree that starts with the n_nodes given nodes (inserted
// into the reference's kd-tree in order), drawing from the stream of n triples `uniforms` until it is used up.  Per triple:
// status 0 = rejected by the sphere / bounds test, 1 = accepted but the edge collides (no node), 2 = node created; for
// created nodes position, parent (index: given nodes first, then created nodes in creation order), num_unmapped_ (gain),
// state_[3] (yaw), distance_, gain_ (score).  *best_index = triple of TreeBase::bestNode_ (-1: none above zero_gain).
int nbvr_iterate_stream(void* c_, const double* root, double vicinity, double extension_range, double overshoot, double v_max,
                        double dyaw_max, double zero_gain, size_t n_nodes, const double* node_state, const double* node_distance,
                        size_t n, const double* uniforms, uint8_t* status, double* pos, int32_t* parent, double* gain, double* yaw,
                        double* distance, double* score, int64_t* best_index) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  nbveplanner::RrtTree& T = c->tree;
  nbveplanner::Params& P = c->params;
  P.extension_range_ = extension_range; P.dist_overshoot_ = overshoot; P.v_max_ = v_max; P.dyaw_max_ = dyaw_max; P.zero_gain_ = zero_gain;
  T.root_vicinity_ = vicinity;
  T.bestGain_ = zero_gain;  // TreeBase::TreeBase, tree.cpp:26
  T.bestNode_ = nullptr;
  T.counter_ = 0;
  T.kdTree_ = kd_create(3);
  std::vector<std::unique_ptr<nbveplanner::Node>> nodes;  // flat ownership (Node::~Node of the reference deletes children)
  std::unordered_map<const nbveplanner::Node*, int64_t> index, triple_of;
  nbveplanner::Node root_node;
  root_node.state_ = nbveplanner::Pose(root[0], root[1], root[2], 0.0);
  T.rootNode_ = &root_node;
  for (size_t j = 0; j < n_nodes; ++j) {
    nodes.emplace_back(new nbveplanner::Node);
    nbveplanner::Node* nd = nodes.back().get();
    nd->state_ = nbveplanner::Pose(node_state[4 * j], node_state[4 * j + 1], node_state[4 * j + 2], node_state[4 * j + 3]);
    nd->distance_ = node_distance[j];
    index[nd] = (int64_t)j;
    kd_insert3(T.kdTree_, nd->state_.x(), nd->state_.y(), nd->state_.z(), nd);
  }
  for (size_t i = 0; i < n; ++i) { status[i] = 0; parent[i] = -1; gain[i] = yaw[i] = distance[i] = score[i] = 0.0; pos[3 * i] = pos[3 * i + 1] = pos[3 * i + 2] = 0.0; }
  g_stream.u = uniforms; g_stream.n = 3 * n; g_stream.next = 0;
  size_t created = 0;
  try {
    for (;;) {
      T.last_published_ = nullptr;
      T.iterate();
      const size_t t = g_stream.next / 3 - 1;  // the triple that passed the sphere / bounds tests
      if (T.last_published_) {
        nbveplanner::Node* nd = T.last_published_;
        nodes.emplace_back(nd);                 // iterate() allocated it with new
        index[nd] = (int64_t)(n_nodes + created++);
        triple_of[nd] = (int64_t)t;
        status[t] = 2;
        pos[3 * t] = nd->state_[0]; pos[3 * t + 1] = nd->state_[1]; pos[3 * t + 2] = nd->state_[2];
        parent[t] = (int32_t)index.at(nd->parent_);
        gain[t] = nd->num_unmapped_; yaw[t] = nd->state_[3]; distance[t] = nd->distance_; score[t] = nd->gain_;
      } else {
        status[t] = 1;
      }
    }
  } catch (const nbv_stream_end&) {
  }
  *best_index = T.bestNode_ ? triple_of.at(T.bestNode_) : -1;
  for (auto& nd : nodes) nd->children_.clear();
  kd_free(T.kdTree_);
  T.kdTree_ = nullptr;
  T.rootNode_ = nullptr;
  return 0;
}
void nbvr_check_segments(void* c_, size_t n, const double* seg, uint8_t* is_free) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  for (size_t i = 0; i < n; ++i) {
    const nbveplanner::Point a(seg[6 * i], seg[6 * i + 1], seg[6 * i + 2]), b(seg[6 * i + 3], seg[6 * i + 4], seg[6 * i + 5]);
    is_free[i] = c->high.checkMotion<nbveplanner::Point>(a, b) ? 1 : 0;
  }
}
// voxblox::castRay (integrator_utils.h:133-143)
size_t nbvr_cast_ray(const float* s, const float* e, int64_t* out, size_t cap) {
  voxblox::AlignedVector<voxblox::GlobalIndex> idx;
  voxblox::castRay(voxblox::Point(s[0], s[1], s[2]), voxblox::Point(e[0], e[1], e[2]), &idx);
  for (size_t i = 0; i < idx.size() && i < cap; ++i) { out[3 * i] = idx[i].x(); out[3 * i + 1] = idx[i].y(); out[3 * i + 2] = idx[i].z(); }
  return idx.size();
}
// per-yaw planes / far-plane points of the real CameraModel for one body pose (rrt.cpp:369-377)
void nbvr_yaw_frame(void* c_, const double* pos, int yaw_deg, double* planes /*[6][4]*/, double* cam /*[3]*/, double* pp /*[3][3]*/) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  const double yaw = M_PI * yaw_deg / 180.0;
  Eigen::Quaterniond orientation;
  orientation = Eigen::AngleAxisd(yaw, nbveplanner::Point::UnitZ());
  const nbveplanner::Transformation T_G_B(nbveplanner::Point(pos[0], pos[1], pos[2]), orientation);
  c->params.camera_model_.setBodyPose(T_G_B);
  const auto& bp = c->params.camera_model_.getBoundingPlanes();
  for (int k = 0; k < 6; ++k) {
    const nbveplanner::Point nrm = bp[k].normal();
    planes[4 * k] = nrm.x(); planes[4 * k + 1] = nrm.y(); planes[4 * k + 2] = nrm.z(); planes[4 * k + 3] = bp[k].distance();
  }
  const nbveplanner::Point cp = c->params.camera_model_.getCameraPose().getPosition();
  cam[0] = cp.x(); cam[1] = cp.y(); cam[2] = cp.z();
  nbveplanner::AlignedVector<nbveplanner::Point> pts;
  c->params.camera_model_.getFarPlanePoints(0, &pts);
  for (int k = 0; k < 3; ++k) { pp[3 * k] = pts[k].x(); pp[3 * k + 1] = pts[k].y(); pp[3 * k + 2] = pts[k].z(); }
}
int nbvr_in_view(void* c_, const double* p) {
  return static_cast<RefCtx*>(c_)->params.camera_model_.isPointInView(nbveplanner::Point(p[0], p[1], p[2])) ? 1 : 0;
}
// index helpers (voxblox/core/common.h, block_hash.h, block_inl.h) for the KAT comparison
void nbvr_block_and_local(const int64_t* g, int vps, int32_t* block, int32_t* local) {
  const voxblox::GlobalIndex gi(g[0], g[1], g[2]);
  const voxblox::BlockIndex b = voxblox::getBlockIndexFromGlobalVoxelIndex(gi, 1.0 / vps);
  const voxblox::VoxelIndex l = voxblox::getLocalFromGlobalVoxelIndex(gi, vps);
  for (int k = 0; k < 3; ++k) { block[k] = b[k]; local[k] = l[k]; }
}
void nbvr_grid_index_from_point(const float* p, float inv, int64_t* out) {
  const voxblox::GlobalIndex g = voxblox::getGridIndexFromPoint<voxblox::GlobalIndex>(voxblox::Point(p[0], p[1], p[2]), inv);
  for (int k = 0; k < 3; ++k) out[k] = g[k];
}
uint64_t nbvr_block_hash(int x, int y, int z) { return voxblox::AnyIndexHash()(voxblox::AnyIndex(x, y, z)); }
int nbvr_voxel_status(void* c_, const int32_t* block, const int32_t* voxel) {
  // oracle codes: 0 unknown, 1 free, 2 occupied
  const nbveplanner::VoxelStatus s = static_cast<RefCtx*>(c_)->low.getVoxelStatusByIndex(
      voxblox::BlockIndex(block[0], block[1], block[2]), voxblox::VoxelIndex(voxel[0], voxel[1], voxel[2]));
  return s == nbveplanner::kUnknown ? 0 : (s == nbveplanner::kFree ? 1 : 2);
}

// History::bfs (history.cpp:103-137) for each vertex position; the reference hard-codes 6.0 m.
void nbvr_frontier_potential(void* c_, size_t n, const double* points, double* gain) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  for (size_t i = 0; i < n; ++i) gain[i] = c->history.bfs(Eigen::Vector3d(points[3 * i], points[3 * i + 1], points[3 * i + 2]));
}
// LowResManager::getVoxelStatus (voxblox_manager.cpp:73-90); oracle codes: 0 unknown, 1 free, 2 occupied
int nbvr_voxel_status_at(void* c_, const double* p) {
  const nbveplanner::VoxelStatus s = static_cast<RefCtx*>(c_)->low.getVoxelStatus(nbveplanner::Point(p[0], p[1], p[2]));
  return s == nbveplanner::kUnknown ? 0 : (s == nbveplanner::kFree ? 1 : 2);
}
// EsdfMap::getDistanceAndGradientAtPosition (esdf_map.cc:22-52) per query, as History::refineVertexPosition calls it
// (history.cpp:236): ok[i] = return value, distance[i], gradient[i][3] (written whatever the return value, as there).
void nbvr_distance_and_gradient(void* c_, size_t n, const double* positions, uint8_t* ok, double* distance, double* gradient) {
  RefCtx* c = static_cast<RefCtx*>(c_);
  for (size_t i = 0; i < n; ++i) {
    double d = 0.0;
    Eigen::Vector3d g = Eigen::Vector3d::Zero();
    ok[i] = c->esdf_map->getDistanceAndGradientAtPosition(Eigen::Vector3d(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]), &d, &g) ? 1 : 0;
    distance[i] = d;
    gradient[3 * i] = g.x(); gradient[3 * i + 1] = g.y(); gradient[3 * i + 2] = g.z();
  }
}
}  // extern "C"
