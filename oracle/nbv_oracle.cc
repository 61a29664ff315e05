// nbv_oracle.cc -- CPU ORACLE for the NBV viewpoint-gain hot path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (libnbvgpu.so, the CUDA
// path behind include/nbvgpu.h) never links, loads or calls anything in here.
//
// What it is: a dependency-free C++17 restatement of the reference's CPU
// algorithm for the path, following the reference line by line in arithmetic
// (double frustum / float Voxblox RayCaster / int64 voxel indices, round to
// nearest, NO fused multiply-add: build with -ffp-contract=off, no -march).
// Every function cites the reference file:line (relative to /root/reference)
// it restates.  Nothing is copied: the reference is Eigen/minkindr/ROS code,
// this file is scalar C++ over plain structs.
//
// PINNING.  The reference holds no test, golden vector or fixture for castRay / optimized_gain /
// checkMotion (its only applicable known answers are the index-math KATs of
// voxblox/voxblox/test/test_tsdf_map.cc, replayed in tests/test_oracle_kats.py), and it cannot be built as
// it is (ROS, Eigen, glog, minkindr, protobuf are absent).  The oracle is therefore pinned against the
// reference's OWN SOURCE instead: oracle/_ref/libnbv_ref.so (oracle/Makefile `ref`, oracle/ref_wrap.cc)
// compiles the reference's integrator_utils.cc and camera_model.cpp verbatim, its Voxblox core headers,
// and the bodies of RrtTree::optimized_gain / gain, getVoxelStatusByIndex, checkCollisionWithRobotAtVoxel
// and checkMotion cut out of rrt.cpp / voxblox_manager.{h,cpp} by line range, and -- for the widened rows --
// LowResManager::getVoxelStatus, History::bfs (history.cpp:96-137), EsdfMap::getDistanceAndGradientAtPosition
// (esdf_map.cc:22-52) with the interpolator headers and evaluation_utils.cc verbatim, against stand-ins for the
// third-party headers only (oracle/ref_shim/).  tests/test_ref_pinning.py compares this file with that
// library bit for bit (ray index sequences incl. the degenerate ones, index math, hash, per-yaw planes /
// camera positions / far-plane points, gains and yaws, edge verdicts, voxel status, the cfg-1 goldens, flood-fill
// gains incl. the -0.0 string quirk, interpolated distances and gradients incl. the partial sums of failures).
// What stays an assumption is the arithmetic of the un-vendored, un-pinned third-party libraries (Eigen's
// dot / normalized / quaternion evaluation order, the summation order of the interpolator's two fixed-size float
// products, minkindr's compose / inverse): it is restated once in
// the stand-ins and once in the "Eigen semantics" section below, as single switchable functions.
//
// Eigen / minkindr are third-party dependencies that are NOT vendored under /root/reference
// (dependencies.rosinstall lists them as bare git URLs with no version pin); their published
// algorithms are restated here.

#include <algorithm>
#include <cstdio>
#include <queue>
#include <string>
#include <unordered_set>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// Eigen semantics (assumed; unpinned -- see header).
// ---------------------------------------------------------------------------
struct V3 {
  double x, y, z;
};
struct Quat {
  double w, x, y, z;
};

inline V3 add(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 scale(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }

// Eigen 3.3 redux on a fixed 3-vector of double with SSE2: one 2-wide packet
// (x,y) then the scalar tail z  =>  (x0*y0 + x1*y1) + x2*y2.
inline double dot(const V3& a, const V3& b) {
#ifdef NBVO_DOT_SCALAR_UNROLLER
  return a.x * b.x + (a.y * b.y + a.z * b.z);
#else
  return (a.x * b.x + a.y * b.y) + a.z * b.z;
#endif
}
inline V3 cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double squared_norm(const V3& a) { return dot(a, a); }
inline double norm(const V3& a) { return std::sqrt(squared_norm(a)); }
// MatrixBase::normalized(): z = squaredNorm(); z > 0 ? v / sqrt(z) : v
// (coefficient-wise true division, Eigen >= 3.3).
inline V3 normalized(const V3& a) {
  const double z = squared_norm(a);
  if (z > 0.0) {
    const double n = std::sqrt(z);
    return {a.x / n, a.y / n, a.z / n};
  }
  return a;
}
// QuaternionBase::_transformVector: uv = q.vec x v; uv += uv;
// v + q.w*uv + q.vec x uv.
inline V3 rotate(const Quat& q, const V3& v) {
  const V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = add(uv, uv);
  const V3 c = cross(qv, uv);
  return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}
// Generic Eigen quaternion product (left to right sums).  With a yaw-only left
// operand (x = y = 0) every component has two non-zero products, so the sum
// order cannot change the rounding.
inline Quat qmul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
          a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Quat conj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }

// ---------------------------------------------------------------------------
// Voxblox index math.
// ---------------------------------------------------------------------------
constexpr float kEpsilon = 1e-6f;  // voxblox/core/common.h:142

// getGridIndexFromPoint(scaled) -- voxblox/core/common.h:166-171.
inline int64_t grid_index_scaled(float v) { return (int64_t)std::floor(v + kEpsilon); }
// getGridIndexFromPoint(point, inv) -- voxblox/core/common.h:151-158.
inline int64_t grid_index(float v, float inv) { return (int64_t)std::floor(v * inv + kEpsilon); }
// signum -- voxblox/core/common.h:258.
inline int signum(float x) { return (x == 0) ? 0 : x < 0 ? -1 : 1; }
// getBlockIndexFromGlobalVoxelIndex -- voxblox/core/common.h:215-224.
inline int32_t block_from_global(int64_t g, float vps_inv) {
  return (int32_t)std::floor(static_cast<float>(g) * vps_inv);
}
// getLocalFromGlobalVoxelIndex -- voxblox/core/common.h:233-243.
inline int32_t local_from_global(int64_t g, int vps) {
  constexpr int offset = 1 << (8 * sizeof(int) - 1);  // INT_MIN, as in the reference
  return (int32_t)((g + offset) & (vps - 1));
}
// Block::computeLinearIndexFromVoxelIndex -- voxblox/core/block_inl.h:12-27.
inline size_t linear_index(int vps, int x, int y, int z) {
  return static_cast<size_t>(x + vps * (y + z * vps));
}
// AnyIndexHash -- voxblox/core/block_hash.h:20-31.
struct Key {
  int32_t x, y, z;
  bool operator==(const Key& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct KeyHash {
  static constexpr size_t sl = 17191;
  static constexpr size_t sl2 = sl * sl;
  size_t operator()(const Key& k) const {
    return static_cast<unsigned int>(k.x + k.y * sl + k.z * sl2);
  }
};

// RayCaster -- voxblox/src/integrator/integrator_utils.cc:111-179.
struct RayCaster {
  int64_t cur[3];
  int sgn[3];
  float t[3], dt[3];
  unsigned len, step;

  RayCaster(const float s[3], const float e[3]) {
    // integrator_utils.cc:129-134: NaN input returns early with current_step_
    // uninitialised (UB in the reference); the ABI rejects non-finite input, the
    // oracle emits nothing for it.
    if (std::isnan(s[0]) || std::isnan(s[1]) || std::isnan(s[2]) || std::isnan(e[0]) ||
        std::isnan(e[1]) || std::isnan(e[2])) {
      len = 0;
      step = 1;
      for (int k = 0; k < 3; ++k) { cur[k] = 0; sgn[k] = 0; t[k] = dt[k] = 0.f; }
      return;
    }
    int64_t diff[3];
    for (int k = 0; k < 3; ++k) {
      cur[k] = grid_index_scaled(s[k]);            // :136
      diff[k] = grid_index_scaled(e[k]) - cur[k];  // :137-138
    }
    step = 0;                                                                  // :140
    len = (unsigned)(std::abs(diff[0]) + std::abs(diff[1]) + std::abs(diff[2]));  // :142-143
    for (int k = 0; k < 3; ++k) {
      const float r = e[k] - s[k];                 // :145
      sgn[k] = signum(r);                          // :147-148
      const int cs = std::max(0, sgn[k]);          // :150-152
      const float shifted = s[k] - (float)cur[k];  // :154-155
      const float dist = (float)cs - shifted;      // :157-158
      // :160-166 / :170-178 -- the `abs(r) < 0.0 ? 2.0 :` guard can never fire,
      // so a zero component divides by zero (+-inf / NaN), as in the reference.
      t[k] = (std::abs(r) < 0.0) ? 2.0 : dist / r;
      dt[k] = (std::abs(r) < 0.0) ? 2.0 : sgn[k] / r;
    }
  }
  // nextRayIndex -- integrator_utils.cc:111-125; minCoeff keeps the first
  // minimum (`value < best`), NaN never wins unless it is element 0.
  bool next(int64_t out[3]) {
    if (step++ > len) return false;
    out[0] = cur[0]; out[1] = cur[1]; out[2] = cur[2];
    int m = 0;
    float best = t[0];
    if (t[1] < best) { best = t[1]; m = 1; }
    if (t[2] < best) { best = t[2]; m = 2; }
    cur[m] += sgn[m];
    t[m] += dt[m];
    return true;
  }
};

// ---------------------------------------------------------------------------
// Layers: unordered_map<BlockIndex, dense vps^3 tile>, voxblox/core/layer.h.
// ---------------------------------------------------------------------------
struct TsdfVoxel { float distance, weight; };  // voxel.h:12-16 (colour unused)
struct Layer {
  int kind;  // 0 TSDF, 1 ESDF
  float voxel_size;
  int vps;
  float vps_inv;
  size_t nvox;
  std::unordered_map<Key, std::unique_ptr<float[]>, KeyHash> blocks;  // TSDF: 2 floats/voxel
  std::unordered_map<Key, std::unique_ptr<uint8_t[]>, KeyHash> observed;  // ESDF: EsdfVoxel::observed (voxel.h:21)
};

struct Camera {
  double hfov_deg, vfov_deg, near_m, far_m;
  double q_CB[4];  // w x y z
  double t_CB[3];
  double bbx_min[3], bbx_max[3];
  double gain_free, gain_occupied, gain_unmapped;
  int32_t sum_all_yaws;
  int32_t reserved;
};

struct Plane {  // camera_model.cpp:5-23
  V3 n;
  double d;
  void set_from_points(const V3& p1, const V3& p2, const V3& p3) {
    const V3 p1p2 = sub(p2, p1);
    const V3 p1p3 = sub(p3, p1);
    n = normalized(cross(p1p2, p1p3));
    d = dot(n, p1);
  }
  bool correct_side(const V3& p) const { return dot(p, n) >= d; }
};

struct Ctx {
  std::vector<std::unique_ptr<Layer>> layers;
  Camera cam{};
  bool cam_set = false;
  V3 corners_C[8];
  Quat q_BC;
  V3 t_BC;
};

// CameraModel::setIntrinsicsFromFoV -- camera_model.cpp:25-83 (6-plane mode).
void set_intrinsics(Ctx* c) {
  const Camera& cam = c->cam;
  // camera_model.cpp:35-42: hfov == 360 builds the first 8 corners for a 90 degree sector.
  const double hfov = (cam.hfov_deg == 360.0 ? 90.0 : cam.hfov_deg) * M_PI / 180.0;
  const double vfov = cam.vfov_deg * M_PI / 180.0;
  const double th = std::tan(hfov / 2.0);
  const double tv = std::tan(vfov / 2.0);
  const double n = cam.near_m, f = cam.far_m;
  c->corners_C[0] = {n, n * th, n * tv};
  c->corners_C[1] = {n, n * th, -n * tv};
  c->corners_C[2] = {n, -n * th, -n * tv};
  c->corners_C[3] = {n, -n * th, n * tv};
  c->corners_C[4] = {f, f * th, f * tv};
  c->corners_C[5] = {f, f * th, -f * tv};
  c->corners_C[6] = {f, -f * th, -f * tv};
  c->corners_C[7] = {f, -f * th, f * tv};
  // T_C_B.inverse() (camera_model.cpp:92-94): minkindr inverts a unit
  // quaternion by conjugation and the translation as -(q^-1 . t).
  const Quat q_CB{cam.q_CB[0], cam.q_CB[1], cam.q_CB[2], cam.q_CB[3]};
  c->q_BC = conj(q_CB);
  const V3 r = rotate(c->q_BC, V3{cam.t_CB[0], cam.t_CB[1], cam.t_CB[2]});
  c->t_BC = {-r.x, -r.y, -r.z};
}

struct YawFrame {
  V3 cam_pos;
  Plane planes[6];
  V3 pp0, pp1, pp2;
};

// rrt.cpp:369-377 + camera_model.cpp:87-94,98-164,280-289 for one yaw.
void yaw_frame(const Ctx* c, const V3& p, int i, YawFrame* f) {
  const double yaw = M_PI * i / 180.0;  // rrt.cpp:369
  // Eigen::AngleAxisd(yaw, UnitZ) -> Quaterniond: w = cos(a/2), vec = sin(a/2)*axis.
  const double ha = 0.5 * yaw;
  const double s = std::sin(ha);
  const Quat q_yaw{std::cos(ha), s * 0.0, s * 0.0, s * 1.0};
  // T_G_C = T_G_B * T_B_C : (qa*qb, ta + qa.rotate(tb)).
  const Quat q_GC = qmul(q_yaw, c->q_BC);
  const V3 t_GC = add(p, rotate(q_yaw, c->t_BC));
  f->cam_pos = t_GC;
  V3 g[8];
  for (int k = 0; k < 8; ++k) g[k] = add(rotate(q_GC, c->corners_C[k]), t_GC);  // :119-121
  f->planes[0].set_from_points(g[0], g[2], g[1]);  // near   :125
  f->planes[1].set_from_points(g[4], g[5], g[6]);  // far    :131
  f->planes[2].set_from_points(g[3], g[6], g[2]);  // left   :138
  f->planes[3].set_from_points(g[0], g[5], g[4]);  // right  :145
  f->planes[4].set_from_points(g[3], g[4], g[7]);  // top    :152
  f->planes[5].set_from_points(g[2], g[6], g[5]);  // bottom :159
  f->pp0 = g[4]; f->pp1 = g[5]; f->pp2 = g[6];     // getFarPlanePoints :280-289
}

// isInsideBounds (nbveplanner/common.h:22-31) + CameraModel::isPointInView
// (camera_model.cpp:191-201).
inline bool in_view(const Camera& cam, const YawFrame& f, const V3& p) {
  if (!(p.x >= cam.bbx_min[0] && p.x <= cam.bbx_max[0] && p.y >= cam.bbx_min[1] &&
        p.y <= cam.bbx_max[1] && p.z >= cam.bbx_min[2] && p.z <= cam.bbx_max[2]))
    return false;
  for (int k = 0; k < 6; ++k)
    if (!f.planes[k].correct_side(p)) return false;
  return true;
}

enum Status { kUnknown = 0, kFree = 1, kOccupied = 2 };

// VoxbloxManager::getVoxelStatusByIndex -- voxblox_manager.cpp:14-33.
inline Status voxel_status(const Layer& L, const Key& b, int lx, int ly, int lz) {
  auto it = L.blocks.find(b);
  if (it == L.blocks.end()) return kUnknown;
  const float* v = it->second.get() + 2 * linear_index(L.vps, lx, ly, lz);
  if (v[1] <= 1e-1) return kUnknown;   // float widened to double vs double literal
  if (v[0] <= 0.0) return kOccupied;
  return kFree;
}

struct GainOut {
  double gain;
  double yaw;          // state[3]
  int32_t yaw_deg;
  uint64_t counts[3];  // unknown, free, occupied over the winning window (or all yaws)
  uint64_t steps;      // executed voxel-steps (SURVEY.md 8d unit of work)
  uint64_t rays;
};

// RrtTree::optimized_gain (rrt.cpp:351-479) and RrtTree::gain (rrt.cpp:481-579).
void eval_gain(const Ctx* c, const Layer& L, const double pos[3], bool sum_all, GainOut* out,
               uint32_t* per_yaw /* [360][3] or null */) {
  const Camera& cam = c->cam;
  const double vs = (double)L.voxel_size;  // rrt.cpp:352 (float layer value widened)
  const double vs_inv = 1.0 / vs;
  const double cube = vs * vs * vs;        // CUBE macro, rrt.h:23
  const V3 p{pos[0], pos[1], pos[2]};
  double vertical_gain[360];
  uint32_t cnt[360][3];
  double total_gain = 0.0;  // gain(): one accumulator over all yaws (rrt.cpp:497)
  uint64_t steps = 0, rays = 0;
  YawFrame f;
  for (int i = -180; i < 180; ++i) {
    yaw_frame(c, p, i, &f);
    const V3 ud = sub(f.pp0, f.pp1);                      // :382
    const V3 us = normalized(ud);                          // :383
    const int u_max = (int)std::ceil(norm(ud) * vs_inv);   // :384
    const V3 d21 = sub(f.pp2, f.pp1);
    const V3 vc{d21.x / 2.0, d21.y / 2.0, d21.z / 2.0};    // :385
    double gain = sum_all ? total_gain : 0.0;              // :388 (gain(): never reset, :497)
    uint32_t* ct = cnt[i + 180];
    ct[0] = ct[1] = ct[2] = 0;
    const float start[3] = {(float)(f.cam_pos.x * vs_inv), (float)(f.cam_pos.y * vs_inv),
                            (float)(f.cam_pos.z * vs_inv)};  // :408-409
    for (int u = 0; u < u_max; ++u) {
      // :396  pos = pp1 + u*u_slope*voxel_size + v_center
      const V3 e{(f.pp1.x + ((double)u * us.x) * vs) + vc.x,
                 (f.pp1.y + ((double)u * us.y) * vs) + vc.y,
                 (f.pp1.z + ((double)u * us.z) * vs) + vc.z};
      const float end[3] = {(float)(e.x * vs_inv), (float)(e.y * vs_inv),
                            (float)(e.z * vs_inv)};  // :410-411
      RayCaster rc(start, end);                       // :413-414
      ++rays;
      int64_t g[3];
      while (rc.next(g)) {                            // :419
        ++steps;
        const Key b{block_from_global(g[0], L.vps_inv), block_from_global(g[1], L.vps_inv),
                    block_from_global(g[2], L.vps_inv)};  // :421
        const V3 rp{(double)g[0] * vs, (double)g[1] * vs, (double)g[2] * vs};  // :431
        if (in_view(cam, f, rp)) {                    // :433
          const Status st = voxel_status(L, b, local_from_global(g[0], L.vps),
                                         local_from_global(g[1], L.vps),
                                         local_from_global(g[2], L.vps));  // :435
          if (st == kUnknown) {
            gain += cam.gain_unmapped; ++ct[0];
          } else if (st == kOccupied) {
            gain += cam.gain_occupied; ++ct[2];
            break;
          } else {
            gain += cam.gain_free; ++ct[1];
          }
        }
      }
    }
    vertical_gain[i + 180] = gain;  // :447
    if (sum_all) total_gain = gain;
  }
  out->steps = steps;
  out->rays = rays;
  if (per_yaw) std::memcpy(per_yaw, cnt, sizeof(cnt));
  if (sum_all) {
    // rrt.cpp:481-579: one running double sum over every step of every yaw,
    // state[3] = 0, return gain * cube.
    uint64_t n[3] = {0, 0, 0};
    for (int y = 0; y < 360; ++y)
      for (int k = 0; k < 3; ++k) n[k] += cnt[y][k];
    out->gain = total_gain * cube;
    out->yaw = 0.0;
    out->yaw_deg = 0;
    out->counts[0] = n[0]; out->counts[1] = n[1]; out->counts[2] = n[2];
    return;
  }
  // Best yaw window -- rrt.cpp:449-478.
  double max_gain = 0.0;
  int max_gain_yaw = 0;
  const int half_hfov = (int)std::floor(cam.hfov_deg / 2.0);
  for (int i = -180; i < 180; i += 5) {
    double current = 0.0;
    int left = (i + 180) - half_hfov;
    if (left < 0) {
      for (int j = 360 + left; j < 360; ++j) current += vertical_gain[j];
      left = 0;
    }
    int right = (i + 180) + half_hfov;
    if (right >= 360) {
      for (int j = 0; j < right - 359; ++j) current += vertical_gain[j];
      right = 359;
    }
    for (int j = left; j <= right; ++j) current += vertical_gain[j];
    if (current > max_gain) {
      max_gain = current;
      max_gain_yaw = i;
    }
  }
  out->yaw_deg = max_gain_yaw;
  out->yaw = max_gain_yaw * M_PI / 180.0;  // :477
  out->gain = max_gain * cube;             // :478
  // Counts over the winning window (diagnostic; same wrap rule as above).
  uint64_t n[3] = {0, 0, 0};
  for (int d = -half_hfov; d <= half_hfov; ++d) {
    const int j = (((max_gain_yaw + 180 + d) % 360) + 360) % 360;
    for (int k = 0; k < 3; ++k) n[k] += cnt[j][k];
  }
  out->counts[0] = n[0]; out->counts[1] = n[1]; out->counts[2] = n[2];
}

// HighResManager::checkMotion (voxblox_manager.h:89-108) +
// checkCollisionWithRobotAtVoxel (voxblox_manager.cpp:121-129) +
// Layer::getVoxelPtrByGlobalIndex (layer.h:215-239).
bool check_motion(const Layer& L, double radius, const double seg[6], uint64_t* steps) {
  float s[3], e[3];
  for (int k = 0; k < 3; ++k) {
    s[k] = (float)seg[k] / L.voxel_size;      // float division by the float voxel size
    e[k] = (float)seg[3 + k] / L.voxel_size;
  }
  RayCaster rc(s, e);
  int64_t g[3];
  while (rc.next(g)) {
    if (steps) ++*steps;
    const Key b{block_from_global(g[0], L.vps_inv), block_from_global(g[1], L.vps_inv),
                block_from_global(g[2], L.vps_inv)};
    auto it = L.blocks.find(b);
    if (it == L.blocks.end()) return false;  // nullptr voxel => collision
    const float d = it->second.get()[linear_index(L.vps, local_from_global(g[0], L.vps),
                                                  local_from_global(g[1], L.vps),
                                                  local_from_global(g[2], L.vps))];
    if (radius >= d) return false;           // double >= float widened
  }
  return true;
}

template <class F>
void parallel_for(size_t n, int n_threads, F&& fn) {
  if (n_threads <= 1 || n <= 1) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  const int nt = (int)std::min<size_t>((size_t)n_threads, n);
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&] {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= n) break;
        fn(i);
      }
    });
  for (auto& t : th) t.join();
}

}  // namespace

// ---------------------------------------------------------------------------
// C API (ctypes).
// ---------------------------------------------------------------------------
extern "C" {

void* nbvo_create() { return new Ctx(); }
void nbvo_destroy(void* c) { delete static_cast<Ctx*>(c); }

int nbvo_layer_create(void* c_, int kind, float voxel_size, int vps) {
  Ctx* c = static_cast<Ctx*>(c_);
  if ((vps & (vps - 1)) != 0 || vps <= 0) return -1;
  auto L = std::make_unique<Layer>();
  L->kind = kind;
  L->voxel_size = voxel_size;
  L->vps = vps;
  L->vps_inv = 1.0f / (float)vps;  // layer.h:34-44
  L->nvox = (size_t)vps * vps * vps;
  c->layers.push_back(std::move(L));
  return (int)c->layers.size() - 1;
}

// Planar payload: TSDF [n][nvox][2] float (distance, weight); ESDF [n][nvox] float.
int nbvo_layer_update(void* c_, int layer, size_t n, const int32_t* keys, const float* data) {
  Ctx* c = static_cast<Ctx*>(c_);
  Layer& L = *c->layers.at(layer);
  const size_t per = L.nvox * (L.kind == 0 ? 2 : 1);
  for (size_t i = 0; i < n; ++i) {
    const Key k{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]};
    auto& slot = L.blocks[k];
    if (!slot) slot.reset(new float[per]);
    std::memcpy(slot.get(), data + i * per, per * sizeof(float));
    if (L.kind == 1) {  // planar ESDF payload carries no flags: a voxel is observed unless its distance is NaN
      auto& ob = L.observed[k];
      if (!ob) ob.reset(new uint8_t[L.nvox]);
      for (size_t v = 0; v < L.nvox; ++v) ob[v] = std::isnan(data[i * per + v]) ? 0 : 1;
    }
  }
  return 0;
}
// ESDF blocks with explicit EsdfVoxel::observed flags: dist [n][nvox] floats, obs [n][nvox] bytes
int nbvo_layer_update_esdf_observed(void* c_, int layer, size_t n, const int32_t* keys, const float* dist, const uint8_t* obs) {
  Ctx* c = static_cast<Ctx*>(c_);
  Layer& L = *c->layers.at(layer);
  if (L.kind != 1) return -3;
  for (size_t i = 0; i < n; ++i) {
    const Key k{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]};
    auto& slot = L.blocks[k];
    if (!slot) slot.reset(new float[L.nvox]);
    std::memcpy(slot.get(), dist + i * L.nvox, L.nvox * sizeof(float));
    auto& ob = L.observed[k];
    if (!ob) ob.reset(new uint8_t[L.nvox]);
    for (size_t v = 0; v < L.nvox; ++v) ob[v] = obs[i * L.nvox + v] ? 1 : 0;
  }
  return 0;
}
int nbvo_layer_remove(void* c_, int layer, size_t n, const int32_t* keys) {
  Ctx* c = static_cast<Ctx*>(c_);
  Layer& L = *c->layers.at(layer);
  for (size_t i = 0; i < n; ++i) {
    L.blocks.erase(Key{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]});
    L.observed.erase(Key{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]});
  }
  return 0;
}
int nbvo_layer_clear(void* c_, int layer) {
  static_cast<Ctx*>(c_)->layers.at(layer)->blocks.clear();
  static_cast<Ctx*>(c_)->layers.at(layer)->observed.clear();
  return 0;
}
size_t nbvo_layer_num_blocks(void* c_, int layer) {
  return static_cast<Ctx*>(c_)->layers.at(layer)->blocks.size();
}

int nbvo_set_camera(void* c_, const Camera* cam) {
  Ctx* c = static_cast<Ctx*>(c_);
  c->cam = *cam;
  c->cam_set = true;
  set_intrinsics(c);
  return 0;
}

// Batched gain. out arrays may be null except gain/yaw_deg.
int nbvo_gain_eval(void* c_, int layer, size_t n, const double* pos, double* gain, int32_t* yaw_deg,
                   double* yaw_rad, uint64_t* counts /*[n][3]*/, uint32_t* per_yaw /*[n][360][3]*/,
                   uint64_t* steps /*[n]*/, uint64_t* rays /*[n]*/, int n_threads) {
  Ctx* c = static_cast<Ctx*>(c_);
  if (!c->cam_set) return -2;
  const Layer& L = *c->layers.at(layer);
  if (L.kind != 0) return -3;
  const bool sum_all = c->cam.sum_all_yaws != 0;
  parallel_for(n, n_threads, [&](size_t i) {
    GainOut o{};
    eval_gain(c, L, pos + 3 * i, sum_all, &o, per_yaw ? per_yaw + i * 1080 : nullptr);
    gain[i] = o.gain;
    yaw_deg[i] = o.yaw_deg;
    if (yaw_rad) yaw_rad[i] = o.yaw;
    if (counts) { counts[3 * i] = o.counts[0]; counts[3 * i + 1] = o.counts[1]; counts[3 * i + 2] = o.counts[2]; }
    if (steps) steps[i] = o.steps;
    if (rays) rays[i] = o.rays;
  });
  return 0;
}

int nbvo_check_segments(void* c_, int layer, double radius, size_t n, const double* seg,
                        uint8_t* is_free, uint64_t* steps /*[n] or null*/, int n_threads) {
  Ctx* c = static_cast<Ctx*>(c_);
  const Layer& L = *c->layers.at(layer);
  if (L.kind != 1) return -3;
  parallel_for(n, n_threads, [&](size_t i) {
    uint64_t st = 0;
    is_free[i] = check_motion(L, radius, seg + 6 * i, &st) ? 1 : 0;
    if (steps) steps[i] = st;
  });
  return 0;
}

// ---- ESDF interpolated distance and gradient ------------------------------------------
// EsdfMap::getDistanceAndGradientAtPosition (voxblox/src/core/esdf_map.cc:22-52) -> Interpolator<EsdfVoxel>
// (voxblox/interpolator/interpolator_inl.h): getInterpDistance :318-330 (setIndexes :160-198, getVoxelsAndQVector
// :224-272, getQVector :200-222, interpMember :447-474) and getGradient :47-77.  All in float.  The two products of
// interpMember, q * (table * data^T), are summed over the inner index from left to right -- the evaluation order of
// the Eigen stand-in the pinning build uses (oracle/ref_shim/Eigen/Core); a build against real Eigen may sum them in
// another order, so values are claimed against the real library within float tolerance only, the success flag and the
// choice of the eight voxels exactly.
struct EsdfLookup {
  const Layer& L;
  float block_size, block_size_inv, voxel_size_inv;
  explicit EsdfLookup(const Layer& l) : L(l) {
    block_size = (float)L.vps * L.voxel_size;                  // layer.h:39 / block.h:39
    block_size_inv = (float)(1.0 / (double)block_size);        // layer.h:41
    voxel_size_inv = (float)(1.0 / (double)L.voxel_size);      // layer.h:37, block.h:38
  }
  // Block::computeCoordinatesFromVoxelIndex (block.h:90-92): origin_ + getCenterPointFromGridIndex(index, voxel_size_),
  // the centre being (float(idx) + 0.5) * voxel_size evaluated in double and rounded to float (common.h:188-193)
  float voxel_centre(int block, int voxel) const {
    const float origin = (float)block * block_size;            // common.h:196-201
    const float centre = (float)(((double)(float)voxel + 0.5) * (double)L.voxel_size);
    return origin + centre;
  }
  // Interpolator::getInterpDistance
  bool interp_distance(const float pos[3], float* distance) const {
    int b[3], v[3];
    for (int k = 0; k < 3; ++k) b[k] = (int)grid_index(pos[k], block_size_inv);  // layer.h:128-131
    if (L.blocks.find(Key{b[0], b[1], b[2]}) == L.blocks.end()) return false;     // setIndexes :165-168
    for (int k = 0; k < 3; ++k) {
      const float origin = (float)b[k] * block_size;
      const int64_t t = grid_index(pos[k] - origin, voxel_size_inv);              // block_inl.h:30-40
      v[k] = (int)std::max<int64_t>(std::min<int64_t>(t, L.vps - 1), 0);
      const float centre_offset = pos[k] - voxel_centre(b[k], v[k]);              // :171-172
      if (centre_offset < 0) {                                                    // :176-184
        v[k]--;
        if (v[k] < 0) {
          b[k]--;
          v[k] += L.vps;
        }
      }
    }
    float q[8], data[8];
    for (int i = 0; i < 8; ++i) {                                                 // getVoxelsAndQVector :232-270
      if (L.blocks.find(Key{b[0], b[1], b[2]}) == L.blocks.end()) return false;   // the (shifted) home block itself
      const int off[3] = {(i >> 2) & 1, (i >> 1) & 1, i & 1};                     // columns of the index table :190-195
      int nb[3], nv[3];
      for (int k = 0; k < 3; ++k) {
        nb[k] = b[k];
        nv[k] = v[k] + off[k];
        if (nv[k] >= L.vps) {                                                     // :241-249
          nb[k]++;
          nv[k] -= L.vps;
        }
      }
      auto it = L.blocks.find(Key{nb[0], nb[1], nb[2]});
      if (it == L.blocks.end()) return false;
      if (i == 0) {                                                               // getQVector :205-221
        float o[3];
        for (int k = 0; k < 3; ++k) o[k] = (pos[k] - voxel_centre(nb[k], nv[k])) * voxel_size_inv;
        q[0] = 1.0f; q[1] = o[0]; q[2] = o[1]; q[3] = o[2];
        q[4] = o[0] * o[1]; q[5] = o[1] * o[2]; q[6] = o[2] * o[0]; q[7] = o[0] * o[1] * o[2];
      }
      const size_t lin = (size_t)linear_index(L.vps, nv[0], nv[1], nv[2]);
      data[i] = it->second.get()[lin];
      auto ob = L.observed.find(Key{nb[0], nb[1], nb[2]});
      if (ob == L.observed.end() || !ob->second.get()[lin]) return false;         // utils::isObservedVoxel :266-268
    }
    static const float table[8][8] = {{1, 0, 0, 0, 0, 0, 0, 0},  {-1, 0, 0, 0, 1, 0, 0, 0}, {-1, 0, 1, 0, 0, 0, 0, 0}, {-1, 1, 0, 0, 0, 0, 0, 0},
                                      {1, 0, -1, 0, -1, 0, 1, 0}, {1, -1, -1, 1, 0, 0, 0, 0}, {1, -1, 0, 0, -1, 1, 0, 0},
                                      {-1, 1, 1, -1, 1, -1, -1, 1}};                // interpMember :460-471
    float y[8];
    for (int r = 0; r < 8; ++r) {
      float acc = table[r][0] * data[0];
      for (int c = 1; c < 8; ++c) acc = acc + table[r][c] * data[c];
      y[r] = acc;
    }
    float acc = q[0] * y[0];
    for (int i = 1; i < 8; ++i) acc = acc + q[i] * y[i];
    *distance = acc;
    return true;
  }
  // Interpolator::getGradient(pos, grad, interpolate = true) :47-77: grad is left partially summed on an early failure
  bool gradient(const float pos[3], float grad[3]) const {
    int b[3];
    for (int k = 0; k < 3; ++k) b[k] = (int)grid_index(pos[k], block_size_inv);
    if (L.blocks.find(Key{b[0], b[1], b[2]}) == L.blocks.end()) return false;
    grad[0] = grad[1] = grad[2] = 0.0f;
    for (int i = 0; i < 3; ++i)
      for (int sign = -1; sign <= 1; sign += 2) {
        float p[3] = {pos[0] + 0.0f, pos[1] + 0.0f, pos[2] + 0.0f};                // pos + offset: all three components are added
        p[i] = pos[i] + (float)sign * L.voxel_size;
        float d;
        if (!interp_distance(p, &d)) return false;
        grad[i] += d * (float)sign;
      }
    const float denom = 2 * L.voxel_size;
    for (int k = 0; k < 3; ++k) grad[k] = grad[k] / denom;
    return true;
  }
};

// one query of EsdfMap::getDistanceAndGradientAtPosition(position, &distance, &gradient), esdf_map.cc:31-52
inline bool distance_and_gradient(const Layer& L, const double position[3], double* distance, double gradient[3]) {
  const EsdfLookup E(L);
  const float pos[3] = {(float)position[0], (float)position[1], (float)position[2]};
  float d = 0.0f, g[3] = {0.0f, 0.0f, 0.0f};
  bool success = E.interp_distance(pos, &d);
  success &= E.gradient(pos, g);  // evaluated whatever the first call returned
  *distance = (double)d;
  for (int k = 0; k < 3; ++k) gradient[k] = (double)g[k];
  return success;
}

// LowResManager::getVoxelStatus(position) -- voxblox_manager.cpp:73-90: the position is cast to float, the
// block comes from Layer::computeBlockIndexFromCoordinates (layer.h:128-131: getGridIndexFromPoint with
// block_size_inv), the voxel from Block::computeTruncatedVoxelIndexFromCoordinates (block_inl.h:30-40:
// getGridIndexFromPoint(coords - origin, voxel_size_inv) clamped to the block), origin = float(index) *
// block_size (common.h:196-201, layer.h:137).  All in float, as the reference.
inline Status voxel_status_at(const Layer& L, const V3& position) {
  const float block_size = (float)L.vps * L.voxel_size;  // layer.h:32  block_size_ = voxel_size_ * voxels_per_side_
  // layer.h:37,41 / block.h:38,40: `1.0 / x` -- a double division rounded to float on assignment (identical to
  // the float quotient: 53 >= 2 * 24 + 2 bits makes the double rounding innocuous; written the reference's way)
  const float block_size_inv = (float)(1.0 / (double)block_size);
  const float voxel_size_inv = (float)(1.0 / (double)L.voxel_size);
  const float pf[3] = {(float)position.x, (float)position.y, (float)position.z};
  int b[3], l[3];
  for (int k = 0; k < 3; ++k) {
    b[k] = (int)grid_index(pf[k], block_size_inv);
    const float origin = (float)b[k] * block_size;
    const int64_t v = grid_index(pf[k] - origin, voxel_size_inv);
    l[k] = (int)std::max<int64_t>(std::min<int64_t>(v, L.vps - 1), 0);
  }
  return voxel_status(L, Key{b[0], b[1], b[2]}, l[0], l[1], l[2]);
}

// History::bfs -- history.cpp:96-137, literally: cells are identified by std::to_string (= "%f", six
// decimals) of their ACCUMULATED double coordinates, kept in FIFO order; a neighbour that is not yet
// visited, within max_distance (6.0 in the reference) of the start and inside the exploration bounds is
// looked up by coordinates: free -> visited + queued, unknown -> counted (once per encounter, i.e. per
// free-cell face), occupied -> nothing.  The start point is expanded whatever its own status.
inline std::string bfs_key(const V3& v) {
  char buf[3 * 400];  // "%f" of a double has at most 1 + 309 + 1 + 6 characters
  std::snprintf(buf, sizeof buf, "%f/%f/%f", v.x, v.y, v.z);
  return std::string(buf);
}
// Order-independent fingerprint of a visited cell: a 64-bit mix of the bit patterns of its three doubles (the
// same function in csrc/frontier_kernels.cuh); the sum over all visited cells tells whether two
// implementations hold the same doubles for every cell, i.e. made the same first arrivals.
inline uint64_t cell_fingerprint(const V3& v) {
  uint64_t b[3];
  std::memcpy(&b[0], &v.x, 8); std::memcpy(&b[1], &v.y, 8); std::memcpy(&b[2], &v.z, 8);
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (int k = 0; k < 3; ++k) {
    h ^= b[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
  }
  return h;
}
inline uint64_t frontier_bfs(const Ctx* c, const Layer& L, const V3& point, double max_distance, uint64_t* visited_out,
                             uint64_t* checksum_out = nullptr) {
  const double vs = (double)L.voxel_size;  // getResolution(): the float voxel size widened (voxblox_manager.h:35)
  const V3 conn[6] = {{vs, 0, 0}, {-vs, 0, 0}, {0, vs, 0}, {0, -vs, 0}, {0, 0, vs}, {0, 0, -vs}};
  std::unordered_set<std::string> visited;
  visited.insert(bfs_key(point));
  std::queue<V3> queue;
  queue.push(point);
  uint64_t frontier = 0, checksum = cell_fingerprint(point);
  while (!queue.empty()) {
    const V3 cur = queue.front();
    queue.pop();
    for (const V3& d : conn) {
      const V3 cand = add(cur, d);
      if (visited.find(bfs_key(cand)) != visited.end()) continue;
      if (!(norm(sub(cand, point)) <= max_distance)) continue;
      if (!(cand.x >= c->cam.bbx_min[0] && cand.x <= c->cam.bbx_max[0] && cand.y >= c->cam.bbx_min[1] &&
            cand.y <= c->cam.bbx_max[1] && cand.z >= c->cam.bbx_min[2] && cand.z <= c->cam.bbx_max[2]))
        continue;  // nbveplanner/common.h:22-31
      const Status st = voxel_status_at(L, cand);
      if (st == kFree) {
        visited.insert(bfs_key(cand));
        queue.push(cand);
        checksum += cell_fingerprint(cand);
      } else if (st == kUnknown) {
        ++frontier;
      }
    }
  }
  if (visited_out) *visited_out = visited.size();
  if (checksum_out) *checksum_out = checksum;
  return frontier;
}

// History::recalculatePotential's input: potential_gain = frontierVoxels * pow(voxel_size, 3.0) per vertex.
int nbvo_frontier_potential(void* c_, int layer, size_t n, const double* points, double max_distance, double* gain,
                            uint64_t* frontier_voxels /*[n] or null*/, uint64_t* visited /*[n] or null*/,
                            uint64_t* checksum /*[n] or null*/, int n_threads) {
  Ctx* c = static_cast<Ctx*>(c_);
  if (!c->cam_set) return -2;
  const Layer& L = *c->layers.at(layer);
  if (L.kind != 0) return -3;
  const double cube = std::pow((double)L.voxel_size, 3.0);
  parallel_for(n, n_threads, [&](size_t i) {
    uint64_t vis = 0, sum = 0;
    const uint64_t f = frontier_bfs(c, L, V3{points[3 * i], points[3 * i + 1], points[3 * i + 2]}, max_distance, &vis, &sum);
    if (checksum) checksum[i] = sum;
    gain[i] = (double)(int)f * cube;  // int frontierVoxels * pow(...) (history.cpp:136)
    if (frontier_voxels) frontier_voxels[i] = f;
    if (visited) visited[i] = vis;
  });
  return 0;
}
int nbvo_distance_and_gradient(void* c_, int layer, size_t n, const double* positions, uint8_t* ok, double* distance, double* gradient,
                               int n_threads) {
  Ctx* c = static_cast<Ctx*>(c_);
  const Layer& L = *c->layers.at(layer);
  if (L.kind != 1) return -3;
  parallel_for(n, n_threads, [&](size_t i) { ok[i] = distance_and_gradient(L, positions + 3 * i, distance + i, gradient + 3 * i) ? 1 : 0; });
  return 0;
}
int nbvo_voxel_status_at(void* c_, int layer, const double* p) {
  Ctx* c = static_cast<Ctx*>(c_);
  return voxel_status_at(*c->layers.at(layer), V3{p[0], p[1], p[2]});
}

// Batched restatement of the body of RrtTree::iterate (rrt.cpp:172-246) for given uniform triples:
// status 0 = sample rejected (outside the vicinity sphere / exploration box), 1 = edge collides,
// 2 = node created.  Parents are the n_nodes existing nodes only (kd_nearest3 by brute force, squared
// distance accumulated x, y, z as kdtree.c does; lowest index on ties).
struct ExpandParams {
  double root[3];
  double vicinity, extension_range, overshoot, robot_radius, v_max, dyaw_max, zero_gain;
};
int nbvo_expand_tree(void* c_, int tsdf_layer, int esdf_layer, const ExpandParams* P, size_t n_nodes, const double* node_state,
                     const double* node_distance, size_t n, const double* uniforms, uint8_t* status, double* pos,
                     int32_t* parent, double* gain, double* yaw, double* distance, double* score, int64_t* best_index,
                     int n_threads) {
  Ctx* c = static_cast<Ctx*>(c_);
  if (!c->cam_set || n_nodes == 0) return -2;
  const Layer& LT = *c->layers.at(tsdf_layer);
  const Layer& LE = *c->layers.at(esdf_layer);
  const Camera& cam = c->cam;
  const bool sum_all = cam.sum_all_yaws != 0;
  parallel_for(n, n_threads, [&](size_t i) {
    status[i] = 0; parent[i] = -1; gain[i] = 0.0; yaw[i] = 0.0; distance[i] = 0.0; score[i] = 0.0;
    double s[3];
    for (int k = 0; k < 3; ++k) s[k] = 2.0 * P->vicinity * (uniforms[3 * i + k] - 0.5);          // :176-178
    pos[3 * i] = pos[3 * i + 1] = pos[3 * i + 2] = 0.0;
    if (s[0] * s[0] + s[1] * s[1] + s[2] * s[2] > P->vicinity * P->vicinity) return;           // :179-180
    for (int k = 0; k < 3; ++k) s[k] += P->root[k];                                             // :182
    if (!(s[0] >= cam.bbx_min[0] && s[0] <= cam.bbx_max[0] && s[1] >= cam.bbx_min[1] && s[1] <= cam.bbx_max[1] &&
          s[2] >= cam.bbx_min[2] && s[2] <= cam.bbx_max[2])) return;                             // :183-185
    size_t best = 0;                                                                            // :190-197
    double best_d = 0.0;
    for (size_t j = 0; j < n_nodes; ++j) {
      double d = 0.0;
      for (int k = 0; k < 3; ++k) { const double df = node_state[4 * j + k] - s[k]; d += df * df; }
      if (j == 0 || d < best_d) { best_d = d; best = j; }
    }
    const V3 origin{node_state[4 * best], node_state[4 * best + 1], node_state[4 * best + 2]};
    V3 dir{s[0] - origin.x, s[1] - origin.y, s[2] - origin.z};                                   // :202-203
    if (norm(dir) > P->extension_range) dir = scale(P->extension_range, normalized(dir));       // :204-206
    const V3 ns = add(origin, dir);                                                             // :207-209
    const V3 un = normalized(dir);
    const V3 end = add(add(dir, origin), V3{un.x * P->overshoot, un.y * P->overshoot, un.z * P->overshoot});  // :210-212
    pos[3 * i] = ns.x; pos[3 * i + 1] = ns.y; pos[3 * i + 2] = ns.z;
    parent[i] = (int32_t)best;
    const double seg[6] = {origin.x, origin.y, origin.z, end.x, end.y, end.z};
    status[i] = 1;
    if (!check_motion(LE, P->robot_radius, seg, nullptr)) return;
    status[i] = 2;
    GainOut o{};
    const double p3[3] = {ns.x, ns.y, ns.z};
    eval_gain(c, LT, p3, sum_all, &o, nullptr);                                                 // :214-219
    gain[i] = o.gain;
    yaw[i] = o.yaw;
    distance[i] = node_distance[best] + norm(dir);                                              // :223
    double dy = o.yaw - node_state[4 * best + 3];                                               // :225
    if (dy > M_PI) dy -= 2.0 * M_PI;
    if (dy < -M_PI) dy += 2.0 * M_PI;
    const double ttr = std::max(dy / P->dyaw_max, distance[i] / P->v_max);                      // :232-233
    score[i] = o.gain / ttr;                                                                    // :234
  });
  double bg = P->zero_gain;                                                                     // :243-246, in sample order
  int64_t bi = -1;
  for (size_t i = 0; i < n; ++i)
    if (status[i] == 2 && score[i] > bg) { bg = score[i]; bi = (int64_t)i; }
  *best_index = bi;
  return 0;
}

// The same stream of triples with the reference's sequential semantics: every RrtTree::iterate() call (rrt.cpp:163-264)
// sees the nodes the earlier calls inserted (kd_insert3 at :236), as nearest-parent candidates and -- through the parent's
// chosen yaw, state_[3] -- in the score (:225).  parent[i] indexes the given nodes first, then the created nodes in creation
// order.  Gains do not depend on other nodes, so they are evaluated in parallel between the two sequential passes.
int nbvo_expand_tree_seq(void* c_, int tsdf_layer, int esdf_layer, const ExpandParams* P, size_t n_nodes, const double* node_state,
                         const double* node_distance, size_t n, const double* uniforms, uint8_t* status, double* pos,
                         int32_t* parent, double* gain, double* yaw, double* distance, double* score, int64_t* best_index,
                         int n_threads) {
  Ctx* c = static_cast<Ctx*>(c_);
  if (!c->cam_set || n_nodes == 0) return -2;
  const Layer& LT = *c->layers.at(tsdf_layer);
  const Layer& LE = *c->layers.at(esdf_layer);
  const Camera& cam = c->cam;
  const bool sum_all = cam.sum_all_yaws != 0;
  std::vector<double> nx(node_state, node_state + 4 * n_nodes), nd(node_distance, node_distance + n_nodes);
  std::vector<size_t> created;  // triples that created a node, in order
  for (size_t i = 0; i < n; ++i) {
    status[i] = 0; parent[i] = -1; gain[i] = 0.0; yaw[i] = 0.0; distance[i] = 0.0; score[i] = 0.0;
    pos[3 * i] = pos[3 * i + 1] = pos[3 * i + 2] = 0.0;
    double s[3];
    for (int k = 0; k < 3; ++k) s[k] = 2.0 * P->vicinity * (uniforms[3 * i + k] - 0.5);          // :176-178
    if (s[0] * s[0] + s[1] * s[1] + s[2] * s[2] > P->vicinity * P->vicinity) continue;          // :179-180
    for (int k = 0; k < 3; ++k) s[k] += P->root[k];                                             // :182
    if (!(s[0] >= cam.bbx_min[0] && s[0] <= cam.bbx_max[0] && s[1] >= cam.bbx_min[1] && s[1] <= cam.bbx_max[1] &&
          s[2] >= cam.bbx_min[2] && s[2] <= cam.bbx_max[2])) continue;                           // :183-185
    const size_t total = nd.size();
    size_t best = 0;                                                                            // :190-197
    double best_d = 0.0;
    for (size_t j = 0; j < total; ++j) {
      double d = 0.0;
      for (int k = 0; k < 3; ++k) { const double df = nx[4 * j + k] - s[k]; d += df * df; }
      if (j == 0 || d < best_d) { best_d = d; best = j; }
    }
    const V3 origin{nx[4 * best], nx[4 * best + 1], nx[4 * best + 2]};
    V3 dir{s[0] - origin.x, s[1] - origin.y, s[2] - origin.z};                                   // :202-203
    if (norm(dir) > P->extension_range) dir = scale(P->extension_range, normalized(dir));       // :204-206
    const V3 ns = add(origin, dir);                                                             // :207-209
    const V3 un = normalized(dir);
    const V3 end = add(add(dir, origin), V3{un.x * P->overshoot, un.y * P->overshoot, un.z * P->overshoot});  // :210-212
    pos[3 * i] = ns.x; pos[3 * i + 1] = ns.y; pos[3 * i + 2] = ns.z;
    parent[i] = (int32_t)best;
    const double seg[6] = {origin.x, origin.y, origin.z, end.x, end.y, end.z};
    status[i] = 1;
    if (!check_motion(LE, P->robot_radius, seg, nullptr)) continue;
    status[i] = 2;
    distance[i] = nd[best] + norm(dir);                                                         // :223
    nx.insert(nx.end(), {ns.x, ns.y, ns.z, 0.0});                                               // :236 (yaw filled in below)
    nd.push_back(distance[i]);
    created.push_back(i);
  }
  parallel_for(created.size(), n_threads, [&](size_t k) {
    const size_t i = created[k];
    GainOut o{};
    const double p3[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    eval_gain(c, LT, p3, sum_all, &o, nullptr);                                                 // :214-219
    gain[i] = o.gain;
    yaw[i] = o.yaw;
  });
  double bg = P->zero_gain;                                                                     // :243-246
  int64_t bi = -1;
  for (size_t k = 0; k < created.size(); ++k) {
    const size_t i = created[k];
    nx[4 * (n_nodes + k) + 3] = yaw[i];
    double dy = yaw[i] - nx[4 * (size_t)parent[i] + 3];                                         // :225
    if (dy > M_PI) dy -= 2.0 * M_PI;
    if (dy < -M_PI) dy += 2.0 * M_PI;
    const double ttr = std::max(dy / P->dyaw_max, distance[i] / P->v_max);                      // :232-233
    score[i] = gain[i] / ttr;                                                                   // :234
    if (score[i] > bg) { bg = score[i]; bi = (int64_t)i; }
  }
  *best_index = bi;
  return 0;
}

// castRay (integrator_utils.h:133-143): returns the number of indices the
// reference would emit; writes at most `cap` of them.
size_t nbvo_cast_ray(const float* start, const float* end, int64_t* out, size_t cap) {
  RayCaster rc(start, end);
  int64_t g[3];
  size_t n = 0;
  while (rc.next(g)) {
    if (n < cap) { out[3 * n] = g[0]; out[3 * n + 1] = g[1]; out[3 * n + 2] = g[2]; }
    ++n;
  }
  return n;
}

// Per-yaw frame dump for tests: planes [6][4] (nx,ny,nz,d), cam[3], pp[3][3], u_max.
int nbvo_yaw_frame(void* c_, int layer, const double* pos, int yaw_deg, double* planes, double* cam,
                   double* pp, int32_t* u_max) {
  Ctx* c = static_cast<Ctx*>(c_);
  const Layer& L = *c->layers.at(layer);
  YawFrame f;
  yaw_frame(c, V3{pos[0], pos[1], pos[2]}, yaw_deg, &f);
  for (int k = 0; k < 6; ++k) {
    planes[4 * k] = f.planes[k].n.x; planes[4 * k + 1] = f.planes[k].n.y;
    planes[4 * k + 2] = f.planes[k].n.z; planes[4 * k + 3] = f.planes[k].d;
  }
  cam[0] = f.cam_pos.x; cam[1] = f.cam_pos.y; cam[2] = f.cam_pos.z;
  const V3* P[3] = {&f.pp0, &f.pp1, &f.pp2};
  for (int k = 0; k < 3; ++k) { pp[3 * k] = P[k]->x; pp[3 * k + 1] = P[k]->y; pp[3 * k + 2] = P[k]->z; }
  const double vs_inv = 1.0 / (double)L.voxel_size;
  *u_max = (int)std::ceil(norm(sub(f.pp0, f.pp1)) * vs_inv);
  return 0;
}

// Index helpers for the test_tsdf_map.cc known answers.
void nbvo_grid_index_from_point(const float* p, float grid_size_inv, int64_t* out) {
  for (int k = 0; k < 3; ++k) out[k] = grid_index(p[k], grid_size_inv);
}
void nbvo_block_and_local(const int64_t* g, int vps, int32_t* block, int32_t* local) {
  const float inv = 1.0f / (float)vps;
  for (int k = 0; k < 3; ++k) { block[k] = block_from_global(g[k], inv); local[k] = local_from_global(g[k], vps); }
}
uint64_t nbvo_linear_index(int vps, int x, int y, int z) { return linear_index(vps, x, y, z); }
uint64_t nbvo_block_hash(int x, int y, int z) { return KeyHash()(Key{x, y, z}); }
// Voxel status of one global voxel index (0 unknown, 1 free, 2 occupied).
int nbvo_voxel_status(void* c_, int layer, const int64_t* g) {
  const Layer& L = *static_cast<Ctx*>(c_)->layers.at(layer);
  const Key b{block_from_global(g[0], L.vps_inv), block_from_global(g[1], L.vps_inv),
              block_from_global(g[2], L.vps_inv)};
  return voxel_status(L, b, local_from_global(g[0], L.vps), local_from_global(g[1], L.vps),
                      local_from_global(g[2], L.vps));
}
// ESDF voxel distance or NaN when the block is not allocated.
float nbvo_esdf_distance(void* c_, int layer, const int64_t* g) {
  const Layer& L = *static_cast<Ctx*>(c_)->layers.at(layer);
  const Key b{block_from_global(g[0], L.vps_inv), block_from_global(g[1], L.vps_inv),
              block_from_global(g[2], L.vps_inv)};
  auto it = L.blocks.find(b);
  if (it == L.blocks.end()) return NAN;
  return it->second.get()[linear_index(L.vps, local_from_global(g[0], L.vps),
                                       local_from_global(g[1], L.vps), local_from_global(g[2], L.vps))];
}

}  // extern "C"
