// synth_gen.cc -- synthetic TSDF / ESDF block generator (bench + test input data).
//
// Produces the *inputs* of the hot path: analytic signed-distance worlds sampled at
// voxel centres, following the recipe of the reference's SimulationWorld
// (voxblox/voxblox/include/voxblox/simulation/simulation_world_inl.h:13-86:
// min over objects, clamp to +-max_dist, weight 1 where generated) with the object
// distance functions of voxblox/simulation/objects.h (Sphere :50-62, Cube :113-142,
// PlaneObject :219-227, Cylinder :271-298) restated in float.  Voxels outside the
// observed region (union of spheres around a trajectory) keep weight 0 / ESDF 0,
// blocks are allocated by the caller.  No part of the gain / collision algorithm
// lives here; both the CUDA path and the oracle consume the same generated arrays.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace {

struct Obj {
  int type;  // 0 sphere, 1 cube, 2 plane, 3 cylinder
  float c[3];
  float a[3];
};

inline float norm3(float x, float y, float z) { return std::sqrt(x * x + y * y + z * z); }

inline float obj_distance(const Obj& o, float px, float py, float pz) {
  switch (o.type) {
    case 0:
      return norm3(o.c[0] - px, o.c[1] - py, o.c[2] - pz) - o.a[0];
    case 1: {
      const float p[3] = {px, py, pz};
      float dv[3], iv[3];
      for (int k = 0; k < 3; ++k) {
        const float lo = o.c[k] - o.a[k] / 2.0f - p[k];
        const float hi = p[k] - o.c[k] - o.a[k] / 2.0f;
        dv[k] = std::max(std::max(lo, 0.0f), hi);
        iv[k] = std::max(lo, hi);
      }
      float d = norm3(dv[0], dv[1], dv[2]);
      if (d < 1e-6f) d = std::max(iv[0], std::max(iv[1], iv[2]));
      return d;
    }
    case 2: {
      const float d = -(o.a[0] * o.c[0] + o.a[1] * o.c[1] + o.a[2] * o.c[2]);
      const float p = d / norm3(o.a[0], o.a[1], o.a[2]);
      return (o.a[0] * px + o.a[1] * py + o.a[2] * pz) + p;
    }
    default: {
      const float r = o.a[0], h = o.a[1];
      const float zmin = o.c[2] - h / 2.0f, zmax = o.c[2] + h / 2.0f;
      const float dx = px - o.c[0], dy = py - o.c[1];
      if (pz >= zmin && pz <= zmax) return std::sqrt(dx * dx + dy * dy) - r;
      const float radial = std::max(dx * dx + dy * dy - r * r, 0.0f);
      const float dz = pz > zmax ? pz - zmax : pz - zmin;
      return std::sqrt(radial + dz * dz);
    }
  }
}

// Lower bound of the object distance over an axis-aligned box (for culling).
inline float obj_distance_lower_bound(const Obj& o, const float lo[3], const float hi[3]) {
  const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
  const float half_diag = 0.5f * norm3(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]);
  return obj_distance(o, cx, cy, cz) - half_diag;  // all objects are 1-Lipschitz
}

}  // namespace

extern "C" int nbvsynth_generate(float voxel_size, int vps, size_t n_blocks, const int32_t* keys,
                                 size_t n_obj, const double* objects /*[n][7]*/, size_t n_traj,
                                 const double* traj /*[n][3]*/, double obs_radius,
                                 const double* bounds_min, const double* bounds_max, float tsdf_trunc,
                                 float esdf_max, float* tsdf_out, float* esdf_out, int n_threads) {
  std::vector<Obj> objs(n_obj);
  for (size_t i = 0; i < n_obj; ++i) {
    objs[i].type = (int)objects[7 * i];
    for (int k = 0; k < 3; ++k) {
      objs[i].c[k] = (float)objects[7 * i + 1 + k];
      objs[i].a[k] = (float)objects[7 * i + 4 + k];
    }
  }
  const size_t nvox = (size_t)vps * vps * vps;
  const float block_size = voxel_size * (float)vps;
  const float far_cut = std::max(tsdf_trunc, esdf_max);
  const float R = (float)obs_radius;
  std::atomic<size_t> next{0};
  auto work = [&]() {
    std::vector<int> near_objs, near_traj;
    for (;;) {
      const size_t b = next.fetch_add(1);
      if (b >= n_blocks) break;
      float lo[3], hi[3];
      for (int k = 0; k < 3; ++k) {
        lo[k] = (float)keys[3 * b + k] * block_size;
        hi[k] = lo[k] + block_size;
      }
      near_objs.clear();
      for (size_t i = 0; i < n_obj; ++i)
        if (obj_distance_lower_bound(objs[i], lo, hi) <= far_cut) near_objs.push_back((int)i);
      near_traj.clear();
      bool all_observed = false;
      const float half_diag = 0.5f * norm3(block_size, block_size, block_size);
      for (size_t t = 0; t < n_traj; ++t) {
        const float d = norm3((float)traj[3 * t] - 0.5f * (lo[0] + hi[0]), (float)traj[3 * t + 1] - 0.5f * (lo[1] + hi[1]),
                              (float)traj[3 * t + 2] - 0.5f * (lo[2] + hi[2]));
        if (d + half_diag <= R) all_observed = true;
        if (d - half_diag <= R) near_traj.push_back((int)t);
      }
      float* tsdf = tsdf_out ? tsdf_out + b * nvox * 2 : nullptr;
      float* esdf = esdf_out ? esdf_out + b * nvox : nullptr;
      size_t lin = 0;
      for (int z = 0; z < vps; ++z)
        for (int y = 0; y < vps; ++y)
          for (int x = 0; x < vps; ++x, ++lin) {
            // Voxel centre: block origin + (idx + 0.5) * voxel_size (block.h:84-92).
            const float px = lo[0] + ((float)x + 0.5f) * voxel_size;
            const float py = lo[1] + ((float)y + 0.5f) * voxel_size;
            const float pz = lo[2] + ((float)z + 0.5f) * voxel_size;
            bool obs = px >= bounds_min[0] && px <= bounds_max[0] && py >= bounds_min[1] &&
                       py <= bounds_max[1] && pz >= bounds_min[2] && pz <= bounds_max[2];
            if (obs && !all_observed) {
              obs = false;
              for (int t : near_traj) {
                const float dx = px - (float)traj[3 * t], dy = py - (float)traj[3 * t + 1],
                            dz = pz - (float)traj[3 * t + 2];
                if (dx * dx + dy * dy + dz * dz <= R * R) { obs = true; break; }
              }
            }
            if (!obs) {
              if (tsdf) { tsdf[2 * lin] = 0.0f; tsdf[2 * lin + 1] = 0.0f; }
              if (esdf) esdf[lin] = 0.0f;
              continue;
            }
            float d = far_cut;
            for (int i : near_objs) d = std::min(d, obj_distance(objs[i], px, py, pz));
            if (tsdf) {
              tsdf[2 * lin] = std::max(std::min(d, tsdf_trunc), -tsdf_trunc);
              tsdf[2 * lin + 1] = 1.0f;
            }
            if (esdf) esdf[lin] = std::max(std::min(d, esdf_max), -esdf_max);
          }
    }
  };
  const int nt = std::max(1, n_threads);
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return 0;
}
