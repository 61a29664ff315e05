// The drop-in as a C++ caller sees it: include/nbvgpu_shim.hpp driven the way RrtTree::iterate drives the reference
// (rrt.cpp:172-246) -- one checkMotion and one optimized_gain per sample, sequentially -- on a map uploaded in the
// reference's own block serialisation (Block::serializeToIntegers).  Reads one input file written by
// tests/test_gpu_cpp_dropin.py, writes gain / yaw / free per sample, prints the per-call latencies as JSON.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "nbvgpu_shim.hpp"

struct Pose {
  double v[4];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct Point {
  double v[3];
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double& operator[](int i) { return v[i]; }
};

template <class T>
static bool rd(std::FILE* f, T* p, size_t n) { return std::fread(p, sizeof(T), n, f) == n; }

static double median(std::vector<double> v) {
  std::sort(v.begin(), v.end());
  return v.empty() ? 0.0 : v[v.size() / 2];
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t hdr[4];  // vps, blocks, samples, repetitions
  float voxel_size;
  double cam[21];  // hfov vfov near far | q wxyz | t | bbx_min | bbx_max | gains free occupied unmapped | robot radius
  if (!rd(f, hdr, 4) || !rd(f, &voxel_size, 1) || !rd(f, cam, 21)) return 2;
  const int vps = hdr[0], nb = hdr[1], ns = hdr[2], reps = hdr[3];
  const size_t nvox = (size_t)vps * vps * vps;
  std::vector<int32_t> keys((size_t)nb * 3);
  std::vector<uint32_t> tsdf((size_t)nb * nvox * 3), esdf((size_t)nb * nvox * 2);
  std::vector<double> samples((size_t)ns * 9);  // position, segment start, segment end
  if (!rd(f, keys.data(), keys.size()) || !rd(f, tsdf.data(), tsdf.size()) || !rd(f, esdf.data(), esdf.size()) ||
      !rd(f, samples.data(), samples.size()))
    return 2;
  std::fclose(f);
  try {
    nbvgpu_shim::GpuMap map(0, voxel_size, vps, (size_t)nb + 8, voxel_size, vps, (size_t)nb + 8);
    map.uploadSerialized(map.tsdf_layer(), (size_t)nb, keys.data(), tsdf.data());
    map.uploadSerialized(map.esdf_layer(), (size_t)nb, keys.data(), esdf.data());
    nbvgpu_shim::setUpCamera(map, cam[0], cam[1], cam[2], cam[3], cam + 4, cam + 8, cam + 11, cam + 14, cam[17], cam[18], cam[19]);
    const double radius = cam[20];
    std::vector<double> out((size_t)ns * 3), t_motion, t_gain;
    using clk = std::chrono::steady_clock;
    for (int rep = 0; rep < reps + 1; ++rep) {  // the first pass warms up
      for (int i = 0; i < ns; ++i) {
        const double* s = &samples[(size_t)i * 9];
        const auto t0 = clk::now();
        const bool free_seg = nbvgpu_shim::checkMotion(map, radius, Point{{s[3], s[4], s[5]}}, Point{{s[6], s[7], s[8]}});
        const auto t1 = clk::now();
        Pose p{{s[0], s[1], s[2], 0.0}};
        const double g = nbvgpu_shim::optimized_gain(map, p);
        const auto t2 = clk::now();
        out[(size_t)i * 3] = g;
        out[(size_t)i * 3 + 1] = p[3];
        out[(size_t)i * 3 + 2] = free_seg ? 1.0 : 0.0;
        if (rep > 0) {
          t_motion.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
          t_gain.push_back(std::chrono::duration<double, std::micro>(t2 - t1).count());
        }
      }
    }
    std::FILE* o = std::fopen(argv[2], "wb");
    if (!o || std::fwrite(out.data(), sizeof(double), out.size(), o) != out.size()) return 2;
    std::fclose(o);
    std::printf("{\"samples\": %d, \"repetitions\": %d, \"check_motion_us\": %.3f, \"optimized_gain_us\": %.3f}\n", ns, reps, median(t_motion),
                median(t_gain));
    return 0;
  } catch (const std::exception& e) {
    std::printf("error: %s\n", e.what());
    return 42;
  }
}
