// One C++ process, several GPUs, through include/nbvgpu_shim.hpp: the map is uploaded once in the reference's block
// serialisation (replicated inside nbvgpu_commit), a batch of samples goes through the batched form of RrtTree::iterate's body
// (nbvgpu_shim::evaluateCandidates), and a few single calls (optimized_gain / checkMotion) follow on the same handle.
// Input file as tests/cpp/shim_iterate_loop.cc; argv[3] = number of GPUs (0 = all visible).  Writes per sample
// gain / yaw_deg / free, then the best record (index, score, gain, yaw); prints timings as JSON.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime_api.h>

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
};

template <class T>
static bool rd(std::FILE* f, T* p, size_t n) { return std::fread(p, sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t hdr[4];
  float voxel_size;
  double cam[21];
  if (!rd(f, hdr, 4) || !rd(f, &voxel_size, 1) || !rd(f, cam, 21)) return 2;
  const int vps = hdr[0], nb = hdr[1], ns = hdr[2], reps = hdr[3];
  const size_t nvox = (size_t)vps * vps * vps;
  std::vector<int32_t> keys((size_t)nb * 3);
  std::vector<uint32_t> tsdf((size_t)nb * nvox * 3), esdf((size_t)nb * nvox * 2);
  std::vector<double> samples((size_t)ns * 9);
  if (!rd(f, keys.data(), keys.size()) || !rd(f, tsdf.data(), tsdf.size()) || !rd(f, esdf.data(), esdf.size()) ||
      !rd(f, samples.data(), samples.size()))
    return 2;
  std::fclose(f);
  int n_gpus = std::atoi(argv[3]);
  if (n_gpus <= 0 && cudaGetDeviceCount(&n_gpus) != cudaSuccess) return 3;
  std::vector<int> devices(n_gpus);
  for (int i = 0; i < n_gpus; ++i) devices[i] = i;
  try {
    nbvgpu_shim::GpuMap map(devices, voxel_size, vps, (size_t)nb + 8, voxel_size, vps, (size_t)nb + 8);
    map.uploadSerialized(map.tsdf_layer(), (size_t)nb, keys.data(), tsdf.data());
    map.uploadSerialized(map.esdf_layer(), (size_t)nb, keys.data(), esdf.data());
    nbvgpu_shim::setUpCamera(map, cam[0], cam[1], cam[2], cam[3], cam + 4, cam + 8, cam + 11, cam + 14, cam[17], cam[18], cam[19]);
    const double radius = cam[20];
    std::vector<double> pos((size_t)ns * 3), seg((size_t)ns * 6), pyaw(ns, 0.0), dist(ns), gain(ns);
    std::vector<int32_t> yaw(ns);
    std::vector<uint8_t> fr(ns);
    for (int i = 0; i < ns; ++i) {
      for (int k = 0; k < 3; ++k) pos[3 * i + k] = samples[9 * i + k];
      for (int k = 0; k < 6; ++k) seg[6 * i + k] = samples[9 * i + 3 + k];
      dist[i] = 1.0 + 0.001 * i;
    }
    using clk = std::chrono::steady_clock;
    nbvgpu_best best{};
    double ms = 0.0;
    for (int rep = 0; rep < reps + 1; ++rep) {
      const auto t0 = clk::now();
      best = nbvgpu_shim::evaluateCandidates(map, radius, (size_t)ns, pos.data(), seg.data(), pyaw.data(), dist.data(), 1.0, 0.5, 0.0,
                                             fr.data(), gain.data(), yaw.data());
      if (rep) ms += std::chrono::duration<double, std::milli>(clk::now() - t0).count();
    }
    // single calls on the same (multi-GPU) handle
    Pose st{{pos[0], pos[1], pos[2], 0.0}};
    const double g1 = nbvgpu_shim::optimized_gain(map, st);
    const bool f1 = nbvgpu_shim::checkMotion(map, radius, Point{{seg[0], seg[1], seg[2]}}, Point{{seg[3], seg[4], seg[5]}});
    std::FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 2;
    for (int i = 0; i < ns; ++i) {
      const double row[3] = {gain[i], (double)yaw[i], (double)fr[i]};
      std::fwrite(row, sizeof(double), 3, o);
    }
    const double tail[8] = {(double)best.index, best.score, best.gain, best.yaw, g1, st[3], f1 ? 1.0 : 0.0, (double)n_gpus};
    std::fwrite(tail, sizeof(double), 8, o);
    std::fclose(o);
    std::printf("{\"gpus\": %d, \"candidates\": %d, \"ms_per_batch\": %.4f}\n", n_gpus, ns, reps ? ms / reps : 0.0);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
