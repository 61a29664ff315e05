// Compile/link test of include/nbvgpu_shim.hpp against libnbvgpu.so with stand-ins for
// Eigen::Vector4d / Vector3d.  Without a CUDA device the constructor must fail loudly
// (NBVGPU_ERR_NO_DEVICE -> exception): there is no CPU fallback to drop into.
#include <cstdio>

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

int main() {
  try {
    nbvgpu_shim::GpuMap map(0, 0.3f, 16, 64, 0.05f, 16, 64);
    const double q[4] = {1, 0, 0, 0}, t[3] = {0, 0, 0}, lo[3] = {-4, -1, 0}, hi[3] = {10, 14, 3};
    nbvgpu_shim::setUpCamera(map, 69.0, 42.0, 0.1, 3.0, q, t, lo, hi, 0.0, 0.0, 1.0);
    Pose s{{1.0, 2.0, 1.0, 0.5}};
    const double g = nbvgpu_shim::optimized_gain(map, s);  // empty map: everything unknown
    const bool free_seg = nbvgpu_shim::checkMotion(map, 0.4, Point{{1, 2, 1}}, Point{{2, 2, 1}});  // unallocated -> collision
    const double potential = nbvgpu_shim::bfs(map, Point{{1, 2, 1}});  // empty map: the six neighbours are unknown
    double batch[2] = {0, 0};
    const double pts[6] = {1, 2, 1, 2, 2, 1};
    nbvgpu_shim::bfs_batch(map, 2, pts, batch);
    double dist = -1.0;
    Point grad{{9, 9, 9}};
    const bool interp_ok = nbvgpu_shim::getDistanceAndGradientAtPosition(map, Point{{1, 2, 1}}, &dist, &grad);  // no block: false, outputs 0
    std::printf("gpu gain=%.9g yaw=%.9g free=%d potential=%.9g interp=%d\n", g, s[3], (int)free_seg, potential, (int)interp_ok);
    return (g > 0.0 && !free_seg && potential > 0.0 && batch[0] == potential && batch[1] == potential && !interp_ok && dist == 0.0 &&
            grad.v[0] == 0.0)
               ? 0
               : 3;
  } catch (const std::exception& e) {
    std::printf("no-device: %s\n", e.what());
    return 42;
  }
}
