// camera_host.h -- position-independent camera tables built on the host (see camera_host.cc).
#ifndef NBVGPU_CAMERA_HOST_H_
#define NBVGPU_CAMERA_HOST_H_

#include "nbvgpu.h"

namespace nbvgpu_host {

struct CameraTables {
  double A[360][8][3];  // q_GC(i).rotate(c_k): frustum corners rotated into the world frame
  double B[360][3];     // q_yaw(i).rotate(t_BC): camera offset from the body position
  float cs[360][2];     // cos / sin of the yaw (single-precision classifier)
  float R_CB[9];        // camera <- body rotation, row major
  float t_CB[3];
  float th, tv, near_f, far_f;
  float margin, margin_h, margin_v;
  double reach_m;       // bound on |voxel - position| over all traversed voxels (without voxel slack)
  int filter_ok;        // single-precision classifier usable (fov < 179 deg)
  int half_hfov;        // floor(hfov / 2)  (rrt.cpp:452)
};

// false when the camera parameters are invalid (non-unit quaternion, near >= far, fov range).
bool build_camera_tables(const nbvgpu_camera& cam, CameraTables* out);

// Integer voxel-index thresholds equivalent to the exploration-box test on (double)g * vs.
void box_thresholds(const nbvgpu_camera& cam, double vs, int gmin[3], int gmax[3]);

// Conservative extent of the voxels visited by the rays of each chunk of `yc` consecutive yaws,
// relative to the voxel that contains the candidate position: chunk_info[chunk][5] = {lower corner
// x, y, z (padded), 16-voxel words along x, voxels along y}; *ez = voxels along z (common);
// *max_words = largest volume in 32-bit words over the chunks.  Sizes the dense shared-memory status
// volume of k_gain's kDense mode.  false if the camera geometry is degenerate.
bool dense_volume_bounds(const CameraTables& tab, double vs, int yc, int* chunk_info, int* ez, long long* max_words);

}  // namespace nbvgpu_host
#endif
