// camera_host.cc -- host half of the camera model: everything in
// CameraModel::setIntrinsicsFromFoV / setBodyPose (nbveplanner/src/camera_model.cpp:25-96)
// and the per-yaw pose of RrtTree::optimized_gain (nbveplanner/src/rrt.cpp:369-372) that does
// NOT depend on the candidate position.  For each of the 360 integer yaws the world-frame
// frustum corner is  corner_k = q_GC(i).rotate(c_k) + (p + q_yaw(i).rotate(t_BC)); the two
// rotations are tabulated here once per camera (libm sin/cos/tan on the host, so there is no
// device-libm ulp drift) and the kernel only performs the two additions, in the reference's
// order.  Compiled by g++ with -ffp-contract=off: products and sums must not fuse.
#include "camera_host.h"

#include <cmath>

namespace nbvgpu_host {
namespace {

struct V3 { double x, y, z; };
struct Q4 { double w, x, y, z; };

inline V3 cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen QuaternionBase::_transformVector: uv = q.vec x v; uv += uv; v + w*uv + q.vec x uv.
inline V3 rotate(const Q4& q, const V3& v) {
  const V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  const V3 c = cross(qv, uv);
  return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}
inline Q4 qmul(const Q4& a, const Q4& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

}  // namespace

bool build_camera_tables(const nbvgpu_camera& cam, CameraTables* out) {
  const double qn = cam.q_CB[0] * cam.q_CB[0] + cam.q_CB[1] * cam.q_CB[1] + cam.q_CB[2] * cam.q_CB[2] +
                    cam.q_CB[3] * cam.q_CB[3];
  if (!(std::fabs(qn - 1.0) < 1e-9)) return false;  // minkindr CHECKs a unit quaternion
  if (!(cam.near_m > 0.0) || !(cam.far_m > cam.near_m)) return false;
  if (!(cam.hfov_deg > 0.0 && cam.hfov_deg <= 360.0 && cam.vfov_deg > 0.0 && cam.vfov_deg < 180.0)) return false;
  // setIntrinsicsFromFoV (camera_model.cpp:25-73); hfov == 360 selects the reference's lidar
  // mode, whose 24-plane test is undefined behaviour there -- the 6-plane frustum of a 90 degree
  // sector (what the reference builds first, :37) is used with the all-yaw sum instead.
  const double hfov_deg = cam.hfov_deg == 360.0 ? 90.0 : cam.hfov_deg;
  const double hfov = hfov_deg * M_PI / 180.0;
  const double vfov = cam.vfov_deg * M_PI / 180.0;
  const double th = std::tan(hfov / 2.0);
  const double tv = std::tan(vfov / 2.0);
  const double n = cam.near_m, f = cam.far_m;
  const V3 c[8] = {{n, n * th, n * tv},  {n, n * th, -n * tv}, {n, -n * th, -n * tv}, {n, -n * th, n * tv},
                   {f, f * th, f * tv},  {f, f * th, -f * tv}, {f, -f * th, -f * tv}, {f, -f * th, f * tv}};
  // T_B_C = T_C_B^-1: conjugate quaternion, translation -(q^-1 . t) (camera_model.cpp:92-94).
  const Q4 q_BC{cam.q_CB[0], -cam.q_CB[1], -cam.q_CB[2], -cam.q_CB[3]};
  const V3 r = rotate(q_BC, V3{cam.t_CB[0], cam.t_CB[1], cam.t_CB[2]});
  const V3 t_BC{-r.x, -r.y, -r.z};
  for (int i = -180; i < 180; ++i) {
    const double yaw = M_PI * i / 180.0;  // rrt.cpp:369
    const double ha = 0.5 * yaw;          // Eigen AngleAxis -> Quaternion
    const double s = std::sin(ha);
    const Q4 q_yaw{std::cos(ha), s * 0.0, s * 0.0, s * 1.0};
    const Q4 q_GC = qmul(q_yaw, q_BC);
    const int y = i + 180;
    for (int k = 0; k < 8; ++k) {
      const V3 a = rotate(q_GC, c[k]);
      out->A[y][k][0] = a.x; out->A[y][k][1] = a.y; out->A[y][k][2] = a.z;
    }
    const V3 b = rotate(q_yaw, t_BC);
    out->B[y][0] = b.x; out->B[y][1] = b.y; out->B[y][2] = b.z;
    out->cs[y][0] = (float)std::cos(yaw);
    out->cs[y][1] = (float)std::sin(yaw);
  }
  // Conservative single-precision frustum classifier (camera frame): P_c = R_CB * Rz(-yaw) * (P - p) + t_CB.
  const double w = cam.q_CB[0], x = cam.q_CB[1], yq = cam.q_CB[2], z = cam.q_CB[3];
  const double R[9] = {1 - 2 * (yq * yq + z * z), 2 * (x * yq - w * z),      2 * (x * z + w * yq),
                       2 * (x * yq + w * z),      1 - 2 * (x * x + z * z),   2 * (yq * z - w * x),
                       2 * (x * z - w * yq),      2 * (yq * z + w * x),      1 - 2 * (x * x + yq * yq)};
  for (int k = 0; k < 9; ++k) out->R_CB[k] = (float)R[k];
  for (int k = 0; k < 3; ++k) out->t_CB[k] = (float)cam.t_CB[k];
  out->th = (float)th;
  out->tv = (float)tv;
  out->near_f = (float)n;
  out->far_f = (float)f;
  const double t_norm = std::sqrt(cam.t_CB[0] * cam.t_CB[0] + cam.t_CB[1] * cam.t_CB[1] + cam.t_CB[2] * cam.t_CB[2]);
  // Every traversed voxel lies within reach_m of the candidate position: |t_BC| + distance to the
  // far-plane centre column end points + one voxel diagonal (added by the caller in voxels).
  out->reach_m = t_norm + f * std::sqrt(1.0 + tv * tv);
  // Classifier margin: the fp32 evaluation error of a signed plane distance is bounded by
  // ~24 ulp(2^-24) * reach; 1e-5 * max(10 m, reach) leaves >= 6x slack.
  const double margin = 1e-5 * std::fmax(10.0, out->reach_m + 1.0);
  out->margin = (float)margin;
  out->margin_h = (float)(margin * std::sqrt(1.0 + th * th));
  out->margin_v = (float)(margin * std::sqrt(1.0 + tv * tv));
  out->filter_ok = (hfov_deg < 179.0 && cam.vfov_deg < 179.0) ? 1 : 0;
  out->half_hfov = (int)std::floor(cam.hfov_deg / 2.0);  // rrt.cpp:452
  return true;
}

void box_thresholds(const nbvgpu_camera& cam, double vs, int gmin[3], int gmax[3]) {
  // isInsideBounds (nbveplanner/common.h:22-31) is evaluated on P = (double)g * vs, which is
  // monotone in the integer g, so  min <= P <= max  <=>  gmin <= g <= gmax  exactly.
  const long long LIM = 1ll << 30;
  for (int k = 0; k < 3; ++k) {
    const double lo = cam.bbx_min[k], hi = cam.bbx_max[k];
    long long a, b;
    if (std::isnan(lo) || std::isnan(hi)) { gmin[k] = 1; gmax[k] = 0; continue; }  // never inside
    if (lo <= -(double)LIM * vs) a = -LIM;
    else if (lo >= (double)LIM * vs) a = LIM;
    else {
      a = (long long)std::floor(lo / vs) - 2;
      while (!((double)a * vs >= lo)) ++a;
    }
    if (hi >= (double)LIM * vs) b = LIM;
    else if (hi <= -(double)LIM * vs) b = -LIM;
    else {
      b = (long long)std::floor(hi / vs) + 2;
      while (!((double)b * vs <= hi)) --b;
    }
    gmin[k] = (int)a;
    gmax[k] = (int)b;
  }
}


bool dense_volume_bounds(const CameraTables& tab, double vs, int yc, int* chunk_info, int* ez_out, long long* max_words) {
  // Every voxel a ray of yaw y visits lies in the box spanned by its start voxel and end voxel
  // (+-1 for the DDA's rounding at the very end).  Relative to the candidate position p, the start is
  // B[y] and the end points of the ray column are A[y][5] + v_center + B[y] + u * u_slope * vs,
  // u in [0, u_max): a straight line, so its extremes are at the two ends.  index(p + rel) - index(p) is
  // floor(rel / vs) or that + 1, hence the asymmetric padding.
  // chunk_info[c] = {min x, min y, min z, words along x (worst-case 16-alignment), voxels along y}.
  const int chunks = 360 / yc;
  int ez = 0;
  long long words = 0;
  for (int c = 0; c < chunks; ++c) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int y = c * yc; y < (c + 1) * yc; ++y) {
      const double* a4 = tab.A[y][4];
      const double* a5 = tab.A[y][5];
      const double* a6 = tab.A[y][6];
      double ud[3], n2 = 0.0;
      for (int k = 0; k < 3; ++k) { ud[k] = a4[k] - a5[k]; n2 += ud[k] * ud[k]; }
      const double n = std::sqrt(n2);
      if (!(n > 0.0)) return false;
      const double umax = std::ceil(n / vs) + 1.0;
      for (int k = 0; k < 3; ++k) {
        const double s = tab.B[y][k];
        const double e0 = a5[k] + (a6[k] - a5[k]) / 2.0 + tab.B[y][k];
        const double e1 = e0 + umax * (ud[k] / n) * vs;
        lo[k] = std::fmin(lo[k], std::fmin(s, std::fmin(e0, e1)));
        hi[k] = std::fmax(hi[k], std::fmax(s, std::fmax(e0, e1)));
      }
    }
    int ext[3];
    for (int k = 0; k < 3; ++k) {
      const double a = std::floor(lo[k] / vs) - 2.0, b = std::floor(hi[k] / vs) + 3.0;
      if (!(std::fabs(a) < 1e6) || !(std::fabs(b) < 1e6)) return false;
      chunk_info[5 * c + k] = (int)a;
      ext[k] = (int)(b - a) + 1;
    }
    chunk_info[5 * c + 3] = (ext[0] + 15) / 16 + 1;
    if (chunk_info[5 * c + 3] > 32) return false;  // kMaxDenseWordsX: the kernel keeps one box mask per x word
    chunk_info[5 * c + 4] = ext[1];
    if (ext[2] > ez) ez = ext[2];
    const long long w = (long long)chunk_info[5 * c + 3] * ext[1];
    if (w > words) words = w;
  }
  *ez_out = ez;
  *max_words = words * ez;
  return true;
}

}  // namespace nbvgpu_host
