// gain_kernels.cuh -- sm_100a kernels for the viewpoint-gain evaluator and its satellites.
//
//   k_gain<V,LV,MODE,FUSE,THREADS>
//                      RrtTree::optimized_gain / gain ray sweep   nbveplanner/src/rrt.cpp:351-447, 481-571
//                      + CameraModel planes / in-view test        nbveplanner/src/camera_model.cpp:5-23,98-164,191-201
//                      + voxblox RayCaster                        voxblox/src/integrator/integrator_utils.cc:111-179
//                      + getVoxelStatusByIndex                    nbveplanner/src/voxblox_manager.cpp:14-33
//   k_best_yaw         best-yaw window scan                       nbveplanner/src/rrt.cpp:449-478
//                      (FUSE: run by the last CTA of a candidate inside k_gain, small batches)
//   k_score_best       node scoring / best node                   nbveplanner/src/rrt.cpp:220-246
//   k_check_segments   HighResManager::checkMotion                nbveplanner/include/nbveplanner/voxblox_manager.h:89-108
//   k_check_segments_warp   the same, one warp per segment (small batches)
//   k_expand_sample    RrtTree::iterate front half (batched)      nbveplanner/src/rrt.cpp:172-212,223
//
// Work decomposition of k_gain: one CTA per (candidate position, chunk of `yc` consecutive
// yaws).  The CTA prologue (a) probes the block hash once for every block the candidate can
// reach and stores the tile slots in a dense shared-memory grid, so the ray walk never hashes,
// and (b) builds, one thread per yaw, the six frustum planes and the ray column frame in FP64
// with the reference's operation order (no FMA: __dmul_rn/__dadd_rn).  Rays are then walked one
// per lane; the 32 lanes of a warp take 32 consecutive yaws of the same ray row u, i.e. rays one
// degree apart that stay within a voxel or two of each other: near-identical trip counts (little
// divergence) and loads that fall into the same sectors.  Per-yaw counts are accumulated with
// shared-memory atomics and written once per CTA.  MODE selects how a voxel's status is found (dense 2-bit
// shared-memory volume, slot grid + status tiles, direct hash probes), THREADS the CTA width (256: four
// CTAs per SM; 512: two, for volumes of up to half an SM's shared memory and for 90-yaw chunks of large
// batches); the host picks them per launch (csrc/nbvgpu.cu: enqueue_gain).
//
// Exactness: the DDA is a literal float transcription (IEEE division, inf/NaN semantics kept);
// the exploration-box test is done on integer voxel thresholds (exactly equivalent, see
// camera_host.cc); the six FP64 plane tests are preceded by a conservative single-precision
// classifier in the camera frame that decides "surely inside" / "surely outside" with a margin
// far above its rounding error and otherwise falls back to the exact FP64 planes.  All variants and
// modes therefore produce identical counts (see the comment above k_gain).
#ifndef NBVGPU_GAIN_KERNELS_CUH_
#define NBVGPU_GAIN_KERNELS_CUH_

#include "map_mirror.cuh"

namespace nbvgpu {

constexpr int kGainThreads = 256;
constexpr int kMaxReachVox = 4096;
constexpr int kMaxDenseWordsX = 32;  // 16-voxel words along x of the dense shared-memory volume

struct FilterConst {
  float R[9];  // camera <- body rotation (row major)
  float t[3];  // t_CB
  float th, tv, nearp, farp;
  float m, mh, mv;  // margins (plain, scaled for the horizontal / vertical side planes)
};

// arguments of the best-yaw scan (best_yaw_cta below)
struct BestYawArgs {
  const uint32_t* per_yaw;  // [n][360][3]
  double* gain;             // [n]
  int32_t* yaw_deg;         // [n] or null
  unsigned long long* counts;  // [n][3] or null
  double g_unmapped, g_free, g_occupied;
  double cube;
  int half_hfov;
  int sum_all;
  int n;
  // single-candidate host calls: after the results are written (to mapped host memory) the writing thread publishes
  // `seq` at `done` (mapped host memory as well), which the host polls instead of synchronising the stream; else null
  unsigned* done;
  unsigned seq;
};

// Publish the completion ticket of a single-item host call: results first, system-wide fence, then the ticket.
__device__ __forceinline__ void publish_done(unsigned* done, unsigned seq) {
  if (done) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(done) = seq;
  }
}

struct GainArgs {
  LayerView L;
  const double* A;    // [360][24]
  const double* B;    // [360][3]
  const float* cs;    // [360][2]
  const double* pos;  // [n][3]
  const uint8_t* valid;  // [n] or null: candidates with valid == 0 are skipped (counts 0)
  uint32_t* per_yaw;  // [n][360][3]
  unsigned long long* steps;   // [n] or null
  unsigned long long* totals;  // [0] rays [1] voxel steps [3] classifier mismatches
  double vs, vs_inv;
  float vs_f;
  int gmin[3];
  unsigned gspan[3];
  int reach_vox;
  int gz_lo;          // lower end of the slot grid along z, in voxels relative to the candidate's voxel (>= -reach_vox: the rays'
                      // own z-extent, so that fine voxels do not pay for a cube of blocks the frustum never reaches)
  int unknown_slot;   // tile index of the all-unknown / zero tile (== max_blocks)
  int gnx, gny, gnz;  // local block grid dimensions
  int yc;             // yaws per CTA (divides 360)
  int n;
  // dense shared-memory status volume (MODE kDense): per-chunk lower corner relative to the
  // candidate's voxel, and the common dimensions {16-voxel words along x, voxels along y, z}
  const int* vmin;    // [360 / yc][5]: lower corner x, y, z; words along x; voxels along y
  int vez, vwords;    // voxels along z (common); shared-memory words reserved for the volume
  // z-clipped volume (fine voxels: the frustum reaches far below / above the exploration box, where nothing can count):
  // a CTA keeps the rows of the box slab +- 3 voxels plus the rows its rays start in; zclip = 0: no clipping
  int zclip, sz_lo, sz_hi, vez_alloc;  // start rows relative to the candidate's voxel; rows the host reserved memory for
  // fused finish (k_gain<..., FUSE = true>, small batches): the last CTA of a candidate to arrive runs the window
  // scan, so that a single-candidate evaluation is one launch.  arrive[v] counts CTAs and is reset by the last one.
  BestYawArgs by;
  int* arrive;
  double pos1[3];     // the position of a single candidate passed by value (used when pos == nullptr)
  FilterConst F;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 d3sub(const D3& a, const D3& b) { return {dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)}; }
__device__ __forceinline__ D3 d3add(const D3& a, const D3& b) { return {dadd(a.x, b.x), dadd(a.y, b.y), dadd(a.z, b.z)}; }
// Eigen 3-vector dot on SSE2: (x0*y0 + x1*y1) + x2*y2.
__device__ __forceinline__ double d3dot(const D3& a, const D3& b) {
  return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z));
}
__device__ __forceinline__ D3 d3cross(const D3& a, const D3& b) {
  return {dsub(dmul(a.y, b.z), dmul(a.z, b.y)), dsub(dmul(a.z, b.x), dmul(a.x, b.z)),
          dsub(dmul(a.x, b.y), dmul(a.y, b.x))};
}
__device__ __forceinline__ D3 d3normalized(const D3& a) {
  const double z = d3dot(a, a);
  if (z > 0.0) {
    const double n = __dsqrt_rn(z);
    return {__ddiv_rn(a.x, n), __ddiv_rn(a.y, n), __ddiv_rn(a.z, n)};
  }
  return a;
}

// Shared-memory carve-up (dynamic): doubles first, then 4-byte arrays.
struct GainSmem {
  double* pl;     // [24][yc]  plane k component c at pl[(4k+c)*yc + j]
  double* rd;     // [9][yc]   pp1 (3), u_slope (3), v_center (3)
  float* st;      // [3][yc]   scaled ray start
  float* csf;     // [2][yc]
  int* um;        // [yc]
  unsigned* cnt;  // [3][yc]
  float* fp;      // [12][yc]  far / top / bottom plane in single precision, relative to the candidate voxel
  float* sl;      // [6][yc]   per-plane corner slack: vs * sum_i max(n_i, 0)
  int* grid;      // [gnz][gny][gnx]
};
// The slot grid sits at offset 0 so that its address is a compile-time constant in the ray loop.
__device__ __forceinline__ GainSmem carve(unsigned char* base, int yc, int gn) {
  GainSmem s;
  s.grid = reinterpret_cast<int*>(base);
  s.pl = reinterpret_cast<double*>(base + (((size_t)gn * 4 + 15) & ~(size_t)15));
  s.rd = s.pl + 24 * yc;
  s.st = reinterpret_cast<float*>(s.rd + 9 * yc);
  s.csf = s.st + 3 * yc;
  s.um = reinterpret_cast<int*>(s.csf + 2 * yc);
  s.cnt = reinterpret_cast<unsigned*>(s.um + yc);
  s.fp = reinterpret_cast<float*>(s.cnt + 3 * yc);
  s.sl = s.fp + 12 * yc;
  return s;
}
inline size_t gain_smem_bytes(int yc, int gn) { return (size_t)yc * (33 * 8 + 27 * 4) + (((size_t)gn * 4 + 15) & ~(size_t)15) + 16; }

// CameraModel::isPointInView plane part (camera_model.cpp:194-199) on P = (double)g * vs.
__device__ __forceinline__ bool exact_in_frustum(const double* pl, int yc, int j, int gx, int gy, int gz, double vs) {
  const double px = dmul((double)gx, vs), py = dmul((double)gy, vs), pz = dmul((double)gz, vs);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double nx = pl[(4 * k + 0) * yc + j], ny = pl[(4 * k + 1) * yc + j], nz = pl[(4 * k + 2) * yc + j];
    const double d = pl[(4 * k + 3) * yc + j];
    if (!(dadd(dadd(dmul(px, nx), dmul(py, ny)), dmul(pz, nz)) >= d)) return false;
  }
  return true;
}

// CTA prologue, part (b): per-yaw frames (rrt.cpp:369-385, camera_model.cpp:87-94,119-160,280-289), one
// thread per (yaw, plane) plus one per yaw for the ray column, so that the FP64 dependency chains of a
// chunk's 40 yaws run side by side on all warps instead of on two.  Kept out of line so that their FP64
// register appetite does not set the register budget of the ray loop.
//
// Plane k of a yaw: Plane::setFromPoints(c[i1], c[i2], c[i3]) with the corner triples of
// CameraModel::calculateBoundingPlanes (near 0,2,1; far 4,5,6; left 3,6,2; right 0,5,4; top 3,4,7;
// bottom 2,6,5), c[i] = A[i] + (p + B).
__device__ __noinline__ void build_yaw_plane(const double* __restrict__ A, const double* __restrict__ B, int k, double* s_pl,
                                             float* s_fp, float* s_sl, double vs, int g0x, int g0y, int g0z, int yc, int j,
                                             double px, double py, double pz) {
  const D3 t = d3add(D3{px, py, pz}, D3{B[0], B[1], B[2]});  // camera position
  const int i1 = (0x230340 >> (4 * k)) & 15, i2 = (0x645652 >> (4 * k)) & 15, i3 = (0x574261 >> (4 * k)) & 15;
  const D3 p1 = d3add(D3{A[3 * i1], A[3 * i1 + 1], A[3 * i1 + 2]}, t);
  const D3 p2 = d3add(D3{A[3 * i2], A[3 * i2 + 1], A[3 * i2 + 2]}, t);
  const D3 p3 = d3add(D3{A[3 * i3], A[3 * i3 + 1], A[3 * i3 + 2]}, t);
  // Plane::setFromPoints (camera_model.cpp:5-11)
  const D3 n = d3normalized(d3cross(d3sub(p2, p1), d3sub(p3, p1)));
  const double d = d3dot(n, p1);
  s_pl[(4 * k + 0) * yc + j] = n.x;
  s_pl[(4 * k + 1) * yc + j] = n.y;
  s_pl[(4 * k + 2) * yc + j] = n.z;
  s_pl[(4 * k + 3) * yc + j] = d;
  // Single-precision copies of the far (1), top (4) and bottom (5) planes for the one-plane tests:
  // s(g) = n.(g*vs) - d = sum_k (n_k*vs)*(g_k - g0_k) + (n.(g0*vs) - d); |error| ~ 3e-6 m << margin.
  if (k == 1 || k >= 4) {
    const int q = k == 1 ? 0 : k - 3;
    s_fp[(4 * q + 0) * yc + j] = __double2float_rn(n.x * vs);
    s_fp[(4 * q + 1) * yc + j] = __double2float_rn(n.y * vs);
    s_fp[(4 * q + 2) * yc + j] = __double2float_rn(n.z * vs);
    s_fp[(4 * q + 3) * yc + j] = __double2float_rn((n.x * (double)g0x + n.y * (double)g0y + n.z * (double)g0z) * vs - d);
  }
  // Corner slack: the tested point is the MIN corner of the voxel that contains the ray point X
  // (g <= X/vs <= g+1), so n.(g*vs) >= n.X - vs * sum_i max(n_i, 0); rounded up.
  const double sx = fmax(n.x, 0.0), sy = fmax(n.y, 0.0), sz = fmax(n.z, 0.0);
  s_sl[k * yc + j] = __double2float_ru((sx + sy + sz) * vs * 1.001);
}

// Ray column of a yaw (rrt.cpp:382-385,408-409, camera_model.cpp:280-289) and the per-yaw counters.
__device__ __noinline__ void build_yaw_rays(const double* __restrict__ A, const double* __restrict__ B,
                                            const float* __restrict__ cs, double vs_inv, double* s_rd, float* s_st,
                                            float* s_csf, int* s_um, unsigned* s_cnt, int yc, int j, double px, double py,
                                            double pz) {
  const D3 t = d3add(D3{px, py, pz}, D3{B[0], B[1], B[2]});  // camera position
  const D3 c4 = d3add(D3{A[12], A[13], A[14]}, t), c5 = d3add(D3{A[15], A[16], A[17]}, t), c6 = d3add(D3{A[18], A[19], A[20]}, t);
  const D3 ud = d3sub(c4, c5);                                         // rrt.cpp:382
  const D3 us = d3normalized(ud);                                      // :383
  const int um = (int)ceil(dmul(__dsqrt_rn(d3dot(ud, ud)), vs_inv));  // :384
  const D3 d21 = d3sub(c6, c5);
  const D3 vc{__ddiv_rn(d21.x, 2.0), __ddiv_rn(d21.y, 2.0), __ddiv_rn(d21.z, 2.0)};  // :385
  s_rd[0 * yc + j] = c5.x; s_rd[1 * yc + j] = c5.y; s_rd[2 * yc + j] = c5.z;
  s_rd[3 * yc + j] = us.x; s_rd[4 * yc + j] = us.y; s_rd[5 * yc + j] = us.z;
  s_rd[6 * yc + j] = vc.x; s_rd[7 * yc + j] = vc.y; s_rd[8 * yc + j] = vc.z;
  s_st[0 * yc + j] = __double2float_rn(dmul(t.x, vs_inv));  // :408-409
  s_st[1 * yc + j] = __double2float_rn(dmul(t.y, vs_inv));
  s_st[2 * yc + j] = __double2float_rn(dmul(t.z, vs_inv));
  s_csf[0 * yc + j] = cs[0];
  s_csf[1 * yc + j] = cs[1];
  s_um[j] = um;
  s_cnt[0 * yc + j] = 0u; s_cnt[1 * yc + j] = 0u; s_cnt[2 * yc + j] = 0u;
}

// Camera-frame coordinates of a point given as (float) voxel offsets from g0 (classifier math).
__device__ __forceinline__ void to_camera(const FilterConst& F, float vs_f, float dx, float dy, float dz, float noffx,
                                          float noffy, float noffz, float cy, float sy, float& xc, float& yk, float& zk) {
  const float rx = fmaf(dx, vs_f, noffx), ry = fmaf(dy, vs_f, noffy), rz = fmaf(dz, vs_f, noffz);
  const float x1 = fmaf(sy, ry, cy * rx);
  const float y1 = fmaf(cy, ry, -(sy * rx));
  xc = fmaf(F.R[0], x1, fmaf(F.R[1], y1, fmaf(F.R[2], rz, F.t[0])));
  yk = fmaf(F.R[3], x1, fmaf(F.R[4], y1, fmaf(F.R[5], rz, F.t[1])));
  zk = fmaf(F.R[6], x1, fmaf(F.R[7], y1, fmaf(F.R[8], rz, F.t[2])));
}

// RayCaster::nextRayIndex advance (integrator_utils.cc:119-123): the axis with the FIRST minimum t steps
// (compare + select, so a NaN in t[1] / t[2] never wins and a NaN in t[0] always does -- the order of the
// reference's minCoeff), t of that axis grows by dt, the voxel index by the step sign.  `tcur` receives the
// ray parameter at which the next voxel is entered.  Written as predicated PTX: 6 compare/select + one
// predicated add per carried quantity (HASP: bit address of the dense volume; HASSV: signed distance to
// the open top/bottom plane).
template <bool HASP, bool HASSV, bool HASG = true>
__device__ __forceinline__ void dda_advance(float (&t)[3], const float (&dt)[3], float& tcur, int (&g)[3], const int (&sg)[3],
                                            unsigned& P, int stx, int sty, int stz, float& sv, float d0, float d1, float d2) {
#define NBV_DDA_SELECT                       \
  "{\n\t"                                    \
  ".reg .pred c1, m0, m1, m2;\n\t"           \
  ".reg .f32 bm;\n\t"                        \
  "setp.lt.f32 c1, %1, %0;\n\t"              \
  "selp.f32 bm, %1, %0, c1;\n\t"             \
  "setp.lt.f32 m2, %2, bm;\n\t"              \
  "selp.f32 %3, %2, bm, m2;\n\t"             \
  "setp.geu.and.f32 m1, %2, bm, c1;\n\t"     \
  "setp.geu.and.f32 m0, %2, bm, !c1;\n\t"
  // Dense volume, voxel index not carried (rebuilt from P when a stretch ends).  Only used where t[0] cannot be NaN (rays without
  // an x component end in the general loop), so the first minimum is a three-input min and two equality tests, and the bit
  // address moves by ONE add of the selected stride.
  if (HASP && !HASG && HASSV) {
    asm("{\n\t"
        ".reg .pred m0, m1, m2;\n\t"
        ".reg .u32 st;\n\t"
        ".reg .f32 ds;\n\t"
        "min.f32 %3, %0, %1, %2;\n\t"
        "setp.eq.f32 m0, %0, %3;\n\t"
        "setp.eq.and.f32 m1, %1, %3, !m0;\n\t"
        "setp.neu.and.f32 m2, %1, %3, !m0;\n\t"
        "@m0 add.rn.f32 %0, %0, %6;\n\t"
        "@m1 add.rn.f32 %1, %1, %7;\n\t"
        "@m2 add.rn.f32 %2, %2, %8;\n\t"
        "selp.b32 st, %10, %11, m1;\n\t"
        "selp.b32 st, %9, st, m0;\n\t"
        "add.s32 %4, %4, st;\n\t"
        "selp.f32 ds, %13, %14, m1;\n\t"
        "selp.f32 ds, %12, ds, m0;\n\t"
        "add.rn.f32 %5, %5, ds;\n\t"
        "}"
        : "+f"(t[0]), "+f"(t[1]), "+f"(t[2]), "=&f"(tcur), "+r"(P), "+f"(sv)
        : "f"(dt[0]), "f"(dt[1]), "f"(dt[2]), "r"(stx), "r"(sty), "r"(stz), "f"(d0), "f"(d1), "f"(d2));
  } else if (HASP && !HASG) {
    asm("{\n\t"
        ".reg .pred m0, m1, m2;\n\t"
        ".reg .u32 st;\n\t"
        "min.f32 %3, %0, %1, %2;\n\t"
        "setp.eq.f32 m0, %0, %3;\n\t"
        "setp.eq.and.f32 m1, %1, %3, !m0;\n\t"
        "setp.neu.and.f32 m2, %1, %3, !m0;\n\t"
        "@m0 add.rn.f32 %0, %0, %5;\n\t"
        "@m1 add.rn.f32 %1, %1, %6;\n\t"
        "@m2 add.rn.f32 %2, %2, %7;\n\t"
        "selp.b32 st, %9, %10, m1;\n\t"
        "selp.b32 st, %8, st, m0;\n\t"
        "add.s32 %4, %4, st;\n\t"
        "}"
        : "+f"(t[0]), "+f"(t[1]), "+f"(t[2]), "=&f"(tcur), "+r"(P)
        : "f"(dt[0]), "f"(dt[1]), "f"(dt[2]), "r"(stx), "r"(sty), "r"(stz));
  } else if (HASP && HASSV) {
    asm(NBV_DDA_SELECT
        "@m0 add.rn.f32 %0, %0, %9;\n\t"
        "@m1 add.rn.f32 %1, %1, %10;\n\t"
        "@m2 add.rn.f32 %2, %2, %11;\n\t"
        "@m0 add.s32 %4, %4, %12;\n\t"
        "@m1 add.s32 %5, %5, %13;\n\t"
        "@m2 add.s32 %6, %6, %14;\n\t"
        "@m0 add.s32 %7, %7, %15;\n\t"
        "@m1 add.s32 %7, %7, %16;\n\t"
        "@m2 add.s32 %7, %7, %17;\n\t"
        "@m0 add.rn.f32 %8, %8, %18;\n\t"
        "@m1 add.rn.f32 %8, %8, %19;\n\t"
        "@m2 add.rn.f32 %8, %8, %20;\n\t"
        "}"
        : "+f"(t[0]), "+f"(t[1]), "+f"(t[2]), "=&f"(tcur), "+r"(g[0]), "+r"(g[1]), "+r"(g[2]), "+r"(P), "+f"(sv)
        : "f"(dt[0]), "f"(dt[1]), "f"(dt[2]), "r"(sg[0]), "r"(sg[1]), "r"(sg[2]), "r"(stx), "r"(sty), "r"(stz), "f"(d0), "f"(d1),
          "f"(d2));
  } else if (HASP) {
    asm(NBV_DDA_SELECT
        "@m0 add.rn.f32 %0, %0, %8;\n\t"
        "@m1 add.rn.f32 %1, %1, %9;\n\t"
        "@m2 add.rn.f32 %2, %2, %10;\n\t"
        "@m0 add.s32 %4, %4, %11;\n\t"
        "@m1 add.s32 %5, %5, %12;\n\t"
        "@m2 add.s32 %6, %6, %13;\n\t"
        "@m0 add.s32 %7, %7, %14;\n\t"
        "@m1 add.s32 %7, %7, %15;\n\t"
        "@m2 add.s32 %7, %7, %16;\n\t"
        "}"
        : "+f"(t[0]), "+f"(t[1]), "+f"(t[2]), "=&f"(tcur), "+r"(g[0]), "+r"(g[1]), "+r"(g[2]), "+r"(P)
        : "f"(dt[0]), "f"(dt[1]), "f"(dt[2]), "r"(sg[0]), "r"(sg[1]), "r"(sg[2]), "r"(stx), "r"(sty), "r"(stz));
  } else if (HASSV) {
    asm(NBV_DDA_SELECT
        "@m0 add.rn.f32 %0, %0, %8;\n\t"
        "@m1 add.rn.f32 %1, %1, %9;\n\t"
        "@m2 add.rn.f32 %2, %2, %10;\n\t"
        "@m0 add.s32 %4, %4, %11;\n\t"
        "@m1 add.s32 %5, %5, %12;\n\t"
        "@m2 add.s32 %6, %6, %13;\n\t"
        "@m0 add.rn.f32 %7, %7, %14;\n\t"
        "@m1 add.rn.f32 %7, %7, %15;\n\t"
        "@m2 add.rn.f32 %7, %7, %16;\n\t"
        "}"
        : "+f"(t[0]), "+f"(t[1]), "+f"(t[2]), "=&f"(tcur), "+r"(g[0]), "+r"(g[1]), "+r"(g[2]), "+f"(sv)
        : "f"(dt[0]), "f"(dt[1]), "f"(dt[2]), "r"(sg[0]), "r"(sg[1]), "r"(sg[2]), "f"(d0), "f"(d1), "f"(d2));
  } else {
    asm(NBV_DDA_SELECT
        "@m0 add.rn.f32 %0, %0, %7;\n\t"
        "@m1 add.rn.f32 %1, %1, %8;\n\t"
        "@m2 add.rn.f32 %2, %2, %9;\n\t"
        "@m0 add.s32 %4, %4, %10;\n\t"
        "@m1 add.s32 %5, %5, %11;\n\t"
        "@m2 add.s32 %6, %6, %12;\n\t"
        "}"
        : "+f"(t[0]), "+f"(t[1]), "+f"(t[2]), "=&f"(tcur), "+r"(g[0]), "+r"(g[1]), "+r"(g[2])
        : "f"(dt[0]), "f"(dt[1]), "f"(dt[2]), "r"(sg[0]), "r"(sg[1]), "r"(sg[2]));
  }
#undef NBV_DDA_SELECT
}

// ---------------------------------------------------------------------------------------------
// k_gain<VARIANT, LV, MODE> -- the ray sweep.
//
// MODE selects how a step reads the map:
//   kDense  the CTA first copies the 2-bit status of every voxel its rays can reach into a dense
//           shared-memory volume [z][y][x words] (x word-aligned to the 16-voxel block rows, so whole
//           status-tile words are copied), with the exploration box folded in as a 4th value.  A lane
//           addresses it with ONE integer, the bit address P = 8*smem_base + ((z*Y + y)*Xw)*32 + 2*x:
//           a DDA step along x/y/z adds +-2 / +-32*Xw / +-32*Xw*Y (the x carry into the next word is
//           free) and the status is (lds[(P>>3)&~3] >> (P&31)) & 3 -- no box test, block lookup, tile
//           address arithmetic or global load in the ray loop.  Needs 16^3 blocks, a volume that fits
//           the shared-memory budget and rays shorter than 255 steps (byte-wide counters).
//   kGrid   integer box test, tile slot from the shared-memory slot grid (one hash probe per reachable
//           block in the prologue), one word of the status tile (or the {distance, weight} tile) through
//           the read-only path.
//   kProbe  as kGrid but the reach is too large for a shared-memory grid: global hash probes with a
//           per-lane last-block cache.
// VARIANT: 0 default; 1 plain transcription (FP64 planes on every step, {distance, weight} tiles);
//   2 default with every shortcut cross-checked against the exact test (mismatch counter);
//   3 default decisions but {distance, weight} tile loads.  LV: log2(voxels per side), 0 = run time.
//
// Voxel codes inside the loop: 0 unknown, 1 free, 2 skip (not in view), 3 occupied (stops the ray).
//
// In-view decision per step, cheapest first (all exactness-preserving):
//   * every voxel is entered at a known ray parameter tcur (the DDA's own t).  The tested point is the
//     MIN corner of the voxel that contains the ray point X, so n.corner >= n.X - vs*sum_i max(n_i,0),
//     and n.X is linear in t.  Per ray: tlo (near, left, right and the farther of top/bottom are sure
//     beyond it), tlo_v (the nearer of top/bottom is sure beyond it), thi (the far plane is sure before
//     it), t_out (past it the ray has left the convex exploration box for good: the remaining voxels
//     can neither count nor stop the ray -- they are accounted as executed and skipped);
//   * between tlo and the all-sure window only the far plane and/or the nearer of top/bottom are open:
//     one single-precision plane evaluation each, decided by the smaller signed distance;
//   * else the single-precision classifier in the camera frame, and the exact FP64 planes when that
//     is within its margin.
// ---------------------------------------------------------------------------------------------
enum { kDense = 0, kGrid = 1, kProbe = 2, kDenseClip = 3 };  // kDenseClip: kDense with the volume clipped in z (own instantiation,
                                                                // so that the common dense kernel carries none of its code)

inline size_t dense_smem_bytes(int yc, int gn, int nvol) {
  return (((size_t)nvol * 4 + 15) & ~(size_t)15) + gain_smem_bytes(yc, gn);
}

// ---------------------------------------------------------------------------------------------
// Best yaw (rrt.cpp:449-478) / all-yaw sum (rrt.cpp:571-578).
// ---------------------------------------------------------------------------------------------
// The whole CTA (a multiple of 32 threads, at least 96) works on candidate v; vg: 360 doubles of shared memory.  The
// counts are read through L2 (__ldcg): in the fused sweep they were written by other CTAs of the same launch.  A window's
// sum must keep the reference's order (floating-point addition), so it is one sequential chain -- but the 72 windows are
// independent: one thread each.  Must be called by all threads of the CTA (barriers inside).
__device__ __forceinline__ void best_yaw_cta(const BestYawArgs& a, int v, double* vg, int tid, int nthreads) {
  __shared__ unsigned long long s_tot[3];
  __shared__ double s_wbest[3];
  __shared__ int s_wc[3];
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t* c = a.per_yaw + (size_t)v * 1080;
  unsigned long long tu = 0, tf = 0, to = 0;
  if (tid < 3) s_tot[tid] = 0ull;
  for (int j = tid; j < 360; j += nthreads) {
    const uint32_t nu = __ldcg(c + 3 * j), nf = __ldcg(c + 3 * j + 1), no = __ldcg(c + 3 * j + 2);
    // vertical_gain[j]: the running sum of rrt.cpp:436-443 in count form.
    vg[j] = dadd(dadd(dmul((double)nu, a.g_unmapped), dmul((double)no, a.g_occupied)), dmul((double)nf, a.g_free));
    tu += nu; tf += nf; to += no;
  }
  __syncthreads();
  if (a.sum_all) {  // RrtTree::gain: integer totals (exact in any order), one product at the end
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tu += __shfl_xor_sync(0xffffffffu, tu, o);
      tf += __shfl_xor_sync(0xffffffffu, tf, o);
      to += __shfl_xor_sync(0xffffffffu, to, o);
    }
    if (lane == 0) {
      atomicAdd(&s_tot[0], tu);
      atomicAdd(&s_tot[1], tf);
      atomicAdd(&s_tot[2], to);
    }
    __syncthreads();
    if (tid == 0) {
      tu = s_tot[0]; tf = s_tot[1]; to = s_tot[2];
      const double g = dadd(dadd(dmul((double)tu, a.g_unmapped), dmul((double)to, a.g_occupied)), dmul((double)tf, a.g_free));
      a.gain[v] = dmul(g, a.cube);
      if (a.yaw_deg) a.yaw_deg[v] = 0;
      if (a.counts) { a.counts[3 * v] = tu; a.counts[3 * v + 1] = tf; a.counts[3 * v + 2] = to; }
      publish_done(a.done, a.seq);
    }
    return;
  }
  // 72 window centres, 5 degrees apart; thread cidx sums its window in the reference's order.
  double best = 0.0;
  int best_c = 1 << 20;  // candidate ordinal; "none" sorts last
  if (tid < 72) {
    const int i = -180 + 5 * tid;
    double cur = 0.0;
    int left = (i + 180) - a.half_hfov;
    if (left < 0) {
      for (int j = 360 + left; j < 360; ++j) cur = dadd(cur, vg[j]);
      left = 0;
    }
    int right = (i + 180) + a.half_hfov;
    if (right >= 360) {
      for (int j = 0; j < right - 359; ++j) cur = dadd(cur, vg[j]);
      right = 359;
    }
    for (int j = left; j <= right; ++j) cur = dadd(cur, vg[j]);
    if (cur > best) { best = cur; best_c = tid; }  // strict, against the initial 0.0 (rrt.cpp:449)
  }
  // arg-max, ties -> lowest ordinal (= the first strict maximum of the sequential scan)
  if (warp < 3) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
      if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
    }
    if (lane == 0) { s_wbest[warp] = best; s_wc[warp] = best_c; }
  }
  __syncthreads();
  if (warp != 0) return;
  best = s_wbest[0];
  best_c = s_wc[0];
#pragma unroll
  for (int w = 1; w < 3; ++w) {
    const double ob = s_wbest[w];
    const int oc = s_wc[w];
    if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
  }
  const int yaw = (best_c < 72) ? (-180 + 5 * best_c) : 0;  // max_gain_yaw starts at 0 (rrt.cpp:450)
  if (a.counts) {
    unsigned long long wu = 0, wf = 0, wo = 0;
    for (int d = -a.half_hfov + lane; d <= a.half_hfov; d += 32) {
      const int j = (((yaw + 180 + d) % 360) + 360) % 360;
      wu += __ldcg(c + 3 * j); wf += __ldcg(c + 3 * j + 1); wo += __ldcg(c + 3 * j + 2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wu += __shfl_xor_sync(0xffffffffu, wu, o);
      wf += __shfl_xor_sync(0xffffffffu, wf, o);
      wo += __shfl_xor_sync(0xffffffffu, wo, o);
    }
    if (lane == 0) { a.counts[3 * v] = wu; a.counts[3 * v + 1] = wf; a.counts[3 * v + 2] = wo; }
  }
  if (lane == 0) {
    a.gain[v] = dmul(best, a.cube);  // rrt.cpp:478
    if (a.yaw_deg) a.yaw_deg[v] = yaw;
    publish_done(a.done, a.seq);
  }
}

// One CTA per candidate.
__global__ void __launch_bounds__(128) k_best_yaw(const BestYawArgs a) {
  __shared__ double vg[360];
  best_yaw_cta(a, (int)blockIdx.x, vg, (int)threadIdx.x, (int)blockDim.x);
}

// Fused finish of k_gain<..., FUSE = true>: every CTA of candidate v arrives here after its per-yaw counts are written;
// the last one runs the window scan (the threadFenceReduction pattern: fence, count, fence, read through L2) and resets
// the counter for the next launch.  `scratch`: 360 doubles of shared memory nobody uses any more.
__device__ __forceinline__ void fused_best_yaw(const GainArgs& a, int v, int chunks, int tid, double* scratch) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.arrive + v, 1) == chunks - 1) ? 1 : 0;
  __syncthreads();
  if (s_last) {  // CTA-uniform
    __threadfence();
    if (tid == 0) a.arrive[v] = 0;
    best_yaw_cta(a.by, v, scratch, tid, (int)blockDim.x);
  }
}
constexpr size_t kFusedScratchBytes = 360 * sizeof(double);  // the launch reserves at least this much dynamic shared memory

// k_gain: the ray sweep described in the comment block above the MODE enum.
// THREADS: 256 (4 CTAs per SM) everywhere except for dense volumes that need more than a quarter of an SM's shared
// memory: those run as 512-thread CTAs, 2 per SM -- the same 32 warps per SM and 64 registers per thread.
template <int VARIANT, int LV, int MODE, bool FUSE = false, int THREADS = kGainThreads>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_gain(const GainArgs a) {
  constexpr bool kFilter = (VARIANT != 1);
  constexpr bool kStatusBits = (VARIANT == 0 || VARIANT == 2);
  constexpr bool kParanoid = (VARIANT == 2);
  constexpr bool kIsDense = (MODE == kDense || MODE == kDenseClip);
  constexpr bool kZClip = (MODE == kDenseClip);
  // The stretch loops of the dense kernel step the volume's bit address P only; the voxel index g is rebuilt from P when a
  // stretch ends (the general loop needs it).  The cross-checking variant and the other modes carry g all the way.
  constexpr bool kKeepG = !kIsDense || kParanoid;
  static_assert(!kIsDense || (LV == 4 && kStatusBits), "the dense volume copies 16-voxel status-tile words");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int yc = a.yc;
  const int gn = (MODE == kProbe) ? 0 : a.gnx * a.gny * a.gnz;
  uint32_t* const vol = reinterpret_cast<uint32_t*>(smem_raw);
  const GainSmem S = carve(smem_raw + (kIsDense ? (((size_t)a.vwords * 4 + 15) & ~(size_t)15) : 0), yc, gn);
  __shared__ int s_umax, s_next;
  __shared__ unsigned long long s_steps;
  __shared__ unsigned int s_mismatch;
  __shared__ uint32_t s_outm[kMaxDenseWordsX];

  const int chunks = 360 / yc;
  const int v = blockIdx.x / chunks;
  const int chunk = blockIdx.x - v * chunks;
  const int tid = threadIdx.x;
  const int lv = LV ? LV : a.L.log2vps;
  const int vm = (1 << lv) - 1;
  const unsigned nvox = LV ? (1u << (3 * LV)) : (unsigned)a.L.nvox;

  const bool by_value = FUSE && a.pos == nullptr;  // single candidate, position in the launch arguments
  const double px = by_value ? a.pos1[0] : a.pos[3 * v], py = by_value ? a.pos1[1] : a.pos[3 * v + 1],
               pz = by_value ? a.pos1[2] : a.pos[3 * v + 2];
  // Guard (the host entry points validate; device-resident inputs are caught here).
  const double lim = 4194304.0 - (double)a.reach_vox - 2.0;
  const bool bad = !(fabs(px * a.vs_inv) < lim) || !(fabs(py * a.vs_inv) < lim) || !(fabs(pz * a.vs_inv) < lim) ||
                   (a.valid != nullptr && a.valid[v] == 0);
  if (tid == 0) {
    s_umax = 0;
    s_next = 0;
    s_steps = 0ull;
    s_mismatch = 0u;
  }
  if (bad) {
    for (int j = tid; j < yc; j += blockDim.x)
      for (int k = 0; k < 3; ++k) a.per_yaw[((size_t)v * 360 + chunk * yc + j) * 3 + k] = 0u;
    if (FUSE) fused_best_yaw(a, v, chunks, tid, reinterpret_cast<double*>(smem_raw));
    return;
  }
  // Voxel containing the candidate position and the origin of the local block grid.
  const int g0x = (int)floor(px * a.vs_inv), g0y = (int)floor(py * a.vs_inv), g0z = (int)floor(pz * a.vs_inv);
  const int b0x = (g0x - a.reach_vox) >> lv, b0y = (g0y - a.reach_vox) >> lv, b0z = (g0z + a.gz_lo) >> lv;

  // (a) block-slot grid: one hash probe per reachable block; (b) per-yaw frames
  for (int i = tid; i < gn; i += blockDim.x) {
    const int ix = i % a.gnx, iy = (i / a.gnx) % a.gny, iz = i / (a.gnx * a.gny);
    const int sl = probe_block(a.L, b0x + ix, b0y + iy, b0z + iz);
    S.grid[i] = sl < 0 ? a.unknown_slot : sl;
  }
  // dense volume: origin along x (global voxel index aligned down to a block row) and, per 16-voxel word,
  // the mask of voxels outside the exploration box along x
  const int dense_x0 = kIsDense ? (((g0x + a.vmin[5 * chunk]) >> 4) << 4) : 0;
  if (kIsDense && tid < kMaxDenseWordsX) {
    const int lo = a.gmin[0] - (dense_x0 + 16 * tid);  // in-box voxels of word `tid`: lo .. lo + gspan
    const long long hi = (long long)lo + (long long)a.gspan[0];
    uint32_t om = 0xFFFFFFFFu;
    if (hi >= 0 && lo <= 15) {
      const int l = lo > 0 ? lo : 0, h = hi < 15 ? (int)hi : 15;
      const uint32_t below = (1u << (2 * l)) - 1u;                                  // voxels < l
      const uint32_t upto = h == 15 ? 0xFFFFFFFFu : ((1u << (2 * (h + 1))) - 1u);  // voxels <= h
      om = below | ~upto;
    }
    s_outm[tid] = om;
  }
  for (int i = tid; i < 6 * yc; i += blockDim.x) {
    const int k = i / yc, j = i - k * yc, y = chunk * yc + j;
    build_yaw_plane(a.A + (size_t)y * 24, a.B + 3 * y, k, S.pl, S.fp, S.sl, a.vs, g0x, g0y, g0z, yc, j, px, py, pz);
  }
  for (int j = (int)blockDim.x - 1 - tid; j < yc; j += blockDim.x) {  // from the far end: those warps have less plane work
    const int y = chunk * yc + j;
    build_yaw_rays(a.A + (size_t)y * 24, a.B + 3 * y, a.cs + 2 * y, a.vs_inv, S.rd, S.st, S.csf, S.um, S.cnt, yc, j, px, py, pz);
  }
  __syncthreads();
  for (int j = tid; j < yc; j += blockDim.x) atomicMax(&s_umax, S.um[j]);

  // (c) dense status volume, filled strip by strip (see below)
  int x0 = 0, y0 = 0, z0 = 0, vxw = 0, vey = 0, nvol = 0;
  if (kIsDense) {
    y0 = g0y + a.vmin[5 * chunk + 1];
    z0 = g0z + a.vmin[5 * chunk + 2];
    x0 = dense_x0;
    vxw = a.vmin[5 * chunk + 3];
    vey = a.vmin[5 * chunk + 4];
    int vez_c = a.vez;
    if (kZClip) {
      // rows that can be fetched: while a ray is inside the slab [gmin_z, gmax_z] (the walk stops at t_out, at most 2 voxels
      // past it: margin 3), and between its start row and the slab when it starts outside
      const int z_lo = max(z0, min(a.gmin[2] - 3, g0z + a.sz_lo));
      const int z_hi = min(z0 + a.vez - 1, max(a.gmin[2] + (int)a.gspan[2] + 3, g0z + a.sz_hi));
      z0 = z_lo;
      vez_c = max(z_hi - z_lo + 1, 1);
      if (vez_c > a.vez_alloc) {  // cannot happen: the host sized the memory from the candidates' own z-range
        if (tid == 0) atomicAdd(a.totals + 3, 1ull);
        vez_c = a.vez_alloc;
      }
    }
    nvol = vxw * vey * vez_c;
    // One item = one x word of up to 16 consecutive y rows that share a block (a "strip"): the box mask of the word, the z test and
    // the block's tile slot are looked up once per strip, a row then costs one load of a status-tile word, four logic operations
    // and one store (the loads of four rows are issued before any is used).
    const int yb0 = y0 >> 4;
    const int nyb = ((y0 + vey - 1) >> 4) - yb0 + 1;
    const int nitems = vxw * nyb * vez_c;
    for (int it = tid; it < nitems; it += blockDim.x) {
      const int q = it / vxw, wx = it - q * vxw;
      const int rz = q / nyb, yb = q - rz * nyb;
      const int gz = z0 + rz;
      const int ylo = max(y0, (yb0 + yb) << 4), yhi = min(y0 + vey - 1, ((yb0 + yb) << 4) + 15);
      const uint32_t om = s_outm[wx];  // voxels of this word outside the exploration box along x
      const bool use_col = (unsigned)(gz - a.gmin[2]) <= a.gspan[2] && om != 0xFFFFFFFFu;
      int slot = a.unknown_slot;
      if (use_col) {
        const int ix = (x0 >> 4) - b0x + wx, iy = yb0 + yb - b0y, iz = (gz >> 4) - b0z;
        if ((unsigned)ix < (unsigned)a.gnx && (unsigned)iy < (unsigned)a.gny && (unsigned)iz < (unsigned)a.gnz) {
          slot = S.grid[(iz * a.gny + iy) * a.gnx + ix];
        } else {
          slot = probe_block(a.L, (x0 >> 4) + wx, yb0 + yb, gz >> 4);
          if (slot < 0) slot = a.unknown_slot;
        }
      }
      const uint32_t* const src = a.L.status + ((unsigned)slot * 256u + (unsigned)((gz & 15) << 4));
      uint32_t* dst = vol + ((rz * vey + (ylo - y0)) * vxw + wx);
      const int ylo_in = max(ylo, a.gmin[1]);  // rows of the strip inside the box along y: ylo_in .. yhi_in
      const long long yhi_box = (long long)a.gmin[1] + (long long)a.gspan[1];
      const int yhi_in = yhi_box < (long long)yhi ? (int)yhi_box : yhi;
      for (int gy = ylo; gy <= yhi; gy += 4) {
        uint32_t word[4];
        bool use[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          use[r] = use_col && gy + r >= ylo_in && gy + r <= yhi_in;
          word[r] = use[r] ? __ldg(src + ((gy + r) & 15)) : 0u;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uint32_t w = word[r];
          w |= (w >> 1) & 0x55555555u;  // tile status 2 (occupied) -> code 3
          w = use[r] ? ((w & ~om) | (om & 0xAAAAAAAAu)) : 0xAAAAAAAAu;  // "skip" outside the box
          if (gy + r <= yhi) dst[r * vxw] = w;
        }
        dst += 4 * vxw;
      }
    }
  }
  __syncthreads();
  const int n_rays = s_umax * yc;
  const int gbase = -((b0z * a.gny + b0y) * a.gnx + b0x);  // grid index of block (bx,by,bz) = (bz*gny + by)*gnx + bx + gbase
  // slot grid through its shared-window address (with gbase folded in), kept opaque like vol_s below: otherwise the
  // window base is re-derived with four uniform-pipe instructions at every voxel
  uint32_t grid_s = (uint32_t)__cvta_generic_to_shared(S.grid) + 4u * (unsigned)gbase;
  asm volatile("mov.u32 %0, %0;" : "+r"(grid_s));

  // Offset of the candidate inside its voxel (classifier works relative to g0).
  const float noffx = -__double2float_rn(px - (double)g0x * a.vs);
  const float noffy = -__double2float_rn(py - (double)g0y * a.vs);
  const float noffz = -__double2float_rn(pz - (double)g0z * a.vs);
  const FilterConst& F = a.F;
  const int stride_y = 32 * vxw, stride_z = 32 * vxw * vey;
  const unsigned mg_xw = vxw > 1 ? 0xFFFFFFFFu / (unsigned)vxw + 1u : 0u, mg_ey = vey > 1 ? 0xFFFFFFFFu / (unsigned)vey + 1u : 0u;
  uint32_t vol_s = (uint32_t)__cvta_generic_to_shared(vol);
  asm volatile("mov.u32 %0, %0;" : "+r"(vol_s));  // opaque: keep the shared-window address in a register

  const int lane = tid & 31;
  unsigned long long my_steps = 0ull;
  unsigned int my_mismatch = 0u;

  // Batches of 32 rays are handed out dynamically (rows differ in ray length by up to 2x, and 8 warps rarely
  // divide the batch count): a warp takes the next batch when it is done with its last.
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&s_next, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n_rays) break;
    // top rows first: with a floor-bounded exploration box the downward rows leave the box early, so the
    // short batches come last and the warps of a CTA finish closer together (order only; any order is exact)
    const int r = (((n_rays + 31) >> 5) << 5) - 32 - base + lane;
    // (32 consecutive rows of one yaw per warp instead -- rays in one vertical plane -- were measured: 0.51 against 0.68 of the
    // roofline on config 3, the lanes then differ in z and collide in the volume's banks)
    const int u = r / yc;
    const int j = r - u * yc;
    // The batch loop is warp-uniform; lanes without a ray (tail of the last batch, rows beyond this yaw's
    // u_max) run the set-up on in-range table entries and never walk, so that the warp-level
    // synchronisation between the stretches of the walk below always sees all 32 lanes.
    const bool active = r < n_rays && u < S.um[j];

    // ---- ray end point (rrt.cpp:396,410-411) and RayCaster setup (integrator_utils.cc:127-179)
    int g[3], sg[3];
    float t[3], dt[3];
    int len = 0;
    float tlo = 0.0f, thi = 3.0e38f, tlo_v = 0.0f, t_out = 3.0e38f;
    float e3[3];
    int adl[3];  // |end voxel - start voxel| per axis
    // Modes that test the exploration box per voxel (the dense volume has it baked in): a ray whose start and end voxels
    // both lie at least two voxels inside the box never needs the test.  The walk moves axis k only towards the end voxel
    // and takes len = sum |end - start| steps; it can overshoot an axis by one voxel at most (a second extra crossing
    // lies 1 / |r_k| later in t, far beyond the accumulated rounding of the t's for the ranges admitted here), so every
    // voxel it emits is within one voxel of the start / end bounding box.
    bool ray_in = !kIsDense && a.reach_vox <= 256;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float sk = S.st[k * yc + j];
      const double pos = dadd(dadd(S.rd[k * yc + j], dmul(dmul((double)u, S.rd[(3 + k) * yc + j]), a.vs)), S.rd[(6 + k) * yc + j]);
      const float ek = __double2float_rn(dmul(pos, a.vs_inv));
      e3[k] = ek;
      g[k] = (int)floorf(__fadd_rn(sk, 1e-6f));
      const int ei = (int)floorf(__fadd_rn(ek, 1e-6f));
      adl[k] = abs(ei - g[k]);
      len += adl[k];
      if (!kIsDense)
        ray_in = ray_in && a.gspan[k] >= 4u && (unsigned)(g[k] - a.gmin[k] - 2) <= a.gspan[k] - 4u &&
                 (unsigned)(ei - a.gmin[k] - 2) <= a.gspan[k] - 4u;
      const float rr = __fsub_rn(ek, sk);
      sg[k] = (rr == 0.0f) ? 0 : (rr < 0.0f ? -1 : 1);
      const float shifted = __fsub_rn(sk, (float)g[k]);
      const float dist = __fsub_rn((float)max(0, sg[k]), shifted);
      t[k] = __fdiv_rn(dist, rr);  // exact IEEE division (+-inf / NaN when rr == 0): the reference's arithmetic
      dt[k] = __fdiv_rn((float)sg[k], rr);
      if (kFilter) {
        // g is surely outside the (convex) exploration box once s + t*r has passed a bound by 1.02 voxels in
        // the direction of travel.  Conservative bounds may use the fast division: padded by 1e-4 relative.
        const float bmin = (float)a.gmin[k], bmax = (float)a.gmin[k] + (float)a.gspan[k];
        if (rr != 0.0f) t_out = fminf(t_out, __fdividef(rr > 0.0f ? (bmax + 1.02f - sk) : (bmin - 0.02f - sk), rr));
      }
    }
    float pa[3] = {0.f, 0.f, 0.f}, pc = 0.f;  // the open one of top/bottom, single precision, relative to g0
    if (kFilter) {
      float xe, ye, ze;
      to_camera(F, a.vs_f, e3[0] - (float)g0x, e3[1] - (float)g0y, e3[2] - (float)g0z, noffx, noffy, noffz, S.csf[j],
                S.csf[yc + j], xe, ye, ze);
      // Signed distances of the ray point to the six planes grow linearly from ~0 at the camera to
      // these values at t = 1 (unnormalised for the side planes): near/far use xe, "left" (plane 2)
      // th*x + y, "right" (3) th*x - y, top (4) tv*x - z, bottom (5) tv*x + z.
      const float dL = fmaf(F.th, xe, ye), dR = fmaf(F.th, xe, -ye);
      const float dT = fmaf(F.tv, xe, -ze), dB = fmaf(F.tv, xe, ze);
      // classifier margin + float slop of the start point + 0.01 voxel for the drift of the DDA's own
      // float t against the exact crossing parameter (<= ~5e-4 voxel over a ray)
      const float m1 = F.m + 1e-4f + 0.01f * a.vs_f;
      const float nh = __fdividef(F.mh, F.m), nv = __fdividef(F.mv, F.m);  // sqrt(1+th^2), sqrt(1+tv^2)
      const bool up = ze >= 0.0f;  // the ray ends in the upper half: top is the nearer plane
      const float d_nearv = up ? dT : dB, d_farv = up ? dB : dT;
      const float sl_nearv = S.sl[(up ? 4 : 5) * yc + j], sl_farv = S.sl[(up ? 5 : 4) * yc + j];
      if (xe > 0.0f && dL > 0.0f && dR > 0.0f && d_farv > 0.0f) {
        const float rx = __fdividef(1.0f, xe);
        tlo = fmaxf(tlo, (F.nearp + S.sl[0 * yc + j] + m1) * rx);
        tlo = fmaxf(tlo, __fdividef((S.sl[2 * yc + j] + m1) * nh, dL));
        tlo = fmaxf(tlo, __fdividef((S.sl[3 * yc + j] + m1) * nh, dR));
        tlo = fmaxf(tlo, __fdividef((sl_farv + m1) * nv, d_farv));
        thi = (F.farp - S.sl[1 * yc + j] - m1) * rx;
        tlo_v = d_nearv > 0.0f ? __fdividef((sl_nearv + m1) * nv, d_nearv) : 3.0e38f;
      } else {
        tlo = 3.0e38f;  // degenerate frustum geometry for this ray: always use the classifier
      }
      tlo = fmaf(fabsf(tlo), 1e-4f, tlo) + 1e-5f;
      tlo_v = fmaf(fabsf(tlo_v), 1e-4f, tlo_v) + 1e-5f;
      thi = fmaf(fabsf(thi), -1e-4f, thi) - 1e-5f;
      t_out = fmaf(fabsf(t_out), 1e-4f, t_out) + 1e-5f;
      const int q = up ? 4 : 8;
      pa[0] = S.fp[(q + 0) * yc + j];
      pa[1] = S.fp[(q + 1) * yc + j];
      pa[2] = S.fp[(q + 2) * yc + j];
      pc = S.fp[(q + 3) * yc + j];
    }
    const float tA = fmaxf(tlo, tlo_v);
    float tB = fminf(thi, t_out);  // (tA, tB): the window in which every test is sure
    // dense volume: bit address of the start voxel (the shared-window byte address of the volume is folded
    // in: word address (P >> 3) & ~3, bit position P & 31; vol_s is 16-byte aligned) and per-axis strides
    unsigned P = 0u;
    int stx = 0, sty = 0, stz = 0;
    float dummy_sv = 0.0f;
    if (kIsDense) {
      P = (unsigned)(((g[2] - z0) * vey + (g[1] - y0)) * stride_y + 2 * (g[0] - x0)) + 8u * vol_s;
      stx = 2 * sg[0];
      sty = stride_y * sg[1];
      stz = stride_z * sg[2];
      // opaque: keep the three strides in registers (ptxas otherwise re-derives them from the step signs
      // inside the loops, on the integer pipe that bounds the kernel)
      asm volatile("" : "+r"(stx), "+r"(sty), "+r"(stz));
    }

    // Per-ray counts by power sums: every walked voxel adds its code c in {0 unknown, 1 free, 2 skip} to
    // S1 = sum c and S2 = sum c^2; with n_acc counted voxels, n_skip = (S2 - S1) / 2, n_free = S1 - 2 n_skip,
    // n_unknown = n_acc - n_free - n_skip.  The high half of S1 counts the voxels walked (the RayCaster's
    // step counter, integrator_utils.cc:111-117), so one add and one compare per step serve both.
    unsigned S1 = 0u, S2 = 0u, occ = 0u;
    const unsigned walk_end = (unsigned)(len + 1) << 16;  // the ray emits len + 1 voxels
    int skipped = 0;  // voxels past t_out: accounted as executed, never looked at
    int cbx = INT_MIN, cby = 0, cbz = 0, cslot = a.unknown_slot;
    float tcur = 0.0f;
    bool run = active;

    // Exploration box on integer thresholds (exact); folded into the volume in dense mode.
    auto in_box = [&]() -> bool {
      return ray_in || ((unsigned)(g[0] - a.gmin[0]) <= a.gspan[0] && (unsigned)(g[1] - a.gmin[1]) <= a.gspan[1] &&
                        (unsigned)(g[2] - a.gmin[2]) <= a.gspan[2]);
    };
    // Voxel code of the current voxel as the map has it: 0 unknown, 1 free, 3 occupied (voxblox_manager.cpp:14-33);
    // the dense volume also answers 2 for voxels outside the exploration box.
    auto tile_code = [&]() -> unsigned {
      const int bx = g[0] >> lv, by = g[1] >> lv, bz = g[2] >> lv;
      int slot;
      if (MODE == kGrid) {
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(slot) : "r"(grid_s + 4u * (unsigned)((bz * a.gny + by) * a.gnx + bx)));
      } else if (MODE == kProbe) {
        if (bx != cbx || by != cby || bz != cbz) {
          cbx = bx; cby = by; cbz = bz;
          cslot = probe_block(a.L, bx, by, bz);
          if (cslot < 0) cslot = a.unknown_slot;  // all-unknown tile
        }
        slot = cslot;
      } else {  // dense mode asks the tiles only to cross-check the volume (variant 2)
        slot = probe_block(a.L, bx, by, bz);
        if (slot < 0) slot = a.unknown_slot;
      }
      const unsigned lin = (unsigned)((g[0] & vm) | ((g[1] & vm) << lv) | ((g[2] & vm) << (2 * lv)));
      unsigned st;
      if (kStatusBits) {
        st = (__ldg(a.L.status + ((unsigned)slot * (nvox >> 4) + (lin >> 4))) >> ((lin & 15u) * 2u)) & 3u;
      } else {
        const float2 vx = __ldg(reinterpret_cast<const float2*>(a.L.tiles) + ((size_t)slot * nvox + lin));
        st = tsdf_status(vx.x, vx.y);
      }
      return st | (st >> 1);  // 2 (occupied) -> 3
    };
    auto volume_code = [&]() -> unsigned {
      if (kParanoid && tcur > t_out) {
        // the default kernel has stopped here; this variant walks on to check that nothing inside the box follows, and
        // must not read the (possibly z-clipped) volume for voxels the default kernel never fetches
        if (in_box()) ++my_mismatch;
        return 2u;
      }
      uint32_t w;  // one word of the dense volume (always in range)
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"((P >> 3) & ~3u));
      const unsigned code = __funnelshift_r(w, 0u, P) & 3u;
      if (kParanoid) {
        const unsigned want = in_box() ? tile_code() : 2u;
        if (((P - 8u * vol_s) >> 5) >= (unsigned)nvol || want != code) ++my_mismatch;
      }
      return code;
    };

    // Voxel index from the bit address (dense kernel; once per stretch, two small integer divisions).
    auto g_from_P = [&]() {
      const unsigned rel = P - 8u * vol_s;
      // divisions by the chunk's row length / row count as multiplications by ceil(2^32 / d): exact while n * d < 2^32
      // (word index < 2^24 and at most 32 words per row; row index * rows per slice < words * rows, checked on the host)
      const unsigned wd = rel >> 5, row = __umulhi(wd, mg_xw), zz = vey > 1 ? __umulhi(row, mg_ey) : row;
      g[0] = x0 + (int)((wd - row * (unsigned)vxw) * 16u + ((rel & 31u) >> 1));
      g[1] = y0 + (int)(row - zz * (unsigned)vey);
      g[2] = z0 + (int)zz;
    };
    // count_safe -- why the stretch loops need not test the RayCaster's step counter per voxel.  The walk ends after len + 1
    // emissions, len = sum_k |end_k - start_k| (integrator_utils.cc:111-117,141-147).  Axis k crosses its j-th boundary at
    // t_k(0) + j dt_k; its |delta_k|-th crossing and every later one lies beyond the end voxel, i.e. at t >= 1 -- except that
    // the 1e-6 voxel nudge of getGridIndexFromPoint can hide one crossing from len, at t >= 1 - 1e-6 / |r_k|.  So while
    // tcur < tB every step taken so far was one of the len counted ones and voxel number len has not been left, provided
    // (a) 1 - tB exceeds the drift of the accumulated float t's (<= len * 2^-24 ~ 1.5e-5 at 250 steps; 1 - tB >= 1.8e-4 from
    // the far-plane margin), (b) no axis has 0 < |r_k| < 2e-6 / (1 - tB) -- a ray that fails it, almost parallel to a grid
    // plane, walks all its voxels in the general loop, which counts every step -- and (c) below.
    if (kFilter) {
      const float room = 1.0f - tB;
      // (the dense window loop's packed accumulator holds 255 voxels; elsewhere the bound keeps the drift of note (a) at 6e-5)
      bool risky = !(room > 1.0e-4f) || len > ((kIsDense && !kParanoid) ? 250 : 1000);
      bool zero_axis = false;
      float dmin = 3.0e38f, tstop = -1.0f;  // the densest stepping axis and the entry time of its third-last counted crossing
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float ar = fabsf(dt[k]);  // = 1 / |r_k| (NaN for r_k = 0: such an axis never moves)
        risky = risky || (sg[k] != 0 && ar * 2.0e-6f > room);
        zero_axis = zero_axis || sg[k] == 0;
        if (sg[k] != 0 && ar < dmin) {
          dmin = ar;
          tstop = adl[k] >= 4 ? fmaf((float)adl[k] - 3.5f, dt[k], t[k]) : -1.0f;
        }
      }
      // (c) an axis WITHOUT a component (the integer yaws 0 / 180: no y; rows without z) starts at t = -inf when the start
      // point is not on a voxel face: the RayCaster "steps" it once, by zero voxels, before it turns NaN for good
      // (integrator_utils.cc:165-179), so such a ray emits its start voxel once more and ends up to two voxels short of its
      // end voxel -- its counter can run out while tcur < tB.  Its stretches therefore end before the last three crossings
      // of its densest axis; the general loop, which counts, walks the rest.
      if (zero_axis) tB = fminf(tB, tstop);
      if (risky) tlo = 3.0e38f;
    }

    // A ray passes through up to three kinds of stretches, in this order: the general one (ray start; again at
    // the very end, where the far plane is open), a stretch where only the nearer of top/bottom is open (whole
    // rays in the rows that hug those planes), and the window in which every plane is sure.  One loop per
    // kind, so that the lanes of a warp -- neighbouring yaws of the same row, with near-equal boundaries --
    // run the same kind together instead of dragging each other through the general code: the warp is
    // re-converged between the loops, and lanes that finished a stretch early wait for the others there.
    while (__any_sync(0xffffffffu, run)) {
      // ---- general loop (camera_model.cpp:194-199 decided cheapest first)
      while (run && (!kFilter || !(tcur > tlo && tcur < tB))) {
        const bool inbox = (!kIsDense || kParanoid) ? in_box() : true;
        bool fin = true;
        if (!kFilter) {
          fin = inbox && exact_in_frustum(S.pl, yc, j, g[0], g[1], g[2], a.vs);
        } else {
          if (tcur > t_out) {  // left the convex exploration box for good
            if (!kParanoid) {
              skipped = len + 1 - (int)(S1 >> 16);  // the rest is accounted as executed but contributes nothing
              run = false;
              break;
            }
            if (inbox) ++my_mismatch;
          }
          bool decided = false;
          fin = false;
          if (kIsDense || inbox) {
            if (tcur > tlo) {
              // near, left/right and the farther of top/bottom are sure; open: the far plane once tcur >= thi,
              // the nearer of top/bottom while tcur <= tlo_v -> one single-precision plane evaluation each
              const float dx = (float)(g[0] - g0x), dy = (float)(g[1] - g0y), dz = (float)(g[2] - g0z);
              float sf = 3.0e38f;
              if (!(tcur < thi)) sf = fmaf(S.fp[0 * yc + j], dx, fmaf(S.fp[1 * yc + j], dy, fmaf(S.fp[2 * yc + j], dz, S.fp[3 * yc + j])));
              const float sv = fmaf(pa[0], dx, fmaf(pa[1], dy, fmaf(pa[2], dz, pc)));
              // the smaller of the open planes' signed distances decides: > m in, < -m out, else exact test
              const float smin = fminf(sf, tcur > tlo_v ? 3.0e38f : sv);
              fin = smin > F.m;
              decided = fin || smin < -F.m;
            }
            if (!decided || kParanoid) {
              float xc, yk, zk;
              to_camera(F, a.vs_f, (float)(g[0] - g0x), (float)(g[1] - g0y), (float)(g[2] - g0z), noffx, noffy, noffz, S.csf[j],
                        S.csf[yc + j], xc, yk, zk);
              const float s0 = xc - F.nearp, s1 = F.farp - xc;
              const float sh = fmaf(F.th, xc, -fabsf(yk)), sv = fmaf(F.tv, xc, -fabsf(zk));
              const bool sure_in = s0 > F.m && s1 > F.m && sh > F.mh && sv > F.mv;
              const bool sure_out = s0 < -F.m || s1 < -F.m || sh < -F.mh || sv < -F.mv;
              bool ex;
              if (kParanoid) {
                ex = exact_in_frustum(S.pl, yc, j, g[0], g[1], g[2], a.vs);
                if ((sure_in && !ex) || (sure_out && ex) || (decided && fin != ex)) ++my_mismatch;
              } else if (sure_in) {
                ex = true;
              } else if (sure_out) {
                ex = false;
              } else {
                ex = exact_in_frustum(S.pl, yc, j, g[0], g[1], g[2], a.vs);
              }
              fin = ex;
            }
          }
          if (!kIsDense) fin = fin && inbox;
        }
        unsigned code = 2u;
        if (kIsDense) {
          code = volume_code();
          code = fin ? code : 2u;
        } else if (fin) {
          code = tile_code();
        }
        if (code == 3u) {
          occ = 1u;
          S1 += 0x10000u;  // this step was executed
          run = false;
          break;
        }
        if (tcur != tcur) {
          // t[0] is NaN: the ray has no x component (integer yaws +-90 deg), its first x "crossing" was at
          // -inf (or 0/0) and grew by dt = 0/0.  minCoeff then picks axis 0 for ever, whose step is 0: the
          // RayCaster emits this same voxel for all remaining steps (integrator_utils.cc:119-123,165-179).
          const unsigned rem = (unsigned)(len + 1) - (S1 >> 16);
          S1 += rem * (code + 0x10000u);
          S2 += rem * code * code;
          run = false;
          break;
        }
        S1 += code + 0x10000u;
        S2 = code * code + S2;
        if (kIsDense) dda_advance<true, false>(t, dt, tcur, g, sg, P, stx, sty, stz, dummy_sv, 0.f, 0.f, 0.f);
        else dda_advance<false, false>(t, dt, tcur, g, sg, P, 0, 0, 0, dummy_sv, 0.f, 0.f, 0.f);
        run = S1 < walk_end;
      }
      if (!kFilter) continue;
      __syncwarp();
      // ---- only the nearer of top/bottom is open (tlo < tcur <= tlo_v, before thi and t_out): its signed
      //      distance is evaluated once on entry and then follows the walk with one add per step.  The n
      //      roundings of that sum are covered by widening the decision margin (|sv| <= reach + 1 m, so each is
      //      below 2^-24 * 1e5 * F.m); inside the margin the exact planes decide.
      //      Neither this loop nor the window loop tests the RayCaster's step counter per voxel: while tcur < tB (< thi < 1,
      //      the far plane cuts the ray's last stretch off) the ray cannot have emitted its len + 1 voxels -- see the
      //      note at `count_safe` above; the counter is compared once when the loop is left.
      if (run && tcur < tB && !(tcur > tlo_v)) {
        float sv = fmaf(pa[0], (float)(g[0] - g0x), fmaf(pa[1], (float)(g[1] - g0y), fmaf(pa[2], (float)(g[2] - g0z), pc)));
        const float d0 = pa[0] * (float)sg[0], d1 = pa[1] * (float)sg[1], d2 = pa[2] * (float)sg[2];
        const float mi = F.m * (1.0f + 0.01f * (float)(len + 1));
        unsigned last;
        do {
          bool fin = sv > 0.0f;
          if (!(fabsf(sv) > mi)) {  // inside the margin (rare): the exact planes decide
            if (!kKeepG) g_from_P();
            fin = exact_in_frustum(S.pl, yc, j, g[0], g[1], g[2], a.vs);
          }
          if (kParanoid && fin != exact_in_frustum(S.pl, yc, j, g[0], g[1], g[2], a.vs) && (kIsDense || in_box())) ++my_mismatch;
          unsigned c = 2u;
          if (kIsDense) {
            c = volume_code();
            c = fin ? c : 2u;
          } else if (fin && in_box()) {
            c = tile_code();
          }
          // an occupied voxel ends the walk; it is accumulated and stepped past like any other (one exit test per voxel,
          // at the bottom) and taken out again below
          S1 += c + 0x10000u;
          S2 = c * c + S2;
          last = c;
          if (kIsDense) dda_advance<true, true, kKeepG>(t, dt, tcur, g, sg, P, stx, sty, stz, sv, d0, d1, d2);
          else dda_advance<false, true>(t, dt, tcur, g, sg, P, 0, 0, 0, sv, d0, d1, d2);
          if (kParanoid && last != 3u && !(S1 < walk_end) && tcur < tB && !(tcur > tlo_v)) ++my_mismatch;  // the counter would have ended the walk here
        } while (last != 3u && (!kParanoid || S1 < walk_end) && tcur < tB && !(tcur > tlo_v));
        if (last == 3u) {
          occ = 1u;
          S1 -= 3u;  // the step was executed, the voxel counts as occupied only
          S2 -= 9u;
          run = false;
        } else {
          run = S1 < walk_end;
          if (!kKeepG) g_from_P();
        }
      }
      __syncwarp();
      // ---- window: every plane is sure, the voxel counts unless it is outside the exploration box
      if (kIsDense && !kParanoid) {
        // The dense kernel's window loop, 74 % of all voxel-steps, written out as PTX so that it stays at the 22 instructions
        // it needs per voxel: word address from the bit address P (2), load, first-minimum select of the RayCaster's t (6: a
        // NaN in t[1] / t[2] never wins, one in t[0] always does -- Eigen's minCoeff), code = (word >> (P & 31)) & 3 (2), the
        // stepping axis' t, and P by the axis' stride (6), S1 += code + 2^16 and S2 += code^2 (2), and ONE exit test: stop on
        // an occupied voxel (code 3) or when the next voxel is entered at or beyond tB (3).  An occupied voxel is accumulated
        // and stepped past like any other and taken out again below.
        if (run && tcur > tA && tcur < tB) {
          // acc += (code + 1024)^2 = 2^20 + 2048 code + code^2 per voxel: voxels walked, sum of codes and sum of squared codes
          // from ONE multiply-add (a stretch is at most 255 voxels: 2048 * 510 + 1020 < 2^20)
          unsigned last, acc = 0u;
          asm volatile(
              "{\n\t"
              ".reg .pred m0, m1, m2, pn, pc;\n\t"
              ".reg .u32 ad, wd;\n\t"
              "NBVW_LOOP:\n\t"
              "shr.u32 ad, %3, 3;\n\t"
              "and.b32 ad, ad, 0xFFFFFFFC;\n\t"
              "ld.shared.u32 wd, [ad];\n\t"
              "min.f32 %5, %0, %1, %2;\n\t"               // t[0] is never NaN here (such rays end in the general loop)
              "setp.eq.f32 m0, %0, %5;\n\t"               // first minimum: axis 0 on ties, then axis 1
              "setp.eq.and.f32 m1, %1, %5, !m0;\n\t"
              "setp.neu.and.f32 m2, %1, %5, !m0;\n\t"
              "shf.r.wrap.b32 %6, wd, 0, %3;\n\t"
              "lop3.b32 %6, %6, 3, %14, 0xEA;\n\t"        // (word >> (P & 31)) & 3 | 1024
              "@m0 add.rn.f32 %0, %0, %7;\n\t"
              "@m1 add.rn.f32 %1, %1, %8;\n\t"
              "@m2 add.rn.f32 %2, %2, %9;\n\t"
              "selp.b32 ad, %11, %12, m1;\n\t"
              "selp.b32 ad, %10, ad, m0;\n\t"
              "add.s32 %3, %3, ad;\n\t"
              "mad.lo.u32 %4, %6, %6, %4;\n\t"
              "setp.ne.u32 pn, %6, 1027;\n\t"
              "setp.lt.and.f32 pc, %5, %13, pn;\n\t"
              "@pc bra NBVW_LOOP;\n\t"
              "}"
              : "+f"(t[0]), "+f"(t[1]), "+f"(t[2]), "+r"(P), "+r"(acc), "+f"(tcur), "=&r"(last)
              : "f"(dt[0]), "f"(dt[1]), "f"(dt[2]), "r"(stx), "r"(sty), "r"(stz), "f"(tB), "r"(1024u)
              : "memory");
          S1 += ((acc >> 20) << 16) + ((acc >> 11) & 0x1FFu);
          S2 += acc & 0x7FFu;
          if (last == 1027u) {  // the occupied voxel that stopped the ray: executed, not counted as anything else
            occ = 1u;
            S1 -= 3u;
            S2 -= 9u;
            run = false;
          } else {
            run = S1 < walk_end;
            g_from_P();
          }
        }
      } else if (run && tcur > tA && tcur < tB) {
        unsigned last;
        do {
          if (kParanoid && (kIsDense || in_box()) && !exact_in_frustum(S.pl, yc, j, g[0], g[1], g[2], a.vs)) {
            if (in_box()) ++my_mismatch;
          }
          unsigned c = 2u;
          if (kIsDense) {
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"((P >> 3) & ~3u));
            c = __funnelshift_r(w, 0u, P) & 3u;
            if (kParanoid && c != volume_code()) ++my_mismatch;  // volume_code() cross-checks against the tiles
          } else if (in_box()) {
            c = tile_code();
          }
          S1 += c + 0x10000u;  // one three-input add: the code and the walk counter's bit
          S2 = c * c + S2;
          last = c;
          if (kIsDense) dda_advance<true, false, kKeepG>(t, dt, tcur, g, sg, P, stx, sty, stz, dummy_sv, 0.f, 0.f, 0.f);
          else dda_advance<false, false>(t, dt, tcur, g, sg, P, 0, 0, 0, dummy_sv, 0.f, 0.f, 0.f);
          if (kParanoid && last != 3u && !(S1 < walk_end) && tcur < tB) ++my_mismatch;  // the counter would have ended the walk here
        } while (last != 3u && (!kParanoid || S1 < walk_end) && tcur < tB);
        if (last == 3u) {  // the occupied voxel that stopped the ray: executed, not counted as anything else
          occ = 1u;
          S1 -= 3u;
          S2 -= 9u;
          run = false;
        } else {
          run = S1 < walk_end;
          if (!kKeepG) g_from_P();
        }
      }
    }
    const unsigned walked = S1 >> 16;  // includes the occupied voxel that stopped the ray
    S1 &= 0xFFFFu;
    S2 &= 0x1FFFFu;
    my_steps += (unsigned long long)(walked + (unsigned)skipped);
    const unsigned n_acc = walked - occ;
    const unsigned n_skip = (S2 - S1) >> 1, nf = S1 - 2u * n_skip, nu = n_acc - nf - n_skip;
    if (active) {
      if (nu) atomicAdd(&S.cnt[0 * yc + j], nu);
      if (nf) atomicAdd(&S.cnt[1 * yc + j], nf);
      if (occ) atomicAdd(&S.cnt[2 * yc + j], occ);
    }
  }
  // ---- CTA epilogue
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
    my_mismatch += __shfl_xor_sync(0xffffffffu, my_mismatch, o);
  }
  if (lane == 0) {
    atomicAdd(&s_steps, my_steps);
    if (kParanoid) atomicAdd(&s_mismatch, my_mismatch);
  }
  __syncthreads();
  for (int i = tid; i < 3 * yc; i += blockDim.x) {
    const int k = i / yc, j = i - k * yc;
    a.per_yaw[((size_t)v * 360 + chunk * yc + j) * 3 + k] = S.cnt[k * yc + j];
  }
  if (tid == 0) {
    unsigned long long rays = 0ull;
    for (int j = 0; j < yc; ++j) rays += (unsigned long long)S.um[j];
    if (a.steps) atomicAdd(a.steps + v, s_steps);
    atomicAdd(a.totals + 0, rays);
    atomicAdd(a.totals + 1, s_steps);
    if (kParanoid && s_mismatch) atomicAdd(a.totals + 3, (unsigned long long)s_mismatch);
  }
  if (FUSE) fused_best_yaw(a, v, chunks, tid, reinterpret_cast<double*>(smem_raw));  // barrier inside: S.um is dead after it
}

// ---------------------------------------------------------------------------------------------
// Node scoring (rrt.cpp:220-246): single CTA arg-max with first-strict-maximum semantics.
// ---------------------------------------------------------------------------------------------
struct BestRecord {
  long long index;
  double score, gain, yaw;
};
// Multi-GPU contexts: where the record of this GPU's shard is also posted -- a 64-byte slot in host memory that every rank
// of the context can read (csrc/multi_gpu.h: BestSlot), index made global, the evaluation's sequence number stored last.
struct BestPost {
  void* slot;               // this rank's slot of parity 0 (null: single-GPU context, nothing is posted)
  unsigned stride;          // bytes from there to its slot of parity 1: evaluation `seq` uses slot (seq & 1)
  long long index_offset;   // first candidate of this GPU's shard
  const unsigned* seq_src;  // host word (mapped) holding the evaluation's sequence number: read by the kernel, so that a
                            // captured graph of the step can be replayed without touching its kernel arguments
};

__global__ void __launch_bounds__(1024) k_score_best(const double* __restrict__ gain, const int32_t* __restrict__ yaw_deg,
                                                     const double* __restrict__ parent_yaw, const double* __restrict__ distance,
                                                     const uint8_t* __restrict__ edge_free, int n, double v_max, double dyaw_max,
                                                     double zero_gain, BestRecord* out, double* score_out = nullptr,
                                                     const BestPost post = BestPost{nullptr, 0u, 0, nullptr}) {
  __shared__ double s_score[32];
  __shared__ int s_idx[32];
  double best = zero_gain;
  int best_i = INT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (edge_free && !edge_free[i]) {
      if (score_out) score_out[i] = 0.0;
      continue;
    }
    const double yaw = __ddiv_rn(dmul((double)yaw_deg[i], 3.14159265358979323846), 180.0);  // rrt.cpp:477
    double dy = dsub(yaw, parent_yaw[i]);                                                   // :225
    if (dy > 3.14159265358979323846) dy = dsub(dy, dmul(2.0, 3.14159265358979323846));
    if (dy < -3.14159265358979323846) dy = dadd(dy, dmul(2.0, 3.14159265358979323846));
    const double ta = __ddiv_rn(dy, dyaw_max), tb = __ddiv_rn(distance[i], v_max);
    const double time = (ta < tb) ? tb : ta;          // std::max (:232-233)
    const double score = __ddiv_rn(gain[i], time);    // :234
    if (score_out) score_out[i] = score;
    if (score > best) { best = score; best_i = i; }   // ascending i per thread: first strict max
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (lane == 0) { s_score[warp] = best; s_idx[warp] = best_i; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? s_score[lane] : zero_gain;
    best_i = lane < nw ? s_idx[lane] : INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    if (lane == 0) {
      BestRecord r;
      if (best_i == INT_MAX) {
        r.index = -1; r.score = zero_gain; r.gain = 0.0; r.yaw = 0.0;
      } else {
        r.index = best_i; r.score = best; r.gain = gain[best_i];
        r.yaw = __ddiv_rn(dmul((double)yaw_deg[best_i], 3.14159265358979323846), 180.0);
      }
      if (out) *out = r;
      if (post.slot) {  // record first, system-wide fence, then the ticket the hosts poll
        const unsigned seq = *reinterpret_cast<const volatile unsigned*>(post.seq_src);
        char* const slot = reinterpret_cast<char*>(post.slot) + (seq & 1u) * post.stride;
        volatile BestRecord* s = reinterpret_cast<volatile BestRecord*>(slot);
        s->index = r.index >= 0 ? r.index + post.index_offset : -1;
        s->score = r.score; s->gain = r.gain; s->yaw = r.yaw;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned*>(slot + 32) = seq;
      }
    }
  }
}

// Window scan and node scoring in one launch (large batches): CTA v finishes candidate v (best_yaw_cta), scores it
// (rrt.cpp:220-234) and takes a ticket; the CTA that draws the last ticket reduces the n scores to the best node
// (rrt.cpp:243-246: first strict maximum over zero_gain, lowest index on ties) and posts it.  A masked candidate scores -inf,
// which never wins (strict comparison).
struct ScoreArgs {
  const double* parent_yaw;   // [n]
  const double* distance;     // [n]
  const uint8_t* edge_free;   // [n] or null
  double v_max, dyaw_max, zero_gain;
  double* score;              // [n] scratch
  int* counter;               // zero between launches
  BestRecord* out;            // device record or null
  BestPost post;
};
__device__ __forceinline__ double node_score(double gain, int yaw_deg, double parent_yaw, double distance, double v_max, double dyaw_max) {
  const double yaw = __ddiv_rn(dmul((double)yaw_deg, 3.14159265358979323846), 180.0);  // rrt.cpp:477
  double dy = dsub(yaw, parent_yaw);                                                   // :225
  if (dy > 3.14159265358979323846) dy = dsub(dy, dmul(2.0, 3.14159265358979323846));
  if (dy < -3.14159265358979323846) dy = dadd(dy, dmul(2.0, 3.14159265358979323846));
  const double ta = __ddiv_rn(dy, dyaw_max), tb = __ddiv_rn(distance, v_max);
  const double time = (ta < tb) ? tb : ta;  // std::max (:232-233)
  return __ddiv_rn(gain, time);             // :234
}
__global__ void __launch_bounds__(128) k_best_yaw_score(const BestYawArgs a, const ScoreArgs sc) {
  __shared__ double vg[360];
  __shared__ double s_score[4];
  __shared__ int s_idx[4];
  __shared__ int s_last;
  const int v = (int)blockIdx.x, tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned seq = 0u;
  best_yaw_cta(a, v, vg, tid, (int)blockDim.x);
  __syncthreads();
  if (tid == 0) {
    const bool ok = !sc.edge_free || sc.edge_free[v];
    sc.score[v] = ok ? node_score(a.gain[v], a.yaw_deg[v], sc.parent_yaw[v], sc.distance[v], sc.v_max, sc.dyaw_max) : -INFINITY;
    __threadfence();
    s_last = (atomicAdd(sc.counter, 1) == a.n - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;  // CTA-uniform
  __threadfence();
  // the last CTA only: the evaluation's sequence number comes from host memory (one PCIe round trip, issued first)
  if (tid == 32 && sc.post.slot) seq = *reinterpret_cast<const volatile unsigned*>(sc.post.seq_src);
  double best = sc.zero_gain;
  int best_i = INT_MAX;
  for (int i = tid; i < a.n; i += (int)blockDim.x) {
    const double s = __ldcg(sc.score + i);
    if (s > best) { best = s; best_i = i; }  // ascending i per thread: first strict maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (lane == 0) { s_score[warp] = best; s_idx[warp] = best_i; }
  seq = __shfl_sync(0xffffffffu, seq, 0);  // warp 1 holds it in its lane 0
  __syncthreads();
  if (warp != 1) return;
  if (lane == 0) {
    best = s_score[0];
    best_i = s_idx[0];
    for (int w = 1; w < 4; ++w)
      if (s_score[w] > best || (s_score[w] == best && s_idx[w] < best_i)) { best = s_score[w]; best_i = s_idx[w]; }
    BestRecord r;
    if (best_i == INT_MAX) {
      r.index = -1; r.score = sc.zero_gain; r.gain = 0.0; r.yaw = 0.0;
    } else {
      r.index = best_i; r.score = best; r.gain = __ldcg(a.gain + best_i);
      r.yaw = __ddiv_rn(dmul((double)__ldcg(a.yaw_deg + best_i), 3.14159265358979323846), 180.0);
    }
    if (sc.out) *sc.out = r;
    *sc.counter = 0;
    if (sc.post.slot) {  // record first, system-wide fence, then the ticket the hosts poll
      char* const slot = reinterpret_cast<char*>(sc.post.slot) + (seq & 1u) * sc.post.stride;
      volatile BestRecord* s = reinterpret_cast<volatile BestRecord*>(slot);
      s->index = r.index >= 0 ? r.index + sc.post.index_offset : -1;
      s->score = r.score; s->gain = r.gain; s->yaw = r.yaw;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(slot + 32) = seq;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Batched RrtTree::iterate front half (rrt.cpp:172-212,223): sample -> reject -> nearest parent ->
// clip -> edge with overshoot.  One thread per uniform triple; FP64 in the reference's operation order.
// status: 0 = sample rejected, 1 = candidate produced (collision check and gain follow).
// ---------------------------------------------------------------------------------------------
struct ExpandArgs {
  double root[3];
  double vicinity, extension_range, overshoot;
  double bbx_min[3], bbx_max[3];
  const double* node_state;     // [n_nodes][4] x y z yaw
  const double* node_distance;  // [n_nodes]
  const double* uniforms;       // [n][3]
  int n_nodes, n;
  uint8_t* status;              // [n]
  double* pos;                  // [n][3]
  double* seg;                  // [n][6]
  int32_t* parent;              // [n]
  double* parent_yaw;           // [n]
  double* distance;             // [n]
};

// rrt.cpp:176-185: the sample of triple i; false when it is rejected (outside the vicinity sphere or the exploration box).
__device__ __forceinline__ bool expand_sample_point(const ExpandArgs& a, int i, double (&s)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) s[k] = dmul(dmul(2.0, a.vicinity), dsub(a.uniforms[3 * i + k], 0.5));  // :176-178
  bool ok = !(dadd(dadd(dmul(s[0], s[0]), dmul(s[1], s[1])), dmul(s[2], s[2])) > dmul(a.vicinity, a.vicinity));  // :179
#pragma unroll
  for (int k = 0; k < 3; ++k) s[k] = dadd(s[k], a.root[k]);                                           // :182
  return ok && s[0] >= a.bbx_min[0] && s[0] <= a.bbx_max[0] && s[1] >= a.bbx_min[1] && s[1] <= a.bbx_max[1] &&
         s[2] >= a.bbx_min[2] && s[2] <= a.bbx_max[2];                                                // :183-185
}
__device__ __forceinline__ void expand_write_rejected(const ExpandArgs& a, int i) {
  a.status[i] = 0;
  for (int k = 0; k < 3; ++k) a.pos[3 * i + k] = 0.0;
  for (int k = 0; k < 6; ++k) a.seg[6 * i + k] = 0.0;
  a.parent[i] = -1;
  a.parent_yaw[i] = 0.0;
  a.distance[i] = 0.0;
}
// kd_nearest3 (kdtree.c): squared distance accumulated x, y, z
__device__ __forceinline__ double expand_dist2(const double* node, const double (&s)[3]) {
  const double dx = dsub(node[0], s[0]), dy = dsub(node[1], s[1]), dz = dsub(node[2], s[2]);
  return dadd(dadd(dadd(0.0, dmul(dx, dx)), dmul(dy, dy)), dmul(dz, dz));
}
// rrt.cpp:200-212,223: clip to extension_range, new state, edge with overshoot, Node::distance_; writes the per-sample outputs.
__device__ __forceinline__ void expand_make_edge(const ExpandArgs& a, int i, const double (&s)[3], int best, const double* node_state,
                                                 const double* node_distance, double (&ns_out)[3], double& dist_out) {
  const D3 origin{node_state[4 * best], node_state[4 * best + 1], node_state[4 * best + 2]};
  D3 dir = d3sub(D3{s[0], s[1], s[2]}, origin);                                                       // :202-203
  if (__dsqrt_rn(d3dot(dir, dir)) > a.extension_range) {                                              // :204-206
    const D3 un = d3normalized(dir);
    dir = D3{dmul(a.extension_range, un.x), dmul(a.extension_range, un.y), dmul(a.extension_range, un.z)};
  }
  const D3 ns = d3add(origin, dir);                                                                   // :207-209
  const D3 un = d3normalized(dir);
  const D3 end = d3add(d3add(dir, origin), D3{dmul(un.x, a.overshoot), dmul(un.y, a.overshoot), dmul(un.z, a.overshoot)});  // :210-212
  a.pos[3 * i] = ns.x; a.pos[3 * i + 1] = ns.y; a.pos[3 * i + 2] = ns.z;
  a.seg[6 * i] = origin.x; a.seg[6 * i + 1] = origin.y; a.seg[6 * i + 2] = origin.z;
  a.seg[6 * i + 3] = end.x; a.seg[6 * i + 4] = end.y; a.seg[6 * i + 5] = end.z;
  a.parent[i] = best;
  a.parent_yaw[i] = node_state[4 * best + 3];
  dist_out = dadd(node_distance[best], __dsqrt_rn(d3dot(dir, dir)));                                  // :223
  a.distance[i] = dist_out;
  ns_out[0] = ns.x; ns_out[1] = ns.y; ns_out[2] = ns.z;
}

__global__ void __launch_bounds__(128) k_expand_sample(const ExpandArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  double s[3];
  const bool ok = expand_sample_point(a, i, s);
  a.status[i] = ok ? 1 : 0;
  if (!ok) {
    expand_write_rejected(a, i);
    return;
  }
  // first minimum over the existing nodes
  int best = 0;
  double best_d = 0.0;
  for (int j = 0; j < a.n_nodes; ++j) {
    const double d = expand_dist2(a.node_state + 4 * j, s);
    if (j == 0 || d < best_d) { best_d = d; best = j; }
  }
  double ns[3], dist;
  expand_make_edge(a, i, s, best, a.node_state, a.node_distance, ns, dist);
}

// status 1 + free edge -> 2 (node created); valid[i] = (status == 2) for the gain / scoring kernels.
__global__ void k_expand_status(uint8_t* status, const uint8_t* __restrict__ edge_free, uint8_t* valid, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t st = status[i] ? (edge_free[i] ? 2 : 1) : 0;
  status[i] = st;
  valid[i] = st == 2 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// HighResManager::checkMotion (voxblox_manager.h:89-108).
// ---------------------------------------------------------------------------------------------
struct SegmentWalk {  // RayCaster state of one segment (integrator_utils.cc:127-179, float arithmetic)
  int g[3], sg[3];
  float t[3], dt[3];
  int len;
  bool ok;  // end points finite and in range
};

__device__ __forceinline__ void segment_setup(const double* s6, float voxel_size, SegmentWalk& w) {
  w.len = 0;
  w.ok = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float s = __fdiv_rn(__double2float_rn(s6[k]), voxel_size);      // :96-98 float division
    const float e = __fdiv_rn(__double2float_rn(s6[3 + k]), voxel_size);
    if (!(fabsf(s) < 4194304.0f) || !(fabsf(e) < 4194304.0f)) w.ok = false;  // non-finite / out of range
    w.g[k] = (int)floorf(__fadd_rn(s, 1e-6f));
    const int ei = (int)floorf(__fadd_rn(e, 1e-6f));
    w.len += abs(ei - w.g[k]);
    const float rr = __fsub_rn(e, s);
    w.sg[k] = (rr == 0.0f) ? 0 : (rr < 0.0f ? -1 : 1);
    const float shifted = __fsub_rn(s, (float)w.g[k]);
    const float dist = __fsub_rn((float)max(0, w.sg[k]), shifted);
    w.t[k] = __fdiv_rn(dist, rr);
    w.dt[k] = __fdiv_rn((float)w.sg[k], rr);
  }
}

__device__ __forceinline__ void segment_advance(SegmentWalk& w) {
  const bool c1 = w.t[1] < w.t[0];
  const float bm = c1 ? w.t[1] : w.t[0];
  const bool c2 = w.t[2] < bm;
  if (c2) { w.g[2] += w.sg[2]; w.t[2] = __fadd_rn(w.t[2], w.dt[2]); }
  else if (c1) { w.g[1] += w.sg[1]; w.t[1] = __fadd_rn(w.t[1], w.dt[1]); }
  else { w.g[0] += w.sg[0]; w.t[0] = __fadd_rn(w.t[0], w.dt[0]); }
}

// One thread per segment (large batches).
__global__ void __launch_bounds__(128) k_check_segments(LayerView L, float voxel_size, double radius,
                                                        const double* __restrict__ seg, int n, uint8_t* is_free,
                                                        unsigned long long* totals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long steps = 0ull;
  if (i < n) {
    const int lv = L.log2vps, vm = (1 << lv) - 1;
    SegmentWalk w;
    segment_setup(seg + 6 * (size_t)i, voxel_size, w);
    bool free_seg = w.ok;
    if (w.ok) {
      int cbx = INT_MIN, cby = 0, cbz = 0, cslot = -1;
      for (int step = 0; step <= w.len; ++step) {
        ++steps;
        const int bx = w.g[0] >> lv, by = w.g[1] >> lv, bz = w.g[2] >> lv;
        if (bx != cbx || by != cby || bz != cbz) {
          cbx = bx; cby = by; cbz = bz;
          cslot = probe_block(L, bx, by, bz);
        }
        if (cslot < 0) { free_seg = false; break; }  // voxel == nullptr -> collision (voxblox_manager.cpp:125-127)
        const int lin = (w.g[0] & vm) | ((w.g[1] & vm) << lv) | ((w.g[2] & vm) << (2 * lv));
        const float d = __ldg(reinterpret_cast<const float*>(L.tiles) + (size_t)cslot * L.nvox + lin);
        if (radius >= (double)d) { free_seg = false; break; }  // :128
        segment_advance(w);
      }
    }
    is_free[i] = free_seg ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(totals + 2, steps);
}

// One warp per segment (small batches, where the latency of one call is what counts): the walk itself is cheap
// arithmetic and sequential, the block probe and the distance load of each voxel are what takes time -- so every lane
// runs the same walk, lane s keeps the voxel of step base + s, and the 32 voxels are then looked up at once.  The first
// failing step decides, exactly as in the sequential loop (same flag, same count of executed steps).
struct SegmentByValue { double v[6]; };  // a single segment passed in the launch arguments (seg == nullptr)

// Warp-collective checkMotion of one segment (all 32 lanes call it with the same s6): true = free; *steps = executed steps.
__device__ __forceinline__ bool segment_free_warp(const LayerView& L, float voxel_size, double radius, const double (&s6)[6], int lane,
                                                  unsigned long long* steps_out) {
  const int lv = L.log2vps, vm = (1 << lv) - 1;
  SegmentWalk w;
  segment_setup(s6, voxel_size, w);
  bool free_seg = w.ok;
  unsigned long long steps = 0ull;
  if (w.ok) {
    for (int base = 0; base <= w.len; base += 32) {
      const int cnt = min(32, w.len + 1 - base);
      int mx = 0, my = 0, mz = 0;
      for (int s = 0; s < cnt; ++s) {
        if (lane == s) { mx = w.g[0]; my = w.g[1]; mz = w.g[2]; }
        segment_advance(w);
      }
      bool fail = false;
      if (lane < cnt) {
        const int slot = probe_block(L, mx >> lv, my >> lv, mz >> lv);
        if (slot < 0) {
          fail = true;  // voxel == nullptr -> collision (voxblox_manager.cpp:125-127)
        } else {
          const int lin = (mx & vm) | ((my & vm) << lv) | ((mz & vm) << (2 * lv));
          const float d = __ldg(reinterpret_cast<const float*>(L.tiles) + (size_t)slot * L.nvox + lin);
          fail = radius >= (double)d;  // :128
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, fail);
      if (m) {
        steps += (unsigned)__ffs(m);
        free_seg = false;
        break;
      }
      steps += (unsigned)cnt;
    }
  }
  *steps_out = steps;
  return free_seg;
}

__global__ void __launch_bounds__(128) k_check_segments_warp(LayerView L, float voxel_size, double radius,
                                                             const double* __restrict__ seg, SegmentByValue one, int n,
                                                             uint8_t* is_free, unsigned long long* totals, unsigned* done,
                                                             unsigned seq) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  double s6[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) s6[k] = seg ? seg[6 * (size_t)i + k] : one.v[k];
  unsigned long long steps = 0ull;
  const bool free_seg = segment_free_warp(L, voxel_size, radius, s6, lane, &steps);
  if (lane == 0) {
    is_free[i] = free_seg ? 1 : 0;
    if (steps) atomicAdd(totals + 2, steps);
    publish_done(done, seq);  // single-segment host calls only (see BestYawArgs::done)
  }
}

// ---------------------------------------------------------------------------------------------
// RrtTree::iterate's front half with the reference's SEQUENTIAL semantics (rrt.cpp:172-212,223,236): sample i finds its
// parent among the given nodes AND the nodes created by samples 0..i-1 of the same batch, so the cheap, order-dependent
// part -- sample, nearest node, clip, checkMotion, insert -- runs sample by sample inside ONE CTA (the nearest-node search
// and the edge's voxel look-ups are what is parallel), and the expensive, order-independent part (the gains) is batched
// afterwards.  all_state / all_dist hold the given nodes followed by room for n created ones; created_triple[k] = the sample
// that created node n_nodes + k.
// ---------------------------------------------------------------------------------------------
struct ExpandSeqArgs {
  ExpandArgs a;          // node_state / node_distance are unused here: all_state / all_dist carry the growing tree
  LayerView E;           // ESDF layer of the edge check
  float esdf_voxel_size;
  double radius;
  double* all_state;     // [(n_nodes + n)][4] x y z yaw (yaw of created nodes filled in by k_expand_parent_yaw)
  double* all_dist;      // [n_nodes + n]
  int* created_triple;   // [n]
  int* n_created;        // [1]
  uint8_t* valid;        // [n] 1 = node created
  unsigned long long* totals;
};
constexpr int kExpandSeqThreads = 1024;

__global__ void __launch_bounds__(kExpandSeqThreads) k_expand_sequential(const ExpandSeqArgs q) {
  const ExpandArgs& a = q.a;
  __shared__ double s_d[32];
  __shared__ int s_j[32];
  __shared__ int s_cnt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  unsigned long long my_steps = 0ull;
  for (int i = 0; i < a.n; ++i) {
    double s[3];
    const bool ok = expand_sample_point(a, i, s);  // every thread computes the same sample: the branch is uniform
    if (!ok) {
      if (tid == 0) {
        expand_write_rejected(a, i);
        q.valid[i] = 0;
      }
      continue;
    }
    // nearest node, first minimum in node order (ties: lowest index)
    const int total = a.n_nodes + s_cnt;
    double best_d = 1.0e300;
    int best = 0x7fffffff;
    for (int j = tid; j < total; j += kExpandSeqThreads) {
      const double d = expand_dist2(q.all_state + 4 * j, s);
      if (d < best_d) { best_d = d; best = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, best_d, o);
      const int oj = __shfl_xor_sync(0xffffffffu, best, o);
      if (od < best_d || (od == best_d && oj < best)) { best_d = od; best = oj; }
    }
    if (lane == 0) { s_d[warp] = best_d; s_j[warp] = best; }
    __syncthreads();
    if (warp == 0) {
      best_d = s_d[lane];
      best = s_j[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best_d, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best, o);
        if (od < best_d || (od == best_d && oj < best)) { best_d = od; best = oj; }
      }
      if (best == 0x7fffffff) best = 0;  // every distance NaN (cannot happen for validated inputs): kd_nearest3 returns a node anyway
      // edge + collision check + insertion by the first warp (lane 0 does the scalar arithmetic)
      double ns[3] = {0.0, 0.0, 0.0}, dist = 0.0, s6[6];
      if (lane == 0) expand_make_edge(a, i, s, best, q.all_state, q.all_dist, ns, dist);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 6; ++k) s6[k] = a.seg[6 * i + k];  // written by lane 0 just above
      unsigned long long steps = 0ull;
      const bool free_seg = segment_free_warp(q.E, q.esdf_voxel_size, q.radius, s6, lane, &steps);
      if (lane == 0) {
        my_steps += steps;
        a.status[i] = free_seg ? 2 : 1;
        q.valid[i] = free_seg ? 1 : 0;
        if (free_seg) {  // kd_insert3 (rrt.cpp:236)
          const int node = a.n_nodes + s_cnt;
          q.all_state[4 * node] = ns[0]; q.all_state[4 * node + 1] = ns[1]; q.all_state[4 * node + 2] = ns[2]; q.all_state[4 * node + 3] = 0.0;
          q.all_dist[node] = dist;
          q.created_triple[s_cnt] = i;
          s_cnt = s_cnt + 1;
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    *q.n_created = s_cnt;
    if (my_steps) atomicAdd(q.totals + 2, my_steps);
  }
}

// After the gains: the yaw each created node chose (state[3] = max_gain_yaw * M_PI / 180.0, rrt.cpp:477) goes into the tree,
// and every sample's parent_yaw is read from there -- the parent may be a node of the same batch (rrt.cpp:225).
__global__ void k_expand_parent_yaw(const int32_t* __restrict__ yaw_deg, const int* __restrict__ created_triple, const int* __restrict__ n_created,
                                    int n_nodes, double* all_state, const int32_t* __restrict__ parent, const uint8_t* __restrict__ valid,
                                    double* parent_yaw, int n, int phase) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (phase == 0) {
    if (i < *n_created)
      all_state[4 * (n_nodes + i) + 3] = __ddiv_rn(dmul((double)yaw_deg[created_triple[i]], 3.14159265358979323846), 180.0);
  } else if (i < n && valid[i]) {
    parent_yaw[i] = all_state[4 * parent[i] + 3];
  }
}

}  // namespace nbvgpu
#endif
