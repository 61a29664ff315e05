// interp_kernels.cuh -- batched ESDF interpolated distance and gradient:
//   HighResManager::getDistanceAndGradientAtPosition   nbveplanner/src/voxblox_manager.cpp:143-148
//     -> EsdfMap::getDistanceAndGradientAtPosition      voxblox/src/core/esdf_map.cc:22-52
//     -> Interpolator<EsdfVoxel>::getInterpDistance     voxblox/interpolator/interpolator_inl.h:318-330
//          (setIndexes :160-198, getVoxelsAndQVector :224-272, getQVector :200-222, interpMember :447-474)
//        Interpolator<EsdfVoxel>::getGradient           interpolator_inl.h:47-77
// One thread per query, everything in float as the reference: block / truncated voxel index by coordinates, shift
// to the lower corner of the eight surrounding voxel centres, neighbouring blocks where an index leaves the block,
// every voxel must be `observed` (the bit tile the ESDF scatter kernel keeps), trilinear weights from the offset to
// the lower corner, gradient = central differences of six more interpolations at +-voxel_size divided by
// 2 * voxel_size.  As in the reference the gradient is evaluated even when the distance failed, and a failing
// gradient leaves its partial sums in the output.  The two products of interpMember are summed left to right,
// like the oracle (and the Eigen stand-in of the pinning build): against a real Eigen build the VALUES are claimed
// within float tolerance only (the summation order of a fixed-size float product belongs to the Eigen version and
// vector flags of that build), the success flag and the choice of voxels exactly.
#ifndef NBVGPU_INTERP_KERNELS_CUH_
#define NBVGPU_INTERP_KERNELS_CUH_

#include "map_mirror.cuh"

namespace nbvgpu {

struct EsdfView {
  LayerView L;  // tiles: float distance [slot][nvox]; status: 1 bit per voxel "observed" [slot][nvox / 32]
  float voxel_size, block_size, block_size_inv, voxel_size_inv;
  int vps;
};

// Block::computeCoordinatesFromVoxelIndex (block.h:90-92): origin_ + getCenterPointFromGridIndex (common.h:188-193:
// (float(idx) + 0.5) * voxel_size in double, rounded to float); origin_ = float(block index) * block_size (:196-201)
__device__ __forceinline__ float esdf_voxel_centre(const EsdfView& E, int block, int voxel) {
  const float origin = __fmul_rn((float)block, E.block_size);
  const float centre = __double2float_rn(__dmul_rn(__dadd_rn((double)(float)voxel, 0.5), (double)E.voxel_size));
  return __fadd_rn(origin, centre);
}
__device__ __forceinline__ int esdf_grid_index(float v, float inv) { return (int)floorf(__fadd_rn(__fmul_rn(v, inv), 1e-6f)); }

// Interpolator::getInterpDistance.  Inlined seven times per query on purpose: ncu shows instruction-fetch stalls for
// that, but the out-of-line version measured 18 % slower (0.300 vs 0.254 ms per million queries, device-resident).
// Issuing the sixteen loads of the eight voxels back to back before any check (the function is pure, only the flag
// matters) was measured as well: 72 registers instead of 56, 3.69 instead of 3.92 G queries/s, 1 us off a single query.
__device__ __forceinline__ bool esdf_interp_distance(const EsdfView& E, float px, float py, float pz, float* distance) {
  const float pos[3] = {px, py, pz};
  int b[3], v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) b[k] = esdf_grid_index(pos[k], E.block_size_inv);
  int home = probe_block(E.L, b[0], b[1], b[2]);
  if (home < 0) return false;
  bool shifted = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float origin = __fmul_rn((float)b[k], E.block_size);
    const int t = esdf_grid_index(__fsub_rn(pos[k], origin), E.voxel_size_inv);
    v[k] = max(min(t, E.vps - 1), 0);
    if (__fsub_rn(pos[k], esdf_voxel_centre(E, b[k], v[k])) < 0.0f) {
      v[k]--;
      if (v[k] < 0) {
        b[k]--;
        v[k] += E.vps;
        shifted = true;
      }
    }
  }
  if (shifted) {
    home = probe_block(E.L, b[0], b[1], b[2]);
    if (home < 0) return false;
  }
  float q[8], data[8];
  const unsigned nvox = (unsigned)E.L.nvox;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int nb[3], nv[3];
    bool other = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int off = k == 0 ? ((i >> 2) & 1) : (k == 1 ? ((i >> 1) & 1) : (i & 1));
      nb[k] = b[k];
      nv[k] = v[k] + off;
      if (nv[k] >= E.vps) {
        nb[k]++;
        nv[k] -= E.vps;
        other = true;
      }
    }
    int slot = home;
    if (other) {
      slot = probe_block(E.L, nb[0], nb[1], nb[2]);
      if (slot < 0) return false;
    }
    if (i == 0) {
      const float o0 = __fmul_rn(__fsub_rn(pos[0], esdf_voxel_centre(E, nb[0], nv[0])), E.voxel_size_inv);
      const float o1 = __fmul_rn(__fsub_rn(pos[1], esdf_voxel_centre(E, nb[1], nv[1])), E.voxel_size_inv);
      const float o2 = __fmul_rn(__fsub_rn(pos[2], esdf_voxel_centre(E, nb[2], nv[2])), E.voxel_size_inv);
      q[0] = 1.0f; q[1] = o0; q[2] = o1; q[3] = o2;
      q[4] = __fmul_rn(o0, o1); q[5] = __fmul_rn(o1, o2); q[6] = __fmul_rn(o2, o0); q[7] = __fmul_rn(__fmul_rn(o0, o1), o2);
    }
    const unsigned lin = (unsigned)(nv[0] + E.vps * (nv[1] + E.vps * nv[2]));
    data[i] = __ldg(reinterpret_cast<const float*>(E.L.tiles) + ((size_t)slot * nvox + lin));
    const uint32_t w = __ldg(E.L.status + ((size_t)slot * (nvox >> 5) + (lin >> 5)));
    if (((w >> (lin & 31u)) & 1u) == 0u) return false;  // utils::isObservedVoxel
  }
  // interpMember: q * (table * data^T), both sums left to right over the inner index
  const float T[8][8] = {{1, 0, 0, 0, 0, 0, 0, 0},  {-1, 0, 0, 0, 1, 0, 0, 0}, {-1, 0, 1, 0, 0, 0, 0, 0}, {-1, 1, 0, 0, 0, 0, 0, 0},
                         {1, 0, -1, 0, -1, 0, 1, 0}, {1, -1, -1, 1, 0, 0, 0, 0}, {1, -1, 0, 0, -1, 1, 0, 0}, {-1, 1, 1, -1, 1, -1, -1, 1}};
  float acc = 0.0f;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float y = __fmul_rn(T[r][0], data[0]);
#pragma unroll
    for (int c = 1; c < 8; ++c) y = __fadd_rn(y, __fmul_rn(T[r][c], data[c]));
    const float term = __fmul_rn(q[r], y);
    acc = r == 0 ? term : __fadd_rn(acc, term);
  }
  *distance = acc;
  return true;
}

// one query of EsdfMap::getDistanceAndGradientAtPosition(position, &distance, &gradient)
__global__ void __launch_bounds__(128) k_esdf_distance_gradient(EsdfView E, const double* __restrict__ positions, int n, uint8_t* ok,
                                                                double* distance, double* gradient) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float px = __double2float_rn(positions[3 * i]), py = __double2float_rn(positions[3 * i + 1]), pz = __double2float_rn(positions[3 * i + 2]);
  float d = 0.0f, g[3] = {0.0f, 0.0f, 0.0f};
  // device-resident inputs are not validated by the host: a non-finite or out-of-range position fails like a position
  // outside the map (the host entry point rejects such input with NBVGPU_ERR_RANGE instead)
  const float lim = 4194000.0f * E.voxel_size;
  if (!(fabsf(px) < lim) || !(fabsf(py) < lim) || !(fabsf(pz) < lim)) {
    ok[i] = 0;
    distance[i] = 0.0;
    gradient[3 * i] = gradient[3 * i + 1] = gradient[3 * i + 2] = 0.0;
    return;
  }
  bool success = esdf_interp_distance(E, px, py, pz, &d);
  // getGradient: needs the block that contains the position, then six interpolations at +-voxel_size
  bool gok = probe_block(E.L, esdf_grid_index(px, E.block_size_inv), esdf_grid_index(py, E.block_size_inv), esdf_grid_index(pz, E.block_size_inv)) >= 0;
  if (gok) {
    for (int a = 0; a < 3 && gok; ++a)
      for (int sign = -1; sign <= 1 && gok; sign += 2) {
        const float step = __fmul_rn((float)sign, E.voxel_size);
        // pos + offset: all three components are added, the idle ones with +0.0
        const float qx = __fadd_rn(px, a == 0 ? step : 0.0f), qy = __fadd_rn(py, a == 1 ? step : 0.0f), qz = __fadd_rn(pz, a == 2 ? step : 0.0f);
        float od;
        if (!esdf_interp_distance(E, qx, qy, qz, &od)) {
          gok = false;
        } else {
          g[a] = __fadd_rn(g[a], __fmul_rn(od, (float)sign));
        }
      }
    if (gok) {
      const float denom = __fmul_rn(2.0f, E.voxel_size);
      for (int a = 0; a < 3; ++a) g[a] = __fdiv_rn(g[a], denom);
    }
  }
  success = success && gok;
  ok[i] = success ? 1 : 0;
  distance[i] = (double)d;
  gradient[3 * i] = (double)g[0];
  gradient[3 * i + 1] = (double)g[1];
  gradient[3 * i + 2] = (double)g[2];
}

// The same query on eight lanes (small batches, where the latency of one call is what counts): lane 0 of a group
// interpolates the distance, lanes 1-6 the six gradient samples (-x, +x, -y, +y, -z, +z: the reference's loop order), lane 7
// probes the block the gradient needs; lane 0 then combines them in the reference's order, stopping at the first failed
// sample exactly like the sequential loop (an interpolation has no side effects, so evaluating the later ones anyway
// changes nothing).  Seven dependent chains of probes and loads become one.  done / seq: see publish_done().
__global__ void __launch_bounds__(128) k_esdf_distance_gradient_group(EsdfView E, const double* __restrict__ positions, int n, uint8_t* ok,
                                                                      double* distance, double* gradient, unsigned* done, unsigned seq) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 3, role = t & 7;
  const bool live = i < n;  // whole groups of eight; the warp stays together for the shuffles
  const int ii = live ? i : 0;
  const float px = __double2float_rn(positions[3 * ii]), py = __double2float_rn(positions[3 * ii + 1]), pz = __double2float_rn(positions[3 * ii + 2]);
  const float lim = 4194000.0f * E.voxel_size;
  const bool in_range = (fabsf(px) < lim) && (fabsf(py) < lim) && (fabsf(pz) < lim);
  float val = 0.0f;
  bool good = false;
  if (live && in_range) {
    if (role == 7) {
      good = probe_block(E.L, esdf_grid_index(px, E.block_size_inv), esdf_grid_index(py, E.block_size_inv), esdf_grid_index(pz, E.block_size_inv)) >= 0;
    } else {
      const int a = (role - 1) >> 1;
      const float step = role == 0 ? 0.0f : __fmul_rn((role & 1) ? -1.0f : 1.0f, E.voxel_size);
      // role 0: the position itself; else pos + offset, all three components added, the idle ones with +0.0
      const float qx = role == 0 ? px : __fadd_rn(px, a == 0 ? step : 0.0f);
      const float qy = role == 0 ? py : __fadd_rn(py, a == 1 ? step : 0.0f);
      const float qz = role == 0 ? pz : __fadd_rn(pz, a == 2 ? step : 0.0f);
      good = esdf_interp_distance(E, qx, qy, qz, &val);
    }
  }
  const unsigned lane = threadIdx.x & 31u, base = lane & ~7u;
  const unsigned goods = __ballot_sync(0xffffffffu, good) >> base;  // bit r: role r of this group
  float sample[7];
#pragma unroll
  for (int r = 1; r < 7; ++r) sample[r] = __shfl_sync(0xffffffffu, val, (int)base + r);
  if (!live || role != 0) return;
  float g[3] = {0.0f, 0.0f, 0.0f};
  if (!in_range) {  // as in the one-thread kernel: fails like a position outside the map
    ok[i] = 0;
    distance[i] = 0.0;
    gradient[3 * i] = gradient[3 * i + 1] = gradient[3 * i + 2] = 0.0;
  } else {
    bool gok = (goods >> 7) & 1u;
    if (gok) {
#pragma unroll
      for (int r = 1; r < 7; ++r) {
        if (gok) {
          if (!((goods >> r) & 1u)) {
            gok = false;
          } else {
            const int a = (r - 1) >> 1;
            g[a] = __fadd_rn(g[a], __fmul_rn(sample[r], (r & 1) ? -1.0f : 1.0f));
          }
        }
      }
      if (gok) {
        const float denom = __fmul_rn(2.0f, E.voxel_size);
        for (int a = 0; a < 3; ++a) g[a] = __fdiv_rn(g[a], denom);
      }
    }
    ok[i] = ((goods & 1u) && gok) ? 1 : 0;
    distance[i] = (double)val;
    gradient[3 * i] = (double)g[0];
    gradient[3 * i + 1] = (double)g[1];
    gradient[3 * i + 2] = (double)g[2];
  }
  if (i == n - 1) publish_done(done, seq);  // single-query host calls only (n == 1)
}

}  // namespace nbvgpu

#endif  // NBVGPU_INTERP_KERNELS_CUH_
