// The upload hook compiled against the REFERENCE'S OWN Voxblox headers: include/nbvgpu_shim.hpp with
// NBVGPU_SHIM_WITH_VOXBLOX, i.e. GpuMap::uploadUpdatedBlocks(voxblox::Layer<VoxelType>*, gpu_layer) -- what a maintainer
// calls after TsdfServer::integratePointcloud (tsdf_server.cc:414) and EsdfServer::updateEsdf (esdf_server.cc:196).
// Built by oracle/Makefile (target `ref`) into oracle/_ref/ from /root/reference/voxblox/voxblox/include, src/core/block.cc
// and oracle/ref_shim (stand-ins for Eigen / glog / protobuf); run by tests/test_gpu_upload_hook.py.
//
// Phase 1: fill voxblox::Layer<TsdfVoxel> / Layer<EsdfVoxel> from the case file (Block::deserializeFromIntegers), mark every
//          block Update::kMap as the integrators do (tsdf_integrator.cc:128, esdf_integrator.cc:147), upload, evaluate.
// Phase 2: change three TSDF blocks (weights to 0 -> unknown), mark only those, upload again, evaluate again.
// Output: per sample gain / yaw_deg / free for both phases, then {blocks still marked after phase 1, blocks uploaded in
// phase 1, blocks still marked after phase 2, blocks uploaded in phase 2}.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "voxblox/core/layer.h"

#define NBVGPU_SHIM_WITH_VOXBLOX
#include "nbvgpu_shim.hpp"

template <class T>
static bool rd(std::FILE* f, T* p, size_t n) { return std::fread(p, sizeof(T), n, f) == n; }

template <class VoxelType>
static size_t count_marked(const voxblox::Layer<VoxelType>& layer) {
  voxblox::BlockIndexList l;
  layer.getAllUpdatedBlocks(voxblox::Update::kMap, &l);
  return l.size();
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t hdr[4];
  float voxel_size;
  double cam[21];
  if (!rd(f, hdr, 4) || !rd(f, &voxel_size, 1) || !rd(f, cam, 21)) return 2;
  const int vps = hdr[0], nb = hdr[1], ns = hdr[2];
  const size_t nvox = (size_t)vps * vps * vps;
  std::vector<int32_t> keys((size_t)nb * 3);
  std::vector<uint32_t> tsdf((size_t)nb * nvox * 3), esdf((size_t)nb * nvox * 2);
  std::vector<double> samples((size_t)ns * 9);
  if (!rd(f, keys.data(), keys.size()) || !rd(f, tsdf.data(), tsdf.size()) || !rd(f, esdf.data(), esdf.size()) ||
      !rd(f, samples.data(), samples.size()))
    return 2;
  std::fclose(f);
  try {
    voxblox::Layer<voxblox::TsdfVoxel> tsdf_layer(voxel_size, vps);
    voxblox::Layer<voxblox::EsdfVoxel> esdf_layer(voxel_size, vps);
    for (int b = 0; b < nb; ++b) {
      const voxblox::BlockIndex bi(keys[3 * b], keys[3 * b + 1], keys[3 * b + 2]);
      auto tb = tsdf_layer.allocateBlockPtrByIndex(bi);
      tb->deserializeFromIntegers(std::vector<uint32_t>(tsdf.begin() + (size_t)b * nvox * 3, tsdf.begin() + (size_t)(b + 1) * nvox * 3));
      tb->updated().set();                                  // tsdf_integrator.cc:128 sets every bit
      auto eb = esdf_layer.allocateBlockPtrByIndex(bi);
      eb->deserializeFromIntegers(std::vector<uint32_t>(esdf.begin() + (size_t)b * nvox * 2, esdf.begin() + (size_t)(b + 1) * nvox * 2));
      eb->updated().set(voxblox::Update::kMap);             // esdf_integrator.cc:147: set_updated(true) = bit 0
    }
    nbvgpu_shim::GpuMap map(0, voxel_size, vps, (size_t)nb + 8, voxel_size, vps, (size_t)nb + 8);
    nbvgpu_shim::setUpCamera(map, cam[0], cam[1], cam[2], cam[3], cam + 4, cam + 8, cam + 11, cam + 14, cam[17], cam[18], cam[19]);
    const double radius = cam[20];
    std::vector<double> pos((size_t)ns * 3), seg((size_t)ns * 6), pyaw(ns, 0.0), dist(ns, 1.0), gain(ns);
    std::vector<int32_t> yaw(ns);
    std::vector<uint8_t> fr(ns);
    for (int i = 0; i < ns; ++i) {
      for (int k = 0; k < 3; ++k) pos[3 * i + k] = samples[9 * i + k];
      for (int k = 0; k < 6; ++k) seg[6 * i + k] = samples[9 * i + 3 + k];
    }
    std::FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 2;
    double tail[4];
    nbvgpu_stats s0{}, s1{};
    for (int phase = 0; phase < 2; ++phase) {
      if (phase == 1)
        for (int b = 0; b < 3 && b < nb; ++b) {             // a new observation wipes three blocks: weight 0 = unknown
          auto tb = tsdf_layer.getBlockPtrByIndex(voxblox::BlockIndex(keys[3 * b], keys[3 * b + 1], keys[3 * b + 2]));
          for (size_t v = 0; v < tb->num_voxels(); ++v) tb->getVoxelByLinearIndex(v).weight = 0.0f;
          tb->updated().set();
        }
      nbvgpu_get_stats(map.ctx(), &s0);
      map.uploadUpdatedBlocks(&tsdf_layer, map.tsdf_layer());
      map.uploadUpdatedBlocks(&esdf_layer, map.esdf_layer());
      nbvgpu_get_stats(map.ctx(), &s1);
      tail[2 * phase] = (double)(count_marked(tsdf_layer) + count_marked(esdf_layer));
      tail[2 * phase + 1] = (double)(s1.blocks_uploaded - s0.blocks_uploaded);
      nbvgpu_shim::evaluateCandidates(map, radius, (size_t)ns, pos.data(), seg.data(), pyaw.data(), dist.data(), 1.0, 0.5, 0.0, fr.data(),
                                      gain.data(), yaw.data());
      for (int i = 0; i < ns; ++i) {
        const double row[3] = {gain[i], (double)yaw[i], (double)fr[i]};
        std::fwrite(row, sizeof(double), 3, o);
      }
    }
    std::fwrite(tail, sizeof(double), 4, o);
    std::fclose(o);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
