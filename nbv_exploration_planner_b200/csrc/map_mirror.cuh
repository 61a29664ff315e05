// map_mirror.cuh -- GPU mirror of voxblox::Layer<TsdfVoxel|EsdfVoxel>
// (voxblox/voxblox/include/voxblox/core/layer.h:23-296, block.h:22-215, block_hash.h:20-31).
//
// Reference: std::unordered_map<BlockIndex, shared_ptr<Block>> with dense vps^3 voxel arrays.
// Here: one open-addressing table (16-byte entries {packed BlockIndex, tile slot}, linear
// probing, power-of-two capacity >= 2 x max_blocks) plus a pool of dense tiles in HBM:
//   TSDF tile   float2[vps^3]  = {distance, weight}, x-fastest (block_inl.h:12-27 linear index)
//   TSDF status uint32[vps^3/16]: 2 bits per voxel (0 unknown / 1 free / 2 occupied), the value
//               of VoxbloxManager::getVoxelStatusByIndex (voxblox_manager.cpp:14-33), rebuilt by
//               the scatter kernel whenever a tile is written (a pure function of the tile)
//   ESDF tile   float[vps^3]   = distance
// Updates are batched: insert kernel (atomicCAS claim + free-list pop) then a scatter kernel
// with 128-bit loads/stores; nothing in the evaluation kernels ever writes the map.
#ifndef NBVGPU_MAP_MIRROR_CUH_
#define NBVGPU_MAP_MIRROR_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace nbvgpu {

struct __align__(16) HashEntry {
  unsigned long long key;
  int slot;
  int pad;
};
static constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
static constexpr unsigned long long KEY_TOMB = 0xFFFFFFFFFFFFFFFEull;

// 21 bits per axis: block indices in [-2^20, 2^20).
__host__ __device__ __forceinline__ unsigned long long pack_key(int x, int y, int z) {
  return ((unsigned long long)((unsigned)x & 0x1FFFFFu)) | ((unsigned long long)((unsigned)y & 0x1FFFFFu) << 21) |
         ((unsigned long long)((unsigned)z & 0x1FFFFFu) << 42);
}
// Reference AnyIndexHash (block_hash.h:27-30: x + y*17191 + z*17191^2 truncated to 32 bit),
// followed by a Fibonacci multiply so that the *high* bits index the power-of-two table.
__host__ __device__ __forceinline__ uint32_t hash_key(int x, int y, int z) {
  const uint32_t h = (uint32_t)x + (uint32_t)y * 17191u + (uint32_t)z * 295530481u;
  return h * 2654435769u;
}

struct LayerView {
  const HashEntry* table;
  const void* tiles;       // float2 (TSDF) or float (ESDF)
  const uint32_t* status;  // TSDF only
  uint32_t mask;           // capacity - 1
  uint32_t shift;          // 32 - log2(capacity)
  int log2vps;
  int nvox;                // vps^3
};

// Layer::getBlockPtrByIndex (layer.h:72-80): tile slot or -1.
__device__ __forceinline__ int probe_block(const LayerView& L, int bx, int by, int bz) {
  const unsigned long long key = pack_key(bx, by, bz);
  uint32_t idx = hash_key(bx, by, bz) >> L.shift;
  for (uint32_t n = 0; n <= L.mask; ++n) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(L.table + idx));
    const unsigned long long k = ((unsigned long long)(uint32_t)raw.y << 32) | (uint32_t)raw.x;
    if (k == key) return raw.z;
    if (k == KEY_EMPTY) return -1;
    idx = (idx + 1) & L.mask;
  }
  return -1;
}

// VoxbloxManager::getVoxelStatusByIndex (voxblox_manager.cpp:23-29): float members compared
// against double literals.
__device__ __forceinline__ uint32_t tsdf_status(float distance, float weight) {
  if ((double)weight <= 1e-1) return 0u;   // kUnknown
  if ((double)distance <= 0.0) return 2u;  // kOccupied
  return 1u;                               // kFree
}

// ---------------------------------------------------------------------------------------------
// Update kernels.
// ---------------------------------------------------------------------------------------------
// err bits: 1 = tile pool exhausted, 2 = table full.
__global__ void k_insert_keys(HashEntry* table, uint32_t mask, uint32_t shift, const int32_t* __restrict__ keys,
                              int n, const int* __restrict__ free_list, int* free_top, int* out_slots, int* err,
                              unsigned long long* n_blocks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int bx = keys[3 * i], by = keys[3 * i + 1], bz = keys[3 * i + 2];
  const unsigned long long key = pack_key(bx, by, bz);
  uint32_t idx = hash_key(bx, by, bz) >> shift;
  int slot = -1;
  for (uint32_t p = 0; p <= mask; ++p) {
    HashEntry* e = table + idx;
    const unsigned long long prev = atomicCAS(&e->key, KEY_EMPTY, key);
    if (prev == KEY_EMPTY) {  // claimed a fresh entry: allocate a tile
      const int top = atomicSub(free_top, 1) - 1;
      if (top < 0) {
        atomicAdd(free_top, 1);
        atomicOr(err, 1);
        e->slot = -1;
        e->key = KEY_TOMB;  // plain store: nobody else matches this key in the batch
      } else {
        slot = free_list[top];
        e->slot = slot;
        atomicAdd(n_blocks, 1ull);
      }
      break;
    }
    if (prev == key) {  // already mapped by an earlier commit: overwrite in place
      slot = e->slot;
      break;
    }
    idx = (idx + 1) & mask;
    if (p == mask) atomicOr(err, 2);
  }
  out_slots[i] = slot;
}

__global__ void k_remove_keys(HashEntry* table, uint32_t mask, uint32_t shift, const int32_t* __restrict__ keys, int n,
                              int* free_list, int* free_top, unsigned long long* n_blocks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int bx = keys[3 * i], by = keys[3 * i + 1], bz = keys[3 * i + 2];
  const unsigned long long key = pack_key(bx, by, bz);
  uint32_t idx = hash_key(bx, by, bz) >> shift;
  for (uint32_t p = 0; p <= mask; ++p) {
    HashEntry* e = table + idx;
    const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&e->key);
    if (k == key) {
      // claim the entry: a key listed twice in one batch (Layer::removeBlock of an absent block is a no-op in the
      // reference, so callers may do that) is freed by exactly one thread
      if (atomicCAS(&e->key, key, KEY_TOMB) != key) return;
      const int slot = e->slot;
      e->slot = -1;
      if (slot >= 0) {
        const int top = atomicAdd(free_top, 1);
        free_list[top] = slot;
        atomicAdd(n_blocks, ~0ull);  // -1
      }
      return;
    }
    if (k == KEY_EMPTY) return;
    idx = (idx + 1) & mask;
  }
}

__global__ void k_reset_table(HashEntry* table, uint32_t cap, int* free_list, int max_blocks, int* free_top,
                              unsigned long long* n_blocks) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) {
    table[i].key = KEY_EMPTY;
    table[i].slot = -1;
    table[i].pad = 0;
  }
  if (i < (uint32_t)max_blocks) free_list[i] = max_blocks - 1 - (int)i;  // pop order 0,1,2,...
  if (i == 0) {
    *free_top = max_blocks;
    *n_blocks = 0ull;
  }
}

// Re-insert every live entry of `old_table` into the (reset) `table` -- drops tombstones.
__global__ void k_rehash(const HashEntry* __restrict__ old_table, uint32_t old_cap, HashEntry* table, uint32_t mask,
                         uint32_t shift) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  const HashEntry e = old_table[i];
  if (e.key == KEY_EMPTY || e.key == KEY_TOMB) return;
  // sign-extend the 21-bit fields back to block indices for the hash
  const int bx = ((int)((uint32_t)(e.key & 0x1FFFFFu) << 11)) >> 11;
  const int by = ((int)((uint32_t)((e.key >> 21) & 0x1FFFFFu) << 11)) >> 11;
  const int bz = ((int)((uint32_t)((e.key >> 42) & 0x1FFFFFu) << 11)) >> 11;
  uint32_t idx = hash_key(bx, by, bz) >> shift;
  for (uint32_t p = 0; p <= mask; ++p) {
    if (atomicCAS(&table[idx].key, KEY_EMPTY, e.key) == KEY_EMPTY) {
      table[idx].slot = e.slot;
      return;
    }
    idx = (idx + 1) & mask;
  }
}

// TSDF payload -> {distance, weight} tile + 2-bit status tile.  One thread per voxel pair:
// PLANAR reads one float4 (two voxels), VOXBLOX reads 6 x u32 (block.cc:159-183 layout) and
// drops the colour word; both write one float4.  8 consecutive lanes own one 32-bit status
// word (16 voxels) and OR-reduce their 4-bit fields with shuffles.
template <int PACKING>
__global__ void k_scatter_tsdf(const void* __restrict__ payload, const int* __restrict__ slots, float2* tiles,
                               uint32_t* status, int nvox) {
  const int b = blockIdx.y;
  const int slot = slots[b];
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  const int npairs = nvox >> 1;
  const bool active = pair < npairs && slot >= 0;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    if (PACKING == 0) {
      v = __ldg(reinterpret_cast<const float4*>(payload) + (size_t)b * npairs + pair);
    } else {
      const uint2* src = reinterpret_cast<const uint2*>(payload) + ((size_t)b * npairs + pair) * 3;
      const uint2 a = __ldg(src), c = __ldg(src + 1), d = __ldg(src + 2);  // d0 w0 | c0 d1 | w1 c1
      v = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(c.y), __uint_as_float(d.x));
    }
    reinterpret_cast<float4*>(tiles + (size_t)slot * nvox)[pair] = v;
  }
  uint32_t bits = 0;
  if (active) bits = (tsdf_status(v.x, v.y) | (tsdf_status(v.z, v.w) << 2)) << ((pair & 7) * 4);
  bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
  bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
  bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
  if (active && (pair & 7) == 0) status[(size_t)slot * (nvox >> 4) + (pair >> 3)] = bits;
}

// ESDF payload -> float distance tile + 1 bit per voxel "observed" (EsdfVoxel::observed, voxel.h:21; the flag the
// interpolator's isObservedVoxel reads).  One thread per 4 voxels, 8 threads per 32-voxel word of the bit tile.
// VOXBLOX = 2 x u32 per voxel (distance bits, parent | flags with observed in bit 0: block.cc:203-234); the planar
// convenience packing carries no flags: a voxel is observed unless its distance is NaN.  checkMotion reads the
// distances only.
template <int PACKING>
__global__ void k_scatter_esdf(const void* __restrict__ payload, const int* __restrict__ slots, float* tiles, uint32_t* observed, int nvox) {
  const int b = blockIdx.y;
  const int slot = slots[b];
  const int quad = blockIdx.x * blockDim.x + threadIdx.x;
  const int nquads = nvox >> 2;
  const bool active = quad < nquads && slot >= 0;
  uint32_t bits = 0u;
  if (active) {
    float4 v;
    if (PACKING == 0) {
      v = __ldg(reinterpret_cast<const float4*>(payload) + (size_t)b * nquads + quad);
      bits = (isnan(v.x) ? 0u : 1u) | (isnan(v.y) ? 0u : 2u) | (isnan(v.z) ? 0u : 4u) | (isnan(v.w) ? 0u : 8u);
    } else {
      const uint4* src = reinterpret_cast<const uint4*>(payload) + ((size_t)b * nquads + quad) * 2;
      const uint4 a = __ldg(src), c = __ldg(src + 1);
      v = make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(c.x), __uint_as_float(c.z));
      bits = (a.y & 1u) | ((a.w & 1u) << 1) | ((c.y & 1u) << 2) | ((c.w & 1u) << 3);
    }
    reinterpret_cast<float4*>(tiles + (size_t)slot * nvox)[quad] = v;
  }
  // layers have 4, 8, 16 or 32 voxels per side: whole 32-voxel words, 8 consecutive threads each
  bits <<= (quad & 7) * 4;
  bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
  bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
  bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
  if (active && (quad & 7) == 0) observed[(size_t)slot * (nvox >> 5) + (quad >> 3)] = bits;
}

// Layer::getVoxelPtrByGlobalIndex (layer.h:215-239) + getVoxelStatusByIndex for test read-back.
__global__ void k_read_voxels(LayerView L, int kind, const long long* __restrict__ g, int n, float* out_values,
                              uint8_t* out_found, uint8_t* out_status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long gx = g[3 * i], gy = g[3 * i + 1], gz = g[3 * i + 2];
  const int vm = (1 << L.log2vps) - 1;
  const int slot = probe_block(L, (int)(gx >> L.log2vps), (int)(gy >> L.log2vps), (int)(gz >> L.log2vps));
  const int lin = ((int)gx & vm) | (((int)gy & vm) << L.log2vps) | (((int)gz & vm) << (2 * L.log2vps));
  float d = 0.f, w = 0.f;
  uint32_t st = 0;
  if (slot >= 0) {
    if (kind == 0) {
      const float2 v = reinterpret_cast<const float2*>(L.tiles)[(size_t)slot * L.nvox + lin];
      d = v.x;
      w = v.y;
      st = (L.status[(size_t)slot * (L.nvox >> 4) + (lin >> 4)] >> ((lin & 15) * 2)) & 3u;
    } else {
      d = reinterpret_cast<const float*>(L.tiles)[(size_t)slot * L.nvox + lin];
    }
  }
  out_values[2 * i] = d;
  out_values[2 * i + 1] = w;
  out_found[i] = slot >= 0 ? 1 : 0;
  if (out_status) out_status[i] = (uint8_t)st;
}

}  // namespace nbvgpu
#endif
