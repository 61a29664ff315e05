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
// Updates are batched: one kernel per commit, one CTA per changed block (atomicCAS claim + free-list pop, then the
// tile written with 128-bit loads/stores); nothing in the evaluation kernels ever writes the map.
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
// A commit is described by a table of 32-byte block descriptors that travels in front of the payload bytes (one
// host->device copy, one broadcast when there are several GPUs): which layer, which block, where its payload starts.
constexpr int kMaxLayers = 16;
enum { kOpUpdate = 0, kOpRemove = 1, kOpSkip = 2 };
struct __align__(16) BlockDesc {
  unsigned long long payload_off;  // byte offset of the block's payload from the payload base (multiple of 16)
  int32_t key[3];                  // BlockIndex
  int16_t layer;
  int8_t packing;                  // NBVGPU_PACK_*
  int8_t op;                       // kOpUpdate / kOpRemove / kOpSkip (superseded by a later entry of the same commit)
  uint32_t pad[2];
};
static_assert(sizeof(BlockDesc) == 32, "descriptor layout");

// Device-side handle of one layer for the update kernels.  state: [0] live blocks (u64); [1] low half = free_top,
// high half = err bits (1 = tile pool exhausted, 2 = table full); both as 32-bit ints at state + 1.
struct LayerDev {
  HashEntry* table;
  void* tiles;
  uint32_t* status;
  int* free_list;
  unsigned long long* state;
  uint32_t mask, shift;
  int nvox, kind;
};
struct ApplyLayers {
  LayerDev L[kMaxLayers];
};
__device__ __forceinline__ int* layer_free_top(const LayerDev& L) { return reinterpret_cast<int*>(L.state + 1); }
__device__ __forceinline__ int* layer_err(const LayerDev& L) { return reinterpret_cast<int*>(L.state + 1) + 1; }

// Layer::allocateBlockPtrByIndex (layer.h:119-134): tile slot of `key`, inserting it if absent.  A key that appears
// twice in one batch is claimed by one thread; the other waits until the winner has published the slot.
__device__ __forceinline__ int insert_or_lookup(const LayerDev& L, int bx, int by, int bz) {
  const unsigned long long key = pack_key(bx, by, bz);
  uint32_t idx = hash_key(bx, by, bz) >> L.shift;
  for (uint32_t p = 0; p <= L.mask; ++p) {
    HashEntry* e = L.table + idx;
    const unsigned long long prev = atomicCAS(&e->key, KEY_EMPTY, key);
    if (prev == KEY_EMPTY) {  // claimed a fresh entry: allocate a tile
      const int top = atomicSub(layer_free_top(L), 1) - 1;
      if (top < 0) {
        atomicAdd(layer_free_top(L), 1);
        atomicOr(layer_err(L), 1);
        atomicExch(&e->slot, -2);  // tell a waiting duplicate that there is no tile
        __threadfence();
        atomicExch(&e->key, KEY_TOMB);
        return -1;
      }
      const int slot = L.free_list[top];
      atomicExch(&e->slot, slot);
      atomicAdd(L.state, 1ull);
      return slot;
    }
    if (prev == key) {  // mapped by an earlier commit (slot >= 0 already) or being claimed right now by a duplicate
      int s;
      while ((s = *reinterpret_cast<volatile int*>(&e->slot)) == -1) {}
      return s >= 0 ? s : -1;
    }
    idx = (idx + 1) & L.mask;
  }
  atomicOr(layer_err(L), 2);
  return -1;
}

// TSDF payload -> {distance, weight} tile + 2-bit status tile.  One thread per voxel pair: PLANAR reads one float4
// (two voxels), VOXBLOX reads 6 x u32 (block.cc:159-183 layout) and drops the colour word; both write one float4.
// 8 consecutive lanes own one 32-bit status word (16 voxels) and OR-reduce their 4-bit fields with shuffles.
__device__ __forceinline__ void scatter_tsdf_block(const char* __restrict__ payload, int packing, int slot, float2* tiles,
                                                   uint32_t* status, int nvox, int tid, int nthreads) {
  const int npairs = nvox >> 1;
  for (int base = 0; base < npairs; base += nthreads) {  // warp-uniform trip count (npairs is a multiple of 32)
    const int pair = base + tid;
    const bool active = pair < npairs;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      if (packing == 0) {
        v = __ldg(reinterpret_cast<const float4*>(payload) + pair);
      } else {
        const uint2* src = reinterpret_cast<const uint2*>(payload) + (size_t)pair * 3;
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
}

// ESDF payload -> float distance tile + 1 bit per voxel "observed" (EsdfVoxel::observed, voxel.h:21; the flag the
// interpolator's isObservedVoxel reads).  One thread per 4 voxels, 8 threads per 32-voxel word of the bit tile.
// VOXBLOX = 2 x u32 per voxel (distance bits, parent | flags with observed in bit 0: block.cc:203-234); the planar
// convenience packing carries no flags: a voxel is observed unless its distance is NaN.  checkMotion reads the
// distances only.
__device__ __forceinline__ void scatter_esdf_block(const char* __restrict__ payload, int packing, int slot, float* tiles,
                                                   uint32_t* observed, int nvox, int tid, int nthreads) {
  const int nquads = nvox >> 2;
  for (int base = 0; base < nquads; base += nthreads) {
    const int quad = base + tid;
    const bool active = quad < nquads;  // 4^3 blocks have 16 quads: the upper half of the warp only takes part in the shuffles
    uint32_t bits = 0u;
    if (active) {
      float4 v;
      if (packing == 0) {
        v = __ldg(reinterpret_cast<const float4*>(payload) + quad);
        bits = (isnan(v.x) ? 0u : 1u) | (isnan(v.y) ? 0u : 2u) | (isnan(v.z) ? 0u : 4u) | (isnan(v.w) ? 0u : 8u);
      } else {
        const uint4* src = reinterpret_cast<const uint4*>(payload) + (size_t)quad * 2;
        const uint4 a = __ldg(src), c = __ldg(src + 1);
        v = make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(c.x), __uint_as_float(c.z));
        bits = (a.y & 1u) | ((a.w & 1u) << 1) | ((c.y & 1u) << 2) | ((c.w & 1u) << 3);
      }
      reinterpret_cast<float4*>(tiles + (size_t)slot * nvox)[quad] = v;
    }
    bits <<= (quad & 7) * 4;
    bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
    if (active && (quad & 7) == 0) observed[(size_t)slot * (nvox >> 5) + (quad >> 3)] = bits;
  }
}

// One commit (or one chunk of a large one) in one launch: CTA b inserts block b of the table into its layer's hash
// (or finds it there) and writes its tile(s).  Any mix of layers and packings; HBM-bound: the payload is read once with
// 128-bit loads and the tile written once with 128-bit stores.
constexpr int kApplyThreads = 256;
__global__ void __launch_bounds__(kApplyThreads) k_apply_blocks(const ApplyLayers layers, const BlockDesc* __restrict__ table,
                                                                const char* __restrict__ payload_base, int n) {
  __shared__ int s_slot;
  const int b = blockIdx.x;
  if (b >= n) return;
  const BlockDesc d = table[b];
  if (d.op != kOpUpdate || d.layer < 0 || d.layer >= kMaxLayers) return;
  const LayerDev& L = layers.L[d.layer];
  if (threadIdx.x == 0) s_slot = insert_or_lookup(L, d.key[0], d.key[1], d.key[2]);
  __syncthreads();
  const int slot = s_slot;
  if (slot < 0) return;
  const char* src = payload_base + d.payload_off;
  if (L.kind == 0) scatter_tsdf_block(src, d.packing, slot, reinterpret_cast<float2*>(L.tiles), L.status, L.nvox, threadIdx.x, blockDim.x);
  else scatter_esdf_block(src, d.packing, slot, reinterpret_cast<float*>(L.tiles), L.status, L.nvox, threadIdx.x, blockDim.x);
}

// Descriptor table for blocks whose keys and payload already sit in device memory (nbvgpu_layer_update_device).
__global__ void k_make_table(const int32_t* __restrict__ keys, int n, int layer, int packing, unsigned long long bytes_per_block,
                             BlockDesc* table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  BlockDesc d;
  d.payload_off = (unsigned long long)i * bytes_per_block;
  d.key[0] = keys[3 * i]; d.key[1] = keys[3 * i + 1]; d.key[2] = keys[3 * i + 2];
  d.layer = (int16_t)layer;
  d.packing = (int8_t)packing;
  // block indices outside the 21-bit key range would alias through pack_key: refused here as on the host path
  const bool ok = d.key[0] >= -(1 << 20) && d.key[0] < (1 << 20) && d.key[1] >= -(1 << 20) && d.key[1] < (1 << 20) &&
                  d.key[2] >= -(1 << 20) && d.key[2] < (1 << 20);
  d.op = ok ? kOpUpdate : kOpSkip;
  d.pad[0] = d.pad[1] = 0u;
  table[i] = d;
}

// Layer::removeBlock (layer.h:171-173) for the kOpRemove entries of a table.
__global__ void k_apply_remove(const ApplyLayers layers, const BlockDesc* __restrict__ table, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const BlockDesc d = table[i];
  if (d.op != kOpRemove || d.layer < 0 || d.layer >= kMaxLayers) return;
  const LayerDev& L = layers.L[d.layer];
  const int bx = d.key[0], by = d.key[1], bz = d.key[2];
  const unsigned long long key = pack_key(bx, by, bz);
  uint32_t idx = hash_key(bx, by, bz) >> L.shift;
  for (uint32_t p = 0; p <= L.mask; ++p) {
    HashEntry* e = L.table + idx;
    const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&e->key);
    if (k == key) {
      // claim the entry: a key listed twice in one batch (Layer::removeBlock of an absent block is a no-op in the
      // reference, so callers may do that) is freed by exactly one thread
      if (atomicCAS(&e->key, key, KEY_TOMB) != key) return;
      const int slot = atomicExch(&e->slot, -1);
      if (slot >= 0) {
        const int top = atomicAdd(layer_free_top(L), 1);
        L.free_list[top] = slot;
        atomicAdd(L.state, ~0ull);  // -1
      }
      return;
    }
    if (k == KEY_EMPTY) return;
    idx = (idx + 1) & L.mask;
  }
}

__global__ void k_reset_table(HashEntry* table, uint32_t cap, int* free_list, int max_blocks, unsigned long long* state) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) {
    table[i].key = KEY_EMPTY;
    table[i].slot = -1;
    table[i].pad = 0;
  }
  if (i < (uint32_t)max_blocks) free_list[i] = max_blocks - 1 - (int)i;  // pop order 0,1,2,...
  if (i == 0) {
    state[0] = 0ull;                                    // live blocks
    reinterpret_cast<int*>(state + 1)[0] = max_blocks;  // free_top
    reinterpret_cast<int*>(state + 1)[1] = 0;           // err
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
