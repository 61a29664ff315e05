// frontier_kernels.cuh -- History::bfs (nbveplanner/src/history.cpp:96-137) on the GPU map mirror:
// the frontier "potential" of a history-graph vertex = number of (visited free cell, unknown neighbour)
// faces within max_distance (6.0 m in the reference) and inside the exploration bounds, times voxel_size^3.
//
// What the reference does, and what therefore has to be reproduced for an identical count:
//   * cells are NOT voxels: a cell is a triple of ACCUMULATED doubles, new = parent + (+-vs, 0, 0) etc. (all
//     three components are added, so a -0.0 turns into +0.0), and two cells are the same iff their
//     std::to_string ("%f": six decimals, correctly rounded, ties to even, "-0.000000" for tiny negatives)
//     agree on all three coordinates;
//   * FIFO order with the neighbour order +x -x +y -y +z -z decides which arrival creates a cell and with
//     which doubles; the later tests (norm <= max_distance, bounds, status) are made on those doubles;
//   * the status comes from LowResManager::getVoxelStatus(position) (voxblox_manager.cpp:73-90): cast to
//     float, block by coordinates, voxel by coordinates truncated into the block;
//   * a free neighbour is inserted and queued, an unknown one is counted -- every time it is met --, the
//     start point is expanded whatever its own status.
//
// Device formulation: one CTA per vertex, level-synchronous.  A level's frontier is an array in FIFO order;
// its 6 * |frontier| (entry, direction) slots are evaluated in parallel.  The sequential order is kept by
// construction: a free cell reached by several slots of one level belongs to the smallest slot index
// (atomicMin on the arrival stamp), the next frontier is written in slot order (block scan), and an
// unknown neighbour is not counted when a smaller slot of the same level inserted that very key (what the
// sequential code would have seen as visited).  The string identity is a 64-bit key: lattice index (i, j, k)
// relative to the start (16 bits each), per axis the difference between the exactly rounded x * 10^6 of the
// accumulated double and of the canonical fma(i, vs, x0) (3 bits: paths differ by a few ulps, i.e. by at most
// one unit and only next to a rounding boundary), and per axis the "-0.000000" flag.
#ifndef NBVGPU_FRONTIER_KERNELS_CUH_
#define NBVGPU_FRONTIER_KERNELS_CUH_

#include <cstdint>
#include <cstring>
#include <type_traits>

#include "map_mirror.cuh"

namespace nbvgpu {

// round_half_even(|x| * 10^6), exactly: |x| = m * 2^(e-1075), 10^6 = 15625 * 2^6, so |x| * 10^6 is the 67-bit
// integer m * 15625 shifted right by 1069 - e.  ~0 for non-finite or |x| >= 2^21 (callers validate).
__host__ __device__ inline unsigned long long dec6_magnitude(double x) {
  unsigned long long bits;
#ifdef __CUDA_ARCH__
  bits = (unsigned long long)__double_as_longlong(x);
#else
  std::memcpy(&bits, &x, sizeof bits);
#endif
  const int e = (int)((bits >> 52) & 0x7ffu);
  if (e == 0) return 0ull;  // zero or subnormal
  const int sh = 1069 - e;
  if (e == 0x7ff || sh < 26) return ~0ull;
  if (sh >= 128) return 0ull;
  const unsigned long long m = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
  unsigned long long lo, hi;
#ifdef __CUDA_ARCH__
  lo = m * 15625ull;
  hi = __umul64hi(m, 15625ull);
#else
  const unsigned __int128 p = (unsigned __int128)m * 15625u;
  lo = (unsigned long long)p;
  hi = (unsigned long long)(p >> 64);
#endif
  unsigned long long q;
  bool gt, eq;  // discarded part greater than / equal to one half
  if (sh < 64) {
    q = (lo >> sh) | (hi << (64 - sh));
    const unsigned long long rem = lo & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
    gt = rem > half;
    eq = rem == half;
  } else if (sh == 64) {
    q = hi;
    gt = lo > (1ull << 63);
    eq = lo == (1ull << 63);
  } else {
    const int s2 = sh - 64;  // 1 .. 63
    q = hi >> s2;
    const unsigned long long rem_hi = hi & ((1ull << s2) - 1ull), half_hi = 1ull << (s2 - 1);
    gt = rem_hi > half_hi || (rem_hi == half_hi && lo != 0ull);
    eq = rem_hi == half_hi && lo == 0ull;
  }
  if (gt || (eq && (q & 1ull))) ++q;
  return q;
}

struct FrontierEntry {  // one queued cell: the doubles of its first arrival and its key
  double x, y, z;
  unsigned long long key;
};
struct FrontierCell {  // hash set of visited cells
  unsigned long long key;      // kFrontierEmpty = free
  unsigned long long arrival;  // (level << 32) | slot of the arrival that created it; smaller = earlier
};
constexpr unsigned long long kFrontierEmpty = ~0ull;

struct FrontierArgs {
  LayerView L;
  const double* points;  // [n][3]
  unsigned long long* frontier_voxels;  // [n]
  unsigned long long* visited;          // [n] or null
  unsigned long long* checksum;         // [n] or null: sum of cell_fingerprint over the visited cells
  int* status;                          // [n]: 0 ok, 1 workspace exhausted, 2 coordinate out of range,
                                        //      3 dense set met a split key: redo this vertex with the hash set
  const int* todo;                      // null, or [n]: the vertices this launch has to do (the others are left alone)
  // workspace of CTA b: cells + b * cell_cap, cur/next + b * frontier_cap, slot_eval + b * 6 * frontier_cap
  FrontierCell* cells;
  FrontierEntry* cur;
  FrontierEntry* next;
  int* slot_eval;
  unsigned cell_cap;      // power of two (hash set)
  unsigned long long* grid;  // dense visited set: workspace of CTA b at grid + b * grid_dim^3
  int grid_r;                // lattice radius covered: indices -grid_r .. grid_r per axis (dense set and canon tables)
  unsigned frontier_cap;
  double vs;              // float voxel size widened (LowResManager::getResolution)
  double max_distance;
  double max_d2;  // largest double x with sqrt_rn(x) <= max_distance
  double bmin[3], bmax[3];
  float block_size, block_size_inv, voxel_size_inv;
  int vps;
  int n;
};

// Order-independent fingerprint of a visited cell (the bit patterns of its three doubles); summed per vertex
// as a diagnostic: equal sums mean the same doubles in every cell, i.e. the same first arrivals as the
// sequential code (the oracle computes the same function).
__device__ __forceinline__ unsigned long long cell_fingerprint(double x, double y, double z) {
  const unsigned long long b[3] = {(unsigned long long)__double_as_longlong(x), (unsigned long long)__double_as_longlong(y),
                                   (unsigned long long)__double_as_longlong(z)};
  unsigned long long h = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    h ^= b[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
  }
  return h;
}

__device__ __forceinline__ unsigned frontier_hash(unsigned long long k) {
  k ^= k >> 29;
  k *= 0xBF58476D1CE4E5B9ull;
  k ^= k >> 32;
  return (unsigned)k;
}

// Key fields.  Lattice index per axis biased by 32768 in bits 16a..16a+15, rounding variant + 4 in bits
// 48+3a..50+3a, "-0.000000" flag in bit 57+a.
__device__ __forceinline__ unsigned long long key_axis_fields(int a, int lattice, int variant, bool negzero) {
  return ((unsigned long long)(unsigned)(lattice + 32768) << (16 * a)) | ((unsigned long long)(unsigned)(variant + 4) << (48 + 3 * a)) |
         ((unsigned long long)(negzero ? 1u : 0u) << (57 + a));
}
__device__ __forceinline__ unsigned long long key_axis_mask(int a) {
  return (0xFFFFull << (16 * a)) | (7ull << (48 + 3 * a)) | (1ull << (57 + a));
}
__device__ __forceinline__ int key_lattice(unsigned long long key, int a) { return (int)((key >> (16 * a)) & 0xFFFFu) - 32768; }

// Canonical value of lattice index i on one axis: the double an axis-monotone path arrives with, x0 + vs + vs + ...
// (or + (-vs) ...), i.e. what almost every first arrival holds, and its rounded decimal.  One table per vertex and
// axis in shared memory, indices -r .. r.
struct AxisCanon {
  double value;
  unsigned long long mag;  // dec6_magnitude(value)
};

// Fields of one axis for the accumulated coordinate v with lattice index i: the rounding variant is the difference
// between the exactly rounded v * 10^6 and the canonical one -- zero without any arithmetic when v IS the canonical
// double (bitwise), computed exactly otherwise.  false if it does not fit (cannot happen: reported, never ignored).
__device__ __forceinline__ bool axis_fields(int a, double v, int lattice, const AxisCanon* canon, int r, unsigned long long* out) {
  if (lattice < -r || lattice > r) return false;
  const AxisCanon c = canon[(size_t)a * (2 * r + 1) + (lattice + r)];
  const long long vb = __double_as_longlong(v), cb = __double_as_longlong(c.value);
  const bool neg = vb < 0;  // sign bit: "-0.000000" also for -0.0 and tiny negatives
  unsigned long long mag = c.mag;
  long long variant = 0;
  if (vb != cb) {
    mag = dec6_magnitude(v);
    if (mag == ~0ull) return false;
    const long long sv = neg ? -(long long)mag : (long long)mag, sc = cb < 0 ? -(long long)c.mag : (long long)c.mag;
    variant = sv - sc;
    if (variant < -3 || variant > 3) return false;
  }
  if (c.mag == ~0ull) return false;
  *out = key_axis_fields(a, lattice, (int)variant, neg && mag == 0ull);
  return true;
}

// LowResManager::getVoxelStatus(position) (voxblox_manager.cpp:73-90) on the mirror: 0 unknown, 1 free, 2 occupied.
__device__ __forceinline__ unsigned status_at(const FrontierArgs& a, double x, double y, double z) {
  const double p[3] = {x, y, z};
  int b[3], l[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float pf = __double2float_rn(p[k]);                                                   // position.cast<float>()
    b[k] = (int)floorf(__fadd_rn(__fmul_rn(pf, a.block_size_inv), 1e-6f));                      // common.h: getGridIndexFromPoint
    const float origin = __fmul_rn((float)b[k], a.block_size);                                  // common.h:196-201
    const int v = (int)floorf(__fadd_rn(__fmul_rn(__fsub_rn(pf, origin), a.voxel_size_inv), 1e-6f));
    l[k] = max(min(v, a.vps - 1), 0);                                                           // block_inl.h:30-40
  }
  const int slot = probe_block(a.L, b[0], b[1], b[2]);
  if (slot < 0) return 0u;
  const unsigned lin = (unsigned)(l[0] + a.vps * (l[1] + a.vps * l[2]));
  return (__ldg(a.L.status + ((size_t)slot * (a.L.nvox >> 4) + (lin >> 4))) >> ((lin & 15u) * 2u)) & 3u;
}

// ---- visited set ------------------------------------------------------------------------------
// Two implementations with one interface.  `stamp` = (level << kStampSlotBits) | slot orders arrivals; an entry created in
// an earlier level has a smaller stamp than any stamp of the current one.
//   (stamp = (level << 26) | slot)
//   seen_before(key, level_stamp)  was this key inserted in an earlier level (or is it the start)?
//   insert(key, stamp)             create the entry or lower its stamp; returns a handle (< 0: no room)
//   owner(handle) / key_of(handle) after the level's barrier: the winning stamp and the key
//   earlier(key, stamp)            after the barrier: was this key inserted with a smaller stamp?
// VisitedHash: open addressing on the full 64-bit key -- exact in every case.
// VisitedDense: one 8-byte word per lattice cell of the cube |i|,|j|,|k| <= R: (stamp << 12) | the key's 12
// rounding / zero-sign bits; one access, no probing, and a wavefront only touches a shell of the cube.  A
// lattice cell can hold one key only: meeting an entry whose 12 bits differ (the same lattice cell under two
// different strings -- next to a "%f" rounding boundary, or -0.0 against +0.0) raises `split`, and the vertex
// is redone with VisitedHash.
constexpr int kStampSlotBits = 26;  // slots per level < 2^26; the dense word keeps 64 - 12 - 26 bits for the level

struct VisitedHash {
  FrontierCell* cells;
  unsigned mask;
  int* split;  // unused
  __device__ void clear(int tid, int nthreads) const {
    for (unsigned i = tid; i <= mask; i += nthreads) cells[i] = FrontierCell{kFrontierEmpty, ~0ull};
  }
  __device__ void insert_start(unsigned long long key) const { cells[frontier_hash(key) & mask] = FrontierCell{key, 0ull}; }
  __device__ bool seen_before(unsigned long long key, unsigned long long level_stamp) const {
    unsigned h = frontier_hash(key) & mask;
    for (unsigned n = 0; n <= mask; ++n) {
      const unsigned long long k = cells[h].key;
      if (k == key) return cells[h].arrival < level_stamp;
      if (k == kFrontierEmpty) return false;
      h = (h + 1) & mask;
    }
    return false;
  }
  __device__ int insert(unsigned long long key, unsigned long long stamp) const {
    unsigned h = frontier_hash(key) & mask;
    for (unsigned n = 0; n <= mask; ++n) {
      unsigned long long k = cells[h].key;
      if (k == kFrontierEmpty) k = atomicCAS(&cells[h].key, kFrontierEmpty, key);
      if (k == kFrontierEmpty || k == key) {
        atomicMin(&cells[h].arrival, stamp);
        return (int)h;
      }
      h = (h + 1) & mask;
    }
    return -1;
  }
  __device__ bool wins(int handle, unsigned long long, unsigned long long stamp) const { return cells[handle].arrival == stamp; }
  __device__ bool earlier(unsigned long long key, unsigned long long stamp) const { return seen_before(key, stamp); }
};

struct VisitedDense {
  unsigned long long* grid;
  int r, d;    // radius, edge = 2 r + 1
  int* split;  // shared flag
  __device__ void clear(int tid, int nthreads) const {
    const unsigned n = (unsigned)d * d * d;
    for (unsigned i = tid; i < n; i += nthreads) grid[i] = ~0ull;
  }
  __device__ int index(unsigned long long key) const {
    return ((key_lattice(key, 2) + r) * d + (key_lattice(key, 1) + r)) * d + (key_lattice(key, 0) + r);
  }
  __device__ static unsigned long long pack(unsigned long long stamp, unsigned long long key) { return (stamp << 12) | ((key >> 48) & 0xFFFull); }
  __device__ void insert_start(unsigned long long key) const { grid[index(key)] = pack(0ull, key); }
  __device__ bool seen_before(unsigned long long key, unsigned long long level_stamp) const {
    const unsigned long long e = grid[index(key)];
    if (e == ~0ull) return false;
    if ((e & 0xFFFull) != ((key >> 48) & 0xFFFull)) {
      *split = 1;
      return false;
    }
    return (e >> 12) < level_stamp;
  }
  __device__ int insert(unsigned long long key, unsigned long long stamp) const {
    const int i = index(key);
    atomicMin(&grid[i], pack(stamp, key));
    return i;
  }
  __device__ bool wins(int handle, unsigned long long key, unsigned long long stamp) const {
    const unsigned long long e = grid[handle];
    if ((e & 0xFFFull) != ((key >> 48) & 0xFFFull)) *split = 1;  // another string took this lattice cell in this level
    return e == pack(stamp, key);
  }
  __device__ bool earlier(unsigned long long key, unsigned long long stamp) const { return seen_before(key, stamp); }
};

__device__ __forceinline__ int vis_index(const VisitedDense& v, unsigned long long key) { return v.index(key); }
__device__ __forceinline__ int vis_index(const VisitedHash&, unsigned long long) { return -1; }  // never used: the hash set keeps handles

constexpr int kFrontierThreads = 1024;
// slot_eval codes
constexpr int kSlotNone = -1, kSlotUnknown = -2;

template <bool DENSE>
__global__ void __launch_bounds__(kFrontierThreads, 1) k_frontier_bfs(const FrontierArgs a) {
  __shared__ unsigned s_warp[2][32];  // per-warp counts / offsets of the resolve scan, double-buffered by chunk parity
  __shared__ unsigned s_total[2];
  __shared__ int s_err, s_split;
  __shared__ unsigned long long s_count, s_sum;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  AxisCanon* const canon = reinterpret_cast<AxisCanon*>(s_dyn);  // [3][2 r + 1]
  const int cr = a.grid_r, cd = 2 * a.grid_r + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  typedef typename std::conditional<DENSE, VisitedDense, VisitedHash>::type Visited;
  Visited vis;
  if constexpr (DENSE) {
    const int d = 2 * a.grid_r + 1;
    vis = VisitedDense{a.grid + (size_t)blockIdx.x * d * d * d, a.grid_r, d, &s_split};
  } else {
    vis = VisitedHash{a.cells + (size_t)blockIdx.x * a.cell_cap, a.cell_cap - 1, &s_split};
  }
  FrontierEntry* cur = a.cur + (size_t)blockIdx.x * a.frontier_cap;
  FrontierEntry* next = a.next + (size_t)blockIdx.x * a.frontier_cap;
  // per-slot outcome of the expand pass: the hash set keeps its entry handle (or a negative code); the dense set only a
  // byte (0 nothing, 1 unknown neighbour, 2 free candidate) -- its handle is the grid index, recomputed from the key
  typedef typename std::conditional<DENSE, signed char, int>::type SlotT;
  SlotT* const slot_eval = reinterpret_cast<SlotT*>(a.slot_eval + (size_t)blockIdx.x * 6 * a.frontier_cap);

  for (int v = blockIdx.x; v < a.n; v += gridDim.x) {
    if (a.todo && !a.todo[v]) continue;
    const double x0 = a.points[3 * v], y0 = a.points[3 * v + 1], z0 = a.points[3 * v + 2];
    vis.clear(tid, blockDim.x);
    if (tid == 0) {
      s_err = 0;
      s_split = 0;
      s_count = 0ull;
      s_sum = cell_fingerprint(x0, y0, z0);
    }
    __syncthreads();
    unsigned n_cur = 1, n_visited = 1;
    // canonical tables: six threads accumulate outwards from the start (sequential adds, as a monotone path does),
    // then everybody rounds
    if (tid < 6) {
      const int axis = tid >> 1;
      const double step = (tid & 1) ? -a.vs : a.vs;
      double v = axis == 0 ? x0 : (axis == 1 ? y0 : z0);
      AxisCanon* t = canon + (size_t)axis * cd + cr;
      if ((tid & 1) == 0) t[0].value = v;
      for (int i = 1; i <= cr; ++i) {
        v = __dadd_rn(v, step);
        t[(tid & 1) ? -i : i].value = v;
      }
    }
    __syncthreads();
    for (int i = tid; i < 3 * cd; i += blockDim.x) canon[i].mag = dec6_magnitude(canon[i].value);
    __syncthreads();
    if (tid == 0) {
      unsigned long long fx = 0, fy = 0, fz = 0;
      const bool ok = axis_fields(0, x0, 0, canon, cr, &fx) && axis_fields(1, y0, 0, canon, cr, &fy) && axis_fields(2, z0, 0, canon, cr, &fz);
      if (!ok) {
        s_err = 2;
      } else {
        const unsigned long long key = fx | fy | fz;
        cur[0] = FrontierEntry{x0, y0, z0, key};
        vis.insert_start(key);  // visited.insert(convertToString(point))
      }
    }
    __syncthreads();
    unsigned long long my_count = 0ull, my_sum = 0ull;
    for (unsigned level = 1; n_cur > 0 && s_err == 0 && s_split == 0; ++level) {
      const unsigned n_slots = 6u * n_cur;
      const unsigned long long stamp0 = (unsigned long long)level << kStampSlotBits;
      // ---- expand: every (entry, direction) of this level
      for (unsigned s = tid; s < n_slots; s += blockDim.x) {
        const FrontierEntry e = cur[s / 6u];
        const unsigned d = s % 6u;
        const int axis = (int)(d >> 1);
        const double step = (d & 1u) ? -a.vs : a.vs;
        // candidate = new_pos + c: all three components are added (Eigen), the idle ones with +0.0
        const double cx = __dadd_rn(e.x, axis == 0 ? step : 0.0), cy = __dadd_rn(e.y, axis == 1 ? step : 0.0),
                     cz = __dadd_rn(e.z, axis == 2 ? step : 0.0);
        int ev = kSlotNone;
        // key: the moved axis is recomputed; an idle axis keeps its fields unless it was -0.0 (now +0.0)
        unsigned long long key = e.key;
        const double cv = axis == 0 ? cx : (axis == 1 ? cy : cz);
        const int lattice = key_lattice(key, axis) + ((d & 1u) ? -1 : 1);
        unsigned long long f = 0;
        bool ok = axis_fields(axis, cv, lattice, canon, cr, &f);
        if (DENSE && ok && (lattice < -a.grid_r || lattice > a.grid_r)) {
          s_split = 1;  // outside the cube (cannot happen for R = floor(max_distance / vs) + 2): let the hash set do this vertex
          ok = false;
        }
        key = (key & ~key_axis_mask(axis)) | f;
        if (axis != 0 && e.x == 0.0) key &= ~(1ull << 57);
        if (axis != 1 && e.y == 0.0) key &= ~(1ull << 58);
        if (axis != 2 && e.z == 0.0) key &= ~(1ull << 59);
        if (!ok) {
          if (!(DENSE && s_split)) s_err = 2;
        } else {
          // visited.find(...) as of the previous levels (same-level arrivals are arbitrated below)
          const bool seen = vis.seen_before(key, stamp0);
          if (!seen) {
            // history.cpp:118-121: (candidate - point).norm() <= max_distance and inside the bounds.  The square root is
            // correctly rounded and monotone, so norm <= max_distance  <=>  squared sum <= max_d2, the largest double whose
            // root does not exceed max_distance (found by the host); the cheap bounds test goes first (both are pure).
            bool near = cx >= a.bmin[0] && cx <= a.bmax[0] && cy >= a.bmin[1] && cy <= a.bmax[1] && cz >= a.bmin[2] && cz <= a.bmax[2];
            if (near) {
              const double dx = __dsub_rn(cx, x0), dy = __dsub_rn(cy, y0), dz = __dsub_rn(cz, z0);
              near = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) <= a.max_d2;
            }
            if (near) {
              const unsigned st = status_at(a, cx, cy, cz);
              if (st == 0u) {
                ev = kSlotUnknown;
              } else if (st == 1u) {
                // insert or find, earliest arrival wins
                ev = vis.insert(key, stamp0 | s);
                if (ev < 0) {
                  ev = kSlotNone;
                  s_err = 1;
                }
              }
            }
          }
        }
        slot_eval[s] = DENSE ? (SlotT)(ev >= 0 ? 2 : (ev == kSlotUnknown ? 1 : 0)) : (SlotT)ev;
      }
      __syncthreads();
      // ---- resolve in slot order: winners form the next frontier, unknown neighbours are counted
      unsigned run = 0u;  // winners before this chunk: the same value in every thread
      for (unsigned base = 0, par = 0; base < n_slots; base += blockDim.x, par ^= 1u) {
        const unsigned s = base + tid;
        bool win = false;
        double cx = 0, cy = 0, cz = 0;
        unsigned long long key = 0;
        if (s < n_slots) {
          const int raw = slot_eval[s];
          const bool is_free = DENSE ? raw == 2 : raw >= 0, is_unknown = DENSE ? raw == 1 : raw == kSlotUnknown;
          if (is_free || is_unknown) {
            const FrontierEntry e = cur[s / 6u];
            const unsigned d = s % 6u;
            const int axis = (int)(d >> 1);
            const double step = (d & 1u) ? -a.vs : a.vs;
            cx = __dadd_rn(e.x, axis == 0 ? step : 0.0);
            cy = __dadd_rn(e.y, axis == 1 ? step : 0.0);
            cz = __dadd_rn(e.z, axis == 2 ? step : 0.0);
            // the candidate's key, as in the expand pass
            key = e.key;
            const double cv = axis == 0 ? cx : (axis == 1 ? cy : cz);
            unsigned long long f = 0;
            axis_fields(axis, cv, key_lattice(key, axis) + ((d & 1u) ? -1 : 1), canon, cr, &f);
            key = (key & ~key_axis_mask(axis)) | f;
            if (axis != 0 && e.x == 0.0) key &= ~(1ull << 57);
            if (axis != 1 && e.y == 0.0) key &= ~(1ull << 58);
            if (axis != 2 && e.z == 0.0) key &= ~(1ull << 59);
            if (is_free) {
              win = vis.wins(DENSE ? vis_index(vis, key) : raw, key, stamp0 | s);
            } else if (!vis.earlier(key, stamp0 | s)) {
              ++my_count;  // unknown neighbour: counted unless an earlier slot of this level inserted this very key
            }
          }
        }
        // exclusive scan of `win` over the chunk, in slot order
        const unsigned ballot = __ballot_sync(0xffffffffu, win);
        if (lane == 0) s_warp[par][warp] = __popc(ballot);
        __syncthreads();
        if (warp == 0) {
          unsigned w = s_warp[par][lane];
          unsigned incl = w;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          s_warp[par][lane] = incl - w;
          if (lane == 31) s_total[par] = incl;
        }
        __syncthreads();
        if (win) {
          const unsigned pos = run + s_warp[par][warp] + __popc(ballot & ((1u << lane) - 1u));
          if (pos < a.frontier_cap) next[pos] = FrontierEntry{cx, cy, cz, key};
          my_sum += cell_fingerprint(cx, cy, cz);
        }
        run += s_total[par];
        // no barrier here: the next chunk writes the other half of the scratch, and its first barrier orders this
        // chunk's reads before the writes of the chunk after that
      }
      __syncthreads();
      const unsigned n_next = run;
      if (n_next > a.frontier_cap) {
        if (tid == 0) s_err = 1;
      }
      __syncthreads();
      n_visited += n_next;
      n_cur = n_next;
      FrontierEntry* t = cur;
      cur = next;
      next = t;
    }
    // ---- vertex result
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      my_count += __shfl_xor_sync(0xffffffffu, my_count, o);
      my_sum += __shfl_xor_sync(0xffffffffu, my_sum, o);
    }
    if (lane == 0 && my_count) atomicAdd(&s_count, my_count);
    if (lane == 0 && my_sum) atomicAdd(&s_sum, my_sum);
    __syncthreads();
    if (tid == 0) {
      a.frontier_voxels[v] = s_count;
      if (a.visited) a.visited[v] = n_visited;
      if (a.checksum) a.checksum[v] = s_sum;
      a.status[v] = s_err ? s_err : (s_split ? 3 : 0);
    }
    __syncthreads();
  }
}

}  // namespace nbvgpu

#endif  // NBVGPU_FRONTIER_KERNELS_CUH_
