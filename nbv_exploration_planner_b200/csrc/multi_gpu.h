// multi_gpu.h -- plumbing of the multi-GPU contexts (nbvgpu_create_multi / nbvgpu_create_rank): NCCL resolved at run
// time, and the block of host memory through which the ranks of a context exchange what is too small for a collective.
//
// The path shards by candidates: every GPU holds a replica of the map mirror and evaluates a contiguous
// block of the batch.  Two exchanges surround it:
//   * nbvgpu_commit: the root copies {descriptor table | payload} of the changed blocks to its GPU and ncclBroadcasts
//     that one buffer over NVLink; every GPU then applies it with the same kernel (csrc/map_mirror.cuh).  Large chunks whose
//     payload lies in memory every rank can read (nbvgpu_shared_alloc) are PULLED instead: every GPU copies 1/world of the
//     payload over its own PCIe link and an ncclAllGather over NVLink completes the chunk on every GPU -- the host links of
//     all GPUs work side by side instead of rank 0's alone;
//   * the per-shard best {score, index, gain, yaw}: each GPU's scoring kernel stores its 32-byte record straight into
//     host memory that every rank can read (page-locked + mapped: cudaHostAlloc in one process, a POSIX shared-memory
//     segment registered with cudaHostRegister across processes), followed by a ticket; the host polls the tickets.
//     No collective, no copy operation and no stream synchronisation on the per-step path.
#ifndef NBVGPU_MULTI_GPU_H_
#define NBVGPU_MULTI_GPU_H_

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>  // types and enums only: the functions are resolved with dlsym (see NcclApi)
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

namespace nbvgpu {

// ---- NCCL, loaded at run time ----------------------------------------------------------------------------------------
// The library must not pull a second NCCL into a process that already has one (PyTorch ships its own libnccl.so.2), and a
// single-GPU user should not need NCCL at all: the handful of entry points used are looked up in whatever libnccl.so.2 the
// process has loaded, else in NBVGPU_NCCL_LIB, else in the system's.
struct NcclApi {
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string error;
};

inline const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (const char* p = std::getenv("NBVGPU_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process (e.g. PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
      return;
    }
    bool all = true;
    auto sym = [&](const char* name) -> void* {
      void* f = dlsym(h, name);
      if (!f) {
        all = false;
        api.error = std::string("NCCL symbol missing: ") + name;
      }
      return f;
    };
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = all;
  });
  return api;
}

// ---- the shared block ------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 64;
constexpr int kMaxChunks = 2048;
constexpr unsigned kSharedMagic = 0x4e425647u;  // "NBVG"

// One rank's best-node record of one evaluation (written by the GPU, read by every rank's host).
struct alignas(64) BestSlot {
  long long index;  // GLOBAL candidate index, -1 if the shard has no candidate above zero_gain
  double score, gain, yaw;
  unsigned ticket;  // sequence number of the evaluation the record belongs to; written last
  unsigned pad[7];
};
static_assert(sizeof(BestSlot) == 64, "slot layout");

// What the root tells the other ranks about a commit (they cannot see its staging area).
constexpr int kMaxPullSegs = 4;
constexpr int kMaxArenas = 256;  // allocations per context lifetime (a freed number is not handed out again)
struct PullSeg {                 // a piece of a chunk's payload that lies in a shared arena (nbvgpu_shared_alloc)
  unsigned arena, pad;
  unsigned long long src_off;    // offset inside the arena
  unsigned long long bytes;
  unsigned long long dst_off;    // offset inside the chunk's payload
};
struct ChunkInfo {
  unsigned long long bytes;  // table + payload bytes of the chunk (what is broadcast; pulled chunks: table + world * slice)
  unsigned table_bytes;      // payload starts here
  unsigned n_update;         // descriptors to apply
  unsigned n_remove;         // > 0: the chunk is the commit's removal table (no payload)
  unsigned n_seg;            // > 0: pulled chunk -- every rank copies its slice of these pieces, an all-gather does the rest
  unsigned long long slice;  // payload bytes per rank of a pulled chunk (a multiple of 256)
  unsigned long long payload;  // payload bytes of the chunk
  PullSeg seg[kMaxPullSegs];
};
struct SharedBlock {
  unsigned magic, world;
  std::atomic<unsigned> attached;  // ranks that have mapped the block
  unsigned pad0[13];
  BestSlot best[2][kMaxRanks];     // alternating by evaluation parity: a fast rank may start evaluation s+1 while a slow one
                                   // still reads the records of evaluation s
  unsigned seq_in[kMaxRanks];      // written by each rank's host before it launches evaluation `seq`, read by its scoring kernel
  // commits
  alignas(64) std::atomic<unsigned long long> commit_seq;   // published by the root after the plan below is complete
  alignas(64) std::atomic<unsigned long long> commit_ack[kMaxRanks];  // last commit each rank has read the plan of
  unsigned n_chunks;
  unsigned incoming[16];           // blocks arriving per layer (steers the hash-table rebuild the same way on every rank)
  // shared arenas (nbvgpu_shared_alloc across processes): the root creates segment number i and publishes its size, the others map it
  alignas(64) std::atomic<unsigned> arena_created;            // segments created so far
  unsigned long long arena_bytes[kMaxArenas];
  std::atomic<unsigned> arena_attached[kMaxArenas];           // ranks that have mapped segment i and tried to page-lock it
  std::atomic<unsigned> arena_failed[kMaxArenas];             // ... of which this many could not: then every rank gives the segment up
  ChunkInfo chunk[kMaxChunks];
};

struct SharedMap {
  SharedBlock* host = nullptr;
  bool posix = false, registered = false, owner = false;
  std::string name;
};

// One process: plain page-locked memory that every device of the process can write.
inline cudaError_t shared_alloc_local(SharedMap* m, int world) {
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, sizeof(SharedBlock), cudaHostAllocPortable | cudaHostAllocMapped);
  if (e != cudaSuccess) return e;
  memset(p, 0, sizeof(SharedBlock));
  m->host = new (p) SharedBlock;
  m->host->magic = kSharedMagic;
  m->host->world = (unsigned)world;
  m->posix = false;
  return cudaSuccess;
}

// Several processes: a POSIX shared-memory segment named after the job's NCCL unique id.  Rank 0 creates it before the
// communicator is initialised (ncclCommInitRank synchronises the ranks, so the others open it afterwards and find it
// there); every rank registers its mapping with CUDA so that its GPU can write into it.
inline std::string shared_name(const void* unique_id, size_t len) {
  unsigned long long h = 1469598103934665603ull;  // FNV-1a over the id bytes
  for (size_t i = 0; i < len; ++i) h = (h ^ reinterpret_cast<const unsigned char*>(unique_id)[i]) * 1099511628211ull;
  char buf[64];
  snprintf(buf, sizeof buf, "/nbvgpu_%016llx", h);
  return buf;
}
inline bool shared_create(SharedMap* m, const std::string& name, int world, std::string* err) {
  shm_unlink(name.c_str());
  const int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) {
    *err = "shm_open(create) failed for " + name;
    return false;
  }
  if (ftruncate(fd, sizeof(SharedBlock)) != 0) {
    close(fd);
    shm_unlink(name.c_str());
    *err = "ftruncate failed for " + name;
    return false;
  }
  void* p = mmap(nullptr, sizeof(SharedBlock), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) {
    shm_unlink(name.c_str());
    *err = "mmap failed for " + name;
    return false;
  }
  memset(p, 0, sizeof(SharedBlock));
  m->host = new (p) SharedBlock;
  m->host->world = (unsigned)world;
  m->host->attached.store(1u);
  std::atomic_thread_fence(std::memory_order_release);
  m->host->magic = kSharedMagic;
  m->posix = true;
  m->owner = true;
  m->name = name;
  return true;
}
inline bool shared_open(SharedMap* m, const std::string& name, std::string* err) {
  int fd = -1;
  for (int tries = 0; tries < 2000 && fd < 0; ++tries) {  // the creator ran before the communicator came up
    fd = shm_open(name.c_str(), O_RDWR, 0600);
    if (fd < 0) std::this_thread::sleep_for(std::chrono::milliseconds(5));
  }
  if (fd < 0) {
    *err = "shm_open failed for " + name;
    return false;
  }
  void* p = mmap(nullptr, sizeof(SharedBlock), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) {
    *err = "mmap failed for " + name;
    return false;
  }
  m->host = reinterpret_cast<SharedBlock*>(p);
  for (int tries = 0; tries < 2000 && m->host->magic != kSharedMagic; ++tries) std::this_thread::sleep_for(std::chrono::milliseconds(1));
  if (m->host->magic != kSharedMagic) {
    *err = "shared block " + name + " was never initialised";
    return false;
  }
  m->host->attached.fetch_add(1u);
  m->posix = true;
  m->name = name;
  return true;
}
inline cudaError_t shared_register(SharedMap* m) {
  if (!m->posix) return cudaSuccess;
  cudaError_t e = cudaHostRegister(m->host, sizeof(SharedBlock), cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e == cudaSuccess) m->registered = true;
  return e;
}
inline void shared_release(SharedMap* m) {
  if (!m->host) return;
  if (m->posix) {
    if (m->registered) cudaHostUnregister(m->host);
    munmap(m->host, sizeof(SharedBlock));
    if (m->owner) shm_unlink(m->name.c_str());
  } else {
    cudaFreeHost(m->host);
  }
  m->host = nullptr;
}

// A named POSIX shared-memory segment of `bytes` bytes: created (create = true) or opened and mapped; null on failure.
inline void* shm_map_segment(const std::string& name, size_t bytes, bool create, std::string* err) {
  int fd = -1;
  if (create) {
    shm_unlink(name.c_str());
    fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd >= 0 && ftruncate(fd, (off_t)bytes) != 0) {
      close(fd);
      shm_unlink(name.c_str());
      fd = -1;
    }
  } else {
    for (int tries = 0; tries < 2000 && fd < 0; ++tries) {
      fd = shm_open(name.c_str(), O_RDWR, 0600);
      if (fd < 0) std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
  }
  if (fd < 0) {
    *err = std::string(create ? "shm_open(create)" : "shm_open") + " failed for " + name;
    return nullptr;
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) {
    if (create) shm_unlink(name.c_str());
    *err = "mmap failed for " + name;
    return nullptr;
  }
  return p;
}

// Spin until pred() holds; false after timeout_s seconds.
template <class Pred>
inline bool spin_until(Pred pred, double timeout_s) {
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned it = 1;; ++it) {
    if (pred()) return true;
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
    if ((it & 0xffffu) == 0u && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return false;
  }
}

}  // namespace nbvgpu
#endif
