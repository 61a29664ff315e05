/* nbvgpu.h -- C ABI of the B200-native viewpoint-gain path.
 *
 * Drop-in boundary for the three reference call sites that make up the hot path
 * of RaccoonlabDev/nbv_exploration_planner (citations relative to the reference
 * tree).  The reference has no plugin / FFI interface: the path is three C++
 * member functions reached through concrete pointers, so each entry point below
 * names the member function / call site it replaces.
 *
 *   RrtTree::optimized_gain(Pose&)        nbveplanner/src/rrt.cpp:351-479   -> nbvgpu_gain_eval*, nbvgpu_optimized_gain
 *   RrtTree::gain(Pose&)                  nbveplanner/src/rrt.cpp:481-579   -> same, camera.sum_all_yaws = 1
 *   HighResManager::checkMotion<T>        nbveplanner/include/nbveplanner/voxblox_manager.h:89-108
 *                                                                           -> nbvgpu_check_segments*, nbvgpu_check_motion
 *   History::bfs(point)                   nbveplanner/src/history.cpp:96-137 -> nbvgpu_frontier_potential
 *   HighResManager::getDistanceAndGradientAtPosition   nbveplanner/src/voxblox_manager.cpp:143-148
 *                                                                           -> nbvgpu_esdf_distance_gradient
 *   VoxbloxManager::getVoxelStatusByIndex nbveplanner/src/voxblox_manager.cpp:14-33 (map read side)
 *   voxblox::Layer<VoxelType> block hash  voxblox/voxblox/include/voxblox/core/layer.h:23-296,
 *                                         core/block_hash.h:20-31           -> nbvgpu_layer_* (GPU mirror)
 *   node scoring in RrtTree::iterate      nbveplanner/src/rrt.cpp:220-246   -> nbvgpu_gain_eval_best*
 *
 * Conventions: plain pointers and sizes only; the caller owns every host buffer;
 * the library owns device memory, pinned staging and its CUDA stream.  All
 * functions return an nbvgpu_status (0 = OK, < 0 = error); nothing throws across
 * the ABI and there is NO CPU fallback: without a usable CUDA device
 * nbvgpu_create fails with NBVGPU_ERR_NO_DEVICE.
 *
 * Threading (mirrors the reference): gain_eval from the planner
 * thread, check_segments from planner and history threads, layer_update/commit
 * from the map (ROS spin) thread -- concurrently.  The host-buffer evaluators are
 * re-entrant: each holds the context's snapshot lock shared and runs on a lane of
 * its own (the context's stream and staging when free, else a lane context with
 * its own high-priority stream), so a checkMotion does not queue behind a gain
 * batch.  layer_update / layer_remove only stage and run beside everything;
 * commit, layer_create / clear, set_camera, set_stream and set_kernel_variant take
 * the snapshot lock exclusively, so an evaluation sees the map exactly as of one
 * commit.
 *
 * "_device" variants take DEVICE pointers (inputs already resident in HBM) and
 * are asynchronous on the context's own stream (nbvgpu_set_stream).
 */
#ifndef NBVGPU_H_
#define NBVGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nbvgpu_ctx nbvgpu_ctx;

typedef enum {
  NBVGPU_OK = 0,
  NBVGPU_ERR_INVALID = -1,   /* null pointer, bad enum, bad size                      */
  NBVGPU_ERR_NO_DEVICE = -2, /* no CUDA device / not sm_100: there is no CPU fallback */
  NBVGPU_ERR_CUDA = -3,      /* a CUDA runtime call failed (see nbvgpu_last_error)    */
  NBVGPU_ERR_CAPACITY = -4,  /* layer block pool or hash table full                   */
  NBVGPU_ERR_STATE = -5,     /* camera not set, wrong layer kind, ...                 */
  NBVGPU_ERR_RANGE = -6      /* non-finite or out-of-range coordinate (|x/voxel| >= 2^22) */
} nbvgpu_status;

/* Layer kinds: TsdfVoxel {distance, weight} / EsdfVoxel {distance}  (voxblox core/voxel.h:12-37). */
enum { NBVGPU_LAYER_TSDF = 0, NBVGPU_LAYER_ESDF = 1 };

/* Payload packings accepted by nbvgpu_layer_update*:
 *  PLANAR : TSDF float[n][vps^3][2] = (distance, weight); ESDF float[n][vps^3] = distance.
 *  VOXBLOX: the reference's own block serialisation (voxblox/src/core/block.cc:159-183 TSDF =
 *           3 x u32 per voxel: distance bits, weight bits, rgba; :203-234 ESDF = 2 x u32 per voxel:
 *           distance bits, parent|flags), i.e. the body of voxblox_msgs/Block.data. */
enum { NBVGPU_PACK_PLANAR = 0, NBVGPU_PACK_VOXBLOX = 1 };

/* Camera + gain parameters: the inputs of CameraModel::setIntrinsicsFromFoV / setExtrinsics /
 * setBoundingBox (nbveplanner/src/camera_model.cpp:25-96, nbvep.cpp:40-61) and Params::gain_*
 * (nbveplanner/include/nbveplanner/params.h:57-62). */
typedef struct {
  double hfov_deg, vfov_deg; /* system/camera/hfov, vfov                         */
  double near_m, far_m;      /* sensor_min_range, nbvep/gain/range               */
  double q_CB[4];            /* extrinsic T_C_B rotation, (w, x, y, z)           */
  double t_CB[3];            /* extrinsic T_C_B translation                      */
  double bbx_min[3], bbx_max[3]; /* exploration box, bbx/min*, bbx/max*          */
  /* The reference adds one of these per counted voxel to a running double in ray / voxel order (rrt.cpp:436-443); this
   * library counts voxels as integers and forms n_u*g_u + n_o*g_o + n_f*g_f once per yaw.  The two are bit-identical
   * when every partial sum is exact -- the defaults (0, 0, 1) and any weights that are multiples of 2^-20 below 1024 --
   * and agree to 1e-6 relative otherwise (north_star's tolerance for the floating-point gain; the counts stay exact,
   * a best-yaw window or best node may differ on ties closer than that: tests/test_gpu_parity.py::
   * test_non_default_gain_weights_within_tolerance). */
  double gain_free, gain_occupied, gain_unmapped;
  int32_t sum_all_yaws;      /* 0: optimized_gain (best yaw window); 1: gain() (sum of all yaws, yaw = 0) */
  int32_t reserved;
} nbvgpu_camera;

/* Counters since context creation (metrics row). */
typedef struct {
  uint64_t rays;             /* rays cast by gain kernels                         */
  uint64_t voxel_steps;      /* executed voxel-steps of gain kernels              */
  uint64_t collision_steps;  /* executed voxel-steps of collision kernels         */
  uint64_t blocks_uploaded;  /* blocks written by layer updates                   */
  uint64_t bytes_uploaded;   /* payload + key bytes copied host->device           */
  uint64_t kernel_launches;  /* kernels launched by this library                  */
} nbvgpu_stats;

/* Best-node record of nbvgpu_gain_eval_best* (rrt.cpp:220-246). */
typedef struct {
  int64_t index;             /* candidate index, -1 if no score > zero_gain       */
  double score;              /* Node::gain_ = num_unmapped / time_to_reach        */
  double gain;               /* Node::num_unmapped_ (return value of optimized_gain) */
  double yaw;                /* state[3], radians                                 */
} nbvgpu_best;

/* ---- context ---------------------------------------------------------------------------- */
int nbvgpu_create(int device, nbvgpu_ctx** out);
void nbvgpu_destroy(nbvgpu_ctx* ctx);

/* ---- multi-GPU contexts (nbvgpu_config{n_gpus, device_ids}) ------------------------------------
 * The path shards by candidates: every GPU keeps a replica of the map mirror and evaluates a contiguous block of
 * ceil(n / world) candidates of each batch; nbvgpu_commit replicates the changed blocks with ONE ncclBroadcast per chunk
 * of {descriptor table | payload} over NVLink; the per-shard best records come back through page-locked host memory that
 * every GPU writes and every rank reads (no collective, no copy, no stream synchronisation per planning step).  NCCL is
 * looked up at run time in the libnccl.so.2 the process already has (PyTorch's), else $NBVGPU_NCCL_LIB, else the system's.
 *
 * Two ways to build one -- the handle then goes through the SAME entry points as a single-GPU context:
 *  nbvgpu_create_multi  one process drives n_gpus devices (what the reference's single C++ planner process needs:
 *                       rrt.cpp:210-219, history.cpp:249,461,555 keep calling one handle).  device_ids may be NULL (0..n-1).
 *  nbvgpu_create_rank   one process per GPU (torchrun / MPI style launches): rank 0 obtains an id with
 *                       nbvgpu_group_unique_id (128 bytes) and hands it to the other ranks by whatever means the launcher has.
 * Semantics on such a handle:
 *  - layer_create, layer_clear, set_camera, set_kernel_variant, commit and the batch evaluators listed next are
 *    COLLECTIVE in the one-process-per-GPU form: every rank makes the same calls in the same order;
 *  - layer_update / layer_update_pinned / layer_remove stage on rank 0 (NBVGPU_ERR_STATE on other ranks);
 *  - nbvgpu_evaluate_candidates and nbvgpu_gain_eval_best take the WHOLE batch on every rank, evaluate this process's
 *    shard(s) and return the best node of the whole batch; per-candidate outputs are filled for the shards evaluated in
 *    this process (all of them with nbvgpu_create_multi);
 *  - nbvgpu_gain_eval / nbvgpu_check_segments shard large batches over the GPUs of this process;
 *  - everything else (n = 1 shims, frontier, interpolation, tree expansion, read-back, _device variants) runs on this
 *    process's first GPU: the map is replicated, so any replica answers. */
int nbvgpu_create_multi(int n_gpus, const int* device_ids, nbvgpu_ctx** out);
int nbvgpu_group_unique_id(uint8_t* id_out, size_t len /* >= 128 */);
int nbvgpu_create_rank(int device, int rank, int world, const uint8_t* unique_id, size_t len, nbvgpu_ctx** out);
/* rank of this process's first GPU, number of GPUs in the context, GPUs driven by this process (any may be NULL). */
int nbvgpu_group_info(nbvgpu_ctx* ctx, int* rank, int* world, int* local_gpus);
/* Use the caller's cudaStream_t (passed as void*) for all subsequent work; NULL restores the
 * library-owned stream (pass cudaStreamLegacy / cudaStreamPerThread to name the default streams). */
int nbvgpu_set_stream(nbvgpu_ctx* ctx, void* cuda_stream);
int nbvgpu_synchronize(nbvgpu_ctx* ctx);
const char* nbvgpu_last_error(nbvgpu_ctx* ctx);
const char* nbvgpu_version(void);

/* ---- map mirror: voxblox::Layer<TsdfVoxel|EsdfVoxel> (core/layer.h) ---------------------- */
/* Layer(voxel_size, voxels_per_side) (layer.h:34-44).  max_blocks sizes the tile pool and the
 * open-addressing table (capacity = next pow2 >= 2 * max_blocks). */
int nbvgpu_layer_create(nbvgpu_ctx* ctx, int kind, float voxel_size, int voxels_per_side,
                        size_t max_blocks, int* layer_out);
/* Insert-or-overwrite n blocks (the blocks Layer::getAllUpdatedBlocks(Update::kMap) returns,
 * layer.h:194-203).  Host buffers are copied before the call returns; the update becomes visible
 * to evaluations at the next nbvgpu_commit. keys = BlockIndex[n] (int32 x 3). */
int nbvgpu_layer_update(nbvgpu_ctx* ctx, int layer, size_t n_blocks, const int32_t* keys,
                        const void* payload, int packing);
/* Same without the staging copy: `payload` must be page-locked host memory (cudaHostAlloc / cudaHostRegister) and must
 * stay untouched until nbvgpu_commit has returned; the commit copies it to the GPU straight from there (the upload hook
 * can serialise Layer::getAllUpdatedBlocks into such a buffer). */
int nbvgpu_layer_update_pinned(nbvgpu_ctx* ctx, int layer, size_t n_blocks, const int32_t* keys,
                               const void* payload, int packing);
/* Page-locked host memory that EVERY GPU of the context can read by DMA: portable cudaHostAlloc memory when one process drives
 * the context (nbvgpu_create, nbvgpu_create_multi); in a one-process-per-GPU context (nbvgpu_create_rank) a POSIX shared-memory
 * segment that every rank maps and page-locks -- there the call is COLLECTIVE: every rank makes it with the same size, in the same
 * order, and gets its own mapping of the same bytes (rank 0, the map's owner, fills it).  Blocks staged out of such memory with
 * nbvgpu_layer_update_pinned are replicated by a parallel pull when a commit chunk carries at least NBVGPU_PULL_MIN_MB (default 8)
 * MiB of them: every GPU copies 1/world of the chunk over its own PCIe link and one ncclAllGather over NVLink completes the chunk
 * on every GPU, instead of rank 0's link carrying everything in front of a broadcast (7 GB map of BASELINE configs[3] on 8 GPUs:
 * see profiles/README.md).  This is where the reference's upload hook should serialise Layer::getAllUpdatedBlocks
 * (voxblox/core/layer.h:194-203, block.cc:159-234) when the planner owns several GPUs.  nbvgpu_shared_free releases one allocation
 * (not while blocks staged from it await their commit; every rank that made the allocation frees it); nbvgpu_destroy releases the
 * rest.  A context hands out at most 256 allocations over its lifetime. */
int nbvgpu_shared_alloc(nbvgpu_ctx* ctx, size_t bytes, void** ptr);
int nbvgpu_shared_free(nbvgpu_ctx* ctx, void* ptr);
/* Same with device buffers (single-GPU contexts); applied in stream order.  A key listed twice is written by either copy. */
int nbvgpu_layer_update_device(nbvgpu_ctx* ctx, int layer, size_t n_blocks, const int32_t* d_keys,
                               const void* d_payload, int packing);
/* Layer::removeBlock (layer.h:171-173) / removeAllBlocks (:175). */
int nbvgpu_layer_remove(nbvgpu_ctx* ctx, int layer, size_t n_blocks, const int32_t* keys);
int nbvgpu_layer_clear(nbvgpu_ctx* ctx, int layer);
/* Publish staged updates/removals of ALL layers to the device snapshot: removals in one small launch, then per chunk of at
 * most 64 MiB of payload ($NBVGPU_CHUNK_MB) one device buffer {32-byte block descriptors | payload}, filled by host->device
 * copies (of chunk k+1 while chunk k is applied), broadcast over NCCL in a multi-GPU context, and applied by ONE kernel
 * (a CTA per block: hash insert-or-find, then the tile written with 128-bit stores).  Synchronous: returns after the
 * snapshot is in place, NBVGPU_ERR_CAPACITY if a tile pool or table overflowed. */
int nbvgpu_commit(nbvgpu_ctx* ctx);
/* Layer::getNumberOfAllocatedBlocks (layer.h:205). */
int nbvgpu_layer_num_blocks(nbvgpu_ctx* ctx, int layer, size_t* out);
/* Layer::getVoxelPtrByGlobalIndex (layer.h:215-239) for n global voxel indices (int64 x 3):
 * out_values[n][2] = (distance, weight) (ESDF: weight slot = 0); out_found[n] = 1 if the block is
 * allocated; out_status[n] (may be NULL) = 0 unknown / 1 free / 2 occupied per
 * getVoxelStatusByIndex (TSDF only). */
int nbvgpu_layer_read_voxels(nbvgpu_ctx* ctx, int layer, size_t n, const int64_t* global_index,
                             float* out_values, uint8_t* out_found, uint8_t* out_status);

/* ---- ingest in the reference's own formats ------------------------------------ */
/* voxblox::io::LoadBlocksFromFile (voxblox/io/layer_io_inl.h:14-127, kReplace, multiple-layer files
 * supported): reads a .voxblox file (varint-framed LayerProto + BlockProto messages), keeps the blocks of
 * the first layer compatible with `layer` (voxel size, voxels per side, "tsdf"/"esdf"), uploads them
 * and commits.  *n_blocks_out (may be NULL) = blocks loaded. */
int nbvgpu_layer_load_voxblox_file(nbvgpu_ctx* ctx, int layer, const char* path, size_t* n_blocks_out);
/* voxblox::deserializeMsgToLayer (voxblox_ros/conversions_inl.h:43-115) on a voxblox_msgs/Layer message
 * in ROS 1 wire serialisation: ACTION_UPDATE (0) overwrites/inserts the blocks, ACTION_RESET (2) clears
 * the layer first; ACTION_MERGE (1) needs the integrators' voxel merge and returns NBVGPU_ERR_INVALID. */
int nbvgpu_layer_apply_layer_msg(nbvgpu_ctx* ctx, int layer, const uint8_t* msg, size_t len, size_t* n_blocks_out);
/* Host-only parsers behind the two calls above (no device needed): kind = NBVGPU_LAYER_*; on success
 * *n_blocks is the number of blocks found; keys_out[n][3] / words_out[n][words per block] are filled when
 * non-NULL and capacity_blocks >= *n_blocks.  format: 0 = .voxblox bytes, 1 = Layer message bytes;
 * *action_out (may be NULL) = the message's action. */
int nbvgpu_parse_blocks(int format, const uint8_t* buf, size_t len, int kind, float voxel_size, int voxels_per_side,
                        size_t* n_blocks, int32_t* keys_out, uint32_t* words_out, size_t capacity_blocks,
                        int* action_out);

/* ---- gain evaluator --------------------------------------------------------------------- */
int nbvgpu_set_camera(nbvgpu_ctx* ctx, const nbvgpu_camera* cam);
/* Batched optimized_gain()/gain(): pos[n][3] = state[0..2].  Outputs (any may be NULL except gain):
 * gain[n] = return value; yaw_deg[n] = state[3] in integer degrees (-180..175; 0 for gain());
 * counts[n][3] = unknown/free/occupied voxel counts over the winning yaw window (all yaws for
 * gain()); per_yaw_counts[n][360][3] = the per-yaw counts behind vertical_gain[] (rrt.cpp:447);
 * steps[n] = executed voxel-steps. */
int nbvgpu_gain_eval(nbvgpu_ctx* ctx, int tsdf_layer, size_t n, const double* pos, double* gain,
                     int32_t* yaw_deg, uint64_t* counts, uint32_t* per_yaw_counts, uint64_t* steps);
int nbvgpu_gain_eval_device(nbvgpu_ctx* ctx, int tsdf_layer, size_t n, const double* d_pos,
                            double* d_gain, int32_t* d_yaw_deg, uint64_t* d_counts,
                            uint32_t* d_per_yaw_counts, uint64_t* d_steps);
/* n = 1 shim with the reference's signature semantics: state[3] is overwritten with the yaw in
 * radians (max_gain_yaw * M_PI / 180), the gain is stored in *gain_out. */
int nbvgpu_optimized_gain(nbvgpu_ctx* ctx, int tsdf_layer, double state[4], double* gain_out);

/* Batched evaluation + node scoring (rrt.cpp:220-246): score = gain / max(dyaw / dyaw_max,
 * distance / v_max) with dyaw = yaw - parent_yaw wrapped once into (-pi, pi] (not abs'd, as in the
 * reference); best = first candidate with score > running best starting at zero_gain.
 * edge_free (may be NULL) masks candidates whose edge failed checkMotion: they are neither swept nor scored (gain = yaw = 0).
 * gain / yaw_deg (may be NULL) receive the per-candidate results as in nbvgpu_gain_eval. */
int nbvgpu_gain_eval_best(nbvgpu_ctx* ctx, int tsdf_layer, size_t n, const double* pos,
                          const double* parent_yaw, const double* distance, const uint8_t* edge_free,
                          double v_max, double dyaw_max, double zero_gain, nbvgpu_best* best,
                          double* gain, int32_t* yaw_deg);
int nbvgpu_gain_eval_best_device(nbvgpu_ctx* ctx, int tsdf_layer, size_t n, const double* d_pos,
                                 const double* d_parent_yaw, const double* d_distance,
                                 const uint8_t* d_edge_free, double v_max, double dyaw_max,
                                 double zero_gain, nbvgpu_best* d_best, double* d_gain,
                                 int32_t* d_yaw_deg);

/* Edge check + gain + node scoring of n candidates in one call: what RrtTree::iterate does with a sample after the
 * sampler (rrt.cpp:205-246) -- checkMotion on seg[i] = {start xyz, end xyz} (voxblox_manager.h:89-108), optimized_gain at
 * pos[i], the score / best-node update -- i.e. nbvgpu_check_segments followed by nbvgpu_gain_eval_best with its verdicts,
 * with identical results.  As in the reference (rrt.cpp:210-216) a candidate whose edge collides is not swept: its gain and
 * yaw read 0.  One call instead of two saves, for host buffers, a copy / synchronise round trip.  edge_free receives the
 * per-candidate verdicts (required for the device variant, optional for the host variant); gain / yaw_deg may be NULL. */
int nbvgpu_evaluate_candidates(nbvgpu_ctx* ctx, int tsdf_layer, int esdf_layer, double robot_radius, size_t n,
                               const double* pos, const double* seg, const double* parent_yaw, const double* distance,
                               double v_max, double dyaw_max, double zero_gain, uint8_t* edge_free, nbvgpu_best* best,
                               double* gain, int32_t* yaw_deg);
int nbvgpu_evaluate_candidates_device(nbvgpu_ctx* ctx, int tsdf_layer, int esdf_layer, double robot_radius, size_t n,
                                      const double* d_pos, const double* d_seg, const double* d_parent_yaw,
                                      const double* d_distance, double v_max, double dyaw_max, double zero_gain,
                                      uint8_t* d_edge_free, nbvgpu_best* d_best, double* d_gain, int32_t* d_yaw_deg);
/* One-process-per-GPU contexts, inputs already in this GPU's memory: this rank's shard -- n_local candidates whose first
 * has global index index_offset -- is evaluated as above and its best record posted to the shared block; the call then
 * waits for the records of every rank and returns the best node of the whole batch in HOST memory (*best).  Collective:
 * every rank calls it once per batch, a rank without candidates with n_local = 0.  d_gain / d_yaw_deg may be NULL. */
int nbvgpu_evaluate_shard_device(nbvgpu_ctx* ctx, int tsdf_layer, int esdf_layer, double robot_radius, size_t n_local,
                                 int64_t index_offset, const double* d_pos, const double* d_seg, const double* d_parent_yaw,
                                 const double* d_distance, double v_max, double dyaw_max, double zero_gain,
                                 uint8_t* d_edge_free, double* d_gain, int32_t* d_yaw_deg, nbvgpu_best* best);

/* ---- batched tree expansion ------------------------------------------------- */
/* Parameters of RrtTree::iterate that are not in nbvgpu_camera (params.h:26-35,93,103; rrt.cpp:175). */
typedef struct {
  double root[3];          /* rootNode_->state_ xyz: centre of the sampling sphere                 */
  double vicinity;         /* root_vicinity_                                                        */
  double extension_range;  /* nbvep/tree/extension_range                                            */
  double overshoot;        /* system/overshoot                                                      */
  double robot_radius;     /* system/robot_radius                                                   */
  double v_max, dyaw_max;  /* system/v_max, system/dyaw_max                                         */
  double zero_gain;        /* nbvep/gain/zero                                                       */
  int32_t sequential;      /* 1: the reference's semantics -- sample i sees the nodes created by samples 0..i-1 of
                              the batch (kd_insert3 at rrt.cpp:236) as parents; 0: parents are the given nodes only */
  int32_t reserved;
} nbvgpu_expand_params;
/* One batched pass of the body of RrtTree::iterate (rrt.cpp:172-246) over n uniform triples u[n][3] in
 * [0,1) (the caller's RNG replaces the reference's clock-seeded mt19937): sample in the vicinity sphere,
 * reject outside it or outside the exploration box; nearest of the n_nodes existing tree nodes
 * (node_state[n_nodes][4] = x y z yaw, node_distance[n_nodes]) as parent; clip to extension_range; edge +
 * overshoot through checkMotion; gain (optimized_gain / gain per the camera) for free edges; node score
 * against the parent's yaw and distance; best = first strict maximum over zero_gain in sample order.
 * params->sequential = 1 is a faithful replay of n consecutive iterate() calls on the caller's stream of triples: the
 * order-dependent front half (sample, nearest node among the given AND the already created nodes, clip, checkMotion,
 * insert) runs sample by sample inside one CTA, the gains are batched, and each score uses its parent's chosen yaw even
 * when the parent was created by the same call (rrt.cpp:225); parent[i] indexes the given nodes first, then the created
 * nodes in creation order.  With sequential = 0 parents are the given nodes only (samples of one batch do not see each
 * other; everything is batched; with n = 1 both modes are exactly one iterate() call).  Per-sample outputs (any may be
 * NULL): status 0 = sample rejected, 1 = edge collides, 2 = node created; pos[n][3]; parent[n]; gain[n]; yaw[n] (radians);
 * distance[n]; score[n]. */
int nbvgpu_expand_tree(nbvgpu_ctx* ctx, int tsdf_layer, int esdf_layer, const nbvgpu_expand_params* params,
                       size_t n_nodes, const double* node_state, const double* node_distance, size_t n,
                       const double* uniforms, uint8_t* status, double* pos, int32_t* parent, double* gain, double* yaw,
                       double* distance, double* score, nbvgpu_best* best);

/* ---- segment collision check ------------------------------------------------------------ */
/* Batched HighResManager::checkMotion: seg[n][6] = (start xyz, end xyz); is_free[n] = 1 when every
 * traversed voxel is allocated and robot_radius < distance. */
int nbvgpu_check_segments(nbvgpu_ctx* ctx, int esdf_layer, double robot_radius, size_t n,
                          const double* seg, uint8_t* is_free);
int nbvgpu_check_segments_device(nbvgpu_ctx* ctx, int esdf_layer, double robot_radius, size_t n,
                                 const double* d_seg, uint8_t* d_is_free);
/* n = 1 shim: *free_out = checkMotion(start, end). */
int nbvgpu_check_motion(nbvgpu_ctx* ctx, int esdf_layer, double robot_radius, const double start[3],
                        const double end[3], int* free_out);

/* ---- ESDF interpolated distance and gradient -----------------------------
 * HighResManager::getDistanceAndGradientAtPosition(pos, &distance, &grad) (nbveplanner/src/voxblox_manager.cpp:143-148,
 * called by History::refineVertexPosition, history.cpp:235) = EsdfMap::getDistanceAndGradientAtPosition
 * (voxblox/src/core/esdf_map.cc:22-52): trilinear interpolation over the eight observed voxels whose centres
 * surround the position and a central-difference gradient from six more interpolations at +-voxel_size
 * (voxblox/interpolator/interpolator_inl.h:47-77,160-330,447-474), in float.  ok[i] = the reference's return value;
 * distance[i] / gradient[i][3] are written in every case, exactly as the reference leaves them (0 / the partial
 * sums of a failing gradient).  The ESDF layer keeps EsdfVoxel::observed per voxel: bit 0 of the flags byte in
 * NBVGPU_PACK_VOXBLOX payloads (block.cc:203-234); a NBVGPU_PACK_PLANAR voxel is observed unless its distance is NaN.
 * Values: the reference sums its two 8-term float products inside Eigen; this library sums them left to right, so
 * against a real Eigen build values agree to float rounding (a few ulp of the largest voxel distance), the ok flag
 * and the choice of voxels exactly.  Host buffers. */
int nbvgpu_esdf_distance_gradient(nbvgpu_ctx* ctx, int esdf_layer, size_t n, const double* positions /* [n][3] */,
                                  uint8_t* ok /* [n] */, double* distance /* [n] */, double* gradient /* [n][3] */);
/* Same with device-resident buffers: stream-ordered on the context's stream, no copies, no synchronisation.  Positions are
 * not validated by the host here: a non-finite or out-of-range position gives ok = 0 and zero outputs. */
int nbvgpu_esdf_distance_gradient_device(nbvgpu_ctx* ctx, int esdf_layer, size_t n, const double* d_positions,
                                         uint8_t* d_ok, double* d_distance, double* d_gradient);

/* ---- history-graph frontier potential ----------------------------------
 * History::bfs(point) (nbveplanner/src/history.cpp:96-137) as History::recalculatePotential calls it per
 * vertex (:211-222): 6-connected flood fill from `point` over free cells -- cells being accumulated double
 * coordinates identified by their std::to_string, FIFO order --, bounded by `max_distance` (the reference
 * hard-codes 6.0) and by the exploration bounds (Params::bbx_min_/bbx_max_: the camera's bounding box here,
 * nbvep.cpp:45), status from LowResManager::getVoxelStatus(position) (voxblox_manager.cpp:73-90) on the
 * TSDF layer.  potential_gain[i] = frontierVoxels * pow(voxel_size, 3.0), frontierVoxels = number of
 * (visited free cell, unknown neighbour) encounters; visited_cells[i] = size of the reference's `visited` set
 * at the end (the start included) and cells_checksum[i] = an order-independent sum over those cells of a
 * 64-bit mix of the bit patterns of their three doubles: diagnostics the reference does not return, used by
 * the parity tests to show that every cell carries the doubles of the reference's first arrival.  Host
 * buffers. */
int nbvgpu_frontier_potential(nbvgpu_ctx* ctx, int tsdf_layer, size_t n, const double* points /* [n][3] */,
                              double max_distance, double* potential_gain /* [n] */,
                              uint64_t* frontier_voxels /* [n] or NULL */, uint64_t* visited_cells /* [n] or NULL */,
                              uint64_t* cells_checksum /* [n] or NULL */);
/* Vertices, since the context was created, whose flood fill met one lattice cell under two different strings (next
 * to a "%f" rounding boundary, or -0.0 against +0.0) and was therefore redone with the exact hash-keyed visited set
 * instead of the dense one; for tests (NBVGPU_FRONTIER_HASH=1 in the environment forces the hash set for all). */
int nbvgpu_debug_frontier_hash_redo(nbvgpu_ctx* ctx, uint64_t* count);
/* Measurement infrastructure (SURVEY.md 8 d): read bandwidth of this GPU's L2 in GB/s -- 128-bit ld.global.cg
 * loads from every SM over a `bytes`-sized buffer (choose it well below the L2 capacity), `iters` timed passes
 * after an untimed warm-up pass, CUDA events on the context's stream.  Used by bench.py for the L2-normalised
 * roofline fraction; not part of the planner interface. */
int nbvgpu_probe_l2_read_bandwidth(nbvgpu_ctx* ctx, size_t bytes, int iters, double* gb_per_s);
/* Ray-sweep launches, since the context was created, that used the z-clipped dense shared-memory volume (fine-voxel layers
 * whose unclipped volume does not fit; NBVGPU_NO_ZCLIP=1 in the environment disables the clipping); for tests. */
int nbvgpu_debug_dense_clipped_launches(nbvgpu_ctx* ctx, uint64_t* count);
/* Ray-sweep launches, since the context was created, that ran the best-yaw window scan in the last CTA of each candidate
 * (batches of at most 64 candidates on the dense-volume sweep: a single-candidate nbvgpu_optimized_gain is then one kernel
 * launch reading and writing mapped pinned host memory; NBVGPU_NO_FUSE=1 / NBVGPU_NO_SMALL_PATH=1 in the environment
 * restore the two-launch / staged-copy paths); for tests. */
int nbvgpu_debug_fused_launches(nbvgpu_ctx* ctx, uint64_t* count);
/* Ray-sweep launches, since the context was created, that ran as 512-thread CTAs (two per SM) because the dense
 * shared-memory volume of a yaw chunk needs more than a quarter but at most half of an SM's shared memory (fine voxels
 * under a tall exploration box; NBVGPU_NO_WIDE_CTA=1 in the environment keeps the slot-grid sweep instead); for tests. */
int nbvgpu_debug_wide_cta_launches(nbvgpu_ctx* ctx, uint64_t* count);
/* Planning steps, since the context was created, that ran as ONE cudaGraphLaunch: nbvgpu_evaluate_candidates[_device],
 * nbvgpu_evaluate_shard_device and nbvgpu_gain_eval_best run eagerly the first time a shape (sizes, buffers, layers, camera,
 * parameters) is seen, are captured into a CUDA graph the second time and replayed from then on (copies of the host-buffer
 * forms included); shapes that synchronise inside (the z-clipped volume) stay eager.  NBVGPU_NO_GRAPH=1 disables it. */
int nbvgpu_debug_graph_replays(nbvgpu_ctx* ctx, uint64_t* count);
/* Commit chunks of a multi-GPU context that were replicated by slice pull + all-gather (nbvgpu_shared_alloc) rather than
 * by a broadcast from rank 0. */
int nbvgpu_debug_pulled_chunks(nbvgpu_ctx* ctx, uint64_t* count);
/* Host-only helper for tests: the largest double x with sqrt(x) <= d (round to nearest).  The flood fill tests
 * `squared distance <= threshold` instead of `(candidate - point).norm() <= max_distance` (history.cpp:118): the two are
 * equivalent because the square root is monotone and correctly rounded.  NBVGPU_ERR_INVALID for negative or non-finite d. */
int nbvgpu_debug_sqrt_threshold(double d, double* threshold);
/* Host-only helper for tests: round-half-even of |x| * 10^6, the integer behind "%f" / std::to_string(double),
 * exactly as the flood fill's cell keys compute it.  NBVGPU_ERR_RANGE for non-finite x or |x| >= 2^21. */
int nbvgpu_debug_round_decimal6(double x, uint64_t* magnitude);

/* ---- introspection ---------------------------------------------------------------------- */
int nbvgpu_get_stats(nbvgpu_ctx* ctx, nbvgpu_stats* out);
/* CUDA-event timing of the ray-sweep kernel (the dominant kernel): when enabled every launch is
 * bracketed by events on the context stream; nbvgpu_profile_collect synchronises, returns the summed
 * device time in ms and the number of launches since the last collect, and resets the list. */
int nbvgpu_set_profiling(nbvgpu_ctx* ctx, int enable);
int nbvgpu_profile_collect(nbvgpu_ctx* ctx, double* gain_kernel_ms, uint64_t* launches);
/* Selects the ray-sweep kernel variant; all produce identical counts (A/B measurement and tests):
 *  0 default: dense shared-memory status volume when it fits (16^3 blocks), else slot-grid path
 *  1 plain transcription: FP64 plane tests on every step + {distance, weight} tile loads
 *  2 variant 0 with every shortcut cross-checked against the exact test (mismatch counter)
 *  3 slot-grid path reading {distance, weight} tiles instead of the 2-bit status tiles
 *  4 slot-grid path (never the dense volume);  5 = 4 with cross-checks */
int nbvgpu_set_kernel_variant(nbvgpu_ctx* ctx, int variant);
/* Number of single-precision frustum-classifier decisions contradicted by the exact FP64 plane test,
 * counted by kernel variants 2 and 5.  Must be 0; the tests assert it. */
int nbvgpu_debug_classifier_mismatches(nbvgpu_ctx* ctx, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* NBVGPU_H_ */
