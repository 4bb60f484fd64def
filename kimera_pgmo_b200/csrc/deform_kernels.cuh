// Device-side data layout and launchers of the deformation path (sm_100a).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace pgmo {

// status byte per processing position
constexpr uint8_t kStatusSkipped = 0;    // reference `continue`d: no control_point_map entry
constexpr uint8_t kStatusDone = 1;       // deformed
constexpr uint8_t kStatusNeedsTie = 2;   // exact float-distance tie (or k too large for the
                                         // register path): finished by the resolver kernel

constexpr int kResolverMaxDepth = 26;    // octree depth the resolver's DFS stack can hold

struct RunFrameDev {
  double min_xyz[3];
  uint32_t depth;
  int32_t shift[3];
};

struct DeformArgs {
  uint64_t first_position;
  uint64_t n_local;
  const float* x;
  const float* y;
  const float* z;
  const uint64_t* subset;     // local vertex index per local position, or null
  // schedule
  const uint64_t* run_start;  // n_runs + 1 entries, run_start[n_runs] = total positions
  uint32_t n_runs;
  const uint4* run_info;      // {lo, hi, k, skip}
  const RunFrameDev* run_frame;
  // control table
  const float4* ctrl_f4;      // float(x), float(y), float(z), del_run (int bits)
  const int4* ctrl_key;       // stable voxel key xyz
  const double* ctrl_pos;     // n_c x 3 (f64 control point)
  const double* poses;        // n_c x 12 (row-major R, t)
  const float* ctrl_pose_f32; // n_c x 16: R(9) t(3) g(3) pad — FP32 mode
  double resolution;
  // outputs
  float* ox;
  float* oy;
  float* oz;
  double* of64;               // 3 per position, or null
  uint32_t* assoc;            // assoc_stride per position, or null
  uint32_t assoc_stride;
  uint8_t* assoc_cnt;         // or null
  uint8_t* status;            // per position (always present)
  unsigned long long* counters;  // [0] = vertices sent to the resolver, [1] = already resolved,
                                 // [2] = work items found stale by the validating blend pass,
                                 // [3] = stale items already redone
  uint32_t* tie_list;            // local positions sent to the resolver (capacity n_local)
  uint32_t* tie_slot;            // their slot in the association cache (kNoCacheSlot: none)
};

// ------------------------------------------------------------------------------------------
// Association cache (deform_blend.cu).  interpPoint's neighbour set and weights
// (kimera_pgmo/src/mesh_deformation.cpp:83-94) depend on the vertex, the control positions /
// stamps / validity and the window schedule - not on the poses.  The search kernel (K3a) leaves
// them in HBM, one record per local position in the order the search grouped the vertices, and
// the blend kernel (K3b) turns records + pose rows into deformed points.  After a new solve
// (poses only) K3b alone runs: it re-reads the vertices, checks a 64-bit fingerprint of the
// item's vertex bits against the one the records were computed for and redoes the search only
// for items that fail it.
constexpr uint32_t kNoCacheSlot = 0xFFFFFFFFu;
constexpr int kCacheMaxNbr = 14;  // blended neighbours a record can hold (k <= 13)
struct AssocCache {
  // plane 0: {vertex offset inside the item, nbr 0..6}, plane 1: {nbr 7..14}; nbr = (control
  // index - run.lo) | row copy << 11 | bank class << 13; slot = local position of the record
  // (item start + rank in search order)
  uint16_t* rec;       // planes x n_slots x 8
  void* weight;        // w_planes x n_slots, double (FP64 mode) or float (FP32 mode)
  uint32_t w_planes;   // neighbours a record holds weights for (>= max_k + 1)
  uint64_t n_slots;    // = n_local of the slab the cache belongs to
  unsigned long long* item_fp;  // per work item: fingerprint of the vertex bits
  uint32_t* item_ok;   // per work item: 1 = records valid for item_fp
  uint32_t* stale_items;  // work items the validating pass rejected (capacity: all items)
  const void* pose_rows;  // FP64 mode: kBlendCopies planes of plane_rows rows of 16 doubles (128 B,
                          // 16-byte chunk c of copy m stored at c ^ blendClass(class, row, m));
                          // FP32 mode: n_c rows of 16 floats.  Row = R(9) t(3) g(3) pad
  uint64_t plane_rows;    // rows per copy plane (= control points the table was built for)
  uint32_t n_copies;      // copies the blend kernel stages (what fits its shared memory)
  // Bank class (0..7) of every control point's row: a colouring in which points that lie close
  // together differ (deform_api.cu: colourControls).  The lanes of a quarter warp read rows of
  // control points that lie close together; different classes put the same chunk of their rows
  // into different bank groups.
  const uint8_t* ctrl_class;
};
// A record's neighbour field: [10:0] control index - run.lo, [12:11] which copy of the pose row
// the blend kernel reads, [15:13] the bank class of that copy (the XOR its chunks are stored under).
// Every row is stored kBlendCopies times under different classes; a placement pass over fresh
// records (placeRecordsKernel) picks, per quarter warp and neighbour slot, copies whose bank groups
// differ, which is what an LDS.128 of 8 lanes needs to be one wavefront.
constexpr uint32_t kCacheRelBits = 11;
constexpr uint32_t kCacheRelMask = (1u << kCacheRelBits) - 1u;
constexpr uint32_t kCacheCopyShift = 11;
constexpr uint32_t kCacheCopyMask = 3u;
constexpr uint32_t kCacheClassShift = 13;
constexpr int kBlendCopies = 3;

// per-run uniform grid over the run's alive control points (deform_grid.cu)
constexpr int kGridMaxCells = 1024;
struct RunGridDev {
  float ox, oy, oz;       // grid origin = bounding-box minimum of the alive set
  float cell, inv_cell;   // cell edge
  int nx, ny, nz;         // nx * ny * nz <= kGridMaxCells, x fastest
  uint32_t point_offset;  // first slot of this run in GridArgs::points
  uint32_t n_alive;
  float r0;               // first search radius: ~1.25 x expected (K+1)-th neighbour distance
};
// up to kDeformChunk consecutive local positions of one run (one CTA of the grid kernels)
#ifndef PGMO_CHUNK
#define PGMO_CHUNK 2048
#endif
constexpr uint32_t kDeformChunk = PGMO_CHUNK;
// runs with at most this many alive control points go through the bucket-ordered kernel
// (deform_sorted.cu); larger alive sets stay on the grid-walk kernel (deform_grid.cu)
constexpr int kSortedMaxPoints = 512;
struct WorkItem {
  uint64_t start;  // local position
  uint32_t run;
  uint32_t count;
};
struct GridArgs {
  const RunGridDev* grids;  // n_runs
  const float4* points;     // alive points sorted by cell: x, y, z, control index (int bits)
  const uint32_t* cells;    // n_runs x (kGridMaxCells + 1) cell starts
  // search-box schedule, in units of the run's r0 / cell: first radius, then growth per round
  float r_scale, r_grow_mul, r_grow_add;
  // bucket-ordered kernel: points farther than sqrt(far_mul) first radii from a slab of vertices
  // are looked at only by the groups whose search radius gets that far
  float far_mul;
};

struct CandEntry {
  float dist;
  int idx;
};

struct ResolveArgs {
  CandEntry* scratch;        // one slab of cand_cap entries per resident warp
  uint32_t cand_cap;
  uint32_t n_warps;          // resident warps (grid is persistent)
  uint32_t max_leaf_points;  // most control points ever sharing one voxel (Schedule)
};

// K1: boundaries of equal-stamp stretches.  Appends (position, stamp) of every position whose
// stamp differs from its predecessor's (position 0 always) to `out`, unordered; *count is the
// number appended (may exceed capacity: entries beyond it are dropped).
cudaError_t launchStampStretches(const uint64_t* stamps, const uint64_t* subset, uint64_t n,
                                 ulonglong2* out, uint64_t capacity,
                                 unsigned long long* count, cudaStream_t stream);

// K3: windowed exact top-(k+2) + blend.  precision: 0 = FP64 blend, 1 = FP32 blend.
cudaError_t launchDeformKnnBlend(const DeformArgs& args, int precision, uint32_t max_k,
                                 cudaStream_t stream);

// K2b: one CTA per run buckets the run's alive control points into its grid.  point_offset[r] =
// first slot of run r (prefix sum of the window sizes hi - lo).
cudaError_t launchBuildRunGrids(uint32_t n_runs, const uint4* run_info, const float4* ctrl_f4,
                                const uint32_t* point_offset, RunGridDev* grids,
                                float4* grid_points, uint32_t* grid_cells, cudaStream_t stream);
// K3 over the per-run grids: same result as launchDeformKnnBlend, ~8x fewer candidates.
// Items whose run has <= skip_alive_le alive points are left untouched (they belong to the
// bucket-ordered kernel); pass 0 to process everything.
cudaError_t launchDeformKnnBlendGrid(const DeformArgs& args, const GridArgs& gargs,
                                     const WorkItem* items, uint32_t n_items, int precision,
                                     uint32_t max_k, uint32_t skip_alive_le, cudaStream_t stream);
// K3, bucket-ordered candidate lists: processes the items whose run has <= kSortedMaxPoints
// alive points and leaves the others untouched.  Same result as the other two K3 kernels.
cudaError_t launchDeformKnnBlendSorted(const DeformArgs& args, const GridArgs& gargs,
                                       const WorkItem* items, uint32_t n_items, int precision,
                                       uint32_t max_k, uint32_t min_k, uint64_t widest_window,
                                       cudaStream_t stream);

// forces the lazy loading of the default K3 kernels (handle creation)
void preloadDeformSortedKernels();

// K3a on the cache: as launchDeformKnnBlendSorted, but instead of blending the kernel writes the
// association records.  from_stale_list: a persistent grid walks cache.stale_items
// [counters[3], counters[2]) instead of `items` (the items are then indexed through it).
cudaError_t launchDeformSearchSorted(const DeformArgs& args, const GridArgs& gargs,
                                     const AssocCache& cache, const WorkItem* items,
                                     uint32_t first_item, uint32_t n_items, int precision,
                                     uint32_t max_k, uint32_t min_k, uint64_t widest_window,
                                     bool from_stale_list, int n_sms, cudaStream_t stream);
// K3b: blend from the association records.  mode 0 = items [first_item, first_item + n_items),
// records trusted; 1 = same items, records validated against the vertices (stale items are
// appended to cache.stale_items and left untouched); 2 = persistent grid over the stale list.
cudaError_t launchDeformBlendCached(const DeformArgs& args, const GridArgs& gargs,
                                    const AssocCache& cache, const WorkItem* items,
                                    uint32_t first_item, uint32_t n_items, int precision,
                                    uint32_t max_k, uint32_t min_k, uint64_t widest_window,
                                    int mode, int n_sms, cudaStream_t stream);
void preloadDeformBlendKernels();
// copies of the pose rows the blend kernel can stage for windows of up to widest_window rows
uint32_t blendCopiesFor(int precision, uint64_t widest_window);
// Placement pass over records the search kernel / the resolver just wrote (cache.n_copies > 1):
// chooses the row copy of every neighbour field.  from_stale_list as in launchDeformSearchSorted.
cudaError_t launchPlaceRecords(const DeformArgs& args, const GridArgs& gargs, const AssocCache& cache,
                               const WorkItem* items, uint32_t first_item, uint32_t n_items,
                               bool from_stale_list, int n_sms, cudaStream_t stream);
// counters[1] = counters[0]; counters[3] = counters[2]
cudaError_t launchDeformStepDone(unsigned long long* counters, cudaStream_t stream);

// K4: PCL-order resolution of the vertices flagged kStatusNeedsTie.
cudaError_t launchDeformResolve(const DeformArgs& args, const ResolveArgs& rargs,
                                const AssocCache* cache, int precision, cudaStream_t stream);
uint32_t resolverResidentWarps(int device);

// K5: rigid predeform by graph index.
// float4 {x, y, z, w} cloud <-> the SoA streams of K3 (device-resident meshes)
cudaError_t launchAosToSoa(uint64_t n, const float4* in, float* x, float* y, float* z,
                           cudaStream_t stream);
cudaError_t launchSoaToAos(uint64_t n, const float4* in, const float* x, const float* y,
                           const float* z, const uint8_t* status, float4* out,
                           cudaStream_t stream);

// KimeraPgmoMesh message arrays <-> {x, y, z, rgba} records + u64 stamps (mesh_msg.cu)
cudaError_t launchMsgToCloud(uint64_t n, const double* pts, const float* colors,
                             const uint32_t* stamps2, float4* cloud, uint64_t* stamps,
                             cudaStream_t stream);
cudaError_t launchCloudToMsg(uint64_t n, const float4* cloud, const uint64_t* stamps, double* pts,
                             float* colors, uint32_t* stamps2, cudaStream_t stream);

cudaError_t launchPredeform(uint64_t n_v, uint64_t start_idx, const float* x, const float* y,
                            const float* z, const int32_t* graph_indices, uint64_t n_c,
                            const float4* ctrl_f4_valid, const uint8_t* valid,
                            const double* ctrl_pos, const double* poses, float* ox, float* oy,
                            float* oz, uint8_t* needs_deform, int* out_of_range,
                            cudaStream_t stream);

// builds ctrl_f4 (only when positions changed) / ctrl_pose_f32 / the chunk-swizzled f64 pose
// rows of the blend kernel from the f64 tables
cudaError_t launchBuildControlTables(uint64_t n_c, const double* ctrl_pos, const double* poses,
                                     float4* ctrl_f4, float* ctrl_pose_f32, double* pose_rows64,
                                     const uint8_t* ctrl_class, bool positions_changed,
                                     int n_planes, cudaStream_t stream);  // n_planes <= kBlendCopies  // pose_rows64: kBlendCopies x n_c rows
// Rows [row_begin, row_end) of the pose table `src` (local or a peer device's memory) -> this
// device's f64 copy (poses_local, may be null or == src), f32 rows and swizzled f64 rows.
// sequenced pose source (deform_kernels.cu): flags in the source's memory order writer and readers
constexpr uint32_t kPoseSrcMaxReaders = 64;
constexpr uint32_t kPoseSrcFlagWords = 2 + kPoseSrcMaxReaders;  // ready[2], ack[kPoseSrcMaxReaders]
cudaError_t launchPoseSourceWaitAcks(const unsigned long long* acks, uint32_t n_readers,
                                     unsigned long long need, unsigned long long* timeouts,
                                     cudaStream_t stream);
cudaError_t launchPoseSourceStore(unsigned long long* flag, unsigned long long seq, cudaStream_t stream);
cudaError_t launchPoseSourceWaitReady(const unsigned long long* ready, unsigned long long seq,
                                      unsigned long long* timeouts, cudaStream_t stream);
cudaError_t launchPullBuildPoseTables(uint64_t row_begin, uint64_t row_end, uint64_t n_c,
                                      const double* src, double* poses_local, const double* ctrl_pos,
                                      float* ctrl_pose_f32, double* pose_rows64,
                                      const uint8_t* ctrl_class, int n_planes, cudaStream_t stream);
// writes del_run into ctrl_f4[].w and the stable keys (after build_schedule)
cudaError_t launchApplySchedule(uint64_t n_c, const int32_t* del_run, const int32_t* stable_key3,
                                float4* ctrl_f4, int4* ctrl_key, cudaStream_t stream);

}  // namespace pgmo
