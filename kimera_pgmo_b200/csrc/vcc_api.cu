// C ABI of VoxelClearingCompression (see include/pgmo_b200.h).
//
// Split of labour: voxel hashing, dedup, index assignment, reference counts, vertex / stamp
// storage live in HBM (vcc_kernels.cu).  The host keeps only block-level bookkeeping that the
// reference keeps in unordered_maps keyed by BlockIndex — update times and the per-block face
// maps — using the same hash and the same sequence of map operations, because the iteration
// order of those maps is what orders `polygons_` and the archive
// (voxel_clearing_compression.cpp:57-66, :90-96, :103-109).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "deform_schedule.h"
#include "vcc_kernels.cuh"
#include "mesh_interface.cuh"

using namespace pgmo;

namespace {

struct BlockKey {
  int32_t x, y, z;
  bool operator==(const BlockKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
// spatial_hash::IndexHash as used by the reference's BlockIndexMap (the voxblox AnyIndexHash it
// derives from): x + y * 17191 + z * 17191^2, truncated to unsigned int
struct BlockKeyHash {
  static constexpr size_t sl = 17191;
  static constexpr size_t sl2 = sl * sl;
  size_t operator()(const BlockKey& k) const {
    return static_cast<unsigned int>(k.x + k.y * sl + k.z * sl2);
  }
};

// Per-block device lists come from one pool per handle: a voxblox stream opens and archives a
// few blocks per message, and a cudaMalloc / cudaFree pair per list (the latter a device-wide
// sync, either of them tens of milliseconds now and then on this platform) would land in the
// message stream.  Power-of-two size classes, free lists per class, 64 MB chunks that are never
// returned.  A released list may be handed out again at once: all work of a handle is ordered
// on one stream.
struct BlockPool {
  static constexpr int kMinClass = 10;  // 1024 entries
  std::vector<DevBuf<uint64_t>*> chunks;
  uint64_t chunk_used = 0;  // entries taken from chunks.back()
  std::vector<uint64_t*> free_lists[48];
  ~BlockPool() {
    for (DevBuf<uint64_t>* c : chunks) delete c;
  }
  static int classOf(uint64_t n) {
    int c = kMinClass;
    while ((1ull << c) < n) ++c;
    return c;
  }
  int alloc(uint64_t n, uint64_t** ptr, uint64_t* cap) {
    const int c = classOf(n);
    const uint64_t need = 1ull << c;
    if (!free_lists[c].empty()) {
      *ptr = free_lists[c].back();
      free_lists[c].pop_back();
    } else {
      if (chunks.empty() || chunk_used + need > chunks.back()->cap) {
        DevBuf<uint64_t>* chunk = new (std::nothrow) DevBuf<uint64_t>();
        PGMO_CHECK_ARG(chunk != nullptr);
        const int rc = chunk->reserve(std::max<uint64_t>(need, 1ull << 23));
        if (rc < 0) {
          delete chunk;
          return rc;
        }
        chunks.push_back(chunk);
        chunk_used = 0;
      }
      *ptr = chunks.back()->ptr + chunk_used;
      chunk_used += need;
    }
    *cap = need;
    return PGMO_OK;
  }
  void release(uint64_t* ptr, uint64_t cap) { free_lists[classOf(cap)].push_back(ptr); }
};

struct PoolBuf {  // the part of DevBuf's interface the block lists use
  BlockPool* pool = nullptr;
  uint64_t* ptr = nullptr;
  uint64_t cap = 0;
  PoolBuf() = default;
  PoolBuf(const PoolBuf&) = delete;
  PoolBuf& operator=(const PoolBuf&) = delete;
  ~PoolBuf() {
    if (ptr) pool->release(ptr, cap);
  }
  int reserve(uint64_t n) {  // contents are not kept: every list is rewritten whole
    if (n <= cap) return PGMO_OK;
    if (ptr) pool->release(ptr, cap);
    ptr = nullptr;
    cap = 0;
    return pool->alloc(n, &ptr, &cap);
  }
};

struct BlockDevState {
  explicit BlockDevState(BlockPool* pool) {
    list[0].pool = list[1].pool = face_map.pool = pool;
  }
  PoolBuf list[2];  // previous / current distinct voxel keys (ping-pong)
  int prev = 0;     // which of list[] holds prev_meshes_[block]
  bool has_prev = false;
  uint64_t prev_count = 0;
  PoolBuf face_map;  // block_face_map_[block] (mesh index per input vertex)
  uint64_t face_count = 0;
};

}  // namespace

struct pgmo_vcc {
  bool poisoned = false;  // a device-side error was reported after state had changed
  // pgmo_vcc_integrate_device may queue on a caller's stream: the event marks the end of that
  // work and every later call makes the stream it uses wait for it (calls on one handle take
  // effect in call order whatever streams they were given)
  cudaEvent_t foreign_done = nullptr;
  cudaStream_t foreign_stream = nullptr;  // non-null: work may still be running there
  int device = 0;
  double resolution = 0.0;
  cudaStream_t stream = nullptr;

  // device state
  DevBuf<uint64_t> t_keys;
  DevBuf<uint32_t> t_vals;
  DevBuf<unsigned long long> t_seq_min, t_seq_max;
  DevBuf<uint64_t> spare_keys;  // second table buffer set (rebuild target, then swapped)
  DevBuf<uint32_t> spare_vals;
  DevBuf<unsigned long long> spare_seq_min, spare_seq_max;
  uint64_t table_cap = 0;
  DevBuf<int32_t> active, inactive;
  DevBuf<uint8_t> live;
  DevBuf<float4> vertices;
  DevBuf<uint64_t> stamps;
  DevBuf<uint32_t> invalid;
  uint64_t index_cap = 0;
  DevBuf<unsigned long long> counters;     // kVccNumGlobal globals + one curr_count per block
  PinnedBuf<unsigned long long> h_counters;
  DevBuf<uint32_t> slot_of_vertex;
  DevBuf<uint32_t> tile_scratch;
  uint32_t pass = 0;

  // mirrored globals (valid after every synchronising call)
  uint64_t max_index = 0, n_live = 0, n_tomb = 0, n_invalid = 0;
  uint64_t used_upper = 0;  // upper bound of occupied (live + tombstone) slots incl. unsynced work
  bool pending = false;     // device API: counters of the last message not read back yet
  std::vector<BlockKey> pending_blocks;

  // host-side block bookkeeping (same containers / hash / operation order as the reference)
  std::unordered_map<BlockKey, double, BlockKeyHash> block_update_times;
  std::unordered_map<BlockKey, int, BlockKeyHash> block_face_map;  // value unused: order only
  std::unordered_map<BlockKey, BlockDevState*, BlockKeyHash> blocks;
  std::vector<uint32_t> archived_polygons;  // 3 per face
  BlockPool block_pool;         // backs every BlockDevState's lists
  PinnedBuf<uint64_t> h_faces;  // prune: face maps of all blocks of a call, one read-back

  // staging for the host entry points
  DevBuf<float> d_xyz;
  DevBuf<uint32_t> d_rgba;
  DevBuf<uint64_t> d_remap;
  DevBuf<int64_t> d_keys;
  DevBuf<uint64_t> d_active_list;
  VoxbloxStage vb_stage;

  ~pgmo_vcc() {
    for (auto& kv : blocks) delete kv.second;  // hands the lists back to block_pool (still alive)
  }
};

namespace {

VccTable tableOf(pgmo_vcc_t* h) {
  VccTable t;
  t.keys = h->t_keys.ptr;
  t.vals = h->t_vals.ptr;
  t.seq_min = h->t_seq_min.ptr;
  t.seq_max = h->t_seq_max.ptr;
  t.mask = static_cast<uint32_t>(h->table_cap - 1);
  return t;
}

VccIndexState indexStateOf(pgmo_vcc_t* h) {
  VccIndexState s;
  s.active = h->active.ptr;
  s.inactive = h->inactive.ptr;
  s.live = h->live.ptr;
  s.vertices = h->vertices.ptr;
  s.stamps = h->stamps.ptr;
  s.invalid = h->invalid.ptr;
  return s;
}

// `st` is about to be used: order it after what an earlier call left on another stream
int orderAfterForeign(pgmo_vcc_t* h, cudaStream_t st) {
  if (h->foreign_stream && h->foreign_stream != st) {
    PGMO_CUDA(cudaStreamWaitEvent(st, h->foreign_done, 0));
    h->foreign_stream = nullptr;
  }
  return PGMO_OK;
}

int syncCounters(pgmo_vcc_t* h, cudaStream_t st) {
  PGMO_TRY(orderAfterForeign(h, st));
  const size_t n = kVccNumGlobal + h->pending_blocks.size();
  PGMO_TRY(h->h_counters.reserve(n));
  PGMO_CUDA(cudaMemcpyAsync(h->h_counters.ptr, h->counters.ptr, n * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  h->max_index = h->h_counters.ptr[kVccMaxIndex];
  h->n_live = h->h_counters.ptr[kVccLive];
  h->n_tomb = h->h_counters.ptr[kVccTombstones];
  h->n_invalid = h->h_counters.ptr[kVccInvalid];
  for (size_t b = 0; b < h->pending_blocks.size(); ++b) {
    auto it = h->blocks.find(h->pending_blocks[b]);
    if (it != h->blocks.end()) it->second->prev_count = h->h_counters.ptr[kVccNumGlobal + b];
  }
  h->pending_blocks.clear();
  h->pending = false;
  h->used_upper = h->n_live + h->n_tomb;
  const unsigned long long err = h->h_counters.ptr[kVccError];
  if (err) {
    // the kernels that hit the error had already changed reference counts / the voxel table:
    // the handle stays in the error state instead of going on with a half-updated mesh
    h->poisoned = true;
    setLastError("voxel clearing: %s%s%s(the handle is unusable from here on: its stored mesh was "
                 "partially updated; create a new compressor)",
                 (err & 1) ? "voxel key outside the packed +-2^20 range; " : "",
                 (err & 2) ? "voxel table full; " : "",
                 (err & 4) ? "previous voxel missing from the table; " : "");
    return PGMO_ERR_UNSUPPORTED;
  }
  return PGMO_OK;
}

// Table sized for `needed` occupied slots at load <= 0.5; rehash drops tombstones.
int ensureTable(pgmo_vcc_t* h, uint64_t extra, cudaStream_t st) {
  const uint64_t needed = h->used_upper + extra;
  if (h->table_cap && 2 * needed <= h->table_cap) return PGMO_OK;
  if (h->pending) PGMO_TRY(syncCounters(h, st));
  // Sized with head room (load <= 1/8 right after a rebuild) so that rebuilds are rare: the
  // stream of 50k-vertex messages is launch-latency-bound and a rebuild costs allocations.  The
  // spare buffer set is kept on the handle: a tombstone-only rebuild of the same size allocates
  // nothing.
  // The floor (4 M slots, 112 MB per buffer set) holds half a million live voxels at that load:
  // the usual horizon-bounded stream never regrows, and the first table is built at creation.
  uint64_t cap = std::max<uint64_t>(h->table_cap, 1ull << 22);
  while (cap < 8 * (h->n_live + extra) + 16) cap <<= 1;
  if (cap > (1ull << 31)) {
    setLastError("voxel table would exceed 2^31 slots");
    return PGMO_ERR_UNSUPPORTED;
  }
  DevBuf<uint64_t>& keys = h->spare_keys;
  DevBuf<uint32_t>& vals = h->spare_vals;
  DevBuf<unsigned long long>& smin = h->spare_seq_min;
  DevBuf<unsigned long long>& smax = h->spare_seq_max;
  PGMO_TRY(keys.reserve(cap));
  PGMO_TRY(vals.reserve(cap));
  PGMO_TRY(smin.reserve(cap));
  PGMO_TRY(smax.reserve(cap));
  VccTable to;
  to.keys = keys.ptr;
  to.vals = vals.ptr;
  to.seq_min = smin.ptr;
  to.seq_max = smax.ptr;
  to.mask = static_cast<uint32_t>(cap - 1);
  PGMO_CUDA(launchVccInitTable(to, st));
  if (h->table_cap) {
    PGMO_CUDA(launchVccRehash(tableOf(h), to, h->counters.ptr, st));
    // tombstones are gone
    const unsigned long long zero = 0;
    PGMO_CUDA(cudaMemcpyAsync(h->counters.ptr + kVccTombstones, &zero, sizeof(zero),
                              cudaMemcpyHostToDevice, st));
  }
  PGMO_CUDA(cudaStreamSynchronize(st));
  std::swap(h->t_keys.ptr, keys.ptr);
  std::swap(h->t_keys.cap, keys.cap);
  std::swap(h->t_vals.ptr, vals.ptr);
  std::swap(h->t_vals.cap, vals.cap);
  std::swap(h->t_seq_min.ptr, smin.ptr);
  std::swap(h->t_seq_min.cap, smin.cap);
  std::swap(h->t_seq_max.ptr, smax.ptr);
  std::swap(h->t_seq_max.cap, smax.cap);
  h->table_cap = cap;
  h->n_tomb = 0;
  h->used_upper = h->n_live;
  // the set that was just retired becomes the target of the next rebuild: bring it to the same
  // size now, inside a call that is slow anyway, instead of inside a later message
  PGMO_TRY(keys.reserve(cap));
  PGMO_TRY(vals.reserve(cap));
  PGMO_TRY(smin.reserve(cap));
  PGMO_TRY(smax.reserve(cap));
  return PGMO_OK;
}

int ensureIndexCapacity(pgmo_vcc_t* h, uint64_t extra, cudaStream_t st) {
  // max_index mirrored on the host is a lower bound while work is pending: add what is in flight
  const uint64_t needed = h->max_index + extra;
  if (needed <= h->index_cap) return PGMO_OK;
  const uint64_t keep = h->index_cap;
  // cudaMalloc / cudaFree cost tens to hundreds of ms on this platform: grow rarely, by a lot
  const uint64_t roomy = std::max<uint64_t>(2 * needed, 1ull << 22);
  PGMO_TRY(h->active.reserve(roomy, st, keep));
  PGMO_TRY(h->inactive.reserve(roomy, st, keep));
  PGMO_TRY(h->live.reserve(roomy, st, keep));
  PGMO_TRY(h->vertices.reserve(roomy, st, keep));
  PGMO_TRY(h->stamps.reserve(roomy, st, keep));
  PGMO_TRY(h->invalid.reserve(roomy, st, keep));
  h->index_cap = std::min({h->active.cap, h->inactive.cap, h->live.cap, h->vertices.cap,
                           h->stamps.cap, h->invalid.cap});
  return PGMO_OK;
}

// updateRemapping over device-resident input; no host synchronisation unless buffers must grow.
int integrateCore(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                  const uint64_t* block_offsets, const float* d_xyz, const uint32_t* d_rgba,
                  double stamp_s, uint64_t* d_remap_out, int64_t* d_keys_out, cudaStream_t st) {
  if (h->poisoned) {
    setLastError("voxel clearing: an earlier call reported a device-side error after the stored mesh "
                 "had changed; this handle is unusable, create a new compressor");
    return PGMO_ERR_INVALID;
  }
  const uint64_t n_total = n_blocks ? block_offsets[n_blocks] : 0;
  for (uint32_t b = 0; b < n_blocks; ++b) {
    PGMO_CHECK_ARG(block_offsets[b] <= block_offsets[b + 1]);
    PGMO_CHECK_ARG(block_offsets[b + 1] - block_offsets[b] < 0xFFFFFFFFull);
  }
  {
    // a duplicate block index inside one message would need the first visit's voxel count
    // before the second visit: not supported (blockIndices() of real interfaces are unique)
    std::unordered_map<BlockKey, int, BlockKeyHash> seen;
    for (uint32_t b = 0; b < n_blocks; ++b) {
      const BlockKey key{block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]};
      if (!seen.emplace(key, 1).second) {
        setLastError("duplicate block index inside one message is not supported");
        return PGMO_ERR_UNSUPPORTED;
      }
    }
  }
  if (h->pending) PGMO_TRY(syncCounters(h, st));  // prev_count of the last message is needed
  PGMO_TRY(ensureTable(h, n_total, st));
  PGMO_TRY(ensureIndexCapacity(h, n_total, st));
  PGMO_TRY(h->slot_of_vertex.reserve(n_total + 1));
  PGMO_TRY(h->tile_scratch.reserve(vccTileScratchEntries(std::max<uint64_t>(n_total, h->index_cap)) + 1));
  PGMO_TRY(h->counters.reserve(kVccNumGlobal + n_blocks + 1, st, kVccNumGlobal));
  if (n_blocks) {
    PGMO_CUDA(cudaMemsetAsync(h->counters.ptr + kVccNumGlobal, 0,
                              n_blocks * sizeof(unsigned long long), st));
  }

  const VccTable t = tableOf(h);
  const VccIndexState s = indexStateOf(h);
  const uint64_t stamp_ns = stampFromSec(stamp_s);
  const double inv_res = 1.0 / h->resolution;

  h->pending_blocks.clear();
  for (uint32_t b = 0; b < n_blocks; ++b) {
    const BlockKey key{block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]};
    h->block_update_times[key] = stamp_s;  // :122
    BlockDevState*& state = h->blocks[key];
    if (!state) state = new BlockDevState(&h->block_pool);
    const uint64_t first = block_offsets[b];
    const uint32_t n = static_cast<uint32_t>(block_offsets[b + 1] - first);
    const int cur = state->prev ^ 1;
    PGMO_TRY(state->list[cur].reserve(std::max<uint32_t>(n, 1)));

    VccBlockArgs args;
    args.xyz = d_xyz;
    args.rgba = d_rgba;
    args.first_vertex = first;
    args.n = n;
    args.inv_resolution = inv_res;
    args.stamp_ns = stamp_ns;
    args.pass = ++h->pass;
    args.slot_of_vertex = h->slot_of_vertex.ptr;
    args.remap_out = d_remap_out;
    args.keys_out = d_keys_out;
    args.curr_list = state->list[cur].ptr;
    args.curr_count = h->counters.ptr + kVccNumGlobal + b;

    PGMO_CUDA(launchVccInsert(t, s, args, h->counters.ptr, st));
    PGMO_CUDA(launchVccAssign(t, s, args, h->tile_scratch.ptr, h->counters.ptr, st));
    PGMO_CUDA(launchVccWrite(t, s, args, st));

    // block_face_map_[block] = face_map (:166)
    h->block_face_map[key] = 1;
    PGMO_TRY(state->face_map.reserve(std::max<uint32_t>(n, 1)));
    if (n) {
      PGMO_CUDA(cudaMemcpyAsync(state->face_map.ptr, d_remap_out + first, n * sizeof(uint64_t),
                                cudaMemcpyDeviceToDevice, st));
    }
    state->face_count = n;

    if (state->has_prev) {
      // decrement every previous voxel of this block (:173-207)
      PGMO_CUDA(launchVccRelease(t, s, state->list[state->prev].ptr, state->prev_count, false,
                                 h->counters.ptr, st));
    }
    state->prev = cur;  // prev_meshes_[block] = curr_voxels
    state->has_prev = true;
    state->prev_count = 0;  // filled in by syncCounters
    h->pending_blocks.push_back(key);
  }
  h->pending = true;
  h->used_upper += n_total;
  h->max_index += 0;  // exact value arrives with syncCounters
  return PGMO_OK;
}

// pruneMeshBlocks (:72-101)
int pruneBlocks(pgmo_vcc_t* h, const std::vector<BlockKey>& to_clear, cudaStream_t st) {
  if (h->pending) PGMO_TRY(syncCounters(h, st));
  if (to_clear.empty()) return PGMO_OK;
  if (!h->table_cap) PGMO_TRY(ensureTable(h, 0, st));
  const VccTable t = tableOf(h);
  const VccIndexState s = indexStateOf(h);
  // release the voxel sets and queue every face map into one pinned buffer: one read-back per call
  uint64_t total_faces = 0;
  for (const BlockKey& key : to_clear) {
    const auto it = h->blocks.find(key);
    if (it != h->blocks.end()) total_faces += it->second->face_count;
  }
  PGMO_TRY(h->h_faces.reserve(std::max<uint64_t>(2 * total_faces, 1ull << 16)));
  uint64_t at = 0;
  for (const BlockKey& key : to_clear) {
    const auto it = h->blocks.find(key);
    if (it == h->blocks.end()) continue;
    BlockDevState* state = it->second;
    if (state->has_prev) {
      PGMO_CUDA(launchVccRelease(t, s, state->list[state->prev].ptr, state->prev_count, true,
                                 h->counters.ptr, st));
    }
    if (state->face_count) {
      PGMO_CUDA(cudaMemcpyAsync(h->h_faces.ptr + at, state->face_map.ptr,
                                state->face_count * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      at += state->face_count;
    }
  }
  if (at) PGMO_CUDA(cudaStreamSynchronize(st));
  at = 0;
  for (const BlockKey& key : to_clear) {
    auto it = h->blocks.find(key);
    if (it != h->blocks.end()) {
      BlockDevState* state = it->second;
      // polygons_.resize(archived); updatePolygons(polygons_, block_face_map_[idx]); archived = size
      const uint64_t n_faces = state->face_count / 3;
      for (uint64_t f = 0; f < 3 * n_faces; ++f) {
        h->archived_polygons.push_back(static_cast<uint32_t>(h->h_faces.ptr[at + f]));
      }
      at += state->face_count;
      delete state;  // lists back to the pool; the release kernel above is ordered before any reuse
      h->blocks.erase(it);
    }
    h->block_face_map[key];  // the reference's operator[] (:91) may rehash the map: mirror it
    h->block_face_map.erase(key);
    h->block_update_times.erase(key);
  }
  h->pending = true;  // refresh the mirrored counters
  PGMO_TRY(syncCounters(h, st));
  return PGMO_OK;
}

}  // namespace

extern "C" {

int pgmo_vcc_create(double resolution, int device, pgmo_vcc_t** out) {
  PGMO_CHECK_ARG(out != nullptr);
  *out = nullptr;
  PGMO_CHECK_ARG(resolution > 0.0);
  PGMO_TRY(ensureDevice(device));
  pgmo_vcc_t* h = new (std::nothrow) pgmo_vcc();
  PGMO_CHECK_ARG(h != nullptr);
  h->device = device;
  h->resolution = resolution;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->foreign_done, cudaEventDisableTiming) != cudaSuccess) {
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    setLastError("cudaStreamCreateWithFlags / cudaEventCreateWithFlags failed");
    return PGMO_ERR_CUDA;
  }
  int rc = h->counters.reserve(kVccNumGlobal + 64);
  if (rc == PGMO_OK &&
      cudaMemsetAsync(h->counters.ptr, 0, h->counters.cap * sizeof(unsigned long long), h->stream) !=
          cudaSuccess) {
    rc = PGMO_ERR_CUDA;
  }
  if (rc == PGMO_OK) rc = ensureTable(h, 1024, h->stream);
  if (rc != PGMO_OK) {
    pgmo_vcc_destroy(h);
    return rc;
  }
  // polygons_ of the archived blocks grows by every pruned block's faces: reserved far ahead
  // (virtual memory until touched) so that no prune call pays for moving the whole list
  h->archived_polygons.reserve(size_t(1) << 26);
  preloadVccKernels();
  preloadMeshInterfaceKernels();
  preloadCommonKernels();
  *out = h;
  return PGMO_OK;
}

void pgmo_vcc_destroy(pgmo_vcc_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->foreign_stream) cudaEventSynchronize(h->foreign_done);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  if (h->foreign_done) cudaEventDestroy(h->foreign_done);
  delete h;
}

int pgmo_vcc_integrate_device(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                              const uint64_t* block_offsets, const float* d_xyz,
                              const uint32_t* d_rgba, double stamp_s, uint64_t* d_remap_out,
                              void* stream) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_CHECK_ARG(n_blocks == 0 || (block_idx_xyz && block_offsets));
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  const uint64_t n_total = n_blocks ? block_offsets[n_blocks] : 0;
  PGMO_CHECK_ARG(n_total == 0 || (d_xyz && d_remap_out));
  PGMO_TRY(orderAfterForeign(h, st));
  const int rc = integrateCore(h, n_blocks, block_idx_xyz, block_offsets, d_xyz, d_rgba, stamp_s,
                               d_remap_out, nullptr, st);
  // no host synchronisation here: remember where the work runs
  PGMO_CUDA(cudaEventRecord(h->foreign_done, st));
  h->foreign_stream = st;
  return rc;
}

int pgmo_vcc_integrate(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                       const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                       double stamp_s, uint64_t* remap_out, int64_t* keys_out) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_CHECK_ARG(n_blocks == 0 || (block_idx_xyz && block_offsets));
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  PGMO_TRY(orderAfterForeign(h, st));
  const uint64_t n_total = n_blocks ? block_offsets[n_blocks] : 0;
  PGMO_CHECK_ARG(n_total == 0 || (xyz && remap_out));
  PGMO_TRY(h->d_xyz.reserve(3 * n_total + 3));
  PGMO_TRY(h->d_rgba.reserve(n_total + 1));
  PGMO_TRY(h->d_remap.reserve(n_total + 1));
  if (keys_out) PGMO_TRY(h->d_keys.reserve(3 * n_total + 3));
  if (n_total) {
    PGMO_CUDA(cudaMemcpyAsync(h->d_xyz.ptr, xyz, 3 * n_total * sizeof(float), cudaMemcpyHostToDevice, st));
    if (rgba) {
      PGMO_CUDA(cudaMemcpyAsync(h->d_rgba.ptr, rgba, n_total * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
  }
  PGMO_TRY(integrateCore(h, n_blocks, block_idx_xyz, block_offsets, h->d_xyz.ptr,
                         rgba ? h->d_rgba.ptr : nullptr, stamp_s, h->d_remap.ptr,
                         keys_out ? h->d_keys.ptr : nullptr, st));
  if (n_total) {
    PGMO_CUDA(cudaMemcpyAsync(remap_out, h->d_remap.ptr, n_total * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (keys_out) {
      PGMO_CUDA(cudaMemcpyAsync(keys_out, h->d_keys.ptr, 3 * n_total * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    }
  }
  return syncCounters(h, st);
}

int pgmo_vcc_integrate_voxblox(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                               const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                               const uint16_t* z, const uint8_t* r, const uint8_t* g,
                               const uint8_t* b, float block_edge_length, double stamp_s,
                               uint64_t* remap_out, int64_t* keys_out) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_CHECK_ARG(n_blocks == 0 || (block_idx_xyz && block_offsets));
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  PGMO_TRY(orderAfterForeign(h, st));
  const uint64_t n_total = n_blocks ? block_offsets[n_blocks] : 0;
  PGMO_CHECK_ARG(n_total == 0 || remap_out);
  PGMO_TRY(h->d_xyz.reserve(3 * n_total + 3));
  PGMO_TRY(h->d_rgba.reserve(n_total + 1));
  PGMO_TRY(h->d_remap.reserve(n_total + 1));
  if (keys_out) PGMO_TRY(h->d_keys.reserve(3 * n_total + 3));
  const VoxbloxInput in{x, y, z, r, g, b, block_edge_length};
  PGMO_TRY(stageVoxblox(&h->vb_stage, n_blocks, block_idx_xyz, block_offsets, in, h->d_xyz.ptr,
                        h->d_rgba.ptr, st));
  PGMO_TRY(integrateCore(h, n_blocks, block_idx_xyz, block_offsets, h->d_xyz.ptr, h->d_rgba.ptr,
                         stamp_s, h->d_remap.ptr, keys_out ? h->d_keys.ptr : nullptr, st));
  if (n_total) {
    PGMO_CUDA(cudaMemcpyAsync(remap_out, h->d_remap.ptr, n_total * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (keys_out) {
      PGMO_CUDA(cudaMemcpyAsync(keys_out, h->d_keys.ptr, 3 * n_total * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    }
  }
  return syncCounters(h, st);
}

int pgmo_vcc_prune(pgmo_vcc_t* h, double earliest_s) {
  PGMO_CHECK_ARG(h != nullptr);
  if (h->poisoned) {
    setLastError("voxel clearing: this handle is unusable after a device-side error");
    return PGMO_ERR_INVALID;
  }
  PGMO_TRY(ensureDevice(h->device));
  std::vector<BlockKey> to_clear;
  for (const auto& kv : h->block_update_times) {  // iteration order of the reference's map
    if (kv.second <= earliest_s) to_clear.push_back(kv.first);
  }
  PGMO_TRY(orderAfterForeign(h, h->stream));
  return pruneBlocks(h, to_clear, h->stream);
}

int pgmo_vcc_clear_blocks(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_CHECK_ARG(n_blocks == 0 || block_idx_xyz != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  std::vector<BlockKey> to_clear(n_blocks);
  for (uint32_t b = 0; b < n_blocks; ++b) {
    to_clear[b] = {block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]};
  }
  PGMO_TRY(orderAfterForeign(h, h->stream));
  return pruneBlocks(h, to_clear, h->stream);
}

int pgmo_vcc_device_view(pgmo_vcc_t* h, const void** d_vertices_xyzw, const uint64_t** d_stamps_ns,
                         uint64_t* n_vertices) {
  PGMO_CHECK_ARG(h != nullptr && n_vertices != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  PGMO_TRY(orderAfterForeign(h, h->stream));
  if (h->pending) PGMO_TRY(syncCounters(h, h->stream));
  PGMO_CUDA(cudaStreamSynchronize(h->stream));  // everything queued so far is visible to any stream
  if (d_vertices_xyzw) *d_vertices_xyzw = h->vertices.ptr;
  if (d_stamps_ns) *d_stamps_ns = h->stamps.ptr;
  *n_vertices = h->max_index;
  return PGMO_OK;
}

int pgmo_vcc_sizes(pgmo_vcc_t* h, uint64_t* sizes4) {
  PGMO_CHECK_ARG(h != nullptr && sizes4 != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  if (h->pending) PGMO_TRY(syncCounters(h, h->stream));
  uint64_t faces = h->archived_polygons.size() / 3;
  for (const auto& kv : h->blocks) faces += kv.second->face_count / 3;
  sizes4[0] = h->max_index;
  sizes4[1] = faces;
  sizes4[2] = h->n_live;
  sizes4[3] = h->n_invalid;
  return PGMO_OK;
}

int pgmo_vcc_snapshot(pgmo_vcc_t* h, float* vertices_xyz, uint32_t* vertices_rgba,
                      uint64_t* stamps_ns, uint32_t* polygons, uint64_t* active,
                      uint64_t* invalid) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  PGMO_TRY(orderAfterForeign(h, st));
  if (h->pending) PGMO_TRY(syncCounters(h, st));
  const uint64_t n = h->max_index;
  if ((vertices_xyz || vertices_rgba) && n) {
    std::vector<float4> v(n);
    PGMO_CUDA(cudaMemcpyAsync(v.data(), h->vertices.ptr, n * sizeof(float4), cudaMemcpyDeviceToHost, st));
    PGMO_CUDA(cudaStreamSynchronize(st));
    for (uint64_t i = 0; i < n; ++i) {
      if (vertices_xyz) {
        vertices_xyz[3 * i + 0] = v[i].x;
        vertices_xyz[3 * i + 1] = v[i].y;
        vertices_xyz[3 * i + 2] = v[i].z;
      }
      if (vertices_rgba) std::memcpy(vertices_rgba + i, &v[i].w, 4);
    }
  }
  if (stamps_ns && n) {
    PGMO_CUDA(cudaMemcpyAsync(stamps_ns, h->stamps.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  }
  if (polygons) {
    // updateVertices (:103-109): archived prefix, then every block of block_face_map_ in the
    // map's iteration order
    uint64_t at = h->archived_polygons.size();
    if (at) std::memcpy(polygons, h->archived_polygons.data(), at * sizeof(uint32_t));
    std::vector<uint64_t> faces;
    for (const auto& kv : h->block_face_map) {
      const auto it = h->blocks.find(kv.first);
      if (it == h->blocks.end() || it->second->face_count == 0) continue;
      const BlockDevState* state = it->second;
      faces.resize(state->face_count);
      PGMO_CUDA(cudaMemcpyAsync(faces.data(), state->face_map.ptr,
                                state->face_count * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      PGMO_CUDA(cudaStreamSynchronize(st));
      const uint64_t n_faces = state->face_count / 3;
      for (uint64_t f = 0; f < 3 * n_faces; ++f) polygons[at++] = static_cast<uint32_t>(faces[f]);
    }
  }
  if (active) {
    // keys of indices_to_active_refs_ (a std::map): ascending live indices (:49-55)
    PGMO_TRY(h->d_active_list.reserve(std::max<uint64_t>(h->n_live, 1)));
    PGMO_TRY(h->tile_scratch.reserve(vccTileScratchEntries(n) + 1));
    PGMO_CUDA(launchVccCompactLive(h->live.ptr, n, h->d_active_list.ptr, h->tile_scratch.ptr,
                                   h->counters.ptr + kVccNumGlobal - 1, st));
    if (h->n_live) {
      PGMO_CUDA(cudaMemcpyAsync(active, h->d_active_list.ptr, h->n_live * sizeof(uint64_t),
                                cudaMemcpyDeviceToHost, st));
    }
  }
  if (invalid && h->n_invalid) {
    std::vector<uint32_t> inv(h->n_invalid);
    PGMO_CUDA(cudaMemcpyAsync(inv.data(), h->invalid.ptr, h->n_invalid * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, st));
    PGMO_CUDA(cudaStreamSynchronize(st));
    std::sort(inv.begin(), inv.end());
    for (uint64_t i = 0; i < h->n_invalid; ++i) invalid[i] = inv[i];
  }
  PGMO_CUDA(cudaStreamSynchronize(st));
  return PGMO_OK;
}

}  // extern "C"
