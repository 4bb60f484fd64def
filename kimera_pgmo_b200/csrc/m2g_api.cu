// Mesh -> graph index bookkeeping of MeshFrontendInterface on the device (sm_100a) and its C ABI
// (pgmo_m2g_*, include/pgmo_b200.h).
//
// Replaces the members msg_to_mesh_idx_ / msg_to_graph_idx_ / mesh_to_graph_idx_
// (kimera_pgmo/include/kimera_pgmo/mesh_frontend_interface.h:118-120), the compressors' writes
// into the two HashedIndexMappings (mesh_compression.cpp:256,288,366-372,
// voxel_clearing_compression.cpp:124,154) and updateMeshToGraphMappings
// (mesh_frontend_interface.cpp:197-211).
//
// Layout in HBM:
//   per block   one segment in each of two pools (msg -> mesh, msg -> graph): dense over the msg
//               vertex index, -1 = no entry.  A segment is as long as the largest version of the
//               block seen so far and is never cleared - exactly the life cycle of the
//               reference's per-block unordered_maps, stale entries included.
//   mesh_to_graph   dense over the mesh index, -1 = no entry ("insert" keeps the first value)
//   claim       scratch of the same length for the ordered insert
// The reference's insert order decides which of several msg vertices of one mesh vertex wins.
// Across blocks that is the message's block order; inside a block it is unordered_map iteration
// order (implementation-defined, parity unpinned) - here ascending msg index, as in
// oracle/frontend_oracle.h.  The ordered insert is "flag -> smallest (block ordinal, msg index)
// claims the mesh index (atomicMin) -> the claimant writes".
#include <algorithm>
#include <array>
#include <new>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "vcc_kernels.cuh"

using namespace pgmo;

namespace {

constexpr long long kAbsent = -1;
constexpr unsigned long long kNoClaim = ~0ull;

struct BlockKey {
  int32_t x, y, z;
  bool operator==(const BlockKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct BlockKeyHash {
  size_t operator()(const BlockKey& k) const {
    return static_cast<unsigned int>(k.x + k.y * 17191ull + k.z * 17191ull * 17191ull);
  }
};

struct Segment {
  uint64_t offset = 0;  // in both pools
  uint64_t cap = 0;     // allocated entries; those beyond `len` are -1 and take part in nothing
  uint64_t len = 0;     // largest version of the block so far (= the reference's key range)
};

struct BlockRow {        // one block of the message, device side
  uint64_t seg_offset;
  uint64_t seg_cap;      // entries to walk (Segment::len)
  uint64_t msg_first;    // first input vertex of the block
  uint64_t entry_first;  // prefix sum of seg_cap over the message's blocks
};

__device__ __forceinline__ uint32_t rowOf(const BlockRow* rows, uint32_t n_rows, uint64_t v,
                                          bool by_entry) {
  uint32_t lo = 0, hi = n_rows;  // last row whose first <= v
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint64_t first = by_entry ? rows[mid].entry_first : rows[mid].msg_first;
    if (first <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// the compressors' writes: kind 0 = none, 1 = map[i] = index, 2 = map.insert({i, index})
__global__ void m2gApplyKernel(const BlockRow* __restrict__ rows, uint32_t n_rows, uint64_t n,
                               const long long* __restrict__ mesh_index,
                               const uint8_t* __restrict__ mesh_kind,
                               const long long* __restrict__ graph_index,
                               const uint8_t* __restrict__ graph_kind, long long* pool_m,
                               long long* pool_g) {
  const uint64_t v = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (v >= n) return;
  const BlockRow r = rows[rowOf(rows, n_rows, v, false)];
  const uint64_t e = r.seg_offset + (v - r.msg_first);
  const uint8_t km = mesh_kind ? mesh_kind[v] : 1, kg = graph_kind ? graph_kind[v] : 1;
  if (km == 1 || (km == 2 && pool_m[e] == kAbsent)) pool_m[e] = mesh_index[v];
  if (kg == 1 || (kg == 2 && pool_g[e] == kAbsent)) pool_g[e] = graph_index[v];
}

// stage 0: claim, 1: commit, 2: release the claims
template <int STAGE>
__global__ void m2gInsertKernel(const BlockRow* __restrict__ rows, uint32_t n_rows, uint64_t n_entries,
                                const long long* __restrict__ pool_m,
                                const long long* __restrict__ pool_g, long long* mesh_to_graph,
                                unsigned long long* claim, unsigned long long* count) {
  const uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (t >= n_entries) return;
  const uint32_t ordinal = rowOf(rows, n_rows, t, true);
  const BlockRow r = rows[ordinal];
  const uint64_t i = t - r.entry_first;
  const long long m = pool_m[r.seg_offset + i];
  if (m == kAbsent) return;
  if (STAGE == 2) {
    claim[m] = kNoClaim;
    return;
  }
  const long long g = pool_g[r.seg_offset + i];
  if (g == kAbsent || mesh_to_graph[m] != kAbsent) return;
  const unsigned long long ticket = (static_cast<unsigned long long>(ordinal) << 40) | i;
  if (STAGE == 0) {
    atomicMin(claim + m, ticket);
  } else if (claim[m] == ticket) {
    mesh_to_graph[m] = g;
    atomicAdd(count, 1ull);
  }
}

__global__ void m2gFlagKernel(const long long* __restrict__ mesh_to_graph, uint64_t n, uint8_t* flag) {
  const uint64_t m = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (m < n) flag[m] = mesh_to_graph[m] != kAbsent;
}

__global__ void m2gGatherKernel(const long long* __restrict__ mesh_to_graph,
                                const uint64_t* __restrict__ mesh_idx, uint64_t n, uint64_t* graph_idx) {
  const uint64_t j = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (j < n) graph_idx[j] = static_cast<uint64_t>(mesh_to_graph[mesh_idx[j]]);
}

}  // namespace

struct pgmo_m2g {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::unordered_map<BlockKey, Segment, BlockKeyHash> segments;
  DevBuf<long long> pool_m, pool_g;
  uint64_t pool_used = 0;
  DevBuf<long long> mesh_to_graph;
  DevBuf<unsigned long long> claim;
  uint64_t mesh_span = 0;  // initialised length of mesh_to_graph / claim
  DevBuf<unsigned long long> count;
  PinnedBuf<unsigned long long> h_count;
  // per-message staging
  DevBuf<long long> d_mesh_index, d_graph_index;
  DevBuf<uint8_t> d_mesh_kind, d_graph_kind;
  DevBuf<BlockRow> d_rows;
  PinnedBuf<BlockRow> h_rows;
  // fetch scratch
  DevBuf<uint8_t> flag;
  DevBuf<uint64_t> list, gathered;
  DevBuf<uint32_t> tiles;
};

namespace {

// room for `extra` more pool entries; both pools keep their contents
int growPools(pgmo_m2g_t* h, uint64_t extra) {
  const uint64_t need = h->pool_used + extra;
  if (need <= h->pool_m.cap) return PGMO_OK;
  // a stream of new blocks keeps appending segments, and a regrowth (cudaMalloc + copy +
  // cudaFree, twice) costs milliseconds: start at 32 MB per pool and grow fourfold
  const uint64_t want = std::max<uint64_t>(4 * need, 1ull << 22);
  PGMO_TRY(h->pool_m.reserve(want, h->stream, h->pool_used));
  PGMO_TRY(h->pool_g.reserve(want, h->stream, h->pool_used));
  return PGMO_OK;
}

// mesh_to_graph / claim cover mesh indices [0, span)
int growMeshSpan(pgmo_m2g_t* h, uint64_t span) {
  if (span <= h->mesh_span) return PGMO_OK;
  if (span > h->mesh_to_graph.cap) {
    const uint64_t want = std::max<uint64_t>(2 * span, 1ull << 20);
    PGMO_TRY(h->mesh_to_graph.reserve(want, h->stream, h->mesh_span));
    PGMO_TRY(h->claim.reserve(want, h->stream, h->mesh_span));
  }
  const uint64_t upto = h->mesh_to_graph.cap;  // initialise everything that was allocated
  PGMO_CUDA(cudaMemsetAsync(h->mesh_to_graph.ptr + h->mesh_span, 0xFF,
                            (upto - h->mesh_span) * sizeof(long long), h->stream));
  PGMO_CUDA(cudaMemsetAsync(h->claim.ptr + h->mesh_span, 0xFF,
                            (upto - h->mesh_span) * sizeof(unsigned long long), h->stream));
  h->mesh_span = upto;
  return PGMO_OK;
}

}  // namespace

extern "C" {

int pgmo_m2g_create(int device, pgmo_m2g_t** out) {
  PGMO_CHECK_ARG(out != nullptr);
  *out = nullptr;
  PGMO_TRY(ensureDevice(device));
  pgmo_m2g_t* h = new (std::nothrow) pgmo_m2g();
  PGMO_CHECK_ARG(h != nullptr);
  h->device = device;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    setLastError("cudaStreamCreateWithFlags failed");
    return PGMO_ERR_CUDA;
  }
  int rc = h->count.reserve(2);  // [0] pairs in mesh_to_graph, [1] fetch scratch
  if (rc == PGMO_OK) rc = h->h_count.reserve(1);
  if (rc == PGMO_OK &&
      cudaMemsetAsync(h->count.ptr, 0, 2 * sizeof(unsigned long long), h->stream) != cudaSuccess) {
    rc = PGMO_ERR_CUDA;
  }
  if (rc != PGMO_OK) {
    pgmo_m2g_destroy(h);
    return rc;
  }
  preloadM2gKernels();
  preloadCommonKernels();
  *out = h;
  return PGMO_OK;
}

void pgmo_m2g_destroy(pgmo_m2g_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  delete h;
}

int pgmo_m2g_update(pgmo_m2g_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                    const uint64_t* block_offsets, const int64_t* mesh_index,
                    const uint8_t* mesh_kind, const int64_t* graph_index,
                    const uint8_t* graph_kind) {
  PGMO_CHECK_ARG(h != nullptr);
  if (n_blocks == 0) return PGMO_OK;
  PGMO_CHECK_ARG(block_idx_xyz != nullptr && block_offsets != nullptr);
  PGMO_CHECK_ARG(n_blocks < (1u << 24));
  const uint64_t n = block_offsets[n_blocks];
  PGMO_CHECK_ARG(n == 0 || (mesh_index != nullptr && graph_index != nullptr));
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;

  // validate, find the mesh index range
  uint64_t span = 0;
  for (uint32_t b = 0; b < n_blocks; ++b) PGMO_CHECK_ARG(block_offsets[b] <= block_offsets[b + 1]);
  {
    // a block index twice in one message: the second occurrence could move the block's segment
    // while the row of the first still points at the old one (unique blocks are assumed, as by
    // the compressors that feed this call)
    std::vector<std::array<int32_t, 3>> seen(n_blocks);
    for (uint32_t b = 0; b < n_blocks; ++b) seen[b] = {block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]};
    std::sort(seen.begin(), seen.end());
    if (std::adjacent_find(seen.begin(), seen.end()) != seen.end()) {
      setLastError("mesh -> graph bookkeeping: a block index appears twice in one message");
      return PGMO_ERR_UNSUPPORTED;
    }
  }
  for (uint64_t v = 0; v < n; ++v) {
    const uint8_t km = mesh_kind ? mesh_kind[v] : 1, kg = graph_kind ? graph_kind[v] : 1;
    PGMO_CHECK_ARG(km <= 2 && kg <= 2);
    if (km != 0) {
      PGMO_CHECK_ARG(mesh_index[v] >= 0);
      span = std::max<uint64_t>(span, static_cast<uint64_t>(mesh_index[v]) + 1);
    }
    if (kg != 0) PGMO_CHECK_ARG(graph_index[v] >= 0);
  }
  PGMO_TRY(growMeshSpan(h, span));

  // segments: a block keeps the key range of its largest version
  PGMO_TRY(h->h_rows.reserve(n_blocks));
  uint64_t n_entries = 0;
  for (uint32_t b = 0; b < n_blocks; ++b) {
    const BlockKey key{block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]};
    const uint64_t size = block_offsets[b + 1] - block_offsets[b];
    PGMO_CHECK_ARG(size < (1ull << 40));
    Segment& seg = h->segments[key];
    if (size > seg.cap) {
      // move the segment to the end of the pools, twice the room it needs now (the old segment
      // is abandoned: with doubling, at most as much again as the block finally occupies; like
      // the reference's maps, nothing here is ever given back)
      const uint64_t room = std::max<uint64_t>(2 * size, 64);
      PGMO_TRY(growPools(h, room));
      const uint64_t to = h->pool_used;
      for (DevBuf<long long>* pool : {&h->pool_m, &h->pool_g}) {
        if (seg.cap > 0) {
          PGMO_CUDA(cudaMemcpyAsync(pool->ptr + to, pool->ptr + seg.offset, seg.cap * sizeof(long long),
                                    cudaMemcpyDeviceToDevice, st));
        }
        PGMO_CUDA(cudaMemsetAsync(pool->ptr + to + seg.cap, 0xFF, (room - seg.cap) * sizeof(long long), st));
      }
      h->pool_used += room;
      seg.offset = to;
      seg.cap = room;
    }
    seg.len = std::max(seg.len, size);
    BlockRow& r = h->h_rows.ptr[b];
    r.seg_offset = seg.offset;
    r.seg_cap = seg.len;
    r.msg_first = block_offsets[b];
    r.entry_first = n_entries;
    n_entries += seg.len;
  }

  // stage the message
  PGMO_TRY(h->d_rows.reserve(n_blocks));
  PGMO_CUDA(cudaMemcpyAsync(h->d_rows.ptr, h->h_rows.ptr, n_blocks * sizeof(BlockRow),
                            cudaMemcpyHostToDevice, st));
  if (n > 0) {
    PGMO_TRY(h->d_mesh_index.reserve(2 * n));
    PGMO_TRY(h->d_graph_index.reserve(2 * n));
    PGMO_TRY(h->d_mesh_kind.reserve(2 * n));
    PGMO_TRY(h->d_graph_kind.reserve(2 * n));
    PGMO_CUDA(cudaMemcpyAsync(h->d_mesh_index.ptr, mesh_index, n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    PGMO_CUDA(cudaMemcpyAsync(h->d_graph_index.ptr, graph_index, n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (mesh_kind) PGMO_CUDA(cudaMemcpyAsync(h->d_mesh_kind.ptr, mesh_kind, n, cudaMemcpyHostToDevice, st));
    if (graph_kind) PGMO_CUDA(cudaMemcpyAsync(h->d_graph_kind.ptr, graph_kind, n, cudaMemcpyHostToDevice, st));
    m2gApplyKernel<<<divUp(n, 256), 256, 0, st>>>(
        h->d_rows.ptr, n_blocks, n, h->d_mesh_index.ptr, mesh_kind ? h->d_mesh_kind.ptr : nullptr,
        h->d_graph_index.ptr, graph_kind ? h->d_graph_kind.ptr : nullptr, h->pool_m.ptr, h->pool_g.ptr);
  }
  if (n_entries > 0) {
    const unsigned grid = divUp(n_entries, 256);
    m2gInsertKernel<0><<<grid, 256, 0, st>>>(h->d_rows.ptr, n_blocks, n_entries, h->pool_m.ptr, h->pool_g.ptr,
                                             h->mesh_to_graph.ptr, h->claim.ptr, h->count.ptr);
    m2gInsertKernel<1><<<grid, 256, 0, st>>>(h->d_rows.ptr, n_blocks, n_entries, h->pool_m.ptr, h->pool_g.ptr,
                                             h->mesh_to_graph.ptr, h->claim.ptr, h->count.ptr);
    m2gInsertKernel<2><<<grid, 256, 0, st>>>(h->d_rows.ptr, n_blocks, n_entries, h->pool_m.ptr, h->pool_g.ptr,
                                             h->mesh_to_graph.ptr, h->claim.ptr, h->count.ptr);
  }
  PGMO_CUDA(cudaGetLastError());
  // the caller's buffers are pageable: the copies above have returned, but keep the contract
  // simple - the update is complete when the call returns
  PGMO_CUDA(cudaStreamSynchronize(st));
  return PGMO_OK;
}

int pgmo_m2g_size(pgmo_m2g_t* h, uint64_t* n_pairs) {
  PGMO_CHECK_ARG(h != nullptr && n_pairs != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  PGMO_CUDA(cudaMemcpyAsync(h->h_count.ptr, h->count.ptr, sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, h->stream));
  PGMO_CUDA(cudaStreamSynchronize(h->stream));
  *n_pairs = h->h_count.ptr[0];
  return PGMO_OK;
}

int pgmo_m2g_fetch(pgmo_m2g_t* h, uint64_t* mesh_idx, uint64_t* graph_idx) {
  PGMO_CHECK_ARG(h != nullptr && mesh_idx != nullptr && graph_idx != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  const uint64_t span = h->mesh_span;
  if (span == 0) return PGMO_OK;
  PGMO_TRY(h->flag.reserve(span));
  PGMO_TRY(h->list.reserve(span));
  PGMO_TRY(h->gathered.reserve(span));
  PGMO_TRY(h->tiles.reserve(span / 1024 + 2));
  m2gFlagKernel<<<divUp(span, 256), 256, 0, st>>>(h->mesh_to_graph.ptr, span, h->flag.ptr);
  PGMO_CUDA(launchVccCompactLive(h->flag.ptr, span, h->list.ptr, h->tiles.ptr, h->count.ptr + 1, st));
  PGMO_CUDA(cudaMemcpyAsync(h->h_count.ptr, h->count.ptr + 1, sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  const uint64_t n = h->h_count.ptr[0];
  if (n == 0) return PGMO_OK;
  m2gGatherKernel<<<divUp(n, 256), 256, 0, st>>>(h->mesh_to_graph.ptr, h->list.ptr, n, h->gathered.ptr);
  PGMO_CUDA(cudaMemcpyAsync(mesh_idx, h->list.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaMemcpyAsync(graph_idx, h->gathered.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  return PGMO_OK;
}

int pgmo_m2g_block_mapping(pgmo_m2g_t* h, const int32_t* block_idx_xyz, int which, int64_t* out,
                           uint64_t cap, uint64_t* length) {
  PGMO_CHECK_ARG(h != nullptr && block_idx_xyz != nullptr && length != nullptr);
  PGMO_CHECK_ARG(which == 0 || which == 1);
  PGMO_TRY(ensureDevice(h->device));
  const auto it = h->segments.find(BlockKey{block_idx_xyz[0], block_idx_xyz[1], block_idx_xyz[2]});
  if (it == h->segments.end()) {
    setLastError("pgmo_m2g_block_mapping: unknown block");
    return PGMO_ERR_INVALID;
  }
  *length = it->second.len;
  const uint64_t n = std::min<uint64_t>(cap, it->second.len);
  if (n > 0) {
    PGMO_CHECK_ARG(out != nullptr);
    const DevBuf<long long>& pool = which == 0 ? h->pool_m : h->pool_g;
    PGMO_CUDA(cudaMemcpyAsync(out, pool.ptr + it->second.offset, n * sizeof(int64_t),
                              cudaMemcpyDeviceToHost, h->stream));
    PGMO_CUDA(cudaStreamSynchronize(h->stream));
  }
  return PGMO_OK;
}

int pgmo_m2g_device_view(pgmo_m2g_t* h, const int64_t** d_mesh_to_graph, uint64_t* span) {
  PGMO_CHECK_ARG(h != nullptr && d_mesh_to_graph != nullptr && span != nullptr);
  PGMO_CUDA(cudaStreamSynchronize(h->stream));
  *d_mesh_to_graph = reinterpret_cast<const int64_t*>(h->mesh_to_graph.ptr);
  *span = h->mesh_span;
  return PGMO_OK;
}

}  // extern "C"

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadM2gKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, m2gApplyKernel);
  cudaFuncGetAttributes(&attr, m2gFlagKernel);
  cudaFuncGetAttributes(&attr, m2gGatherKernel);
  cudaGetLastError();
}
}  // namespace pgmo
