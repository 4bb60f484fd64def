// C ABI of DeltaCompression (pgmo_delta_*): host orchestration around delta_kernels.cu.
//
// kimera_pgmo::DeltaCompression::update / archiveBlocksByTime / clearArchivedBlocks
// (kimera_pgmo/src/compression/delta_compression.cpp:211-360).  The host keeps what the
// reference keeps per block (block_info_map_ in insertion order: update time, sequence number,
// where the block's index list and voxel list live in the device pools); everything per vertex,
// voxel and face runs on the GPU.  One private stream per handle, no CPU fallback.
#include <algorithm>
#include <array>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "common.cuh"
#include "delta_kernels.cuh"

using namespace pgmo;

namespace {

struct BlockRec {
  int32_t idx[3];
  uint64_t update_time_ns = 0;
  uint32_t seq = 0;
  uint32_t f_off = 0, f_n = 0;  // face pool segment (BlockInfo::indices)
  uint32_t x_off = 0, x_cap = 0;  // voxel pool segment (BlockInfo::vertices), capacity
  uint32_t cnt_slot = 0;          // where its voxel count lives on the device
};

using BlockKey = std::array<int32_t, 3>;

}  // namespace

struct pgmo_delta {
  int device = 0;
  cudaStream_t stream = nullptr;
  double resolution = 0.0, index_scale = 0.0;

  // voxel table + records
  DevBuf<unsigned long long> t_keys, t_first;
  DevBuf<uint32_t> t_vslot;
  uint32_t t_cap = 0;
  DevBuf<unsigned long long> r_key, r_stamp;
  DevBuf<float4> r_point;
  DevBuf<uint32_t> r_mesh_index, r_seq, r_label;
  DevBuf<int32_t> r_refs;
  DevBuf<uint8_t> r_needs_archive, r_is_new, r_alive;
  uint64_t n_records = 0;  // exact after every call
  DevBuf<uint32_t> new_id;
  // second record buffer set: compaction target, then swapped in (no allocation in steady state)
  DevBuf<unsigned long long> s_key, s_stamp;
  DevBuf<float4> s_point;
  DevBuf<uint32_t> s_mesh_index, s_seq, s_label;
  DevBuf<int32_t> s_refs;
  DevBuf<uint8_t> s_needs_archive, s_is_new, s_alive;
  uint64_t record_compactions = 0;
  uint64_t compact_slack = 1u << 19;  // PGMO_DELTA_COMPACT_SLACK: erased records tolerated

  // block_info_map_
  std::vector<BlockRec> blocks;  // insertion order, erased ones removed
  std::map<BlockKey, size_t> block_index;  // -> position in `blocks`
  DevBuf<uint32_t> f_pool, x_pool, blk_counts, f_spare, x_spare;
  DevBuf<uint4> d_segs;
  uint64_t f_used = 0, x_used = 0, f_live = 0, x_live = 0;
  std::vector<uint32_t> free_cnt_slots;
  uint32_t next_cnt_slot = 0;

  // archived_faces_
  DevBuf<uint32_t> pending, pending2;
  uint64_t n_pending = 0;

  // delta under construction / last finished delta
  bool delta_open = false;
  uint32_t vertex_start = 0, face_start = 0, prev_total = 0;
  DevBuf<float4> o_point;
  DevBuf<unsigned long long> o_stamp;
  DevBuf<uint32_t> o_label, d_label;
  bool labeled = false;  // some update carried semantic labels
  DevBuf<uint8_t> o_new, o_obs, o_del;
  DevBuf<uint32_t> o_p2c, o_faces_upd, o_faces_arch;
  uint64_t d_arch = 0, d_vertices = 0, d_faces_upd = 0, d_faces_arch = 0;  // host mirrors
  bool have_delta = false;

  uint32_t seq = 1, pass = 0;
  uint64_t num_archived_vertices = 0, num_archived_faces = 0, last_total = 0;

  // scratch
  DevBuf<float> d_xyz;
  DevBuf<uint32_t> d_rgba, tslot_of_vertex, active_remapping, tiles, work_faces;
  DevBuf<uint8_t> flags, valid, acc_a, acc_b, f1, f2;
  DevBuf<unsigned long long> d_remap, e_keys, counters;
  DevBuf<uint32_t> e_minseq;
  DevBuf<DeltaFaceBlock> d_fblocks;
  PinnedBuf<unsigned long long> h_counters;
  uint32_t last_rounds = 0;
};

namespace {

// Allocation sizes: cudaMalloc / cudaFree take tens to hundreds of milliseconds on this platform,
// so buffers are sized with generous head room (HBM is not the scarce resource here) and the
// steady state never reallocates.
inline uint64_t room(uint64_t n) { return std::max<uint64_t>(2 * n, 1ull << 22); }
template <typename Buf>
int roomy(Buf& buf, uint64_t n, cudaStream_t st = nullptr, uint64_t keep = 0) {
  if (n <= buf.cap) return PGMO_OK;
  return buf.reserve(std::max<uint64_t>(2 * n, 1ull << 20), st, keep);
}

DeltaTable tableOf(pgmo_delta* h) {
  return DeltaTable{h->t_keys.ptr, h->t_vslot.ptr, h->t_first.ptr, h->t_cap - 1};
}
DeltaRecords recordsOf(pgmo_delta* h) {
  return DeltaRecords{h->r_key.ptr, h->r_point.ptr, h->r_stamp.ptr, h->r_mesh_index.ptr, h->r_seq.ptr,
                      h->r_refs.ptr, h->r_needs_archive.ptr, h->r_is_new.ptr, h->r_alive.ptr,
                      h->r_label.ptr};
}
DeltaOut outOf(pgmo_delta* h) {
  DeltaOut o;
  o.point = h->o_point.ptr;
  o.stamp = h->o_stamp.ptr;
  o.label = h->o_label.ptr;
  o.is_new = h->o_new.ptr;
  o.observed = h->o_obs.ptr;
  o.p2c = h->o_p2c.ptr;
  o.deleted = h->o_del.ptr;
  o.prev_total = h->prev_total;
  o.vertex_start = h->vertex_start;
  o.faces_upd = h->o_faces_upd.ptr;
  o.faces_arch = h->o_faces_arch.ptr;
  return o;
}

int reserveRecords(pgmo_delta* h, uint64_t n, cudaStream_t st) {
  const uint64_t keep = h->n_records;
  PGMO_TRY(h->r_key.reserve(n, st, keep));
  PGMO_TRY(h->r_stamp.reserve(n, st, keep));
  PGMO_TRY(h->r_point.reserve(n, st, keep));
  PGMO_TRY(h->r_label.reserve(n, st, keep));
  PGMO_TRY(h->r_mesh_index.reserve(n, st, keep));
  PGMO_TRY(h->r_seq.reserve(n, st, keep));
  PGMO_TRY(h->r_refs.reserve(n, st, keep));
  PGMO_TRY(h->r_needs_archive.reserve(n, st, keep));
  PGMO_TRY(h->r_is_new.reserve(n, st, keep));
  PGMO_TRY(h->r_alive.reserve(n, st, keep));
  // the compaction target set and its scratch grow with the primary set, inside the call that
  // is slow anyway, not inside the first garbage collection of a stream
  const uint64_t cap = h->r_key.cap;
  PGMO_TRY(h->s_key.reserve(cap));
  PGMO_TRY(h->s_stamp.reserve(cap));
  PGMO_TRY(h->s_point.reserve(cap));
  PGMO_TRY(h->s_label.reserve(cap));
  PGMO_TRY(h->s_mesh_index.reserve(cap));
  PGMO_TRY(h->s_seq.reserve(cap));
  PGMO_TRY(h->s_refs.reserve(cap));
  PGMO_TRY(h->s_needs_archive.reserve(cap));
  PGMO_TRY(h->s_is_new.reserve(cap));
  PGMO_TRY(h->s_alive.reserve(cap));
  PGMO_TRY(h->new_id.reserve(cap + 1));
  return PGMO_OK;
}

// table with room for `entries` keys at load <= 0.5; rebuilt from the alive records when it grows
// (tombstones and the keys of erased records disappear)
int ensureTable(pgmo_delta* h, uint64_t entries, cudaStream_t st) {
  // floor: 4 M slots (80 MB) after the doubling below, a million voxel records at load 1/4 - a
  // horizon-bounded stream never regrows its table in the middle of the stream
  uint64_t want = 1ull << 21;
  while (want < 2 * entries + 2) want <<= 1;
  if (want <= h->t_cap) return PGMO_OK;
  PGMO_CHECK_ARG(want <= (1ull << 31));
  want <<= 1;  // head room: rebuilds are rare
  h->t_keys.release();
  h->t_first.release();
  h->t_vslot.release();
  PGMO_TRY(h->t_keys.reserve(want));
  PGMO_TRY(h->t_first.reserve(want));
  PGMO_TRY(h->t_vslot.reserve(want));
  h->t_cap = static_cast<uint32_t>(want);
  PGMO_CUDA(launchDeltaInitTable(tableOf(h), st));
  PGMO_CUDA(launchDeltaRehash(tableOf(h), recordsOf(h), h->counters.ptr,
                              static_cast<uint32_t>(h->n_records), h->counters.ptr, st));
  return PGMO_OK;
}

int readCounters(pgmo_delta* h, cudaStream_t st) {
  PGMO_CUDA(cudaMemcpyAsync(h->h_counters.ptr, h->counters.ptr, kDcCount * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  const unsigned long long err = h->h_counters.ptr[kDcError];
  if (err) {
    PGMO_CUDA(cudaMemsetAsync(h->counters.ptr + kDcError, 0, sizeof(unsigned long long), st));
    if (err & 1ull) {
      setLastError("delta compression: voxel key beyond +-2^20 voxels");
      return PGMO_ERR_UNSUPPORTED;
    }
    if (err & 4ull) {
      setLastError("delta compression: map lookup of a missing key (the reference's .at() throws)");
      return PGMO_ERR_OUT_OF_RANGE;
    }
    setLastError("delta compression: voxel table full");
    return PGMO_ERR_CUDA;
  }
  h->n_records = h->h_counters.ptr[kDcNumRecords];
  h->d_arch = h->h_counters.ptr[kDcArchVerts];
  h->d_faces_upd = h->h_counters.ptr[kDcFacesUpd];
  h->d_faces_arch = h->h_counters.ptr[kDcFacesArch];
  h->n_pending = h->h_counters.ptr[kDcPending];
  return PGMO_OK;
}

// delta_.reset(new MeshDelta(num_archived_vertices_, num_archived_faces_)) (:217-222, :321-323)
int openDelta(pgmo_delta* h, cudaStream_t st) {
  if (h->delta_open) return PGMO_OK;
  h->vertex_start = static_cast<uint32_t>(h->num_archived_vertices);
  h->face_start = static_cast<uint32_t>(h->num_archived_faces);
  h->prev_total = static_cast<uint32_t>(h->last_total);
  if (h->prev_total + 1ull > h->o_p2c.cap) PGMO_TRY(h->o_p2c.reserve(room(h->prev_total + 1ull)));
  if (h->prev_total + 1ull > h->o_del.cap) PGMO_TRY(h->o_del.reserve(room(h->prev_total + 1ull)));
  PGMO_CUDA(cudaMemsetAsync(h->o_p2c.ptr, 0xFF, (h->prev_total + 1ull) * sizeof(uint32_t), st));
  PGMO_CUDA(cudaMemsetAsync(h->o_del.ptr, 0, h->prev_total + 1ull, st));
  PGMO_CUDA(cudaMemsetAsync(h->counters.ptr + kDcArchVerts, 0, 3 * sizeof(unsigned long long), st));
  h->d_arch = h->d_faces_upd = h->d_faces_arch = 0;
  h->delta_open = true;
  h->have_delta = false;
  return PGMO_OK;
}

int reserveOutVertices(pgmo_delta* h, uint64_t n, cudaStream_t st) {
  PGMO_TRY(h->o_point.reserve(n, st, h->d_arch));
  PGMO_TRY(h->o_stamp.reserve(n, st, h->d_arch));
  PGMO_TRY(h->o_label.reserve(n, st, h->d_arch));
  return PGMO_OK;
}

int reserveTiles(pgmo_delta* h, uint64_t n) {
  if (deltaTileEntries(n) > h->tiles.cap) PGMO_TRY(h->tiles.reserve(std::max<uint64_t>(2 * deltaTileEntries(n), 1u << 16)));
  if ((n / kDeltaTile + 1) * kDeltaTile > h->flags.cap) PGMO_TRY(h->flags.reserve(room((n / kDeltaTile + 1) * kDeltaTile)));
  return PGMO_OK;
}

// RedunancyChecker over the ordered work list: the unique fixed point of "a valid face is
// accepted unless both of its first two edges come from faces accepted before it" is what the
// sequential check()/add() loop computes; Jacobi rounds from "everything valid is accepted"
// extend the correct prefix by at least one face each.
int redundancyFixedPoint(pgmo_delta* h, const uint32_t* faces, const uint8_t* valid, uint32_t n_faces,
                         const uint8_t** accepted, cudaStream_t st) {
  *accepted = valid;
  if (n_faces == 0) return PGMO_OK;
  uint64_t cap = 1024;
  while (cap < 8ull * n_faces) cap <<= 1;
  PGMO_TRY(roomy(h->e_keys, cap));
  PGMO_TRY(roomy(h->e_minseq, cap));
  PGMO_TRY(roomy(h->acc_a, n_faces));
  PGMO_TRY(roomy(h->acc_b, n_faces));
  DeltaEdgeTable e{h->e_keys.ptr, h->e_minseq.ptr, static_cast<uint32_t>(cap - 1)};
  PGMO_CUDA(cudaMemcpyAsync(h->acc_a.ptr, valid, n_faces, cudaMemcpyDeviceToDevice, st));
  uint8_t* in = h->acc_a.ptr;
  uint8_t* out = h->acc_b.ptr;
  for (uint32_t round = 0;; ++round) {
    PGMO_CUDA(launchDeltaEdgeRound(e, faces, valid, in, out, n_faces, h->counters.ptr, st));
    PGMO_CUDA(cudaMemcpyAsync(h->h_counters.ptr + kDcChanged, h->counters.ptr + kDcChanged,
                              sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    PGMO_CUDA(cudaStreamSynchronize(st));
    ++h->last_rounds;
    std::swap(in, out);
    if (h->h_counters.ptr[kDcChanged] == 0) break;
    if (round > n_faces + 2) {
      setLastError("delta compression: redundancy check did not converge");
      return PGMO_ERR_CUDA;
    }
  }
  *accepted = in;
  return PGMO_OK;
}

struct PhaseTimer {  // PGMO_DEBUG=1: wall time per phase of slow updates (stream-synchronised)
  bool on = std::getenv("PGMO_DEBUG") != nullptr;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  std::vector<std::pair<const char*, double>> marks;
  explicit PhaseTimer(cudaStream_t s) : st(s) {}
  void mark(const char* name) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    marks.emplace_back(name, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
  ~PhaseTimer() {
    double total = 0;
    for (auto& m : marks) total += m.second;
    if (on && total > 10.0) {
      fprintf(stderr, "pgmo debug: slow delta update %.1f ms:", total);
      for (auto& m : marks) fprintf(stderr, " %s %.1f", m.first, m.second);
      fprintf(stderr, "\n");
    }
  }
};

uint32_t allocCountSlot(pgmo_delta* h) {
  if (!h->free_cnt_slots.empty()) {
    const uint32_t s = h->free_cnt_slots.back();
    h->free_cnt_slots.pop_back();
    return s;
  }
  return h->next_cnt_slot++;
}

// drops the garbage of replaced / erased blocks from the two pools: the live segments are copied
// in block order into a retained second pair of pools (no allocation in steady state)
int compactPools(pgmo_delta* h, cudaStream_t st) {
  if (h->f_used <= 2 * h->f_live + h->compact_slack && h->x_used <= 2 * h->x_live + h->compact_slack) {
    return PGMO_OK;
  }
  // (sized by what is alive, never by the capacity of the pool being replaced: the two swap roles)
  if (h->f_live + 1 > h->f_spare.cap) PGMO_TRY(h->f_spare.reserve(room(h->f_live + 1)));
  if (h->x_live + 1 > h->x_spare.cap) PGMO_TRY(h->x_spare.reserve(room(h->x_live + 1)));
  std::vector<uint4> segs;
  uint64_t f = 0, x = 0;
  for (auto& b : h->blocks) {
    if (b.f_n) segs.push_back(make_uint4(b.f_off, static_cast<uint32_t>(f), b.f_n, 0));
    if (b.x_cap) segs.push_back(make_uint4(b.x_off, static_cast<uint32_t>(x), b.x_cap, 1));
    b.f_off = static_cast<uint32_t>(f);
    b.x_off = static_cast<uint32_t>(x);
    f += b.f_n;
    x += b.x_cap;
  }
  PGMO_TRY(roomy(h->d_segs, segs.size() + 1));
  PGMO_CUDA(cudaMemcpyAsync(h->d_segs.ptr, segs.data(), segs.size() * sizeof(uint4), cudaMemcpyHostToDevice, st));
  PGMO_CUDA(launchDeltaCopySegments(h->d_segs.ptr, static_cast<uint32_t>(segs.size()), h->f_pool.ptr,
                                    h->f_spare.ptr, h->x_pool.ptr, h->x_spare.ptr, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  std::swap(h->f_pool.ptr, h->f_spare.ptr);
  std::swap(h->f_pool.cap, h->f_spare.cap);
  std::swap(h->x_pool.ptr, h->x_spare.ptr);
  std::swap(h->x_pool.cap, h->x_spare.cap);
  h->f_used = f;
  h->x_used = x;
  return PGMO_OK;
}

// Drops the erased voxel records (they only accumulate: every voxel that leaves the mesh and
// comes back is a new insertion) once they outnumber the alive ones.
int compactRecords(pgmo_delta* h, uint64_t alive, cudaStream_t st) {
  if (h->n_records <= 2 * alive + h->compact_slack) return PGMO_OK;
  const uint64_t n_old = h->n_records;
  auto& key = h->s_key;
  auto& stamp = h->s_stamp;
  auto& point = h->s_point;
  auto& mesh_index = h->s_mesh_index;
  auto& seq = h->s_seq;
  auto& refs = h->s_refs;
  auto& needs_archive = h->s_needs_archive;
  auto& is_new = h->s_is_new;
  auto& alive_flag = h->s_alive;
  auto& label = h->s_label;
  if (alive + 1 > key.cap) {  // sized by what is alive, never by the set being replaced
    const uint64_t cap = room(alive + 1);
    PGMO_TRY(key.reserve(cap));
    PGMO_TRY(stamp.reserve(cap));
    PGMO_TRY(point.reserve(cap));
    PGMO_TRY(mesh_index.reserve(cap));
    PGMO_TRY(seq.reserve(cap));
    PGMO_TRY(refs.reserve(cap));
    PGMO_TRY(needs_archive.reserve(cap));
    PGMO_TRY(is_new.reserve(cap));
    PGMO_TRY(alive_flag.reserve(cap));
    PGMO_TRY(label.reserve(cap));
  }
  PGMO_TRY(roomy(h->new_id, n_old + 1));
  PGMO_TRY(reserveTiles(h, n_old));
  const DeltaRecords to{key.ptr, point.ptr, stamp.ptr, mesh_index.ptr, seq.ptr,
                        refs.ptr, needs_archive.ptr, is_new.ptr, alive_flag.ptr, label.ptr};
  PGMO_CUDA(launchDeltaCompactRecords(recordsOf(h), to, static_cast<uint32_t>(n_old), h->new_id.ptr,
                                      h->x_pool.ptr, h->x_used, h->tiles.ptr, h->counters.ptr, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  auto swapBuf = [](auto& a, auto& b) {
    std::swap(a.ptr, b.ptr);
    std::swap(a.cap, b.cap);
  };
  swapBuf(h->r_key, key);
  swapBuf(h->r_stamp, stamp);
  swapBuf(h->r_point, point);
  swapBuf(h->r_mesh_index, mesh_index);
  swapBuf(h->r_seq, seq);
  swapBuf(h->r_refs, refs);
  swapBuf(h->r_needs_archive, needs_archive);
  swapBuf(h->r_is_new, is_new);
  swapBuf(h->r_alive, alive_flag);
  swapBuf(h->r_label, label);
  h->n_records = alive;
  ++h->record_compactions;
  // fresh table over the compacted records (same capacity: tombstones are gone)
  PGMO_CUDA(launchDeltaInitTable(tableOf(h), st));
  PGMO_CUDA(launchDeltaRehash(tableOf(h), recordsOf(h), h->counters.ptr, static_cast<uint32_t>(alive),
                              h->counters.ptr, st));
  return readCounters(h, st);
}

int archiveChosen(pgmo_delta* h, const std::vector<size_t>& chosen) {
  cudaStream_t st = h->stream;
  PGMO_TRY(openDelta(h, st));
  if (chosen.empty()) return PGMO_OK;
  // ---- vertex loop (:329-349), block by block in block_info_map_ order
  uint64_t x_total = 0, f_total = 0;
  for (const size_t bi : chosen) {
    x_total += h->blocks[bi].x_cap;
    f_total += h->blocks[bi].f_n / 3;
  }
  PGMO_TRY(reserveOutVertices(h, h->d_arch + x_total + 1, st));
  for (const size_t bi : chosen) {
    const BlockRec& b = h->blocks[bi];
    PGMO_TRY(reserveTiles(h, b.x_cap));
    PGMO_CUDA(launchDeltaRelease(tableOf(h), recordsOf(h), outOf(h), h->x_pool.ptr + b.x_off,
                                 h->blk_counts.ptr + b.cnt_slot, b.x_cap, true, h->flags.ptr,
                                 h->tiles.ptr, h->counters.ptr, st));
  }
  // ---- archiveBlockFaces over the chosen blocks with one shared checker (:351-356, :362-414)
  const uint32_t T = static_cast<uint32_t>(f_total);
  if (T > 0) {
    std::vector<DeltaFaceBlock> fb;
    uint32_t first = 0;
    for (const size_t bi : chosen) {
      const BlockRec& b = h->blocks[bi];
      if (b.f_n < 3) continue;
      fb.push_back(DeltaFaceBlock{b.f_off, b.f_n, first, -1});
      first += b.f_n / 3;
    }
    PGMO_TRY(roomy(h->d_fblocks, fb.size()));
    PGMO_CUDA(cudaMemcpyAsync(h->d_fblocks.ptr, fb.data(), fb.size() * sizeof(DeltaFaceBlock),
                              cudaMemcpyHostToDevice, st));
    PGMO_CUDA(cudaStreamSynchronize(st));  // fb is a stack vector
    PGMO_TRY(roomy(h->work_faces, 3ull * T));
    PGMO_TRY(roomy(h->f1, T));
    PGMO_TRY(roomy(h->f2, T));
    PGMO_CUDA(launchDeltaArchiveFaces(h->d_fblocks.ptr, static_cast<uint32_t>(fb.size()), T, h->f_pool.ptr,
                                      outOf(h), static_cast<uint32_t>(h->num_archived_vertices),
                                      h->counters.ptr, h->work_faces.ptr, h->f1.ptr, h->f2.ptr, st));
    const uint8_t* acc = nullptr;
    PGMO_TRY(redundancyFixedPoint(h, h->work_faces.ptr, h->f2.ptr, T, &acc, st));
    PGMO_TRY(reserveTiles(h, T));
    PGMO_TRY(roomy(h->o_faces_arch, 3 * (h->d_faces_arch + T) + 3, st, 3 * h->d_faces_arch));
    PGMO_CUDA(launchDeltaAppendFaces(h->work_faces.ptr, acc, T, h->o_faces_arch.ptr,
                                     h->counters.ptr + kDcFacesArch, h->tiles.ptr, st));
    PGMO_TRY(roomy(h->pending, 3 * (h->n_pending + T) + 3, st, 3 * h->n_pending));
    PGMO_CUDA(launchDeltaAppendFaces(h->work_faces.ptr, h->f1.ptr, T, h->pending.ptr,
                                     h->counters.ptr + kDcPending, h->tiles.ptr, st));
  }
  // ---- block_info_map_.erase(idx)
  std::vector<char> dead(h->blocks.size(), 0);
  for (const size_t bi : chosen) dead[bi] = 1;
  std::vector<BlockRec> kept;
  h->block_index.clear();
  for (size_t i = 0; i < h->blocks.size(); ++i) {
    const BlockRec& b = h->blocks[i];
    if (dead[i]) {
      h->free_cnt_slots.push_back(b.cnt_slot);
      h->f_live -= b.f_n;
      h->x_live -= b.x_cap;
      continue;
    }
    h->block_index[BlockKey{b.idx[0], b.idx[1], b.idx[2]}] = kept.size();
    kept.push_back(b);
  }
  h->blocks.swap(kept);
  return readCounters(h, st);
}

}  // namespace

namespace {

// Every buffer of the handle at its floor size, once, at creation: what a 50k-vertex / 10 s
// horizon stream needs is then in place before the first message (first update 11 ms -> the
// usual 0.4 ms; the first archive and the first garbage collections no longer allocate).
int prewarm(pgmo_delta* h) {
  cudaStream_t st = h->stream;
  PGMO_TRY(reserveRecords(h, room(1), st));
  PGMO_TRY(ensureTable(h, 1, st));
  PGMO_TRY(h->f_pool.reserve(room(1)));
  PGMO_TRY(h->x_pool.reserve(room(1)));
  PGMO_TRY(h->f_spare.reserve(h->f_pool.cap));
  PGMO_TRY(h->x_spare.reserve(h->x_pool.cap));
  PGMO_TRY(h->o_p2c.reserve(room(1)));
  PGMO_TRY(h->o_del.reserve(room(1)));
  PGMO_TRY(h->o_new.reserve(room(1)));
  PGMO_TRY(h->o_obs.reserve(room(1)));
  PGMO_TRY(reserveOutVertices(h, room(1), st));
  PGMO_TRY(reserveTiles(h, room(1)));
  PGMO_TRY(roomy(h->e_keys, 1));
  PGMO_TRY(roomy(h->e_minseq, 1));
  PGMO_TRY(roomy(h->acc_a, 1));
  PGMO_TRY(roomy(h->acc_b, 1));
  PGMO_TRY(roomy(h->d_segs, 1));
  PGMO_TRY(roomy(h->d_fblocks, 1));
  PGMO_TRY(roomy(h->work_faces, 1));
  PGMO_TRY(roomy(h->f1, 1));
  PGMO_TRY(roomy(h->f2, 1));
  PGMO_TRY(roomy(h->valid, 1));
  PGMO_TRY(roomy(h->o_faces_arch, 1));
  PGMO_TRY(roomy(h->o_faces_upd, 1));
  PGMO_TRY(roomy(h->pending, 1));
  PGMO_TRY(roomy(h->pending2, 1));
  PGMO_TRY(roomy(h->d_xyz, 1));
  PGMO_TRY(roomy(h->d_rgba, 1));
  PGMO_TRY(roomy(h->tslot_of_vertex, 1));
  PGMO_TRY(roomy(h->active_remapping, 1));
  PGMO_TRY(roomy(h->d_remap, 1));
  PGMO_TRY(roomy(h->blk_counts, 1));
  PGMO_CUDA(cudaStreamSynchronize(st));
  return PGMO_OK;
}

}  // namespace

extern "C" {

int pgmo_delta_create(double resolution, int device, pgmo_delta_t** out) {
  PGMO_CHECK_ARG(out != nullptr && resolution > 0.0);
  *out = nullptr;
  PGMO_TRY(ensureDevice(device));
  pgmo_delta* h = new (std::nothrow) pgmo_delta();
  PGMO_CHECK_ARG(h != nullptr);
  h->device = device;
  h->resolution = resolution;
  h->index_scale = 1.0 / resolution;  // :49
  if (const char* e = std::getenv("PGMO_DELTA_COMPACT_SLACK")) h->compact_slack = std::strtoull(e, nullptr, 10);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    setLastError("cudaStreamCreate failed");
    return PGMO_ERR_CUDA;
  }
  if (h->counters.reserve(kDcCount) < 0 || h->h_counters.reserve(kDcCount) < 0) {
    delete h;
    return PGMO_ERR_CUDA;
  }
  cudaMemsetAsync(h->counters.ptr, 0, kDcCount * sizeof(unsigned long long), h->stream);
  if (prewarm(h) != PGMO_OK) {
    pgmo_delta_destroy(h);
    return PGMO_ERR_CUDA;
  }
  preloadDeltaKernels();
  preloadCommonKernels();
  *out = h;
  return PGMO_OK;
}

void pgmo_delta_destroy(pgmo_delta_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  delete h;
}

int pgmo_delta_update(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                      const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                      uint64_t timestamp_ns, uint64_t* remap_out) {
  return pgmo_delta_update_labeled(h, n_blocks, block_idx_xyz, block_offsets, xyz, rgba, nullptr,
                                   timestamp_ns, remap_out);
}

int pgmo_delta_update_labeled(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                              const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                              const uint32_t* labels, uint64_t timestamp_ns, uint64_t* remap_out) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_CHECK_ARG(n_blocks == 0 || (block_idx_xyz && block_offsets));
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  const uint64_t n_total = n_blocks ? block_offsets[n_blocks] : 0;
  PGMO_CHECK_ARG(n_total == 0 || xyz != nullptr);
  PGMO_CHECK_ARG(n_total < 0x7FFFFFFFull);
  for (uint32_t b = 0; b < n_blocks; ++b) {
    if ((block_offsets[b + 1] - block_offsets[b]) % 3 != 0) {
      setLastError("delta compression: a block's vertex count is not a multiple of 3 (Face::at throws)");
      return PGMO_ERR_OUT_OF_RANGE;
    }
    for (uint32_t c = 0; c < b; ++c) {
      if (std::memcmp(block_idx_xyz + 3 * b, block_idx_xyz + 3 * c, 3 * sizeof(int32_t)) == 0) {
        setLastError("delta compression: a block index appears twice in one message");
        return PGMO_ERR_UNSUPPORTED;
      }
    }
  }
  h->last_rounds = 0;
  PhaseTimer timer(st);
  PGMO_TRY(openDelta(h, st));  // 1) archive delta, if any, is the starting point (:215-222)

  // ---- inputs and capacities
  PGMO_TRY(roomy(h->d_xyz, 3 * n_total + 3));
  PGMO_TRY(roomy(h->d_rgba, n_total + 1));
  if (n_total) {
    PGMO_CUDA(cudaMemcpyAsync(h->d_xyz.ptr, xyz, 3 * n_total * sizeof(float), cudaMemcpyHostToDevice, st));
    if (rgba) PGMO_CUDA(cudaMemcpyAsync(h->d_rgba.ptr, rgba, n_total * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  }
  if (labels) {
    PGMO_TRY(roomy(h->d_label, n_total + 1));
    if (n_total) PGMO_CUDA(cudaMemcpyAsync(h->d_label.ptr, labels, n_total * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    h->labeled = true;
  }
  if (h->n_records + n_total + 1 > h->r_key.cap) PGMO_TRY(reserveRecords(h, room(h->n_records + n_total + 1), st));
  PGMO_TRY(ensureTable(h, h->n_records + n_total, st));
  PGMO_TRY(roomy(h->tslot_of_vertex, n_total + 1));
  PGMO_TRY(roomy(h->active_remapping, n_total + 1));
  PGMO_TRY(roomy(h->d_remap, n_total + 1));
  PGMO_CUDA(cudaMemsetAsync(h->d_remap.ptr, 0xFF, (n_total + 1) * sizeof(unsigned long long), st));
  PGMO_CUDA(cudaMemsetAsync(h->counters.ptr + kDcActiveSlots, 0, sizeof(unsigned long long), st));
  const uint64_t out_cap = h->d_arch + 2 * h->n_records + n_total + 1;
  if (out_cap > h->o_point.cap) PGMO_TRY(reserveOutVertices(h, room(out_cap), st));
  if (out_cap > h->o_new.cap) PGMO_TRY(h->o_new.reserve(room(out_cap)));
  if (out_cap > h->o_obs.cap) PGMO_TRY(h->o_obs.reserve(room(out_cap)));
  PGMO_CUDA(cudaMemsetAsync(h->o_new.ptr, 0, out_cap, st));
  PGMO_CUDA(cudaMemsetAsync(h->o_obs.ptr, 0, out_cap, st));

  timer.mark("reserve");
  // ---- 2) updateRemapping (:261-296), block by block
  std::vector<int64_t> msg_offset_of_block(h->blocks.size(), -1);
  for (uint32_t b = 0; b < n_blocks; ++b) {
    const uint32_t n_b = static_cast<uint32_t>(block_offsets[b + 1] - block_offsets[b]);
    const BlockKey key{block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]};
    auto it = h->block_index.find(key);
    const bool is_block_new = it == h->block_index.end();
    if (is_block_new) {
      BlockRec rec;
      rec.idx[0] = key[0], rec.idx[1] = key[1], rec.idx[2] = key[2];
      it = h->block_index.emplace(key, h->blocks.size()).first;
      h->blocks.push_back(rec);
      msg_offset_of_block.push_back(-1);
    }
    BlockRec& rec = h->blocks[it->second];
    const BlockRec prev = rec;
    msg_offset_of_block[it->second] = static_cast<int64_t>(block_offsets[b]);
    rec.update_time_ns = timestamp_ns;
    rec.seq = h->seq;
    // new segments at the end of the pools (the old ones become garbage)
    const auto tr0 = std::chrono::steady_clock::now();
    if (h->f_used + n_b + 1 > h->f_pool.cap) PGMO_TRY(h->f_pool.reserve(room(h->f_used + n_b + 1), st, h->f_used));
    if (h->x_used + n_b + 1 > h->x_pool.cap) PGMO_TRY(h->x_pool.reserve(room(h->x_used + n_b + 1), st, h->x_used));
    PGMO_TRY(h->f_spare.reserve(h->f_pool.cap));  // the garbage collection's target pair
    PGMO_TRY(h->x_spare.reserve(h->x_pool.cap));
    if (timer.on) {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count();
      if (ms > 5.0) fprintf(stderr, "pgmo debug: pool reserve %.1f ms (f_used %llu cap %zu, x_used %llu cap %zu)\n", ms,
                            (unsigned long long)h->f_used, h->f_pool.cap, (unsigned long long)h->x_used, h->x_pool.cap);
    }
    rec.f_off = static_cast<uint32_t>(h->f_used);
    rec.f_n = n_b;
    rec.x_off = static_cast<uint32_t>(h->x_used);
    rec.x_cap = n_b;
    rec.cnt_slot = allocCountSlot(h);
    h->f_used += n_b;
    h->x_used += n_b;
    h->f_live += n_b;
    h->x_live += n_b;
    PGMO_TRY(roomy(h->blk_counts, h->next_cnt_slot + 1, st, h->next_cnt_slot));
    PGMO_TRY(reserveTiles(h, std::max<uint64_t>(n_b, prev.x_cap)));

    DeltaBlockArgs a;
    a.xyz = h->d_xyz.ptr;
    a.rgba = rgba ? h->d_rgba.ptr : nullptr;
    a.label = labels ? h->d_label.ptr : nullptr;
    a.first_vertex = block_offsets[b];
    a.n = n_b;
    a.index_scale = h->index_scale;
    a.stamp_ns = timestamp_ns;
    a.pass = ++h->pass;
    a.seq = h->seq;
    a.tslot_of_vertex = h->tslot_of_vertex.ptr;
    a.face_map = h->f_pool.ptr + rec.f_off;
    a.voxels = h->x_pool.ptr + rec.x_off;
    a.voxel_count = h->blk_counts.ptr + rec.cnt_slot;
    a.active_remapping = h->active_remapping.ptr;
    if (n_b <= static_cast<uint32_t>(kDeltaTile) && prev.x_cap <= static_cast<uint32_t>(kDeltaTile)) {
      PGMO_CUDA(launchDeltaSmallBlock(tableOf(h), recordsOf(h), outOf(h), a,
                                      is_block_new ? nullptr : h->x_pool.ptr + prev.x_off,
                                      is_block_new ? nullptr : h->blk_counts.ptr + prev.cnt_slot,
                                      h->counters.ptr, st));
    } else {
      PGMO_CUDA(launchDeltaAddBlock(tableOf(h), recordsOf(h), a, h->tiles.ptr, h->counters.ptr, st));
      if (!is_block_new) {  // removeBlockObservations(block_info.vertices) (:291-293)
        PGMO_CUDA(launchDeltaRelease(tableOf(h), recordsOf(h), outOf(h), h->x_pool.ptr + prev.x_off,
                                     h->blk_counts.ptr + prev.cnt_slot, prev.x_cap, false, h->flags.ptr,
                                     h->tiles.ptr, h->counters.ptr, st));
      }
    }
    if (!is_block_new) {
      h->free_cnt_slots.push_back(prev.cnt_slot);
      h->f_live -= prev.f_n;
      h->x_live -= prev.x_cap;
    }
  }

  timer.mark("blocks");
  // ---- 3) addActiveVertices (:128-154)
  const uint64_t rec_upper = h->n_records + n_total;
  PGMO_TRY(reserveTiles(h, rec_upper));
  PGMO_CUDA(launchDeltaActiveVertices(recordsOf(h), outOf(h), static_cast<uint32_t>(rec_upper), h->seq,
                                      h->active_remapping.ptr, h->tiles.ptr, h->counters.ptr, st));

  // ---- addActiveFaces (:156-209): every face of every block in block_info_map_ order
  std::vector<DeltaFaceBlock> fb;
  uint32_t T = 0;
  for (size_t i = 0; i < h->blocks.size(); ++i) {
    const BlockRec& b = h->blocks[i];
    if (b.f_n < 3) continue;
    fb.push_back(DeltaFaceBlock{b.f_off, b.f_n, T, msg_offset_of_block[i]});
    T += b.f_n / 3;
  }
  PGMO_TRY(roomy(h->o_faces_upd, 3 * (static_cast<uint64_t>(T) + h->n_pending) + 3));
  if (T > 0) {
    PGMO_TRY(roomy(h->d_fblocks, fb.size()));
    PGMO_CUDA(cudaMemcpyAsync(h->d_fblocks.ptr, fb.data(), fb.size() * sizeof(DeltaFaceBlock),
                              cudaMemcpyHostToDevice, st));
    PGMO_TRY(roomy(h->work_faces, 3ull * T));
    PGMO_TRY(roomy(h->valid, T));
    PGMO_CUDA(launchDeltaRemapFaces(h->d_fblocks.ptr, static_cast<uint32_t>(fb.size()), T, h->f_pool.ptr,
                                    h->active_remapping.ptr, outOf(h), h->d_remap.ptr,
                                    h->work_faces.ptr, h->valid.ptr, h->counters.ptr, st));
    const uint8_t* acc = nullptr;
    PGMO_TRY(redundancyFixedPoint(h, h->work_faces.ptr, h->valid.ptr, T, &acc, st));  // syncs: fb is done
    PGMO_TRY(reserveTiles(h, T));
    PGMO_CUDA(launchDeltaAppendFaces(h->work_faces.ptr, acc, T, h->o_faces_upd.ptr,
                                     h->counters.ptr + kDcFacesUpd, h->tiles.ptr, st));
  }

  timer.mark("vertices+faces");
  // ---- 4) updateAndAddArchivedFaces (:416-446)
  if (h->n_pending > 0) {
    const uint32_t P = static_cast<uint32_t>(h->n_pending);
    PGMO_TRY(roomy(h->valid, P));
    PGMO_TRY(roomy(h->f1, P));
    PGMO_TRY(roomy(h->f2, P));
    PGMO_TRY(roomy(h->pending2, 3ull * P + 3));
    PGMO_CUDA(launchDeltaPendingRemap(h->pending.ptr, P, outOf(h),
                                      static_cast<uint32_t>(h->num_archived_vertices), h->valid.ptr, st));
    const uint8_t* acc = nullptr;
    PGMO_TRY(redundancyFixedPoint(h, h->pending.ptr, h->valid.ptr, P, &acc, st));
    PGMO_CUDA(launchDeltaClassifyPending(h->pending.ptr, acc, P, h->counters.ptr, outOf(h), h->f1.ptr,
                                         h->f2.ptr, st));
    PGMO_TRY(reserveTiles(h, P));
    PGMO_TRY(roomy(h->o_faces_arch, 3 * (h->d_faces_arch + P) + 3, st, 3 * h->d_faces_arch));
    PGMO_CUDA(launchDeltaAppendFaces(h->pending.ptr, h->f1.ptr, P, h->o_faces_arch.ptr,
                                     h->counters.ptr + kDcFacesArch, h->tiles.ptr, st));
    PGMO_CUDA(launchDeltaAppendFaces(h->pending.ptr, h->f2.ptr, P, h->o_faces_upd.ptr,
                                     h->counters.ptr + kDcFacesUpd, h->tiles.ptr, st));
    PGMO_CUDA(cudaMemsetAsync(h->counters.ptr + kDcPending, 0, sizeof(unsigned long long), st));
    PGMO_CUDA(launchDeltaAppendFaces(h->pending.ptr, h->f2.ptr, P, h->pending2.ptr,
                                     h->counters.ptr + kDcPending, h->tiles.ptr, st));
    std::swap(h->pending.ptr, h->pending2.ptr);
    std::swap(h->pending.cap, h->pending2.cap);
  }

  // ---- 5) book-keeping (:245-257)
  if (remap_out && n_total) {
    PGMO_CUDA(cudaMemcpyAsync(remap_out, h->d_remap.ptr, n_total * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  }
  PGMO_TRY(readCounters(h, st));
  timer.mark("pending+sync");
  h->d_vertices = h->d_arch + h->h_counters.ptr[kDcScratch];
  h->num_archived_vertices = h->vertex_start + h->d_arch;
  h->num_archived_faces = h->face_start + h->d_faces_arch;
  h->last_total = h->vertex_start + h->d_vertices;
  h->seq = (h->seq + 1) & 0xFFFFu;
  if (h->seq == 0) h->seq = 1;
  h->delta_open = false;
  h->have_delta = true;
  PGMO_TRY(compactPools(h, st));
  timer.mark("compact pools");
  const int rc = compactRecords(h, h->d_vertices - h->d_arch, st);
  timer.mark("compact records");
  return rc;
}

int pgmo_delta_archive_by_time(pgmo_delta_t* h, uint64_t earliest_time_ns) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  std::vector<size_t> chosen;
  for (size_t i = 0; i < h->blocks.size(); ++i) {
    if (h->blocks[i].update_time_ns <= earliest_time_ns) chosen.push_back(i);  // :306-310
  }
  return archiveChosen(h, chosen);
}

int pgmo_delta_archive_blocks(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz) {
  PGMO_CHECK_ARG(h != nullptr && (n_blocks == 0 || block_idx_xyz != nullptr));
  PGMO_TRY(ensureDevice(h->device));
  std::vector<size_t> chosen;
  for (size_t i = 0; i < h->blocks.size(); ++i) {
    for (uint32_t b = 0; b < n_blocks; ++b) {
      if (std::memcmp(h->blocks[i].idx, block_idx_xyz + 3 * b, 3 * sizeof(int32_t)) == 0) {
        chosen.push_back(i);
        break;
      }
    }
  }
  return archiveChosen(h, chosen);
}

int pgmo_delta_sizes(pgmo_delta_t* h, uint64_t* s) {
  PGMO_CHECK_ARG(h != nullptr && s != nullptr);
  for (int i = 0; i < 12; ++i) s[i] = 0;
  if (h->have_delta) {
    s[0] = h->vertex_start;
    s[1] = h->face_start;
    s[2] = h->d_vertices;
    s[3] = h->d_arch;
    s[4] = h->d_faces_upd;
    s[5] = h->d_faces_arch;
    s[6] = h->prev_total;  // capacity for prev_to_curr / deleted (their index space)
    s[7] = h->prev_total;
    s[8] = h->d_vertices;  // capacity for observed / new
    s[9] = h->d_vertices;
  }
  s[10] = h->have_delta ? h->d_vertices - h->d_arch : 0;
  s[11] = h->n_pending;
  return PGMO_OK;
}

int pgmo_delta_fetch_labels(pgmo_delta_t* h, uint32_t* labels, int* any_label) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  if (!h->have_delta) {
    setLastError("pgmo_delta_fetch_labels: no finished delta (call pgmo_delta_update)");
    return PGMO_ERR_INVALID;
  }
  if (any_label) *any_label = h->labeled ? 1 : 0;
  if (labels && h->d_vertices) {
    PGMO_CUDA(cudaMemcpyAsync(labels, h->o_label.ptr, h->d_vertices * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                              h->stream));
    PGMO_CUDA(cudaStreamSynchronize(h->stream));
  }
  return PGMO_OK;
}

int pgmo_delta_fetch(pgmo_delta_t* h, float* vertices_xyz, uint32_t* vertices_rgba, uint64_t* stamps_ns,
                     uint32_t* faces, uint32_t* faces_archive, uint64_t* prev_to_curr2,
                     uint64_t* n_prev_to_curr, uint64_t* deleted, uint64_t* n_deleted,
                     uint64_t* observed, uint64_t* n_observed, uint64_t* new_indices, uint64_t* n_new) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  if (!h->have_delta) {
    setLastError("pgmo_delta_fetch: no finished delta (call pgmo_delta_update)");
    return PGMO_ERR_INVALID;
  }
  cudaStream_t st = h->stream;
  const uint64_t nv = h->d_vertices;
  std::vector<float4> pts(nv);
  std::vector<uint8_t> f_new(nv), f_obs(nv), f_del(h->prev_total);
  std::vector<uint32_t> p2c(h->prev_total);
  if (nv) {
    PGMO_CUDA(cudaMemcpyAsync(pts.data(), h->o_point.ptr, nv * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (stamps_ns) PGMO_CUDA(cudaMemcpyAsync(stamps_ns, h->o_stamp.ptr, nv * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PGMO_CUDA(cudaMemcpyAsync(f_new.data(), h->o_new.ptr, nv, cudaMemcpyDeviceToHost, st));
    PGMO_CUDA(cudaMemcpyAsync(f_obs.data(), h->o_obs.ptr, nv, cudaMemcpyDeviceToHost, st));
  }
  if (h->prev_total) {
    PGMO_CUDA(cudaMemcpyAsync(f_del.data(), h->o_del.ptr, h->prev_total, cudaMemcpyDeviceToHost, st));
    PGMO_CUDA(cudaMemcpyAsync(p2c.data(), h->o_p2c.ptr, h->prev_total * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  if (faces && h->d_faces_upd) {
    PGMO_CUDA(cudaMemcpyAsync(faces, h->o_faces_upd.ptr, 3 * h->d_faces_upd * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  if (faces_archive && h->d_faces_arch) {
    PGMO_CUDA(cudaMemcpyAsync(faces_archive, h->o_faces_arch.ptr, 3 * h->d_faces_arch * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  PGMO_CUDA(cudaStreamSynchronize(st));
  for (uint64_t i = 0; i < nv; ++i) {
    if (vertices_xyz) {
      vertices_xyz[3 * i] = pts[i].x;
      vertices_xyz[3 * i + 1] = pts[i].y;
      vertices_xyz[3 * i + 2] = pts[i].z;
    }
    if (vertices_rgba) std::memcpy(vertices_rgba + i, &pts[i].w, sizeof(uint32_t));
  }
  // std::map / std::set contents in key order
  uint64_t k = 0;
  for (uint64_t i = 0; i < h->prev_total; ++i) {
    if (p2c[i] == kDeltaNone) continue;
    if (prev_to_curr2) {
      prev_to_curr2[2 * k] = i;
      prev_to_curr2[2 * k + 1] = p2c[i];
    }
    ++k;
  }
  if (n_prev_to_curr) *n_prev_to_curr = k;
  k = 0;
  for (uint64_t i = 0; i < h->prev_total; ++i) {
    if (!f_del[i]) continue;
    if (deleted) deleted[k] = i;
    ++k;
  }
  if (n_deleted) *n_deleted = k;
  uint64_t ko = 0, kn = 0;
  for (uint64_t i = 0; i < nv; ++i) {
    if (f_obs[i]) {
      if (observed) observed[ko] = h->vertex_start + i;
      ++ko;
    }
    if (f_new[i]) {
      if (new_indices) new_indices[kn] = h->vertex_start + i;
      ++kn;
    }
  }
  if (n_observed) *n_observed = ko;
  if (n_new) *n_new = kn;
  return PGMO_OK;
}

int pgmo_delta_blocks(pgmo_delta_t* h, uint64_t capacity, int32_t* block_idx_xyz,
                      uint64_t* update_time_ns, uint32_t* sequence_number, uint64_t* n_blocks) {
  PGMO_CHECK_ARG(h != nullptr && n_blocks != nullptr);
  *n_blocks = h->blocks.size();
  for (uint64_t i = 0; i < h->blocks.size() && i < capacity; ++i) {
    if (block_idx_xyz) std::memcpy(block_idx_xyz + 3 * i, h->blocks[i].idx, 3 * sizeof(int32_t));
    if (update_time_ns) update_time_ns[i] = h->blocks[i].update_time_ns;
    if (sequence_number) sequence_number[i] = h->blocks[i].seq;
  }
  return PGMO_OK;
}

int pgmo_delta_last_stats(pgmo_delta_t* h, uint64_t* stats4) {
  PGMO_CHECK_ARG(h != nullptr && stats4 != nullptr);
  stats4[0] = h->last_rounds;
  stats4[1] = h->n_records;
  stats4[2] = h->blocks.size();
  stats4[3] = h->record_compactions;
  return PGMO_OK;
}

}  // extern "C"
