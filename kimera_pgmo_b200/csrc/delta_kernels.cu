// DeltaCompression kernels (sm_100a).  See delta_kernels.cuh for the data layout.
//
// Everything the reference does vertex by vertex in container-iteration order is done here as
// "flag, ordered rank (tile scan), apply": ranks come from an exclusive scan over the elements
// in the reference's order, so indices are assigned exactly as the sequential code assigns them.
#include "delta_kernels.cuh"

namespace pgmo {

namespace {

constexpr unsigned long long kEmpty = 0ull;
constexpr unsigned long long kTomb = ~0ull;
constexpr int kKeyBits = 21;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// LongIndex(round(p.x * scale), ...) — float promoted to double, one rounded product,
// std::round (half away from zero), conversion to int64 (delta_compression.cpp:61-63)
__device__ __forceinline__ bool voxelKey(const float* __restrict__ xyz, uint64_t v, double scale,
                                         unsigned long long* packed) {
  const long long lim = 1ll << (kKeyBits - 1);
  long long k[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    k[a] = static_cast<long long>(round(__dmul_rn(static_cast<double>(xyz[3 * v + a]), scale)));
    if (k[a] < -lim || k[a] >= lim) return false;
  }
  *packed = (static_cast<unsigned long long>(k[0] + lim) |
             (static_cast<unsigned long long>(k[1] + lim) << kKeyBits) |
             (static_cast<unsigned long long>(k[2] + lim) << (2 * kKeyBits))) + 1;
  return true;
}

// inserts only into empty slots (tombstones are skipped and disappear at the next rehash), so
// that concurrent inserts of one key always meet in the same slot
__device__ __forceinline__ bool findOrInsert(const DeltaTable& t, unsigned long long key, uint32_t* slot) {
  uint32_t h = static_cast<uint32_t>(mix64(key)) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    const unsigned long long cur = t.keys[h];
    if (cur == key) {
      *slot = h;
      return true;
    }
    if (cur == kEmpty) {
      const unsigned long long old = atomicCAS(t.keys + h, kEmpty, key);
      if (old == kEmpty || old == key) {
        *slot = h;
        return true;
      }
    }
    h = (h + 1) & t.mask;
  }
  return false;
}

__device__ __forceinline__ bool findSlot(const DeltaTable& t, unsigned long long key, uint32_t* slot) {
  uint32_t h = static_cast<uint32_t>(mix64(key)) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    const unsigned long long cur = t.keys[h];
    if (cur == key) {
      *slot = h;
      return true;
    }
    if (cur == kEmpty) return false;
    h = (h + 1) & t.mask;
  }
  return false;
}

__device__ __forceinline__ void eraseRecord(const DeltaTable& t, const DeltaRecords& r, uint32_t vs,
                                            unsigned long long* counters) {
  uint32_t slot;
  if (findSlot(t, r.key[vs], &slot)) {
    t.vslot[slot] = kDeltaNone;
    t.keys[slot] = kTomb;
  } else {
    atomicOr(counters + kDcError, 4ull);
  }
  r.alive[vs] = 0;
}

// exclusive scan over the 1024 threads of a CTA of a word packing three counts (21 bits each)
__device__ __forceinline__ unsigned long long blockScanPacked(unsigned long long v,
                                                             unsigned long long* total) {
  __shared__ unsigned long long warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, w, o);
      if (lane >= o) w += up;
    }
    warp_sums[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const unsigned long long before = warp > 0 ? warp_sums[warp - 1] : 0ull;
  *total = warp_sums[31];
  __syncthreads();
  return before + incl - v;
}

__device__ __forceinline__ unsigned long long pack3(bool a, bool b, bool c) {
  return (a ? 1ull : 0ull) | (b ? (1ull << 21) : 0ull) | (c ? (1ull << 42) : 0ull);
}
__device__ __forceinline__ uint32_t field(unsigned long long w, int k) {
  return static_cast<uint32_t>((w >> (21 * k)) & 0x1FFFFFull);
}

// tiles: 3 u32 per tile.  Turns per-tile counts into exclusive offsets, writes the old values of
// the three target counters (the bases) to base_out[0..2] and adds the totals to them; a null
// target keeps a private count (base 0).  totals_out (optional) receives the three totals.
// One warp per field, 32 tiles per step.
__global__ void scanTiles3Kernel(uint32_t* tiles, uint32_t n_tiles, unsigned long long* c0,
                                 unsigned long long* c1, unsigned long long* c2,
                                 uint32_t* base_out, unsigned long long* totals_out) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k >= 3) return;
  uint32_t run = 0;
  for (uint32_t t0 = 0; t0 < n_tiles; t0 += 32) {
    const uint32_t t = t0 + lane;
    const uint32_t c = t < n_tiles ? tiles[3 * t + k] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += up;
    }
    if (t < n_tiles) tiles[3 * t + k] = run + incl - c;
    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  if (lane != 0) return;
  unsigned long long* target = k == 0 ? c0 : (k == 1 ? c1 : c2);
  uint32_t base = 0;
  if (target) {
    base = static_cast<uint32_t>(*target);
    *target += run;
  }
  base_out[k] = base;
  if (totals_out) totals_out[k] = run;
}

// ---------------------------------------------------------------------------------- table
__global__ void initTableKernel(DeltaTable t) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > t.mask) return;
  t.keys[i] = kEmpty;
  t.vslot[i] = kDeltaNone;
  t.first[i] = ~0ull;
}

__global__ void rehashKernel(DeltaTable t, DeltaRecords r, const unsigned long long* counters,
                             uint32_t n_upper, unsigned long long* err) {
  const uint32_t vs = blockIdx.x * blockDim.x + threadIdx.x;
  if (vs >= n_upper || vs >= counters[kDcNumRecords] || !r.alive[vs]) return;
  uint32_t slot;
  if (!findOrInsert(t, r.key[vs], &slot)) {
    atomicOr(err + kDcError, 2ull);
    return;
  }
  t.vslot[slot] = vs;
}

// ---------------------------------------------------------------------------------- add block
__global__ void __launch_bounds__(256)
addInsertKernel(const DeltaTable t, const DeltaBlockArgs b, unsigned long long* counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  const uint64_t v = b.first_vertex + i;
  unsigned long long key;
  uint32_t slot = kDeltaNone;
  if (!voxelKey(b.xyz, v, b.index_scale, &key)) {
    atomicOr(counters + kDcError, 1ull);
  } else if (!findOrInsert(t, key, &slot)) {
    atomicOr(counters + kDcError, 2ull);
    slot = kDeltaNone;
  }
  b.tslot_of_vertex[v] = slot;
  if (slot == kDeltaNone) return;
  // first occurrence in this block: newer passes win through the inverted pass id
  atomicMin(t.first + slot, (static_cast<unsigned long long>(~b.pass) << 32) | i);
}

struct AddFlags {
  bool first, fresh, pass_slot;
  uint32_t tslot;
};
__device__ __forceinline__ AddFlags addFlags(const DeltaTable& t, const DeltaRecords& r,
                                             const DeltaBlockArgs& b, uint32_t i) {
  AddFlags f{false, false, false, kDeltaNone};
  if (i >= b.n) return f;
  f.tslot = b.tslot_of_vertex[b.first_vertex + i];
  if (f.tslot == kDeltaNone) return f;
  f.first = t.first[f.tslot] == ((static_cast<unsigned long long>(~b.pass) << 32) | i);
  if (!f.first) return f;
  const uint32_t vs = t.vslot[f.tslot];
  f.fresh = vs == kDeltaNone;                       // vertices_map_.find() == end()
  f.pass_slot = f.fresh || r.seq[vs] != b.seq;      // info.sequence_number != sequence_number_
  return f;
}

__global__ void __launch_bounds__(kDeltaTile)
addCountKernel(const DeltaTable t, const DeltaRecords r, const DeltaBlockArgs b, uint32_t* tiles) {
  const uint32_t i = blockIdx.x * kDeltaTile + threadIdx.x;
  const AddFlags f = addFlags(t, r, b, i);
  unsigned long long total;
  blockScanPacked(pack3(f.first, f.fresh, f.pass_slot), &total);
  if (threadIdx.x < 3) tiles[3 * blockIdx.x + threadIdx.x] = field(total, threadIdx.x);
}

__global__ void __launch_bounds__(kDeltaTile)
addApplyKernel(const DeltaTable t, const DeltaRecords r, const DeltaBlockArgs b,
               const uint32_t* __restrict__ tiles, const uint32_t* __restrict__ bases) {
  const uint32_t i = blockIdx.x * kDeltaTile + threadIdx.x;
  const AddFlags f = addFlags(t, r, b, i);
  unsigned long long total;
  const unsigned long long ex = blockScanPacked(pack3(f.first, f.fresh, f.pass_slot), &total);
  if (!f.first) return;
  const uint32_t pos = tiles[3 * blockIdx.x + 0] + field(ex, 0);  // in curr_voxels
  uint32_t vs;
  if (f.fresh) {
    vs = bases[1] + tiles[3 * blockIdx.x + 1] + field(ex, 1);  // next record, insertion order
    t.vslot[f.tslot] = vs;
    r.key[vs] = t.keys[f.tslot];
    r.mesh_index[vs] = 0;
    r.seq[vs] = 0;
    r.refs[vs] = 0;
    r.needs_archive[vs] = 0;
    r.is_new[vs] = 1;
    r.alive[vs] = 1;
  } else {
    vs = t.vslot[f.tslot];
  }
  if (f.pass_slot) {  // first observation of the voxel in this pass (:72-84)
    const uint64_t v = b.first_vertex + i;
    const uint32_t a = bases[2] + tiles[3 * blockIdx.x + 2] + field(ex, 2);
    r.stamp[vs] = b.stamp_ns;
    r.point[vs] = make_float4(b.xyz[3 * v], b.xyz[3 * v + 1], b.xyz[3 * v + 2],
                              __uint_as_float(b.rgba ? b.rgba[v] : 0xFF000000u));
    r.label[vs] = b.label ? b.label[v] : kDeltaNoLabel;  // info.label = semantic_label (:76)
    b.active_remapping[a] = r.mesh_index[vs];  // cache previous index
    r.mesh_index[vs] = a;
    r.seq[vs] = b.seq;
  }
  r.refs[vs] += 1;  // one observation per block (:87-90); one thread per voxel and block
  b.voxels[pos] = vs;
}

__global__ void __launch_bounds__(256)
addFaceMapKernel(const DeltaTable t, const DeltaRecords r, const DeltaBlockArgs b) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  const uint32_t tslot = b.tslot_of_vertex[b.first_vertex + i];
  b.face_map[i] = tslot == kDeltaNone ? 0u : r.mesh_index[t.vslot[tslot]];
}

// ---------------------------------------------------------------------------------- release
__global__ void __launch_bounds__(kDeltaTile)
releaseFlagKernel(const DeltaTable t, const DeltaRecords r, const DeltaOut o,
                  const uint32_t* __restrict__ voxels, const uint32_t* __restrict__ voxel_count,
                  bool archive, uint8_t* __restrict__ flags, uint32_t* tiles,
                  unsigned long long* counters) {
  const uint32_t j = blockIdx.x * kDeltaTile + threadIdx.x;
  bool to_archive = false;
  if (j < *voxel_count) {
    const uint32_t vs = voxels[j];
    if (!r.alive[vs]) {
      atomicOr(counters + kDcError, 4ull);  // vertices_map_.at() would throw
    } else {
      if (archive) r.needs_archive[vs] = 1;  // archiveObservation
      const int refs = --r.refs[vs];
      if (!r.needs_archive[vs] && refs <= 0) {  // notObserved(): gone from the mesh
        const uint32_t prev = r.mesh_index[vs];
        if (prev < o.prev_total) o.deleted[prev] = 1;
        eraseRecord(t, r, vs, counters);
      } else if (r.needs_archive[vs] && refs <= 0) {
        to_archive = true;
      }
    }
  }
  flags[j] = to_archive ? 1 : 0;
  unsigned long long total;
  blockScanPacked(pack3(to_archive, false, false), &total);
  if (threadIdx.x < 3) tiles[3 * blockIdx.x + threadIdx.x] = field(total, threadIdx.x);
}

__global__ void __launch_bounds__(kDeltaTile)
releaseApplyKernel(const DeltaTable t, const DeltaRecords r, const DeltaOut o,
                   const uint32_t* __restrict__ voxels, const uint8_t* __restrict__ flags,
                   const uint32_t* __restrict__ tiles, const uint32_t* __restrict__ bases,
                   unsigned long long* counters) {
  const uint32_t j = blockIdx.x * kDeltaTile + threadIdx.x;
  const bool to_archive = flags[j] != 0;
  unsigned long long total;
  const unsigned long long ex = blockScanPacked(pack3(to_archive, false, false), &total);
  if (!to_archive) return;
  const uint32_t vs = voxels[j];
  const uint32_t local = bases[0] + tiles[3 * blockIdx.x] + field(ex, 0);  // delta->addVertex(archive)
  o.point[local] = r.point[vs];
  o.stamp[local] = r.stamp[vs];
    o.label[local] = r.label[vs];
  const uint32_t prev = r.mesh_index[vs];
  if (prev < o.prev_total) o.p2c[prev] = o.vertex_start + local;
  eraseRecord(t, r, vs, counters);
}

// ---------------------------------------------------------------------------------- vertices
__global__ void __launch_bounds__(kDeltaTile)
activeCountKernel(const DeltaRecords r, const unsigned long long* counters, uint32_t* tiles) {
  const uint32_t vs = blockIdx.x * kDeltaTile + threadIdx.x;
  const bool alive = vs < counters[kDcNumRecords] && r.alive[vs];
  unsigned long long total;
  blockScanPacked(pack3(alive, false, false), &total);
  if (threadIdx.x < 3) tiles[3 * blockIdx.x + threadIdx.x] = field(total, threadIdx.x);
}

__global__ void __launch_bounds__(kDeltaTile)
activeApplyKernel(const DeltaRecords r, const DeltaOut o, uint32_t seq, uint32_t* active_remapping,
                  const uint32_t* __restrict__ tiles, const unsigned long long* counters) {
  const uint32_t vs = blockIdx.x * kDeltaTile + threadIdx.x;
  const bool alive = vs < counters[kDcNumRecords] && r.alive[vs];
  unsigned long long total;
  const unsigned long long ex = blockScanPacked(pack3(alive, false, false), &total);
  if (!alive) return;
  const uint32_t local = static_cast<uint32_t>(counters[kDcArchVerts]) + tiles[3 * blockIdx.x] + field(ex, 0);
  const uint32_t new_idx = o.vertex_start + local;
  o.point[local] = r.point[vs];
  o.stamp[local] = r.stamp[vs];
    o.label[local] = r.label[vs];
  const uint32_t mi = r.mesh_index[vs];
  if (r.seq[vs] != seq) {
    if (mi < o.prev_total) o.p2c[mi] = new_idx;  // not seen this pass (:131-133)
  } else {
    if (!r.is_new[vs]) {
      const uint32_t prev = active_remapping[mi];
      if (prev < o.prev_total) o.p2c[prev] = new_idx;
    } else {
      o.is_new[local] = 1;
      r.is_new[vs] = 0;
    }
    o.observed[local] = 1;
    active_remapping[mi] = new_idx;
  }
  r.mesh_index[vs] = new_idx;
}

// ---------------------------------------------------------------------------------- faces
__device__ __forceinline__ uint32_t blockOfFace(const DeltaFaceBlock* blocks, uint32_t n_blocks, uint32_t f) {
  uint32_t lo = 0, hi = n_blocks;  // last block with first_face <= f and faces
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (blocks[mid].first_face <= f) lo = mid;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
remapFacesKernel(const DeltaFaceBlock* __restrict__ blocks, uint32_t n_blocks, uint32_t n_faces,
                 uint32_t* pool, const uint32_t* __restrict__ active_remapping, const DeltaOut o,
                 unsigned long long* remap_out, uint32_t* work_faces, uint8_t* valid,
                 unsigned long long* counters) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_faces) return;
  const DeltaFaceBlock blk = blocks[blockOfFace(blocks, n_blocks, f)];
  const uint32_t i = 3 * (f - blk.first_face);
  uint32_t v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint32_t old = pool[blk.face_offset + i + c];
    uint32_t now;
    if (blk.msg_offset >= 0) {
      now = active_remapping[old];
    } else {
      now = old < o.prev_total ? o.p2c[old] : kDeltaNone;  // prev_to_curr.at()
      if (now == kDeltaNone) {
        atomicOr(counters + kDcError, 4ull);
        now = old;
      }
    }
    pool[blk.face_offset + i + c] = now;
    if (blk.msg_offset >= 0 && remap_out) remap_out[blk.msg_offset + i + c] = now;
    v[c] = now;
    work_faces[3 * f + c] = now;
  }
  valid[f] = (v[0] != v[1] && v[0] != v[2] && v[1] != v[2]) ? 1 : 0;
}

__device__ __forceinline__ void edgeInsert(const DeltaEdgeTable& e, uint32_t a, uint32_t b, uint32_t seq) {
  const unsigned long long key = ((static_cast<unsigned long long>(a) << 32) | b) + 1;
  uint32_t h = static_cast<uint32_t>(mix64(key)) & e.mask;
  for (uint32_t probe = 0; probe <= e.mask; ++probe) {
    const unsigned long long cur = e.keys[h];
    if (cur == key) break;
    if (cur == kEmpty) {
      const unsigned long long old = atomicCAS(e.keys + h, kEmpty, key);
      if (old == kEmpty || old == key) break;
    }
    h = (h + 1) & e.mask;
  }
  atomicMin(e.minseq + h, seq);
}
__device__ __forceinline__ uint32_t edgeSeq(const DeltaEdgeTable& e, uint32_t a, uint32_t b) {
  const unsigned long long key = ((static_cast<unsigned long long>(a) << 32) | b) + 1;
  uint32_t h = static_cast<uint32_t>(mix64(key)) & e.mask;
  for (uint32_t probe = 0; probe <= e.mask; ++probe) {
    const unsigned long long cur = e.keys[h];
    if (cur == key) return e.minseq[h];
    if (cur == kEmpty) return kDeltaNone;
    h = (h + 1) & e.mask;
  }
  return kDeltaNone;
}

__global__ void __launch_bounds__(256)
edgeInsertKernel(const DeltaEdgeTable e, const uint32_t* __restrict__ faces,
                 const uint8_t* __restrict__ acc, uint32_t n_faces) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_faces || !acc[f]) return;
  const uint32_t a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
  edgeInsert(e, a, b, f);  // RedunancyChecker::add
  edgeInsert(e, b, c, f);
  edgeInsert(e, c, a, f);
}

__global__ void __launch_bounds__(256)
edgeDecideKernel(const DeltaEdgeTable e, const uint32_t* __restrict__ faces,
                 const uint8_t* __restrict__ valid, const uint8_t* __restrict__ acc_in,
                 uint8_t* __restrict__ acc_out, uint32_t n_faces, unsigned long long* counters) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_faces) return;
  uint8_t acc = 0;
  if (valid[f]) {
    const uint32_t a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
    // check(): present = hasEdge(v1, v2) && hasEdge(v2, v3) [&& hasEdge(v2, v3)], edges of the
    // faces accepted before this one
    const bool present = edgeSeq(e, a, b) < f && edgeSeq(e, b, c) < f;
    acc = present ? 0 : 1;
  }
  acc_out[f] = acc;
  if (acc != acc_in[f]) atomicOr(counters + kDcChanged, 1ull);
}

__global__ void __launch_bounds__(kDeltaTile)
flagCountKernel(const uint8_t* __restrict__ flags, uint32_t n, uint32_t* tiles) {
  const uint32_t i = blockIdx.x * kDeltaTile + threadIdx.x;
  const bool on = i < n && flags[i];
  unsigned long long total;
  blockScanPacked(pack3(on, false, false), &total);
  if (threadIdx.x < 3) tiles[3 * blockIdx.x + threadIdx.x] = field(total, threadIdx.x);
}

__global__ void __launch_bounds__(kDeltaTile)
appendFacesKernel(const uint32_t* __restrict__ faces, const uint8_t* __restrict__ flags, uint32_t n,
                  uint32_t* out, const uint32_t* __restrict__ tiles, const uint32_t* __restrict__ bases) {
  const uint32_t i = blockIdx.x * kDeltaTile + threadIdx.x;
  const bool on = i < n && flags[i];
  unsigned long long total;
  const unsigned long long ex = blockScanPacked(pack3(on, false, false), &total);
  if (!on) return;
  const uint32_t at = bases[0] + tiles[3 * blockIdx.x] + field(ex, 0);
  out[3 * at] = faces[3 * i];
  out[3 * at + 1] = faces[3 * i + 1];
  out[3 * at + 2] = faces[3 * i + 2];
}

__global__ void __launch_bounds__(256)
pendingRemapKernel(uint32_t* faces, uint32_t n_faces, const DeltaOut o, uint32_t num_archived,
                   uint8_t* valid) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_faces) return;
  uint32_t v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t idx = faces[3 * f + c];
    // getRemappedIndex(prev_to_curr, idx, num_archived_vertices_) (:26-33)
    if (idx >= num_archived && idx < o.prev_total && o.p2c[idx] != kDeltaNone) idx = o.p2c[idx];
    faces[3 * f + c] = idx;
    v[c] = idx;
  }
  valid[f] = (v[0] != v[1] && v[0] != v[2] && v[1] != v[2]) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
classifyPendingKernel(const uint32_t* __restrict__ faces, const uint8_t* __restrict__ acc,
                      uint32_t n_faces, const unsigned long long* counters, const DeltaOut o,
                      uint8_t* to_arch, uint8_t* to_upd) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_faces) return;
  const uint32_t threshold = o.vertex_start + static_cast<uint32_t>(counters[kDcArchVerts]);
  const bool can = faces[3 * f] < threshold && faces[3 * f + 1] < threshold && faces[3 * f + 2] < threshold;
  to_arch[f] = (acc[f] && can) ? 1 : 0;
  to_upd[f] = (acc[f] && !can) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
archiveFacesKernel(const DeltaFaceBlock* __restrict__ blocks, uint32_t n_blocks, uint32_t n_faces,
                   const uint32_t* __restrict__ pool, const DeltaOut o, uint32_t num_archived,
                   const unsigned long long* counters, uint32_t* work_faces, uint8_t* to_pending,
                   uint8_t* candidate) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_faces) return;
  const DeltaFaceBlock blk = blocks[blockOfFace(blocks, n_blocks, f)];
  const uint32_t i = 3 * (f - blk.first_face);
  const uint32_t threshold = o.vertex_start + static_cast<uint32_t>(counters[kDcArchVerts]);
  uint32_t v[3], m[3];
  bool all_can = true;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    v[c] = pool[blk.face_offset + i + c];
    const bool mapped = v[c] < o.prev_total && o.p2c[v[c]] != kDeltaNone;
    m[c] = mapped ? o.p2c[v[c]] : v[c];
    all_can = all_can && (mapped || v[c] < num_archived);
  }
  bool pend = false, cand = false;
  uint32_t* w = work_faces + 3 * f;
  if (!all_can) {  // not remapped: cannot tell later whether an index was remapped (:383-389)
    pend = true;
    w[0] = v[0], w[1] = v[1], w[2] = v[2];
  } else {
    w[0] = m[0], w[1] = m[1], w[2] = m[2];
    const bool valid = m[0] != m[1] && m[0] != m[2] && m[1] != m[2];
    if (valid) {
      if (!(m[0] < threshold && m[1] < threshold && m[2] < threshold)) pend = true;
      else cand = true;
    }
  }
  to_pending[f] = pend ? 1 : 0;
  candidate[f] = cand ? 1 : 0;
}

// ---------------------------------------------------------------------------------- small blocks
// A voxblox-like message is dozens of blocks of a few hundred vertices.  For a block (and its
// previous voxel list) of at most 1024 entries one CTA does the whole of updateRemapping's body
// — addPoint for every vertex, then removeBlockObservations — in a single launch; the phases
// are the kernels above, separated by __syncthreads() instead of launch boundaries.
__global__ void __launch_bounds__(kDeltaTile)
smallBlockKernel(const DeltaTable t, const DeltaRecords r, const DeltaOut o, const DeltaBlockArgs b,
                 const uint32_t* __restrict__ prev_voxels, const uint32_t* __restrict__ prev_count,
                 unsigned long long* counters) {
  const uint32_t i = threadIdx.x;
  // ---- addPoint: hash insert + first occurrence
  if (i < b.n) {
    const uint64_t v = b.first_vertex + i;
    unsigned long long key;
    uint32_t slot = kDeltaNone;
    if (!voxelKey(b.xyz, v, b.index_scale, &key)) {
      atomicOr(counters + kDcError, 1ull);
    } else if (!findOrInsert(t, key, &slot)) {
      atomicOr(counters + kDcError, 2ull);
      slot = kDeltaNone;
    }
    b.tslot_of_vertex[v] = slot;
    if (slot != kDeltaNone) atomicMin(t.first + slot, (static_cast<unsigned long long>(~b.pass) << 32) | i);
  }
  __syncthreads();
  const uint32_t base_rec = static_cast<uint32_t>(counters[kDcNumRecords]);
  const uint32_t base_act = static_cast<uint32_t>(counters[kDcActiveSlots]);
  const uint32_t base_arch = static_cast<uint32_t>(counters[kDcArchVerts]);
  // ---- ordered ranks + apply
  {
    const AddFlags f = addFlags(t, r, b, i);
    unsigned long long total;
    const unsigned long long ex = blockScanPacked(pack3(f.first, f.fresh, f.pass_slot), &total);
    if (f.first) {
      const uint32_t pos = field(ex, 0);
      uint32_t vs;
      if (f.fresh) {
        vs = base_rec + field(ex, 1);
        t.vslot[f.tslot] = vs;
        r.key[vs] = t.keys[f.tslot];
        r.mesh_index[vs] = 0;
        r.seq[vs] = 0;
        r.refs[vs] = 0;
        r.needs_archive[vs] = 0;
        r.is_new[vs] = 1;
        r.alive[vs] = 1;
      } else {
        vs = t.vslot[f.tslot];
      }
      if (f.pass_slot) {
        const uint64_t v = b.first_vertex + i;
        const uint32_t a = base_act + field(ex, 2);
        r.stamp[vs] = b.stamp_ns;
        r.point[vs] = make_float4(b.xyz[3 * v], b.xyz[3 * v + 1], b.xyz[3 * v + 2],
                                  __uint_as_float(b.rgba ? b.rgba[v] : 0xFF000000u));
        r.label[vs] = b.label ? b.label[v] : kDeltaNoLabel;
        b.active_remapping[a] = r.mesh_index[vs];
        r.mesh_index[vs] = a;
        r.seq[vs] = b.seq;
      }
      r.refs[vs] += 1;
      b.voxels[pos] = vs;
    }
    if (i == 0) {
      *b.voxel_count = field(total, 0);
      counters[kDcNumRecords] = base_rec + field(total, 1);
      counters[kDcActiveSlots] = base_act + field(total, 2);
    }
  }
  __syncthreads();
  // ---- face map
  if (i < b.n) {
    const uint32_t tslot = b.tslot_of_vertex[b.first_vertex + i];
    b.face_map[i] = tslot == kDeltaNone ? 0u : r.mesh_index[t.vslot[tslot]];
  }
  if (!prev_voxels) return;
  __syncthreads();
  // ---- removeBlockObservations over the previous voxel list
  bool to_archive = false;
  uint32_t vs = 0;
  if (i < *prev_count) {
    vs = prev_voxels[i];
    if (!r.alive[vs]) {
      atomicOr(counters + kDcError, 4ull);
    } else {
      const int refs = --r.refs[vs];
      if (!r.needs_archive[vs] && refs <= 0) {
        const uint32_t prev = r.mesh_index[vs];
        if (prev < o.prev_total) o.deleted[prev] = 1;
        eraseRecord(t, r, vs, counters);
      } else if (r.needs_archive[vs] && refs <= 0) {
        to_archive = true;
      }
    }
  }
  unsigned long long total;
  const unsigned long long ex = blockScanPacked(pack3(to_archive, false, false), &total);
  if (to_archive) {
    const uint32_t local = base_arch + field(ex, 0);
    o.point[local] = r.point[vs];
    o.stamp[local] = r.stamp[vs];
    o.label[local] = r.label[vs];
    const uint32_t prev = r.mesh_index[vs];
    if (prev < o.prev_total) o.p2c[prev] = o.vertex_start + local;
    eraseRecord(t, r, vs, counters);
  }
  if (i == 0) counters[kDcArchVerts] = base_arch + field(total, 0);
}

// ---------------------------------------------------------------------------------- compaction
// Erased records leave holes; insertion order is all that matters, so the alive ones are moved
// to the front in order (stable), the block voxel lists are renumbered and the table is rebuilt.
__global__ void __launch_bounds__(kDeltaTile)
compactMoveKernel(const DeltaRecords from, const DeltaRecords to, const uint32_t* __restrict__ tiles,
                  const unsigned long long* counters, uint32_t* __restrict__ new_id) {
  const uint32_t vs = blockIdx.x * kDeltaTile + threadIdx.x;
  const bool in_range = vs < counters[kDcNumRecords];
  const bool alive = in_range && from.alive[vs];
  unsigned long long total;
  const unsigned long long ex = blockScanPacked(pack3(alive, false, false), &total);
  if (in_range) new_id[vs] = kDeltaNone;
  if (!alive) return;
  const uint32_t n = tiles[3 * blockIdx.x] + field(ex, 0);
  new_id[vs] = n;
  to.key[n] = from.key[vs];
  to.point[n] = from.point[vs];
  to.stamp[n] = from.stamp[vs];
  to.label[n] = from.label[vs];
  to.mesh_index[n] = from.mesh_index[vs];
  to.seq[n] = from.seq[vs];
  to.refs[n] = from.refs[vs];
  to.needs_archive[n] = from.needs_archive[vs];
  to.is_new[n] = from.is_new[vs];
  to.alive[n] = 1;
}

__global__ void renumberVoxelsKernel(uint32_t* pool, uint64_t n, const uint32_t* __restrict__ new_id,
                                     uint32_t n_old) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t old = pool[i];
  const uint32_t now = old < n_old ? new_id[old] : kDeltaNone;
  pool[i] = now == kDeltaNone ? 0u : now;  // entries beyond a list's count / of replaced lists are garbage
}

__global__ void setCounterKernel(unsigned long long* dst, const unsigned long long* src) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *dst = *src;
}

__global__ void copySegmentsKernel(const uint4* __restrict__ segs, const uint32_t* __restrict__ src_a,
                                   uint32_t* __restrict__ dst_a, const uint32_t* __restrict__ src_b,
                                   uint32_t* __restrict__ dst_b) {
  // segs[i] = {src offset, dst offset, count, pool (0 = a, 1 = b)}: one CTA per segment
  const uint4 sg = segs[blockIdx.x];
  const uint32_t* src = (sg.w ? src_b : src_a) + sg.x;
  uint32_t* dst = (sg.w ? dst_b : dst_a) + sg.y;
  for (uint32_t i = threadIdx.x; i < sg.z; i += blockDim.x) dst[i] = src[i];
}

inline unsigned tilesFor(uint64_t n) { return static_cast<unsigned>((n + kDeltaTile - 1) / kDeltaTile); }

}  // namespace

uint64_t deltaTileEntries(uint64_t n) { return 3ull * tilesFor(n) + 16; }

cudaError_t launchDeltaInitTable(const DeltaTable& t, cudaStream_t s) {
  initTableKernel<<<(t.mask + 256) / 256, 256, 0, s>>>(t);
  return cudaGetLastError();
}

cudaError_t launchDeltaRehash(const DeltaTable& t, const DeltaRecords& r, const unsigned long long* counters,
                              uint32_t n_upper, unsigned long long* err, cudaStream_t s) {
  if (n_upper == 0) return cudaSuccess;
  rehashKernel<<<(n_upper + 255) / 256, 256, 0, s>>>(t, r, counters, n_upper, err);
  return cudaGetLastError();
}

cudaError_t launchDeltaAddBlock(const DeltaTable& t, const DeltaRecords& r, const DeltaBlockArgs& b,
                                uint32_t* tiles, unsigned long long* counters, cudaStream_t s) {
  const unsigned n_tiles = tilesFor(b.n);
  uint32_t* bases = tiles + 3 * n_tiles;  // 3 bases + room for the totals (as u64 x 3 behind them)
  unsigned long long* totals =
      reinterpret_cast<unsigned long long*>(tiles + ((3 * n_tiles + 4 + 1) & ~1u));  // 8-byte aligned
  if (b.n == 0) {
    cudaMemsetAsync(b.voxel_count, 0, sizeof(uint32_t), s);
    return cudaGetLastError();
  }
  addInsertKernel<<<(b.n + 255) / 256, 256, 0, s>>>(t, b, counters);
  addCountKernel<<<n_tiles, kDeltaTile, 0, s>>>(t, r, b, tiles);
  scanTiles3Kernel<<<1, 96, 0, s>>>(tiles, n_tiles, nullptr, counters + kDcNumRecords,
                                    counters + kDcActiveSlots, bases, totals);
  addApplyKernel<<<n_tiles, kDeltaTile, 0, s>>>(t, r, b, tiles, bases);
  cudaMemcpyAsync(b.voxel_count, totals, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
  addFaceMapKernel<<<(b.n + 255) / 256, 256, 0, s>>>(t, r, b);
  return cudaGetLastError();
}

cudaError_t launchDeltaRelease(const DeltaTable& t, const DeltaRecords& r, const DeltaOut& o,
                               const uint32_t* voxels, const uint32_t* voxel_count, uint32_t n_upper,
                               bool archive, uint8_t* flags, uint32_t* tiles,
                               unsigned long long* counters, cudaStream_t s) {
  if (n_upper == 0) return cudaSuccess;
  const unsigned n_tiles = tilesFor(n_upper);
  uint32_t* bases = tiles + 3 * n_tiles;
  releaseFlagKernel<<<n_tiles, kDeltaTile, 0, s>>>(t, r, o, voxels, voxel_count, archive, flags, tiles,
                                                  counters);
  scanTiles3Kernel<<<1, 96, 0, s>>>(tiles, n_tiles, counters + kDcArchVerts, nullptr, nullptr, bases,
                                    nullptr);
  releaseApplyKernel<<<n_tiles, kDeltaTile, 0, s>>>(t, r, o, voxels, flags, tiles, bases, counters);
  return cudaGetLastError();
}

cudaError_t launchDeltaActiveVertices(const DeltaRecords& r, const DeltaOut& o, uint32_t n_upper,
                                      uint32_t seq, uint32_t* active_remapping, uint32_t* tiles,
                                      unsigned long long* counters, cudaStream_t s) {
  if (n_upper == 0) {
    cudaMemsetAsync(counters + kDcScratch, 0, sizeof(unsigned long long), s);
    return cudaGetLastError();
  }
  const unsigned n_tiles = tilesFor(n_upper);
  uint32_t* bases = tiles + 3 * n_tiles;
  activeCountKernel<<<n_tiles, kDeltaTile, 0, s>>>(r, counters, tiles);
  // the number of active vertices goes to counters[kDcScratch]
  scanTiles3Kernel<<<1, 96, 0, s>>>(tiles, n_tiles, nullptr, nullptr, nullptr, bases,
                                    counters + kDcScratch);
  activeApplyKernel<<<n_tiles, kDeltaTile, 0, s>>>(r, o, seq, active_remapping, tiles, counters);
  return cudaGetLastError();
}

cudaError_t launchDeltaRemapFaces(const DeltaFaceBlock* blocks, uint32_t n_blocks, uint32_t n_faces,
                                  uint32_t* pool, const uint32_t* active_remapping, const DeltaOut& o,
                                  unsigned long long* remap_out, uint32_t* work_faces, uint8_t* valid,
                                  unsigned long long* counters, cudaStream_t s) {
  if (n_faces == 0) return cudaSuccess;
  remapFacesKernel<<<(n_faces + 255) / 256, 256, 0, s>>>(blocks, n_blocks, n_faces, pool,
                                                        active_remapping, o, remap_out, work_faces,
                                                        valid, counters);
  return cudaGetLastError();
}

cudaError_t launchDeltaEdgeRound(const DeltaEdgeTable& e, const uint32_t* faces, const uint8_t* valid,
                                 const uint8_t* acc_in, uint8_t* acc_out, uint32_t n_faces,
                                 unsigned long long* counters, cudaStream_t s) {
  if (n_faces == 0) return cudaSuccess;
  cudaMemsetAsync(e.keys, 0, (static_cast<size_t>(e.mask) + 1) * sizeof(unsigned long long), s);
  cudaMemsetAsync(e.minseq, 0xFF, (static_cast<size_t>(e.mask) + 1) * sizeof(uint32_t), s);
  cudaMemsetAsync(counters + kDcChanged, 0, sizeof(unsigned long long), s);
  edgeInsertKernel<<<(n_faces + 255) / 256, 256, 0, s>>>(e, faces, acc_in, n_faces);
  edgeDecideKernel<<<(n_faces + 255) / 256, 256, 0, s>>>(e, faces, valid, acc_in, acc_out, n_faces,
                                                        counters);
  return cudaGetLastError();
}

cudaError_t launchDeltaAppendFaces(const uint32_t* faces, const uint8_t* flags, uint32_t n_faces,
                                   uint32_t* out, unsigned long long* count, uint32_t* tiles,
                                   cudaStream_t s) {
  if (n_faces == 0) return cudaSuccess;
  const unsigned n_tiles = tilesFor(n_faces);
  uint32_t* bases = tiles + 3 * n_tiles;
  flagCountKernel<<<n_tiles, kDeltaTile, 0, s>>>(flags, n_faces, tiles);
  scanTiles3Kernel<<<1, 96, 0, s>>>(tiles, n_tiles, count, nullptr, nullptr, bases, nullptr);
  appendFacesKernel<<<n_tiles, kDeltaTile, 0, s>>>(faces, flags, n_faces, out, tiles, bases);
  return cudaGetLastError();
}

cudaError_t launchDeltaPendingRemap(uint32_t* faces, uint32_t n_faces, const DeltaOut& o,
                                    uint32_t num_archived_vertices, uint8_t* valid, cudaStream_t s) {
  if (n_faces == 0) return cudaSuccess;
  pendingRemapKernel<<<(n_faces + 255) / 256, 256, 0, s>>>(faces, n_faces, o, num_archived_vertices, valid);
  return cudaGetLastError();
}

cudaError_t launchDeltaClassifyPending(const uint32_t* faces, const uint8_t* acc, uint32_t n_faces,
                                       const unsigned long long* counters, const DeltaOut& o,
                                       uint8_t* to_arch, uint8_t* to_upd, cudaStream_t s) {
  if (n_faces == 0) return cudaSuccess;
  classifyPendingKernel<<<(n_faces + 255) / 256, 256, 0, s>>>(faces, acc, n_faces, counters, o, to_arch, to_upd);
  return cudaGetLastError();
}

cudaError_t launchDeltaArchiveFaces(const DeltaFaceBlock* blocks, uint32_t n_blocks, uint32_t n_faces,
                                    const uint32_t* pool, const DeltaOut& o,
                                    uint32_t num_archived_vertices, const unsigned long long* counters,
                                    uint32_t* work_faces, uint8_t* to_pending, uint8_t* candidate,
                                    cudaStream_t s) {
  if (n_faces == 0) return cudaSuccess;
  archiveFacesKernel<<<(n_faces + 255) / 256, 256, 0, s>>>(blocks, n_blocks, n_faces, pool, o,
                                                          num_archived_vertices, counters, work_faces,
                                                          to_pending, candidate);
  return cudaGetLastError();
}

cudaError_t launchDeltaCompactRecords(const DeltaRecords& from, const DeltaRecords& to, uint32_t n_old,
                                      uint32_t* new_id, uint32_t* voxel_pool, uint64_t pool_used,
                                      uint32_t* tiles, unsigned long long* counters, cudaStream_t s) {
  if (n_old == 0) return cudaSuccess;
  const unsigned n_tiles = tilesFor(n_old);
  uint32_t* bases = tiles + 3 * n_tiles;
  activeCountKernel<<<n_tiles, kDeltaTile, 0, s>>>(from, counters, tiles);
  scanTiles3Kernel<<<1, 96, 0, s>>>(tiles, n_tiles, nullptr, nullptr, nullptr, bases,
                                    counters + kDcScratch + 4);
  compactMoveKernel<<<n_tiles, kDeltaTile, 0, s>>>(from, to, tiles, counters, new_id);
  if (pool_used) {
    renumberVoxelsKernel<<<static_cast<unsigned>((pool_used + 255) / 256), 256, 0, s>>>(voxel_pool, pool_used,
                                                                                       new_id, n_old);
  }
  setCounterKernel<<<1, 32, 0, s>>>(counters + kDcNumRecords, counters + kDcScratch + 4);
  return cudaGetLastError();
}

cudaError_t launchDeltaSmallBlock(const DeltaTable& t, const DeltaRecords& r, const DeltaOut& o,
                                  const DeltaBlockArgs& b, const uint32_t* prev_voxels,
                                  const uint32_t* prev_count, unsigned long long* counters,
                                  cudaStream_t s) {
  smallBlockKernel<<<1, kDeltaTile, 0, s>>>(t, r, o, b, prev_voxels, prev_count, counters);
  return cudaGetLastError();
}

cudaError_t launchDeltaCopySegments(const uint4* segs, uint32_t n_segs, const uint32_t* src_a,
                                    uint32_t* dst_a, const uint32_t* src_b, uint32_t* dst_b,
                                    cudaStream_t s) {
  if (n_segs == 0) return cudaSuccess;
  copySegmentsKernel<<<n_segs, 256, 0, s>>>(segs, src_a, dst_a, src_b, dst_b);
  return cudaGetLastError();
}

}  // namespace pgmo

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadDeltaKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, scanTiles3Kernel);
  cudaFuncGetAttributes(&attr, initTableKernel);
  cudaFuncGetAttributes(&attr, rehashKernel);
  cudaFuncGetAttributes(&attr, addInsertKernel);
  cudaFuncGetAttributes(&attr, addCountKernel);
  cudaFuncGetAttributes(&attr, addApplyKernel);
  cudaFuncGetAttributes(&attr, addFaceMapKernel);
  cudaFuncGetAttributes(&attr, releaseFlagKernel);
  cudaFuncGetAttributes(&attr, releaseApplyKernel);
  cudaFuncGetAttributes(&attr, activeCountKernel);
  cudaFuncGetAttributes(&attr, activeApplyKernel);
  cudaFuncGetAttributes(&attr, remapFacesKernel);
  cudaFuncGetAttributes(&attr, edgeInsertKernel);
  cudaFuncGetAttributes(&attr, edgeDecideKernel);
  cudaFuncGetAttributes(&attr, flagCountKernel);
  cudaFuncGetAttributes(&attr, appendFacesKernel);
  cudaFuncGetAttributes(&attr, pendingRemapKernel);
  cudaFuncGetAttributes(&attr, classifyPendingKernel);
  cudaFuncGetAttributes(&attr, archiveFacesKernel);
  cudaFuncGetAttributes(&attr, smallBlockKernel);
  cudaFuncGetAttributes(&attr, compactMoveKernel);
  cudaFuncGetAttributes(&attr, renumberVoxelsKernel);
  cudaFuncGetAttributes(&attr, setCounterKernel);
  cudaFuncGetAttributes(&attr, copySegmentsKernel);
  cudaGetLastError();
}
}  // namespace pgmo
