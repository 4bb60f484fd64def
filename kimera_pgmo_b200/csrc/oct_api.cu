// OctreeCompression C ABI (see include/pgmo_b200.h) — MeshCompression::compressAndIntegrate /
// pruneStoredMesh (kimera_pgmo/src/compression/mesh_compression.cpp:216-464) with
// OctreeCompression's structure tests (src/compression/octree_compression.cpp:17-81).
// The octree itself lives in oct_kernels.cu; this file holds the mesh bookkeeping kernels
// (face checks, ordered acceptance, remap record, polygon set) and the per-message host flow.
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "deform_schedule.h"  // stampFromSec
#include "oct_kernels.cuh"
#include "mesh_interface.cuh"
#include "vcc_kernels.cuh"    // launchVccCompactLive / vccTileScratchEntries (ordered compaction)

using namespace pgmo;

namespace {

struct PolyTable {
  unsigned long long* keys;  // 64-bit mix of the ordered triple; kOctEmptyKey = empty
  uint32_t* tri;             // 3 per slot: the triple itself (collision check)
  uint32_t* val;             // 0 = committed, i + 1 = pending for input vertex i, ~0 = fresh
  uint32_t mask;
};

__device__ __forceinline__ unsigned long long mixTriple(uint32_t a, uint32_t b, uint32_t c) {
  unsigned long long k = (static_cast<unsigned long long>(a) << 32) | b;
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k += static_cast<unsigned long long>(c) * 0x9e3779b97f4a7c15ull;
  k ^= k >> 29;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 32;
  return k == kOctEmptyKey ? 0ull : k;
}

__device__ __forceinline__ uint32_t polyInsert(const PolyTable& t, unsigned long long key) {
  uint32_t slot = static_cast<uint32_t>(key) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    const unsigned long long old = atomicCAS(t.keys + slot, kOctEmptyKey, key);
    if (old == kOctEmptyKey || old == key) return slot;
    slot = (slot + 1) & t.mask;
  }
  return kOctNone;
}

// i % 3 == 2 inside its block (mesh_compression.cpp:304)
__global__ void octTriEndKernel(const unsigned long long* __restrict__ offsets, uint32_t n_blocks,
                                uint32_t n, uint8_t* __restrict__ tri_end) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t lo = 0, hi = n_blocks;  // last b with offsets[b] <= i
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= i) lo = mid;
    else hi = mid;
  }
  // empty blocks share offsets: the last such b is the one that owns i
  tri_end[i] = ((i - offsets[lo]) % 3ull) == 2ull ? 1 : 0;
}

__global__ void octCandFlagsKernel(const int32_t* __restrict__ reobs, uint32_t n,
                                   uint8_t* __restrict__ flag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = reobs[i] < 0 ? 1 : 0;
}

__global__ void octGatherCandKernel(const unsigned long long* __restrict__ cand_list, uint32_t n_cand,
                                    const float* __restrict__ xyz, float4* __restrict__ cand_xyz,
                                    int32_t* __restrict__ cand_of_vertex) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const uint32_t v = static_cast<uint32_t>(cand_list[c]);
  cand_xyz[c] = make_float4(xyz[3 * v], xyz[3 * v + 1], xyz[3 * v + 2], 0.f);
  cand_of_vertex[v] = static_cast<int32_t>(c);
}

__global__ void octScatterRankKernel(const unsigned long long* __restrict__ list, uint32_t n,
                                     int32_t* __restrict__ rank_of) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) rank_of[list[j]] = static_cast<int32_t>(j);
}

// temp_reindex (mesh_compression.cpp:274-292): old global index for re-observations,
// num_original + temp index for temp-unique / converged vertices
__global__ void octTempReindexKernel(uint32_t n, const int32_t* __restrict__ reobs,
                                     const uint32_t* __restrict__ act_index,
                                     const int32_t* __restrict__ cand_of_vertex,
                                     const int32_t* __restrict__ decision,
                                     const int32_t* __restrict__ pot_of_cand,
                                     unsigned long long n_orig, long long* __restrict__ tr) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = reobs[i];
  if (r >= 0) {
    tr[i] = static_cast<long long>(act_index[r]);
    return;
  }
  const int32_t c = cand_of_vertex[i];
  const int32_t target = decision[c] < 0 ? c : decision[c];
  tr[i] = static_cast<long long>(n_orig) + pot_of_cand[target];
}

// the face check of the vertex loop (:304-343)
__global__ void octFaceCheckKernel(uint32_t n, const uint8_t* __restrict__ tri_end,
                                   const long long* __restrict__ tr, unsigned long long n_orig,
                                   uint8_t* __restrict__ surf, uint8_t* __restrict__ check) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !tri_end[i]) return;
  const long long r0 = tr[i - 2], r1 = tr[i - 1], r2 = tr[i];
  if (r2 == r1) return;
  const long long no = static_cast<long long>(n_orig);
  if (!(r0 >= no || r1 >= no || r2 >= no)) return;
  if (r0 == r1 || r0 == r2 || r1 == r2) return;
  surf[i] = 1;
  if (r0 >= no) check[r0 - no] = 1;
  if (r1 >= no) check[r1 - no] = 1;
  if (r2 >= no) check[r2 - no] = 1;
}

// accepted potential vertices, in creation order (:349-378)
__global__ void octCommitVerticesKernel(uint32_t n_acc, const unsigned long long* __restrict__ acc_list,
                                        const unsigned long long* __restrict__ pot_list,
                                        const unsigned long long* __restrict__ cand_list,
                                        const float* __restrict__ xyz, const uint32_t* __restrict__ rgba,
                                        unsigned long long n_orig, unsigned long long n_act,
                                        unsigned long long stamp_ns, double stamp_s,
                                        float4* __restrict__ all_vertices,
                                        unsigned long long* __restrict__ all_stamps,
                                        float4* __restrict__ act_xyz, uint32_t* __restrict__ act_index,
                                        double* __restrict__ act_stamp, float4* __restrict__ new_vertices) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_acc) return;
  const uint32_t v = static_cast<uint32_t>(cand_list[pot_list[acc_list[a]]]);
  const float4 p = make_float4(xyz[3 * v], xyz[3 * v + 1], xyz[3 * v + 2],
                               __uint_as_float(rgba ? rgba[v] : 0u));
  const unsigned long long g = n_orig + a;
  all_vertices[g] = p;
  all_stamps[g] = stamp_ns;
  act_xyz[n_act + a] = make_float4(p.x, p.y, p.z, 0.f);
  act_index[n_act + a] = static_cast<uint32_t>(g);
  act_stamp[n_act + a] = stamp_s;
  new_vertices[a] = p;
}

// reindex + the dense record of what the reference does to `remapping` (:284-292, :363-373)
__global__ void octReindexKernel(uint32_t n, const int32_t* __restrict__ reobs,
                                 const uint32_t* __restrict__ act_index,
                                 const int32_t* __restrict__ cand_of_vertex,
                                 const int32_t* __restrict__ decision,
                                 const int32_t* __restrict__ pot_of_cand,
                                 const int32_t* __restrict__ acc_rank, unsigned long long n_orig,
                                 double stamp_s, double* __restrict__ act_stamp,
                                 uint8_t* __restrict__ reobs_flag, long long* __restrict__ reindex,
                                 long long* __restrict__ remap_index, uint8_t* __restrict__ remap_kind) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = reobs[i];
  if (r >= 0) {
    const uint32_t g = act_index[r];
    act_stamp[r] = stamp_s;
    reobs_flag[g] = 1;
    reindex[i] = g;
    remap_index[i] = g;
    remap_kind[i] = 1;
    return;
  }
  const int32_t c = cand_of_vertex[i];
  const int32_t target = decision[c] < 0 ? c : decision[c];
  const int32_t rank = acc_rank[pot_of_cand[target]];
  if (rank >= 0) {
    const long long g = static_cast<long long>(n_orig) + rank;
    reindex[i] = g;
    remap_index[i] = g;
    remap_kind[i] = 2;
  } else {
    reindex[i] = -1;
    remap_index[i] = -1;
    remap_kind[i] = 0;
  }
}

__global__ void octAppendNewIndicesKernel(uint32_t n_acc, const unsigned long long* __restrict__ n_reobs,
                                          unsigned long long n_orig,
                                          unsigned long long* __restrict__ new_indices) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n_acc) new_indices[*n_reobs + a] = n_orig + a;
}

// second pass over the candidate faces (:382-416), part 1: classify
__global__ void octFaceClassifyKernel(uint32_t n, const uint8_t* __restrict__ surf,
                                      const long long* __restrict__ reindex, unsigned long long n_orig,
                                      PolyTable pt, uint8_t* __restrict__ keep,
                                      uint32_t* __restrict__ face_slot, uint32_t* __restrict__ error) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !surf[i]) return;
  const long long r0 = reindex[i - 2], r1 = reindex[i - 1], r2 = reindex[i];
  if (r0 < 0 || r1 < 0 || r2 < 0) return;
  if (r0 == r1 || r1 == r2 || r2 == r0) return;  // never stored, never changes any state
  const long long no = static_cast<long long>(n_orig);
  const uint32_t a = static_cast<uint32_t>(r0), b = static_cast<uint32_t>(r1),
                 c = static_cast<uint32_t>(r2);
  if (r0 >= no || r1 >= no || r2 >= no) {
    keep[i] = 1;  // a face with a new vertex is always added
    return;
  }
  // all-old face: added iff SurfaceExists is false, i.e. the ordered triple is not stored yet;
  // among equal faces of this message the first one wins
  const unsigned long long key = mixTriple(a, b, c);
  const uint32_t slot = polyInsert(pt, key);
  if (slot == kOctNone) {
    atomicOr(error, 4u);
    return;
  }
  atomicMin(pt.val + slot, i + 1);
  pt.tri[3 * slot] = a;
  pt.tri[3 * slot + 1] = b;
  pt.tri[3 * slot + 2] = c;
  face_slot[i] = slot;
  keep[i] = 2;  // pending
}

__global__ void octFaceFinalizeKernel(uint32_t n, const long long* __restrict__ reindex, PolyTable pt,
                                      uint8_t* __restrict__ keep, const uint32_t* __restrict__ face_slot,
                                      uint32_t* __restrict__ error) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || keep[i] != 2) return;
  const uint32_t slot = face_slot[i];
  const uint32_t a = static_cast<uint32_t>(reindex[i - 2]), b = static_cast<uint32_t>(reindex[i - 1]),
                 c = static_cast<uint32_t>(reindex[i]);
  if (pt.tri[3 * slot] != a || pt.tri[3 * slot + 1] != b || pt.tri[3 * slot + 2] != c) {
    atomicOr(error, 8u);  // 64-bit hash collision between different triples: fail loudly
    keep[i] = 0;
    return;
  }
  keep[i] = pt.val[slot] == i + 1 ? 1 : 0;
}

__global__ void octCommitFacesKernel(uint32_t n_faces, const unsigned long long* __restrict__ face_list,
                                     const long long* __restrict__ reindex, unsigned long long n_poly,
                                     PolyTable pt, uint32_t* __restrict__ polygons,
                                     uint32_t* __restrict__ new_triangles, uint32_t* __restrict__ error) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_faces) return;
  const uint32_t i = static_cast<uint32_t>(face_list[j]);
  const uint32_t a = static_cast<uint32_t>(reindex[i - 2]), b = static_cast<uint32_t>(reindex[i - 1]),
                 c = static_cast<uint32_t>(reindex[i]);
  polygons[3 * (n_poly + j)] = a;
  polygons[3 * (n_poly + j) + 1] = b;
  polygons[3 * (n_poly + j) + 2] = c;
  new_triangles[3 * j] = a;
  new_triangles[3 * j + 1] = b;
  new_triangles[3 * j + 2] = c;
  const unsigned long long key = mixTriple(a, b, c);
  const uint32_t slot = polyInsert(pt, key);
  if (slot == kOctNone) {
    atomicOr(error, 4u);
    return;
  }
  pt.tri[3 * slot] = a;
  pt.tri[3 * slot + 1] = b;
  pt.tri[3 * slot + 2] = c;
  pt.val[slot] = 0;  // committed
}

__global__ void octPolyRehashKernel(const uint32_t* __restrict__ polygons, unsigned long long n_poly,
                                    PolyTable pt, uint32_t* __restrict__ error) {
  const unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_poly) return;
  const uint32_t a = polygons[3 * j], b = polygons[3 * j + 1], c = polygons[3 * j + 2];
  const uint32_t slot = polyInsert(pt, mixTriple(a, b, c));
  if (slot == kOctNone) {
    atomicOr(error, 4u);
    return;
  }
  pt.tri[3 * slot] = a;
  pt.tri[3 * slot + 1] = b;
  pt.tri[3 * slot + 2] = c;
  pt.val[slot] = 0;
}

__global__ void octPruneFlagsKernel(const double* __restrict__ act_stamp, uint32_t n, double earliest,
                                    uint8_t* __restrict__ keep) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = act_stamp[i] > earliest ? 1 : 0;
}

__global__ void octPruneGatherKernel(const unsigned long long* __restrict__ list, uint32_t n,
                                     const float4* __restrict__ xyz, const uint32_t* __restrict__ index,
                                     const double* __restrict__ stamp, float4* __restrict__ xyz_out,
                                     uint32_t* __restrict__ index_out, double* __restrict__ stamp_out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long i = list[j];
  xyz_out[j] = xyz[i];
  index_out[j] = index[i];
  stamp_out[j] = stamp[i];
}

inline unsigned blocksFor(uint64_t n, unsigned threads = 256) {
  return static_cast<unsigned>((n + threads - 1) / threads);
}

uint64_t nextPow2(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct TreeStore {
  // keys (u64), minseq (u32) and head (u32) share one allocation so that emptying the table is
  // one memset: slab[0, cap) keys, then cap minseq words, then cap head words
  DevBuf<unsigned long long> slab;
  DevBuf<OctFrames> frames;
  uint64_t cap = 0;
  uint64_t factor = 8;  // table slots per point; doubled when a build reports "table full"
  uint64_t min_alloc_slots = 1024;  // floor of the allocation (not of the hashed range)
  OctTable table() const {
    OctTable t;
    t.keys = slab.ptr;
    t.minseq = reinterpret_cast<uint32_t*>(slab.ptr + cap);
    t.head = t.minseq + cap;
    t.mask = static_cast<uint32_t>(cap - 1);
    return t;
  }
  void releaseTable() {
    slab.release();
    cap = 0;
  }
  // empties the table (and optionally the frames) with room for `points` points
  int reset(uint64_t points, bool reset_frames, cudaStream_t st) {
    const uint64_t want = std::max<uint64_t>(nextPow2(factor * std::max<uint64_t>(points, 1)), 1024);
    if (want > cap) {
      // the table grows with the active set over the first messages of a stream, and every
      // cudaFree + cudaMalloc pair costs milliseconds here: allocate for one or two more
      // doublings (at least min_alloc_slots), so `cap` - the hashed and memset range - can grow
      // in place
      if (2 * want > slab.cap) {
        slab.release();
        const uint64_t room = want <= (1ull << 24) ? 4 * want : 2 * want;  // 16 B per slot
        PGMO_TRY(slab.reserve(2 * std::max<uint64_t>(room, min_alloc_slots)));
      }
      cap = want;
    }
    PGMO_TRY(frames.reserve(1));
    PGMO_CUDA(cudaMemsetAsync(slab.ptr, 0xFF, 2 * cap * sizeof(unsigned long long), st));
    if (reset_frames) PGMO_CUDA(cudaMemsetAsync(frames.ptr, 0, sizeof(OctFrames), st));
    return PGMO_OK;
  }
};

}  // namespace

struct pgmo_oct {
  // A failure reported after the stored mesh started to change (table full, depth / range limits,
  // CUDA errors inside the commit phase) leaves stamps refreshed and faces half inserted: the
  // handle refuses further work instead of continuing on that state.
  bool committing = false;  // the current call has started to change persistent state
  bool poisoned = false;
  int device = 0;
  int mode = 0;  // 0 = OctreeCompression, 1 = BlockCompression (cell hash)
  double resolution = 0.0;
  float voxel_size_inv = 1.f;
  cudaStream_t stream = nullptr;

  // MeshCompression state (mesh_compression.h:133-147)
  DevBuf<float4> all_vertices;             // xyz + rgba bits
  DevBuf<unsigned long long> all_stamps;
  uint64_t n_all = 0;
  DevBuf<float4> act_xyz, act_xyz2;        // active_vertices_xyz_
  DevBuf<uint32_t> act_index, act_index2;  // active_vertices_index_
  DevBuf<double> act_stamp, act_stamp2;    // active_vertex_stamps_
  DevBuf<uint32_t> act_next;
  DevBuf<int4> act_sigma;
  uint64_t n_act = 0;
  DevBuf<uint32_t> polygons;               // polygons_
  uint64_t n_poly = 0;
  DevBuf<unsigned long long> pt_keys;      // set of stored ordered triples (SurfaceExists)
  DevBuf<uint32_t> pt_tri, pt_val;
  uint64_t pt_cap = 0;

  TreeStore ptree;  // octree_
  bool ptree_built = false;
  TreeStore ttree;  // temp_octree_

  // last call's outputs
  DevBuf<float4> new_vertices;
  DevBuf<uint32_t> new_triangles;
  DevBuf<unsigned long long> new_indices;
  uint64_t n_new_vertices = 0, n_new_triangles = 0, n_new_indices = 0;

  // per-message scratch
  DevBuf<float> d_xyz;
  DevBuf<uint32_t> d_rgba;
  DevBuf<unsigned long long> d_offsets;
  VoxbloxStage vb_stage;
  DevBuf<uint8_t> tri_end, flag, member_a, member_b, surf, check, keep, reobs_flag;
  DevBuf<int32_t> reobs, cand_of_vertex, decision, pot_of_cand, acc_rank;
  DevBuf<unsigned long long> cand_list, pot_list, acc_list, face_list;
  DevBuf<float4> cand_xyz;
  DevBuf<uint32_t> cand_next, face_slot, tiles;
  DevBuf<int4> cand_sigma;
  DevBuf<long long> tr, reindex, d_remap_index;
  DevBuf<uint8_t> d_remap_kind;
  DevBuf<unsigned long long> counts;  // device counters
  DevBuf<uint32_t> flags32;           // [0] changed, [1] error
  PinnedBuf<unsigned long long> h_counts;
  PinnedBuf<uint32_t> h_flags;
  PinnedBuf<OctFrames> h_frames;
  uint64_t last_rounds = 0, last_cand = 0, last_pot = 0;
};

namespace {

PolyTable polyTable(pgmo_oct_t* h) {
  PolyTable t;
  t.keys = h->pt_keys.ptr;
  t.tri = h->pt_tri.ptr;
  t.val = h->pt_val.ptr;
  t.mask = static_cast<uint32_t>(h->pt_cap - 1);
  return t;
}

int readCount(pgmo_oct_t* h, int which, cudaStream_t st, uint64_t* out) {
  PGMO_TRY(h->h_counts.reserve(8));
  PGMO_CUDA(cudaMemcpyAsync(h->h_counts.ptr, h->counts.ptr + which, sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  *out = h->h_counts.ptr[0];
  return PGMO_OK;
}

// Grow rarely, by a lot: an allocation costs milliseconds (sometimes hundreds) inside the message
// stream, and buffers sized by the whole mesh would otherwise regrow all along the stream.
inline uint64_t roomy(uint64_t need, size_t cap) {
  return need <= cap ? need : std::max<uint64_t>(2 * need, 1ull << 21);
}

int compact(pgmo_oct_t* h, const uint8_t* flags, uint64_t n, DevBuf<unsigned long long>* list,
            int counter, cudaStream_t st, uint64_t* count) {
  PGMO_TRY(list->reserve(roomy(n + 1, list->cap)));
  PGMO_TRY(h->tiles.reserve(roomy(vccTileScratchEntries(n) + 1, h->tiles.cap)));
  PGMO_CUDA(launchVccCompactLive(flags, n, reinterpret_cast<uint64_t*>(list->ptr), h->tiles.ptr,
                                 h->counts.ptr + counter, st));
  return readCount(h, counter, st, count);
}

int checkTreeError(pgmo_oct_t* h, TreeStore* tree, cudaStream_t st, uint32_t* error) {
  PGMO_TRY(h->h_frames.reserve(1));
  PGMO_CUDA(cudaMemcpyAsync(h->h_frames.ptr, tree->frames.ptr, sizeof(OctFrames),
                            cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  *error = h->h_frames.ptr->error;
  return PGMO_OK;
}

int treeErrorToRc(uint32_t error) {
  if (error & 1u) {
    setLastError("octree: bounding box grew past the supported depth");
    return PGMO_ERR_UNSUPPORTED;
  }
  if (error & 2u) {
    setLastError("octree: vertex farther than 2^18 voxels from the first vertex (block mode: cell "
                 "index beyond +-2^20)");
    return PGMO_ERR_UNSUPPORTED;
  }
  if (error & 8u) {
    setLastError("polygon set: 64-bit hash collision");
    return PGMO_ERR_UNSUPPORTED;
  }
  if (error & 4u) {
    setLastError("octree / polygon table full");
    return PGMO_ERR_CUDA;
  }
  return PGMO_OK;
}

int buildPersistentOnce(pgmo_oct_t* h, uint64_t first, cudaStream_t st, uint32_t* error);

// (re)builds the persistent structure over active points [first, n_act): first == 0 re-anchors
// (OctreeCompression::reInitializeStructure), first > 0 is updateStructure for the new points.
int buildPersistent(pgmo_oct_t* h, uint64_t first, cudaStream_t st) {
  while (true) {
    uint32_t error = 0;
    PGMO_TRY(buildPersistentOnce(h, first, st, &error));
    if (error == 4u && h->ptree.factor < 256) {
      // table full: re-register everybody with twice the room.  The frames already hold every
      // growth event, so this neither re-anchors nor changes any key.
      h->ptree.factor *= 2;
      h->ptree.releaseTable();
      if (first == 0) h->ptree_built = false;
      else first = 1;  // forces the "re-register all" branch below without re-anchoring
      continue;
    }
    return treeErrorToRc(error);
  }
}

int buildPersistentOnce(pgmo_oct_t* h, uint64_t first, cudaStream_t st, uint32_t* error_out) {
  const uint64_t n = h->n_act;
  if (h->mode == 1) {
    // cell hash: reInitializeStructure = emplace everybody in index order, updateStructure =
    // emplace the new points; `emplace` keeps the first point of a cell = the smallest index
    if (first == 0 || !h->ptree_built || h->ptree.factor * n > h->ptree.cap) {
      PGMO_TRY(h->ptree.reset(2 * n, true, st));
      h->ptree_built = true;
      first = 0;
    }
    PGMO_CUDA(launchCellInsert(h->ptree.frames.ptr, h->ptree.table(), h->act_xyz.ptr,
                               static_cast<uint32_t>(first), static_cast<uint32_t>(n),
                               h->voxel_size_inv, st));
    PGMO_TRY(checkTreeError(h, &h->ptree, st, error_out));
    if (*error_out == 4u) {
      PGMO_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(h->ptree.frames.ptr) + offsetof(OctFrames, error), 0,
                                sizeof(uint32_t), st));
    }
    return PGMO_OK;
  }
  PGMO_TRY(h->act_next.reserve(h->act_xyz.cap));
  PGMO_TRY(h->act_sigma.reserve(h->act_xyz.cap));
  const bool full = first == 0 || !h->ptree_built || h->ptree.factor * n > h->ptree.cap;
  OctPoints pts;
  pts.xyz = h->act_xyz.ptr;
  pts.next = h->act_next.ptr;
  pts.sigma = h->act_sigma.ptr;
  if (first == 0 || !h->ptree_built) {
    PGMO_TRY(h->ptree.reset(2 * n, true, st));
    h->ptree_built = true;
    if (n == 0) return PGMO_OK;
    PGMO_CUDA(launchOctGrow(h->ptree.frames.ptr, h->act_xyz.ptr, nullptr, 0, static_cast<uint32_t>(n),
                            h->resolution, st));
    PGMO_CUDA(launchOctInsert(h->ptree.frames.ptr, h->ptree.table(), pts, nullptr, 0,
                              static_cast<uint32_t>(n), h->resolution, st));
  } else {
    // growth for the new points (in order), then either insert them or, when the table is
    // getting crowded, re-register everybody (keeping the frames: no re-anchoring)
    PGMO_CUDA(launchOctGrow(h->ptree.frames.ptr, h->act_xyz.ptr, nullptr, static_cast<uint32_t>(first),
                            static_cast<uint32_t>(n), h->resolution, st));
    if (full) {
      PGMO_TRY(h->ptree.reset(2 * n, false, st));
      PGMO_CUDA(launchOctInsert(h->ptree.frames.ptr, h->ptree.table(), pts, nullptr, 0,
                                static_cast<uint32_t>(n), h->resolution, st));
    } else {
      PGMO_CUDA(launchOctInsert(h->ptree.frames.ptr, h->ptree.table(), pts, nullptr,
                                static_cast<uint32_t>(first), static_cast<uint32_t>(n), h->resolution, st));
    }
  }
  PGMO_CUDA(launchOctInsertRootChain(h->ptree.frames.ptr, h->ptree.table(), st));
  PGMO_TRY(checkTreeError(h, &h->ptree, st, error_out));
  if (*error_out == 4u) {
    // clear the sticky flag for the retry
    PGMO_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(h->ptree.frames.ptr) + offsetof(OctFrames, error), 0,
                              sizeof(uint32_t), st));
  }
  return PGMO_OK;
}

int ensurePolyTable(pgmo_oct_t* h, uint64_t extra, cudaStream_t st) {
  const uint64_t want = std::max<uint64_t>(nextPow2(2 * (h->n_poly + extra) + 16), 1024);
  if (want <= h->pt_cap) return PGMO_OK;
  const uint64_t cap = std::max<uint64_t>(want * 2, 1ull << 21);  // grow with head room
  h->pt_keys.release();
  h->pt_tri.release();
  h->pt_val.release();
  PGMO_TRY(h->pt_keys.reserve(cap));
  PGMO_TRY(h->pt_tri.reserve(3 * cap));
  PGMO_TRY(h->pt_val.reserve(cap));
  h->pt_cap = cap;
  PGMO_CUDA(cudaMemsetAsync(h->pt_keys.ptr, 0xFF, cap * sizeof(unsigned long long), st));
  PGMO_CUDA(cudaMemsetAsync(h->pt_val.ptr, 0xFF, cap * sizeof(uint32_t), st));
  if (h->n_poly > 0) {
    octPolyRehashKernel<<<blocksFor(h->n_poly), 256, 0, st>>>(h->polygons.ptr, h->n_poly, polyTable(h),
                                                             h->flags32.ptr + 1);
    PGMO_CUDA(cudaGetLastError());
  }
  return PGMO_OK;
}

}  // namespace

extern "C" {

int pgmo_oct_create(double resolution, int device, pgmo_oct_t** out) {
  PGMO_CHECK_ARG(out != nullptr);
  *out = nullptr;
  PGMO_CHECK_ARG(resolution > 0.0);
  PGMO_TRY(ensureDevice(device));
  pgmo_oct_t* h = new (std::nothrow) pgmo_oct();
  PGMO_CHECK_ARG(h != nullptr);
  h->device = device;
  h->resolution = resolution;
  // persistent structure: 64 MB of table up front (4 M slots) covers an active set of
  // 250 k vertices, the steady state of a 50 k-vertex / 10 s-horizon stream, without a regrowth
  h->ptree.min_alloc_slots = 1ull << 22;
  h->ttree.min_alloc_slots = 1ull << 19;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    setLastError("cudaStreamCreateWithFlags failed");
    return PGMO_ERR_CUDA;
  }
  if (h->counts.reserve(16) < 0 || h->flags32.reserve(4) < 0) {
    pgmo_oct_destroy(h);
    return PGMO_ERR_CUDA;
  }
  cudaMemsetAsync(h->counts.ptr, 0, 16 * sizeof(unsigned long long), h->stream);
  cudaMemsetAsync(h->flags32.ptr, 0, 4 * sizeof(uint32_t), h->stream);
  cudaStreamSynchronize(h->stream);
  preloadOctTreeKernels();
  preloadOctApiKernels();
  preloadMeshInterfaceKernels();
  preloadCommonKernels();
  *out = h;
  return PGMO_OK;
}

int pgmo_oct_create_block(double resolution, int device, pgmo_oct_t** out) {
  PGMO_TRY(pgmo_oct_create(resolution, device, out));
  (*out)->mode = 1;
  (*out)->voxel_size_inv = 1.0f / static_cast<float>(resolution);
  return PGMO_OK;
}

void pgmo_oct_destroy(pgmo_oct_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  delete h;
}

static int octIntegrate(pgmo_oct_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                        const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                        const VoxbloxInput* vb, double stamp_s, int64_t* remap_index,
                        uint8_t* remap_kind) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_CHECK_ARG(n_blocks == 0 || (block_idx_xyz && block_offsets));
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  h->n_new_vertices = h->n_new_triangles = h->n_new_indices = 0;
  const uint64_t n64 = n_blocks ? block_offsets[n_blocks] : 0;
  if (n64 == 0) return PGMO_OK;
  PGMO_CHECK_ARG(n64 < 0x7FFFFFFFull);
  PGMO_CHECK_ARG((xyz || vb) && remap_index && remap_kind);
  const uint32_t n = static_cast<uint32_t>(n64);
  const double res = h->resolution;
  const uint64_t n_orig = h->n_all;
  // A block index twice in one message: the reference resolves both occurrences to one block
  // through its block map; the per-block bookkeeping here assumes unique blocks (as
  // pgmo_vcc_integrate and pgmo_delta_update do).
  for (uint32_t b = 0; b < n_blocks; ++b) {
    for (uint32_t c = 0; c < b; ++c) {
      if (std::memcmp(block_idx_xyz + 3 * b, block_idx_xyz + 3 * c, 3 * sizeof(int32_t)) == 0) {
        setLastError("octree compression: a block index appears twice in one message");
        return PGMO_ERR_UNSUPPORTED;
      }
    }
  }

  // ---- inputs
  PGMO_TRY(h->d_xyz.reserve(3ull * n));
  PGMO_TRY(h->d_rgba.reserve(n));
  PGMO_TRY(h->d_offsets.reserve(n_blocks + 1));
  if (vb) {
    PGMO_TRY(stageVoxblox(&h->vb_stage, n_blocks, block_idx_xyz, block_offsets, *vb, h->d_xyz.ptr,
                          h->d_rgba.ptr, st));
  } else {
    PGMO_CUDA(cudaMemcpyAsync(h->d_xyz.ptr, xyz, 3ull * n * sizeof(float), cudaMemcpyHostToDevice, st));
    if (rgba) PGMO_CUDA(cudaMemcpyAsync(h->d_rgba.ptr, rgba, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  }
  PGMO_CUDA(cudaMemcpyAsync(h->d_offsets.ptr, block_offsets, (n_blocks + 1) * sizeof(uint64_t),
                            cudaMemcpyHostToDevice, st));
  PGMO_TRY(h->tri_end.reserve(n));
  PGMO_TRY(h->flag.reserve(n));
  PGMO_TRY(h->reobs.reserve(n));
  PGMO_TRY(h->cand_of_vertex.reserve(n));
  PGMO_TRY(h->surf.reserve(n));
  PGMO_TRY(h->keep.reserve(n));
  PGMO_TRY(h->face_slot.reserve(n));
  PGMO_TRY(h->tr.reserve(n));
  PGMO_TRY(h->reindex.reserve(n));
  PGMO_TRY(h->d_remap_index.reserve(n));
  PGMO_TRY(h->d_remap_kind.reserve(n));
  PGMO_TRY(h->h_flags.reserve(4));
  PGMO_CUDA(cudaMemsetAsync(h->flags32.ptr, 0, 4 * sizeof(uint32_t), st));
  octTriEndKernel<<<blocksFor(n), 256, 0, st>>>(h->d_offsets.ptr, n_blocks, n, h->tri_end.ptr);

  // ---- persistent-structure test for every vertex (:268)
  if (!h->ptree_built) PGMO_TRY(buildPersistent(h, 0, st));
  if (h->mode == 1) {
    PGMO_CUDA(launchCellQueryPersistent(h->ptree.table(), h->d_xyz.ptr, n, h->voxel_size_inv,
                                        h->reobs.ptr, st));
  } else {
    PGMO_CUDA(launchOctQueryPersistent(h->ptree.frames.ptr, h->ptree.table(), h->act_xyz.ptr,
                                       h->act_next.ptr, h->d_xyz.ptr, n, res, h->reobs.ptr, st));
  }
  octCandFlagsKernel<<<blocksFor(n), 256, 0, st>>>(h->reobs.ptr, n, h->flag.ptr);
  uint64_t n_cand = 0;
  PGMO_TRY(compact(h, h->flag.ptr, n, &h->cand_list, 0, st, &n_cand));

  // ---- temp-structure greedy (:269-283) as a fixed point over "who becomes a temp vertex"
  uint64_t n_pot = 0;
  PGMO_TRY(h->cand_xyz.reserve(n_cand + 1));
  PGMO_TRY(h->cand_next.reserve(n_cand + 1));
  PGMO_TRY(h->cand_sigma.reserve(n_cand + 1));
  PGMO_TRY(h->member_a.reserve(n_cand + 1));
  PGMO_TRY(h->member_b.reserve(n_cand + 1));
  PGMO_TRY(h->decision.reserve(n_cand + 1));
  PGMO_TRY(h->pot_of_cand.reserve(n_cand + 1));
  PGMO_CUDA(cudaMemsetAsync(h->cand_of_vertex.ptr, 0xFF, n * sizeof(int32_t), st));
  uint8_t* member = h->member_a.ptr;
  uint8_t* member_next = h->member_b.ptr;
  h->last_rounds = 0;
  if (n_cand > 0) {
    const uint32_t nc = static_cast<uint32_t>(n_cand);
    octGatherCandKernel<<<blocksFor(nc), 256, 0, st>>>(h->cand_list.ptr, nc, h->d_xyz.ptr,
                                                       h->cand_xyz.ptr, h->cand_of_vertex.ptr);
    PGMO_CUDA(cudaMemsetAsync(member, 1, nc, st));
    OctPoints cpts;
    cpts.xyz = h->cand_xyz.ptr;
    cpts.next = h->cand_next.ptr;
    cpts.sigma = h->cand_sigma.ptr;
    while (h->mode == 1) {
      // BlockCompression: the first candidate of a cell is the temp vertex, no iteration needed
      h->last_rounds = 1;
      PGMO_TRY(h->ttree.reset(n_cand, true, st));
      PGMO_CUDA(launchCellInsert(h->ttree.frames.ptr, h->ttree.table(), h->cand_xyz.ptr, 0, nc,
                                 h->voxel_size_inv, st));
      PGMO_CUDA(launchCellDecideTemp(h->ttree.table(), h->cand_xyz.ptr, nc, h->voxel_size_inv, member,
                                     h->decision.ptr, st));
      uint32_t error = 0;
      PGMO_TRY(checkTreeError(h, &h->ttree, st, &error));
      if (error == 4u && h->ttree.factor < 256) {
        h->ttree.factor *= 2;
        continue;
      }
      PGMO_TRY(treeErrorToRc(error));
      break;
    }
    while (h->mode == 0) {
      ++h->last_rounds;
      PGMO_TRY(h->ttree.reset(n_cand, true, st));
      PGMO_CUDA(launchOctGrow(h->ttree.frames.ptr, h->cand_xyz.ptr, member, 0, nc, res, st));
      PGMO_CUDA(launchOctInsert(h->ttree.frames.ptr, h->ttree.table(), cpts, member, 0, nc, res, st));
      // the "membership changed" flag is the spare word of the frames header: zeroed by the
      // reset above, read back with the header by checkTreeError
      PGMO_CUDA(launchOctDecideTemp(h->ttree.frames.ptr, h->ttree.table(), h->cand_xyz.ptr,
                                    h->cand_next.ptr, nc, res, member, member_next, h->decision.ptr,
                                    &h->ttree.frames.ptr->pad, st));
      uint32_t error = 0;
      PGMO_TRY(checkTreeError(h, &h->ttree, st, &error));  // synchronises
      if (error == 4u && h->ttree.factor < 256) {  // table full: same round again, twice the room
        h->ttree.factor *= 2;
        --h->last_rounds;
        continue;
      }
      PGMO_TRY(treeErrorToRc(error));
      if (h->h_frames.ptr->pad == 0) break;  // the hypothesis reproduced itself
      std::swap(member, member_next);
      if (h->last_rounds > n_cand + 2) {
        setLastError("octree temp pass did not converge");
        return PGMO_ERR_CUDA;
      }
    }
    PGMO_TRY(compact(h, member, n_cand, &h->pot_list, 1, st, &n_pot));
    if (n_pot > 0) {
      octScatterRankKernel<<<blocksFor(n_pot), 256, 0, st>>>(h->pot_list.ptr, static_cast<uint32_t>(n_pot),
                                                             h->pot_of_cand.ptr);
    }
  }

  // ---- face check of the vertex loop (:304-343) and ordered acceptance (:349-378)
  PGMO_TRY(h->check.reserve(n_pot + 1));
  PGMO_TRY(h->acc_rank.reserve(n_pot + 1));
  PGMO_CUDA(cudaMemsetAsync(h->surf.ptr, 0, n, st));
  PGMO_CUDA(cudaMemsetAsync(h->keep.ptr, 0, n, st));
  PGMO_CUDA(cudaMemsetAsync(h->check.ptr, 0, n_pot + 1, st));
  PGMO_CUDA(cudaMemsetAsync(h->acc_rank.ptr, 0xFF, (n_pot + 1) * sizeof(int32_t), st));
  octTempReindexKernel<<<blocksFor(n), 256, 0, st>>>(n, h->reobs.ptr, h->act_index.ptr,
                                                     h->cand_of_vertex.ptr, h->decision.ptr,
                                                     h->pot_of_cand.ptr, n_orig, h->tr.ptr);
  octFaceCheckKernel<<<blocksFor(n), 256, 0, st>>>(n, h->tri_end.ptr, h->tr.ptr, n_orig, h->surf.ptr,
                                                   h->check.ptr);
  uint64_t n_acc = 0;
  if (n_pot > 0) {
    PGMO_TRY(compact(h, h->check.ptr, n_pot, &h->acc_list, 2, st, &n_acc));
    if (n_acc > 0) {
      octScatterRankKernel<<<blocksFor(n_acc), 256, 0, st>>>(h->acc_list.ptr, static_cast<uint32_t>(n_acc),
                                                             h->acc_rank.ptr);
    }
  }
  // room for the accepted vertices
  PGMO_TRY(h->all_vertices.reserve(roomy(n_orig + n_acc + 1, h->all_vertices.cap), st, n_orig));
  PGMO_TRY(h->all_stamps.reserve(roomy(n_orig + n_acc + 1, h->all_stamps.cap), st, n_orig));
  PGMO_TRY(h->act_xyz.reserve(roomy(h->n_act + n_acc + 1, h->act_xyz.cap), st, h->n_act));
  PGMO_TRY(h->act_index.reserve(roomy(h->n_act + n_acc + 1, h->act_index.cap), st, h->n_act));
  PGMO_TRY(h->act_stamp.reserve(roomy(h->n_act + n_acc + 1, h->act_stamp.cap), st, h->n_act));
  PGMO_TRY(h->act_next.reserve(h->act_xyz.cap, st, h->n_act));
  PGMO_TRY(h->act_sigma.reserve(h->act_xyz.cap, st, h->n_act));
  // prune's compaction targets follow at once (they would otherwise be allocated inside the
  // first prune that removes something, in the middle of the stream)
  PGMO_TRY(h->act_xyz2.reserve(h->act_xyz.cap));
  PGMO_TRY(h->act_index2.reserve(h->act_index.cap));
  PGMO_TRY(h->act_stamp2.reserve(h->act_stamp.cap));
  PGMO_TRY(h->new_vertices.reserve(n_acc + 1));
  PGMO_TRY(h->reobs_flag.reserve(roomy(n_orig + 1, h->reobs_flag.cap)));
  PGMO_CUDA(cudaMemsetAsync(h->reobs_flag.ptr, 0, n_orig + 1, st));
  h->committing = true;  // from here on the stored mesh changes
  if (n_acc > 0) {
    octCommitVerticesKernel<<<blocksFor(n_acc), 256, 0, st>>>(
        static_cast<uint32_t>(n_acc), h->acc_list.ptr, h->pot_list.ptr, h->cand_list.ptr, h->d_xyz.ptr,
        (rgba || vb) ? h->d_rgba.ptr : nullptr, n_orig, h->n_act, stampFromSec(stamp_s), stamp_s,
        h->all_vertices.ptr, h->all_stamps.ptr, h->act_xyz.ptr, h->act_index.ptr, h->act_stamp.ptr,
        h->new_vertices.ptr);
  }
  octReindexKernel<<<blocksFor(n), 256, 0, st>>>(n, h->reobs.ptr, h->act_index.ptr, h->cand_of_vertex.ptr,
                                                 h->decision.ptr, h->pot_of_cand.ptr, h->acc_rank.ptr,
                                                 n_orig, stamp_s, h->act_stamp.ptr, h->reobs_flag.ptr,
                                                 h->reindex.ptr, h->d_remap_index.ptr, h->d_remap_kind.ptr);
  // new_indices: re-observed indices ascending (sorted-unique insert, :293-299), then the new ones
  uint64_t n_reobs = 0;
  PGMO_TRY(h->new_indices.reserve(roomy(n_orig + n_acc + 1, h->new_indices.cap)));
  if (n_orig > 0) {
    PGMO_TRY(h->tiles.reserve(roomy(vccTileScratchEntries(n_orig) + 1, h->tiles.cap)));
    PGMO_CUDA(launchVccCompactLive(h->reobs_flag.ptr, n_orig, reinterpret_cast<uint64_t*>(h->new_indices.ptr),
                                   h->tiles.ptr, h->counts.ptr + 3, st));
  } else {
    PGMO_CUDA(cudaMemsetAsync(h->counts.ptr + 3, 0, sizeof(unsigned long long), st));
  }
  if (n_acc > 0) {
    octAppendNewIndicesKernel<<<blocksFor(n_acc), 256, 0, st>>>(static_cast<uint32_t>(n_acc), h->counts.ptr + 3,
                                                                n_orig, h->new_indices.ptr);
  }

  // ---- second pass over the candidate faces (:382-416)
  PGMO_TRY(ensurePolyTable(h, n / 3 + 1, st));
  {
    const uint64_t need = 3 * (h->n_poly + n / 3 + 1);
    PGMO_TRY(h->polygons.reserve(need <= h->polygons.cap ? need : std::max<uint64_t>(2 * need, 1ull << 22), st,
                                 3 * h->n_poly));
  }
  PGMO_TRY(h->new_triangles.reserve(3ull * (n / 3 + 1)));
  octFaceClassifyKernel<<<blocksFor(n), 256, 0, st>>>(n, h->surf.ptr, h->reindex.ptr, n_orig, polyTable(h),
                                                      h->keep.ptr, h->face_slot.ptr, h->flags32.ptr + 1);
  octFaceFinalizeKernel<<<blocksFor(n), 256, 0, st>>>(n, h->reindex.ptr, polyTable(h), h->keep.ptr,
                                                      h->face_slot.ptr, h->flags32.ptr + 1);
  uint64_t n_faces = 0;
  PGMO_TRY(compact(h, h->keep.ptr, n, &h->face_list, 4, st, &n_faces));
  if (n_faces > 0) {
    octCommitFacesKernel<<<blocksFor(n_faces), 256, 0, st>>>(
        static_cast<uint32_t>(n_faces), h->face_list.ptr, h->reindex.ptr, h->n_poly, polyTable(h),
        h->polygons.ptr, h->new_triangles.ptr, h->flags32.ptr + 1);
  }
  PGMO_CUDA(cudaGetLastError());

  // ---- outputs, then updateStructure for the accepted vertices (:358)
  PGMO_CUDA(cudaMemcpyAsync(remap_index, h->d_remap_index.ptr, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaMemcpyAsync(remap_kind, h->d_remap_kind.ptr, n, cudaMemcpyDeviceToHost, st));
  PGMO_TRY(readCount(h, 3, st, &n_reobs));  // synchronises
  PGMO_CUDA(cudaMemcpyAsync(h->h_flags.ptr, h->flags32.ptr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  PGMO_CUDA(cudaStreamSynchronize(st));
  PGMO_TRY(treeErrorToRc(h->h_flags.ptr[1]));
  h->last_cand = n_cand;
  h->last_pot = n_pot;
  const uint64_t first_new = h->n_act;
  h->n_all = n_orig + n_acc;
  h->n_act += n_acc;
  h->n_poly += n_faces;
  h->n_new_vertices = n_acc;
  h->n_new_triangles = n_faces;
  h->n_new_indices = n_reobs + n_acc;
  if (n_acc > 0) PGMO_TRY(buildPersistent(h, first_new, st));
  return PGMO_OK;
}

// refuses a poisoned handle; poisons it when a call fails after it started to commit
static int guardedIntegrate(pgmo_oct_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                            const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                            const VoxbloxInput* vb, double stamp_s, int64_t* remap_index,
                            uint8_t* remap_kind) {
  PGMO_CHECK_ARG(h != nullptr);
  if (h->poisoned) {
    setLastError("octree compression: an earlier integrate failed while the stored mesh was being "
                 "updated; this handle is unusable, create a new compressor");
    return PGMO_ERR_INVALID;
  }
  h->committing = false;
  const int rc = octIntegrate(h, n_blocks, block_idx_xyz, block_offsets, xyz, rgba, vb, stamp_s,
                              remap_index, remap_kind);
  if (rc < 0 && h->committing) h->poisoned = true;
  h->committing = false;
  return rc;
}

int pgmo_oct_integrate(pgmo_oct_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                       const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                       double stamp_s, int64_t* remap_index, uint8_t* remap_kind) {
  return guardedIntegrate(h, n_blocks, block_idx_xyz, block_offsets, xyz, rgba, nullptr, stamp_s,
                          remap_index, remap_kind);
}

int pgmo_oct_integrate_voxblox(pgmo_oct_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                               const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                               const uint16_t* z, const uint8_t* r, const uint8_t* g,
                               const uint8_t* b, float block_edge_length, double stamp_s,
                               int64_t* remap_index, uint8_t* remap_kind) {
  const VoxbloxInput in{x, y, z, r, g, b, block_edge_length};
  return guardedIntegrate(h, n_blocks, block_idx_xyz, block_offsets, nullptr, nullptr, &in, stamp_s,
                          remap_index, remap_kind);
}

int pgmo_oct_prune(pgmo_oct_t* h, double earliest_s) {
  PGMO_CHECK_ARG(h != nullptr);
  if (h->poisoned) {
    setLastError("octree compression: this handle is unusable after a failed integrate");
    return PGMO_ERR_INVALID;
  }
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  if (h->n_act == 0) return PGMO_OK;  // :420-422
  const uint32_t n = static_cast<uint32_t>(h->n_act);
  PGMO_TRY(h->flag.reserve(n));
  octPruneFlagsKernel<<<blocksFor(n), 256, 0, st>>>(h->act_stamp.ptr, n, earliest_s, h->flag.ptr);
  uint64_t n_keep = 0;
  PGMO_TRY(compact(h, h->flag.ptr, n, &h->acc_list, 5, st, &n_keep));
  if (n_keep >= h->n_act) return PGMO_OK;  // nothing removed: structure untouched (:450)
  PGMO_TRY(h->act_xyz2.reserve(h->act_xyz.cap));
  PGMO_TRY(h->act_index2.reserve(h->act_index.cap));
  PGMO_TRY(h->act_stamp2.reserve(h->act_stamp.cap));
  if (n_keep > 0) {
    octPruneGatherKernel<<<blocksFor(n_keep), 256, 0, st>>>(
        h->acc_list.ptr, static_cast<uint32_t>(n_keep), h->act_xyz.ptr, h->act_index.ptr, h->act_stamp.ptr,
        h->act_xyz2.ptr, h->act_index2.ptr, h->act_stamp2.ptr);
    PGMO_CUDA(cudaGetLastError());
  }
  PGMO_CUDA(cudaStreamSynchronize(st));
  std::swap(h->act_xyz.ptr, h->act_xyz2.ptr);
  std::swap(h->act_xyz.cap, h->act_xyz2.cap);
  std::swap(h->act_index.ptr, h->act_index2.ptr);
  std::swap(h->act_index.cap, h->act_index2.cap);
  std::swap(h->act_stamp.ptr, h->act_stamp2.ptr);
  std::swap(h->act_stamp.cap, h->act_stamp2.cap);
  h->n_act = n_keep;
  return buildPersistent(h, 0, st);  // reInitializeStructure: re-anchored at the first survivor
}

int pgmo_oct_device_view(pgmo_oct_t* h, const void** d_vertices_xyzw, const uint64_t** d_stamps_ns,
                         uint64_t* n_vertices) {
  PGMO_CHECK_ARG(h != nullptr && n_vertices != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  PGMO_CUDA(cudaStreamSynchronize(h->stream));  // everything queued so far is visible to any stream
  if (d_vertices_xyzw) *d_vertices_xyzw = h->all_vertices.ptr;
  if (d_stamps_ns) *d_stamps_ns = reinterpret_cast<const uint64_t*>(h->all_stamps.ptr);
  *n_vertices = h->n_all;
  return PGMO_OK;
}

int pgmo_oct_sizes(pgmo_oct_t* h, uint64_t* sizes8) {
  PGMO_CHECK_ARG(h != nullptr && sizes8 != nullptr);
  sizes8[0] = h->n_all;
  sizes8[1] = h->n_poly;
  sizes8[2] = h->n_act;
  sizes8[3] = 0;
  sizes8[4] = h->n_new_vertices;
  sizes8[5] = h->n_new_triangles;
  sizes8[6] = h->n_new_indices;
  sizes8[7] = h->n_act;
  return PGMO_OK;
}

int pgmo_oct_snapshot(pgmo_oct_t* h, float* vertices_xyz, uint32_t* vertices_rgba,
                      uint64_t* stamps_ns, uint32_t* polygons, uint64_t* active_index,
                      float* active_xyz, double* active_stamps_s) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  std::vector<float4> v;
  if ((vertices_xyz || vertices_rgba) && h->n_all) {
    v.resize(h->n_all);
    PGMO_CUDA(cudaMemcpyAsync(v.data(), h->all_vertices.ptr, h->n_all * sizeof(float4), cudaMemcpyDeviceToHost, st));
  }
  if (stamps_ns && h->n_all) {
    PGMO_CUDA(cudaMemcpyAsync(stamps_ns, h->all_stamps.ptr, h->n_all * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  }
  if (polygons && h->n_poly) {
    PGMO_CUDA(cudaMemcpyAsync(polygons, h->polygons.ptr, 3 * h->n_poly * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  std::vector<uint32_t> ai;
  std::vector<float4> ax;
  if (active_index && h->n_act) {
    ai.resize(h->n_act);
    PGMO_CUDA(cudaMemcpyAsync(ai.data(), h->act_index.ptr, h->n_act * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  if (active_xyz && h->n_act) {
    ax.resize(h->n_act);
    PGMO_CUDA(cudaMemcpyAsync(ax.data(), h->act_xyz.ptr, h->n_act * sizeof(float4), cudaMemcpyDeviceToHost, st));
  }
  if (active_stamps_s && h->n_act) {
    PGMO_CUDA(cudaMemcpyAsync(active_stamps_s, h->act_stamp.ptr, h->n_act * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  PGMO_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < v.size(); ++i) {
    if (vertices_xyz) {
      vertices_xyz[3 * i] = v[i].x;
      vertices_xyz[3 * i + 1] = v[i].y;
      vertices_xyz[3 * i + 2] = v[i].z;
    }
    if (vertices_rgba) std::memcpy(vertices_rgba + i, &v[i].w, sizeof(uint32_t));
  }
  for (size_t i = 0; i < ai.size(); ++i) active_index[i] = ai[i];
  for (size_t i = 0; i < ax.size(); ++i) {
    active_xyz[3 * i] = ax[i].x;
    active_xyz[3 * i + 1] = ax[i].y;
    active_xyz[3 * i + 2] = ax[i].z;
  }
  return PGMO_OK;
}

int pgmo_oct_last_stats(pgmo_oct_t* h, uint64_t* stats4) {
  PGMO_CHECK_ARG(h != nullptr && stats4 != nullptr);
  stats4[0] = h->last_rounds;
  stats4[1] = h->last_cand;
  stats4[2] = h->last_pot;
  stats4[3] = h->n_new_vertices;
  return PGMO_OK;
}

int pgmo_oct_last_outputs(pgmo_oct_t* h, float* new_vertices_xyz, uint32_t* new_vertices_rgba,
                          uint32_t* new_triangles, uint64_t* new_indices) {
  PGMO_CHECK_ARG(h != nullptr);
  PGMO_TRY(ensureDevice(h->device));
  cudaStream_t st = h->stream;
  std::vector<float4> v;
  if ((new_vertices_xyz || new_vertices_rgba) && h->n_new_vertices) {
    v.resize(h->n_new_vertices);
    PGMO_CUDA(cudaMemcpyAsync(v.data(), h->new_vertices.ptr, v.size() * sizeof(float4), cudaMemcpyDeviceToHost, st));
  }
  if (new_triangles && h->n_new_triangles) {
    PGMO_CUDA(cudaMemcpyAsync(new_triangles, h->new_triangles.ptr, 3 * h->n_new_triangles * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, st));
  }
  if (new_indices && h->n_new_indices) {
    PGMO_CUDA(cudaMemcpyAsync(new_indices, h->new_indices.ptr, h->n_new_indices * sizeof(uint64_t),
                              cudaMemcpyDeviceToHost, st));
  }
  PGMO_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < v.size(); ++i) {
    if (new_vertices_xyz) {
      new_vertices_xyz[3 * i] = v[i].x;
      new_vertices_xyz[3 * i + 1] = v[i].y;
      new_vertices_xyz[3 * i + 2] = v[i].z;
    }
    if (new_vertices_rgba) std::memcpy(new_vertices_rgba + i, &v[i].w, sizeof(uint32_t));
  }
  return PGMO_OK;
}

}  // extern "C"

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadOctApiKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, octTriEndKernel);
  cudaFuncGetAttributes(&attr, octCandFlagsKernel);
  cudaFuncGetAttributes(&attr, octGatherCandKernel);
  cudaFuncGetAttributes(&attr, octScatterRankKernel);
  cudaFuncGetAttributes(&attr, octTempReindexKernel);
  cudaFuncGetAttributes(&attr, octFaceCheckKernel);
  cudaFuncGetAttributes(&attr, octCommitVerticesKernel);
  cudaFuncGetAttributes(&attr, octReindexKernel);
  cudaFuncGetAttributes(&attr, octAppendNewIndicesKernel);
  cudaFuncGetAttributes(&attr, octFaceClassifyKernel);
  cudaFuncGetAttributes(&attr, octFaceFinalizeKernel);
  cudaFuncGetAttributes(&attr, octCommitFacesKernel);
  cudaFuncGetAttributes(&attr, octPolyRehashKernel);
  cudaFuncGetAttributes(&attr, octPruneFlagsKernel);
  cudaFuncGetAttributes(&attr, octPruneGatherKernel);
  cudaGetLastError();
}
}  // namespace pgmo
