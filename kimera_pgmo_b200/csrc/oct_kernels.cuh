// OctreeCompression on the GPU (sm_100a): device-side layout and launchers.
//
// Replaces MeshCompression::compressAndIntegrate / pruneStoredMesh
// (kimera_pgmo/src/compression/mesh_compression.cpp:216-464) specialised by OctreeCompression
// (kimera_pgmo/src/compression/octree_compression.cpp:17-81, include/.../octree_compression.h:
// 20-27), including the slice of pcl::octree::OctreePointCloudSearch the reference relies on:
// bounding-box growth (adoptBoundingBoxToPoint), leaf keys, getLeafCount() == 0, and
// approxNearestSearch's greedy descent by voxel-centre distance.
//
// Implicit octree.  PCL's pointer octree is replaced by a hash set of occupied nodes.  Growth
// only ever doubles the box, moving `min` by whole old-box sides, so grid lines never move: with
// S = cumulative leaf-key shift since the (re)build and sigma = key_at_insertion - S_at_insertion
// (a point's "stable" leaf coordinate), the node that holds the point at height h (h = 0: leaf)
// is  nu_h = floor((sigma + base_h) / 2^h),  base_h = S mod 2^h  (fixed from the moment the tree
// is deeper than h), and its key in the frame at time T is  nu_h + (S_T - base_h) / 2^h.
// The table maps (h, nu_h) -> the smallest sequence number of a point below that node, so one
// table answers "did this branch exist when point c was examined?" for every c at once.  Leaves
// also head a linked list of their points.
//
// Sequential parts of the reference and how they are parallelised (bit-exact results):
//   * persistent-structure test (mesh_compression.cpp:268): read-only during the vertex loop
//     -> one thread per input vertex;
//   * temp-structure greedy (:269-283): vertex c's outcome depends only on EARLIER vertices.
//     Start from the hypothesis "every candidate becomes a temp vertex", build the time-stamped
//     tree for the hypothesis, let every candidate decide against the tree as of its own
//     sequence number, and iterate until the hypothesis reproduces itself.  A fixed point
//     satisfies the reference's recurrence vertex by vertex, hence equals the sequential result;
//     the correct prefix grows by at least one vertex per round (2-3 rounds in practice);
//   * box growth (adoptBoundingBoxToPoint per accepted point, in order): "first point outside
//     the current box" searches, one per doubling (<= ~30);
//   * face bookkeeping: per-triangle kernels + ordered compaction (tile scans).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace pgmo {

constexpr int kOctMaxFrames = 48;   // anchor + doublings
constexpr int kOctMaxDepth = 30;
constexpr uint64_t kOctEmptyKey = ~0ull;
constexpr uint32_t kOctNone = 0xFFFFFFFFu;
constexpr int kOctCoordBits = 19;   // node coordinates are offset by 2^18: +-262143 leaves

struct OctFrame {
  double min[3];
  double max[3];
  long long shift[3];  // S: cumulative leaf-key shift since the (re)build
  uint32_t depth;
  uint32_t seq;        // sequence number of the point whose insertion produced this frame
};

struct OctFrames {
  OctFrame f[kOctMaxFrames];  // f[e]: frame after event e (event 0 = the anchor point)
  long long base[32][3];      // base_h per height
  long long anchor_sigma[3];  // stable leaf coordinate of the anchor point
  uint32_t n;                 // events so far (0 = no box yet: getLeafCount() == 0)
  uint32_t error;             // bit 0: too many events / depth; bit 1: coordinate out of range;
                              // bit 2: table full; bit 3: polygon hash collision
  uint32_t cursor;            // growth search position
  uint32_t pad;
};

struct OctTable {
  unsigned long long* keys;   // packed (h, nu); kOctEmptyKey = empty
  uint32_t* minseq;
  uint32_t* head;             // leaves: newest point of the leaf's list
  uint32_t mask;              // capacity - 1
};

// points of a tree: xyz (+ unused w), linked-list links, stable leaf coordinates
struct OctPoints {
  const float4* xyz;
  uint32_t* next;
  int4* sigma;
};

// Box growth for points [first, n) of `pts` that are members (member == nullptr: all), in
// order; appends events to frames.  One CTA.
cudaError_t launchOctGrow(OctFrames* frames, const float4* xyz, const uint8_t* member,
                          uint32_t first, uint32_t n, double resolution, cudaStream_t stream);
// Registers points [first, n) (members only) at every height of the final frame, links them
// into their leaf lists and stores sigma.
cudaError_t launchOctInsert(OctFrames* frames, OctTable table, OctPoints pts,
                            const uint8_t* member, uint32_t first, uint32_t n, double resolution,
                            cudaStream_t stream);
// The old-root chain: (h, nu_h(anchor)) for every height created by growth (idempotent).
cudaError_t launchOctInsertRootChain(OctFrames* frames, OctTable table, cudaStream_t stream);

// inOctree(octree_, v) for every input vertex (octree_compression.cpp:17-40) against the whole
// persistent tree: result[i] = matched active position or -1.
cudaError_t launchOctQueryPersistent(const OctFrames* frames, OctTable table, const float4* act_xyz,
                                     const uint32_t* next, const float* xyz, uint32_t n,
                                     double resolution, int32_t* result, cudaStream_t stream);
// inOctree(temp_octree_, v) for every candidate c against the temp tree as of sequence c.
// decision[c] = -1 (temp-unique) or the ordinal of the temp point it converges to;
// member_out[c] = (decision < 0); *changed |= member_out != member_in.
cudaError_t launchOctDecideTemp(const OctFrames* frames, OctTable table, const float4* cand_xyz,
                                const uint32_t* next, uint32_t n_cand, double resolution,
                                const uint8_t* member_in, uint8_t* member_out, int32_t* decision,
                                uint32_t* changed, cudaStream_t stream);

// ---- BlockCompression (kimera_pgmo/src/compression/block_compression.cpp:19-81): the structure
// is a cell hash, cell = floor(p * (1 / voxel_size)) per axis in float (spatial_hash::Grid::toIndex
// as restated by the oracle), value = the first point of the cell (`emplace` never overwrites).
// Re-uses OctTable: keys = packed cell, minseq = smallest point index, frames->error for errors.

// emplace(toIndex(p_i), i) for points [first, n)
cudaError_t launchCellInsert(OctFrames* frames, OctTable table, const float4* xyz, uint32_t first,
                             uint32_t n, float voxel_size_inv, cudaStream_t stream);
// checkIfVertexUnique for every input vertex: result[i] = matched active position or -1
cudaError_t launchCellQueryPersistent(OctTable table, const float* xyz, uint32_t n,
                                      float voxel_size_inv, int32_t* result, cudaStream_t stream);
// checkIfVertexTempUnique in sequence order: the first candidate of a cell is temp-unique, every
// later one converges to it.  decision[c] = -1 or the ordinal of the first candidate of c's cell.
cudaError_t launchCellDecideTemp(OctTable table, const float4* cand_xyz, uint32_t n_cand,
                                 float voxel_size_inv, uint8_t* member_out, int32_t* decision,
                                 cudaStream_t stream);

}  // namespace pgmo
