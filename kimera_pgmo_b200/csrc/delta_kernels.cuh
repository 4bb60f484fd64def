// DeltaCompression on the device (sm_100a): data layout and launchers.
//
// Replaces kimera_pgmo::DeltaCompression::update / archiveBlocks
// (kimera_pgmo/src/compression/delta_compression.cpp:54-458).  Iteration order of the reference's
// unordered containers is fixed to insertion order, exactly like oracle/delta_oracle.h.
//
// Persistent state in HBM:
//   voxel hash    open addressing: keys[] (packed voxel key + 1; 0 = empty, ~0 = tombstone),
//                 vslot[] (index of the voxel's record), first[] (per-block first-occurrence word)
//   voxel records append-only SoA in insertion order (= the reference's vertices_map_ order):
//                 key, point, stamp, mesh_index, sequence number, active refs, needs_archive,
//                 is_new, alive
//   block pools   per block its flat face index list (BlockInfo::indices) and the records of its
//                 distinct voxels in first-observation order (BlockInfo::vertices)
//   pending       archived_faces_ (faces waiting for all their vertices to be archived)
// The delta under construction (archive_delta_ / delta_) lives in device buffers too: vertices
// (archived ones first), the two face lists, prev_to_curr as a dense array over the previous
// delta's index space, deleted / observed / new as flag arrays.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace pgmo {

constexpr uint32_t kDeltaNone = 0xFFFFFFFFu;
constexpr int kDeltaTile = 1024;

enum DeltaCounter {
  kDcNumRecords = 0,   // voxel records ever created
  kDcActiveSlots = 1,  // active_remapping_.size() of this pass
  kDcArchVerts = 2,    // archived vertices in the delta under construction
  kDcFacesUpd = 3,     // delta.face_updates.size()
  kDcFacesArch = 4,    // delta.face_archive_updates.size()
  kDcPending = 5,      // archived_faces_.size()
  kDcError = 6,        // bit 0: key out of range, bit 1: table full, bit 2: .at() would throw
  kDcChanged = 7,      // fixed-point flag
  kDcScratch = 8,      // per-call scratch counts [8..15]
  kDcCount = 16
};

struct DeltaTable {
  unsigned long long* keys;
  uint32_t* vslot;
  unsigned long long* first;
  uint32_t mask;
};

struct DeltaRecords {
  unsigned long long* key;
  float4* point;  // xyz + rgba bits
  unsigned long long* stamp;
  uint32_t* mesh_index;
  uint32_t* seq;
  int32_t* refs;
  uint8_t* needs_archive;
  uint8_t* is_new;
  uint8_t* alive;
  uint32_t* label;  // VoxelInfo::label (delta_compression.h:23-24); kDeltaNoLabel = nullopt
};
constexpr uint32_t kDeltaNoLabel = 0xFFFFFFFFu;

struct DeltaOut {      // the delta under construction
  float4* point;       // [archived..., active...]
  unsigned long long* stamp;
  uint32_t* label;     // semantic_updates (kDeltaNoLabel where the reference pushed nothing)
  uint8_t* is_new;     // per delta vertex (local index)
  uint8_t* observed;
  uint32_t* p2c;       // dense over the previous delta's global indices, kDeltaNone = absent
  uint8_t* deleted;    // same index space
  uint32_t prev_total; // size of that index space
  uint32_t vertex_start;
  uint32_t* faces_upd;   // 3 per face
  uint32_t* faces_arch;  // 3 per face
};

struct DeltaBlockArgs {
  const float* xyz;      // message vertices, interleaved
  const uint32_t* rgba;  // may be null
  const uint32_t* label; // MeshInterface::getActiveSemantics per vertex, may be null
  uint64_t first_vertex;
  uint32_t n;
  double index_scale;    // 1.0 / resolution
  unsigned long long stamp_ns;
  uint32_t pass;         // unique per (update, block)
  uint32_t seq;          // sequence_number_
  uint32_t* tslot_of_vertex;  // scratch, message-sized
  uint32_t* face_map;    // BlockInfo::indices of this block (n entries)
  uint32_t* voxels;      // BlockInfo::vertices of this block (capacity n)
  uint32_t* voxel_count; // its size (device)
  uint32_t* active_remapping;
};

cudaError_t launchDeltaInitTable(const DeltaTable& t, cudaStream_t s);
cudaError_t launchDeltaRehash(const DeltaTable& t, const DeltaRecords& r, const unsigned long long* counters,
                              uint32_t n_upper, unsigned long long* err, cudaStream_t s);
// updateRemapping for one block: addPoint for every vertex (:54-92)
cudaError_t launchDeltaAddBlock(const DeltaTable& t, const DeltaRecords& r, const DeltaBlockArgs& b,
                                uint32_t* tiles, unsigned long long* counters, cudaStream_t s);
// both of the above for a block and a previous voxel list of at most kDeltaTile entries, in one
// launch (prev_voxels == nullptr: a new block)
cudaError_t launchDeltaSmallBlock(const DeltaTable& t, const DeltaRecords& r, const DeltaOut& o,
                                  const DeltaBlockArgs& b, const uint32_t* prev_voxels,
                                  const uint32_t* prev_count, unsigned long long* counters,
                                  cudaStream_t s);
// removeBlockObservations (:94-126, archive == false) / the vertex loop of archiveBlocks
// (:336-349, archive == true) over one block's voxel list
cudaError_t launchDeltaRelease(const DeltaTable& t, const DeltaRecords& r, const DeltaOut& o,
                               const uint32_t* voxels, const uint32_t* voxel_count, uint32_t n_upper,
                               bool archive, uint8_t* flags, uint32_t* tiles,
                               unsigned long long* counters, cudaStream_t s);
// addActiveVertices (:128-154)
cudaError_t launchDeltaActiveVertices(const DeltaRecords& r, const DeltaOut& o, uint32_t n_upper,
                                      uint32_t seq, uint32_t* active_remapping, uint32_t* tiles,
                                      unsigned long long* counters, cudaStream_t s);

struct DeltaFaceBlock {  // one block of block_info_map_, in insertion order
  uint32_t face_offset;  // start in the face pool
  uint32_t n;            // indices (3 per face)
  uint32_t first_face;   // prefix sum of n / 3
  int64_t msg_offset;    // >= 0: updated this pass, start of the block in the message; < 0: not
};
// addActiveFaces' remap (:176-196): indices rewritten in place, remap_out filled for updated
// blocks, faces + validity written to the ordered work list
cudaError_t launchDeltaRemapFaces(const DeltaFaceBlock* blocks, uint32_t n_blocks, uint32_t n_faces,
                                  uint32_t* pool, const uint32_t* active_remapping, const DeltaOut& o,
                                  unsigned long long* remap_out, uint32_t* work_faces, uint8_t* valid,
                                  unsigned long long* counters, cudaStream_t s);

struct DeltaEdgeTable {
  unsigned long long* keys;  // (src << 32 | dst) + 1
  uint32_t* minseq;
  uint32_t mask;
};
// one round of the RedunancyChecker fixed point (redundancy_checker.cpp:12-23): insert the edges
// of the currently accepted faces, then re-decide every valid face against edges of earlier ones
cudaError_t launchDeltaEdgeRound(const DeltaEdgeTable& e, const uint32_t* faces, const uint8_t* valid,
                                 const uint8_t* acc_in, uint8_t* acc_out, uint32_t n_faces,
                                 unsigned long long* counters, cudaStream_t s);
// ordered append of the flagged faces (3 u32 each) to out[*count ...]; bumps *count
cudaError_t launchDeltaAppendFaces(const uint32_t* faces, const uint8_t* flags, uint32_t n_faces,
                                   uint32_t* out, unsigned long long* count, uint32_t* tiles,
                                   cudaStream_t s);
// updateAndAddArchivedFaces' remap (:424-428) / archiveBlockFaces' classification (:376-404)
cudaError_t launchDeltaPendingRemap(uint32_t* faces, uint32_t n_faces, const DeltaOut& o,
                                    uint32_t num_archived_vertices, uint8_t* valid, cudaStream_t s);
cudaError_t launchDeltaClassifyPending(const uint32_t* faces, const uint8_t* acc, uint32_t n_faces,
                                       const unsigned long long* counters, const DeltaOut& o,
                                       uint8_t* to_arch, uint8_t* to_upd, cudaStream_t s);
cudaError_t launchDeltaArchiveFaces(const DeltaFaceBlock* blocks, uint32_t n_blocks, uint32_t n_faces,
                                    const uint32_t* pool, const DeltaOut& o,
                                    uint32_t num_archived_vertices, const unsigned long long* counters,
                                    uint32_t* work_faces, uint8_t* to_pending, uint8_t* candidate,
                                    cudaStream_t s);
// stable compaction of the alive records into `to`; new_id[old] = new index (kDeltaNone = erased);
// the block voxel lists in voxel_pool[0, pool_used) are renumbered; counters[kDcNumRecords] =
// the number of alive records afterwards
cudaError_t launchDeltaCompactRecords(const DeltaRecords& from, const DeltaRecords& to, uint32_t n_old,
                                      uint32_t* new_id, uint32_t* voxel_pool, uint64_t pool_used,
                                      uint32_t* tiles, unsigned long long* counters, cudaStream_t s);
// pool garbage collection: segs[i] = {src offset, dst offset, count, pool (0 = a, 1 = b)}
cudaError_t launchDeltaCopySegments(const uint4* segs, uint32_t n_segs, const uint32_t* src_a,
                                    uint32_t* dst_a, const uint32_t* src_b, uint32_t* dst_b,
                                    cudaStream_t s);
uint64_t deltaTileEntries(uint64_t n);

}  // namespace pgmo
