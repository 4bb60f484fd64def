// Device-side layout and launchers of VoxelClearingCompression (sm_100a).
//
// Replaces VoxelClearingCompression::updateRemapping / pruneMeshBlocks
// (kimera_pgmo/src/compression/voxel_clearing_compression.cpp:72-101, :111-210).
//
// Persistent state in HBM:
//   voxel hash   open addressing, linear probing: keys[] (packed voxel key + 1; 0 = empty,
//                ~0 = tombstone) and vals[] (mesh index, 0xFFFFFFFF until one is assigned)
//   per slot     seq_min / seq_max: (block-pass id, sequence number) words that give, without
//                any clearing between passes, the first and the last occurrence of the voxel in
//                the block being integrated (first occurrence decides the new index, the last
//                occurrence is the writer of all_vertices_[index])
//   per index    active / inactive reference counts, live flag, vertex (xyz, rgba), stamp
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace pgmo {

constexpr uint64_t kVccEmpty = 0ull;
constexpr uint64_t kVccTomb = ~0ull;
constexpr uint32_t kVccNoIndex = 0xFFFFFFFFu;
constexpr int kVccKeyBits = 21;  // per axis, two's complement offset: |key| < 2^20

// global counters (device, u64)
enum VccCounter {
  kVccMaxIndex = 0,    // max_index_
  kVccLive = 1,        // live voxel->index entries (== indices_to_active_refs_.size())
  kVccTombstones = 2,
  kVccInvalid = 3,     // empty_slots_.size()
  kVccError = 4,       // bit 0: voxel key out of the packed range; bit 1: table full;
                       // bit 2: a previous voxel was not found (state corrupted)
  kVccNumGlobal = 8
};

struct VccTable {
  uint64_t* keys;
  uint32_t* vals;
  unsigned long long* seq_min;
  unsigned long long* seq_max;
  uint32_t mask;  // capacity - 1
};

struct VccIndexState {
  int32_t* active;
  int32_t* inactive;
  uint8_t* live;
  float4* vertices;   // xyz + rgba bits
  uint64_t* stamps;
  uint32_t* invalid;  // empty_slots_
};

struct VccBlockArgs {
  const float* xyz;        // message vertices, interleaved
  const uint32_t* rgba;    // may be null
  uint64_t first_vertex;   // block = vertices [first_vertex, first_vertex + n)
  uint32_t n;
  double inv_resolution;   // 1.0 / resolution_
  uint64_t stamp_ns;
  uint32_t pass;           // unique per (message, block)
  uint32_t* slot_of_vertex;  // scratch, message-sized
  uint64_t* remap_out;     // message-sized
  int64_t* keys_out;       // optional, 3 per vertex
  uint64_t* curr_list;     // distinct packed keys of this block (capacity >= n)
  unsigned long long* curr_count;
};

cudaError_t launchVccInsert(const VccTable& t, const VccIndexState& s, const VccBlockArgs& b,
                            unsigned long long* counters, cudaStream_t stream);
// tile scratch (u32) the ordered kernels need for n elements
uint64_t vccTileScratchEntries(uint64_t n);
cudaError_t launchVccAssign(const VccTable& t, const VccIndexState& s, const VccBlockArgs& b,
                            uint32_t* tile_scratch, unsigned long long* counters,
                            cudaStream_t stream);
cudaError_t launchVccWrite(const VccTable& t, const VccIndexState& s, const VccBlockArgs& b,
                           cudaStream_t stream);
// updateRemapping's pass over the block's previous voxel set (:168-208): archive == false;
// pruneMeshBlocks (:72-101): archive == true (only the active count is decremented).
cudaError_t launchVccRelease(const VccTable& t, const VccIndexState& s, const uint64_t* prev_list,
                             uint64_t prev_count, bool archive, unsigned long long* counters,
                             cudaStream_t stream);
cudaError_t launchVccRehash(const VccTable& from, const VccTable& to,
                            unsigned long long* counters, cudaStream_t stream);
cudaError_t launchVccInitTable(const VccTable& t, cudaStream_t stream);
// ascending list of live indices (getActiveVerticesIndex)
cudaError_t launchVccCompactLive(const uint8_t* live, uint64_t n, uint64_t* out,
                                 uint32_t* tile_scratch, unsigned long long* out_count,
                                 cudaStream_t stream);

}  // namespace pgmo
