// VoxelClearingCompression kernels (sm_100a): voxel-key hashing, dedup, index assignment in
// reference order, reference-count release.  See vcc_kernels.cuh for the data layout.
#include "vcc_kernels.cuh"

namespace pgmo {

namespace {

constexpr int kTile = 1024;  // threads per CTA in the ordered (scan) kernels

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// LongIndex(round(p.x * inv), round(p.y * inv), round(p.z * inv)) — float promoted to double,
// one rounded product, std::round (half away from zero), conversion to int64
// (voxel_clearing_compression.cpp:134-136)
__device__ __forceinline__ void voxelKey(const float* __restrict__ xyz, uint64_t v, double inv,
                                         long long key[3]) {
  key[0] = static_cast<long long>(round(__dmul_rn(static_cast<double>(xyz[3 * v + 0]), inv)));
  key[1] = static_cast<long long>(round(__dmul_rn(static_cast<double>(xyz[3 * v + 1]), inv)));
  key[2] = static_cast<long long>(round(__dmul_rn(static_cast<double>(xyz[3 * v + 2]), inv)));
}

__device__ __forceinline__ bool packKey(const long long key[3], uint64_t* packed) {
  const long long lim = 1ll << (kVccKeyBits - 1);
  if (key[0] < -lim || key[0] >= lim || key[1] < -lim || key[1] >= lim || key[2] < -lim ||
      key[2] >= lim) {
    return false;
  }
  const uint64_t x = static_cast<uint64_t>(key[0] + lim);
  const uint64_t y = static_cast<uint64_t>(key[1] + lim);
  const uint64_t z = static_cast<uint64_t>(key[2] + lim);
  *packed = (x | (y << kVccKeyBits) | (z << (2 * kVccKeyBits))) + 1;  // never 0, never ~0
  return true;
}

// finds the slot of `key`, inserting it when absent; returns false when the table is full
__device__ __forceinline__ bool findOrInsert(const VccTable& t, uint64_t key, uint32_t* slot,
                                             bool* inserted) {
  uint32_t h = static_cast<uint32_t>(mix64(key)) & t.mask;
  *inserted = false;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    uint64_t cur = t.keys[h];
    if (cur == key) {
      *slot = h;
      return true;
    }
    if (cur == kVccEmpty) {
      const unsigned long long old =
          atomicCAS(reinterpret_cast<unsigned long long*>(t.keys + h), kVccEmpty, key);
      if (old == kVccEmpty) {
        *slot = h;
        *inserted = true;
        return true;
      }
      if (old == key) {
        *slot = h;
        return true;
      }
    }
    h = (h + 1) & t.mask;
  }
  return false;
}

__device__ __forceinline__ bool findSlot(const VccTable& t, uint64_t key, uint32_t* slot) {
  uint32_t h = static_cast<uint32_t>(mix64(key)) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    const uint64_t cur = t.keys[h];
    if (cur == key) {
      *slot = h;
      return true;
    }
    if (cur == kVccEmpty) return false;
    h = (h + 1) & t.mask;
  }
  return false;
}

// K6: key + hash lookup/insert + first/last occurrence + per-block first-seen bookkeeping
__global__ void __launch_bounds__(256)
vccInsertKernel(const VccTable t, const VccIndexState s, const VccBlockArgs b,
                unsigned long long* __restrict__ counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  const uint64_t v = b.first_vertex + i;
  long long key[3];
  voxelKey(b.xyz, v, b.inv_resolution, key);
  if (b.keys_out) {
    b.keys_out[3 * v + 0] = key[0];
    b.keys_out[3 * v + 1] = key[1];
    b.keys_out[3 * v + 2] = key[2];
  }
  uint64_t packed;
  if (!packKey(key, &packed)) {
    atomicOr(counters + kVccError, 1ull);
    b.slot_of_vertex[v] = kVccNoIndex;
    return;
  }
  uint32_t slot;
  bool inserted;
  if (!findOrInsert(t, packed, &slot, &inserted)) {
    atomicOr(counters + kVccError, 2ull);
    b.slot_of_vertex[v] = kVccNoIndex;
    return;
  }
  b.slot_of_vertex[v] = slot;
  // first occurrence: newer pass wins through the inverted pass id, then the smallest sequence
  const unsigned long long min_word =
      (static_cast<unsigned long long>(~b.pass) << 32) | static_cast<unsigned long long>(i);
  atomicMin(t.seq_min + slot, min_word);
  // last occurrence; the thread that moves the slot into this pass is the voxel's first visitor
  const unsigned long long max_word =
      (static_cast<unsigned long long>(b.pass) << 32) | static_cast<unsigned long long>(i);
  const unsigned long long old = atomicMax(t.seq_max + slot, max_word);
  if (static_cast<uint32_t>(old >> 32) != b.pass) {
    // !curr_voxels.count(vertex_index): once per voxel per block (:157-162)
    const unsigned long long at = atomicAdd(b.curr_count, 1ull);
    b.curr_list[at] = packed;
    const uint32_t idx = t.vals[slot];
    if (idx != kVccNoIndex) {
      atomicAdd(s.active + idx, 1);
      atomicAdd(s.inactive + idx, 1);
    }
  }
}

// K7a: per tile, how many vertices create a new voxel (they are their voxel's first occurrence)
__device__ __forceinline__ bool isCreator(const VccTable& t, const VccBlockArgs& b, uint32_t i) {
  if (i >= b.n) return false;
  const uint32_t slot = b.slot_of_vertex[b.first_vertex + i];
  if (slot == kVccNoIndex) return false;
  if (t.vals[slot] != kVccNoIndex) return false;
  return static_cast<uint32_t>(t.seq_min[slot] & 0xFFFFFFFFull) == i;
}

__global__ void __launch_bounds__(kTile)
vccCountCreatorsKernel(const VccTable t, const VccBlockArgs b, uint32_t* __restrict__ tile_counts) {
  const uint32_t i = blockIdx.x * kTile + threadIdx.x;
  const bool c = isCreator(t, b, i);
  const int total = __syncthreads_count(c);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = static_cast<uint32_t>(total);
}

// exclusive scan of the tile counts (single CTA; tiles <= a few thousand), adds max_index_
__global__ void __launch_bounds__(kTile)
vccScanTilesKernel(uint32_t* __restrict__ tile_counts, uint32_t n_tiles,
                   unsigned long long* __restrict__ counters) {
  __shared__ uint32_t warp_sums[kTile / 32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_tiles; base += kTile) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_tiles ? tile_counts[i] : 0;
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += up;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = threadIdx.x < kTile / 32 ? warp_sums[threadIdx.x] : 0;
      uint32_t wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        if (threadIdx.x >= o) wi += up;
      }
      warp_sums[threadIdx.x] = wi - w;  // exclusive
    }
    __syncthreads();
    const uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
    if (i < n_tiles) tile_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == kTile - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // tile_counts[n_tiles] = total; the base index is applied by the assign kernel
    tile_counts[n_tiles] = carry;
  }
  (void)counters;
}

// K7b: new voxels get max_index_ + rank in first-occurrence order (:143-152)
__global__ void __launch_bounds__(kTile)
vccAssignKernel(const VccTable t, const VccIndexState s, const VccBlockArgs b,
                const uint32_t* __restrict__ tile_offsets,
                const unsigned long long* __restrict__ counters) {
  __shared__ uint32_t warp_sums[kTile / 32];
  const uint32_t i = blockIdx.x * kTile + threadIdx.x;
  const bool c = isCreator(t, b, i);
  const unsigned ballot = __ballot_sync(0xFFFFFFFFu, c);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = __popc(ballot);
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t w = warp_sums[threadIdx.x];
    uint32_t wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (threadIdx.x >= o) wi += up;
    }
    warp_sums[threadIdx.x] = wi - w;
  }
  __syncthreads();
  if (!c) return;
  const uint32_t rank = tile_offsets[blockIdx.x] + warp_sums[warp] + __popc(ballot & ((1u << lane) - 1));
  const uint32_t idx = static_cast<uint32_t>(counters[kVccMaxIndex]) + rank;
  const uint32_t slot = b.slot_of_vertex[b.first_vertex + i];
  t.vals[slot] = idx;
  // refs start at 0 and the voxel's first visit in this block increments both (:150-151, :160-161)
  s.active[idx] = 1;
  s.inactive[idx] = 1;
  s.live[idx] = 1;
  s.stamps[idx] = b.stamp_ns;
}

__global__ void vccCommitNewKernel(const uint32_t* __restrict__ tile_offsets, uint32_t n_tiles,
                                   unsigned long long* __restrict__ counters) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long total = tile_offsets[n_tiles];
    counters[kVccMaxIndex] += total;
    counters[kVccLive] += total;
  }
}

// remapping[block][i] = index; all_vertices_[index] = p by the voxel's last occurrence (:154-156)
__global__ void __launch_bounds__(256)
vccWriteKernel(const VccTable t, const VccIndexState s, const VccBlockArgs b) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  const uint64_t v = b.first_vertex + i;
  const uint32_t slot = b.slot_of_vertex[v];
  if (slot == kVccNoIndex) {
    b.remap_out[v] = ~0ull;
    return;
  }
  const uint32_t idx = t.vals[slot];
  b.remap_out[v] = idx;
  if (static_cast<uint32_t>(t.seq_max[slot] & 0xFFFFFFFFull) == i) {
    const uint32_t rgba = b.rgba ? b.rgba[v] : 0u;
    s.vertices[idx] = make_float4(b.xyz[3 * v + 0], b.xyz[3 * v + 1], b.xyz[3 * v + 2],
                                  __uint_as_float(rgba));
  }
}

// K8: release the references a block held on its previous voxel set
__global__ void __launch_bounds__(256)
vccReleaseKernel(const VccTable t, const VccIndexState s, const uint64_t* __restrict__ prev_list,
                 uint64_t prev_count, int archive, unsigned long long* __restrict__ counters) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= prev_count) return;
  const uint64_t key = prev_list[i];
  uint32_t slot;
  if (!findSlot(t, key, &slot)) {
    atomicOr(counters + kVccError, 4ull);
    return;
  }
  const uint32_t idx = t.vals[slot];
  // every previous voxel is distinct, so nobody else touches this index in this launch
  const int active = --s.active[idx];
  if (archive) {
    // pruneMeshBlocks (:76-88): only the active count drops
    if (active != 0) return;
  } else {
    const int inactive = --s.inactive[idx];
    if (inactive == 0) {
      // no block references the vertex any more: invalidate the slot (:185-193)
      t.keys[slot] = kVccTomb;
      s.live[idx] = 0;
      s.vertices[idx] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0xFF000000u));
      const unsigned long long at = atomicAdd(counters + kVccInvalid, 1ull);
      s.invalid[at] = idx;
      atomicAdd(counters + kVccTombstones, 1ull);
      atomicAdd(counters + kVccLive, static_cast<unsigned long long>(-1ll));
      return;
    }
    if (active != 0) return;
  }
  // archive: drop the voxel -> index association so a revisit creates a new vertex
  t.keys[slot] = kVccTomb;
  s.live[idx] = 0;
  atomicAdd(counters + kVccTombstones, 1ull);
  atomicAdd(counters + kVccLive, static_cast<unsigned long long>(-1ll));
}

__global__ void vccInitTableKernel(const VccTable t) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i > t.mask) return;
  t.keys[i] = kVccEmpty;
  t.vals[i] = kVccNoIndex;
  t.seq_min[i] = ~0ull;
  t.seq_max[i] = 0ull;
}

__global__ void vccRehashKernel(const VccTable from, const VccTable to,
                                unsigned long long* __restrict__ counters) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i > from.mask) return;
  const uint64_t key = from.keys[i];
  if (key == kVccEmpty || key == kVccTomb) return;
  uint32_t slot;
  bool inserted;
  if (!findOrInsert(to, key, &slot, &inserted)) {
    atomicOr(counters + kVccError, 2ull);
    return;
  }
  to.vals[slot] = from.vals[i];
}

__global__ void __launch_bounds__(kTile)
countLiveKernel(const uint8_t* __restrict__ live, uint64_t n, uint32_t* __restrict__ tile_counts) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kTile + threadIdx.x;
  const bool c = i < n && live[i];
  const int total = __syncthreads_count(c);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = static_cast<uint32_t>(total);
}

__global__ void __launch_bounds__(kTile)
writeLiveKernel(const uint8_t* __restrict__ live, uint64_t n,
                const uint32_t* __restrict__ tile_offsets, uint64_t* __restrict__ out) {
  __shared__ uint32_t warp_sums[kTile / 32];
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kTile + threadIdx.x;
  const bool c = i < n && live[i];
  const unsigned ballot = __ballot_sync(0xFFFFFFFFu, c);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = __popc(ballot);
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t w = warp_sums[threadIdx.x];
    uint32_t wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (threadIdx.x >= o) wi += up;
    }
    warp_sums[threadIdx.x] = wi - w;
  }
  __syncthreads();
  if (c) out[tile_offsets[blockIdx.x] + warp_sums[warp] + __popc(ballot & ((1u << lane) - 1))] = i;
}

__global__ void copyTotalKernel(const uint32_t* __restrict__ tile_offsets, uint32_t n_tiles,
                                unsigned long long* __restrict__ out_count) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out_count = tile_offsets[n_tiles];
}

}  // namespace

cudaError_t launchVccInsert(const VccTable& t, const VccIndexState& s, const VccBlockArgs& b,
                            unsigned long long* counters, cudaStream_t stream) {
  if (b.n == 0) return cudaSuccess;
  vccInsertKernel<<<(b.n + 255) / 256, 256, 0, stream>>>(t, s, b, counters);
  return cudaGetLastError();
}

uint64_t vccTileScratchEntries(uint64_t n) { return (n + kTile - 1) / kTile + 1; }

cudaError_t launchVccAssign(const VccTable& t, const VccIndexState& s, const VccBlockArgs& b,
                            uint32_t* tiles, unsigned long long* counters, cudaStream_t stream) {
  if (b.n == 0) return cudaSuccess;
  const uint32_t n_tiles = (b.n + kTile - 1) / kTile;
  vccCountCreatorsKernel<<<n_tiles, kTile, 0, stream>>>(t, b, tiles);
  vccScanTilesKernel<<<1, kTile, 0, stream>>>(tiles, n_tiles, counters);
  vccAssignKernel<<<n_tiles, kTile, 0, stream>>>(t, s, b, tiles, counters);
  vccCommitNewKernel<<<1, 32, 0, stream>>>(tiles, n_tiles, counters);
  return cudaGetLastError();
}

cudaError_t launchVccWrite(const VccTable& t, const VccIndexState& s, const VccBlockArgs& b,
                           cudaStream_t stream) {
  if (b.n == 0) return cudaSuccess;
  vccWriteKernel<<<(b.n + 255) / 256, 256, 0, stream>>>(t, s, b);
  return cudaGetLastError();
}

cudaError_t launchVccRelease(const VccTable& t, const VccIndexState& s, const uint64_t* prev_list,
                             uint64_t prev_count, bool archive, unsigned long long* counters,
                             cudaStream_t stream) {
  if (prev_count == 0) return cudaSuccess;
  vccReleaseKernel<<<static_cast<unsigned>((prev_count + 255) / 256), 256, 0, stream>>>(
      t, s, prev_list, prev_count, archive ? 1 : 0, counters);
  return cudaGetLastError();
}

cudaError_t launchVccRehash(const VccTable& from, const VccTable& to,
                            unsigned long long* counters, cudaStream_t stream) {
  const uint64_t n = static_cast<uint64_t>(from.mask) + 1;
  vccRehashKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(from, to, counters);
  return cudaGetLastError();
}

cudaError_t launchVccInitTable(const VccTable& t, cudaStream_t stream) {
  const uint64_t n = static_cast<uint64_t>(t.mask) + 1;
  vccInitTableKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(t);
  return cudaGetLastError();
}

cudaError_t launchVccCompactLive(const uint8_t* live, uint64_t n, uint64_t* out, uint32_t* tiles,
                                 unsigned long long* out_count, cudaStream_t stream) {
  const uint32_t n_tiles = static_cast<uint32_t>((n + kTile - 1) / kTile);
  if (n_tiles > 0) {
    countLiveKernel<<<n_tiles, kTile, 0, stream>>>(live, n, tiles);
  }
  vccScanTilesKernel<<<1, kTile, 0, stream>>>(tiles, n_tiles, nullptr);
  if (n_tiles > 0) {
    writeLiveKernel<<<n_tiles, kTile, 0, stream>>>(live, n, tiles, out);
  }
  copyTotalKernel<<<1, 32, 0, stream>>>(tiles, n_tiles, out_count);
  return cudaGetLastError();
}

}  // namespace pgmo

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadVccKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, vccInsertKernel);
  cudaFuncGetAttributes(&attr, vccCountCreatorsKernel);
  cudaFuncGetAttributes(&attr, vccScanTilesKernel);
  cudaFuncGetAttributes(&attr, vccAssignKernel);
  cudaFuncGetAttributes(&attr, vccCommitNewKernel);
  cudaFuncGetAttributes(&attr, vccWriteKernel);
  cudaFuncGetAttributes(&attr, vccReleaseKernel);
  cudaFuncGetAttributes(&attr, vccInitTableKernel);
  cudaFuncGetAttributes(&attr, vccRehashKernel);
  cudaFuncGetAttributes(&attr, countLiveKernel);
  cudaFuncGetAttributes(&attr, writeLiveKernel);
  cudaFuncGetAttributes(&attr, copyTotalKernel);
  cudaGetLastError();
}
}  // namespace pgmo
