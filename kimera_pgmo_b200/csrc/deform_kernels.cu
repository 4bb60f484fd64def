// Deformation kernels for sm_100a.  Replaces the per-vertex body of deformation::deformPoints
// (kimera_pgmo/include/kimera_pgmo/mesh_deformation.h:161-212) and interpPoint
// (kimera_pgmo/src/mesh_deformation.cpp:73-106) including PCL's nearestKSearch semantics.
//
// Arithmetic contract (bit-exact against the oracle):
//   d^2      float, dx*dx + (dy*dy + dz*dz), every op rounded separately (no FMA)
//   sqrt     float, round-to-nearest
//   weights  double: 1 - double(sqrtf(d2_j)) / double(sqrtf(d2_last))
//   blend    double: acc += w * ((R0*p0 + R1*p1) + R2*p2 + t), acc /= sum(w), cast to float
// All of it is written with explicit round-to-nearest intrinsics so that nvcc cannot contract
// a*b+c into an FMA (which would change the last bit).
#include "deform_kernels.cuh"

#include <cfloat>

#include "deform_device.cuh"
#include "stdsort_restated.cuh"

namespace pgmo {

namespace {

using namespace dev;

// ------------------------------------------------------------------------------------------
// K1
__global__ void stampStretchKernel(const uint64_t* __restrict__ stamps,
                                   const uint64_t* __restrict__ subset, uint64_t n,
                                   ulonglong2* __restrict__ out, uint64_t capacity,
                                   unsigned long long* __restrict__ count) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t p = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n;
       p += stride) {
    const uint64_t s = subset ? stamps[subset[p]] : stamps[p];
    bool boundary = (p == 0);
    if (p > 0) {
      const uint64_t prev = subset ? stamps[subset[p - 1]] : stamps[p - 1];
      boundary = prev != s;
    }
    if (boundary) {
      const unsigned long long slot = atomicAdd(count, 1ull);
      if (slot < capacity) out[slot] = make_ulonglong2(p, s);
      // order-independent fingerprint of the boundary set (count[1]): lets the host skip the
      // download + sort when the stretches are the ones it already holds
      unsigned long long m = (p + 0x9e3779b97f4a7c15ull) * 0xbf58476d1ce4e5b9ull;
      m ^= m >> 29;
      m = (m ^ s) * 0x94d049bb133111ebull;
      m ^= m >> 32;
      atomicAdd(count + 1, m);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K3: brute force over the run's window [lo, hi), exact top-(K+1) in registers.
template <int KCAP>
__global__ void __launch_bounds__(256)
deformKnnBlendKernel(const DeformArgs a, const int precision) {
  const uint64_t p = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= a.n_local) return;
  const uint32_t r = findRun(a.run_start, a.n_runs, a.first_position + p);
  const uint4 info = __ldg(a.run_info + r);
  if (info.w) {
    a.status[p] = kStatusSkipped;
    return;
  }
  const int K = static_cast<int>(info.z) + 2;
  if (K + 1 > KCAP) {
    sendToResolver(a, p);
    return;
  }
  const uint64_t ii = a.subset ? a.subset[p] : p;
  const float qx = a.x[ii], qy = a.y[ii], qz = a.z[ii];

  float best_d[KCAP];
  int best_i[KCAP];
#pragma unroll
  for (int i = 0; i < KCAP; ++i) {
    best_d[i] = FLT_MAX;
    best_i[i] = -1;
  }
  int n_alive = 0;
  const int run = static_cast<int>(r);
  for (uint32_t j = info.x; j < info.y; ++j) {
    const float4 c = __ldg(a.ctrl_f4 + j);
    if (__float_as_int(c.w) <= run) continue;  // voxel deleted before this run, or invalid
    ++n_alive;
    const float d2 = dist2Exact(c.x, c.y, c.z, qx, qy, qz);
    if (d2 < best_d[KCAP - 1] || best_i[KCAP - 1] < 0) {
      best_d[KCAP - 1] = d2;
      best_i[KCAP - 1] = static_cast<int>(j);
#pragma unroll
      for (int i = KCAP - 1; i > 0; --i) {
        // strict "<" (or an empty slot ahead): equal distances keep ascending control index
        if (best_i[i - 1] < 0 || best_d[i] < best_d[i - 1]) {
          const float td = best_d[i];
          best_d[i] = best_d[i - 1];
          best_d[i - 1] = td;
          const int ti = best_i[i];
          best_i[i] = best_i[i - 1];
          best_i[i - 1] = ti;
        }
      }
    }
  }
  const int n = n_alive < K ? n_alive : K;             // neighbours nearestKSearch returns
  const int n_check = n_alive < K + 1 ? n_alive : K + 1;  // plus the runner-up for the boundary
  bool tie = false;
#pragma unroll
  for (int i = 0; i + 1 < KCAP; ++i) {
    if (i + 1 < n_check && best_d[i] == best_d[i + 1]) tie = true;
  }
  if (tie) {
    sendToResolver(a, p);
    return;
  }
  writeResultReg<KCAP>(a, p, ii, precision, qx, qy, qz, best_d, best_i, n);
  a.status[p] = kStatusDone;
}

// ------------------------------------------------------------------------------------------
// K4: exact replay of PCL's getKNearestNeighborRecursive for one flagged vertex per warp.
// The octree of the run is implicit: node (level L, key prefix P) exists iff an alive control
// point's leaf key has prefix P; children are visited by ascending float distance to the voxel
// centre, ties by descending child index (std::sort of 8 entries is an insertion sort, stable),
// the loop over children ends at the first child failing the pruning bound.
struct DfsLevel {
  float dist[8];
  uint8_t child[8];
  int cnt;
  int pos;
  uint32_t key[3];
};

constexpr int kResolveWarpsPerBlock = 4;
constexpr int kResolveStage = 320;  // alive points of the vertex's run staged per warp

struct ResolveStage {
  float4 pts[kResolveStage];    // x, y, z, control index (int bits), ascending control index
  uint32_t key[kResolveStage][3];  // leaf key in the run's octree frame
};

__device__ __forceinline__ bool aliveInRun(const float4& c, int run) {
  return __float_as_int(c.w) > run;
}

__device__ void resolveOne(const DeformArgs& a, const ResolveArgs& ra, DfsLevel* levels,
                           ResolveStage& stage, CandEntry* cand, uint64_t p, int precision,
                           int lane) {
  const uint32_t r = findRun(a.run_start, a.n_runs, a.first_position + p);
  const uint4 info = a.run_info[r];
  const RunFrameDev frame = a.run_frame[r];
  const int run = static_cast<int>(r);
  const uint64_t ii = a.subset ? a.subset[p] : p;
  const float qx = a.x[ii], qy = a.y[ii], qz = a.z[ii];
  const uint32_t K = info.z + 2;
  const int depth = static_cast<int>(frame.depth);
  const unsigned full = 0xFFFFFFFFu;

  // stage the run's alive points (ascending control index, like the PCL leaves hold them) with
  // their leaf keys: every later pass of the DFS then runs out of shared memory
  int n_staged = 0;
  bool staged = true;
  for (uint32_t base = info.x; base < info.y; base += 32) {
    const uint32_t j = base + lane;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    bool alive = false;
    if (j < info.y) {
      c = __ldg(a.ctrl_f4 + j);
      alive = aliveInRun(c, run);
    }
    const unsigned m = __ballot_sync(full, alive);
    const int at = n_staged + __popc(m & ((1u << lane) - 1u));
    if (alive && at < kResolveStage) {
      const int4 sk = __ldg(a.ctrl_key + j);
      stage.pts[at] = make_float4(c.x, c.y, c.z, __int_as_float(static_cast<int>(j)));
      stage.key[at][0] = static_cast<uint32_t>(sk.x + frame.shift[0]);
      stage.key[at][1] = static_cast<uint32_t>(sk.y + frame.shift[1]);
      stage.key[at][2] = static_cast<uint32_t>(sk.z + frame.shift[2]);
    }
    n_staged += __popc(m);
  }
  if (n_staged > kResolveStage) staged = false;  // large alive set: scan the window in place
  __syncwarp();

  uint32_t n_cand = 0;
  double best = DBL_MAX;

  int level = 0;
  if (lane == 0) {
    levels[0].key[0] = levels[0].key[1] = levels[0].key[2] = 0;
  }
  __syncwarp();
  bool entering = true;
  while (level >= 0) {
    DfsLevel& L = levels[level];
    if (entering) {
      // which of the 8 children of (level, key) hold an alive control point?
      const uint32_t kx = L.key[0], ky = L.key[1], kz = L.key[2];
      const int up = depth - level;  // leaf key >> up == node key
      unsigned mask = 0;
      if (staged) {
        for (int i = lane; i < n_staged; i += 32) {
          const uint32_t x = stage.key[i][0], y = stage.key[i][1], z = stage.key[i][2];
          const bool in_node = up >= 32 || ((x >> up) == kx && (y >> up) == ky && (z >> up) == kz);
          if (!in_node) continue;
          const int bit = up - 1;
          mask |= 1u << ((((x >> bit) & 1) << 2) | (((y >> bit) & 1) << 1) | ((z >> bit) & 1));
        }
      }
      for (uint32_t j = info.x + lane; !staged && j < info.y; j += 32) {
        const float4 c = __ldg(a.ctrl_f4 + j);
        if (!aliveInRun(c, run)) continue;
        const int4 sk = __ldg(a.ctrl_key + j);
        const uint32_t x = static_cast<uint32_t>(sk.x + frame.shift[0]);
        const uint32_t y = static_cast<uint32_t>(sk.y + frame.shift[1]);
        const uint32_t z = static_cast<uint32_t>(sk.z + frame.shift[2]);
        const bool in_node = up >= 32 || ((x >> up) == kx && (y >> up) == ky && (z >> up) == kz);
        if (!in_node) continue;
        const int bit = up - 1;
        const int child = (((x >> bit) & 1) << 2) | (((y >> bit) & 1) << 1) | ((z >> bit) & 1);
        mask |= 1u << child;
      }
      mask = __reduce_or_sync(full, mask);
      {
        // lanes 0..7 each take one child: voxel centre (genVoxelCenterFromOctreeKey: double
        // arithmetic, float result), float distance, then its rank in the visiting order —
        // ascending distance, ties -> higher child index first
        const int tree_depth = level + 1;
        const double side = a.resolution * static_cast<double>(1u << (depth - tree_depth));
        const int c = lane & 7;
        const bool exists = lane < 8 && (mask & (1u << c));
        const uint32_t ccx = (kx << 1) + ((c >> 2) & 1);
        const uint32_t ccy = (ky << 1) + ((c >> 1) & 1);
        const uint32_t ccz = (kz << 1) + (c & 1);
        const float vx = __double2float_rn(
            __dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(ccx), 0.5), side), frame.min_xyz[0]));
        const float vy = __double2float_rn(
            __dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(ccy), 0.5), side), frame.min_xyz[1]));
        const float vz = __double2float_rn(
            __dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(ccz), 0.5), side), frame.min_xyz[2]));
        const float d = dist2Exact(vx, vy, vz, qx, qy, qz);
        int rank = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const float od = __shfl_sync(full, d, o);
          const bool other = (mask >> o) & 1u;
          rank += (other && (od < d || (od == d && o > c))) ? 1 : 0;
        }
        if (exists) {
          L.dist[rank] = d;
          L.child[rank] = static_cast<uint8_t>(c);
        }
        if (lane == 0) {
          L.cnt = __popc(mask & 0xFFu);
          L.pos = 0;
        }
      }
      __syncwarp();
      entering = false;
    }
    if (L.pos >= L.cnt) {
      --level;
      if (level >= 0) {
        __syncwarp();
        if (lane == 0) levels[level].pos += 1;
        __syncwarp();
      }
      continue;
    }
    const int tree_depth = level + 1;
    double side = a.resolution * static_cast<double>(1u << (depth - tree_depth));
    side = __dmul_rn(side, side);
    const double diam2 = __dmul_rn(side, 3.0);
    const float cd = L.dist[L.pos];
    const double bound =
        __dadd_rn(__dadd_rn(best, __ddiv_rn(diam2, 4.0)), __dsqrt_rn(__dmul_rn(best, diam2)));
    if (!(static_cast<double>(cd) < bound)) {
      // `while` condition failed: this node is finished
      --level;
      if (level >= 0) {
        __syncwarp();
        if (lane == 0) levels[level].pos += 1;
        __syncwarp();
      }
      continue;
    }
    const int c = L.child[L.pos];
    const uint32_t cx = (L.key[0] << 1) + ((c >> 2) & 1);
    const uint32_t cy = (L.key[1] << 1) + ((c >> 1) & 1);
    const uint32_t cz = (L.key[2] << 1) + (c & 1);
    if (tree_depth < depth) {
      if (lane == 0) {
        levels[level + 1].key[0] = cx;
        levels[level + 1].key[1] = cy;
        levels[level + 1].key[2] = cz;
      }
      __syncwarp();
      ++level;
      entering = true;
      continue;
    }
    // leaf: append its points (ascending index) that beat the current K-th distance
    for (int base = 0; staged && base < n_staged; base += 32) {
      const int i = base + lane;
      bool accept = false;
      float d2 = 0.f;
      int j = 0;
      if (i < n_staged && stage.key[i][0] == cx && stage.key[i][1] == cy && stage.key[i][2] == cz) {
        const float4 cp = stage.pts[i];
        j = __float_as_int(cp.w);
        d2 = dist2Exact(cp.x, cp.y, cp.z, qx, qy, qz);
        accept = static_cast<double>(d2) < best;
      }
      const unsigned m = __ballot_sync(full, accept);
      if (accept) {
        const uint32_t slot = n_cand + __popc(m & ((1u << lane) - 1));
        if (slot < ra.cand_cap) {
          cand[slot].dist = d2;
          cand[slot].idx = j;
        }
      }
      n_cand += __popc(m);
    }
    for (uint32_t base = info.x; !staged && base < info.y; base += 32) {
      const uint32_t j = base + lane;
      bool accept = false;
      float d2 = 0.f;
      if (j < info.y) {
        const float4 cp = __ldg(a.ctrl_f4 + j);
        if (aliveInRun(cp, run)) {
          const int4 sk = __ldg(a.ctrl_key + j);
          if (static_cast<uint32_t>(sk.x + frame.shift[0]) == cx &&
              static_cast<uint32_t>(sk.y + frame.shift[1]) == cy &&
              static_cast<uint32_t>(sk.z + frame.shift[2]) == cz) {
            d2 = dist2Exact(cp.x, cp.y, cp.z, qx, qy, qz);
            accept = static_cast<double>(d2) < best;
          }
        }
      }
      const unsigned m = __ballot_sync(full, accept);
      if (accept) {
        const uint32_t slot = n_cand + __popc(m & ((1u << lane) - 1));
        if (slot < ra.cand_cap) {
          cand[slot].dist = d2;
          cand[slot].idx = static_cast<int>(j);
        }
      }
      n_cand += __popc(m);
    }
    if (n_cand > ra.cand_cap) n_cand = ra.cand_cap;  // cannot happen: cap = K + max leaf size
    __syncwarp();
    if (lane == 0) {
      stdsort::sort(cand, cand + n_cand,
                    [](const CandEntry& l, const CandEntry& r2) { return l.dist < r2.dist; });
      L.pos += 1;
    }
    __syncwarp();
    if (n_cand > K) n_cand = K;
    if (n_cand == K) best = static_cast<double>(cand[K - 1].dist);
  }

  __syncwarp();
  if (lane == 0) {
    if (n_cand == 0) {
      a.status[p] = kStatusSkipped;  // unreachable: the schedule rejects empty trees
    } else {
      // blend straight from the candidate list
      constexpr int kChunk = 16;
      if (static_cast<int>(n_cand) <= kChunk) {
        float d2[kChunk];
        int idx[kChunk];
        for (uint32_t i = 0; i < n_cand; ++i) {
          d2[i] = cand[i].dist;
          idx[i] = cand[i].idx;
        }
        writeResult(a, p, ii, precision, qx, qy, qz, d2, idx, static_cast<int>(n_cand));
      } else {
        // large-k path: split the struct array into the two views writeResult expects, in place
        // (dist and idx share a slab: copy idx out behind the entries)
        float* d2 = reinterpret_cast<float*>(cand + ra.cand_cap);
        int* idx = reinterpret_cast<int*>(d2 + ra.cand_cap);
        for (uint32_t i = 0; i < n_cand; ++i) {
          d2[i] = cand[i].dist;
          idx[i] = cand[i].idx;
        }
        writeResult(a, p, ii, precision, qx, qy, qz, d2, idx, static_cast<int>(n_cand));
      }
      a.status[p] = kStatusDone;
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kResolveWarpsPerBlock * 32)
deformResolveKernel(const DeformArgs a, const ResolveArgs ra, const int precision) {
  __shared__ DfsLevel s_levels[kResolveWarpsPerBlock][kResolverMaxDepth + 1];
  __shared__ ResolveStage s_stage[kResolveWarpsPerBlock];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const uint32_t warp = blockIdx.x * kResolveWarpsPerBlock + warp_in_block;
  // slab per warp: cand_cap entries + room for the split views used by large k
  CandEntry* cand = ra.scratch + static_cast<uint64_t>(warp) * ra.cand_cap * 2;
  // work list filled by the K3 kernels: entries [counters[1], counters[0]) are new
  const unsigned long long begin = a.counters[1], end = a.counters[0];
  for (unsigned long long i = begin + warp; i < end; i += ra.n_warps) {
    resolveOne(a, ra, s_levels[warp_in_block], s_stage[warp_in_block], cand, a.tie_list[i], precision,
               lane);
  }
}

__global__ void deformResolveDoneKernel(unsigned long long* counters) {
  if (threadIdx.x == 0 && blockIdx.x == 0) counters[1] = counters[0];
}

// ------------------------------------------------------------------------------------------
// K5
__global__ void predeformKernel(uint64_t n_v, uint64_t start_idx, const float* __restrict__ x,
                                const float* __restrict__ y, const float* __restrict__ z,
                                const int32_t* __restrict__ graph_indices, uint64_t n_c,
                                const uint8_t* __restrict__ valid,
                                const double* __restrict__ ctrl_pos,
                                const double* __restrict__ poses, float* __restrict__ ox,
                                float* __restrict__ oy, float* __restrict__ oz,
                                uint8_t* __restrict__ needs_deform) {
  const uint64_t i = start_idx + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_v) return;
  const int32_t index = graph_indices[i];
  // values.exists(Symbol(prefix, index)): an index past the table has no pose
  if (index < 0 || static_cast<uint64_t>(index) >= n_c || (valid && !valid[index])) {
    needs_deform[i] = 1;
    return;
  }
  needs_deform[i] = 0;
  const double* g = ctrl_pos + 3ull * index;
  const double* T = poses + 12ull * index;
  const double p0 = __dsub_rn(static_cast<double>(x[i]), g[0]);
  const double p1 = __dsub_rn(static_cast<double>(y[i]), g[1]);
  const double p2 = __dsub_rn(static_cast<double>(z[i]), g[2]);
  const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(T[0], p0), __dmul_rn(T[1], p1)), __dmul_rn(T[2], p2));
  const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(T[3], p0), __dmul_rn(T[4], p1)), __dmul_rn(T[5], p2));
  const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(T[6], p0), __dmul_rn(T[7], p1)), __dmul_rn(T[8], p2));
  ox[i] = __double2float_rn(__dadd_rn(r0, T[9]));
  oy[i] = __double2float_rn(__dadd_rn(r1, T[10]));
  oz[i] = __double2float_rn(__dadd_rn(r2, T[11]));
}

__global__ void buildControlTablesKernel(uint64_t n_c, const double* __restrict__ ctrl_pos,
                                         const double* __restrict__ poses,
                                         float4* __restrict__ ctrl_f4,
                                         float* __restrict__ pose_f32, int positions_changed) {
  const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_c) return;
  const double gx = ctrl_pos[3 * j], gy = ctrl_pos[3 * j + 1], gz = ctrl_pos[3 * j + 2];
  if (positions_changed) {
    // .w (the schedule's del_run) is rewritten by applyScheduleKernel after the next build
    ctrl_f4[j] = make_float4(__double2float_rn(gx), __double2float_rn(gy), __double2float_rn(gz),
                             __int_as_float(0));
  }
  float* o = pose_f32 + 16 * j;
  for (int i = 0; i < 12; ++i) o[i] = poses ? __double2float_rn(poses[12 * j + i]) : 0.f;
  o[12] = __double2float_rn(gx);
  o[13] = __double2float_rn(gy);
  o[14] = __double2float_rn(gz);
  o[15] = 0.f;
}

__global__ void applyScheduleKernel(uint64_t n_c, const int32_t* __restrict__ del_run,
                                    const int32_t* __restrict__ stable_key3,
                                    float4* __restrict__ ctrl_f4, int4* __restrict__ ctrl_key) {
  const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_c) return;
  ctrl_f4[j].w = __int_as_float(del_run[j]);
  ctrl_key[j] = make_int4(stable_key3[3 * j], stable_key3[3 * j + 1], stable_key3[3 * j + 2], 0);
}

}  // namespace

cudaError_t launchStampStretches(const uint64_t* stamps, const uint64_t* subset, uint64_t n,
                                 ulonglong2* out, uint64_t capacity,
                                 unsigned long long* count, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned threads = 256;
  uint64_t blocks = (n + threads - 1) / threads;
  if (blocks > 148ull * 16) blocks = 148ull * 16;
  stampStretchKernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(stamps, subset, n, out,
                                                                           capacity, count);
  return cudaGetLastError();
}

cudaError_t launchDeformKnnBlend(const DeformArgs& args, int precision, uint32_t max_k,
                                 cudaStream_t stream) {
  if (args.n_local == 0) return cudaSuccess;
  const unsigned threads = 256;
  const unsigned blocks = static_cast<unsigned>((args.n_local + threads - 1) / threads);
  if (max_k + 3 <= 8) {
    deformKnnBlendKernel<8><<<blocks, threads, 0, stream>>>(args, precision);
  } else {
    deformKnnBlendKernel<16><<<blocks, threads, 0, stream>>>(args, precision);
  }
  return cudaGetLastError();
}

uint32_t resolverResidentWarps(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return static_cast<uint32_t>(sms) * 4 * kResolveWarpsPerBlock;
}

cudaError_t launchDeformResolve(const DeformArgs& args, const ResolveArgs& rargs, int precision,
                                cudaStream_t stream) {
  if (args.n_local == 0) return cudaSuccess;
  const unsigned blocks = rargs.n_warps / kResolveWarpsPerBlock;
  deformResolveKernel<<<blocks, kResolveWarpsPerBlock * 32, 0, stream>>>(args, rargs, precision);
  deformResolveDoneKernel<<<1, 32, 0, stream>>>(args.counters);
  return cudaGetLastError();
}

cudaError_t launchPredeform(uint64_t n_v, uint64_t start_idx, const float* x, const float* y,
                            const float* z, const int32_t* graph_indices, uint64_t n_c,
                            const float4*, const uint8_t* valid, const double* ctrl_pos,
                            const double* poses, float* ox, float* oy, float* oz,
                            uint8_t* needs_deform, int*, cudaStream_t stream) {
  if (start_idx >= n_v) return cudaSuccess;
  const uint64_t n = n_v - start_idx;
  const unsigned threads = 256;
  predeformKernel<<<static_cast<unsigned>((n + threads - 1) / threads), threads, 0, stream>>>(
      n_v, start_idx, x, y, z, graph_indices, n_c, valid, ctrl_pos, poses, ox, oy, oz, needs_deform);
  return cudaGetLastError();
}

cudaError_t launchBuildControlTables(uint64_t n_c, const double* ctrl_pos, const double* poses,
                                     float4* ctrl_f4, float* ctrl_pose_f32, bool positions_changed,
                                     cudaStream_t stream) {
  if (n_c == 0) return cudaSuccess;
  const unsigned threads = 256;
  buildControlTablesKernel<<<static_cast<unsigned>((n_c + threads - 1) / threads), threads, 0,
                             stream>>>(n_c, ctrl_pos, poses, ctrl_f4, ctrl_pose_f32,
                                       positions_changed ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launchApplySchedule(uint64_t n_c, const int32_t* del_run, const int32_t* stable_key3,
                                float4* ctrl_f4, int4* ctrl_key, cudaStream_t stream) {
  if (n_c == 0) return cudaSuccess;
  const unsigned threads = 256;
  applyScheduleKernel<<<static_cast<unsigned>((n_c + threads - 1) / threads), threads, 0, stream>>>(
      n_c, del_run, stable_key3, ctrl_f4, ctrl_key);
  return cudaGetLastError();
}

}  // namespace pgmo
