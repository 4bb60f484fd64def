// K4 device code: PCL's nearestKSearch result, tie order included, for one vertex by one warp
// (deform_kernels.cu: deformResolveKernel).
#pragma once

#include <cfloat>

#include "deform_device.cuh"
#include "stdsort_restated.cuh"

namespace pgmo {
namespace {

using namespace dev;

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
  float d2[kResolveStage];      // exact d^2 to the vertex (order-of-visit path)
};

// Which of two alive points does PCL's depth-first search reach first?  Same leaf: the leaf
// holds its points in insertion order (ascending control index = staged order).  Otherwise the
// two root-to-leaf paths split at one node, whose children are visited by ascending float
// distance of the voxel centre to the query, equal distances by descending child index
// (getKNearestNeighborRecursive; the first child's whole subtree precedes the second's).
__device__ __forceinline__ bool visitedBefore(const ResolveStage& st, int ia, int ib,
                                              const RunFrameDev& frame, double resolution, float qx,
                                              float qy, float qz) {
  const uint32_t ax = st.key[ia][0], ay = st.key[ia][1], az = st.key[ia][2];
  const uint32_t bx = st.key[ib][0], by = st.key[ib][1], bz = st.key[ib][2];
  const uint32_t x = (ax ^ bx) | (ay ^ by) | (az ^ bz);
  if (x == 0) return ia < ib;
  const int m = 31 - __clz(x);  // the children that hold a and b are the nodes of height m
  const double side = __dmul_rn(resolution, static_cast<double>(1u << m));
  // genVoxelCenterFromOctreeKey: double arithmetic, float result
  auto centreDist = [&](uint32_t kx, uint32_t ky, uint32_t kz) {
    const float vx = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(kx >> m), 0.5), side), frame.min_xyz[0]));
    const float vy = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(ky >> m), 0.5), side), frame.min_xyz[1]));
    const float vz = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(kz >> m), 0.5), side), frame.min_xyz[2]));
    return dist2Exact(vx, vy, vz, qx, qy, qz);
  };
  const float da = centreDist(ax, ay, az), db = centreDist(bx, by, bz);
  if (da != db) return da < db;
  const int ca = (((ax >> m) & 1) << 2) | (((ay >> m) & 1) << 1) | ((az >> m) & 1);
  const int cb = (((bx >> m) & 1) << 2) | (((by >> m) & 1) << 1) | ((bz >> m) & 1);
  return ca > cb;
}

__device__ __forceinline__ bool aliveInRun(const float4& c, int run) {
  return __float_as_int(c.w) > run;
}

// the resolver's neighbour list as an association record (what K3a writes for untied vertices)
__device__ __forceinline__ void cacheFromResolver(const DeformArgs&, const AssocCache& c,
                                                  uint32_t slot, uint32_t lo, const float* d2,
                                                  const int* idx, int n, int precision) {
  if (n - 1 > kCacheMaxNbr) return;  // the item is not cacheable (K3b skips it)
  const float d_max_f = __fsqrt_rn(d2[n - 1]);
  const bool const_w = constWeight(d2[0], d_max_f);
  const float inv_dmax = const_w ? 0.0f : __frcp_rn(d_max_f);
  for (int j = 0; j < n - 1; ++j) {
    cacheStoreNbr(c, slot, j, (static_cast<uint32_t>(idx[j]) - lo) |
                                  (static_cast<uint32_t>(__ldg(c.ctrl_class + idx[j]) & 7u) << kCacheClassShift));
    cacheStoreWeight(c, slot, j, precision, weightFp64(d2[j], d_max_f, const_w),
                     weightFp32(d2[j], inv_dmax, const_w));
  }
}

__device__ __noinline__ void resolveOne(const DeformArgs& a, const ResolveArgs& ra, DfsLevel* levels,
                           ResolveStage& stage, CandEntry* cand, uint64_t p, int precision,
                           int lane, const AssocCache* cache, uint32_t slot) {
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

  // Order-of-visit path.  The candidate vector never holds more than K + (points of one leaf)
  // entries; up to 16 of them std::sort is an insertion sort, i.e. stable, so equal distances
  // stay in the order they were pushed = the order the search visits them, and a tied point is
  // dropped at the K-th place exactly when it is visited later (it fails `d2 < smallest` once
  // the list is full, or loses the stable sort's truncation).  The pruning bound never cuts a
  // point closer than the current K-th distance, so the result is "the K smallest under
  // (d^2, visit order)" — computed directly, without replaying the traversal.
  const bool by_order = staged && K + ra.max_leaf_points <= 16 && K <= ra.cand_cap;
  if (by_order) {
    for (int i = lane; i < n_staged; i += 32) {
      const float4 cp = stage.pts[i];
      stage.d2[i] = dist2Exact(cp.x, cp.y, cp.z, qx, qy, qz);
    }
    __syncwarp();
    uint32_t taken = 0;  // bit t: this lane's point lane + 32 t is already selected
    const uint32_t want = static_cast<uint32_t>(n_staged) < K ? static_cast<uint32_t>(n_staged) : K;
    for (uint32_t rnk = 0; rnk < want; ++rnk) {
      int bi = -1;
      float bd = 0.f;
      for (int t = 0, i = lane; i < n_staged; ++t, i += 32) {
        if (taken & (1u << t)) continue;
        const float d = stage.d2[i];
        if (bi < 0 || d < bd || (d == bd && visitedBefore(stage, i, bi, frame, a.resolution, qx, qy, qz))) {
          bi = i;
          bd = d;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int oi = __shfl_xor_sync(full, bi, o);
        const float od = __shfl_xor_sync(full, bd, o);
        if (oi >= 0 && (bi < 0 || od < bd ||
                        (od == bd && visitedBefore(stage, oi, bi, frame, a.resolution, qx, qy, qz)))) {
          bi = oi;
          bd = od;
        }
      }
      if ((bi & 31) == lane) taken |= 1u << (bi >> 5);
      if (lane == 0) {
        cand[rnk].dist = bd;
        cand[rnk].idx = __float_as_int(stage.pts[bi].w);
      }
    }
    n_cand = want;
  }

  int level = 0;
  if (lane == 0) {
    levels[0].key[0] = levels[0].key[1] = levels[0].key[2] = 0;
  }
  __syncwarp();
  bool entering = true;
  if (by_order) level = -1;  // nothing left for the traversal
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
        if (cache && slot != kNoCacheSlot) {
          cacheFromResolver(a, *cache, slot, info.x, d2, idx, static_cast<int>(n_cand), precision);
        }
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
        if (cache && slot != kNoCacheSlot) {
          cacheFromResolver(a, *cache, slot, info.x, d2, idx, static_cast<int>(n_cand), precision);
        }
      }
      a.status[p] = kStatusDone;
    }
  }
  __syncwarp();
}

}  // namespace
}  // namespace pgmo
