// K3a, bucket-ordered candidate lists (sm_100a): the main knn kernel.
//
// deformation::interpPoint (kimera_pgmo/src/mesh_deformation.cpp:73-106) needs, per vertex, the
// K = k + 2 nearest alive control points in ascending float-d^2 order.  The profile of the
// grid-walk kernel (profiles/r01b_*) showed the search to be ALU-issue-bound: ~6000 warp
// instructions per 32 vertices, almost all of them 64-bit compare/select chains and cell-loop
// bookkeeping.  This kernel removes both:
//
//   * One CTA = one work item (<= 2048 consecutive positions of ONE run, i.e. one alive set of
//     a few hundred control points, staged in shared memory).  The vertices are split, by two
//     counting sorts in shared memory, into equal-count slabs along the longest axis of their
//     box and then into groups of 32 along the second longest, so that every warp pass is a
//     compact column B.
//   * Per group, the warp buckets ALL alive points by a lower bound of their squared distance
//     to B's bounding box (32 buckets of width r0^2/8) and walks the buckets outwards.  Lanes
//     keep their (K+1) smallest keys in registers; the walk stops as soon as every lane's
//     (K+1)-th key is below the lower bound of the next bucket, i.e. nothing unscanned can
//     enter (or tie with) anybody's list.  Every candidate read is one broadcast LDS.128.
//   * Keys are 32 bits: the exact float d^2 bit pattern with the low `ord_bits` mantissa bits
//     replaced by the candidate's position in the warp's list, so that the sorted insert is a
//     branch-free chain of integer min/max pairs (2 VIMNMX per slot instead of the 6-7
//     instructions of a 64-bit compare + payload select).  The winners' exact d^2 and control
//     index are recomputed from the list afterwards.  Two winners whose keys agree above the
//     ordinal bits are "near ties" (|delta d^2| < 2^ord_bits ulp, ~3e-4 of the vertices):
//     they are re-ordered on their exact d^2 in registers; exact ties (and near ties that
//     touch the last tracked slot) are handed to the PCL-order resolver (K4), which is exact.
//
//   * The kernel does not blend.  Neighbour set and weights depend on the vertex, the control
//     positions and the schedule, not on the poses: each vertex leaves an association record
//     (AssocCache, deform_kernels.cuh: vertex offset, k + 1 control indices relative to run.lo,
//     k + 1 weights) in HBM, in the order the search grouped the vertices, and the blend kernel
//     (deform_blend.cu, K3b) turns records + pose rows into points — after every search, and
//     alone when only the poses changed since.
//
// Search + blend are bit-identical to the brute-force and grid-walk kernels
// (tests/test_deform_gpu.py::test_search_modes_agree).  Runs with more than kSMaxPts alive
// points stay on the grid-walk kernel (deform_grid.cu), which blends by itself.
#include <cfloat>
#include <type_traits>

#include "deform_device.cuh"

namespace pgmo {

namespace {

using namespace dev;

constexpr int kSThreads = 256;
constexpr int kSWarps = kSThreads / 32;
constexpr int kSChunk = kDeformChunk;  // positions per work item
constexpr int kSMaxCells = 1024;       // sort bins of a pass
#ifndef PGMO_SLAB_SHIFT
#define PGMO_SLAB_SHIFT 8
#endif
constexpr int kSlabShift = PGMO_SLAB_SHIFT;  // vertices per slab (first split axis) = 1 << kSlabShift
constexpr int kSlab = 1 << kSlabShift;
constexpr int kSVBins = kSMaxCells / (kSChunk / kSlab);  // bins along the second axis inside a slab
static_assert((kSChunk / kSlab) * kSVBins <= kSMaxCells, "slab x bin keys");
#ifndef PGMO_K3_CTAS
#define PGMO_K3_CTAS 4
#endif
constexpr int kSmallMaxPts = 256;  // alive sets up to here use the leaner variant

template <int MAXPTS>
struct SortedSmem {
  float4 pts[MAXPTS + 4];                   // alive control points: x, y, z, index (int bits);
                                            // pts[n_alive] = a point at infinity (list padding)
  float vx[kSChunk], vy[kSChunk], vz[kSChunk];
  uint16_t perm[kSChunk];                   // sorted position -> vertex of the item
  union alignas(16) {
    struct {                                // vertex sort only
      uint32_t hist[kSMaxCells + 1];
      uint16_t perm1[kSChunk];              // order along the first split axis
    } sort;
    struct {
      uint16_t list[kSWarps][MAXPTS + 8];   // per warp: candidate slots in the order they are walked
    } w;
  } u;
  // per warp and slab: the alive slots, split into those that can matter for the slab ("near",
  // from the front) and the rest ("far", from the back)
  typename std::conditional<(MAXPTS <= 256), uint8_t, uint16_t>::type near[kSWarps][MAXPTS];
  float red[6][kSWarps];
  uint32_t warp_tot[kSWarps];
  float grid[8];                            // split axes: origin and 1 / bin width, axis ids
  unsigned long long fp_w[kSWarps];         // per warp: fingerprint of its vertices' bits
};

template <typename Smem>
__device__ __forceinline__ unsigned long long itemFingerprint(const Smem& S) {
  unsigned long long f = 0ull;
#pragma unroll
  for (int w = 0; w < kSWarps; ++w) f += S.fp_w[w];
  return f;
}

// Inserts two keys into the ascending list at once.  The i-th smallest of the union of two
// ascending sequences b and (lo, hi) is min(b[i], max(b[i-1], lo), max(b[i-2], hi)): per slot two
// VIMNMX and one three-input VIMNMX3 (sm_100a), no dependence between slots — 10 instructions
// per candidate instead of 13, and a dependency depth of 3 instead of 13.
template <int KCAP>
__device__ __forceinline__ void pairInsert(uint32_t (&best)[KCAP], uint32_t x, uint32_t y) {
  const uint32_t lo = min(x, y), hi = max(x, y);
#pragma unroll
  for (int i = KCAP - 1; i >= 2; --i) {
    best[i] = __vimin3_u32(best[i], max(best[i - 1], lo), max(best[i - 2], hi));
  }
  if (KCAP > 1) best[1] = __vimin3_u32(best[1], max(best[0], lo), hi);
  best[0] = min(best[0], lo);
}

// asynchronous global -> shared copies (LDGSTS): the staging of the alive points and their pose
// rows overlaps the vertex loads and the vertex sort; waited for right before the first group
template <int BYTES>
__device__ __forceinline__ void cpAsync(void* smem_dst, const void* gsrc) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES));
  }
}
__device__ __forceinline__ void cpAsyncWaitAll() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// warp-wide float min / max in one instruction (CREDUX.MIN/MAX.F32, sm_100a); NaN inputs are
// ignored like fminf / fmaxf do
__device__ __forceinline__ float warpMinF(float v) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warpMaxF(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// FULL: every vertex blends exactly KCAP - 2 neighbours (k + 1 with KCAP = k + 3) — the caller
// asked for one k (no mid-sweep change) and the item's alive set holds at least KCAP points.  The
// neighbour counts are then compile-time constants and every per-slot guard disappears; items
// that do not qualify (a handful of vertices at the start of a trajectory) go to the resolver.
template <int KCAP, int PREC, int MAXPTS, bool FULL>
__device__ __forceinline__ void searchItem(const DeformArgs& a, const GridArgs& ga,
                                           const AssocCache& cache, const WorkItem& item,
                                           uint32_t item_index, SortedSmem<MAXPTS>& S) {
  constexpr int kSMaxPts = MAXPTS;
  const uint32_t r = item.run;
  const uint4 info = __ldg(a.run_info + r);
  const uint64_t p_base = item.start;
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int count = static_cast<int>(item.count);
  if (info.w) return;  // skipped run: K3b writes the status bytes
  const RunGridDev g = ga.grids[r];
  const int n_alive = static_cast<int>(g.n_alive);
  if (n_alive > kSortedMaxPoints) return;  // the grid-walk kernel's item
  const int K = static_cast<int>(info.z) + 2;
  const bool cacheable = runCacheable(info);
  if (!cacheable || n_alive > kSMaxPts || K + 1 > KCAP || (FULL && (K + 1 != KCAP || n_alive < KCAP))) {
    // not this variant's item (a handful of vertices at the start of a trajectory, or a k this
    // build has no register path for): the resolver computes the neighbours, in item order
    unsigned long long fp = 0ull;
    for (int i = t; i < count; i += kSThreads) {
      const uint64_t p = p_base + i;
      const uint64_t ii = a.subset ? a.subset[p] : p;
      fp += vertexFingerprint(a.x[ii], a.y[ii], a.z[ii], static_cast<uint32_t>(i),
                              a.subset ? ii : 0ull);
      if (cacheable) cache.rec[p * 8] = static_cast<uint16_t>(i);
      sendToResolver(a, p, cacheable ? static_cast<uint32_t>(p) : kNoCacheSlot);
    }
    fp = warpSumU64(fp);
    if (lane == 0) S.fp_w[warp] = fp;
    __syncthreads();
    if (t == 0) {
      cache.item_fp[item_index] = itemFingerprint(S);
      cache.item_ok[item_index] = cacheable ? 1u : 0u;
    }
    return;
  }

  // Items of at most one group (the head vertex of a stamp stretch is a run of its own) skip
  // the vertex sort.
  const bool tiny = count <= 32;

  // ---- stage the run's alive points; load the item's vertices and their bounding box
  {
    const float4* src = ga.points + g.point_offset;
    for (int i = t; i < n_alive; i += kSThreads) cpAsync<16>(&S.pts[i], src + i);
    // list padding (see the walk): d^2 to this point is +inf for every finite query
    if (t == 0) S.pts[n_alive] = make_float4(3e30f, 3e30f, 3e30f, __int_as_float(-1));
  }
  unsigned long long fp = 0ull;
  float mn0 = FLT_MAX, mn1 = FLT_MAX, mn2 = FLT_MAX, mx0 = -FLT_MAX, mx1 = -FLT_MAX,
        mx2 = -FLT_MAX;
  for (int i = t; i < count; i += kSThreads) {
    const uint64_t ii = a.subset ? a.subset[p_base + i] : p_base + i;
    const float x = a.x[ii], y = a.y[ii], z = a.z[ii];
    fp += vertexFingerprint(x, y, z, static_cast<uint32_t>(i), a.subset ? ii : 0ull);
    S.vx[i] = x;
    S.vy[i] = y;
    S.vz[i] = z;
    mn0 = fminf(mn0, x);
    mx0 = fmaxf(mx0, x);
    mn1 = fminf(mn1, y);
    mx1 = fmaxf(mx1, y);
    mn2 = fminf(mn2, z);
    mx2 = fmaxf(mx2, z);
  }
  fp = warpSumU64(fp);
  if (lane == 0) S.fp_w[warp] = fp;
  if (tiny) {
    if (t < 32) S.perm[t] = static_cast<uint16_t>(t);
    cpAsyncWaitAll();
    __syncthreads();
  } else {
  mn0 = warpMinF(mn0);
  mn1 = warpMinF(mn1);
  mn2 = warpMinF(mn2);
  mx0 = warpMaxF(mx0);
  mx1 = warpMaxF(mx1);
  mx2 = warpMaxF(mx2);
  if (lane == 0) {
    S.red[0][warp] = mn0;
    S.red[1][warp] = mn1;
    S.red[2][warp] = mn2;
    S.red[3][warp] = mx0;
    S.red[4][warp] = mx1;
    S.red[5][warp] = mx2;
  }
  for (int i = t; i <= kSMaxCells; i += kSThreads) S.u.sort.hist[i] = 0;
  __syncthreads();
  if (t == 0) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int w = 0; w < kSWarps; ++w) {
      for (int ax = 0; ax < 3; ++ax) {
        lo[ax] = fminf(lo[ax], S.red[ax][w]);
        hi[ax] = fmaxf(hi[ax], S.red[3 + ax][w]);
      }
    }
    // Grouping = a two-level equal-count split over the two longest axes of the item's box:
    // slabs of 256 vertices along u, then groups of 32 along v inside each slab, so that every
    // warp pass is a compact column (the shortest axis — height, for a building-scale mesh — is
    // left unsplit: the alive control points span it anyway).  NaN / inf coordinates degrade
    // the ordering only (bins are clamped), never the result.
    float ext[3];
    for (int ax = 0; ax < 3; ++ax) {
      ext[ax] = hi[ax] - lo[ax];
      if (!(ext[ax] >= 0.f) || ext[ax] > 1e30f) ext[ax] = 0.f;
    }
    int u = 0;
    if (ext[1] > ext[u]) u = 1;
    if (ext[2] > ext[u]) u = 2;
    int v = (u + 1) % 3, w3 = (u + 2) % 3;
    if (ext[w3] > ext[v]) v = w3;
    S.grid[0] = lo[u];
    S.grid[1] = ext[u] > 0.f ? static_cast<float>(kSMaxCells) / ext[u] : 0.f;
    S.grid[2] = lo[v];
    S.grid[3] = ext[v] > 0.f ? static_cast<float>(kSVBins) / ext[v] : 0.f;
    S.grid[4] = __int_as_float(u);
    S.grid[5] = __int_as_float(v);
  }
  __syncthreads();
  const float u_lo = S.grid[0], u_inv = S.grid[1], v_lo = S.grid[2], v_inv = S.grid[3];
  const int axis_u = __float_as_int(S.grid[4]), axis_v = __float_as_int(S.grid[5]);
  const float* cu = axis_u == 0 ? S.vx : (axis_u == 1 ? S.vy : S.vz);
  const float* cv = axis_v == 0 ? S.vx : (axis_v == 1 ? S.vy : S.vz);
  uint16_t* perm1 = S.u.sort.perm1;

  // block-wide exclusive scan of hist[1..n_bins] (in place: hist[i] = #entries in bins < i)
  auto scanHist = [&](int n_bins) {
    constexpr int per = (kSMaxCells + kSThreads - 1) / kSThreads;
    uint32_t local[per];
    uint32_t sum = 0;
#pragma unroll
    for (int e = 0; e < per; ++e) {
      const int i = 1 + t * per + e;
      local[e] = i <= n_bins ? S.u.sort.hist[i] : 0;
      sum += local[e];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) S.warp_tot[warp] = incl;
    __syncthreads();
    uint32_t run_sum = incl - sum;
    for (int w = 0; w < warp; ++w) run_sum += S.warp_tot[w];
#pragma unroll
    for (int e = 0; e < per; ++e) {
      const int i = 1 + t * per + e;
      run_sum += local[e];
      if (i <= n_bins) S.u.sort.hist[i] = run_sum;
    }
    __syncthreads();
  };
  auto clampBin = [](float c, float lo, float inv, int n) {
    int b = static_cast<int>((c - lo) * inv);
    b = b < 0 ? 0 : b;
    return b >= n ? n - 1 : b;
  };

  // ---- pass 1: counting sort along u (1024 bins) -> perm1
  for (int i = t; i < count; i += kSThreads) {
    atomicAdd(&S.u.sort.hist[clampBin(cu[i], u_lo, u_inv, kSMaxCells) + 1], 1u);
  }
  __syncthreads();
  scanHist(kSMaxCells);
  for (int i = t; i < count; i += kSThreads) {
    const uint32_t at = atomicAdd(&S.u.sort.hist[clampBin(cu[i], u_lo, u_inv, kSMaxCells)], 1u);
    perm1[at] = static_cast<uint16_t>(i);
  }
  __syncthreads();
  // ---- pass 2: inside each slab of 256 (by rank along u) counting sort along v (128 bins)
  for (int i = t; i <= kSMaxCells; i += kSThreads) S.u.sort.hist[i] = 0;
  __syncthreads();
  for (int rk = t; rk < count; rk += kSThreads) {
    const int key = (rk >> kSlabShift) * kSVBins + clampBin(cv[perm1[rk]], v_lo, v_inv, kSVBins);
    atomicAdd(&S.u.sort.hist[key + 1], 1u);
  }
  __syncthreads();
  scanHist(kSMaxCells);
  for (int rk = t; rk < count; rk += kSThreads) {
    const int i = perm1[rk];
    const int key = (rk >> kSlabShift) * kSVBins + clampBin(cv[i], v_lo, v_inv, kSVBins);
    const uint32_t at = atomicAdd(&S.u.sort.hist[key], 1u);
    S.perm[at] = static_cast<uint16_t>(i);
  }
  cpAsyncWaitAll();
  __syncthreads();  // hist is dead from here on: the candidate lists alias it
  }

  // ---- per group of 32 sorted vertices: select the alive points around the group's box, walk
  const unsigned full = 0xFFFFFFFFu;
  uint16_t* list = S.u.w.list[warp];
  const int n_need = FULL ? KCAP : (n_alive < K + 1 ? n_alive : K + 1);
  const int n = FULL ? KCAP - 1 : (n_alive < K ? n_alive : K);
  // ordinal bits: enough for every list position of this run (two padding entries included)
  const int ord_bits = n_alive + 2 <= 128 ? 7 : (n_alive + 2 <= 256 ? 8 : (n_alive + 2 <= 512 ? 9 : 10));
  const uint32_t ord_mask = (1u << ord_bits) - 1u;
  // first selection radius^2 (box distance) and the radius^2 beyond which a point is "far"
  const float t0 = fmaxf(g.r0 * g.r0 * 1.3f * ga.r_scale, 1e-20f);
  const float t_far = fmaxf(ga.far_mul, 1.f) * t0;
  const bool full_k = FULL || n_need == KCAP;

  // A warp takes whole slabs (256 sorted vertices = 8 groups).  Per slab the alive points are
  // split once: points whose lower-bound distance to the SLAB's box is beyond t_far are "far"
  // for every group of the slab (a group's box lies inside the slab's and every step of the
  // bound is monotone in float); the others ("near") are the only ones a group looks at unless
  // its search radius reaches t_far.
  using NearT = typename std::conditional<(MAXPTS <= 256), uint8_t, uint16_t>::type;
  NearT* near = S.near[warp];
  const int n_slabs = (count + kSlab - 1) >> kSlabShift;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int slab = warp; slab < n_slabs; slab += kSWarps) {
    const int s_begin = slab << kSlabShift;
    const int s_end = count < s_begin + kSlab ? count : s_begin + kSlab;
    float sl0 = FLT_MAX, sl1 = FLT_MAX, sl2 = FLT_MAX, sh0 = -FLT_MAX, sh1 = -FLT_MAX, sh2 = -FLT_MAX;
    for (int rk = s_begin + lane; rk < s_end; rk += 32) {
      const int vv = S.perm[rk];
      sl0 = fminf(sl0, S.vx[vv]);
      sh0 = fmaxf(sh0, S.vx[vv]);
      sl1 = fminf(sl1, S.vy[vv]);
      sh1 = fmaxf(sh1, S.vy[vv]);
      sl2 = fminf(sl2, S.vz[vv]);
      sh2 = fmaxf(sh2, S.vz[vv]);
    }
    sl0 = warpMinF(sl0);
    sl1 = warpMinF(sl1);
    sl2 = warpMinF(sl2);
    sh0 = warpMaxF(sh0);
    sh1 = warpMaxF(sh1);
    sh2 = warpMaxF(sh2);
    int n_near = 0, n_far = 0;
    __syncwarp();  // the previous slab's groups are done with near[]
    for (int i0 = 0; i0 < n_alive; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < n_alive;
      bool is_near = false;
      if (valid) {
        const float4 c = S.pts[i];
        const float ax = fmaxf(fmaxf(sl0 - c.x, c.x - sh0), 0.f);
        const float ay = fmaxf(fmaxf(sl1 - c.y, c.y - sh1), 0.f);
        const float az = fmaxf(fmaxf(sl2 - c.z, c.z - sh2), 0.f);
        const float m = (ax * ax + ay * ay + az * az) * 0.9999f;
        is_near = !(m >= t_far);  // NaN: near
      }
      const unsigned m_near = __ballot_sync(full, is_near);
      const unsigned m_far = __ballot_sync(full, valid && !is_near);
      if (is_near) {
        near[n_near + __popc(m_near & lt_mask)] = static_cast<NearT>(i);
      } else if (valid) {
        near[n_alive - 1 - (n_far + __popc(m_far & lt_mask))] = static_cast<NearT>(i);
      }
      n_near += __popc(m_near);
      n_far += __popc(m_far);
    }
    __syncwarp();

  for (int group = s_begin >> 5; group * 32 < s_end; ++group) {
    const int slot = group * 32 + lane;
    const bool active = slot < count;
    int v = S.perm[active ? slot : group * 32];  // idle lanes shadow the group's first vertex
    float qx = S.vx[v], qy = S.vy[v], qz = S.vz[v];
    const float lox = warpMinF(qx), hix = warpMaxF(qx);
    const float loy = warpMinF(qy), hiy = warpMaxF(qy);
    const float loz = warpMinF(qz), hiz = warpMaxF(qz);
#ifndef PGMO_NO_LANE_REGROUP
    // Lanes regrouped by octant of the group's box (the shortest axis is the lowest key bit, so
    // that a quarter warp = two octants that differ along it).  Records are written in lane
    // order: the quarter warps of the blend kernel then read the pose rows of vertices that lie
    // close together — fewer distinct rows per shared-memory phase.  The search result does not
    // depend on the lane a vertex sits in.
    {
      const float ex = hix - lox, ey = hiy - loy, ez = hiz - loz;
      const uint32_t bx = qx > 0.5f * (lox + hix) ? 1u : 0u;
      const uint32_t by = qy > 0.5f * (loy + hiy) ? 1u : 0u;
      const uint32_t bz = qz > 0.5f * (loz + hiz) ? 1u : 0u;
      uint32_t key;
      if (ex <= ey && ex <= ez) key = (ey >= ez ? (by << 2) | (bz << 1) : (bz << 2) | (by << 1)) | bx;
      else if (ey <= ez) key = (ex >= ez ? (bx << 2) | (bz << 1) : (bz << 2) | (bx << 1)) | by;
      else key = (ex >= ey ? (bx << 2) | (by << 1) : (by << 2) | (bx << 1)) | bz;
      if (!active) key = 8u;
      uint32_t rank = __popc(__match_any_sync(full, key) & lt_mask);
#pragma unroll
      for (uint32_t c = 0; c < 8; ++c) {
        const unsigned m = __ballot_sync(full, key == c);
        rank += key > c ? __popc(m) : 0;
      }
      __syncwarp();
      if (active) S.perm[group * 32 + rank] = static_cast<uint16_t>(v);
      __syncwarp();
      v = S.perm[active ? slot : group * 32];
      qx = S.vx[v];
      qy = S.vy[v];
      qz = S.vz[v];
    }
#endif

    // lower bound (with head room for the separately rounded exact d^2 of every lane) of the
    // squared distance between near point i and the group's box
    auto boxBound = [&](int i) {
      const float4 c = S.pts[near[i]];
      const float ax = fmaxf(fmaxf(lox - c.x, c.x - hix), 0.f);
      const float ay = fmaxf(fmaxf(loy - c.y, c.y - hiy), 0.f);
      const float az = fmaxf(fmaxf(loz - c.z, c.z - hiz), 0.f);
      return (ax * ax + ay * ay + az * az) * 0.9999f;
    };
    uint32_t best[KCAP];
#pragma unroll
    for (int i = 0; i < KCAP; ++i) best[i] = 0xFFFFFFFFu;
    // Two selection phases, each followed by a walk of the selected slots (exact d^2 of every
    // lane's vertex to the candidate — one broadcast LDS.128 each —, keys inserted two at a
    // time).  Phase 0: the near points within t0 of the box, in slot order.  Phase 1: the other
    // near points that are not farther from the box than the largest (K+1)-th distance found in
    // phase 0 — nothing beyond that can enter (or tie with) anybody's list.
    int n_list = 0;
    float k_max = 0.f;
    __syncwarp();  // the previous group's winners have been read back from the list
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      const int c0 = n_list;
      for (int i0 = 0; i0 < n_near; i0 += 32) {
        const int i = i0 + lane;
        bool sel = false;
        if (i < n_near) {
          const float m = boxBound(i);
          sel = phase == 0 ? !(m >= t0) : (m >= t0 && m <= k_max);
        }
        const unsigned msel = __ballot_sync(full, sel);
        if (sel) list[n_list + __popc(msel & lt_mask)] = near[i];
        n_list += __popc(msel);
      }
      // too few to fill the lists (phase 0), or the search radius reaches the slab's far points
      // (phase 1): take everything that is left (near[] holds all alive slots)
      const bool everything = phase == 0 ? n_list < n_need : k_max >= t_far;
      if (everything) {
        const int from = phase == 0 ? 0 : n_near;
        const int to = n_alive;
        n_list = phase == 0 ? 0 : n_list;
        for (int i = from + lane; i < to; i += 32) list[n_list + (i - from)] = near[i];
        n_list += to - from;
      }
      if (n_list & 1) {
        if (lane == 0) list[n_list] = static_cast<uint16_t>(n_alive);
        ++n_list;
      }
      __syncwarp();
      int c = c0;
      for (; c + 4 <= n_list; c += 4) {
        const uint32_t e01 = *reinterpret_cast<const uint32_t*>(list + c);
        const uint32_t e23 = *reinterpret_cast<const uint32_t*>(list + c + 2);
        const uint32_t e[4] = {e01 & 0xFFFFu, e01 >> 16, e23 & 0xFFFFu, e23 >> 16};
        uint32_t key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 p = S.pts[e[u]];  // same address in all lanes: broadcast
          const float d2 = dist2Exact(p.x, p.y, p.z, qx, qy, qz);
          key[u] = (__float_as_uint(d2) & ~ord_mask) | static_cast<uint32_t>(c + u);
        }
        pairInsert<KCAP>(best, key[0], key[1]);
        pairInsert<KCAP>(best, key[2], key[3]);
      }
      if (c < n_list) {
        const uint32_t e01 = *reinterpret_cast<const uint32_t*>(list + c);
        const float4 p0 = S.pts[e01 & 0xFFFFu], p1 = S.pts[e01 >> 16];
        const float d0 = dist2Exact(p0.x, p0.y, p0.z, qx, qy, qz);
        const float d1 = dist2Exact(p1.x, p1.y, p1.z, qx, qy, qz);
        pairInsert<KCAP>(best, (__float_as_uint(d0) & ~ord_mask) | static_cast<uint32_t>(c),
                         (__float_as_uint(d1) & ~ord_mask) | static_cast<uint32_t>(c + 1));
      }
#ifdef PGMO_DEBUG_COUNTERS
      if (lane == 0) atomicAdd(a.counters + 5, static_cast<unsigned long long>(n_list - c0));
#endif
      if (everything) break;
      if (phase == 0) {
        uint32_t kth;  // best[n_need - 1]: the maximum of the first n_need (ascending) entries
        if (full_k) {
          kth = best[KCAP - 1];
        } else {
          kth = 0u;
#pragma unroll
          for (int i = 0; i < KCAP; ++i) kth = i < n_need ? max(kth, best[i]) : kth;
        }
        // (key | ord_mask) >= the entry's exact d^2 bits
        const float kth_f = __uint_as_float(kth | ord_mask);
        k_max = warpMaxF(kth_f);
        if (__any_sync(full, !(kth_f == kth_f))) k_max = __int_as_float(0x7F800000);  // NaN / inf d^2
        if (k_max < t0) break;  // the usual case: phase 0 already covered every lane's radius
      }
    }
#ifdef PGMO_DEBUG_COUNTERS
    if (lane == 0) {
      atomicAdd(a.counters + 4, 1ull);
      atomicAdd(a.counters + 6, static_cast<unsigned long long>((hix - lox) * (hiy - loy) * 1000.f));
      atomicAdd(a.counters + 7, static_cast<unsigned long long>((hiz - loz) * 1000.f));
    }
#endif
    if (!active) continue;

    // winners: exact d^2 and control index back from the list
    const uint64_t p = p_base + v;
    float nbr_d2[KCAP];
    int nbr_idx[KCAP];
    uint32_t nbr_cls[KCAP];  // bank class of the neighbour's pose row (deform_blend.cu)
    bool near_tie = false;   // two winners agree above the ordinal bits
    bool near_last = false;  // ... and one of them is the last tracked slot
#pragma unroll
    for (int i = 0; i < KCAP; ++i) {
      nbr_d2[i] = FLT_MAX;
      nbr_idx[i] = -1;
      nbr_cls[i] = 0u;
      if (i < n_need) {
        const int slot_i = list[best[i] & ord_mask];
        const float4 c = S.pts[slot_i];
        nbr_d2[i] = dist2Exact(c.x, c.y, c.z, qx, qy, qz);
        nbr_idx[i] = __float_as_int(c.w);
        nbr_cls[i] = __ldg(cache.ctrl_class + nbr_idx[i]);
      }
      if (i > 0 && i < n_need && (best[i] >> ord_bits) == (best[i - 1] >> ord_bits)) {
        near_tie = true;
        near_last = near_last || (i == n_need - 1);
      }
    }
    bool tie = false;
    if (near_tie) {
      // Keys that agree above the ordinal bits were ordered by list position, not by d^2.
      // Every point sharing that prefix is tracked unless the group reaches the last slot
      // (untracked points have keys >= the last slot's), so below the last slot the exact
      // order is recovered by sorting the winners on their exact d^2; what remains equal is an
      // exact tie.  A group touching the last slot is checked for untracked members first.
      // Exact ties, groups with untracked members and anything two bubble passes do not sort
      // (>= 4 near-equal winners) go to the resolver.
      if (near_last && n_alive > n_need) {
        // does an untracked point share the last slot's prefix?  Count all of them (rare path:
        // ~1e-4 of the vertices) and compare with the number of tracked ones.
        uint32_t last_key = 0u;  // best[n_need - 1] without dynamic indexing (ascending keys)
#pragma unroll
        for (int i = 0; i < KCAP; ++i) last_key = i < n_need ? max(last_key, best[i]) : last_key;
        const uint32_t prefix = last_key >> ord_bits;
        int tracked = 0;
#pragma unroll
        for (int i = 0; i < KCAP; ++i) tracked += (i < n_need && (best[i] >> ord_bits) == prefix) ? 1 : 0;
        int all = 0;
        for (int j = 0; j < n_alive; ++j) {
          const float4 c = S.pts[j];
          all += ((__float_as_uint(dist2Exact(c.x, c.y, c.z, qx, qy, qz)) >> ord_bits) == prefix) ? 1 : 0;
        }
        tie = all != tracked;
      }
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int i = 1; i < KCAP; ++i) {
          if (i < n_need && nbr_d2[i] < nbr_d2[i - 1]) {
            const float td = nbr_d2[i];
            nbr_d2[i] = nbr_d2[i - 1];
            nbr_d2[i - 1] = td;
            const int ti = nbr_idx[i];
            nbr_idx[i] = nbr_idx[i - 1];
            nbr_idx[i - 1] = ti;
            const uint32_t tc = nbr_cls[i];
            nbr_cls[i] = nbr_cls[i - 1];
            nbr_cls[i - 1] = tc;
          }
        }
      }
#pragma unroll
      for (int i = 1; i < KCAP; ++i) {
        if (i < n_need && nbr_d2[i] <= nbr_d2[i - 1]) tie = true;
      }
    }
    // ---- association record: vertex offset + neighbours (relative to run.lo) + weights
    const uint64_t rec_slot = p_base + static_cast<uint64_t>(slot);
    uint32_t rel[KCAP];
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
      rel[j] = (j < n - 1) ? (static_cast<uint32_t>(nbr_idx[j]) - info.x) | ((nbr_cls[j] & 7u) << kCacheClassShift) : 0u;
    }
    {
      uint4 rec0;
      rec0.x = static_cast<uint32_t>(v) | (rel[0] << 16);
      rec0.y = (KCAP > 1 ? rel[1] : 0u) | ((KCAP > 2 ? rel[2] : 0u) << 16);
      rec0.z = (KCAP > 3 ? rel[3] : 0u) | ((KCAP > 4 ? rel[4] : 0u) << 16);
      rec0.w = (KCAP > 5 ? rel[5] : 0u) | ((KCAP > 6 ? rel[6] : 0u) << 16);
      reinterpret_cast<uint4*>(cache.rec)[rec_slot] = rec0;
      if (KCAP >= 15 && cache.w_planes > 7) {  // more than 7 blended neighbours: plane 1 = nbr 7..14
        const uint4 rec1 = make_uint4(rel[7] | (rel[8] << 16), rel[9] | (rel[10] << 16),
                                      rel[11] | (rel[12] << 16), rel[13] | (rel[14] << 16));
        reinterpret_cast<uint4*>(cache.rec)[cache.n_slots + rec_slot] = rec1;
      }
    }
    if (tie) {  // exact tie (or an unresolvable near tie): the resolver is exact for both
      sendToResolver(a, p, static_cast<uint32_t>(rec_slot));
      continue;
    }
    {
      // nbr_d2[n - 1] without dynamic indexing: distances ascend
      float last_d2 = nbr_d2[0];
#pragma unroll
      for (int j = 1; j < KCAP; ++j) last_d2 = j < n ? fmaxf(last_d2, nbr_d2[j]) : last_d2;
      const float d_max_f = __fsqrt_rn(last_d2);
      const bool const_w = constWeight(nbr_d2[0], d_max_f);
      if constexpr (PREC == 0) {
        double* wp = static_cast<double*>(cache.weight) + rec_slot;
#pragma unroll
        for (int j = 0; j < KCAP; ++j) {
          if (j < n - 1) wp[static_cast<uint64_t>(j) * cache.n_slots] = weightFp64(nbr_d2[j], d_max_f, const_w);
        }
      } else {
        const float inv_dmax = const_w ? 0.0f : __frcp_rn(d_max_f);
        float* wp = static_cast<float*>(cache.weight) + rec_slot;
#pragma unroll
        for (int j = 0; j < KCAP; ++j) {
          if (j < n - 1) wp[static_cast<uint64_t>(j) * cache.n_slots] = weightFp32(nbr_d2[j], inv_dmax, const_w);
        }
      }
    }
  }
  }
  __syncthreads();  // every record of the item is on its way: publish the item
  if (t == 0) {
    cache.item_fp[item_index] = itemFingerprint(S);
    cache.item_ok[item_index] = 1u;
  }
}

// one CTA per work item
template <int KCAP, int PREC, int MAXPTS, bool FULL>
__global__ void __launch_bounds__(kSThreads, PGMO_K3_CTAS)
deformSearchSortedKernel(const DeformArgs a, const GridArgs ga, const AssocCache cache,
                         const WorkItem* __restrict__ items, const uint32_t first_item) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortedSmem<MAXPTS>& S = *reinterpret_cast<SortedSmem<MAXPTS>*>(smem_raw);
  const uint32_t item_index = first_item + blockIdx.x;
  searchItem<KCAP, PREC, MAXPTS, FULL>(a, ga, cache, items[item_index], item_index, S);
}

// persistent grid over the items the validating blend pass found stale
template <int KCAP, int PREC, int MAXPTS, bool FULL>
__global__ void __launch_bounds__(kSThreads, PGMO_K3_CTAS)
deformSearchStaleKernel(const DeformArgs a, const GridArgs ga, const AssocCache cache,
                        const WorkItem* __restrict__ items) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortedSmem<MAXPTS>& S = *reinterpret_cast<SortedSmem<MAXPTS>*>(smem_raw);
  const unsigned long long begin = a.counters[3], end = a.counters[2];
  for (unsigned long long i = begin + blockIdx.x; i < end; i += gridDim.x) {
    const uint32_t item_index = cache.stale_items[i];
    searchItem<KCAP, PREC, MAXPTS, FULL>(a, ga, cache, items[item_index], item_index, S);
    __syncthreads();  // shared memory is reused by the next item
  }
}

template <int KCAP, int PREC, int MAXPTS, bool FULL>
cudaError_t launchOne(const DeformArgs& args, const GridArgs& gargs, const AssocCache& cache,
                      const WorkItem* items, uint32_t first_item, uint32_t n_items,
                      bool from_stale_list, int n_sms, cudaStream_t stream) {
  const size_t smem = sizeof(SortedSmem<MAXPTS>);
  static thread_local int configured_device = -1;
  int device = 0;
  cudaGetDevice(&device);
  if (configured_device != device) {
    cudaError_t e = cudaFuncSetAttribute(deformSearchSortedKernel<KCAP, PREC, MAXPTS, FULL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(deformSearchStaleKernel<KCAP, PREC, MAXPTS, FULL>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    configured_device = device;
  }
  if (from_stale_list) {
    deformSearchStaleKernel<KCAP, PREC, MAXPTS, FULL>
        <<<static_cast<unsigned>(n_sms * PGMO_K3_CTAS), kSThreads, smem, stream>>>(args, gargs, cache, items);
  } else {
    deformSearchSortedKernel<KCAP, PREC, MAXPTS, FULL>
        <<<n_items, kSThreads, smem, stream>>>(args, gargs, cache, items, first_item);
  }
  return cudaGetLastError();
}

template <int KCAP, int PREC, bool FULL>
cudaError_t launchVariant(const DeformArgs& args, const GridArgs& gargs, const AssocCache& cache,
                          const WorkItem* items, uint32_t first_item, uint32_t n_items, bool small,
                          bool from_stale_list, int n_sms, cudaStream_t stream) {
  return small ? launchOne<KCAP, PREC, kSmallMaxPts, FULL>(args, gargs, cache, items, first_item, n_items,
                                                          from_stale_list, n_sms, stream)
               : launchOne<KCAP, PREC, kSortedMaxPoints, FULL>(args, gargs, cache, items, first_item,
                                                               n_items, from_stale_list, n_sms, stream);
}

template <int KCAP, int PREC>
void preloadVariants() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, deformSearchSortedKernel<KCAP, PREC, kSmallMaxPts, true>);
  cudaFuncGetAttributes(&attr, deformSearchStaleKernel<KCAP, PREC, kSmallMaxPts, true>);
  cudaFuncGetAttributes(&attr, deformSearchSortedKernel<KCAP, PREC, kSortedMaxPoints, true>);
}

}  // namespace

// CUDA loads kernels lazily at their first launch (hundreds of ms for the large K3 variants):
// pull the common ones in when the handle is created instead of inside the first deformMesh.
void preloadDeformSortedKernels() {
  preloadVariants<7, 0>();
  preloadVariants<7, 1>();
}

cudaError_t launchDeformSearchSorted(const DeformArgs& args, const GridArgs& gargs,
                                     const AssocCache& cache, const WorkItem* items,
                                     uint32_t first_item, uint32_t n_items, int precision,
                                     uint32_t max_k, uint32_t min_k, uint64_t widest_window,
                                     bool from_stale_list, int n_sms, cudaStream_t stream) {
  if (n_items == 0 && !from_stale_list) return cudaSuccess;
  // every alive set fits the lean variant iff the widest window does
  const bool small = widest_window <= static_cast<uint64_t>(kSmallMaxPts);
  if (max_k == 4 && min_k == 4) {  // the reference's default k, never changed mid-sweep
    return precision == 0 ? launchVariant<7, 0, true>(args, gargs, cache, items, first_item, n_items, small, from_stale_list, n_sms, stream)
                          : launchVariant<7, 1, true>(args, gargs, cache, items, first_item, n_items, small, from_stale_list, n_sms, stream);
  }
  if (max_k + 3 <= 7) {
    return precision == 0 ? launchVariant<7, 0, false>(args, gargs, cache, items, first_item, n_items, small, from_stale_list, n_sms, stream)
                          : launchVariant<7, 1, false>(args, gargs, cache, items, first_item, n_items, small, from_stale_list, n_sms, stream);
  }
  return precision == 0 ? launchVariant<16, 0, false>(args, gargs, cache, items, first_item, n_items, small, from_stale_list, n_sms, stream)
                        : launchVariant<16, 1, false>(args, gargs, cache, items, first_item, n_items, small, from_stale_list, n_sms, stream);
}

}  // namespace pgmo
