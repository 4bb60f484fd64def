// K3, bucket-ordered candidate lists (sm_100a): the main knn + blend kernel.
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
//   * The pose table rows of the run's alive points (R, t, g: 15 doubles, or floats in FP32
//     mode) are staged in shared memory as SoA when the alive set has <= 256 points, so that
//     the blend's 75 gathered loads per vertex are conflict-free LDS instead of L2 round trips
//     (profiles/r01e: the gathers were 1/3 of all stall samples and most of the LSU traffic).
//
// The result is bit-identical to the brute-force and grid-walk kernels (tests/test_deform_gpu.py
// ::test_search_modes_agree).  Runs with more than kSMaxPts alive points stay on the grid-walk
// kernel (deform_grid.cu).
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
#ifndef PGMO_STAGED_MAX
#define PGMO_STAGED_MAX 256
#endif
#ifndef PGMO_K3_CTAS
#define PGMO_K3_CTAS 3
#endif
constexpr int kStagedMaxPts = PGMO_STAGED_MAX;  // alive-set size up to which the pose rows are staged

// pose rows of the alive points, SoA: component c of slot s at row[c][s]
template <typename T, int MAXPTS>
struct PoseRows {
  T row[15][MAXPTS];  // R0..R8, t0..t2, g0..g2
};
struct NoPoseRows {};

template <int MAXPTS, typename POSE>
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
  alignas(16) POSE pose;
};

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

// interpPoint's blend with the pose rows read from shared memory (same arithmetic, operation
// for operation, as dev::writeResultReg).  nbr_slot: slots of the neighbours in S.pts / S.pose.
template <int KCAP, int PREC, typename Smem>
__device__ __forceinline__ void writeResultStaged(const DeformArgs& a, const Smem& S, uint64_t p,
                                                  uint64_t ii, float qx, float qy, float qz,
                                                  const float (&nbr_d2)[KCAP],
                                                  const int (&nbr_slot)[KCAP], int n) {
  float last_d2 = nbr_d2[0];
#pragma unroll
  for (int j = 1; j < KCAP; ++j) last_d2 = j < n ? fmaxf(last_d2, nbr_d2[j]) : last_d2;
  if constexpr (PREC == 0) {
    const float d_max_f = __fsqrt_rn(last_d2);
    const double d_max = static_cast<double>(d_max_f);
    const bool const_w = (__fsqrt_rn(nbr_d2[0]) == d_max_f) || d_max == 0.0;
    double wsum = 0.0, acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    const double v0 = static_cast<double>(qx), v1 = static_cast<double>(qy),
                 v2 = static_cast<double>(qz);
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
      if (j < n - 1) {
        const int s = nbr_slot[j];
        const double w =
            const_w ? 1.0
                    : __dsub_rn(1.0, __ddiv_rn(static_cast<double>(__fsqrt_rn(nbr_d2[j])), d_max));
        wsum = __dadd_rn(wsum, w);
        const double p0 = __dsub_rn(v0, S.pose.row[12][s]);
        const double p1 = __dsub_rn(v1, S.pose.row[13][s]);
        const double p2 = __dsub_rn(v2, S.pose.row[14][s]);
        const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(S.pose.row[0][s], p0), __dmul_rn(S.pose.row[1][s], p1)), __dmul_rn(S.pose.row[2][s], p2));
        const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(S.pose.row[3][s], p0), __dmul_rn(S.pose.row[4][s], p1)), __dmul_rn(S.pose.row[5][s], p2));
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(S.pose.row[6][s], p0), __dmul_rn(S.pose.row[7][s], p1)), __dmul_rn(S.pose.row[8][s], p2));
        acc0 = __dadd_rn(acc0, __dmul_rn(w, __dadd_rn(r0, S.pose.row[9][s])));
        acc1 = __dadd_rn(acc1, __dmul_rn(w, __dadd_rn(r1, S.pose.row[10][s])));
        acc2 = __dadd_rn(acc2, __dmul_rn(w, __dadd_rn(r2, S.pose.row[11][s])));
      }
    }
    const double o0 = __ddiv_rn(acc0, wsum), o1 = __ddiv_rn(acc1, wsum), o2 = __ddiv_rn(acc2, wsum);
    a.ox[ii] = __double2float_rn(o0);
    a.oy[ii] = __double2float_rn(o1);
    a.oz[ii] = __double2float_rn(o2);
    if (a.of64) {
      a.of64[3 * p + 0] = o0;
      a.of64[3 * p + 1] = o1;
      a.of64[3 * p + 2] = o2;
    }
  } else {
    const float d_max = __fsqrt_rn(last_d2);
    const bool const_w = (__fsqrt_rn(nbr_d2[0]) == d_max) || d_max == 0.0f;
    const float inv_dmax = const_w ? 0.0f : __frcp_rn(d_max);
    float wsum = 0.f, acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
      if (j < n - 1) {
        const int s = nbr_slot[j];
        const float w = const_w ? 1.0f : 1.0f - __fsqrt_rn(nbr_d2[j]) * inv_dmax;
        wsum += w;
        const float p0 = qx - S.pose.row[12][s], p1 = qy - S.pose.row[13][s],
                    p2 = qz - S.pose.row[14][s];
        acc0 += w * (S.pose.row[0][s] * p0 + S.pose.row[1][s] * p1 + S.pose.row[2][s] * p2 + S.pose.row[9][s]);
        acc1 += w * (S.pose.row[3][s] * p0 + S.pose.row[4][s] * p1 + S.pose.row[5][s] * p2 + S.pose.row[10][s]);
        acc2 += w * (S.pose.row[6][s] * p0 + S.pose.row[7][s] * p1 + S.pose.row[8][s] * p2 + S.pose.row[11][s]);
      }
    }
    const float o0 = acc0 / wsum, o1 = acc1 / wsum, o2 = acc2 / wsum;
    a.ox[ii] = o0;
    a.oy[ii] = o1;
    a.oz[ii] = o2;
    if (a.of64) {
      a.of64[3 * p + 0] = o0;
      a.of64[3 * p + 1] = o1;
      a.of64[3 * p + 2] = o2;
    }
  }
  if (a.assoc) {
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
      if (static_cast<uint32_t>(j) < a.assoc_stride) {
        a.assoc[p * a.assoc_stride + j] =
            (j < n - 1) ? static_cast<uint32_t>(__float_as_int(S.pts[nbr_slot[j]].w)) : 0xFFFFFFFFu;
      }
    }
    for (uint32_t j = KCAP; j < a.assoc_stride; ++j) a.assoc[p * a.assoc_stride + j] = 0xFFFFFFFFu;
  }
  if (a.assoc_cnt) a.assoc_cnt[p] = static_cast<uint8_t>(n - 1 > 255 ? 255 : n - 1);
}

template <int PREC, int MAXPTS, bool STAGE>
struct SmemOf {
  using Pose = typename std::conditional<
      STAGE, typename std::conditional<PREC == 0, PoseRows<double, MAXPTS>, PoseRows<float, MAXPTS>>::type,
      NoPoseRows>::type;
  using type = SortedSmem<MAXPTS, Pose>;
};

// FULL: every vertex blends exactly KCAP - 2 neighbours (k + 1 with KCAP = k + 3) — the caller
// asked for one k (no mid-sweep change) and the item's alive set holds at least KCAP points.  The
// neighbour counts are then compile-time constants and every per-slot guard disappears; items
// that do not qualify (a handful of vertices at the start of a trajectory) go to the resolver.
template <int KCAP, int PREC, int MAXPTS, bool STAGE, bool FULL>
__global__ void __launch_bounds__(kSThreads, STAGE ? PGMO_K3_CTAS : 4)
deformKnnBlendSortedKernel(const DeformArgs a, const GridArgs ga,
                           const WorkItem* __restrict__ items) {
  using Smem = typename SmemOf<PREC, MAXPTS, STAGE>::type;
  constexpr int kSMaxPts = MAXPTS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);

  const WorkItem item = items[blockIdx.x];
  const uint32_t r = item.run;
  const uint4 info = __ldg(a.run_info + r);
  const uint64_t p_base = item.start;
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int count = static_cast<int>(item.count);
  if (info.w) {
    for (int i = t; i < count; i += kSThreads) a.status[p_base + i] = kStatusSkipped;
    return;
  }
  const RunGridDev g = ga.grids[r];
  const int n_alive = static_cast<int>(g.n_alive);
  if (n_alive > kSMaxPts) return;  // another kernel's item
  const int K = static_cast<int>(info.z) + 2;
  if (K + 1 > KCAP || (FULL && (K + 1 != KCAP || n_alive < KCAP))) {
    for (int i = t; i < count; i += kSThreads) sendToResolver(a, p_base + i);
    return;
  }

  // Items of at most one group (the head vertex of a stamp stretch is a run of its own) skip
  // the vertex sort and stage only the pose rows of their winners.
  const bool tiny = count <= 32;

  // ---- stage the run's alive points; load the item's vertices and their bounding box
  {
    const float4* src = ga.points + g.point_offset;
    for (int i = t; i < n_alive; i += kSThreads) cpAsync<16>(&S.pts[i], src + i);
    // list padding (see the walk): d^2 to this point is +inf for every finite query
    if (t == 0) S.pts[n_alive] = make_float4(3e30f, 3e30f, 3e30f, __int_as_float(-1));
    if constexpr (STAGE) {
      // pose rows of the alive points: component-major so that lanes reading one component of
      // different slots never collide on a bank
      if (!tiny) {
        for (int e = t; e < n_alive * 15; e += kSThreads) {
          const int s_ = e / 15, c = e - s_ * 15;
          const int j = __float_as_int(__ldg(src + s_).w);
          if constexpr (PREC == 0) {
            cpAsync<8>(&S.pose.row[c][s_], c < 12 ? a.poses + 12ull * j + c : a.ctrl_pos + 3ull * j + (c - 12));
          } else {
            cpAsync<4>(&S.pose.row[c][s_], a.ctrl_pose_f32 + 16ull * j + c);  // R(9) t(3) g(3)
          }
        }
      }
    }
  }
  float mn0 = FLT_MAX, mn1 = FLT_MAX, mn2 = FLT_MAX, mx0 = -FLT_MAX, mx1 = -FLT_MAX,
        mx2 = -FLT_MAX;
  for (int i = t; i < count; i += kSThreads) {
    const uint64_t ii = a.subset ? a.subset[p_base + i] : p_base + i;
    const float x = a.x[ii], y = a.y[ii], z = a.z[ii];
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
    const int v = S.perm[active ? slot : group * 32];  // idle lanes shadow the group's first vertex
    const float qx = S.vx[v], qy = S.vy[v], qz = S.vz[v];
    const float lox = warpMinF(qx), hix = warpMaxF(qx);
    const float loy = warpMinF(qy), hiy = warpMaxF(qy);
    const float loz = warpMinF(qz), hiz = warpMaxF(qz);

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
    const uint64_t ii = a.subset ? a.subset[p] : p;
    float nbr_d2[KCAP];
    int nbr_idx[KCAP];
    bool near_tie = false;   // two winners agree above the ordinal bits
    bool near_last = false;  // ... and one of them is the last tracked slot
#pragma unroll
    for (int i = 0; i < KCAP; ++i) {
      nbr_d2[i] = FLT_MAX;
      nbr_idx[i] = -1;
      if (i < n_need) {
        const int slot_i = list[best[i] & ord_mask];
        const float4 c = S.pts[slot_i];
        nbr_d2[i] = dist2Exact(c.x, c.y, c.z, qx, qy, qz);
        nbr_idx[i] = STAGE ? slot_i : __float_as_int(c.w);  // staged: the slot, resolved at the end
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
          }
        }
      }
#pragma unroll
      for (int i = 1; i < KCAP; ++i) {
        if (i < n_need && nbr_d2[i] <= nbr_d2[i - 1]) tie = true;
      }
    }
    if (tie) {  // exact tie (or an unresolvable near tie): the resolver is exact for both
      sendToResolver(a, p);
      continue;
    }
    if constexpr (STAGE) {
      if (tiny) {
        // the pose rows were not staged: fetch those of this vertex's own neighbours (lanes
        // that share a neighbour write identical values)
#pragma unroll
        for (int j = 0; j < KCAP; ++j) {
          if (j < n - 1) {
            const int s_ = nbr_idx[j];
            const int cj = __float_as_int(S.pts[s_].w);
#pragma unroll 1
            for (int c = 0; c < 15; ++c) {
              if constexpr (PREC == 0) {
                S.pose.row[c][s_] = c < 12 ? a.poses[12ull * cj + c] : a.ctrl_pos[3ull * cj + (c - 12)];
              } else {
                S.pose.row[c][s_] = a.ctrl_pose_f32[16ull * cj + c];
              }
            }
          }
        }
      }
      writeResultStaged<KCAP, PREC>(a, S, p, ii, qx, qy, qz, nbr_d2, nbr_idx, n);
    } else {
      writeResultReg<KCAP>(a, p, ii, PREC, qx, qy, qz, nbr_d2, nbr_idx, n);
    }
    a.status[p] = kStatusDone;
  }
  }
}

template <int KCAP, int PREC, int MAXPTS, bool STAGE, bool FULL>
cudaError_t launchOne(const DeformArgs& args, const GridArgs& gargs, const WorkItem* items,
                      uint32_t n_items, cudaStream_t stream) {
  const size_t smem = sizeof(typename SmemOf<PREC, MAXPTS, STAGE>::type);
  static thread_local int configured_device = -1;
  int device = 0;
  cudaGetDevice(&device);
  if (configured_device != device) {
    const cudaError_t e =
        cudaFuncSetAttribute(deformKnnBlendSortedKernel<KCAP, PREC, MAXPTS, STAGE, FULL>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    configured_device = device;
  }
  deformKnnBlendSortedKernel<KCAP, PREC, MAXPTS, STAGE, FULL>
      <<<n_items, kSThreads, smem, stream>>>(args, gargs, items);
  return cudaGetLastError();
}

template <int KCAP, int PREC, bool FULL>
cudaError_t launchVariant(const DeformArgs& args, const GridArgs& gargs, const WorkItem* items,
                          uint32_t n_items, bool staged, cudaStream_t stream) {
  return staged ? launchOne<KCAP, PREC, kStagedMaxPts, true, FULL>(args, gargs, items, n_items, stream)
                : launchOne<KCAP, PREC, kSortedMaxPoints, false, FULL>(args, gargs, items, n_items, stream);
}

template <int KCAP, int PREC>
void preloadVariants() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, deformKnnBlendSortedKernel<KCAP, PREC, kStagedMaxPts, true, true>);
  cudaFuncGetAttributes(&attr, deformKnnBlendSortedKernel<KCAP, PREC, kSortedMaxPoints, false, true>);
}

}  // namespace

// CUDA loads kernels lazily at their first launch (hundreds of ms for the large K3 variants):
// pull the common ones in when the handle is created instead of inside the first deformMesh.
void preloadDeformSortedKernels() {
  preloadVariants<7, 0>();
  preloadVariants<7, 1>();
}

cudaError_t launchDeformKnnBlendSorted(const DeformArgs& args, const GridArgs& gargs,
                                       const WorkItem* items, uint32_t n_items, int precision,
                                       uint32_t max_k, uint32_t min_k, uint64_t widest_window,
                                       cudaStream_t stream) {
  if (n_items == 0) return cudaSuccess;
  // every alive set fits the staged variant (pose rows in shared memory) iff the widest window does
  const bool staged = widest_window <= static_cast<uint64_t>(kStagedMaxPts);
  if (max_k == 4 && min_k == 4) {  // the reference's default k, never changed mid-sweep
    return precision == 0 ? launchVariant<7, 0, true>(args, gargs, items, n_items, staged, stream)
                          : launchVariant<7, 1, true>(args, gargs, items, n_items, staged, stream);
  }
  if (max_k + 3 <= 7) {
    return precision == 0 ? launchVariant<7, 0, false>(args, gargs, items, n_items, staged, stream)
                          : launchVariant<7, 1, false>(args, gargs, items, n_items, staged, stream);
  }
  return precision == 0 ? launchVariant<16, 0, false>(args, gargs, items, n_items, staged, stream)
                        : launchVariant<16, 1, false>(args, gargs, items, n_items, staged, stream);
}

}  // namespace pgmo
