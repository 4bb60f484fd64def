// K3b, blend from the association cache (sm_100a).
//
// deformation::interpPoint (kimera_pgmo/src/mesh_deformation.cpp:91-105) after the search:
//   acc += w_j * (R_j * (double(v) - g_j) + t_j)  for the k + 1 nearest, acc / sum(w), cast to
// float — with the neighbour indices and weights read from the records the search kernel (K3a,
// deform_sorted.cu) or the tie resolver (K4) left in HBM.  Those depend on the vertex, the
// control positions / stamps / validity and the window schedule, not on the poses, so after a
// new solve this kernel alone redoes a deformMesh: per vertex it reads 12 B of coordinates, a
// 16-byte record, k + 1 weights and writes 12 B.
//
//   * One CTA = one work item (<= 2048 consecutive positions of one run).  The pose rows of the
//     run's window [lo, hi) — R(9) t(3) g(3) as 128-byte rows, one contiguous block of the
//     table — arrive by ONE bulk asynchronous copy (cp.async.bulk, TMA engine; completion on an
//     mbarrier) that overlaps the vertex loads.
//   * The kernel is bound by the shared-memory data pipe: per vertex and neighbour a lane gathers
//     a 120-byte row (7 LDS.128 + 1 LDS.64), an LDS.128 is served per quarter warp, and the 8
//     lanes of a quarter warp want ~4.5 different rows (scripts/cache_stats.py): one wavefront
//     only if those rows sit in different 16-byte bank groups.  The table therefore stores 16-byte
//     chunk c of row j at c ^ class(j) (class = the colour of the control point in a colouring of
//     the table, deform_api.cu: colourControls) — and stores every row kBlendCopies times under
//     different classes.  A placement pass (placeRecordsKernel, below) over fresh records picks,
//     per quarter warp and neighbour slot, copies whose bank groups differ: measured on the 10 M
//     benchmark 1.69 -> ~1.1 wavefronts per quarter warp and LDS.128.
//   * Vertices are loaded coalesced, fingerprinted (validating mode) and staged in shared
//     memory; a lane takes the vertex of its record (records are in search order: neighbouring
//     lanes share neighbours, equal addresses are broadcast), blends in the reference's
//     operation order (explicit round-to-nearest intrinsics, no FMA) and leaves the result in
//     the vertex's staging slot; the stores to HBM are coalesced again.
//   * Validating mode: an item whose fingerprint differs from the one its records were computed
//     for (or that has no records yet) is appended to the stale list and left untouched; the
//     search kernel, the resolver and this kernel then run over that list.
#include <algorithm>
#include <cfloat>

#include "deform_device.cuh"

namespace pgmo {

namespace {

using namespace dev;

// CTA shape: 256 threads x 4 CTAs per SM (PGMO_K3B_THREADS / PGMO_K3B_CTAS).  An item goes
// through phases (vertices in, barrier, rows landed, blend, barrier, results out) and only other
// CTAs of the SM fill the shared-memory pipe while one of them loads or stores: 512 x 2 — which
// would leave room for kBlendCopies staged copies of the window's rows — measured 0.52 ms against
// 0.42 ms per 10 M vertices with one copy, 0.46 ms with three.
#ifndef PGMO_K3B_THREADS
#define PGMO_K3B_THREADS 256
#endif
#ifndef PGMO_K3B_CTAS
#define PGMO_K3B_CTAS 4
#endif
constexpr int kBThreads = PGMO_K3B_THREADS;
constexpr int kBWarps = kBThreads / 32;
constexpr int kBChunk = kDeformChunk;
// shared memory one CTA may use so that PGMO_K3B_CTAS of them fit an SM (228 KB, 1 KB per CTA
// reserved by the driver)
constexpr size_t kBSmemBudget = (228u * 1024u - PGMO_K3B_CTAS * 1024u) / PGMO_K3B_CTAS;

struct alignas(128) BlendSmemHead {
  float vx[kBChunk], vy[kBChunk], vz[kBChunk];  // vertices in, results out (same slots)
  unsigned long long fp_w[kBWarps];
  unsigned long long bar;                         // mbarrier of the row copy
  uint32_t stale;
};
static_assert(sizeof(BlendSmemHead) % 128 == 0, "pose rows start 128-byte aligned");

__device__ __forceinline__ uint32_t smemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbarInit(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// bulk copies global -> shared: the barrier is told the total byte count once, every copy
// signals its share on it
__device__ __forceinline__ void expectTx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, uint32_t bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smemAddr(dst)),
      "l"(src), "r"(bytes), "r"(smemAddr(bar))
      : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smemAddr(bar)),
      "r"(parity)
      : "memory");
}

// one neighbour's term of the FP64 blend; row = 8 chunks of 2 doubles, chunk c at c ^ sw
__device__ __forceinline__ void accumulate64(const double2* __restrict__ row, int sw, double w,
                                             double v0, double v1, double v2, double& wsum,
                                             double& acc0, double& acc1, double& acc2) {
  const double2 c6 = row[6 ^ sw];                                             // g0 g1
  const double c7x = *reinterpret_cast<const double*>(row + (7 ^ sw));  // g2
  const double2 c0 = row[0 ^ sw], c1 = row[1 ^ sw];  // R0 R1 | R2 R3
  const double2 c2 = row[2 ^ sw], c3 = row[3 ^ sw];  // R4 R5 | R6 R7
  const double2 c4 = row[4 ^ sw], c5 = row[5 ^ sw];  // R8 t0 | t1 t2
  wsum = __dadd_rn(wsum, w);
  const double p0 = __dsub_rn(v0, c6.x);
  const double p1 = __dsub_rn(v1, c6.y);
  const double p2 = __dsub_rn(v2, c7x);
  const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(c0.x, p0), __dmul_rn(c0.y, p1)), __dmul_rn(c1.x, p2));
  const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(c1.y, p0), __dmul_rn(c2.x, p1)), __dmul_rn(c2.y, p2));
  const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(c3.x, p0), __dmul_rn(c3.y, p1)), __dmul_rn(c4.x, p2));
  acc0 = __dadd_rn(acc0, __dmul_rn(w, __dadd_rn(r0, c4.y)));
  acc1 = __dadd_rn(acc1, __dmul_rn(w, __dadd_rn(r1, c5.x)));
  acc2 = __dadd_rn(acc2, __dmul_rn(w, __dadd_rn(r2, c5.y)));
}
// FP32 mode: row = 16 floats R(9) t(3) g(3) pad
__device__ __forceinline__ void accumulate32(const float4* __restrict__ row, float w, float qx,
                                             float qy, float qz, float& wsum, float& acc0,
                                             float& acc1, float& acc2) {
  const float4 t0 = row[0], t1 = row[1], t2 = row[2], t3 = row[3];
  wsum += w;
  const float p0 = qx - t3.x, p1 = qy - t3.y, p2 = qz - t3.z;
  acc0 += w * (t0.x * p0 + t0.y * p1 + t0.z * p2 + t2.y);
  acc1 += w * (t0.w * p0 + t1.x * p1 + t1.y * p2 + t2.z);
  acc2 += w * (t1.z * p0 + t1.w * p1 + t2.x * p2 + t2.w);
}

// the blend proper: lane <-> record (search order).  NB > 0: compile-time neighbour count.
// IN_PLACE: pose rows read from the table in global memory instead of the staged copy.
template <int PREC, int NB, bool IN_PLACE>
__device__ __forceinline__ void blendVertices(const DeformArgs& a, const AssocCache& cache,
                                              BlendSmemHead& S, const unsigned char* rows,
                                              uint64_t p_base, int count, int nb_runtime,
                                              uint32_t lo, uint32_t staged_rows) {
  const int t = threadIdx.x;
  const int nb = NB > 0 ? NB : nb_runtime;
  const uint4* rec0 = reinterpret_cast<const uint4*>(cache.rec);
  for (int q = t; q < count; q += kBThreads) {
    const uint64_t slot = p_base + static_cast<uint64_t>(q);
    const uint4 r0 = __ldg(rec0 + slot);
    const int v = static_cast<int>(r0.x & 0xFFFFu);
    uint32_t rel[kCacheMaxNbr + 1];
    rel[0] = r0.x >> 16;
    rel[1] = r0.y & 0xFFFFu;
    rel[2] = r0.y >> 16;
    rel[3] = r0.z & 0xFFFFu;
    rel[4] = r0.z >> 16;
    rel[5] = r0.w & 0xFFFFu;
    rel[6] = r0.w >> 16;
    if (NB == 0 && nb > 7) {
      const uint4 r1 = __ldg(rec0 + cache.n_slots + slot);
      rel[7] = r1.x & 0xFFFFu;
      rel[8] = r1.x >> 16;
      rel[9] = r1.y & 0xFFFFu;
      rel[10] = r1.y >> 16;
      rel[11] = r1.z & 0xFFFFu;
      rel[12] = r1.z >> 16;
      rel[13] = r1.w & 0xFFFFu;
      rel[14] = r1.w >> 16;
    } else {
#pragma unroll
      for (int j = 7; j <= kCacheMaxNbr; ++j) rel[j] = 0u;
    }
    const float qx = S.vx[v], qy = S.vy[v], qz = S.vz[v];
    const uint64_t p = p_base + v;
    if constexpr (PREC == 0) {
      const double* wp = static_cast<const double*>(cache.weight) + slot;
      const double v0 = static_cast<double>(qx), v1 = static_cast<double>(qy),
                   v2 = static_cast<double>(qz);
      double wsum = 0.0, acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
      // row of neighbour field f: copy plane (f >> 11 & 3) of the staged window / of the table
      const double2* table = IN_PLACE ? static_cast<const double2*>(cache.pose_rows) + 8ull * lo
                                      : reinterpret_cast<const double2*>(rows);
      const uint64_t plane = IN_PLACE ? cache.plane_rows : static_cast<uint64_t>(staged_rows);
      auto rowOf = [&](uint32_t f) {
        const uint32_t copy = (f >> kCacheCopyShift) & kCacheCopyMask;
        if (IN_PLACE) return table + 8ull * ((f & kCacheRelMask) + copy * plane);
        return table + 8u * ((f & kCacheRelMask) + copy * static_cast<uint32_t>(plane));
      };
      if (NB > 0) {
        double w[NB > 0 ? NB : 1];
#pragma unroll
        for (int j = 0; j < NB; ++j) w[j] = __ldg(wp + static_cast<uint64_t>(j) * cache.n_slots);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          accumulate64(rowOf(rel[j]), static_cast<int>(rel[j] >> kCacheClassShift), w[j], v0, v1, v2,
                       wsum, acc0, acc1, acc2);
        }
      } else {
#pragma unroll
        for (int j = 0; j < kCacheMaxNbr; ++j) {
          if (j < nb) {
            const double w = __ldg(wp + static_cast<uint64_t>(j) * cache.n_slots);
            accumulate64(rowOf(rel[j]), static_cast<int>(rel[j] >> kCacheClassShift), w, v0, v1, v2,
                         wsum, acc0, acc1, acc2);
          }
        }
      }
      const double o0 = __ddiv_rn(acc0, wsum), o1 = __ddiv_rn(acc1, wsum), o2 = __ddiv_rn(acc2, wsum);
      S.vx[v] = __double2float_rn(o0);
      S.vy[v] = __double2float_rn(o1);
      S.vz[v] = __double2float_rn(o2);
      if (a.of64) {
        a.of64[3 * p + 0] = o0;
        a.of64[3 * p + 1] = o1;
        a.of64[3 * p + 2] = o2;
      }
    } else {
      const float* wp = static_cast<const float*>(cache.weight) + slot;
      float wsum = 0.f, acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
      const float4* table = IN_PLACE ? static_cast<const float4*>(cache.pose_rows) + 4ull * lo
                                     : reinterpret_cast<const float4*>(rows);
#pragma unroll
      for (int j = 0; j < (NB > 0 ? NB : kCacheMaxNbr); ++j) {
        if (NB > 0 || j < nb) {
          const float w = __ldg(wp + static_cast<uint64_t>(j) * cache.n_slots);
          accumulate32(table + 4u * (rel[j] & kCacheRelMask), w, qx, qy, qz, wsum, acc0, acc1, acc2);
        }
      }
      const float o0 = acc0 / wsum, o1 = acc1 / wsum, o2 = acc2 / wsum;
      S.vx[v] = o0;
      S.vy[v] = o1;
      S.vz[v] = o2;
      if (a.of64) {
        a.of64[3 * p + 0] = o0;
        a.of64[3 * p + 1] = o1;
        a.of64[3 * p + 2] = o2;
      }
    }
    if (a.assoc) {
#pragma unroll
      for (int j = 0; j <= kCacheMaxNbr; ++j) {
        if (static_cast<uint32_t>(j) < a.assoc_stride) {
          a.assoc[p * a.assoc_stride + j] = j < nb ? lo + (rel[j] & kCacheRelMask) : 0xFFFFFFFFu;
        }
      }
      for (uint32_t j = kCacheMaxNbr + 1; j < a.assoc_stride; ++j) a.assoc[p * a.assoc_stride + j] = 0xFFFFFFFFu;
    }
    if (a.assoc_cnt) a.assoc_cnt[p] = static_cast<uint8_t>(nb);
  }
}

// MODE 0: records trusted; 1: validate, stale items go to the list; 2: item from the stale list
// (records just rebuilt).
template <int PREC, int MODE>
__device__ __forceinline__ void blendItem(const DeformArgs& a, const GridArgs& ga,
                                          const AssocCache& cache, const WorkItem& item,
                                          uint32_t item_index, unsigned char* smem_raw,
                                          uint32_t max_rows, uint32_t& bar_phase) {
  BlendSmemHead& S = *reinterpret_cast<BlendSmemHead*>(smem_raw);
  unsigned char* rows = smem_raw + sizeof(BlendSmemHead);
  constexpr uint32_t kRowBytes = PREC == 0 ? 128u : 64u;

  const uint32_t r = item.run;
  const uint4 info = __ldg(a.run_info + r);
  const uint64_t p_base = item.start;
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int count = static_cast<int>(item.count);
  if (info.w) {  // mesh_deformation.h:184-186: the reference `continue`s
    for (int i = t; i < count; i += kBThreads) a.status[p_base + i] = kStatusSkipped;
    return;
  }
  const int n_alive = static_cast<int>(ga.grids[r].n_alive);
  if (n_alive > kSortedMaxPoints) return;  // the grid-walk kernel's item (it blends by itself)
  const int K = static_cast<int>(info.z) + 2;
  if (!runCacheable(info)) {
    // no records possible: the resolver's own blend is final, and it has to run on every call
    if (MODE == 1 && t == 0) {
      const unsigned long long at = atomicAdd(a.counters + 2, 1ull);
      cache.stale_items[at] = item_index;
    }
    return;
  }
  const int nb = (n_alive < K ? n_alive : K) - 1;  // neighbours every vertex of the item blends
  const uint32_t lo = info.x;
  const uint32_t n_rows = info.y - info.x;
  // one group (the head vertex of a stretch), or a window wider than the staging area (most of
  // its control points are dead): rows read in place
  const bool tiny = count <= 32 || n_rows > max_rows;

  if (!tiny && t == 0) {
    // the window's rows, every copy plane the launch has room for (FP32 rows have one)
    const uint32_t n_copies = PREC == 0 ? cache.n_copies : 1u;
    expectTx(&S.bar, n_copies * n_rows * kRowBytes);
    for (uint32_t m = 0; m < n_copies; ++m) {
      bulkLoad(rows + static_cast<size_t>(m) * n_rows * kRowBytes,
               static_cast<const unsigned char*>(cache.pose_rows) +
                   (static_cast<uint64_t>(m) * cache.plane_rows + lo) * kRowBytes,
               n_rows * kRowBytes, &S.bar);
    }
  }
  // Validating mode: an item without records is known at once (one flag); whether its vertices
  // are still the ones the records were computed for is known only after all of them have been
  // read — that comparison waits until the blend is done (the barrier before the results go out
  // is there anyway), so that a valid item, the usual case, pays no barrier for being checked.
  // A stale item's blend read valid records of other coordinates: its output is garbage that
  // the stale-list pass overwrites.
  if (MODE == 1 && t == 0) {
    const bool missing = cache.item_ok[item_index] == 0u;
    S.stale = missing ? 1u : 0u;
    if (missing) {
      const unsigned long long at = atomicAdd(a.counters + 2, 1ull);
      cache.stale_items[at] = item_index;
    }
  }
  // ---- vertices in (coalesced), fingerprint
  unsigned long long fp = 0ull;
  for (int i = t; i < count; i += kBThreads) {
    const uint64_t ii = a.subset ? a.subset[p_base + i] : p_base + i;
    const float x = a.x[ii], y = a.y[ii], z = a.z[ii];
    if (MODE == 1) fp += vertexFingerprint(x, y, z, static_cast<uint32_t>(i), a.subset ? ii : 0ull);
    S.vx[i] = x;
    S.vy[i] = y;
    S.vz[i] = z;
  }
  if (MODE == 1) {
    fp = warpSumU64(fp);
    if (lane == 0) S.fp_w[warp] = fp;
  }
  __syncthreads();
  if (MODE == 1 && S.stale) {  // no records: nothing to blend from
    if (!tiny) {  // the copy must have landed before the CTA may leave or reuse the rows
      mbarWait(&S.bar, bar_phase);
      bar_phase ^= 1u;
    }
    return;
  }
  if (!tiny) {
    mbarWait(&S.bar, bar_phase);
    bar_phase ^= 1u;
  }

  if (tiny) blendVertices<PREC, 0, true>(a, cache, S, rows, p_base, count, nb, lo, n_rows);
  else if (nb == 5) blendVertices<PREC, 5, false>(a, cache, S, rows, p_base, count, nb, lo, n_rows);
  else blendVertices<PREC, 0, false>(a, cache, S, rows, p_base, count, nb, lo, n_rows);
  __syncthreads();
  if (MODE == 1 && t == 0) {
    unsigned long long f = 0ull;
#pragma unroll
    for (int w = 0; w < kBWarps; ++w) f += S.fp_w[w];
    if (cache.item_fp[item_index] != f) {
      const unsigned long long at = atomicAdd(a.counters + 2, 1ull);
      cache.stale_items[at] = item_index;
    }
  }
  // ---- results out (coalesced)
  for (int i = t; i < count; i += kBThreads) {
    const uint64_t ii = a.subset ? a.subset[p_base + i] : p_base + i;
    a.ox[ii] = S.vx[i];
    a.oy[ii] = S.vy[i];
    a.oz[ii] = S.vz[i];
    a.status[p_base + i] = kStatusDone;
  }
}

template <int PREC, int MODE>
__global__ void __launch_bounds__(kBThreads, PGMO_K3B_CTAS)
deformBlendCachedKernel(const DeformArgs a, const GridArgs ga, const AssocCache cache,
                        const WorkItem* __restrict__ items, const uint32_t first_item,
                        const uint32_t max_rows) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BlendSmemHead& S = *reinterpret_cast<BlendSmemHead*>(smem_raw);
  if (threadIdx.x == 0) mbarInit(&S.bar, 1);
  __syncthreads();
  uint32_t phase = 0;
  if (MODE != 2) {
    const uint32_t item_index = first_item + blockIdx.x;
    blendItem<PREC, MODE>(a, ga, cache, items[item_index], item_index, smem_raw, max_rows, phase);
  } else {
    const unsigned long long begin = a.counters[3], end = a.counters[2];
    for (unsigned long long i = begin + blockIdx.x; i < end; i += gridDim.x) {
      const uint32_t item_index = cache.stale_items[i];
      blendItem<PREC, 2>(a, ga, cache, items[item_index], item_index, smem_raw, max_rows, phase);
      __syncthreads();  // shared memory is reused by the next item
    }
  }
}


// ---- placement pass --------------------------------------------------------------------------
// Which copy of its pose row a neighbour field reads.  For one neighbour slot of the 32 records
// of a warp, per quarter warp (the unit an LDS.128 is served in): distinct rows are represented
// by their first lane; round m lets every representative still without a bank group ask for the
// group of its copy m, and the first asker of a free group gets it.  Representatives left over
// after the last round keep copy 0 (that quarter warp pays a second wavefront).  All four
// quarter warps run in the same instructions: match.any on (group, quarter) keys, one redux.or
// over one-hot group bytes.
__device__ __forceinline__ uint32_t placeSlot(uint32_t field, bool active, uint32_t lo,
                                              const AssocCache& cache, int lane) {
  const unsigned full = 0xFFFFFFFFu;
  const uint32_t q = static_cast<uint32_t>(lane) >> 3;
  const uint32_t rel = field & kCacheRelMask;
  const uint64_t row = static_cast<uint64_t>(lo) + rel;
  const uint32_t class0 = active ? (__ldg(cache.ctrl_class + row) & 7u) : 0u;
  // idle lanes get keys nobody shares (bit 31 / bit 30 + lane)
  const unsigned same_row = __match_any_sync(full, active ? (rel | (q << 16)) : (0x80000000u | lane));
  const int rep_lane = __ffs(same_row) - 1;
  const bool rep = active && lane == rep_lane;
  uint32_t cls[kBlendCopies];
#pragma unroll
  for (int m = 0; m < kBlendCopies; ++m) cls[m] = blendClass(class0, row, m);
  uint32_t taken = 0u;  // bank groups taken in this lane's quarter warp (one-hot byte)
  int copy = -1;
#pragma unroll
  for (int m = 0; m < kBlendCopies; ++m) {
    if (static_cast<uint32_t>(m) < cache.n_copies) {
      const bool asks = rep && copy < 0 && !((taken >> cls[m]) & 1u);
      const unsigned askers = __match_any_sync(full, asks ? (cls[m] | (q << 3)) : (0x40000000u | lane));
      const bool wins = asks && lane == __ffs(askers) - 1;
      if (wins) copy = m;
      const uint32_t all = __reduce_or_sync(full, wins ? (1u << cls[m]) << (8u * q) : 0u);
      taken |= (all >> (8u * q)) & 0xFFu;
    }
  }
  if (copy < 0) copy = 0;
  copy = __shfl_sync(full, copy, rep_lane);  // every lane of a row reads the same address
  const uint32_t c = copy == 0 ? cls[0] : (copy == 1 ? cls[1] : cls[2]);
  return rel | (static_cast<uint32_t>(copy) << kCacheCopyShift) | (c << kCacheClassShift);
}

__device__ __forceinline__ void placeItem(const DeformArgs& a, const GridArgs& ga,
                                          const AssocCache& cache, const WorkItem& item) {
  const uint32_t r = item.run;
  const uint4 info = __ldg(a.run_info + r);
  if (info.w) return;
  const int n_alive = static_cast<int>(ga.grids[r].n_alive);
  if (n_alive > kSortedMaxPoints || !runCacheable(info)) return;  // no records (see blendItem)
  const int K = static_cast<int>(info.z) + 2;
  const int nb = (n_alive < K ? n_alive : K) - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int count = static_cast<int>(item.count);
  uint4* rec0 = reinterpret_cast<uint4*>(cache.rec);
  for (int g = warp; g * 32 < count; g += n_warps) {
    const int q = g * 32 + lane;
    const bool active = q < count;
    const uint64_t slot = item.start + static_cast<uint64_t>(active ? q : 0);
    uint4 r0 = rec0[slot];
    uint32_t f[7] = {r0.x >> 16, r0.y & 0xFFFFu, r0.y >> 16, r0.z & 0xFFFFu, r0.z >> 16,
                     r0.w & 0xFFFFu, r0.w >> 16};
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j < nb) f[j] = placeSlot(f[j], active, info.x, cache, lane);
    }
    if (active) {
      r0.x = (r0.x & 0xFFFFu) | (f[0] << 16);
      r0.y = f[1] | (f[2] << 16);
      r0.z = f[3] | (f[4] << 16);
      r0.w = f[5] | (f[6] << 16);
      rec0[slot] = r0;
    }
    if (nb > 7) {  // plane 1: neighbours 7..14
      uint4 r1 = rec0[cache.n_slots + slot];
      uint32_t h[8] = {r1.x & 0xFFFFu, r1.x >> 16, r1.y & 0xFFFFu, r1.y >> 16,
                       r1.z & 0xFFFFu, r1.z >> 16, r1.w & 0xFFFFu, r1.w >> 16};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (7 + j < nb) h[j] = placeSlot(h[j], active, info.x, cache, lane);
      }
      if (active) {
        rec0[cache.n_slots + slot] = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16),
                                                h[4] | (h[5] << 16), h[6] | (h[7] << 16));
      }
    }
  }
}

constexpr int kPlaceThreads = 256;
template <bool FROM_LIST>
__global__ void __launch_bounds__(kPlaceThreads)
placeRecordsKernel(const DeformArgs a, const GridArgs ga, const AssocCache cache,
                   const WorkItem* __restrict__ items, const uint32_t first_item) {
  if (!FROM_LIST) {
    placeItem(a, ga, cache, items[first_item + blockIdx.x]);
  } else {
    const unsigned long long begin = a.counters[3], end = a.counters[2];
    for (unsigned long long i = begin + blockIdx.x; i < end; i += gridDim.x) {
      placeItem(a, ga, cache, items[cache.stale_items[i]]);
    }
  }
}

uint32_t blendRows(uint64_t widest_window) {
  return static_cast<uint32_t>(widest_window < static_cast<uint64_t>(kSortedMaxPoints)
                                   ? widest_window : static_cast<uint64_t>(kSortedMaxPoints));
}
size_t blendSmemBytes(int precision, uint64_t widest_window, uint32_t n_copies) {
  return sizeof(BlendSmemHead) +
         static_cast<size_t>(blendRows(widest_window)) * (precision == 0 ? 128u * n_copies : 64u);
}

template <int PREC, int MODE>
cudaError_t launchBlend(const DeformArgs& args, const GridArgs& gargs, const AssocCache& cache,
                        const WorkItem* items, uint32_t first_item, uint32_t n_items,
                        uint64_t widest_window, int n_sms, cudaStream_t stream) {
  static thread_local int configured_device = -1;
  int device = 0;
  cudaGetDevice(&device);
  if (configured_device != device) {
    const cudaError_t e = cudaFuncSetAttribute(
        deformBlendCachedKernel<PREC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(std::max(kBSmemBudget, blendSmemBytes(PREC, kSortedMaxPoints, 1))));
    if (e != cudaSuccess) return e;
    configured_device = device;
  }
  const size_t smem = blendSmemBytes(PREC, widest_window, PREC == 0 ? cache.n_copies : 1u);
  const unsigned grid = MODE == 2 ? static_cast<unsigned>(n_sms * PGMO_K3B_CTAS) : n_items;
  deformBlendCachedKernel<PREC, MODE><<<grid, kBThreads, smem, stream>>>(args, gargs, cache, items,
                                                                        first_item, blendRows(widest_window));
  return cudaGetLastError();
}

template <int PREC>
cudaError_t launchBlendMode(const DeformArgs& args, const GridArgs& gargs, const AssocCache& cache,
                            const WorkItem* items, uint32_t first_item, uint32_t n_items,
                            uint64_t widest_window, int mode, int n_sms, cudaStream_t stream) {
  switch (mode) {
    case 0: return launchBlend<PREC, 0>(args, gargs, cache, items, first_item, n_items, widest_window, n_sms, stream);
    case 1: return launchBlend<PREC, 1>(args, gargs, cache, items, first_item, n_items, widest_window, n_sms, stream);
    default: return launchBlend<PREC, 2>(args, gargs, cache, items, first_item, n_items, widest_window, n_sms, stream);
  }
}

template <int PREC>
void preloadBlend() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, deformBlendCachedKernel<PREC, 0>);
  cudaFuncGetAttributes(&attr, deformBlendCachedKernel<PREC, 1>);
  cudaFuncGetAttributes(&attr, deformBlendCachedKernel<PREC, 2>);
}

}  // namespace

void preloadDeformBlendKernels() {
  preloadBlend<0>();
  preloadBlend<1>();
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, placeRecordsKernel<false>);
  cudaFuncGetAttributes(&attr, placeRecordsKernel<true>);
}

uint32_t blendCopiesFor(int precision, uint64_t widest_window) {
  if (precision != 0) return 1u;
  const size_t per_copy = static_cast<size_t>(blendRows(widest_window)) * 128u;
  if (per_copy == 0) return 1u;
  const size_t room = kBSmemBudget > sizeof(BlendSmemHead) ? kBSmemBudget - sizeof(BlendSmemHead) : 0;
  const size_t fit = room / per_copy;
  return static_cast<uint32_t>(fit < 1 ? 1 : (fit > static_cast<size_t>(kBlendCopies) ? kBlendCopies : fit));
}

cudaError_t launchPlaceRecords(const DeformArgs& args, const GridArgs& gargs, const AssocCache& cache,
                               const WorkItem* items, uint32_t first_item, uint32_t n_items,
                               bool from_stale_list, int n_sms, cudaStream_t stream) {
  if (cache.n_copies <= 1) return cudaSuccess;
  if (from_stale_list) {
    placeRecordsKernel<true><<<static_cast<unsigned>(n_sms * 4), kPlaceThreads, 0, stream>>>(
        args, gargs, cache, items, first_item);
  } else {
    if (n_items == 0) return cudaSuccess;
    placeRecordsKernel<false><<<n_items, kPlaceThreads, 0, stream>>>(args, gargs, cache, items, first_item);
  }
  return cudaGetLastError();
}

cudaError_t launchDeformBlendCached(const DeformArgs& args, const GridArgs& gargs,
                                    const AssocCache& cache, const WorkItem* items,
                                    uint32_t first_item, uint32_t n_items, int precision,
                                    uint32_t max_k, uint32_t min_k, uint64_t widest_window,
                                    int mode, int n_sms, cudaStream_t stream) {
  if (n_items == 0 && mode != 2) return cudaSuccess;
  (void)max_k;
  (void)min_k;
  if (precision == 0) {
    return launchBlendMode<0>(args, gargs, cache, items, first_item, n_items, widest_window, mode, n_sms, stream);
  }
  return launchBlendMode<1>(args, gargs, cache, items, first_item, n_items, widest_window, mode, n_sms, stream);
}

}  // namespace pgmo
