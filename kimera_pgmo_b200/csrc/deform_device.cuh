// Device helpers shared by the deformation kernels: the reference's arithmetic, bit for bit.
#pragma once

#include <cfloat>

#include "deform_kernels.cuh"

namespace pgmo {
namespace dev {

__device__ __forceinline__ float dist2Exact(float ax, float ay, float az, float bx, float by,
                                            float bz) {
  const float dx = __fsub_rn(ax, bx);
  const float dy = __fsub_rn(ay, by);
  const float dz = __fsub_rn(az, bz);
  return __fadd_rn(__fmul_rn(dx, dx), __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz)));
}

// hands local position p to the PCL-order resolver (K4); slot = where its association record
// lives (kNoCacheSlot: the vertex has none, the resolver's own blend is final)
__device__ __forceinline__ void sendToResolver(const DeformArgs& a, uint64_t p,
                                               uint32_t slot = kNoCacheSlot) {
  a.status[p] = kStatusNeedsTie;
  const unsigned long long at = atomicAdd(a.counters, 1ull);
  a.tie_list[at] = static_cast<uint32_t>(p);
  a.tie_slot[at] = slot;
}

// interpPoint's weights (mesh_deformation.cpp:85-94) for neighbour j of n, distances ascending:
// d_max = double(sqrtf(d2[n-1])); 1 if sqrtf(d2[0]) == d_max or d_max == 0, else
// 1 - double(sqrtf(d2[j])) / d_max.
__device__ __forceinline__ double weightFp64(float d2_j, float d_max_f, bool const_w) {
  return const_w ? 1.0
                 : __dsub_rn(1.0, __ddiv_rn(static_cast<double>(__fsqrt_rn(d2_j)),
                                            static_cast<double>(d_max_f)));
}
__device__ __forceinline__ bool constWeight(float d2_first, float d_max_f) {
  return (__fsqrt_rn(d2_first) == d_max_f) || d_max_f == 0.0f;
}
// FP32 mode: same formula in float with a reciprocal (no FMA: the cached value must not depend
// on the kernel it was computed in)
__device__ __forceinline__ float weightFp32(float d2_j, float inv_dmax, bool const_w) {
  return const_w ? 1.0f : __fsub_rn(1.0f, __fmul_rn(__fsqrt_rn(d2_j), inv_dmax));
}

// Fingerprint of a work item's vertices: the wrapping sum over its vertices of a 64-bit mix of
// the coordinate bits, the offset inside the item and (with an indices list) the vertex index.
// The association records of an item are valid for exactly the vertex bits they were computed
// from; K3b recomputes the sum from the vertices it is given and compares.
__device__ __forceinline__ unsigned long long vertexFingerprint(float x, float y, float z,
                                                                uint32_t offset, uint64_t index) {
  unsigned long long m = (static_cast<unsigned long long>(__float_as_uint(y)) << 32) | __float_as_uint(x);
  m = (m + 0x9e3779b97f4a7c15ull * (offset + 1u)) * 0xbf58476d1ce4e5b9ull;
  m ^= m >> 29;
  m = (m ^ __float_as_uint(z) ^ (index * 0x2545f4914f6cdd1dull)) * 0x94d049bb133111ebull;
  m ^= m >> 32;
  return m;
}
__device__ __forceinline__ unsigned long long warpSumU64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// Bank class of copy `copy` of pose row `row` (class0 = the colouring's class of the row): the three
// copies of a row sit under three different XOR masks, so that a quarter warp that needs several
// rows of one class can read some of them from another bank group.
__host__ __device__ __forceinline__ uint32_t blendClass(uint32_t class0, uint64_t row, int copy) {
  if (copy == 0) return class0 & 7u;
  const uint32_t a = 1u + static_cast<uint32_t>(row % 7u);                  // 1..7
  if (copy == 1) return (class0 ^ a) & 7u;
  const uint32_t b = 1u + (a + static_cast<uint32_t>((row / 7u) % 6u)) % 7u;  // 1..7, != a
  return (class0 ^ b) & 7u;
}

// can the vertices of a run {lo, hi, k, skip} have association records?  (neighbour count and
// the 16-bit relative control index of the record layout)
__device__ __forceinline__ bool runCacheable(const uint4& info) {
  return static_cast<int>(info.z) + 1 <= kCacheMaxNbr && info.y - info.x <= (1u << kCacheRelBits);
}
// record slot `slot` of the association cache: neighbour j's control index relative to run.lo
// (| bank class << kCacheClassShift; copy 0 until the placement pass has seen the record)
__device__ __forceinline__ void cacheStoreNbr(const AssocCache& c, uint64_t slot, int j, uint32_t rel) {
  const uint64_t plane = j < 7 ? 0 : 1;
  c.rec[(plane * c.n_slots + slot) * 8 + (j < 7 ? 1 + j : j - 7)] = static_cast<uint16_t>(rel);
}
__device__ __forceinline__ void cacheStoreWeight(const AssocCache& c, uint64_t slot, int j, int precision,
                                                 double w64, float w32) {
  if (precision == 0) static_cast<double*>(c.weight)[static_cast<uint64_t>(j) * c.n_slots + slot] = w64;
  else static_cast<float*>(c.weight)[static_cast<uint64_t>(j) * c.n_slots + slot] = w32;
}

__device__ __forceinline__ uint32_t findRun(const uint64_t* __restrict__ run_start,
                                            uint32_t n_runs, uint64_t position) {
  // last r with run_start[r] <= position
  uint32_t lo = 0, hi = n_runs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(run_start + mid) <= position) lo = mid;
    else hi = mid;
  }
  return lo;
}

// interpPoint's blend in the reference's arithmetic.  nbr_d2 / nbr_idx: n neighbours ascending.
__device__ __forceinline__ void blendFp64(const DeformArgs& a, float qx, float qy, float qz,
                                          const float* nbr_d2, const int* nbr_idx, int n,
                                          double out[3]) {
  const float d_max_f = __fsqrt_rn(nbr_d2[n - 1]);
  const double d_max = static_cast<double>(d_max_f);
  const bool const_w = (__fsqrt_rn(nbr_d2[0]) == d_max_f) || d_max == 0.0;
  double wsum = 0.0, acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  const double v0 = static_cast<double>(qx), v1 = static_cast<double>(qy),
               v2 = static_cast<double>(qz);
  for (int j = 0; j < n - 1; ++j) {
    const int c = nbr_idx[j];
    const double w =
        const_w ? 1.0
                : __dsub_rn(1.0, __ddiv_rn(static_cast<double>(__fsqrt_rn(nbr_d2[j])), d_max));
    wsum = __dadd_rn(wsum, w);
    const double* g = a.ctrl_pos + 3ull * c;
    const double* T = a.poses + 12ull * c;
    const double p0 = __dsub_rn(v0, g[0]);
    const double p1 = __dsub_rn(v1, g[1]);
    const double p2 = __dsub_rn(v2, g[2]);
    const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(T[0], p0), __dmul_rn(T[1], p1)), __dmul_rn(T[2], p2));
    const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(T[3], p0), __dmul_rn(T[4], p1)), __dmul_rn(T[5], p2));
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(T[6], p0), __dmul_rn(T[7], p1)), __dmul_rn(T[8], p2));
    acc0 = __dadd_rn(acc0, __dmul_rn(w, __dadd_rn(r0, T[9])));
    acc1 = __dadd_rn(acc1, __dmul_rn(w, __dadd_rn(r1, T[10])));
    acc2 = __dadd_rn(acc2, __dmul_rn(w, __dadd_rn(r2, T[11])));
  }
  out[0] = __ddiv_rn(acc0, wsum);
  out[1] = __ddiv_rn(acc1, wsum);
  out[2] = __ddiv_rn(acc2, wsum);
}

// FP32 mode: same neighbours, same weight formula, float arithmetic (FMA allowed).
__device__ __forceinline__ void blendFp32(const DeformArgs& a, float qx, float qy, float qz,
                                          const float* nbr_d2, const int* nbr_idx, int n,
                                          float out[3]) {
  const float d_max = __fsqrt_rn(nbr_d2[n - 1]);
  const bool const_w = (__fsqrt_rn(nbr_d2[0]) == d_max) || d_max == 0.0f;
  const float inv_dmax = const_w ? 0.0f : __frcp_rn(d_max);
  float wsum = 0.f, acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  for (int j = 0; j < n - 1; ++j) {
    const int c = nbr_idx[j];
    const float w = const_w ? 1.0f : 1.0f - __fsqrt_rn(nbr_d2[j]) * inv_dmax;
    wsum += w;
    const float4* T = reinterpret_cast<const float4*>(a.ctrl_pose_f32 + 16ull * c);
    const float4 t0 = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2), t3 = __ldg(T + 3);
    // layout: R0..R8, t0..t2, g0..g2, pad
    const float p0 = qx - t3.x, p1 = qy - t3.y, p2 = qz - t3.z;
    const float r0 = t0.x * p0 + t0.y * p1 + t0.z * p2 + t2.y;
    const float r1 = t0.w * p0 + t1.x * p1 + t1.y * p2 + t2.z;
    const float r2 = t1.z * p0 + t1.w * p1 + t2.x * p2 + t2.w;
    acc0 += w * r0;
    acc1 += w * r1;
    acc2 += w * r2;
  }
  out[0] = acc0 / wsum;
  out[1] = acc1 / wsum;
  out[2] = acc2 / wsum;
}

__device__ __forceinline__ void writeResult(const DeformArgs& a, uint64_t p, uint64_t ii,
                                            int precision, float qx, float qy, float qz,
                                            const float* nbr_d2, const int* nbr_idx, int n) {
  if (precision == 0) {
    double o[3];
    blendFp64(a, qx, qy, qz, nbr_d2, nbr_idx, n, o);
    a.ox[ii] = __double2float_rn(o[0]);
    a.oy[ii] = __double2float_rn(o[1]);
    a.oz[ii] = __double2float_rn(o[2]);
    if (a.of64) {
      a.of64[3 * p + 0] = o[0];
      a.of64[3 * p + 1] = o[1];
      a.of64[3 * p + 2] = o[2];
    }
  } else {
    float o[3];
    blendFp32(a, qx, qy, qz, nbr_d2, nbr_idx, n, o);
    a.ox[ii] = o[0];
    a.oy[ii] = o[1];
    a.oz[ii] = o[2];
    if (a.of64) {
      a.of64[3 * p + 0] = o[0];
      a.of64[3 * p + 1] = o[1];
      a.of64[3 * p + 2] = o[2];
    }
  }
  if (a.assoc) {
    for (uint32_t j = 0; j < a.assoc_stride; ++j) {
      a.assoc[p * a.assoc_stride + j] = (static_cast<int>(j) < n - 1) ? nbr_idx[j] : 0xFFFFFFFFu;
    }
  }
  if (a.assoc_cnt) a.assoc_cnt[p] = static_cast<uint8_t>(n - 1 > 255 ? 255 : n - 1);
}

// Register-resident variants for the K3 kernels: every loop is fully unrolled over the
// compile-time capacity and guarded by the runtime neighbour count, so that the neighbour
// arrays are never indexed dynamically (dynamic indexing would push them to local memory).
template <int KCAP>
__device__ __forceinline__ void writeResultReg(const DeformArgs& a, uint64_t p, uint64_t ii,
                                               int precision, float qx, float qy, float qz,
                                               const float (&nbr_d2)[KCAP],
                                               const int (&nbr_idx)[KCAP], int n) {
  // nbr_d2[n - 1] without dynamic indexing: distances ascend, so it is the max of the first n
  float last_d2 = nbr_d2[0];
#pragma unroll
  for (int j = 1; j < KCAP; ++j) {
    last_d2 = j < n ? fmaxf(last_d2, nbr_d2[j]) : last_d2;
  }
  if (precision == 0) {
    const float d_max_f = __fsqrt_rn(last_d2);
    const double d_max = static_cast<double>(d_max_f);
    const bool const_w = (__fsqrt_rn(nbr_d2[0]) == d_max_f) || d_max == 0.0;
    double wsum = 0.0, acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    const double v0 = static_cast<double>(qx), v1 = static_cast<double>(qy),
                 v2 = static_cast<double>(qz);
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
      if (j < n - 1) {
        const int c = nbr_idx[j];
        const double w =
            const_w ? 1.0
                    : __dsub_rn(1.0, __ddiv_rn(static_cast<double>(__fsqrt_rn(nbr_d2[j])), d_max));
        wsum = __dadd_rn(wsum, w);
        const double* g = a.ctrl_pos + 3ull * c;
        const double* T = a.poses + 12ull * c;
        const double p0 = __dsub_rn(v0, g[0]);
        const double p1 = __dsub_rn(v1, g[1]);
        const double p2 = __dsub_rn(v2, g[2]);
        const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(T[0], p0), __dmul_rn(T[1], p1)), __dmul_rn(T[2], p2));
        const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(T[3], p0), __dmul_rn(T[4], p1)), __dmul_rn(T[5], p2));
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(T[6], p0), __dmul_rn(T[7], p1)), __dmul_rn(T[8], p2));
        acc0 = __dadd_rn(acc0, __dmul_rn(w, __dadd_rn(r0, T[9])));
        acc1 = __dadd_rn(acc1, __dmul_rn(w, __dadd_rn(r1, T[10])));
        acc2 = __dadd_rn(acc2, __dmul_rn(w, __dadd_rn(r2, T[11])));
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
        const int c = nbr_idx[j];
        const float w = const_w ? 1.0f : 1.0f - __fsqrt_rn(nbr_d2[j]) * inv_dmax;
        wsum += w;
        const float4* T = reinterpret_cast<const float4*>(a.ctrl_pose_f32 + 16ull * c);
        const float4 t0 = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2), t3 = __ldg(T + 3);
        const float p0 = qx - t3.x, p1 = qy - t3.y, p2 = qz - t3.z;
        acc0 += w * (t0.x * p0 + t0.y * p1 + t0.z * p2 + t2.y);
        acc1 += w * (t0.w * p0 + t1.x * p1 + t1.y * p2 + t2.z);
        acc2 += w * (t1.z * p0 + t1.w * p1 + t2.x * p2 + t2.w);
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
        a.assoc[p * a.assoc_stride + j] = (j < n - 1) ? static_cast<uint32_t>(nbr_idx[j]) : 0xFFFFFFFFu;
      }
    }
    for (uint32_t j = KCAP; j < a.assoc_stride; ++j) a.assoc[p * a.assoc_stride + j] = 0xFFFFFFFFu;
  }
  if (a.assoc_cnt) a.assoc_cnt[p] = static_cast<uint8_t>(n - 1 > 255 ? 255 : n - 1);
}

}  // namespace dev
}  // namespace pgmo
