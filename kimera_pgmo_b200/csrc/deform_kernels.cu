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
#include "deform_resolve.cuh"
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
// K4 kernel: one flagged vertex per warp (resolveOne: deform_resolve.cuh)
__global__ void __launch_bounds__(kResolveWarpsPerBlock * 32)
deformResolveKernel(const DeformArgs a, const ResolveArgs ra, const AssocCache cache,
                    const int use_cache, const int precision) {
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
               lane, use_cache ? &cache : nullptr, a.tie_slot[i]);
  }
}

__global__ void deformStepDoneKernel(unsigned long long* counters) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    counters[1] = counters[0];
    counters[3] = counters[2];
  }
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

// Eight threads per control point, one 16-byte chunk of its rows each: the 128-byte rows of the
// blend table (and the 64-byte rows of the FP32 table) are written as whole, coalesced lines.
// Only the first n_planes copy planes of the blend table are written (the ones the blend kernel
// stages: deform_blend.cu blendCopiesFor).
__global__ void buildControlTablesKernel(uint64_t n_c, const double* __restrict__ ctrl_pos,
                                         const double* __restrict__ poses,
                                         float4* __restrict__ ctrl_f4,
                                         float* __restrict__ pose_f32,
                                         double* __restrict__ pose_rows64,
                                         const uint8_t* __restrict__ ctrl_class,
                                         int positions_changed, int n_planes) {
  const uint64_t gid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t j = gid >> 3;
  const int c = static_cast<int>(gid & 7u);
  if (j >= n_c) return;
  // row = R(9) t(3) g(3) pad; this thread's elements 2c and 2c + 1
  auto element = [&](int i) -> double {
    if (i < 12) return poses ? poses[12 * j + i] : 0.0;
    if (i < 15) return ctrl_pos[3 * j + (i - 12)];
    return 0.0;
  };
  const double e0 = element(2 * c), e1 = element(2 * c + 1);
  if (positions_changed && c == 0) {
    // .w (the schedule's del_run) is rewritten by applyScheduleKernel after the next build
    ctrl_f4[j] = make_float4(__double2float_rn(ctrl_pos[3 * j]), __double2float_rn(ctrl_pos[3 * j + 1]),
                             __double2float_rn(ctrl_pos[3 * j + 2]), __int_as_float(0));
  }
  reinterpret_cast<float2*>(pose_f32 + 16 * j)[c] = make_float2(__double2float_rn(e0), __double2float_rn(e1));
  if (pose_rows64) {
    // chunk c stored at c ^ class so that lanes reading the same chunk of the rows of control
    // points that lie close together spread over the shared-memory banks
    const uint32_t class0 = ctrl_class ? (ctrl_class[j] & 7u) : 0u;
    for (int m = 0; m < n_planes; ++m) {  // one plane of n_c rows per copy
      double2* dst = reinterpret_cast<double2*>(pose_rows64 + 16 * (static_cast<uint64_t>(m) * n_c + j));
      const int sw = static_cast<int>(dev::blendClass(class0, j, m));
      dst[c ^ sw] = make_double2(e0, e1);
    }
  }
}

// Pose rows [row_begin, row_end) pulled from `src` — the table one device received the solve
// result in, possibly a PEER device's memory mapped over NVLink — and turned into this device's
// tables in the same kernel: the local f64 copy (tie resolver, grid-walk kernel), the f32 rows
// and the chunk-swizzled f64 rows of the blend kernel.  A block takes 128 rows: the 12 x 128
// doubles are read flat (consecutive threads, consecutive 16-byte words: full NVLink / HBM
// sectors) into shared memory, then one thread builds one row.
constexpr int kPullRows = 128;
__global__ void __launch_bounds__(kPullRows)
pullBuildPoseTablesKernel(uint64_t row_begin, uint64_t row_end, uint64_t n_c,
                          const double* __restrict__ src,
                          double* __restrict__ poses_local, const double* __restrict__ ctrl_pos,
                          float* __restrict__ pose_f32, double* __restrict__ pose_rows64,
                          const uint8_t* __restrict__ ctrl_class, int n_planes) {
  __shared__ double2 stage[kPullRows * 6];
  const uint64_t r0 = row_begin + static_cast<uint64_t>(blockIdx.x) * kPullRows;
  const uint64_t n_rows = row_end - r0 < static_cast<uint64_t>(kPullRows) ? row_end - r0 : kPullRows;
  const double2* s2 = reinterpret_cast<const double2*>(src + 12 * r0);     // 96-byte rows: 16-byte aligned
  double2* l2 = poses_local ? reinterpret_cast<double2*>(poses_local + 12 * r0) : nullptr;
  for (uint32_t e = threadIdx.x; e < n_rows * 6; e += kPullRows) {
    const double2 v = s2[e];
    stage[e] = v;
    if (l2 && l2 != s2) l2[e] = v;
  }
  __syncthreads();
  // one 16-byte chunk per thread and step: rows leave as whole, coalesced lines
  const double* staged = reinterpret_cast<const double*>(stage);
  for (uint32_t e = threadIdx.x; e < n_rows * 8; e += kPullRows) {
    const uint32_t local = e >> 3;
    const int c = static_cast<int>(e & 7u);
    const uint64_t j = r0 + local;
    auto element = [&](int i) -> double {
      if (i < 12) return staged[12 * local + i];
      if (i < 15) return ctrl_pos[3 * j + (i - 12)];
      return 0.0;
    };
    const double e0 = element(2 * c), e1 = element(2 * c + 1);
    reinterpret_cast<float2*>(pose_f32 + 16 * j)[c] = make_float2(__double2float_rn(e0), __double2float_rn(e1));
    const uint32_t class0 = ctrl_class ? (ctrl_class[j] & 7u) : 0u;
    for (int m = 0; m < n_planes; ++m) {
      double2* dst = reinterpret_cast<double2*>(pose_rows64 + 16 * (static_cast<uint64_t>(m) * n_c + j));
      const int sw = static_cast<int>(dev::blendClass(class0, j, m));
      dst[c ^ sw] = make_double2(e0, e1);
    }
  }
}

__global__ void applyScheduleKernel(uint64_t n_c, const int32_t* __restrict__ del_run,
                                    const int32_t* __restrict__ stable_key3,
                                    float4* __restrict__ ctrl_f4, int4* __restrict__ ctrl_key) {
  const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_c) return;
  ctrl_f4[j].w = __int_as_float(del_run[j]);
  ctrl_key[j] = make_int4(stable_key3[3 * j], stable_key3[3 * j + 1], stable_key3[3 * j + 2], 0);
}

// ------------------------------------------------------------------------------------------
// Device-resident meshes: the compressors keep their vertices as float4 {x, y, z, rgba bits}
// (the reference's PointXYZRGBA cloud, deformation_graph.cpp:541,557 converts blob <-> cloud on
// the host); K3 streams SoA.  Both directions are plain coalesced copies (28 B / vertex each).
__global__ void aosToSoaKernel(uint64_t n, const float4* __restrict__ in, float* __restrict__ x,
                               float* __restrict__ y, float* __restrict__ z) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = in[i];
  x[i] = v.x;
  y[i] = v.y;
  z[i] = v.z;
}

// out[i] = {deformed xyz, in[i].w} for processed vertices, in[i] otherwise (in == out allowed)
__global__ void soaToAosKernel(uint64_t n, const float4* in, const float* __restrict__ x,
                               const float* __restrict__ y, const float* __restrict__ z,
                               const uint8_t* __restrict__ status, float4* out) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = in[i];
  if (status[i] == kStatusDone) {
    v.x = x[i];
    v.y = y[i];
    v.z = z[i];
  }
  out[i] = v;
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

cudaError_t launchDeformResolve(const DeformArgs& args, const ResolveArgs& rargs,
                                const AssocCache* cache, int precision, cudaStream_t stream) {
  if (args.n_local == 0) return cudaSuccess;
  const unsigned blocks = rargs.n_warps / kResolveWarpsPerBlock;
  deformResolveKernel<<<blocks, kResolveWarpsPerBlock * 32, 0, stream>>>(
      args, rargs, cache ? *cache : AssocCache{}, cache ? 1 : 0, precision);
  return cudaGetLastError();
}

cudaError_t launchDeformStepDone(unsigned long long* counters, cudaStream_t stream) {
  deformStepDoneKernel<<<1, 32, 0, stream>>>(counters);
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
                                     float4* ctrl_f4, float* ctrl_pose_f32, double* pose_rows64,
                                     const uint8_t* ctrl_class, bool positions_changed,
                                     int n_planes, cudaStream_t stream) {
  if (n_c == 0) return cudaSuccess;
  const unsigned threads = 256;
  buildControlTablesKernel<<<static_cast<unsigned>((8 * n_c + threads - 1) / threads), threads, 0,
                             stream>>>(n_c, ctrl_pos, poses, ctrl_f4, ctrl_pose_f32, pose_rows64,
                                       ctrl_class, positions_changed ? 1 : 0, n_planes);
  return cudaGetLastError();
}

// ---- sequenced pose source: the ordering between the writer of a solve result and the devices
// that pull it is kept by flags inside the source's own memory (sequence numbers, system-scope
// release / acquire over NVLink) instead of a collective between the processes.
namespace {
__device__ __forceinline__ unsigned long long loadAcquireSys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void storeReleaseSys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globalTimerNs() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// waits until *p >= need; gives up after kPoseWaitNs and counts the timeout (the caller's data is
// then whatever the buffer holds: pgmo_deform_pose_source_timeouts reports it)
constexpr unsigned long long kPoseWaitNs = 2000000000ull;
__device__ __forceinline__ void waitAtLeast(const unsigned long long* p, unsigned long long need,
                                            unsigned long long* timeouts) {
  const unsigned long long t0 = globalTimerNs();
  while (loadAcquireSys(p) < need) {
    if (globalTimerNs() - t0 > kPoseWaitNs) {
      atomicAdd(timeouts, 1ull);
      return;
    }
    __nanosleep(100);
  }
}
}  // namespace

// writer, before it overwrites a buffer: every reader has acknowledged the buffer's old content
__global__ void poseSourceWaitAcksKernel(const unsigned long long* __restrict__ acks, uint32_t n_readers,
                                         unsigned long long need, unsigned long long* timeouts) {
  if (threadIdx.x < n_readers) waitAtLeast(acks + threadIdx.x, need, timeouts);
}
__global__ void poseSourceStoreKernel(unsigned long long* flag, unsigned long long seq) {
  __threadfence_system();
  storeReleaseSys(flag, seq);
}
// reader, before it pulls: the buffer holds sequence number `seq` (or a later one)
__global__ void poseSourceWaitReadyKernel(const unsigned long long* ready, unsigned long long seq,
                                          unsigned long long* timeouts) {
  waitAtLeast(ready, seq, timeouts);
}

cudaError_t launchPoseSourceWaitAcks(const unsigned long long* acks, uint32_t n_readers,
                                     unsigned long long need, unsigned long long* timeouts,
                                     cudaStream_t stream) {
  if (n_readers == 0 || need == 0) return cudaSuccess;
  poseSourceWaitAcksKernel<<<1, 64, 0, stream>>>(acks, n_readers, need, timeouts);
  return cudaGetLastError();
}
cudaError_t launchPoseSourceStore(unsigned long long* flag, unsigned long long seq, cudaStream_t stream) {
  poseSourceStoreKernel<<<1, 1, 0, stream>>>(flag, seq);
  return cudaGetLastError();
}
cudaError_t launchPoseSourceWaitReady(const unsigned long long* ready, unsigned long long seq,
                                      unsigned long long* timeouts, cudaStream_t stream) {
  poseSourceWaitReadyKernel<<<1, 1, 0, stream>>>(ready, seq, timeouts);
  return cudaGetLastError();
}

cudaError_t launchPullBuildPoseTables(uint64_t row_begin, uint64_t row_end, uint64_t n_c,
                                      const double* src, double* poses_local, const double* ctrl_pos,
                                      float* ctrl_pose_f32, double* pose_rows64,
                                      const uint8_t* ctrl_class, int n_planes, cudaStream_t stream) {
  if (row_end <= row_begin) return cudaSuccess;
  const uint64_t blocks = (row_end - row_begin + kPullRows - 1) / kPullRows;
  pullBuildPoseTablesKernel<<<static_cast<unsigned>(blocks), kPullRows, 0, stream>>>(
      row_begin, row_end, n_c, src, poses_local, ctrl_pos, ctrl_pose_f32, pose_rows64, ctrl_class,
      n_planes);
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

cudaError_t launchAosToSoa(uint64_t n, const float4* in, float* x, float* y, float* z,
                           cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  aosToSoaKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(n, in, x, y, z);
  return cudaGetLastError();
}

cudaError_t launchSoaToAos(uint64_t n, const float4* in, const float* x, const float* y,
                           const float* z, const uint8_t* status, float4* out,
                           cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  soaToAosKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(n, in, x, y, z,
                                                                            status, out);
  return cudaGetLastError();
}

}  // namespace pgmo

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadDeformMiscKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, stampStretchKernel);
  cudaFuncGetAttributes(&attr, deformResolveKernel);
  cudaFuncGetAttributes(&attr, deformStepDoneKernel);
  cudaFuncGetAttributes(&attr, predeformKernel);
  cudaFuncGetAttributes(&attr, buildControlTablesKernel);
  cudaFuncGetAttributes(&attr, pullBuildPoseTablesKernel);
  cudaFuncGetAttributes(&attr, applyScheduleKernel);
  cudaFuncGetAttributes(&attr, aosToSoaKernel);
  cudaFuncGetAttributes(&attr, soaToAosKernel);
  cudaGetLastError();
}
}  // namespace pgmo
