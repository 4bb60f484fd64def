// Implicit-octree kernels of OctreeCompression (see oct_kernels.cuh for the design).
#include "oct_kernels.cuh"

#include <cfloat>

#include "deform_device.cuh"  // dist2Exact: (a - b).squaredNorm() as PCL / Eigen evaluate it

namespace pgmo {

namespace {

using dev::dist2Exact;

constexpr long long kCoordOffset = 1ll << (kOctCoordBits - 1);
constexpr long long kCoordLimit = 1ll << kOctCoordBits;

__device__ __forceinline__ bool packKey(int h, long long x, long long y, long long z,
                                        unsigned long long* key) {
  x += kCoordOffset;
  y += kCoordOffset;
  z += kCoordOffset;
  if (x < 0 || y < 0 || z < 0 || x >= kCoordLimit || y >= kCoordLimit || z >= kCoordLimit) {
    return false;
  }
  *key = (static_cast<unsigned long long>(h) << (3 * kOctCoordBits)) |
         (static_cast<unsigned long long>(x) << (2 * kOctCoordBits)) |
         (static_cast<unsigned long long>(y) << kOctCoordBits) | static_cast<unsigned long long>(z);
  return true;
}

__device__ __forceinline__ uint32_t hashKey(unsigned long long k) {
  k ^= k >> 30;
  k *= 0xbf58476d1ce4e5b9ull;
  k ^= k >> 27;
  k *= 0x94d049bb133111ebull;
  k ^= k >> 31;
  return static_cast<uint32_t>(k);
}

// Test before test-and-set: the nodes near the root are shared by every point of a build, and
// same-address atomics serialise in L2.  A key never changes once set and minseq only falls, so
// a plain (volatile: L2) read that already shows this key / a small enough minseq makes the
// atomic redundant; a stale read only costs the atomic it would have saved.
__device__ __forceinline__ uint32_t tableInsert(const OctTable& t, unsigned long long key,
                                                uint32_t seq) {
  uint32_t slot = hashKey(key) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(t.keys + slot);
    if (k == kOctEmptyKey) {
      k = atomicCAS(t.keys + slot, kOctEmptyKey, key);
      if (k == kOctEmptyKey) k = key;
    }
    if (k == key) {
      if (*reinterpret_cast<volatile uint32_t*>(t.minseq + slot) > seq) atomicMin(t.minseq + slot, seq);
      return slot;
    }
    slot = (slot + 1) & t.mask;
  }
  return kOctNone;
}

__device__ __forceinline__ uint32_t tableFind(const OctTable& t, unsigned long long key) {
  uint32_t slot = hashKey(key) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    const unsigned long long k = t.keys[slot];
    if (k == key) return slot;
    if (k == kOctEmptyKey) return kOctNone;
    slot = (slot + 1) & t.mask;
  }
  return kOctNone;
}

// OctreePointCloud::adoptBoundingBoxToPoint's test: lower bounds inclusive, upper exclusive
__device__ __forceinline__ bool outsideBox(const OctFrame& f, float x, float y, float z) {
  const double px = x, py = y, pz = z;
  return px < f.min[0] || py < f.min[1] || pz < f.min[2] || px >= f.max[0] || py >= f.max[1] ||
         pz >= f.max[2];
}

__device__ __forceinline__ uint32_t ceilLog2(uint32_t v) {  // v >= 1
  return v <= 1 ? 0u : 32u - __clz(v - 1);
}

// Growth working state of one octGrowKernel launch: the newest frame, the event count and the
// error word live in shared memory between events; every new frame is written through to the
// global event list, which nothing re-reads inside the kernel (a global round trip per field
// and event was what the kernel's time went to).
struct GrowState {
  OctFrame frame;
  uint32_t n;
  uint32_t error;
};

// One point's adoptBoundingBoxToPoint (+ getKeyBitSize for the very first point); appends one
// event per doubling.  Single thread.
// Every lane of the calling warp runs it on its own copy of the state (same inputs, same result);
// `writer` names the one lane that stores the new frames to the global event list.
__device__ void growForPoint(OctFrames* F, GrowState& g, float x, float y, float z, uint32_t seq,
                             double res, bool writer) {
  const double eps = static_cast<double>(FLT_EPSILON);
  const double p[3] = {static_cast<double>(x), static_cast<double>(y), static_cast<double>(z)};
  if (g.n == 0) {
    OctFrame f;
    for (int a = 0; a < 3; ++a) {
      f.min[a] = p[a] - res / 2;
      f.max[a] = p[a] + res / 2;
      f.shift[a] = 0;
    }
    uint32_t max_voxels = 2;
    for (int a = 0; a < 3; ++a) {
      const uint32_t mk = static_cast<uint32_t>(ceil((f.max[a] - f.min[a] - eps) / res));
      max_voxels = mk > max_voxels ? mk : max_voxels;
    }
    // ceil(log2(max_voxels) - eps) of the reference: max_voxels is 2 for a single point
    uint32_t depth = ceilLog2(max_voxels);
    depth = depth > 32 ? 32 : depth;
    const double side = static_cast<double>(1u << depth) * res;
    for (int a = 0; a < 3; ++a) {
      const double over = (side - (f.max[a] - f.min[a])) / 2.0;
      if (over > eps) {
        f.min[a] -= over;
        f.max[a] += over;
      }
    }
    f.depth = depth;
    f.seq = seq;
    if (writer) {
      F->f[0] = f;
      for (int h = 0; h < 32; ++h) {
        F->base[h][0] = F->base[h][1] = F->base[h][2] = 0;
      }
      for (int a = 0; a < 3; ++a) {
        F->anchor_sigma[a] = static_cast<long long>(static_cast<uint32_t>((p[a] - f.min[a]) / res));
      }
    }
    g.frame = f;
    g.n = 1;
    return;
  }
  OctFrame f = g.frame;
  uint32_t n = g.n;
  while (outsideBox(f, x, y, z)) {
    if (n >= kOctMaxFrames || f.depth >= kOctMaxDepth) {
      g.error |= 1u;
      break;
    }
    const bool up[3] = {p[0] >= f.max[0], p[1] >= f.max[1], p[2] >= f.max[2]};
    double side = static_cast<double>(1u << f.depth) * res;
    for (int a = 0; a < 3; ++a) {
      if (!up[a]) {
        f.min[a] -= side;
        f.shift[a] += 1ll << f.depth;
      }
      // height f.depth comes into existence: its partition is fixed from now on
      if (writer) F->base[f.depth][a] = f.shift[a] & ((1ll << f.depth) - 1);
    }
    ++f.depth;
    side = static_cast<double>(1u << f.depth) * res - eps;
    for (int a = 0; a < 3; ++a) f.max[a] = f.min[a] + side;
    f.seq = seq;
    if (writer) F->f[n] = f;
    ++n;
  }
  g.frame = f;
  g.n = n;
}

// The growth events of the members [first, n) in index order: an event is a member outside the
// box as it stands when that member's turn comes (OctreePointCloud::addPointIdx tests every
// point once, against the box of its time).  One CTA walks the members in windows:
//   A  all threads flag the window's members that lie outside the box of the window's start, or
//      within kGrowMargin of its faces, and compact them in index order into shared memory.
//      Nothing else can become an event inside this window: the box never shrinks by more than
//      FLT_EPSILON (the first doubling replaces max = min + side by min + side - eps) plus
//      rounding, which the margin covers.  Most windows flag nothing and cost one barrier.
//   B  warp 0 replays the flagged members in order with the exact test, every lane growing its
//      own register copy of the box (no barrier, no memory round trip per event); lane 0 stores
//      the new frames.
constexpr int kGrowPerThread = 2;
constexpr uint32_t kGrowWindow = 1024 * kGrowPerThread;
constexpr double kGrowMargin = 1e-6;

__device__ __forceinline__ bool nearOrOutsideBox(const OctFrame& f, float x, float y, float z) {
  const double px = x, py = y, pz = z;
  return px < f.min[0] + kGrowMargin || py < f.min[1] + kGrowMargin || pz < f.min[2] + kGrowMargin ||
         px >= f.max[0] - kGrowMargin || py >= f.max[1] - kGrowMargin || pz >= f.max[2] - kGrowMargin;
}

__global__ void __launch_bounds__(1024)
octGrowKernel(OctFrames* F, const float4* __restrict__ xyz, const uint8_t* __restrict__ member,
              uint32_t first, uint32_t n, double res) {
  __shared__ GrowState s_g;
  __shared__ uint32_t s_wcount[2][32];
  __shared__ float4 s_cand[kGrowWindow];
  __shared__ uint32_t s_idx[kGrowWindow];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s_g.n = F->n;
    s_g.error = F->error;
    if (s_g.n > 0) s_g.frame = F->f[s_g.n - 1];
  }
  // a thread owns kGrowPerThread consecutive members of the window
  float4 cur[kGrowPerThread], nxt[kGrowPerThread];
  bool cur_m[kGrowPerThread], nxt_m[kGrowPerThread];
  auto load = [&](float4* p, bool* m, uint32_t base) {
#pragma unroll
    for (int u = 0; u < kGrowPerThread; ++u) {
      const uint32_t i = base + kGrowPerThread * threadIdx.x + u;
      m[u] = i < n && (!member || member[i]);
      p[u] = i < n ? xyz[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load(cur, cur_m, first);
  __syncthreads();
  int buf = 0;
  for (uint32_t base = first; base < n; base += kGrowWindow, buf ^= 1) {
    load(nxt, nxt_m, base + kGrowWindow);
    const bool defined = s_g.n > 0;
    bool flag[kGrowPerThread];
    unsigned ballots[kGrowPerThread];
    unsigned before = 0, mine_total = 0;
#pragma unroll
    for (int u = 0; u < kGrowPerThread; ++u) {
      flag[u] = cur_m[u] && (!defined || nearOrOutsideBox(s_g.frame, cur[u].x, cur[u].y, cur[u].z));
      ballots[u] = __ballot_sync(0xFFFFFFFFu, flag[u]);
      before += __popc(ballots[u] & ((1u << lane) - 1u));
      mine_total += __popc(ballots[u]);
    }
    if (lane == 0) s_wcount[buf][warp] = mine_total;
    __syncthreads();
    const uint32_t wc = s_wcount[buf][lane];
    const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, wc);
    if (total != 0) {  // uniform over the CTA
      uint32_t pos = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? wc : 0u) + before;
#pragma unroll
      for (int u = 0; u < kGrowPerThread; ++u) {
        if (flag[u]) {
          s_cand[pos] = cur[u];
          s_idx[pos] = base + kGrowPerThread * threadIdx.x + u;
          ++pos;
        }
      }
      __syncthreads();
      if (warp == 0) {
        GrowState g = s_g;
        const uint32_t n_before = g.n, error_before = g.error;
        for (uint32_t b = 0; b < total && g.error == error_before; b += 32) {
          const bool have = b + lane < total;
          const float4 c = have ? s_cand[b + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
          const uint32_t ci = have ? s_idx[b + lane] : 0u;
          unsigned done = 0;  // lanes of this batch whose turn has passed
          while (g.error == error_before) {
            const bool out = have && !((done >> lane) & 1u) &&
                             (g.n == 0 || outsideBox(g.frame, c.x, c.y, c.z));
            const unsigned mask = __ballot_sync(0xFFFFFFFFu, out);
            if (mask == 0) break;
            const int src = __ffs(mask) - 1;
            growForPoint(F, g, __shfl_sync(0xFFFFFFFFu, c.x, src), __shfl_sync(0xFFFFFFFFu, c.y, src),
                         __shfl_sync(0xFFFFFFFFu, c.z, src), __shfl_sync(0xFFFFFFFFu, ci, src), res,
                         lane == 0);
            done |= (2u << src) - 1u;  // members up to the event point are never tested again
          }
        }
        if (lane == 0 && (g.n != n_before || g.error != error_before)) s_g = g;
      }
      __syncthreads();
      if (s_g.error != 0 && (s_g.error & 1u)) break;  // box past the supported depth
    }
#pragma unroll
    for (int u = 0; u < kGrowPerThread; ++u) {
      cur[u] = nxt[u];
      cur_m[u] = nxt_m[u];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    F->n = s_g.n;
    F->error = s_g.error;
  }
}

// frame in force when point `seq` is inserted: the last event with f.seq <= seq
__device__ __forceinline__ int insertionFrame(const OctFrames* F, uint32_t seq) {
  int e = -1;
  for (uint32_t i = 0; i < F->n; ++i) {
    if (F->f[i].seq <= seq) e = static_cast<int>(i);
  }
  return e;
}
// frame in force when point `seq` is examined: the last event caused by an earlier point
__device__ __forceinline__ int queryFrame(const OctFrames* F, uint32_t seq) {
  int e = -1;
  for (uint32_t i = 0; i < F->n; ++i) {
    if (F->f[i].seq < seq) e = static_cast<int>(i);
  }
  return e;
}

__global__ void __launch_bounds__(256)
octInsertKernel(OctFrames* F, const OctTable t, const OctPoints pts,
                const uint8_t* __restrict__ member, uint32_t first, uint32_t n, double res) {
  // 8 lanes per point: lane l inserts the nodes at heights l, l + 8, ... (independent keys), so
  // a point costs ceil(depth / 8) dependent atomic round trips instead of `depth`
  const uint32_t i = first + ((blockIdx.x * blockDim.x + threadIdx.x) >> 3);
  const int lane8 = threadIdx.x & 7;
  if (i >= n) return;
  if (member && !member[i]) return;
  const int e = insertionFrame(F, i);
  if (e < 0) return;  // unreachable: growth ran over the same members
  const OctFrame& fi = F->f[e];
  const float4 p = pts.xyz[i];
  // OctreePointCloud::genOctreeKeyforPoint at insertion time, then made stable
  const long long sx = static_cast<long long>(static_cast<uint32_t>((static_cast<double>(p.x) - fi.min[0]) / res)) - fi.shift[0];
  const long long sy = static_cast<long long>(static_cast<uint32_t>((static_cast<double>(p.y) - fi.min[1]) / res)) - fi.shift[1];
  const long long sz = static_cast<long long>(static_cast<uint32_t>((static_cast<double>(p.z) - fi.min[2]) / res)) - fi.shift[2];
  if (lane8 == 0) pts.sigma[i] = make_int4(static_cast<int>(sx), static_cast<int>(sy), static_cast<int>(sz), 0);
  const int depth = static_cast<int>(F->f[F->n - 1].depth);
  for (int h = lane8; h < depth; h += 8) {
    unsigned long long key;
    if (!packKey(h, (sx + F->base[h][0]) >> h, (sy + F->base[h][1]) >> h,
                 (sz + F->base[h][2]) >> h, &key)) {
      atomicOr(&F->error, 2u);
      return;
    }
    const uint32_t slot = tableInsert(t, key, i);
    if (slot == kOctNone) {
      atomicOr(&F->error, 4u);
      return;
    }
    if (h == 0) pts.next[i] = atomicExch(t.head + slot, i);
  }
}

// one lane per height: the keys are independent
__global__ void octRootChainKernel(OctFrames* F, const OctTable t) {
  if (blockIdx.x != 0 || F->n == 0) return;
  const int depth = static_cast<int>(F->f[F->n - 1].depth);
  const uint32_t seq = F->f[0].seq;
  const int h = static_cast<int>(F->f[0].depth) + static_cast<int>(threadIdx.x);
  if (h >= depth) return;
  unsigned long long key;
  if (!packKey(h, (F->anchor_sigma[0] + F->base[h][0]) >> h,
               (F->anchor_sigma[1] + F->base[h][1]) >> h,
               (F->anchor_sigma[2] + F->base[h][2]) >> h, &key)) {
    atomicOr(&F->error, 2u);
    return;
  }
  if (tableInsert(t, key, seq) == kOctNone) atomicOr(&F->error, 4u);
}

// OctreePointCloudSearch::approxNearestSearchRecursive on the implicit tree as of `limit_seq`:
// per level the existing child with the strictly smallest voxel-centre distance (first child
// index wins ties), at the leaf the point with the strictly smallest d^2 in insertion order.
//
// Eight lanes (an aligned octet of a warp) serve one query: lane c probes child c of the current
// node, so a level costs one hash-probe latency instead of eight, and a 3-step shuffle picks the
// winner (smallest distance, then smallest child index == the reference's `d >= best` scan in
// child order).  All eight lanes of an octet take the same path; `octet_mask` names them.
__device__ int approxNearest(const OctFrames* F, const OctFrame& fr, const OctTable& t,
                             const float4* __restrict__ pts, const uint32_t* __restrict__ next,
                             float qx, float qy, float qz, uint32_t limit_seq, double res,
                             float* out_d2, unsigned octet_mask) {
  long long kx = 0, ky = 0, kz = 0;
  const int depth = static_cast<int>(fr.depth);
  const int c = threadIdx.x & 7;
  uint32_t leaf_slot = kOctNone;
  for (int td = 1; td <= depth; ++td) {
    const int h = depth - td;
    const double side = __dmul_rn(res, static_cast<double>(1u << h));
    const long long ox = (fr.shift[0] - F->base[h][0]) >> h;
    const long long oy = (fr.shift[1] - F->base[h][1]) >> h;
    const long long oz = (fr.shift[2] - F->base[h][2]) >> h;
    const long long nx = (kx << 1) + ((c >> 2) & 1);
    const long long ny = (ky << 1) + ((c >> 1) & 1);
    const long long nz = (kz << 1) + (c & 1);
    double d = DBL_MAX;
    int best_c = 8;  // 8 = this child does not exist
    uint32_t best_slot = kOctNone;
    unsigned long long key;
    if (packKey(h, nx - ox, ny - oy, nz - oz, &key)) {
      const uint32_t slot = tableFind(t, key);
      if (slot != kOctNone && t.minseq[slot] < limit_seq) {
        // genVoxelCenterFromOctreeKey: double arithmetic, float result
        const float cx = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(nx), 0.5), side), fr.min[0]));
        const float cy = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(ny), 0.5), side), fr.min[1]));
        const float cz = __double2float_rn(__dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(nz), 0.5), side), fr.min[2]));
        d = static_cast<double>(dist2Exact(cx, cy, cz, qx, qy, qz));
        best_c = c;
        best_slot = slot;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(octet_mask, d, o, 8);
      const int oc = __shfl_xor_sync(octet_mask, best_c, o, 8);
      const uint32_t os = __shfl_xor_sync(octet_mask, best_slot, o, 8);
      // an existing child always beats a missing one (DBL_MAX from a real distance cannot
      // happen: float d^2 <= FLT_MAX); equal distances: the smaller child index
      if (oc < 8 && (best_c == 8 || od < d || (od == d && oc < best_c))) {
        d = od;
        best_c = oc;
        best_slot = os;
      }
    }
    if (best_c == 8) return -1;  // PCL asserts; unreachable below an existing root
    kx = (kx << 1) + ((best_c >> 2) & 1);
    ky = (ky << 1) + ((best_c >> 1) & 1);
    kz = (kz << 1) + (best_c & 1);
    leaf_slot = best_slot;
  }
  if (leaf_slot == kOctNone) return -1;
  // the leaf's few points: every lane of the octet walks the list (same addresses: broadcast)
  float smallest = FLT_MAX;
  uint32_t result = kOctNone;
  for (uint32_t j = t.head[leaf_slot]; j != kOctNone; j = next[j]) {
    if (j >= limit_seq) continue;
    const float4 p = pts[j];
    const float d2 = dist2Exact(p.x, p.y, p.z, qx, qy, qz);
    // `if (d2 >= smallest) continue` in insertion order == argmin over (d2, insertion order)
    if (d2 < smallest || (d2 == smallest && j < result)) {
      smallest = d2;
      result = j;
    }
  }
  if (result == kOctNone) return -1;
  *out_d2 = smallest;
  return static_cast<int>(result);
}

// OctreeCompression::inOctree (octree_compression.cpp:17-40); called by all 8 lanes of an octet
__device__ int inOctree(const OctFrames* F, int frame, const OctTable& t,
                        const float4* __restrict__ pts, const uint32_t* __restrict__ next, float qx,
                        float qy, float qz, uint32_t limit_seq, double res, unsigned octet_mask) {
  if (frame < 0) return -1;  // getLeafCount() == 0
  const OctFrame& fr = F->f[frame];
  const double px = qx, py = qy, pz = qz;
  const bool in_box = (px >= fr.min[0] && px <= fr.max[0]) && (py >= fr.min[1] && py <= fr.max[1]) &&
                      (pz >= fr.min[2] && pz <= fr.max[2]);
  if (!in_box) return -1;
  const double threshold = __dmul_rn(res, res);
  float sqr_distance = static_cast<float>(threshold + 1.0);
  float d2 = 0.f;
  const int idx = approxNearest(F, fr, t, pts, next, qx, qy, qz, limit_seq, res, &d2, octet_mask);
  if (idx >= 0) sqr_distance = d2;
  if (static_cast<double>(sqr_distance) >= threshold) return -1;
  return idx;
}

__device__ __forceinline__ unsigned octetMask() { return 0xFFu << (threadIdx.x & 24); }

__global__ void __launch_bounds__(128)
octQueryPersistentKernel(const OctFrames* F, const OctTable t, const float4* __restrict__ act_xyz,
                         const uint32_t* __restrict__ next, const float* __restrict__ xyz,
                         uint32_t n, double res, int32_t* __restrict__ result) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // 8 lanes per query
  if (i >= n) return;  // whole octets leave together
  const int frame = static_cast<int>(F->n) - 1;
  const int r = inOctree(F, frame, t, act_xyz, next, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2],
                         kOctNone, res, octetMask());
  if ((threadIdx.x & 7) == 0) result[i] = r;
}

__global__ void __launch_bounds__(128)
octDecideTempKernel(const OctFrames* F, const OctTable t, const float4* __restrict__ cand_xyz,
                    const uint32_t* __restrict__ next, uint32_t n_cand, double res,
                    const uint8_t* __restrict__ member_in, uint8_t* __restrict__ member_out,
                    int32_t* __restrict__ decision, uint32_t* __restrict__ changed) {
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // 8 lanes per candidate
  if (c >= n_cand) return;
  const float4 q = cand_xyz[c];
  const int d = inOctree(F, queryFrame(F, c), t, cand_xyz, next, q.x, q.y, q.z, c, res, octetMask());
  if ((threadIdx.x & 7) != 0) return;
  decision[c] = d;
  const uint8_t m = d < 0 ? 1 : 0;
  member_out[c] = m;
  if (m != member_in[c]) atomicOr(changed, 1u);
}

// ---- BlockCompression cell hash
constexpr long long kCellOffset = 1ll << 20;
constexpr long long kCellLimit = 1ll << 21;

__device__ __forceinline__ bool cellKey(float x, float y, float z, float inv,
                                        unsigned long long* key) {
  const long long cx = static_cast<long long>(floorf(__fmul_rn(x, inv))) + kCellOffset;
  const long long cy = static_cast<long long>(floorf(__fmul_rn(y, inv))) + kCellOffset;
  const long long cz = static_cast<long long>(floorf(__fmul_rn(z, inv))) + kCellOffset;
  if (cx < 0 || cy < 0 || cz < 0 || cx >= kCellLimit || cy >= kCellLimit || cz >= kCellLimit) {
    return false;
  }
  *key = (static_cast<unsigned long long>(cx) << 42) | (static_cast<unsigned long long>(cy) << 21) |
         static_cast<unsigned long long>(cz);
  return true;
}

__global__ void __launch_bounds__(256)
cellInsertKernel(OctFrames* F, const OctTable t, const float4* __restrict__ xyz, uint32_t first,
                 uint32_t n, float inv) {
  const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = xyz[i];
  unsigned long long key;
  if (!cellKey(p.x, p.y, p.z, inv, &key)) {
    atomicOr(&F->error, 2u);
    return;
  }
  if (tableInsert(t, key, i) == kOctNone) atomicOr(&F->error, 4u);
}

__global__ void __launch_bounds__(256)
cellQueryPersistentKernel(const OctTable t, const float* __restrict__ xyz, uint32_t n, float inv,
                          int32_t* __restrict__ result) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long key;
  int32_t r = -1;
  if (cellKey(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], inv, &key)) {
    const uint32_t slot = tableFind(t, key);
    if (slot != kOctNone) r = static_cast<int32_t>(t.minseq[slot]);
  }
  result[i] = r;
}

__global__ void __launch_bounds__(256)
cellDecideTempKernel(const OctTable t, const float4* __restrict__ cand_xyz, uint32_t n_cand,
                     float inv, uint8_t* __restrict__ member_out, int32_t* __restrict__ decision) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const float4 p = cand_xyz[c];
  unsigned long long key;
  int32_t d = -1;
  if (cellKey(p.x, p.y, p.z, inv, &key)) {
    const uint32_t slot = tableFind(t, key);
    if (slot != kOctNone && t.minseq[slot] != c) d = static_cast<int32_t>(t.minseq[slot]);
  }
  decision[c] = d;
  member_out[c] = d < 0 ? 1 : 0;
}

}  // namespace

cudaError_t launchCellInsert(OctFrames* frames, OctTable table, const float4* xyz, uint32_t first,
                             uint32_t n, float voxel_size_inv, cudaStream_t stream) {
  if (first >= n) return cudaSuccess;
  cellInsertKernel<<<(n - first + 255) / 256, 256, 0, stream>>>(frames, table, xyz, first, n,
                                                                voxel_size_inv);
  return cudaGetLastError();
}

cudaError_t launchCellQueryPersistent(OctTable table, const float* xyz, uint32_t n,
                                      float voxel_size_inv, int32_t* result, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  cellQueryPersistentKernel<<<(n + 255) / 256, 256, 0, stream>>>(table, xyz, n, voxel_size_inv, result);
  return cudaGetLastError();
}

cudaError_t launchCellDecideTemp(OctTable table, const float4* cand_xyz, uint32_t n_cand,
                                 float voxel_size_inv, uint8_t* member_out, int32_t* decision,
                                 cudaStream_t stream) {
  if (n_cand == 0) return cudaSuccess;
  cellDecideTempKernel<<<(n_cand + 255) / 256, 256, 0, stream>>>(table, cand_xyz, n_cand,
                                                                 voxel_size_inv, member_out, decision);
  return cudaGetLastError();
}

cudaError_t launchOctGrow(OctFrames* frames, const float4* xyz, const uint8_t* member,
                          uint32_t first, uint32_t n, double resolution, cudaStream_t stream) {
  if (first >= n) return cudaSuccess;
  octGrowKernel<<<1, 1024, 0, stream>>>(frames, xyz, member, first, n, resolution);
  return cudaGetLastError();
}

cudaError_t launchOctInsert(OctFrames* frames, OctTable table, OctPoints pts,
                            const uint8_t* member, uint32_t first, uint32_t n, double resolution,
                            cudaStream_t stream) {
  if (first >= n) return cudaSuccess;
  octInsertKernel<<<(n - first + 31) / 32, 256, 0, stream>>>(frames, table, pts, member, first, n,
                                                               resolution);
  return cudaGetLastError();
}

cudaError_t launchOctInsertRootChain(OctFrames* frames, OctTable table, cudaStream_t stream) {
  octRootChainKernel<<<1, 32, 0, stream>>>(frames, table);
  return cudaGetLastError();
}

cudaError_t launchOctQueryPersistent(const OctFrames* frames, OctTable table, const float4* act_xyz,
                                     const uint32_t* next, const float* xyz, uint32_t n,
                                     double resolution, int32_t* result, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  octQueryPersistentKernel<<<(n + 15) / 16, 128, 0, stream>>>(frames, table, act_xyz, next, xyz, n,
                                                               resolution, result);
  return cudaGetLastError();
}

cudaError_t launchOctDecideTemp(const OctFrames* frames, OctTable table, const float4* cand_xyz,
                                const uint32_t* next, uint32_t n_cand, double resolution,
                                const uint8_t* member_in, uint8_t* member_out, int32_t* decision,
                                uint32_t* changed, cudaStream_t stream) {
  if (n_cand == 0) return cudaSuccess;
  octDecideTempKernel<<<(n_cand + 15) / 16, 128, 0, stream>>>(
      frames, table, cand_xyz, next, n_cand, resolution, member_in, member_out, decision, changed);
  return cudaGetLastError();
}

}  // namespace pgmo

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadOctTreeKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, octGrowKernel);
  cudaFuncGetAttributes(&attr, octInsertKernel);
  cudaFuncGetAttributes(&attr, octRootChainKernel);
  cudaFuncGetAttributes(&attr, octQueryPersistentKernel);
  cudaFuncGetAttributes(&attr, octDecideTempKernel);
  cudaFuncGetAttributes(&attr, cellInsertKernel);
  cudaFuncGetAttributes(&attr, cellQueryPersistentKernel);
  cudaFuncGetAttributes(&attr, cellDecideTempKernel);
  cudaGetLastError();
}
}  // namespace pgmo
