// Per-run spatial grids for K3 (sm_100a).
//
// The window schedule gives every run an *alive set* of control points (a few hundred in the
// reference's configurations).  Scanning all of them per vertex makes K3 issue-bound
// (profiles/r01_*).  Because the schedule is cached across solves — it depends on control
// positions / stamps / validity and on the mesh stamps, not on the poses — we can afford to
// bucket each run's alive set once into a small uniform grid (<= kGridMaxCells cells, cell size
// chosen from the set's density) and let every vertex scan only the cells around it, growing
// the searched box until the (K+1)-th best distance is provably inside it.  The candidate set is
// always a superset of the exact top-(K+1), distances use the same exact arithmetic, so the
// result is bit-identical to the brute-force kernel.
#include <cfloat>

#include "deform_device.cuh"

namespace pgmo {

namespace {

using namespace dev;

constexpr int kBuildThreads = 256;

__device__ __forceinline__ int cellCoord(float p, float origin, float inv_cell, int n) {
  const float u = (p - origin) * inv_cell;
  int c = static_cast<int>(floorf(u));
  c = c < 0 ? 0 : c;
  return c >= n ? n - 1 : c;
}

// One CTA per run: bounding box + count of the alive set, grid parameters, counting sort.
__global__ void __launch_bounds__(kBuildThreads)
buildRunGridsKernel(const uint4* __restrict__ run_info, const float4* __restrict__ ctrl_f4,
                    const uint32_t* __restrict__ point_offset, RunGridDev* __restrict__ grids,
                    float4* __restrict__ grid_points, uint32_t* __restrict__ grid_cells) {
  __shared__ float s_min[3][kBuildThreads / 32], s_max[3][kBuildThreads / 32];
  __shared__ uint32_t s_cnt[kBuildThreads / 32];
  __shared__ uint32_t s_hist[kGridMaxCells + 1];
  __shared__ RunGridDev s_grid;
  const uint32_t r = blockIdx.x;
  const uint4 info = run_info[r];
  const int run = static_cast<int>(r);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  uint32_t cnt = 0;
  if (!info.w) {
    for (uint32_t j = info.x + threadIdx.x; j < info.y; j += kBuildThreads) {
      const float4 c = ctrl_f4[j];
      if (__float_as_int(c.w) <= run) continue;
      ++cnt;
      mn[0] = fminf(mn[0], c.x);
      mn[1] = fminf(mn[1], c.y);
      mn[2] = fminf(mn[2], c.z);
      mx[0] = fmaxf(mx[0], c.x);
      mx[1] = fmaxf(mx[1], c.y);
      mx[2] = fmaxf(mx[2], c.z);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xFFFFFFFFu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xFFFFFFFFu, mx[a], o));
    }
  }
  if (lane == 0) {
    s_cnt[warp] = cnt;
    for (int a = 0; a < 3; ++a) {
      s_min[a][warp] = mn[a];
      s_max[a][warp] = mx[a];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t n = 0;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int w = 0; w < kBuildThreads / 32; ++w) {
      n += s_cnt[w];
      for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(lo[a], s_min[a][w]);
        hi[a] = fmaxf(hi[a], s_max[a][w]);
      }
    }
    RunGridDev g;
    g.n_alive = n;
    g.point_offset = point_offset[r];
    g.nx = g.ny = g.nz = 1;
    g.ox = g.oy = g.oz = 0.f;
    g.cell = 1.f;
    g.inv_cell = 1.f;
    g.r0 = 1.f;
    if (n > 0) {
      g.ox = lo[0];
      g.oy = lo[1];
      g.oz = lo[2];
      float ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
      const float longest = fmaxf(fmaxf(ext[0], ext[1]), ext[2]);
      if (n > 16 && longest > 0.f) {
        // estimated distance r of the (K+1)-th neighbour at the set's density (extents thinner
        // than 2r do not count as a dimension); first search radius r0 = 1.25 r, cell = r0 / 2
        const float kk = static_cast<float>(info.z) + 3.f;
        float rr = 0.f;
        for (int it = 0; it < 4; ++it) {
          float vol = 1.f;
          int dims_in_play = 0;
          for (int a = 0; a < 3; ++a) {
            if (ext[a] > 2.f * rr) {
              vol *= ext[a];
              ++dims_in_play;
            }
          }
          if (dims_in_play == 0) {
            rr = longest;
            break;
          }
          const float per = vol * kk / static_cast<float>(n);
          const float next = dims_in_play == 3 ? cbrtf(per * 0.2387f)
                                               : (dims_in_play == 2 ? sqrtf(per * 0.3183f) : 0.5f * per);
          if (fabsf(next - rr) <= 1e-3f * next) {
            rr = next;
            break;
          }
          rr = next;
        }
        g.r0 = 1.25f * fminf(rr, longest);
        float cell = 0.5f * g.r0;
        cell = fminf(cell, longest);
        cell = fmaxf(cell, longest / 256.f);
        int nx, ny, nz;
        while (true) {
          nx = static_cast<int>(ext[0] / cell) + 1;
          ny = static_cast<int>(ext[1] / cell) + 1;
          nz = static_cast<int>(ext[2] / cell) + 1;
          if (static_cast<long long>(nx) * ny * nz <= kGridMaxCells) break;
          cell *= 1.26f;
        }
        g.nx = nx;
        g.ny = ny;
        g.nz = nz;
        g.cell = cell;
        g.inv_cell = 1.f / cell;
      }
    }
    s_grid = g;
    grids[r] = g;
  }
  for (int i = threadIdx.x; i <= kGridMaxCells; i += kBuildThreads) s_hist[i] = 0;
  __syncthreads();
  const RunGridDev g = s_grid;
  if (g.n_alive == 0) {
    for (int i = threadIdx.x; i <= kGridMaxCells; i += kBuildThreads) {
      grid_cells[static_cast<uint64_t>(r) * (kGridMaxCells + 1) + i] = 0;
    }
    return;
  }
  const int n_cells = g.nx * g.ny * g.nz;
  for (uint32_t j = info.x + threadIdx.x; j < info.y; j += kBuildThreads) {
    const float4 c = ctrl_f4[j];
    if (__float_as_int(c.w) <= run) continue;
    const int cell = (cellCoord(c.z, g.oz, g.inv_cell, g.nz) * g.ny +
                      cellCoord(c.y, g.oy, g.inv_cell, g.ny)) * g.nx +
                     cellCoord(c.x, g.ox, g.inv_cell, g.nx);
    atomicAdd(&s_hist[cell + 1], 1u);
  }
  __syncthreads();
  // inclusive scan of s_hist[1..n_cells] by one warp (n_cells <= 1024)
  if (warp == 0) {
    uint32_t carry = 0;
    for (int base = 1; base <= n_cells; base += 32) {
      const int i = base + lane;
      uint32_t v = i <= n_cells ? s_hist[i] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += up;
      }
      v += carry;
      if (i <= n_cells) s_hist[i] = v;
      carry = __shfl_sync(0xFFFFFFFFu, v, 31);
    }
  }
  __syncthreads();
  uint32_t* cells_out = grid_cells + static_cast<uint64_t>(r) * (kGridMaxCells + 1);
  for (int i = threadIdx.x; i <= n_cells; i += kBuildThreads) cells_out[i] = s_hist[i];
  __syncthreads();
  // scatter: s_hist[cell] doubles as the write cursor of `cell` (cells_out keeps the starts)
  for (uint32_t j = info.x + threadIdx.x; j < info.y; j += kBuildThreads) {
    const float4 c = ctrl_f4[j];
    if (__float_as_int(c.w) <= run) continue;
    const int cell = (cellCoord(c.z, g.oz, g.inv_cell, g.nz) * g.ny +
                      cellCoord(c.y, g.oy, g.inv_cell, g.ny)) * g.nx +
                     cellCoord(c.x, g.ox, g.inv_cell, g.nx);
    const uint32_t at = atomicAdd(&s_hist[cell], 1u);
    grid_points[g.point_offset + at] = make_float4(c.x, c.y, c.z, __int_as_float(static_cast<int>(j)));
  }
}

// sorted insertion of (d2, index) keys; ascending, ties by ascending control index
template <int KCAP>
__device__ __forceinline__ void insertKey(unsigned long long (&best)[KCAP], unsigned long long key) {
  if (key >= best[KCAP - 1]) return;
  best[KCAP - 1] = key;
#pragma unroll
  for (int i = KCAP - 1; i > 0; --i) {
    const unsigned long long a = best[i - 1], b = best[i];
    const bool sw = b < a;
    best[i - 1] = sw ? b : a;
    best[i] = sw ? a : b;
  }
}

constexpr int kStagePoints = 1536;  // alive points of one run staged in shared memory (24 KB)
constexpr int kChunk = kDeformChunk;  // positions per work item (one CTA)
constexpr int kThreads = 256;
constexpr int kMaxTiles = 1024;

struct BinnedSmem {
  float4 pts[kStagePoints];
  uint32_t cells[kGridMaxCells + 1];
  uint32_t hist[kMaxTiles + 1];
  float vx[kChunk], vy[kChunk], vz[kChunk];
  uint16_t perm[kChunk];
  uint32_t warp_tot[kThreads / 32];
};

// K3 over the per-run grids, with spatially coherent warps.
//
// A work item is up to kChunk consecutive positions of ONE run.  The CTA stages that run's grid
// in shared memory, loads its vertices (coalesced), counting-sorts them in shared memory by
// 2x2-cell column (tile) of the grid, and hands groups of 32 tile-adjacent vertices to its
// warps.  The lanes of a warp then share one candidate box (the cells overlapping the group's
// query bounding box grown by R) and walk it in lock step: every candidate point is one
// broadcast shared-memory read, loop trip counts are uniform, and far candidates are rejected
// by all lanes at once.  R starts at the run's estimated (K+1)-th neighbour distance and grows
// by 1.5x until every lane can prove that nothing outside the box can enter its top-(K+1).
template <int KCAP>
__global__ void __launch_bounds__(kThreads, 3)
deformKnnBlendBinnedKernel(const DeformArgs a, const GridArgs ga, const WorkItem* __restrict__ items,
                           const int precision, const uint32_t skip_alive_le) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BinnedSmem& S = *reinterpret_cast<BinnedSmem*>(smem_raw);

  const WorkItem item = items[blockIdx.x];
  const uint32_t r = item.run;
  const uint4 info = __ldg(a.run_info + r);
  const uint64_t p_base = item.start;  // local position of the item's first vertex
  const int t = threadIdx.x;
  const int lane = t & 31;
  const int count = static_cast<int>(item.count);
  if (info.w) {
    for (int i = t; i < count; i += kThreads) a.status[p_base + i] = kStatusSkipped;
    return;
  }
  const RunGridDev g = ga.grids[r];
  if (g.n_alive <= skip_alive_le) return;  // the bucket-ordered kernel's item
  const int K = static_cast<int>(info.z) + 2;
  if (K + 1 > KCAP) {
    for (int i = t; i < count; i += kThreads) sendToResolver(a, p_base + i);
    return;
  }
  const int n_cells = g.nx * g.ny * g.nz;
  const int tiles_x = (g.nx + 1) >> 1, tiles_y = (g.ny + 1) >> 1;
  const int n_tiles = tiles_x * tiles_y;  // <= kGridMaxCells
  const bool staged = g.n_alive <= kStagePoints;
  const uint32_t* gcells = ga.cells + static_cast<uint64_t>(r) * (kGridMaxCells + 1);
  if (staged) {
    const float4* src = ga.points + g.point_offset;
    for (uint32_t i = t; i < g.n_alive; i += kThreads) S.pts[i] = __ldg(src + i);
  }
  for (int i = t; i <= n_cells; i += kThreads) S.cells[i] = __ldg(gcells + i);
  for (int i = t; i <= n_tiles; i += kThreads) S.hist[i] = 0;
  __syncthreads();

  // ---- load the item's vertices and bin them by tile
  for (int i = t; i < count; i += kThreads) {
    const uint64_t ii = a.subset ? a.subset[p_base + i] : p_base + i;
    const float x = a.x[ii], y = a.y[ii], z = a.z[ii];
    S.vx[i] = x;
    S.vy[i] = y;
    S.vz[i] = z;
    const int tile = (cellCoord(y, g.oy, g.inv_cell, g.ny) >> 1) * tiles_x +
                     (cellCoord(x, g.ox, g.inv_cell, g.nx) >> 1);
    atomicAdd(&S.hist[tile + 1], 1u);
  }
  __syncthreads();
  {
    // inclusive scan of hist[1..n_tiles]: 4 entries per thread, then across threads
    constexpr int per = (kMaxTiles + kThreads - 1) / kThreads;
    uint32_t local[per];
    uint32_t sum = 0;
#pragma unroll
    for (int e = 0; e < per; ++e) {
      const int i = 1 + t * per + e;
      local[e] = i <= n_tiles ? S.hist[i] : 0;
      sum += local[e];
    }
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) S.warp_tot[t >> 5] = incl;
    __syncthreads();
    uint32_t run_sum = incl - sum;
    for (int w = 0; w < (t >> 5); ++w) run_sum += S.warp_tot[w];
#pragma unroll
    for (int e = 0; e < per; ++e) {
      const int i = 1 + t * per + e;
      run_sum += local[e];
      if (i <= n_tiles) S.hist[i] = run_sum;  // hist[i] = #vertices in tiles < i
    }
  }
  __syncthreads();
  for (int i = t; i < count; i += kThreads) {
    const int tile = (cellCoord(S.vy[i], g.oy, g.inv_cell, g.ny) >> 1) * tiles_x +
                     (cellCoord(S.vx[i], g.ox, g.inv_cell, g.nx) >> 1);
    const uint32_t at = atomicAdd(&S.hist[tile], 1u);
    S.perm[at] = static_cast<uint16_t>(i);
  }
  __syncthreads();

  // ---- groups of 32 tile-adjacent vertices, one shared candidate box per group
  const unsigned full = 0xFFFFFFFFu;
  const float4* pts = staged ? S.pts : ga.points + g.point_offset;
  const int n_groups = (count + 31) >> 5;
  const int n_need = static_cast<int>(g.n_alive) < K + 1 ? static_cast<int>(g.n_alive) : K + 1;
  for (int group = t >> 5; group < n_groups; group += kThreads / 32) {
    const int slot = group * 32 + lane;
    const bool active = slot < count;
    const int v = active ? S.perm[slot] : 0;
    const float qx = active ? S.vx[v] : 0.f, qy = active ? S.vy[v] : 0.f,
                qz = active ? S.vz[v] : 0.f;
    // continuous cell coordinates (may lie outside the grid)
    const float ux = (qx - g.ox) * g.inv_cell, uy = (qy - g.oy) * g.inv_cell,
                uz = (qz - g.oz) * g.inv_cell;
    float lox = active ? ux : FLT_MAX, hix = active ? ux : -FLT_MAX;
    float loy = active ? uy : FLT_MAX, hiy = active ? uy : -FLT_MAX;
    float loz = active ? uz : FLT_MAX, hiz = active ? uz : -FLT_MAX;
    for (int o = 16; o > 0; o >>= 1) {
      lox = fminf(lox, __shfl_xor_sync(full, lox, o));
      hix = fmaxf(hix, __shfl_xor_sync(full, hix, o));
      loy = fminf(loy, __shfl_xor_sync(full, loy, o));
      hiy = fmaxf(hiy, __shfl_xor_sync(full, hiy, o));
      loz = fminf(loz, __shfl_xor_sync(full, loz, o));
      hiz = fmaxf(hiz, __shfl_xor_sync(full, hiz, o));
    }

    unsigned long long best[KCAP];
#pragma unroll
    for (int i = 0; i < KCAP; ++i) best[i] = ~0ull;

    int px0 = 0, px1 = -1, py0 = 0, py1 = -1, pz0 = 0, pz1 = -1;  // box scanned so far (empty)
    float R = g.r0 * g.inv_cell * ga.r_scale;  // in cell units
    while (true) {
      // cells overlapping [lo - R, hi + R], clamped into the grid
      const float fx = static_cast<float>(g.nx - 1), fy = static_cast<float>(g.ny - 1),
                  fz = static_cast<float>(g.nz - 1);
      int bx0 = static_cast<int>(floorf(fminf(fmaxf(lox - R, 0.f), fx)));
      int by0 = static_cast<int>(floorf(fminf(fmaxf(loy - R, 0.f), fy)));
      int bz0 = static_cast<int>(floorf(fminf(fmaxf(loz - R, 0.f), fz)));
      int bx1 = static_cast<int>(floorf(fminf(fmaxf(hix + R, 0.f), fx)));
      int by1 = static_cast<int>(floorf(fminf(fmaxf(hiy + R, 0.f), fy)));
      int bz1 = static_cast<int>(floorf(fminf(fmaxf(hiz + R, 0.f), fz)));
      if (px1 >= px0) {  // keep the boxes nested (they are: R only grows)
        bx0 = min(bx0, px0);
        by0 = min(by0, py0);
        bz0 = min(bz0, pz0);
        bx1 = max(bx1, px1);
        by1 = max(by1, py1);
        bz1 = max(bz1, pz1);
      }
      for (int z = bz0; z <= bz1; ++z) {
        for (int y = by0; y <= by1; ++y) {
          const uint32_t row = static_cast<uint32_t>((z * g.ny + y) * g.nx);
          const bool inside_old = z >= pz0 && z <= pz1 && y >= py0 && y <= py1;
          // old rows contribute the two x-slabs outside the old box, new rows the whole span
          for (int part = 0; part < 2; ++part) {
            uint32_t cur, end;
            if (!inside_old) {
              if (part == 1) break;
              cur = S.cells[row + bx0];
              end = S.cells[row + bx1 + 1];
            } else if (part == 0) {
              if (bx0 >= px0) continue;
              cur = S.cells[row + bx0];
              end = S.cells[row + px0];
            } else {
              if (bx1 <= px1) continue;
              cur = S.cells[row + px1 + 1];
              end = S.cells[row + bx1 + 1];
            }
            for (; cur < end; ++cur) {
              const float4 c = pts[cur];  // same address for the whole warp: broadcast
              const float d2 = dist2Exact(c.x, c.y, c.z, qx, qy, qz);
              const unsigned long long key =
                  (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) |
                  static_cast<unsigned long long>(__float_as_uint(c.w));
              insertKey<KCAP>(best, key);
            }
          }
        }
      }
      px0 = bx0; px1 = bx1; py0 = by0; py1 = by1; pz0 = bz0; pz1 = bz1;
      const bool whole = bx0 == 0 && by0 == 0 && bz0 == 0 && bx1 == g.nx - 1 &&
                         by1 == g.ny - 1 && bz1 == g.nz - 1;
      if (whole) break;
      bool ok = !active;
      // best[n_need - 1] without dynamic indexing (which would push best[] to local memory):
      // best[] is ascending, so it is the maximum of the first n_need entries
      unsigned long long kth_key = 0ull;
#pragma unroll
      for (int i = 0; i < KCAP; ++i) {
        const unsigned long long cand = i < n_need ? best[i] : 0ull;
        kth_key = cand > kth_key ? cand : kth_key;
      }
      if (active && n_need > 0 && kth_key != ~0ull) {
        // nearest face of the scanned box with unscanned cells behind it, in cell units
        float dout = FLT_MAX;
        if (bx0 > 0) dout = fminf(dout, ux - static_cast<float>(bx0));
        if (bx1 < g.nx - 1) dout = fminf(dout, static_cast<float>(bx1 + 1) - ux);
        if (by0 > 0) dout = fminf(dout, uy - static_cast<float>(by0));
        if (by1 < g.ny - 1) dout = fminf(dout, static_cast<float>(by1 + 1) - uy);
        if (bz0 > 0) dout = fminf(dout, uz - static_cast<float>(bz0));
        if (bz1 < g.nz - 1) dout = fminf(dout, static_cast<float>(bz1 + 1) - uz);
        // 1 % of a cell of slack covers the float rounding of the binning (<= 4e-4 cell)
        const float safe = (dout - 0.01f) * g.cell;
        if (safe > 0.f) {
          const float kth = __uint_as_float(static_cast<uint32_t>(kth_key >> 32));
          ok = kth <= safe * safe;
        }
      }
      if (__all_sync(full, ok)) break;
      R = fmaxf(R * ga.r_grow_mul, R + ga.r_grow_add);
    }
    if (!active) continue;

    const uint64_t p = p_base + v;
    const uint64_t ii = a.subset ? a.subset[p] : p;
    const int n_alive = static_cast<int>(g.n_alive);
    const int n = n_alive < K ? n_alive : K;
    const int n_check = n_alive < K + 1 ? n_alive : K + 1;
    float nbr_d2[KCAP];
    int nbr_idx[KCAP];
    bool tie = false;
#pragma unroll
    for (int i = 0; i < KCAP; ++i) {
      nbr_d2[i] = __uint_as_float(static_cast<uint32_t>(best[i] >> 32));
      nbr_idx[i] = static_cast<int>(static_cast<uint32_t>(best[i]));
    }
#pragma unroll
    for (int i = 0; i + 1 < KCAP; ++i) {
      if (i + 1 < n_check && nbr_d2[i] == nbr_d2[i + 1]) tie = true;
    }
    if (tie) {
      sendToResolver(a, p);
      continue;
    }
    writeResultReg<KCAP>(a, p, ii, precision, qx, qy, qz, nbr_d2, nbr_idx, n);
    a.status[p] = kStatusDone;
  }
}

}  // namespace

cudaError_t launchBuildRunGrids(uint32_t n_runs, const uint4* run_info, const float4* ctrl_f4,
                                const uint32_t* point_offset, RunGridDev* grids,
                                float4* grid_points, uint32_t* grid_cells, cudaStream_t stream) {
  if (n_runs == 0) return cudaSuccess;
  buildRunGridsKernel<<<n_runs, kBuildThreads, 0, stream>>>(run_info, ctrl_f4, point_offset, grids,
                                                           grid_points, grid_cells);
  return cudaGetLastError();
}

cudaError_t launchDeformKnnBlendGrid(const DeformArgs& args, const GridArgs& gargs,
                                     const WorkItem* items, uint32_t n_items, int precision,
                                     uint32_t max_k, uint32_t skip_alive_le, cudaStream_t stream) {
  if (n_items == 0) return cudaSuccess;
  const size_t smem = sizeof(BinnedSmem);
  static thread_local int configured_device = -1;
  int device = 0;
  cudaGetDevice(&device);
  if (configured_device != device) {
    cudaError_t e = cudaFuncSetAttribute(deformKnnBlendBinnedKernel<8>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(deformKnnBlendBinnedKernel<16>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    configured_device = device;
  }
  if (max_k + 3 <= 8) {
    deformKnnBlendBinnedKernel<8><<<n_items, kThreads, smem, stream>>>(args, gargs, items, precision,
                                                                       skip_alive_le);
  } else {
    deformKnnBlendBinnedKernel<16><<<n_items, kThreads, smem, stream>>>(args, gargs, items, precision,
                                                                        skip_alive_le);
  }
  return cudaGetLastError();
}

}  // namespace pgmo

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadDeformGridKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, buildRunGridsKernel);
  cudaGetLastError();
}
}  // namespace pgmo
