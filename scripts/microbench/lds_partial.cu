// What a gather of 16-byte chunks from shared memory (LDS.128, rows of 128 bytes, chunk c of a
// row stored at c ^ class(row) like the blend kernel's pose rows) costs when only a fraction of
// the lanes takes part (the others predicated off) — the question behind "reload a row only
// for the lanes whose neighbour changed".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lds_partial scripts/microbench/lds_partial.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 2048;
constexpr int kRows = 256;

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// pool: rows a warp draws from (small pool = lanes share rows); active_pct: lanes taking part
__global__ void gatherKernel(int pool, int active_pct, unsigned long long* cycles, float* sink) {
  extern __shared__ float4 rows[];  // kRows x 8 chunks
  for (int i = threadIdx.x; i < kRows * 8; i += blockDim.x) rows[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < kIters; ++i) {
    const unsigned h = hash32(i * 1315423911u + warp * 2654435761u);
    const unsigned hl = hash32(h ^ (lane * 0x9E3779B9u));
    const int row = (h + hl % pool) % kRows;          // pool consecutive rows starting at a random one
    const int cls = hash32(row * 77u + 5u) & 7;          // "colouring": pseudo-random class per row
    const bool on = static_cast<int>(hash32(hl + 11u) % 100u) < active_pct;
    if (on) {
      const float4* r = rows + row * 8;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = r[c ^ cls];
        acc += v.x + v.w;
      }
    }
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// the same loop without the loads: what the address arithmetic costs
__global__ void emptyKernel(int pool, int active_pct, unsigned long long* cycles, float* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < kIters; ++i) {
    const unsigned h = hash32(i * 1315423911u + warp * 2654435761u);
    const unsigned hl = hash32(h ^ (lane * 0x9E3779B9u));
    const int row = (h + hl % pool) % kRows;
    const int cls = hash32(row * 77u + 5u) & 7;
    const bool on = static_cast<int>(hash32(hl + 11u) % 100u) < active_pct;
    if (on) acc += row + cls;
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  unsigned long long* d_cycles;
  float* d_sink;
  cudaMalloc(&d_cycles, 1024 * sizeof(unsigned long long));
  cudaMalloc(&d_sink, 1024 * 1024 * sizeof(float));
  unsigned long long h[256], h0[256];
  const int warps = 32;
  for (int pool : {1, 4, 8, 32}) {
    for (int pct : {100, 75, 50, 30, 15, 5}) {
      gatherKernel<<<148, warps * 32, kRows * 128>>>(pool, pct, d_cycles, d_sink);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d_cycles, 148 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      emptyKernel<<<148, warps * 32>>>(pool, pct, d_cycles, d_sink);
      cudaDeviceSynchronize();
      cudaMemcpy(h0, d_cycles, 148 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      printf("rows drawn from %2d, %3d %% of the lanes active: %.2f cycles per warp-level LDS.128 (SM-wide; loop overhead %.2f per row of 8 loads)\n",
             pool, pct, double(h[0]) / (double(kIters) * warps * 8), double(h0[0]) / (double(kIters) * warps));
    }
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
