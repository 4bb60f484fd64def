// Microbenchmarks behind DESIGN 2.9: how many cycles a shared-memory LDS.128 costs for a given
// number of distinct 16-byte segments per warp, and the FP64 add / mul issue rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_fp64 smem_fp64.cu && ./smem_fp64
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

// pattern p: lane -> 16-byte segment index
__device__ __forceinline__ int segOf(int p, int lane) {
  switch (p) {
    case 0: return 0;                       // one address for the warp
    case 1: return lane;                    // 32 distinct, contiguous
    case 2: return (lane & 1);              // 2 distinct, neighbouring bank groups
    case 3: return (lane & 3);              // 4 distinct
    case 4: return (lane & 7);              // 8 distinct, one per bank group
    case 5: return (lane >> 2);             // 8 distinct, 2 per quarter warp
    case 6: return (lane & 3) * 8;          // 4 distinct in ONE bank group (conflict)
    case 7: return (lane & 15);             // 16 distinct, 2 per bank group
    case 8: return (lane >> 3) * 9;         // 4 distinct, one per quarter warp, different groups
    case 9: return ((lane * 5) & 7) + 8 * ((lane >> 3) & 1);  // 16 distinct, permuted
    case 10: return (lane & 7) % 3;                           // 3 distinct per quarter warp
    case 11: return (lane & 1) * 4;                           // 2 distinct, bank groups 0 and 4
    case 12: return (lane & 1) * 8;                           // 2 distinct, same bank group
    case 13: return (lane >> 3) * 2 + (lane & 1);             // 2 distinct per quarter, 8 in the warp
    case 14: return ((lane & 1) ? 9 : 0) + 17 * (lane >> 3);  // 2 per quarter, scattered rows
    case 15: return (lane & 7) < 6 ? 0 : (lane & 7) - 5;      // 3 distinct per quarter, 6 + 1 + 1
    case 16: return (lane & 7) == 7 ? 1 : 0;                  // 2 distinct per quarter, 7 + 1
    case 17: return (lane & 7) < 4 ? 0 : ((lane & 7) < 6 ? 1 : ((lane & 7) < 7 ? 2 : 3));  // 4 per quarter 4+2+1+1
    case 18: return ((lane & 7) % 3) * 11 + 29 * (lane >> 3);  // 3 per quarter, scattered rows
    case 19: return (lane & 3) + 4 * ((lane >> 3) & 1);        // 4 per quarter, 8 in the warp
    default: return lane;
  }
}

template <int VEC>
__global__ void ldsKernel(int pattern, unsigned long long* cycles, float* sink) {
  __shared__ float4 buf[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int seg = segOf(pattern, lane);
  float acc = 0.f;
  const unsigned long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < kIters; ++i) {
    if (VEC == 4) {
      const float4 v = buf[(seg + (i & 15) * 64) & 2047];
      acc += (v.x + v.y) + (v.z + v.w);
    } else if (VEC == 2) {
      const float2 v = reinterpret_cast<const float2*>(buf)[(seg + (i & 15) * 64) & 4095];
      acc += v.x + v.y;
    } else {
      const float v = reinterpret_cast<const float*>(buf)[(seg + (i & 15) * 64) & 8191];
      acc += v;
    }
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
__global__ void fp64Kernel(unsigned long long* cycles, double* sink, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001;
  const unsigned long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < kIters; ++i) {
    if (OP == 0) {
      a0 = __dadd_rn(a0, m); a1 = __dadd_rn(a1, m); a2 = __dadd_rn(a2, m); a3 = __dadd_rn(a3, m);
      a4 = __dadd_rn(a4, m); a5 = __dadd_rn(a5, m); a6 = __dadd_rn(a6, m); a7 = __dadd_rn(a7, m);
    } else if (OP == 1) {
      a0 = __dmul_rn(a0, m); a1 = __dmul_rn(a1, m); a2 = __dmul_rn(a2, m); a3 = __dmul_rn(a3, m);
      a4 = __dmul_rn(a4, m); a5 = __dmul_rn(a5, m); a6 = __dmul_rn(a6, m); a7 = __dmul_rn(a7, m);
    } else if (OP == 2) {
      a0 = __fma_rn(a0, m, m); a1 = __fma_rn(a1, m, m); a2 = __fma_rn(a2, m, m); a3 = __fma_rn(a3, m, m);
      a4 = __fma_rn(a4, m, m); a5 = __fma_rn(a5, m, m); a6 = __fma_rn(a6, m, m); a7 = __fma_rn(a7, m, m);
    } else {
      a0 = __ddiv_rn(a0, m); a1 = __ddiv_rn(a1, m); a2 = __ddiv_rn(a2, m); a3 = __ddiv_rn(a3, m);
      a4 = __ddiv_rn(a4, m); a5 = __ddiv_rn(a5, m); a6 = __ddiv_rn(a6, m); a7 = __ddiv_rn(a7, m);
    }
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// 64-bit selects: the cost of routing a value into one of 5 register slots by a per-lane rank
__global__ void selKernel(unsigned long long* cycles, double* sink, int r) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  double v = threadIdx.x;
  int rank = (threadIdx.x + r) % 5;
  const unsigned long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < kIters; ++i) {
    v = __dadd_rn(v, 1.0);
    s0 = rank == 0 ? v : s0;
    s1 = rank == 1 ? v : s1;
    s2 = rank == 2 ? v : s2;
    s3 = rank == 3 ? v : s3;
    s4 = rank == 4 ? v : s4;
    rank = rank == 4 ? 0 : rank + 1;
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + s4;
}

int main() {
  unsigned long long* d_cycles;
  float* d_sink;
  cudaMalloc(&d_cycles, 1024 * sizeof(unsigned long long));
  cudaMalloc(&d_sink, 1024 * 1024 * sizeof(double));
  unsigned long long h[1024];
  const int warps_list[3] = {4, 16, 32};
  for (int wi = 0; wi < 3; ++wi) {
    const int threads = warps_list[wi] * 32;
    for (int vec = 4; vec >= 1; vec >>= 1) {
      for (int p = 0; p < 20; ++p) {
        if (vec == 4) ldsKernel<4><<<148, threads>>>(p, d_cycles, d_sink);
        else if (vec == 2) ldsKernel<2><<<148, threads>>>(p, d_cycles, d_sink);
        else ldsKernel<1><<<148, threads>>>(p, d_cycles, d_sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d_cycles, 148 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        // cycles per warp-level load on one SM = cycles / (iters * warps)
        printf("lds%-3d warps/SM %2d pattern %2d: %.2f cycles per warp-load (SM-wide)\n", vec * 32,
               warps_list[wi], p, double(h[0]) / (double(kIters) * warps_list[wi]));
      }
    }
    const char* names[4] = {"dadd", "dmul", "dfma", "ddiv"};
    for (int op = 0; op < 4; ++op) {
      double* ds = reinterpret_cast<double*>(d_sink);
      if (op == 0) fp64Kernel<0><<<148, threads>>>(d_cycles, ds, 1.0);
      if (op == 1) fp64Kernel<1><<<148, threads>>>(d_cycles, ds, 1.0);
      if (op == 2) fp64Kernel<2><<<148, threads>>>(d_cycles, ds, 1.0);
      if (op == 3) fp64Kernel<3><<<148, threads>>>(d_cycles, ds, 1.0);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d_cycles, 148 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      printf("%s warps/SM %2d: %.2f cycles per warp-instruction (SM-wide)\n", names[op],
             warps_list[wi], double(h[0]) / (double(kIters) * 8 * warps_list[wi]));
    }
    selKernel<<<148, threads>>>(d_cycles, reinterpret_cast<double*>(d_sink), wi);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_cycles, 148 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    printf("sel5x64 warps/SM %2d: %.2f cycles per (dadd + 5 x 64-bit select + rank step) (SM-wide)\n",
           warps_list[wi], double(h[0]) / (double(kIters) * warps_list[wi]));
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
