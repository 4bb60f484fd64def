// PCIe duplex rate against the size of the DMA operations: 120 MB up and 120 MB down at once,
// each direction split into operations of S bytes (one stream per direction), plus the same
// with 3-row cudaMemcpy2DAsync operations (x / y / z planes of one allocation in one operation).
//   nvcc -O2 -o build/pcie_chunks scripts/microbench/pcie_chunks.cu && build/pcie_chunks
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

// stands in for the deformation kernels of one slab: streams `n4` float4 through the SMs
__global__ void busyKernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n4, int rounds) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 v = in[i];
    for (int r = 0; r < rounds; ++r) v.x = v.x * 1.0001f + v.y;
    out[i] = v;
  }
}

int main() {
  const size_t n = 120u << 20;
  char *h_in, *h_out, *d_in, *d_out;
  cudaHostAlloc(&h_in, n, cudaHostAllocDefault);
  cudaHostAlloc(&h_out, n, cudaHostAllocDefault);
  cudaMalloc(&d_in, n);
  cudaMalloc(&d_out, n);
  cudaStream_t up, down;
  cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking);
  auto run = [&](size_t op, int mode) {  // mode 0: up only, 1: down only, 2: both, 3: both, 2D ops of 3 rows
    double best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
      cudaDeviceSynchronize();
      const auto t0 = std::chrono::steady_clock::now();
      if (mode == 3) {
        const size_t plane = n / 3, rows_op = op / 3;
        for (size_t o = 0; o < plane; o += rows_op) {
          const size_t w = o + rows_op <= plane ? rows_op : plane - o;
          cudaMemcpy2DAsync(d_in + o, plane, h_in + o, plane, w, 3, cudaMemcpyHostToDevice, up);
          cudaMemcpy2DAsync(h_out + o, plane, d_out + o, plane, w, 3, cudaMemcpyDeviceToHost, down);
        }
      } else {
        for (size_t o = 0; o < n; o += op) {
          const size_t w = o + op <= n ? op : n - o;
          if (mode != 1) cudaMemcpyAsync(d_in + o, h_in + o, w, cudaMemcpyHostToDevice, up);
          if (mode != 0) cudaMemcpyAsync(h_out + o, d_out + o, w, cudaMemcpyDeviceToHost, down);
        }
      }
      cudaDeviceSynchronize();
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms < best) best = ms;
    }
    return best;
  };
  const char* names[] = {"H2D alone", "D2H alone", "both", "both, 3-row 2D ops"};
  for (size_t mb : {1, 2, 5, 10, 15, 30, 60, 120}) {
    for (int mode = 0; mode < 4; ++mode) {
      const double ms = run(mb << 20, mode);
      const double bytes = (mode >= 2 ? 2.0 : 1.0) * n;
      printf("ops of %3zu MB, %-20s: %.3f ms = %.1f GB/s\n", mb, names[mode], ms, bytes / ms / 1e6);
    }
  }
  // The shape of the host-pointer deformation step without its kernels: chunk c goes up on one
  // stream, a second stream waits for it (the kernels would run here), a third stream waits for
  // that and brings chunk c of the results down.  n_chunks 3-row operations each way.
  cudaStream_t mid;
  cudaStreamCreateWithFlags(&mid, cudaStreamNonBlocking);
  cudaEvent_t in_done[64], k_done[64];
  for (int i = 0; i < 64; ++i) {
    cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&k_done[i], cudaEventDisableTiming);
  }
  char *d_scratch_in, *d_scratch_out;
  cudaMalloc(&d_scratch_in, n);
  cudaMalloc(&d_scratch_out, n);
  for (int with_kernel = 0; with_kernel < 3; ++with_kernel)
  for (int n_chunks : {1, 2, 4, 6, 8, 12, 16}) {
    double best = 1e9;
    const size_t plane = n / 3, per = plane / n_chunks;
    for (int rep = 0; rep < 6; ++rep) {
      cudaDeviceSynchronize();
      const auto t0 = std::chrono::steady_clock::now();
      for (int c = 0; c < n_chunks; ++c) {
        const size_t o = per * c, w = c + 1 == n_chunks ? plane - o : per;
        cudaMemcpy2DAsync(d_in + o, plane, h_in + o, plane, w, 3, cudaMemcpyHostToDevice, up);
        cudaEventRecord(in_done[c], up);
        cudaStreamWaitEvent(mid, in_done[c], 0);
        if (with_kernel) {  // 1: ~HBM-bound pass over 120 MB / n_chunks; 2: the same, 200 dependent FMAs per element
          busyKernel<<<148 * 8, 256, 0, mid>>>(reinterpret_cast<const float4*>(d_scratch_in),
                                               reinterpret_cast<float4*>(d_scratch_out), n / 16 / n_chunks,
                                               with_kernel == 1 ? 1 : 200);
        }
        cudaEventRecord(k_done[c], mid);
        cudaStreamWaitEvent(down, k_done[c], 0);
        cudaMemcpy2DAsync(h_out + o, plane, d_out + o, plane, w, 3, cudaMemcpyDeviceToHost, down);
      }
      cudaDeviceSynchronize();
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms < best) best = ms;
    }
    printf("pipeline of %2d chunks (up -> %s -> down, 3-row operations): %.3f ms\n", n_chunks,
           with_kernel == 0 ? "wait" : with_kernel == 1 ? "streaming kernel" : "compute-heavy kernel", best);
  }
  return 0;
}
