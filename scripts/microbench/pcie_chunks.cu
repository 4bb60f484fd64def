// PCIe duplex rate against the size of the DMA operations: 120 MB up and 120 MB down at once,
// each direction split into operations of S bytes (one stream per direction), plus the same
// with 3-row cudaMemcpy2DAsync operations (x / y / z planes of one allocation in one operation).
//   nvcc -O2 -o build/pcie_chunks scripts/microbench/pcie_chunks.cu && build/pcie_chunks
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

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
  return 0;
}
