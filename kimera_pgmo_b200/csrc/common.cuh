// Shared plumbing for libpgmo_b200: error reporting, device buffers, launch helpers.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "pgmo_b200.h"

namespace pgmo {

void setLastError(const char* fmt, ...);
const char* lastError();

// Selects the device and verifies that a kernel image for it exists (sm_100a only).
int ensureDevice(int device);

#define PGMO_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      ::pgmo::setLastError("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                   \
                           cudaGetErrorString(err__));                                    \
      return PGMO_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

#define PGMO_CHECK_ARG(cond)                                                              \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::pgmo::setLastError("%s:%d invalid argument: %s", __FILE__, __LINE__, #cond);       \
      return PGMO_ERR_INVALID;                                                            \
    }                                                                                     \
  } while (0)

#define PGMO_TRY(expr)                   \
  do {                                   \
    const int rc__ = (expr);             \
    if (rc__ < 0) return rc__;           \
  } while (0)

// Grow-only device buffer (never shrinks: handles are long-lived and re-used per message).
// cudaMalloc / cudaFree cost tens to hundreds of milliseconds on this platform: callers on a
// message stream ask for generous head room when they have to grow.
// PGMO_TRACE_ALLOC=1: every device / pinned allocation made after start-up is reported on stderr
// (cudaMalloc and cudaFree cost 0.1 - 100 ms here: a steady-state message must not allocate)
inline bool traceAlloc() {
  static const bool on = std::getenv("PGMO_TRACE_ALLOC") != nullptr;
  return on;
}

template <typename T>
struct DevBuf {
  T* ptr = nullptr;
  size_t cap = 0;
  // a buffer that had to grow once will grow again (a mesh that gains vertices between loop
  // closures): with this set, a regrowth allocates twice what is asked for, so that a steadily
  // growing caller pays cudaMalloc + cudaFree (tens to hundreds of ms here) log2 times
  bool regrow_double = false;
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  // keep == true preserves the first `keep_n` elements across a growth
  int reserve(size_t n, cudaStream_t stream = nullptr, size_t keep_n = 0) {
    if (n <= cap) return PGMO_OK;
    size_t new_cap = cap ? cap : 256;
    while (new_cap < n) new_cap += new_cap / 2 + 256;
    if (regrow_double && cap > 0 && new_cap < 2 * n) new_cap = 2 * n;
    T* fresh = nullptr;
    if (traceAlloc()) fprintf(stderr, "pgmo alloc: device %zu -> %zu bytes\n", cap * sizeof(T), new_cap * sizeof(T));
    PGMO_CUDA(cudaMalloc(&fresh, new_cap * sizeof(T)));
    if (ptr && keep_n) {
      PGMO_CUDA(cudaMemcpyAsync(fresh, ptr, keep_n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
      PGMO_CUDA(cudaStreamSynchronize(stream));
    }
    if (ptr) cudaFree(ptr);
    ptr = fresh;
    cap = new_cap;
    return PGMO_OK;
  }
};

// Pinned host staging buffer.
template <typename T>
struct PinnedBuf {
  T* ptr = nullptr;
  size_t cap = 0;
  ~PinnedBuf() {
    if (ptr) cudaFreeHost(ptr);
  }
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  int reserve(size_t n) {
    if (n <= cap) return PGMO_OK;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
    size_t new_cap = n + n / 4 + 64;
    if (traceAlloc()) fprintf(stderr, "pgmo alloc: pinned %zu bytes\n", new_cap * sizeof(T));
    PGMO_CUDA(cudaMallocHost(&ptr, new_cap * sizeof(T)));
    cap = new_cap;
    return PGMO_OK;
  }
};

// per-file kernel preloading (cudaFuncGetAttributes on every kernel: forces the lazy module load
// at handle creation, not inside a message stream)
void preloadCommonKernels();
void preloadDeltaKernels();
void preloadVccKernels();
void preloadOctTreeKernels();
void preloadOctApiKernels();
void preloadM2gKernels();
void preloadMeshInterfaceKernels();
void preloadMeshMsgKernels();
void preloadDeformMiscKernels();
void preloadDeformGridKernels();

inline unsigned divUp(uint64_t a, uint64_t b) { return static_cast<unsigned>((a + b - 1) / b); }

}  // namespace pgmo
