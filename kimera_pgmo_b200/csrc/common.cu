// Error reporting, device selection and the small stateless C entry points.
#include "common.cuh"

#include <cstring>

#include "deform_schedule.h"

namespace pgmo {

namespace {
thread_local char g_last_error[512] = "";
__global__ void probeKernel(int* out) {
  if (out) *out = 100;
}
}  // namespace

void setLastError(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

const char* lastError() { return g_last_error; }

int ensureDevice(int device) {
  int count = 0;
  const cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count <= 0) {
    cudaGetLastError();
    setLastError("no CUDA device available (%s); libpgmo_b200 has no CPU fallback",
                 err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    return PGMO_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    setLastError("device %d out of range (have %d)", device, count);
    return PGMO_ERR_INVALID;
  }
  int current = -1;
  cudaGetDevice(&current);
  if (current != device) PGMO_CUDA(cudaSetDevice(device));
  // the library only carries sm_100a code: fail loudly on anything else
  static thread_local int verified_device = -1;
  if (verified_device != device) {
    cudaFuncAttributes attr;
    const cudaError_t fe = cudaFuncGetAttributes(&attr, probeKernel);
    if (fe != cudaSuccess) {
      cudaGetLastError();
      setLastError("no sm_100a kernel image for device %d (%s)", device, cudaGetErrorString(fe));
      return PGMO_ERR_NO_DEVICE;
    }
    verified_device = device;
  }
  return PGMO_OK;
}

}  // namespace pgmo

extern "C" {

const char* pgmo_last_error(void) { return pgmo::lastError(); }
const char* pgmo_version(void) { return "pgmo_b200 0.1 (sm_100a)"; }
int pgmo_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}
uint64_t pgmo_stamp_from_sec(double sec) { return pgmo::stampFromSec(sec); }

int pgmo_host_alloc(uint64_t bytes, void** out) {
  PGMO_CHECK_ARG(out != nullptr);
  *out = nullptr;
  if (bytes == 0) return PGMO_OK;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    pgmo::setLastError("no CUDA device available; libpgmo_b200 has no CPU fallback");
    return PGMO_ERR_NO_DEVICE;
  }
  const cudaError_t err = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
  if (err != cudaSuccess) {
    cudaGetLastError();  // not sticky: the caller may go on with pageable memory
    *out = nullptr;
    pgmo::setLastError("cudaHostAlloc(%llu bytes) -> %s", static_cast<unsigned long long>(bytes),
                       cudaGetErrorString(err));
    return PGMO_ERR_CUDA;
  }
  return PGMO_OK;
}
int pgmo_host_free(void* ptr) {
  if (ptr) PGMO_CUDA(cudaFreeHost(ptr));
  return PGMO_OK;
}

}  // extern "C"

// CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for a module): pull
// every kernel of this file in when a handle is created instead of inside a message stream.
namespace pgmo {
void preloadCommonKernels() {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, probeKernel);
  cudaGetLastError();
}
}  // namespace pgmo
