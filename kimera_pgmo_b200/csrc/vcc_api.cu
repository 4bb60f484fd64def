// VoxelClearingCompression C ABI — placeholder until the hash kernels land (fails loudly).
#include "common.cuh"

struct pgmo_vcc {
  int device;
};

extern "C" {
int pgmo_vcc_create(double, int, pgmo_vcc_t** out) {
  if (out) *out = nullptr;
  pgmo::setLastError("pgmo_vcc_*: not implemented yet");
  return PGMO_ERR_UNSUPPORTED;
}
void pgmo_vcc_destroy(pgmo_vcc_t*) {}
int pgmo_vcc_integrate(pgmo_vcc_t*, uint32_t, const int32_t*, const uint64_t*, const float*,
                       const uint32_t*, double, uint64_t*, int64_t*) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_vcc_prune(pgmo_vcc_t*, double) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_vcc_clear_blocks(pgmo_vcc_t*, uint32_t, const int32_t*) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_vcc_sizes(pgmo_vcc_t*, uint64_t*) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_vcc_snapshot(pgmo_vcc_t*, float*, uint32_t*, uint64_t*, uint32_t*, uint64_t*, uint64_t*) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_vcc_integrate_device(pgmo_vcc_t*, uint32_t, const int32_t*, const uint64_t*, const float*,
                              const uint32_t*, double, uint64_t*, void*) { return PGMO_ERR_UNSUPPORTED; }
}
