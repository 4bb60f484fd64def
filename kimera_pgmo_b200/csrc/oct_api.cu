// OctreeCompression C ABI — placeholder until the exact-replay kernel lands (fails loudly).
#include "common.cuh"

struct pgmo_oct {
  int device;
};

extern "C" {
int pgmo_oct_create(double, int, pgmo_oct_t** out) {
  if (out) *out = nullptr;
  pgmo::setLastError("pgmo_oct_*: not implemented yet");
  return PGMO_ERR_UNSUPPORTED;
}
void pgmo_oct_destroy(pgmo_oct_t*) {}
int pgmo_oct_integrate(pgmo_oct_t*, uint32_t, const int32_t*, const uint64_t*, const float*,
                       const uint32_t*, double, int64_t*, uint8_t*) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_oct_prune(pgmo_oct_t*, double) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_oct_sizes(pgmo_oct_t*, uint64_t*) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_oct_snapshot(pgmo_oct_t*, float*, uint32_t*, uint64_t*, uint32_t*, uint64_t*, float*,
                      double*) { return PGMO_ERR_UNSUPPORTED; }
int pgmo_oct_last_outputs(pgmo_oct_t*, float*, uint32_t*, uint32_t*, uint64_t*) { return PGMO_ERR_UNSUPPORTED; }
}
