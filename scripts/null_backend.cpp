// Development aid for scripts/bench_shim.cpp: the deformation entry points of the C ABI as
// no-ops (output = input, every vertex "processed"), to time what the C++ shim itself costs on
// the host — copies, gather / scatter, cache — without a device.  Never shipped, never tested
// against; numbers from it say nothing about the kernels.
#include <cstdlib>
#include <cstring>

#include "pgmo_b200.h"

struct pgmo_deform {};
extern "C" {
const char* pgmo_last_error(void) { return ""; }
int pgmo_device_count(void) { return 1; }
uint64_t pgmo_stamp_from_sec(double sec) { return static_cast<uint64_t>(sec * 1e9); }
int pgmo_host_alloc(uint64_t bytes, void** out) {
  *out = bytes ? std::malloc(bytes) : nullptr;
  return PGMO_OK;
}
int pgmo_host_free(void* ptr) {
  std::free(ptr);
  return PGMO_OK;
}
int pgmo_deform_create(int, pgmo_deform_t** out) {
  *out = new pgmo_deform();
  return PGMO_OK;
}
void pgmo_deform_destroy(pgmo_deform_t* h) { delete h; }
int pgmo_deform_set_controls(pgmo_deform_t*, uint64_t, const double*, const uint64_t*,
                             const uint8_t*, const double*) {
  return PGMO_OK;
}
int pgmo_deform_points(pgmo_deform_t*, uint64_t n_v, const float* x, const float* y,
                       const float* z, const uint64_t*, const uint64_t* subset_idx,
                       uint64_t n_subset, uint32_t, double, int, float* out_x, float* out_y,
                       float* out_z, double*, uint32_t*, uint32_t, uint8_t*, uint8_t* processed) {
  std::memcpy(out_x, x, n_v * sizeof(float));
  std::memcpy(out_y, y, n_v * sizeof(float));
  std::memcpy(out_z, z, n_v * sizeof(float));
  if (processed) std::memset(processed, 1, subset_idx ? n_subset : n_v);
  return PGMO_OK;
}
int pgmo_deform_predeform(pgmo_deform_t*, uint64_t, const float*, const float*, const float*,
                          const int32_t*, uint64_t, float*, float*, float*, uint64_t*,
                          uint64_t* n_to_deform) {
  *n_to_deform = 0;
  return PGMO_OK;
}
int pgmo_deform_mesh_device(pgmo_deform_t*, uint64_t, const void*, const uint64_t*, uint32_t,
                            double, int, void*, uint8_t*, uint64_t*, void*) {
  return PGMO_ERR_UNSUPPORTED;
}
// the multi-device backend of the shim (linked, not used by bench_shim's single-device run)
struct pgmo_deform_group {};
int pgmo_deform_group_create(int, const int*, pgmo_deform_group_t** out) {
  *out = new pgmo_deform_group();
  return PGMO_OK;
}
void pgmo_deform_group_destroy(pgmo_deform_group_t* g) { delete g; }
int pgmo_deform_group_size(pgmo_deform_group_t*) { return 1; }
int pgmo_deform_group_set_controls(pgmo_deform_group_t*, uint64_t, const double*, const uint64_t*,
                                   const uint8_t*, const double*) {
  return PGMO_OK;
}
int pgmo_deform_group_update_poses(pgmo_deform_group_t*, const double*) { return PGMO_OK; }
int pgmo_deform_group_points(pgmo_deform_group_t*, uint64_t n_v, const float* x, const float* y,
                             const float* z, const uint64_t*, uint32_t, double, int, float* out_x,
                             float* out_y, float* out_z, uint8_t* processed) {
  std::memcpy(out_x, x, n_v * sizeof(float));
  std::memcpy(out_y, y, n_v * sizeof(float));
  std::memcpy(out_z, z, n_v * sizeof(float));
  if (processed) std::memset(processed, 1, n_v);
  return PGMO_OK;
}
int pgmo_deform_group_repeat(pgmo_deform_group_t* g, uint64_t n_v, const float* x, const float* y,
                             const float* z, int precision, float* out_x, float* out_y, float* out_z,
                             uint8_t* processed) {
  return pgmo_deform_group_points(g, n_v, x, y, z, nullptr, 0, 0.0, precision, out_x, out_y, out_z,
                                  processed);
}
}
