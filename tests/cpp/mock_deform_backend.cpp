// TEST DOUBLE — not part of the product, never shipped, never measured.
//
// Implements the few libpgmo_b200 entry points the deformation half of the C++ shim calls
// (include/kimera_pgmo_b200/{mesh_deformation,deformation_graph}.h) on top of the CPU oracle
// (oracle/liboracle.so), with the argument meaning and return codes include/pgmo_b200.h states.
// tests/test_shim.py links it INSTEAD of libpgmo_b200.so into a second build of
// tests/cpp/test_shim.cpp, so that the CPU suite (`-m "not gpu"`) exercises the shim's HOST
// logic — start index, cache, predeform scatter, staging, threaded gather / scatter, set
// rebuilding, early returns — against the reference's own unit tests.  The GPU suite runs the
// same tests against the real library; libpgmo_b200.so itself has no CPU path.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pgmo_b200.h"

extern "C" {
// oracle/oracle_capi.cpp
int oracle_deform_points(uint64_t n_v, const float* x, const float* y, const float* z,
                         const uint64_t* stamps, uint64_t n_c, const double* ctrl_xyz,
                         const uint64_t* ctrl_stamps, const uint8_t* valid, const double* poses12,
                         uint32_t k, double tol_t, const uint64_t* indices, uint64_t n_indices,
                         int want_trace, float* out_x, float* out_y, float* out_z, double* out_f64,
                         int32_t* assoc, uint32_t assoc_stride, int32_t* assoc_cnt,
                         uint8_t* processed, double* elapsed_s);
uint64_t oracle_stamp_from_sec(double sec);
}

struct pgmo_deform {
  uint64_t n_c = 0;
  bool stamped = false;
  std::vector<double> pos, poses;
  std::vector<uint64_t> stamps;
  std::vector<uint8_t> valid;
};

static thread_local const char* g_error = "";

extern "C" {

const char* pgmo_last_error(void) { return g_error; }
int pgmo_device_count(void) { return 2; }  // lets the shim's device-group path be driven
uint64_t pgmo_stamp_from_sec(double sec) { return oracle_stamp_from_sec(sec); }

int pgmo_host_alloc(uint64_t bytes, void** out) {
  if (std::getenv("MOCK_PINNED_ALLOC_FAILS")) {  // the shim must go on with pageable staging
    *out = nullptr;
    g_error = "test double: page-locking refused";
    return PGMO_ERR_CUDA;
  }
  *out = bytes ? std::malloc(bytes) : nullptr;
  return bytes && !*out ? PGMO_ERR_CUDA : PGMO_OK;
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

int pgmo_deform_set_controls(pgmo_deform_t* h, uint64_t n_c, const double* pos_xyz,
                             const uint64_t* stamps_ns, const uint8_t* valid,
                             const double* poses12) {
  h->n_c = n_c;
  h->stamped = stamps_ns != nullptr;
  h->pos.assign(pos_xyz, pos_xyz + 3 * n_c);
  h->poses.assign(poses12, poses12 + 12 * n_c);
  h->stamps.assign(n_c, 0);
  if (stamps_ns) h->stamps.assign(stamps_ns, stamps_ns + n_c);
  h->valid.assign(n_c, 1);
  if (valid) h->valid.assign(valid, valid + n_c);
  return PGMO_OK;
}

int pgmo_deform_points(pgmo_deform_t* h, uint64_t n_v, const float* x, const float* y,
                       const float* z, const uint64_t* stamps_ns, const uint64_t* subset_idx,
                       uint64_t n_subset, uint32_t k, double tol_t, int /* precision */,
                       float* out_x, float* out_y, float* out_z, double* out_xyz_f64,
                       uint32_t* assoc, uint32_t assoc_stride, uint8_t* assoc_cnt,
                       uint8_t* processed) {
  const uint64_t n_pos = subset_idx ? n_subset : n_v;
  for (uint64_t p = 0; subset_idx && p < n_subset; ++p) {
    if (subset_idx[p] >= n_v) {  // indices->at() / cloud.at()
      g_error = "vertex index out of range";
      return PGMO_ERR_OUT_OF_RANGE;
    }
  }
  std::vector<int32_t> cnt32(assoc_cnt ? n_pos : 0);
  const int rc = oracle_deform_points(
      n_v, x, y, z, stamps_ns, h->n_c, h->pos.data(), h->stamped ? h->stamps.data() : nullptr,
      h->valid.data(), h->poses.data(), k, tol_t, subset_idx, n_subset, 1, out_x, out_y, out_z,
      out_xyz_f64, reinterpret_cast<int32_t*>(assoc), assoc_stride,
      assoc_cnt ? cnt32.data() : nullptr, processed, nullptr);
  if (rc != 0) return PGMO_NOOP;
  for (uint64_t p = 0; p < cnt32.size(); ++p) assoc_cnt[p] = static_cast<uint8_t>(cnt32[p]);
  return PGMO_OK;  // the oracle's -1 padding is the ABI's 0xFFFFFFFF
}

// deformation_graph.h:779-806, the arithmetic of oracle/deform_oracle.cpp:244-266
int pgmo_deform_predeform(pgmo_deform_t* h, uint64_t n_v, const float* x, const float* y,
                          const float* z, const int32_t* graph_indices, uint64_t start_idx,
                          float* out_x, float* out_y, float* out_z, uint64_t* to_deform,
                          uint64_t* n_to_deform) {
  uint64_t n_to = 0;
  for (uint64_t i = start_idx; i < n_v; ++i) {
    const int32_t index = graph_indices[i];
    // values.exists(Symbol(prefix, index)): an index past the table has no pose
    if (index < 0 || static_cast<uint64_t>(index) >= h->n_c || !h->valid[index]) {
      to_deform[n_to++] = i;
      continue;
    }
    const double* T = &h->poses[12 * static_cast<size_t>(index)];
    const double* g = &h->pos[3 * static_cast<size_t>(index)];
    const double p[3] = {double(x[i]) - g[0], double(y[i]) - g[1], double(z[i]) - g[2]};
    float* o[3] = {&out_x[i], &out_y[i], &out_z[i]};
    for (int r = 0; r < 3; ++r) {
      const double rot = (T[3 * r + 0] * p[0] + T[3 * r + 1] * p[1]) + T[3 * r + 2] * p[2];
      *o[r] = static_cast<float>(rot + T[9 + r]);
    }
  }
  *n_to_deform = n_to;
  return PGMO_OK;
}

// pgmo_deform_group_*: the double deforms the whole cloud in one go (what the real group's slabs
// add up to), so that the shim's routing and argument handling are exercised without a device
struct pgmo_deform_group {
  pgmo_deform table;
  int n_devices = 0;
  uint64_t calls = 0;
};
int pgmo_deform_group_create(int n_devices, const int*, pgmo_deform_group_t** out) {
  *out = new pgmo_deform_group();
  (*out)->n_devices = n_devices == 0 ? pgmo_device_count() : n_devices;
  return PGMO_OK;
}
void pgmo_deform_group_destroy(pgmo_deform_group_t* g) { delete g; }
int pgmo_deform_group_size(pgmo_deform_group_t* g) { return g->n_devices; }
int pgmo_deform_group_set_controls(pgmo_deform_group_t* g, uint64_t n_c, const double* pos_xyz,
                                   const uint64_t* stamps_ns, const uint8_t* valid,
                                   const double* poses12) {
  return pgmo_deform_set_controls(&g->table, n_c, pos_xyz, stamps_ns, valid, poses12);
}
int pgmo_deform_group_points(pgmo_deform_group_t* g, uint64_t n_v, const float* x, const float* y,
                             const float* z, const uint64_t* stamps_ns, uint32_t k, double tol_t,
                             int precision, float* out_x, float* out_y, float* out_z,
                             uint8_t* processed) {
  ++g->calls;
  return pgmo_deform_points(&g->table, n_v, x, y, z, stamps_ns, nullptr, 0, k, tol_t, precision, out_x,
                            out_y, out_z, nullptr, nullptr, 0, nullptr, processed);
}

// not reachable from the tests that link this file
int pgmo_deform_mesh_device(pgmo_deform_t*, uint64_t, const void*, const uint64_t*, uint32_t,
                            double, int, void*, uint8_t*, uint64_t*, void*) {
  g_error = "test double: no device path";
  return PGMO_ERR_UNSUPPORTED;
}

}  // extern "C"
