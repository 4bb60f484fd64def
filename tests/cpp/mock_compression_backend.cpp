// TEST DOUBLE — not part of the product, never shipped, never measured.
//
// The pgmo_vcc_* / pgmo_oct_* / pgmo_voxblox_* / pgmo_delta_* / pgmo_m2g_* entry points of the C
// ABI on top of the CPU oracle (oracle/liboracle.so), with the argument meaning include/pgmo_b200.h states.  Linked
// together with mock_deform_backend.cpp INSTEAD of libpgmo_b200.so into the CPU suite's build
// of tests/cpp/test_shim.cpp (tests/test_shim.py), so that the shim's host logic for the
// compressors and the frontend — MeshInterface flattening, `remapping` maintenance, MeshDelta
// reconstruction, block filters, getters — runs the
// reference's unit tests without a GPU.  The GPU suite runs the same file against the real
// library; libpgmo_b200.so itself has no CPU path.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "pgmo_b200.h"

extern "C" {
// oracle/oracle_capi.cpp
void* oracle_comp_create(int kind, double resolution);
void oracle_comp_destroy(void* h);
void oracle_comp_integrate(void* h, uint32_t n_blocks, const int32_t* block_idx,
                           const uint64_t* offsets, const float* xyz, const uint32_t* rgba,
                           double stamp_s, int64_t* remap_index, uint8_t* remap_kind,
                           int64_t* keys, double* elapsed_s);
void oracle_comp_prune(void* h, double earliest_s);
void oracle_comp_clear_blocks(void* h, uint32_t n_blocks, const int32_t* block_idx);
void oracle_comp_sizes(void* h, uint64_t* sizes8);
void oracle_comp_snapshot(void* h, float* vertices_xyz, uint32_t* vertices_rgba, uint64_t* stamps,
                          uint32_t* polygons, uint64_t* active, uint64_t* invalid,
                          float* new_vertices_xyz, uint32_t* new_triangles, uint64_t* new_indices,
                          float* active_xyz, double* active_stamps);
void oracle_voxblox_decode(uint32_t n_blocks, const int32_t* block_idx_xyz,
                           const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                           const uint16_t* z, const uint8_t* r, const uint8_t* g, const uint8_t* b,
                           float block_edge_length, float* xyz, uint32_t* rgba);
void* oracle_delta_create(double resolution);
void oracle_delta_destroy(void* h);
int oracle_delta_update(void* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                        const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                        uint64_t stamp_ns, uint64_t* remap_out, double* elapsed_s);
int oracle_delta_update_labeled(void* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                                const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                                const uint32_t* labels, uint64_t stamp_ns, uint64_t* remap_out,
                                double* elapsed_s);
uint64_t oracle_delta_fetch_labels(void* h, uint32_t* labels);
void oracle_delta_archive_by_time(void* h, uint64_t earliest_ns);
void oracle_delta_archive_blocks(void* h, uint32_t n_blocks, const int32_t* block_idx_xyz);
uint64_t oracle_delta_blocks(void* h, uint64_t capacity, int32_t* block_idx_xyz,
                             uint64_t* update_time_ns, uint32_t* sequence_number);
void oracle_delta_sizes(void* h, uint64_t* sizes12);
void oracle_delta_fetch(void* h, float* xyz, uint32_t* rgba, uint64_t* stamps, uint32_t* faces,
                        uint32_t* faces_archive, uint64_t* prev_to_curr2, uint64_t* deleted,
                        uint64_t* observed, uint64_t* new_indices);
void* oracle_m2g_create();
void oracle_m2g_destroy(void* h);
void oracle_m2g_update(void* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                       const uint64_t* block_offsets, const int64_t* mesh_index,
                       const uint8_t* mesh_kind, const int64_t* graph_index,
                       const uint8_t* graph_kind);
uint64_t oracle_m2g_size(void* h);
void oracle_m2g_fetch(void* h, uint64_t* mesh_idx, uint64_t* graph_idx);
}

struct pgmo_vcc {
  void* oracle;
};
struct pgmo_oct {
  void* oracle;
};
struct pgmo_delta {
  void* oracle;
  bool labeled = false;
};
struct pgmo_m2g {
  void* oracle;
};

namespace {
struct Decoded {
  std::vector<float> xyz;
  std::vector<uint32_t> rgba;
  Decoded(uint32_t n_blocks, const int32_t* block_idx, const uint64_t* offsets, const uint16_t* x,
          const uint16_t* y, const uint16_t* z, const uint8_t* r, const uint8_t* g,
          const uint8_t* b, float edge)
      : xyz(3 * offsets[n_blocks]), rgba(offsets[n_blocks]) {
    oracle_voxblox_decode(n_blocks, block_idx, offsets, x, y, z, r, g, b, edge, xyz.data(),
                          rgba.data());
  }
};
}  // namespace

extern "C" {

// ---------------------------------------------------------------- VoxelClearingCompression
int pgmo_vcc_create(double resolution, int, pgmo_vcc_t** out) {
  *out = new pgmo_vcc{oracle_comp_create(0, resolution)};
  return PGMO_OK;
}
void pgmo_vcc_destroy(pgmo_vcc_t* h) {
  if (h) oracle_comp_destroy(h->oracle);
  delete h;
}
int pgmo_vcc_integrate(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                       const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                       double stamp_s, uint64_t* remap_out, int64_t* keys_out) {
  const uint64_t n = block_offsets[n_blocks];
  std::vector<int64_t> index(n);
  std::vector<uint8_t> kind(n);
  oracle_comp_integrate(h->oracle, n_blocks, block_idx_xyz, block_offsets, xyz, rgba, stamp_s,
                        index.data(), kind.data(), keys_out, nullptr);
  for (uint64_t i = 0; i < n; ++i) remap_out[i] = static_cast<uint64_t>(index[i]);
  return PGMO_OK;
}
int pgmo_vcc_integrate_voxblox(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                               const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                               const uint16_t* z, const uint8_t* r, const uint8_t* g,
                               const uint8_t* b, float block_edge_length, double stamp_s,
                               uint64_t* remap_out, int64_t* keys_out) {
  const Decoded d(n_blocks, block_idx_xyz, block_offsets, x, y, z, r, g, b, block_edge_length);
  return pgmo_vcc_integrate(h, n_blocks, block_idx_xyz, block_offsets, d.xyz.data(), d.rgba.data(),
                            stamp_s, remap_out, keys_out);
}
int pgmo_vcc_prune(pgmo_vcc_t* h, double earliest_s) {
  oracle_comp_prune(h->oracle, earliest_s);
  return PGMO_OK;
}
int pgmo_vcc_clear_blocks(pgmo_vcc_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz) {
  oracle_comp_clear_blocks(h->oracle, n_blocks, block_idx_xyz);
  return PGMO_OK;
}
int pgmo_vcc_sizes(pgmo_vcc_t* h, uint64_t* sizes4) {
  uint64_t s[8];
  oracle_comp_sizes(h->oracle, s);
  std::copy(s, s + 4, sizes4);
  return PGMO_OK;
}
int pgmo_vcc_snapshot(pgmo_vcc_t* h, float* vertices_xyz, uint32_t* vertices_rgba,
                      uint64_t* stamps_ns, uint32_t* polygons, uint64_t* active,
                      uint64_t* invalid) {
  oracle_comp_snapshot(h->oracle, vertices_xyz, vertices_rgba, stamps_ns, polygons, active, invalid,
                       nullptr, nullptr, nullptr, nullptr, nullptr);
  if (invalid) {  // the ABI returns them ascending
    uint64_t s[8];
    oracle_comp_sizes(h->oracle, s);
    std::sort(invalid, invalid + s[3]);
  }
  return PGMO_OK;
}
int pgmo_vcc_device_view(pgmo_vcc_t*, const void**, const uint64_t**, uint64_t*) {
  return PGMO_ERR_UNSUPPORTED;  // no device behind the double
}

// ------------------------------------------------------ OctreeCompression / BlockCompression
int pgmo_oct_create(double resolution, int, pgmo_oct_t** out) {
  *out = new pgmo_oct{oracle_comp_create(1, resolution)};
  return PGMO_OK;
}
int pgmo_oct_create_block(double resolution, int, pgmo_oct_t** out) {
  *out = new pgmo_oct{oracle_comp_create(2, resolution)};
  return PGMO_OK;
}
void pgmo_oct_destroy(pgmo_oct_t* h) {
  if (h) oracle_comp_destroy(h->oracle);
  delete h;
}
int pgmo_oct_integrate(pgmo_oct_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                       const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                       double stamp_s, int64_t* remap_index, uint8_t* remap_kind) {
  oracle_comp_integrate(h->oracle, n_blocks, block_idx_xyz, block_offsets, xyz, rgba, stamp_s,
                        remap_index, remap_kind, nullptr, nullptr);
  return PGMO_OK;
}
int pgmo_oct_integrate_voxblox(pgmo_oct_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                               const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                               const uint16_t* z, const uint8_t* r, const uint8_t* g,
                               const uint8_t* b, float block_edge_length, double stamp_s,
                               int64_t* remap_index, uint8_t* remap_kind) {
  const Decoded d(n_blocks, block_idx_xyz, block_offsets, x, y, z, r, g, b, block_edge_length);
  return pgmo_oct_integrate(h, n_blocks, block_idx_xyz, block_offsets, d.xyz.data(), d.rgba.data(),
                            stamp_s, remap_index, remap_kind);
}
int pgmo_oct_prune(pgmo_oct_t* h, double earliest_s) {
  oracle_comp_prune(h->oracle, earliest_s);
  return PGMO_OK;
}
int pgmo_oct_sizes(pgmo_oct_t* h, uint64_t* sizes8) {
  oracle_comp_sizes(h->oracle, sizes8);
  sizes8[3] = 0;
  return PGMO_OK;
}
int pgmo_oct_snapshot(pgmo_oct_t* h, float* vertices_xyz, uint32_t* vertices_rgba,
                      uint64_t* stamps_ns, uint32_t* polygons, uint64_t* active_index,
                      float* active_xyz, double* active_stamps_s) {
  oracle_comp_snapshot(h->oracle, vertices_xyz, vertices_rgba, stamps_ns, polygons, active_index,
                       nullptr, nullptr, nullptr, nullptr, active_xyz, active_stamps_s);
  return PGMO_OK;
}
int pgmo_oct_last_outputs(pgmo_oct_t* h, float* new_vertices_xyz, uint32_t* new_vertices_rgba,
                          uint32_t* new_triangles, uint64_t* new_indices) {
  uint64_t s[8];
  oracle_comp_sizes(h->oracle, s);
  std::vector<uint64_t> indices(s[6]);
  oracle_comp_snapshot(h->oracle, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                       new_vertices_xyz, new_triangles, indices.data(), nullptr, nullptr);
  if (new_indices) std::copy(indices.begin(), indices.end(), new_indices);
  if (new_vertices_rgba) {  // new_vertices[i] is all_vertices_[new_indices[i]] (:356-362)
    std::vector<uint32_t> rgba(s[0]);
    oracle_comp_snapshot(h->oracle, nullptr, rgba.data(), nullptr, nullptr, nullptr, nullptr,
                         nullptr, nullptr, nullptr, nullptr, nullptr);
    for (uint64_t i = 0; i < s[4] && i < indices.size(); ++i) {
      new_vertices_rgba[i] = rgba[indices[i]];
    }
  }
  return PGMO_OK;
}
int pgmo_oct_device_view(pgmo_oct_t*, const void**, const uint64_t**, uint64_t*) {
  return PGMO_ERR_UNSUPPORTED;
}

int pgmo_voxblox_decode(int, uint32_t n_blocks, const int32_t* block_idx_xyz,
                        const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                        const uint16_t* z, const uint8_t* r, const uint8_t* g, const uint8_t* b,
                        float block_edge_length, float* xyz_out, uint32_t* rgba_out) {
  oracle_voxblox_decode(n_blocks, block_idx_xyz, block_offsets, x, y, z, r, g, b,
                        block_edge_length, xyz_out, rgba_out);
  return PGMO_OK;
}

// ------------------------------------------------------------------------ DeltaCompression
int pgmo_delta_create(double resolution, int, pgmo_delta_t** out) {
  *out = new pgmo_delta{oracle_delta_create(resolution)};
  return PGMO_OK;
}
void pgmo_delta_destroy(pgmo_delta_t* h) {
  if (h) oracle_delta_destroy(h->oracle);
  delete h;
}
int pgmo_delta_update(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                      const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                      uint64_t timestamp_ns, uint64_t* remap_out) {
  const int rc = oracle_delta_update(h->oracle, n_blocks, block_idx_xyz, block_offsets, xyz, rgba,
                                     timestamp_ns, remap_out, nullptr);
  return rc == 0 ? PGMO_OK : PGMO_ERR_OUT_OF_RANGE;
}
int pgmo_delta_update_labeled(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                              const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                              const uint32_t* labels, uint64_t timestamp_ns, uint64_t* remap_out) {
  if (labels) h->labeled = true;
  const int rc = oracle_delta_update_labeled(h->oracle, n_blocks, block_idx_xyz, block_offsets, xyz, rgba,
                                             labels, timestamp_ns, remap_out, nullptr);
  return rc == 0 ? PGMO_OK : PGMO_ERR_OUT_OF_RANGE;
}
int pgmo_delta_fetch_labels(pgmo_delta_t* h, uint32_t* labels, int* any_label) {
  if (any_label) *any_label = h->labeled ? 1 : 0;
  if (labels) oracle_delta_fetch_labels(h->oracle, labels);
  return PGMO_OK;
}
int pgmo_delta_archive_by_time(pgmo_delta_t* h, uint64_t earliest_time_ns) {
  oracle_delta_archive_by_time(h->oracle, earliest_time_ns);
  return PGMO_OK;
}
int pgmo_delta_archive_blocks(pgmo_delta_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz) {
  oracle_delta_archive_blocks(h->oracle, n_blocks, block_idx_xyz);
  return PGMO_OK;
}
int pgmo_delta_sizes(pgmo_delta_t* h, uint64_t* sizes12) {
  oracle_delta_sizes(h->oracle, sizes12);
  return PGMO_OK;
}
int pgmo_delta_fetch(pgmo_delta_t* h, float* vertices_xyz, uint32_t* vertices_rgba,
                     uint64_t* stamps_ns, uint32_t* faces, uint32_t* faces_archive,
                     uint64_t* prev_to_curr2, uint64_t* n_prev_to_curr, uint64_t* deleted,
                     uint64_t* n_deleted, uint64_t* observed, uint64_t* n_observed,
                     uint64_t* new_indices, uint64_t* n_new) {
  uint64_t s[12];
  oracle_delta_sizes(h->oracle, s);
  oracle_delta_fetch(h->oracle, vertices_xyz, vertices_rgba, stamps_ns, faces, faces_archive,
                     prev_to_curr2, deleted, observed, new_indices);
  if (n_prev_to_curr) *n_prev_to_curr = s[6];
  if (n_deleted) *n_deleted = s[7];
  if (n_observed) *n_observed = s[8];
  if (n_new) *n_new = s[9];
  return PGMO_OK;
}
int pgmo_delta_blocks(pgmo_delta_t* h, uint64_t capacity, int32_t* block_idx_xyz,
                      uint64_t* update_time_ns, uint32_t* sequence_number, uint64_t* n_blocks) {
  *n_blocks = oracle_delta_blocks(h->oracle, capacity, block_idx_xyz, update_time_ns,
                                  sequence_number);
  return PGMO_OK;
}

// ------------------------------------------- MeshFrontendInterface's mesh -> graph bookkeeping
int pgmo_m2g_create(int, pgmo_m2g_t** out) {
  *out = new pgmo_m2g{oracle_m2g_create()};
  return PGMO_OK;
}
void pgmo_m2g_destroy(pgmo_m2g_t* h) {
  if (h) oracle_m2g_destroy(h->oracle);
  delete h;
}
int pgmo_m2g_update(pgmo_m2g_t* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                    const uint64_t* block_offsets, const int64_t* mesh_index,
                    const uint8_t* mesh_kind, const int64_t* graph_index,
                    const uint8_t* graph_kind) {
  oracle_m2g_update(h->oracle, n_blocks, block_idx_xyz, block_offsets, mesh_index, mesh_kind,
                    graph_index, graph_kind);
  return PGMO_OK;
}
int pgmo_m2g_size(pgmo_m2g_t* h, uint64_t* n_pairs) {
  *n_pairs = oracle_m2g_size(h->oracle);
  return PGMO_OK;
}
int pgmo_m2g_fetch(pgmo_m2g_t* h, uint64_t* mesh_idx, uint64_t* graph_idx) {
  oracle_m2g_fetch(h->oracle, mesh_idx, graph_idx);
  return PGMO_OK;
}
int pgmo_m2g_device_view(pgmo_m2g_t*, const int64_t**, uint64_t*) { return PGMO_ERR_UNSUPPORTED; }

}  // extern "C"
