// ORACLE — TEST INFRASTRUCTURE ONLY.
// Flat C entry points over the CPU restatement so that tests/ and bench.py's cpu_baseline /
// --impl reference legs can drive it through ctypes.  Never linked into the product library.
#include <chrono>
#include <limits>
#include <cstring>
#include <memory>

#include "compression_oracle.h"
#include "delta_oracle.h"
#include "deform_oracle.h"
#include "frontend_oracle.h"

using namespace pgmo_oracle;

namespace {

struct ControlTable {
  std::vector<Point3d> pos;
  std::vector<Timestamp> stamps;
  std::vector<uint8_t> valid;
  std::vector<Pose3> poses;
  void fill(uint64_t n_c, const double* xyz, const uint64_t* st, const uint8_t* v,
            const double* poses12) {
    pos.resize(n_c);
    stamps.assign(n_c, 0);
    valid.assign(n_c, 1);
    poses.resize(n_c);
    for (uint64_t j = 0; j < n_c; ++j) {
      pos[j] = {xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]};
      if (st) stamps[j] = st[j];
      if (v) valid[j] = v[j];
      if (poses12) {
        std::memcpy(poses[j].R, poses12 + 12 * j, 9 * sizeof(double));
        std::memcpy(poses[j].t, poses12 + 12 * j + 9, 3 * sizeof(double));
      }
    }
  }
  ValuesView view() const { return {valid.data(), poses.data(), poses.size()}; }
};

void exportTrace(const DeformTrace& trace, uint64_t n_pos, double* out_f64, int32_t* assoc,
                 uint32_t assoc_stride, int32_t* assoc_cnt, uint8_t* processed) {
  if (processed) std::memset(processed, 0, n_pos);
  if (assoc_cnt) std::memset(assoc_cnt, 0, n_pos * sizeof(int32_t));
  if (assoc) {
    for (uint64_t i = 0; i < n_pos * assoc_stride; ++i) assoc[i] = -1;
  }
  for (size_t e = 0; e < trace.position.size(); ++e) {
    const size_t pos = trace.position[e];
    if (processed) processed[pos] = 1;
    if (assoc_cnt) assoc_cnt[pos] = static_cast<int32_t>(trace.order[e].size());
    if (assoc) {
      for (size_t j = 0; j < trace.order[e].size() && j < assoc_stride; ++j) {
        assoc[pos * assoc_stride + j] = trace.order[e][j];
      }
    }
    if (out_f64) {
      out_f64[3 * pos + 0] = trace.pre_cast[3 * e + 0];
      out_f64[3 * pos + 1] = trace.pre_cast[3 * e + 1];
      out_f64[3 * pos + 2] = trace.pre_cast[3 * e + 2];
    }
  }
}

struct GraphHandle {
  DeformationGraphMeshPath graph;
  std::map<char, ControlTable> tables;
};

struct CompHandle {
  std::unique_ptr<MeshCompression> comp;
  VoxelClearingCompression* vcc = nullptr;
  HashedIndexMapping remapping;  // the caller-owned persistent map of the reference
  std::vector<PointXYZRGBA> new_vertices;
  std::vector<Face> new_triangles;
  std::vector<size_t> new_indices;
};

FlatBlockMesh makeMesh(uint32_t n_blocks, const int32_t* block_idx, const uint64_t* offsets,
                       const float* xyz, const uint32_t* rgba) {
  FlatBlockMesh mesh;
  mesh.blocks.resize(n_blocks);
  mesh.offsets.resize(n_blocks + 1);
  for (uint32_t b = 0; b < n_blocks; ++b) {
    mesh.blocks[b] = {block_idx[3 * b], block_idx[3 * b + 1], block_idx[3 * b + 2]};
  }
  for (uint32_t b = 0; b <= n_blocks; ++b) mesh.offsets[b] = offsets[b];
  const uint64_t n = offsets[n_blocks];
  mesh.vertices.resize(n);
  for (uint64_t i = 0; i < n; ++i) {
    mesh.vertices[i].x = xyz[3 * i];
    mesh.vertices[i].y = xyz[3 * i + 1];
    mesh.vertices[i].z = xyz[3 * i + 2];
    mesh.vertices[i].rgba = rgba ? rgba[i] : 0;
  }
  return mesh;
}

}  // namespace

extern "C" {

// ---- DeltaCompression (delta_oracle.h) ---------------------------------------------------------
struct DeltaHandle {
  explicit DeltaHandle(double res) : comp(res) {}
  DeltaCompression comp;
  std::shared_ptr<MeshDelta> last;
};

void* oracle_delta_create(double resolution) { return new DeltaHandle(resolution); }
void oracle_delta_destroy(void* h) { delete static_cast<DeltaHandle*>(h); }

// DeltaCompression::update.  remap_out: per input vertex the delta index written to
// remapping[block][i].  Returns 0, or -1 if the reference would have thrown (.at()).
int oracle_delta_update_labeled(void* hv, uint32_t n_blocks, const int32_t* block_idx_xyz,
                                const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                                const uint32_t* labels, uint64_t stamp_ns, uint64_t* remap_out,
                                double* elapsed_s);
int oracle_delta_update(void* hv, uint32_t n_blocks, const int32_t* block_idx_xyz,
                        const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                        uint64_t stamp_ns, uint64_t* remap_out, double* elapsed_s) {
  return oracle_delta_update_labeled(hv, n_blocks, block_idx_xyz, block_offsets, xyz, rgba, nullptr,
                                     stamp_ns, remap_out, elapsed_s);
}
// labels (optional): MeshInterface::getActiveSemantics per input vertex, 0xFFFFFFFF = nullopt
int oracle_delta_update_labeled(void* hv, uint32_t n_blocks, const int32_t* block_idx_xyz,
                                const uint64_t* block_offsets, const float* xyz, const uint32_t* rgba,
                                const uint32_t* labels, uint64_t stamp_ns, uint64_t* remap_out,
                                double* elapsed_s) {
  DeltaHandle* h = static_cast<DeltaHandle*>(hv);
  FlatBlockMesh mesh;
  mesh.offsets.push_back(0);
  for (uint32_t b = 0; b < n_blocks; ++b) {
    mesh.blocks.push_back({block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]});
    for (uint64_t i = block_offsets[b]; i < block_offsets[b + 1]; ++i) {
      PointXYZRGBA p;
      p.x = xyz[3 * i];
      p.y = xyz[3 * i + 1];
      p.z = xyz[3 * i + 2];
      p.rgba = rgba ? rgba[i] : 0xFF000000u;
      mesh.vertices.push_back(p);
      if (labels) mesh.labels.push_back(labels[i]);
    }
    mesh.offsets.push_back(mesh.vertices.size());
  }
  std::vector<uint64_t> remap;
  const auto t0 = std::chrono::steady_clock::now();
  try {
    h->last = h->comp.update(mesh, stamp_ns, &remap);
  } catch (const std::out_of_range&) {
    return -1;
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
  if (remap_out) std::copy(remap.begin(), remap.end(), remap_out);
  return 0;
}

void oracle_delta_archive_by_time(void* hv, uint64_t earliest_ns) {
  static_cast<DeltaHandle*>(hv)->comp.archiveBlocksByTime(earliest_ns);
}

void oracle_delta_archive_blocks(void* hv, uint32_t n_blocks, const int32_t* block_idx_xyz) {
  BlockIndices blocks;
  for (uint32_t b = 0; b < n_blocks; ++b) {
    blocks.push_back({block_idx_xyz[3 * b], block_idx_xyz[3 * b + 1], block_idx_xyz[3 * b + 2]});
  }
  static_cast<DeltaHandle*>(hv)->comp.clearArchivedBlocks(blocks);
}

// block_info_map_ in iteration order: index, update_time_ns, sequence_number of up to `capacity`
// blocks; returns how many there are
uint64_t oracle_delta_blocks(void* hv, uint64_t capacity, int32_t* block_idx_xyz,
                             uint64_t* update_time_ns, uint32_t* sequence_number) {
  uint64_t n = 0;
  static_cast<DeltaHandle*>(hv)->comp.forEachBlock(
      [&](const BlockIndex& idx, const DeltaBlockInfo& info) {
        if (n < capacity) {
          if (block_idx_xyz) {
            block_idx_xyz[3 * n] = idx.x;
            block_idx_xyz[3 * n + 1] = idx.y;
            block_idx_xyz[3 * n + 2] = idx.z;
          }
          if (update_time_ns) update_time_ns[n] = info.update_time_ns;
          if (sequence_number) sequence_number[n] = info.sequence_number;
        }
        ++n;
      });
  return n;
}

// sizes12: vertex_start, face_start, n vertices, n archived vertices, n faces, n archived faces,
// n prev_to_curr, n deleted, n observed, n new, active voxels, pending archived faces
void oracle_delta_sizes(void* hv, uint64_t* s) {
  DeltaHandle* h = static_cast<DeltaHandle*>(hv);
  for (int i = 0; i < 12; ++i) s[i] = 0;
  if (h->last) {
    const MeshDelta& d = *h->last;
    s[0] = d.vertex_start;
    s[1] = d.face_start;
    s[2] = d.vertex_updates.size();
    s[3] = d.num_archived_vertices;
    s[4] = d.face_updates.size();
    s[5] = d.face_archive_updates.size();
    s[6] = d.prev_to_curr.size();
    s[7] = d.deleted_indices.size();
    s[8] = d.observed_indices.size();
    s[9] = d.new_indices.size();
  }
  s[10] = h->comp.numActiveVertices();
  s[11] = h->comp.numPendingArchivedFaces();
}

// per delta vertex the label handed to addVertex (0xFFFFFFFF = none); returns
// semantic_updates.size() (== the vertex count iff MeshDelta::hasSemantics())
uint64_t oracle_delta_fetch_labels(void* hv, uint32_t* labels) {
  DeltaHandle* h = static_cast<DeltaHandle*>(hv);
  if (!h->last) return 0;
  const MeshDelta& d = *h->last;
  for (size_t i = 0; labels && i < d.label_of_vertex.size(); ++i) labels[i] = d.label_of_vertex[i];
  return d.semantic_updates.size();
}

void oracle_delta_fetch(void* hv, float* xyz, uint32_t* rgba, uint64_t* stamps, uint32_t* faces,
                        uint32_t* faces_archive, uint64_t* prev_to_curr2, uint64_t* deleted,
                        uint64_t* observed, uint64_t* new_indices) {
  DeltaHandle* h = static_cast<DeltaHandle*>(hv);
  if (!h->last) return;
  const MeshDelta& d = *h->last;
  for (size_t i = 0; i < d.vertex_updates.size(); ++i) {
    if (xyz) {
      xyz[3 * i] = d.vertex_updates[i].x;
      xyz[3 * i + 1] = d.vertex_updates[i].y;
      xyz[3 * i + 2] = d.vertex_updates[i].z;
    }
    if (rgba) rgba[i] = d.vertex_updates[i].rgba;
    if (stamps) stamps[i] = d.stamp_updates[i];
  }
  for (size_t i = 0; faces && i < d.face_updates.size(); ++i) {
    faces[3 * i] = d.face_updates[i].v1;
    faces[3 * i + 1] = d.face_updates[i].v2;
    faces[3 * i + 2] = d.face_updates[i].v3;
  }
  for (size_t i = 0; faces_archive && i < d.face_archive_updates.size(); ++i) {
    faces_archive[3 * i] = d.face_archive_updates[i].v1;
    faces_archive[3 * i + 1] = d.face_archive_updates[i].v2;
    faces_archive[3 * i + 2] = d.face_archive_updates[i].v3;
  }
  size_t k = 0;
  for (const auto& kv : d.prev_to_curr) {
    if (prev_to_curr2) {
      prev_to_curr2[2 * k] = kv.first;
      prev_to_curr2[2 * k + 1] = kv.second;
    }
    ++k;
  }
  k = 0;
  for (const auto v : d.deleted_indices) {
    if (deleted) deleted[k] = v;
    ++k;
  }
  k = 0;
  for (const auto v : d.observed_indices) {
    if (observed) observed[k] = v;
    ++k;
  }
  k = 0;
  for (const auto v : d.new_indices) {
    if (new_indices) new_indices[k] = v;
    ++k;
  }
}

// VoxbloxMsgInterface::getActiveVertex (kimera_pgmo_ros/src/voxblox_msg_interface.cpp:43-62):
// vertex i of block b = (float(u16) * (2.0f / 65535) + float(block index)) * block_edge_length,
// every operation in float (no FMA: the reference is built without -march=native), colour
// (r, g, b, 255).  rgba uses PCL's packing: a << 24 | r << 16 | g << 8 | b.
void oracle_voxblox_decode(uint32_t n_blocks, const int32_t* block_idx_xyz,
                           const uint64_t* block_offsets, const uint16_t* x, const uint16_t* y,
                           const uint16_t* z, const uint8_t* r, const uint8_t* g, const uint8_t* b,
                           float block_edge_length, float* xyz, uint32_t* rgba) {
  const float conv = 2.0f / std::numeric_limits<uint16_t>::max();
  for (uint32_t blk = 0; blk < n_blocks; ++blk) {
    const int32_t* index = block_idx_xyz + 3 * blk;
    for (uint64_t i = block_offsets[blk]; i < block_offsets[blk + 1]; ++i) {
      xyz[3 * i + 0] = (static_cast<float>(x[i]) * conv + static_cast<float>(index[0])) * block_edge_length;
      xyz[3 * i + 1] = (static_cast<float>(y[i]) * conv + static_cast<float>(index[1])) * block_edge_length;
      xyz[3 * i + 2] = (static_cast<float>(z[i]) * conv + static_cast<float>(index[2])) * block_edge_length;
      rgba[i] = 0xFF000000u | (static_cast<uint32_t>(r[i]) << 16) | (static_cast<uint32_t>(g[i]) << 8) | b[i];
    }
  }
}

// conversions::fillFromMsg (kimera_pgmo_ros/include/kimera_pgmo_ros/conversion/mesh_conversion.h:
// 153-199) per vertex: Pos(p.x, p.y, p.z) narrows the doubles to float; colour bytes are
// static_cast<uint8_t>(255.0f * c); stamp = ros::Time::toNSec() = sec * 1e9 + nsec.  Without
// colours the point keeps pcl::PointXYZRGBA's default (0, 0, 0, 255).  rgba in PCL's packing.
void oracle_msg_to_cloud(uint64_t n, const double* vertices_xyz, const float* colors_rgba,
                         const uint32_t* stamps_sec_nsec, float* xyz, uint32_t* rgba,
                         uint64_t* stamps_ns) {
  constexpr float color_conv_factor = 1.0f * std::numeric_limits<uint8_t>::max();
  for (uint64_t i = 0; i < n; ++i) {
    for (int a = 0; a < 3; ++a) xyz[3 * i + a] = static_cast<float>(vertices_xyz[3 * i + a]);
    uint32_t packed = 0xFF000000u;
    if (colors_rgba) {
      const float* c = colors_rgba + 4 * i;
      const uint8_t r = static_cast<uint8_t>(color_conv_factor * c[0]);
      const uint8_t g = static_cast<uint8_t>(color_conv_factor * c[1]);
      const uint8_t b = static_cast<uint8_t>(color_conv_factor * c[2]);
      const uint8_t al = static_cast<uint8_t>(color_conv_factor * c[3]);
      packed = (static_cast<uint32_t>(al) << 24) | (static_cast<uint32_t>(r) << 16) |
               (static_cast<uint32_t>(g) << 8) | b;
    }
    rgba[i] = packed;
    if (stamps_sec_nsec && stamps_ns) {
      stamps_ns[i] = static_cast<uint64_t>(stamps_sec_nsec[2 * i]) * 1000000000ull + stamps_sec_nsec[2 * i + 1];
    }
  }
}

// conversions::fillMsg (:69-139): p.x = pos.x(); colour = (1.0f / 255) * byte;
// ros::Time::fromNSec: sec = ns / 1e9, nsec = ns % 1e9.
void oracle_cloud_to_msg(uint64_t n, const float* xyz, const uint32_t* rgba, const uint64_t* stamps_ns,
                         double* vertices_xyz, float* colors_rgba, uint32_t* stamps_sec_nsec) {
  constexpr float color_conv_factor = 1.0f / std::numeric_limits<uint8_t>::max();
  for (uint64_t i = 0; i < n; ++i) {
    for (int a = 0; a < 3; ++a) vertices_xyz[3 * i + a] = xyz[3 * i + a];
    if (colors_rgba && rgba) {
      const uint32_t p = rgba[i];
      colors_rgba[4 * i + 0] = color_conv_factor * static_cast<uint8_t>(p >> 16);
      colors_rgba[4 * i + 1] = color_conv_factor * static_cast<uint8_t>(p >> 8);
      colors_rgba[4 * i + 2] = color_conv_factor * static_cast<uint8_t>(p);
      colors_rgba[4 * i + 3] = color_conv_factor * static_cast<uint8_t>(p >> 24);
    }
    if (stamps_sec_nsec && stamps_ns) {
      stamps_sec_nsec[2 * i] = static_cast<uint32_t>(stamps_ns[i] / 1000000000ull);
      stamps_sec_nsec[2 * i + 1] = static_cast<uint32_t>(stamps_ns[i] % 1000000000ull);
    }
  }
}

// deformation::deformPoints (mesh_deformation.h:63-213).  out_* must be pre-filled by the
// caller (the reference leaves skipped vertices untouched); in-place (out == in) is allowed.
// Returns 0 = ran, 1 = reference early return.  elapsed_s (optional) = time inside the call.
int oracle_deform_points(uint64_t n_v, const float* x, const float* y, const float* z,
                         const uint64_t* stamps, uint64_t n_c, const double* ctrl_xyz,
                         const uint64_t* ctrl_stamps, const uint8_t* valid,
                         const double* poses12, uint32_t k, double tol_t,
                         const uint64_t* indices, uint64_t n_indices, int want_trace,
                         float* out_x, float* out_y, float* out_z, double* out_f64,
                         int32_t* assoc, uint32_t assoc_stride, int32_t* assoc_cnt,
                         uint8_t* processed, double* elapsed_s) {
  ControlTable table;
  table.fill(n_c, ctrl_xyz, ctrl_stamps, valid, poses12);
  DeformIO io;
  io.num_vertices = n_v;
  io.in_x = x;
  io.in_y = y;
  io.in_z = z;
  io.stamps = stamps;
  io.out_x = out_x;
  io.out_y = out_y;
  io.out_z = out_z;
  std::vector<size_t> idx;
  if (indices) idx.assign(indices, indices + n_indices);
  std::vector<std::set<size_t>> map;
  DeformTrace trace;
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = deformPoints(io, map, table.pos, table.stamps, table.view(), k, tol_t,
                              indices ? &idx : nullptr, want_trace ? &trace : nullptr);
  const auto t1 = std::chrono::steady_clock::now();
  if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
  if (want_trace) {
    exportTrace(trace, indices ? n_indices : n_v, out_f64, assoc, assoc_stride, assoc_cnt,
                processed);
  }
  return rc;
}

// DeformationGraph mesh path (deformation_graph.h:715-870)
void* oracle_dgraph_create() { return new GraphHandle(); }
void oracle_dgraph_destroy(void* h) { delete static_cast<GraphHandle*>(h); }
void oracle_dgraph_set_controls(void* h, char prefix, uint64_t n_c, const double* ctrl_xyz,
                                const uint64_t* ctrl_stamps, const uint8_t* valid,
                                const double* poses12) {
  auto* g = static_cast<GraphHandle*>(h);
  auto& t = g->tables[prefix];
  t.fill(n_c, ctrl_xyz, ctrl_stamps, valid, poses12);
  g->graph.setControlPoints(prefix, t.pos, t.stamps);
}
void oracle_dgraph_set_recalculate(void* h, int v) {
  static_cast<GraphHandle*>(h)->graph.setRecalculate(v != 0);
}
// returns the start index used, or -1 when the prefix is unknown (reference: silent return)
int64_t oracle_dgraph_deform(void* h, char prefix, uint64_t n_v, const float* x, const float* y,
                             const float* z, const uint64_t* stamps, uint32_t k, double tol_t,
                             const int32_t* graph_indices, int start_index_hint, float* out_x,
                             float* out_y, float* out_z, int32_t* assoc, uint32_t assoc_stride,
                             int32_t* assoc_cnt) {
  auto* g = static_cast<GraphHandle*>(h);
  const auto it = g->tables.find(prefix);
  if (it == g->tables.end()) return -1;
  DeformIO io;
  io.num_vertices = n_v;
  io.in_x = x;
  io.in_y = y;
  io.in_z = z;
  io.stamps = stamps;
  io.out_x = out_x;
  io.out_y = out_y;
  io.out_z = out_z;
  std::vector<int> gi;
  if (graph_indices) gi.assign(graph_indices, graph_indices + n_v);
  std::vector<std::set<size_t>> vgm;
  DeformTrace trace;
  const size_t start = g->graph.deformPoints(io, prefix, it->second.view(), k, tol_t,
                                             graph_indices ? &gi : nullptr, start_index_hint,
                                             assoc ? &vgm : nullptr, &trace);
  if (assoc) {
    // vertex_graph_map as the reference builds it (deformation_graph.h:857-866), in blend order
    for (uint64_t i = 0; i < n_v * assoc_stride; ++i) assoc[i] = -1;
    if (assoc_cnt) std::memset(assoc_cnt, 0, n_v * sizeof(int32_t));
    for (size_t e = 0; e < trace.vertex.size(); ++e) {
      // start == 0: entry e of the map; start != 0: scattered to the vertex index
      const size_t slot = start == 0 ? e : trace.vertex[e];
      if (assoc_cnt) assoc_cnt[slot] = static_cast<int32_t>(trace.order[e].size());
      for (size_t j = 0; j < trace.order[e].size() && j < assoc_stride; ++j) {
        assoc[slot * assoc_stride + j] = trace.order[e][j];
      }
    }
  }
  return static_cast<int64_t>(start);
}

// MeshCompression family.  kind: 0 = VoxelClearingCompression, 1 = OctreeCompression
void* oracle_comp_create(int kind, double resolution) {
  auto* h = new CompHandle();
  if (kind == 0) {
    h->vcc = new VoxelClearingCompression(resolution);
    h->comp.reset(h->vcc);
  } else if (kind == 2) {
    h->comp.reset(new BlockCompression(resolution));
  } else {
    h->comp.reset(new OctreeCompression(resolution));
  }
  return h;
}
void oracle_comp_destroy(void* h) { delete static_cast<CompHandle*>(h); }

// compressAndIntegrate with the handle's persistent remapping.  remap_index/remap_kind: dense
// per-input-vertex record (see RemapRecord); keys: 3 int64 per input vertex (voxel clearing).
void oracle_comp_integrate(void* hv, uint32_t n_blocks, const int32_t* block_idx,
                           const uint64_t* offsets, const float* xyz, const uint32_t* rgba,
                           double stamp_s, int64_t* remap_index, uint8_t* remap_kind,
                           int64_t* keys, double* elapsed_s) {
  auto* h = static_cast<CompHandle*>(hv);
  const FlatBlockMesh mesh = makeMesh(n_blocks, block_idx, offsets, xyz, rgba);
  RemapRecord rec;
  const auto t0 = std::chrono::steady_clock::now();
  h->comp->compressAndIntegrate(mesh, h->new_vertices, h->new_triangles, h->new_indices,
                                h->remapping, stamp_s, &rec);
  const auto t1 = std::chrono::steady_clock::now();
  if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
  const uint64_t n = offsets[n_blocks];
  for (uint64_t i = 0; i < n && i < rec.index.size(); ++i) {
    if (remap_index) remap_index[i] = rec.index[i];
    if (remap_kind) remap_kind[i] = rec.kind[i];
  }
  if (keys && h->vcc) {
    const auto& k = h->vcc->lastKeys();
    for (uint64_t i = 0; i < n && i < k.size(); ++i) {
      keys[3 * i] = k[i].x;
      keys[3 * i + 1] = k[i].y;
      keys[3 * i + 2] = k[i].z;
    }
  }
}
void oracle_comp_prune(void* h, double earliest_s) {
  static_cast<CompHandle*>(h)->comp->pruneStoredMesh(earliest_s);
}
void oracle_comp_clear_blocks(void* hv, uint32_t n_blocks, const int32_t* block_idx) {
  BlockIndices b(n_blocks);
  for (uint32_t i = 0; i < n_blocks; ++i) b[i] = {block_idx[3 * i], block_idx[3 * i + 1], block_idx[3 * i + 2]};
  static_cast<CompHandle*>(hv)->comp->clearArchivedBlocks(b);
}

// sizes: [num_vertices, num_polygons, num_active, num_invalid, num_new_vertices,
//         num_new_triangles, num_new_indices, num_active_xyz]
void oracle_comp_sizes(void* hv, uint64_t* sizes) {
  auto* h = static_cast<CompHandle*>(hv);
  sizes[0] = h->comp->getVertices().size();
  sizes[1] = h->comp->getStoredPolygons().size();
  sizes[2] = h->comp->getActiveVerticesIndex().size();
  sizes[3] = h->comp->getInvalidIndices().size();
  sizes[4] = h->new_vertices.size();
  sizes[5] = h->new_triangles.size();
  sizes[6] = h->new_indices.size();
  sizes[7] = h->comp->getActiveVertices().size();
}
// any pointer may be null; buffers sized from oracle_comp_sizes
void oracle_comp_snapshot(void* hv, float* vertices_xyz, uint32_t* vertices_rgba,
                          uint64_t* stamps, uint32_t* polygons, uint64_t* active,
                          uint64_t* invalid, float* new_vertices_xyz, uint32_t* new_triangles,
                          uint64_t* new_indices, float* active_xyz, double* active_stamps) {
  auto* h = static_cast<CompHandle*>(hv);
  const auto& v = h->comp->getVertices();
  for (size_t i = 0; i < v.size(); ++i) {
    if (vertices_xyz) {
      vertices_xyz[3 * i] = v[i].x;
      vertices_xyz[3 * i + 1] = v[i].y;
      vertices_xyz[3 * i + 2] = v[i].z;
    }
    if (vertices_rgba) vertices_rgba[i] = v[i].rgba;
    if (stamps) stamps[i] = h->comp->getTimestamps()[i];
  }
  const auto& p = h->comp->getStoredPolygons();
  if (polygons) {
    for (size_t i = 0; i < p.size(); ++i) {
      polygons[3 * i] = p[i][0];
      polygons[3 * i + 1] = p[i][1];
      polygons[3 * i + 2] = p[i][2];
    }
  }
  const auto& a = h->comp->getActiveVerticesIndex();
  if (active) {
    for (size_t i = 0; i < a.size(); ++i) active[i] = a[i];
  }
  const auto inv = h->comp->getInvalidIndices();
  if (invalid) {
    for (size_t i = 0; i < inv.size(); ++i) invalid[i] = inv[i];
  }
  if (new_vertices_xyz) {
    for (size_t i = 0; i < h->new_vertices.size(); ++i) {
      new_vertices_xyz[3 * i] = h->new_vertices[i].x;
      new_vertices_xyz[3 * i + 1] = h->new_vertices[i].y;
      new_vertices_xyz[3 * i + 2] = h->new_vertices[i].z;
    }
  }
  if (new_triangles) {
    for (size_t i = 0; i < h->new_triangles.size(); ++i) {
      new_triangles[3 * i] = h->new_triangles[i][0];
      new_triangles[3 * i + 1] = h->new_triangles[i][1];
      new_triangles[3 * i + 2] = h->new_triangles[i][2];
    }
  }
  if (new_indices) {
    for (size_t i = 0; i < h->new_indices.size(); ++i) new_indices[i] = h->new_indices[i];
  }
  const auto& ax = h->comp->getActiveVertices();
  if (active_xyz) {
    for (size_t i = 0; i < ax.size(); ++i) {
      active_xyz[3 * i] = ax[i].x;
      active_xyz[3 * i + 1] = ax[i].y;
      active_xyz[3 * i + 2] = ax[i].z;
    }
  }
  if (active_stamps) {
    const auto& as = h->comp->getActiveVerticesTimestamps();
    for (size_t i = 0; i < as.size(); ++i) active_stamps[i] = as[i];
  }
}
// persistent remapping lookup: -1 if the block or the local index has no entry
int64_t oracle_comp_remap_lookup(void* hv, int32_t bx, int32_t by, int32_t bz, uint64_t local) {
  auto* h = static_cast<CompHandle*>(hv);
  const auto b = h->remapping.find(BlockIndex{bx, by, bz});
  if (b == h->remapping.end()) return -1;
  const auto e = b->second.find(local);
  if (e == b->second.end()) return -1;
  return static_cast<int64_t>(e->second);
}
int64_t oracle_comp_remap_block_size(void* hv, int32_t bx, int32_t by, int32_t bz) {
  auto* h = static_cast<CompHandle*>(hv);
  const auto b = h->remapping.find(BlockIndex{bx, by, bz});
  if (b == h->remapping.end()) return -1;
  return static_cast<int64_t>(b->second.size());
}

uint64_t oracle_stamp_from_sec(double sec) { return stampFromSec(sec); }

// ---- mesh -> graph index bookkeeping (frontend_oracle.h) ----------------------------------------
void* oracle_m2g_create() { return new MeshToGraphOracle(); }
void oracle_m2g_destroy(void* h) { delete static_cast<MeshToGraphOracle*>(h); }
void oracle_m2g_update(void* h, uint32_t n_blocks, const int32_t* block_idx_xyz,
                       const uint64_t* block_offsets, const int64_t* mesh_index,
                       const uint8_t* mesh_kind, const int64_t* graph_index,
                       const uint8_t* graph_kind) {
  static_cast<MeshToGraphOracle*>(h)->update(n_blocks, block_idx_xyz, block_offsets, mesh_index,
                                             mesh_kind, graph_index, graph_kind);
}
uint64_t oracle_m2g_size(void* h) { return static_cast<MeshToGraphOracle*>(h)->meshToGraph().size(); }
// pairs in ascending mesh index
void oracle_m2g_fetch(void* h, uint64_t* mesh_idx, uint64_t* graph_idx) {
  size_t i = 0;
  for (const auto& kv : static_cast<MeshToGraphOracle*>(h)->meshToGraph()) {
    mesh_idx[i] = kv.first;
    graph_idx[i] = kv.second;
    ++i;
  }
}
// dense msg_to_mesh_idx_[block] (which 0) / msg_to_graph_idx_[block] (which 1); returns its
// length (-1: unknown block), copies min(length, cap) entries (-1 = no entry)
int64_t oracle_m2g_block_mapping(void* h, int32_t bx, int32_t by, int32_t bz, int which,
                                 int64_t* out, uint64_t cap) {
  const auto* v = static_cast<MeshToGraphOracle*>(h)->blockMapping({bx, by, bz}, which);
  if (!v) return -1;
  for (size_t i = 0; i < v->size() && i < cap; ++i) out[i] = (*v)[i];
  return static_cast<int64_t>(v->size());
}

}  // extern "C"
