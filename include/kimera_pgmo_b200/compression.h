// Drop-in for Kimera-PGMO's mesh-frontend compressors:
//   MeshInterface / PclMeshInterface   include/kimera_pgmo/utils/mesh_interface.h:17-37,
//                                      utils/pcl_mesh_interface.h:17-45
//   MeshCompression (public surface)   include/kimera_pgmo/compression/mesh_compression.h:25-148
//   VoxelClearingCompression           include/.../voxel_clearing_compression.h:14-97
//   OctreeCompression                  include/.../octree_compression.h:29-66
// Same class names, constructor (double resolution), compressAndIntegrate / pruneStoredMesh /
// clearArchivedBlocks / getters; state lives in HBM behind pgmo_vcc_t / pgmo_oct_t handles.
#pragma once

#include <array>
#include <limits>
#include <map>

#include <memory>
#include <optional>
#include <stdexcept>
#include <vector>

#include "kimera_pgmo_b200/mesh_deformation.h"

namespace kimera_pgmo_b200 {

// utils/mesh_interface.h:17-37
struct MeshInterface {
  using Ptr = std::shared_ptr<MeshInterface>;
  virtual ~MeshInterface() = default;
  virtual const BlockIndices& blockIndices() const = 0;
  virtual void markBlockActive(const BlockIndex& block) const = 0;
  virtual size_t activeBlockSize() const = 0;
  virtual pcl::PointXYZRGBA getActiveVertex(size_t i) const = 0;
  virtual bool hasSemantics() const { return false; }
  virtual std::optional<uint32_t> getActiveSemantics(size_t /* index */) const {
    return std::nullopt;
  }
  virtual Ptr clone() const = 0;
};

// utils/pcl_mesh_interface.h:17-45: one block (0,0,0) of 3 * faces unindexed vertices
class PclMeshInterface : public MeshInterface {
 public:
  explicit PclMeshInterface(const pcl::PolygonMesh& mesh)
      : cloud_(new pcl::PointCloud<pcl::PointXYZRGBA>()), faces_(mesh.polygons) {
    pcl::fromPCLPointCloud2(mesh.cloud, *cloud_);
    indices_.push_back(BlockIndex::Zero());
  }
  const BlockIndices& blockIndices() const override { return indices_; }
  void markBlockActive(const BlockIndex&) const override {}
  size_t activeBlockSize() const override { return 3 * faces_.size(); }
  pcl::PointXYZRGBA getActiveVertex(size_t i) const override {
    return cloud_->at(faces_.at(i / 3).vertices.at(i % 3));
  }
  MeshInterface::Ptr clone() const override { return std::make_shared<PclMeshInterface>(*this); }

 protected:
  pcl::PointCloud<pcl::PointXYZRGBA>::Ptr cloud_;
  std::vector<pcl::Vertices> faces_;
  BlockIndices indices_;
};

// kimera_pgmo_ros/include/kimera_pgmo_ros/voxblox_msg_interface.h + src/voxblox_msg_interface.cpp.
// The message types mirror the fields of voxblox_msgs::Mesh / MeshBlock that the interface reads
// (a ROS build passes the real message through a two-line adapter).  The
// compressors recognise this interface and upload the packed u16 / u8 arrays as they are: the
// per-vertex decode runs on the GPU (pgmo_*_integrate_voxblox).
struct VoxbloxMeshBlock {
  int64_t index[3] = {0, 0, 0};
  std::vector<uint16_t> x, y, z;
  std::vector<uint8_t> r, g, b;
};
struct VoxbloxMesh {
  float block_edge_length = 0.f;
  std::vector<VoxbloxMeshBlock> mesh_blocks;
};

class VoxbloxMsgInterface : public MeshInterface {
 public:
  explicit VoxbloxMsgInterface(const VoxbloxMesh& mesh) : active_mesh_block_(nullptr), mesh_(mesh) {
    size_t count = 0;
    for (const auto& block : mesh_.mesh_blocks) {
      mesh_blocks_.emplace_back(block.index[0], block.index[1], block.index[2]);
      block_lookup_[{block.index[0], block.index[1], block.index[2]}] = count;
      ++count;
    }
  }
  const BlockIndices& blockIndices() const override { return mesh_blocks_; }
  void markBlockActive(const BlockIndex& block) const override {
    const auto iter = block_lookup_.find({block.x(), block.y(), block.z()});
    active_mesh_block_ = iter == block_lookup_.end() ? nullptr : &mesh_.mesh_blocks.at(iter->second);
  }
  size_t activeBlockSize() const override {
    return active_mesh_block_ ? active_mesh_block_->x.size() : 0;
  }
  pcl::PointXYZRGBA getActiveVertex(size_t i) const override {  // :43-62
    const VoxbloxMeshBlock& blk = *active_mesh_block_;
    pcl::PointXYZRGBA point;
    point.x = (static_cast<float>(blk.x.at(i)) * point_conv_factor_ + static_cast<float>(blk.index[0])) *
              mesh_.block_edge_length;
    point.y = (static_cast<float>(blk.y.at(i)) * point_conv_factor_ + static_cast<float>(blk.index[1])) *
              mesh_.block_edge_length;
    point.z = (static_cast<float>(blk.z.at(i)) * point_conv_factor_ + static_cast<float>(blk.index[2])) *
              mesh_.block_edge_length;
    point.r = blk.r[i];
    point.g = blk.g[i];
    point.b = blk.b[i];
    point.a = std::numeric_limits<uint8_t>::max();
    return point;
  }
  MeshInterface::Ptr clone() const override { return std::make_shared<VoxbloxMsgInterface>(*this); }
  const VoxbloxMesh& message() const { return mesh_; }

 protected:
  mutable const VoxbloxMeshBlock* active_mesh_block_;
  std::map<std::array<int64_t, 3>, size_t> block_lookup_;  // BlockIndexMap<size_t> upstream
  BlockIndices mesh_blocks_;
  const VoxbloxMesh& mesh_;
  inline static constexpr float point_conv_factor_ = 2.0f / std::numeric_limits<uint16_t>::max();
};

namespace detail {
// MeshInterface -> the flat arrays of the C ABI (blockIndices() order, cursor protocol kept)
struct FlatMesh {
  std::vector<int32_t> block_idx;
  std::vector<uint64_t> offsets;
  std::vector<float> xyz;
  std::vector<uint32_t> rgba;
  std::vector<uint32_t> labels;  // getActiveSemantics (0xFFFFFFFF = nullopt) when hasSemantics()
  // voxblox messages stay packed (decoded on the GPU)
  bool packed = false;
  std::vector<uint16_t> x16, y16, z16;
  std::vector<uint8_t> r8, g8, b8;
  float block_edge_length = 0.f;
  // allow_packed = false: every interface goes through getActiveVertex (compressors without a
  // packed entry point); with_semantics: also collect getActiveSemantics
  explicit FlatMesh(const MeshInterface& mesh, bool allow_packed = true, bool with_semantics = false) {
    offsets.push_back(0);
    const auto* vb = allow_packed ? dynamic_cast<const VoxbloxMsgInterface*>(&mesh) : nullptr;
    if (vb) {
      packed = true;
      block_edge_length = vb->message().block_edge_length;
      for (const auto& blk : vb->message().mesh_blocks) {
        for (int a = 0; a < 3; ++a) block_idx.push_back(static_cast<int32_t>(blk.index[a]));
        const size_t n = blk.x.size();
        if (blk.y.size() != n || blk.z.size() != n || blk.r.size() < n || blk.g.size() < n ||
            blk.b.size() < n) {
          throw std::out_of_range("voxblox mesh block: y/z/.at(i) or colour arrays shorter than x");
        }
        x16.insert(x16.end(), blk.x.begin(), blk.x.end());
        y16.insert(y16.end(), blk.y.begin(), blk.y.end());
        z16.insert(z16.end(), blk.z.begin(), blk.z.end());
        r8.insert(r8.end(), blk.r.begin(), blk.r.begin() + n);
        g8.insert(g8.end(), blk.g.begin(), blk.g.begin() + n);
        b8.insert(b8.end(), blk.b.begin(), blk.b.begin() + n);
        offsets.push_back(x16.size());
      }
      return;
    }
    for (const auto& block : mesh.blockIndices()) {
      block_idx.push_back(block.x());
      block_idx.push_back(block.y());
      block_idx.push_back(block.z());
      mesh.markBlockActive(block);
      const size_t n = mesh.activeBlockSize();
      for (size_t i = 0; i < n; ++i) {
        const pcl::PointXYZRGBA p = mesh.getActiveVertex(i);
        xyz.push_back(p.x);
        xyz.push_back(p.y);
        xyz.push_back(p.z);
        rgba.push_back(p.rgba);
        if (with_semantics && mesh.hasSemantics()) {
          const std::optional<uint32_t> label = mesh.getActiveSemantics(i);
          labels.push_back(label ? *label : 0xFFFFFFFFu);
        }
      }
      offsets.push_back(rgba.size());
    }
  }
  uint32_t numBlocks() const { return static_cast<uint32_t>(offsets.size() - 1); }
  size_t numVertices() const { return static_cast<size_t>(offsets.back()); }
};

inline pcl::PointXYZRGBA makePoint(const float* xyz, uint32_t rgba) {
  pcl::PointXYZRGBA p;
  p.x = xyz[0];
  p.y = xyz[1];
  p.z = xyz[2];
  p.rgba = rgba;
  return p;
}
inline void fillPolygons(const std::vector<uint32_t>& flat, std::vector<pcl::Vertices>* out) {
  out->clear();
  out->resize(flat.size() / 3);
  for (size_t f = 0; f < out->size(); ++f) {
    (*out)[f].vertices = {flat[3 * f], flat[3 * f + 1], flat[3 * f + 2]};
  }
}
}  // namespace detail

// mesh_compression.h:25-148 — the public surface shared by the compressors
class MeshCompression {
 public:
  using PointCloud = pcl::PointCloud<pcl::PointXYZRGBA>;
  using PointCloudXYZ = pcl::PointCloud<pcl::PointXYZ>;

  explicit MeshCompression(double resolution) : resolution_(resolution) {}
  virtual ~MeshCompression() = default;

  virtual void compressAndIntegrate(const MeshInterface& mesh, PointCloud& new_vertices,
                                    std::vector<pcl::Vertices>& new_triangles,
                                    std::vector<size_t>& new_indices,
                                    HashedIndexMapping& remapping, double stamp_in_sec) = 0;
  virtual void pruneStoredMesh(double earliest_time_sec) = 0;
  virtual std::vector<size_t> getInvalidIndices() const { return {}; }
  virtual void clearArchivedBlocks(const BlockIndices& /* blocks */) {}

  virtual void getVertices(PointCloud::Ptr vertices) = 0;
  virtual void getTimestamps(std::shared_ptr<std::vector<Timestamp>> vertex_stamps) = 0;
  virtual void getStoredPolygons(std::shared_ptr<std::vector<pcl::Vertices>> polygons) = 0;
  virtual size_t getNumVertices() const = 0;
  virtual const std::vector<size_t>& getActiveVerticesIndex() const = 0;

  // Extension (no reference counterpart): getVertices / getTimestamps without the copy —
  // all_vertices_ as 16-byte {x, y, z, rgba} records and all_vertex_stamps_ in HBM, valid until
  // the next compressAndIntegrate / pruneStoredMesh.  DeformationGraphMeshPath::
  // deformMeshResident consumes it so that the compressed mesh never crosses PCIe.
  struct DeviceView {
    const void* vertices_xyzw = nullptr;
    const uint64_t* stamps_ns = nullptr;
    size_t num_vertices = 0;
  };
  virtual DeviceView deviceView() const = 0;

  // Extension (no reference counterpart): what the last compressAndIntegrate did to `remapping`,
  // as a dense record per input vertex in blockIndices() order - kind 0 = no write, 1 =
  // overwrite, 2 = insert-if-absent; an empty kind vector = all overwrite.
  // MeshFrontendInterface hands it to pgmo_m2g_update.
  const std::vector<int64_t>& lastRemapIndex() const { return last_index_; }
  const std::vector<uint8_t>& lastRemapKind() const { return last_kind_; }

 protected:
  double resolution_;
  std::vector<int64_t> last_index_;
  std::vector<uint8_t> last_kind_;
};
using MeshCompressionPtr = std::shared_ptr<MeshCompression>;

class VoxelClearingCompression : public MeshCompression {
 public:
  explicit VoxelClearingCompression(double resolution, int device = 0)
      : MeshCompression(resolution) {
    pgmo_vcc_t* h = nullptr;
    detail::throwOnError(pgmo_vcc_create(resolution, device, &h));
    handle_.reset(h, [](pgmo_vcc_t* p) { pgmo_vcc_destroy(p); });
  }

  // voxel_clearing_compression.cpp:34-47: outputs are `remapping` and the getters; the three
  // new_* arguments are cleared and left empty, as in the reference
  void compressAndIntegrate(const MeshInterface& mesh, PointCloud& new_vertices,
                            std::vector<pcl::Vertices>& new_triangles,
                            std::vector<size_t>& new_indices, HashedIndexMapping& remapping,
                            double stamp_in_sec) override {
    new_vertices.clear();
    new_triangles.clear();
    new_indices.clear();
    const detail::FlatMesh flat(mesh);
    std::vector<uint64_t> remap(flat.numVertices());
    if (flat.packed) {
      detail::throwOnError(pgmo_vcc_integrate_voxblox(
          handle_.get(), flat.numBlocks(), flat.block_idx.data(), flat.offsets.data(),
          flat.x16.data(), flat.y16.data(), flat.z16.data(), flat.r8.data(), flat.g8.data(),
          flat.b8.data(), flat.block_edge_length, stamp_in_sec, remap.data(), nullptr));
    } else {
      detail::throwOnError(pgmo_vcc_integrate(handle_.get(), flat.numBlocks(), flat.block_idx.data(),
                                              flat.offsets.data(), flat.xyz.data(), flat.rgba.data(),
                                              stamp_in_sec, remap.data(), nullptr));
    }
    last_index_.assign(remap.begin(), remap.end());
    last_kind_.clear();
    size_t b = 0;
    for (const auto& block : mesh.blockIndices()) {
      remapping.insert({block, IndexMapping()});  // :123
      auto& map = remapping.at(block);
      for (uint64_t i = flat.offsets[b]; i < flat.offsets[b + 1]; ++i) {
        map[i - flat.offsets[b]] = remap[i];  // :154 operator[] (overwrite)
      }
      ++b;
    }
    refreshActive();
  }

  void pruneStoredMesh(double earliest_time_sec) override {
    detail::throwOnError(pgmo_vcc_prune(handle_.get(), earliest_time_sec));
    refreshActive();
  }

  void clearArchivedBlocks(const BlockIndices& blocks) override {
    std::vector<int32_t> idx;
    for (const auto& b : blocks) {
      idx.push_back(b.x());
      idx.push_back(b.y());
      idx.push_back(b.z());
    }
    detail::throwOnError(
        pgmo_vcc_clear_blocks(handle_.get(), static_cast<uint32_t>(blocks.size()), idx.data()));
    refreshActive();
  }

  std::vector<size_t> getInvalidIndices() const override {
    uint64_t sizes[4];
    detail::throwOnError(pgmo_vcc_sizes(handle_.get(), sizes));
    std::vector<uint64_t> inv(sizes[3]);
    detail::throwOnError(pgmo_vcc_snapshot(handle_.get(), nullptr, nullptr, nullptr, nullptr,
                                           nullptr, inv.data()));
    return std::vector<size_t>(inv.begin(), inv.end());
  }

  void getVertices(PointCloud::Ptr vertices) override {
    uint64_t sizes[4];
    detail::throwOnError(pgmo_vcc_sizes(handle_.get(), sizes));
    std::vector<float> xyz(3 * sizes[0]);
    std::vector<uint32_t> rgba(sizes[0]);
    detail::throwOnError(pgmo_vcc_snapshot(handle_.get(), xyz.data(), rgba.data(), nullptr,
                                           nullptr, nullptr, nullptr));
    vertices->clear();
    for (uint64_t i = 0; i < sizes[0]; ++i) {
      vertices->push_back(detail::makePoint(xyz.data() + 3 * i, rgba[i]));
    }
  }

  void getTimestamps(std::shared_ptr<std::vector<Timestamp>> vertex_stamps) override {
    uint64_t sizes[4];
    detail::throwOnError(pgmo_vcc_sizes(handle_.get(), sizes));
    vertex_stamps->assign(sizes[0], 0);
    detail::throwOnError(pgmo_vcc_snapshot(handle_.get(), nullptr, nullptr, vertex_stamps->data(),
                                           nullptr, nullptr, nullptr));
  }

  void getStoredPolygons(std::shared_ptr<std::vector<pcl::Vertices>> polygons) override {
    uint64_t sizes[4];
    detail::throwOnError(pgmo_vcc_sizes(handle_.get(), sizes));
    std::vector<uint32_t> flat(3 * sizes[1]);
    detail::throwOnError(pgmo_vcc_snapshot(handle_.get(), nullptr, nullptr, nullptr, flat.data(),
                                           nullptr, nullptr));
    detail::fillPolygons(flat, polygons.get());
  }

  DeviceView deviceView() const override {
    DeviceView v;
    uint64_t n = 0;
    detail::throwOnError(pgmo_vcc_device_view(handle_.get(), &v.vertices_xyzw, &v.stamps_ns, &n));
    v.num_vertices = n;
    return v;
  }
  size_t getNumVertices() const override {
    uint64_t sizes[4];
    detail::throwOnError(pgmo_vcc_sizes(handle_.get(), sizes));
    return sizes[0];
  }

  const std::vector<size_t>& getActiveVerticesIndex() const override {
    return active_vertices_index_;
  }

 protected:
  void refreshActive() {
    uint64_t sizes[4];
    detail::throwOnError(pgmo_vcc_sizes(handle_.get(), sizes));
    std::vector<uint64_t> act(sizes[2]);
    detail::throwOnError(pgmo_vcc_snapshot(handle_.get(), nullptr, nullptr, nullptr, nullptr,
                                           act.data(), nullptr));
    active_vertices_index_.assign(act.begin(), act.end());
  }

  std::shared_ptr<pgmo_vcc_t> handle_;
  std::vector<size_t> active_vertices_index_;
};
using VoxelClearingCompressionPtr = std::shared_ptr<VoxelClearingCompression>;

class OctreeCompression : public MeshCompression {
 public:
  explicit OctreeCompression(double resolution, int device = 0) : MeshCompression(resolution) {
    pgmo_oct_t* h = nullptr;
    detail::throwOnError(pgmo_oct_create(resolution, device, &h));
    handle_.reset(h, [](pgmo_oct_t* p) { pgmo_oct_destroy(p); });
  }

 protected:
  // BlockCompression: same MeshCompression base flow, cell-hash structure tests
  struct BlockTag {};
  OctreeCompression(double resolution, int device, BlockTag) : MeshCompression(resolution) {
    pgmo_oct_t* h = nullptr;
    detail::throwOnError(pgmo_oct_create_block(resolution, device, &h));
    handle_.reset(h, [](pgmo_oct_t* p) { pgmo_oct_destroy(p); });
  }

 public:

  // mesh_compression.cpp:216-417.  The dense per-vertex record from the GPU is replayed onto the
  // caller's persistent `remapping` with the reference's exact container operations:
  // reobservations overwrite (:288), new / converged vertices insert-if-absent (:366-372),
  // potential vertices that failed the face check get no entry.
  void compressAndIntegrate(const MeshInterface& mesh, PointCloud& new_vertices,
                            std::vector<pcl::Vertices>& new_triangles,
                            std::vector<size_t>& new_indices, HashedIndexMapping& remapping,
                            double stamp_in_sec) override {
    new_vertices.clear();
    new_triangles.clear();
    new_indices.clear();
    const detail::FlatMesh flat(mesh);
    std::vector<int64_t> idx(flat.numVertices());
    std::vector<uint8_t> kind(flat.numVertices());
    if (flat.packed) {
      detail::throwOnError(pgmo_oct_integrate_voxblox(
          handle_.get(), flat.numBlocks(), flat.block_idx.data(), flat.offsets.data(),
          flat.x16.data(), flat.y16.data(), flat.z16.data(), flat.r8.data(), flat.g8.data(),
          flat.b8.data(), flat.block_edge_length, stamp_in_sec, idx.data(), kind.data()));
    } else {
      detail::throwOnError(pgmo_oct_integrate(handle_.get(), flat.numBlocks(), flat.block_idx.data(),
                                              flat.offsets.data(), flat.xyz.data(), flat.rgba.data(),
                                              stamp_in_sec, idx.data(), kind.data()));
    }
    last_index_ = idx;
    last_kind_ = kind;
    size_t b = 0;
    for (const auto& block : mesh.blockIndices()) {
      remapping.insert({block, IndexMapping()});  // :256
      ++b;
    }
    // pass 1: reobservations, in input order; pass 2: inserts.  (Overwrites and inserts touch
    // different keys within one message, so the relative order between the passes is free.)
    b = 0;
    for (const auto& block : mesh.blockIndices()) {
      auto& map = remapping.at(block);
      for (uint64_t i = flat.offsets[b]; i < flat.offsets[b + 1]; ++i) {
        if (kind[i] == 1) map[i - flat.offsets[b]] = static_cast<size_t>(idx[i]);
      }
      ++b;
    }
    b = 0;
    for (const auto& block : mesh.blockIndices()) {
      auto& map = remapping.at(block);
      for (uint64_t i = flat.offsets[b]; i < flat.offsets[b + 1]; ++i) {
        if (kind[i] == 2) map.insert({i - flat.offsets[b], static_cast<size_t>(idx[i])});
      }
      ++b;
    }
    uint64_t sizes[8];
    detail::throwOnError(pgmo_oct_sizes(handle_.get(), sizes));
    std::vector<float> nv(3 * sizes[4]);
    std::vector<uint32_t> nc(sizes[4]), nt(3 * sizes[5]);
    std::vector<uint64_t> ni(sizes[6]);
    detail::throwOnError(pgmo_oct_last_outputs(handle_.get(), nv.data(), nc.data(), nt.data(), ni.data()));
    for (uint64_t i = 0; i < sizes[4]; ++i) new_vertices.push_back(detail::makePoint(nv.data() + 3 * i, nc[i]));
    detail::fillPolygons(nt, &new_triangles);
    new_indices.assign(ni.begin(), ni.end());
    refreshActive();
  }

  void pruneStoredMesh(double earliest_time_sec) override {
    detail::throwOnError(pgmo_oct_prune(handle_.get(), earliest_time_sec));
    refreshActive();
  }

  void getVertices(PointCloud::Ptr vertices) override {
    uint64_t sizes[8];
    detail::throwOnError(pgmo_oct_sizes(handle_.get(), sizes));
    std::vector<float> xyz(3 * sizes[0]);
    std::vector<uint32_t> rgba(sizes[0]);
    detail::throwOnError(pgmo_oct_snapshot(handle_.get(), xyz.data(), rgba.data(), nullptr, nullptr,
                                           nullptr, nullptr, nullptr));
    vertices->clear();
    for (uint64_t i = 0; i < sizes[0]; ++i) {
      vertices->push_back(detail::makePoint(xyz.data() + 3 * i, rgba[i]));
    }
  }
  void getTimestamps(std::shared_ptr<std::vector<Timestamp>> vertex_stamps) override {
    uint64_t sizes[8];
    detail::throwOnError(pgmo_oct_sizes(handle_.get(), sizes));
    vertex_stamps->assign(sizes[0], 0);
    detail::throwOnError(pgmo_oct_snapshot(handle_.get(), nullptr, nullptr, vertex_stamps->data(),
                                           nullptr, nullptr, nullptr, nullptr));
  }
  void getStoredPolygons(std::shared_ptr<std::vector<pcl::Vertices>> polygons) override {
    uint64_t sizes[8];
    detail::throwOnError(pgmo_oct_sizes(handle_.get(), sizes));
    std::vector<uint32_t> flat(3 * sizes[1]);
    detail::throwOnError(pgmo_oct_snapshot(handle_.get(), nullptr, nullptr, nullptr, flat.data(),
                                           nullptr, nullptr, nullptr));
    detail::fillPolygons(flat, polygons.get());
  }
  void getActiveVertices(PointCloudXYZ::Ptr vertices) {
    uint64_t sizes[8];
    detail::throwOnError(pgmo_oct_sizes(handle_.get(), sizes));
    std::vector<float> xyz(3 * sizes[2]);
    detail::throwOnError(pgmo_oct_snapshot(handle_.get(), nullptr, nullptr, nullptr, nullptr,
                                           nullptr, xyz.data(), nullptr));
    vertices->clear();
    for (uint64_t i = 0; i < sizes[2]; ++i) {
      vertices->push_back(pcl::PointXYZ(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    }
  }
  void getActiveVerticesTimestamps(std::shared_ptr<std::vector<double>> timestamps) {
    uint64_t sizes[8];
    detail::throwOnError(pgmo_oct_sizes(handle_.get(), sizes));
    timestamps->assign(sizes[2], 0.0);
    detail::throwOnError(pgmo_oct_snapshot(handle_.get(), nullptr, nullptr, nullptr, nullptr,
                                           nullptr, nullptr, timestamps->data()));
  }
  DeviceView deviceView() const override {
    DeviceView v;
    uint64_t n = 0;
    detail::throwOnError(pgmo_oct_device_view(handle_.get(), &v.vertices_xyzw, &v.stamps_ns, &n));
    v.num_vertices = n;
    return v;
  }
  size_t getNumVertices() const override {
    uint64_t sizes[8];
    detail::throwOnError(pgmo_oct_sizes(handle_.get(), sizes));
    return sizes[0];
  }
  const std::vector<size_t>& getActiveVerticesIndex() const override {
    return active_vertices_index_;
  }

 protected:
  void refreshActive() {
    uint64_t sizes[8];
    detail::throwOnError(pgmo_oct_sizes(handle_.get(), sizes));
    std::vector<uint64_t> act(sizes[2]);
    detail::throwOnError(pgmo_oct_snapshot(handle_.get(), nullptr, nullptr, nullptr, nullptr,
                                           act.data(), nullptr, nullptr));
    active_vertices_index_.assign(act.begin(), act.end());
  }

  std::shared_ptr<pgmo_oct_t> handle_;
  std::vector<size_t> active_vertices_index_;
};
using OctreeCompressionPtr = std::shared_ptr<OctreeCompression>;

// compression/block_compression.h:14-58: identical interface and base flow; a vertex is a
// re-observation iff its cell floor(p / resolution) already holds an active vertex.
class BlockCompression : public OctreeCompression {
 public:
  explicit BlockCompression(double resolution, int device = 0)
      : OctreeCompression(resolution, device, BlockTag()) {}
};
using BlockCompressionPtr = std::shared_ptr<BlockCompression>;

}  // namespace kimera_pgmo_b200
