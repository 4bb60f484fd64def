// Drop-in for the compressor / index-bookkeeping part of kimera_pgmo::MeshFrontendInterface
// (include/kimera_pgmo/mesh_frontend_interface.h:20-143, src/mesh_frontend_interface.cpp):
//   Config (the compression fields), update(), getVoxbloxMsgToGraphMapping(), the protected
//   members vertices_ / triangles_ / vertex_stamps_ / graph_* / msg_to_*_idx_ /
//   mesh_to_graph_idx_ as accessors.
// Not here (stays upstream): simplified_mesh_graph_ / makePoseGraph (the
// pose-graph message for the optimizer) and the CSV logs.
//
// update() runs the two compressors on their own threads like the reference (:100-110; each
// handle owns a CUDA stream, so their kernels overlap), then does updateMeshToGraphMappings
// (:197-211) on the GPU through pgmo_m2g_update: msg_to_mesh_idx_ / msg_to_graph_idx_ /
// mesh_to_graph_idx_ live in HBM and are materialised as host maps only when asked for.
#pragma once

#include <memory>
#include <stdexcept>
#include <thread>
#include <vector>

#include "kimera_pgmo_b200/compression.h"

namespace kimera_pgmo_b200 {

class MeshFrontendInterface {
 public:
  enum class FullCompressionMethod { OCTREE, BLOCK, VOXEL_CLEARING };
  enum class GraphCompressionMethod { OCTREE, BLOCK };
  struct Config {
    int robot_id = 0;
    double time_horizon = 10.0;
    bool track_mesh_graph_mapping = true;
    FullCompressionMethod full_compression_method = FullCompressionMethod::BLOCK;
    GraphCompressionMethod graph_compression_method = GraphCompressionMethod::BLOCK;
    double d_graph_resolution = 3.0;
    double mesh_resolution = 0.2;
    int device = 0;  // extension: the GPU that holds the state
  } const config;

  explicit MeshFrontendInterface(const Config& cfg)
      : config(cfg),
        vertices_(new pcl::PointCloud<pcl::PointXYZRGBA>()),
        triangles_(new std::vector<pcl::Vertices>()),
        vertex_stamps_(new std::vector<Timestamp>()),
        graph_vertices_(new pcl::PointCloud<pcl::PointXYZRGBA>()),
        graph_triangles_(new std::vector<pcl::Vertices>()),
        graph_stamps_(new std::vector<Timestamp>()),
        msg_to_graph_idx_(new HashedIndexMapping()),
        msg_to_mesh_idx_(new HashedIndexMapping()) {
    // createFullCompression / createGraphCompression (mesh_frontend_interface.cpp:44-68)
    switch (cfg.full_compression_method) {
      case FullCompressionMethod::OCTREE:
        full_mesh_compression_ = std::make_shared<OctreeCompression>(cfg.mesh_resolution, cfg.device);
        break;
      case FullCompressionMethod::BLOCK:
        full_mesh_compression_ = std::make_shared<BlockCompression>(cfg.mesh_resolution, cfg.device);
        break;
      case FullCompressionMethod::VOXEL_CLEARING:
        full_mesh_compression_ = std::make_shared<VoxelClearingCompression>(cfg.mesh_resolution, cfg.device);
        break;
    }
    switch (cfg.graph_compression_method) {
      case GraphCompressionMethod::OCTREE:
        d_graph_compression_ = std::make_shared<OctreeCompression>(cfg.d_graph_resolution, cfg.device);
        break;
      case GraphCompressionMethod::BLOCK:
        d_graph_compression_ = std::make_shared<BlockCompression>(cfg.d_graph_resolution, cfg.device);
        break;
    }
    pgmo_m2g_t* h = nullptr;
    detail::throwOnError(pgmo_m2g_create(cfg.device, &h));
    m2g_.reset(h, [](pgmo_m2g_t* p) { pgmo_m2g_destroy(p); });
  }
  virtual ~MeshFrontendInterface() = default;

  // mesh_frontend_interface.cpp:100-115
  void update(const MeshInterface& mesh, double time_in_sec) {
    // markBlockActive mutates the interface: every thread works on its own clone (:103-106)
    const MeshInterface::Ptr full_mesh = mesh.clone(), graph_mesh = mesh.clone();
    std::exception_ptr full_error, graph_error;
    std::thread full_thread([&] {
      try {
        if (full_mesh) updateFullMesh(*full_mesh, time_in_sec);
      } catch (...) {
        full_error = std::current_exception();
      }
    });
    std::thread graph_thread([&] {
      try {
        if (graph_mesh) updateGraph(*graph_mesh, time_in_sec);
      } catch (...) {
        graph_error = std::current_exception();
      }
    });
    latest_blocks_ = mesh.blockIndices();
    full_thread.join();
    graph_thread.join();
    if (full_error) std::rethrow_exception(full_error);
    if (graph_error) std::rethrow_exception(graph_error);
    if (config.track_mesh_graph_mapping) updateMeshToGraphMappings(mesh);
  }

  const HashedIndexMapping& getVoxbloxMsgToGraphMapping() const { return *msg_to_graph_idx_; }
  const HashedIndexMapping& getMsgToMeshMapping() const { return *msg_to_mesh_idx_; }

  // mesh_to_graph_idx_, read back from HBM (ascending mesh index)
  IndexMapping getMeshToGraphMapping() const {
    uint64_t n = 0;
    detail::throwOnError(pgmo_m2g_size(m2g_.get(), &n));
    std::vector<uint64_t> mesh_idx(n), graph_idx(n);
    if (n > 0) detail::throwOnError(pgmo_m2g_fetch(m2g_.get(), mesh_idx.data(), graph_idx.data()));
    IndexMapping out;
    out.reserve(n);
    for (uint64_t i = 0; i < n; ++i) out.insert({mesh_idx[i], graph_idx[i]});
    return out;
  }
  // the same as a dense device array over the mesh index (-1 = no graph vertex), e.g. for a
  // kernel that ties mesh vertices to control points
  struct DeviceMapping {
    const int64_t* mesh_to_graph = nullptr;
    size_t span = 0;
  };
  DeviceMapping meshToGraphDevice() const {
    DeviceMapping v;
    uint64_t span = 0;
    detail::throwOnError(pgmo_m2g_device_view(m2g_.get(), &v.mesh_to_graph, &span));
    v.span = span;
    return v;
  }

  const pcl::PointCloud<pcl::PointXYZRGBA>::Ptr& vertices() const { return vertices_; }
  const std::shared_ptr<std::vector<pcl::Vertices>>& triangles() const { return triangles_; }
  const std::shared_ptr<std::vector<Timestamp>>& vertexStamps() const { return vertex_stamps_; }
  const pcl::PointCloud<pcl::PointXYZRGBA>::Ptr& graphVertices() const { return graph_vertices_; }
  const std::shared_ptr<std::vector<pcl::Vertices>>& graphTriangles() const { return graph_triangles_; }
  const std::shared_ptr<std::vector<Timestamp>>& graphStamps() const { return graph_stamps_; }
  const std::vector<size_t>& activeIndices() const { return active_indices_; }
  const std::vector<size_t>& invalidIndices() const { return invalid_indices_; }
  const std::vector<size_t>& lastNewGraphIndices() const { return new_graph_indices_; }
  const std::vector<pcl::Vertices>& lastNewGraphTriangles() const { return new_graph_triangles_; }
  const MeshCompressionPtr& fullMeshCompression() const { return full_mesh_compression_; }
  const MeshCompressionPtr& graphCompression() const { return d_graph_compression_; }

 protected:
  // :118-150
  void updateFullMesh(const MeshInterface& mesh, double time_in_sec) {
    full_mesh_compression_->pruneStoredMesh(time_in_sec - config.time_horizon);
    std::vector<size_t> new_indices;
    std::vector<pcl::Vertices> new_triangles;
    pcl::PointCloud<pcl::PointXYZRGBA> new_vertices;
    full_mesh_compression_->compressAndIntegrate(mesh, new_vertices, new_triangles, new_indices,
                                                 *msg_to_mesh_idx_, time_in_sec);
    full_mesh_compression_->getVertices(vertices_);
    full_mesh_compression_->getStoredPolygons(triangles_);
    full_mesh_compression_->getTimestamps(vertex_stamps_);
    active_indices_ = full_mesh_compression_->getActiveVerticesIndex();
    invalid_indices_ = full_mesh_compression_->getInvalidIndices();
  }

  // :152-195 without the pose-graph message
  void updateGraph(const MeshInterface& mesh, double time_in_sec) {
    d_graph_compression_->pruneStoredMesh(time_in_sec - config.time_horizon);
    pcl::PointCloud<pcl::PointXYZRGBA> new_graph_vertices;
    new_graph_indices_.clear();
    new_graph_triangles_.clear();
    d_graph_compression_->compressAndIntegrate(mesh, new_graph_vertices, new_graph_triangles_,
                                               new_graph_indices_, *msg_to_graph_idx_, time_in_sec);
    d_graph_compression_->getVertices(graph_vertices_);
    d_graph_compression_->getStoredPolygons(graph_triangles_);
    d_graph_compression_->getTimestamps(graph_stamps_);
  }

  // :197-211, on the device
  void updateMeshToGraphMappings(const MeshInterface& mesh) {
    const detail::FlatMesh flat(mesh);  // block order and extents only
    const auto& mi = full_mesh_compression_->lastRemapIndex();
    const auto& mk = full_mesh_compression_->lastRemapKind();
    const auto& gi = d_graph_compression_->lastRemapIndex();
    const auto& gk = d_graph_compression_->lastRemapKind();
    if (mi.size() != flat.numVertices() || gi.size() != flat.numVertices()) {
      throw std::logic_error("MeshFrontendInterface: compressor records do not match the mesh");
    }
    detail::throwOnError(pgmo_m2g_update(m2g_.get(), flat.numBlocks(), flat.block_idx.data(),
                                         flat.offsets.data(), mi.data(), mk.empty() ? nullptr : mk.data(),
                                         gi.data(), gk.empty() ? nullptr : gk.data()));
  }

  MeshCompressionPtr full_mesh_compression_;
  MeshCompressionPtr d_graph_compression_;
  pcl::PointCloud<pcl::PointXYZRGBA>::Ptr vertices_;
  std::shared_ptr<std::vector<pcl::Vertices>> triangles_;
  std::shared_ptr<std::vector<Timestamp>> vertex_stamps_;
  pcl::PointCloud<pcl::PointXYZRGBA>::Ptr graph_vertices_;
  std::shared_ptr<std::vector<pcl::Vertices>> graph_triangles_;
  std::shared_ptr<std::vector<Timestamp>> graph_stamps_;
  std::shared_ptr<HashedIndexMapping> msg_to_graph_idx_;
  std::shared_ptr<HashedIndexMapping> msg_to_mesh_idx_;
  std::shared_ptr<pgmo_m2g_t> m2g_;  // msg_to_*_idx_ and mesh_to_graph_idx_ in HBM
  BlockIndices latest_blocks_;
  std::vector<size_t> active_indices_;
  std::vector<size_t> invalid_indices_;
  std::vector<size_t> new_graph_indices_;
  std::vector<pcl::Vertices> new_graph_triangles_;
};

}  // namespace kimera_pgmo_b200
