// Drop-in for the mesh path of kimera_pgmo::DeformationGraph:
//   deformMesh x2            src/deformation_graph.cpp:523-559
//   deformPoints (XYZRGBA)   src/deformation_graph.cpp:561-581
//   deformPoints<In,Out>     include/kimera_pgmo/deformation_graph.h:808-870
//   findStartIndex / fillPreviousPoints / cacheNewPoints / predeformPoints   :715-806
// Everything that belongs to the solver (factors, optimize(), values_) stays upstream: this class
// only holds what the mesh path reads — vertex_positions_, vertex_stamps_, the cache of the
// last deformed vertices and the recalculate flag — and forwards the heavy lifting to the GPU.
//
// Integration: DeformationGraph keeps a DeformationGraphMeshPath member, mirrors
// addNewMeshNode() into addControlPoint(), sets setRecalculateVertices() where it sets
// recalculate_vertices_, and its deformMesh/deformPoints bodies become one-line forwards
// (INTEGRATION.md).
#pragma once

#include <algorithm>
#include <cstring>
#include <map>
#include <numeric>

#include "kimera_pgmo_b200/mesh_deformation.h"

namespace kimera_pgmo_b200 {

class DeformationGraphMeshPath {
 public:
  // n_devices: 1 = `device` only (default); 0 = every visible device; n = devices 0 .. n - 1.
  // With several devices a full deformMesh is sharded by vertex range over all of them
  // (deformation::Backend); results are bit-identical.
  explicit DeformationGraphMeshPath(int device = 0, int n_devices = 1)
      : device_(device), n_devices_(n_devices) {}

  // DeformationGraph::addNewMeshNode (deformation_graph.cpp:362-378) appends in index order
  void addControlPoint(char prefix, const gtsam::Point3& position, Timestamp stamp) {
    vertex_positions_[prefix].push_back(position);
    vertex_stamps_[prefix].push_back(stamp);
  }
  void setControlPoints(char prefix, const std::vector<gtsam::Point3>& positions,
                        const std::vector<Timestamp>& stamps) {
    vertex_positions_[prefix] = positions;
    vertex_stamps_[prefix] = stamps;
  }
  void setRecalculateVertices(bool v = true) { recalculate_vertices_ = v; }
  bool recalculateVertices() const { return recalculate_vertices_; }
  size_t getNumVertices(char prefix) const {
    const auto it = vertex_positions_.find(prefix);
    return it == vertex_positions_.end() ? 0 : it->second.size();
  }

  // deformation_graph.cpp:532-559
  pcl::PolygonMesh deformMesh(const pcl::PolygonMesh& original_mesh,
                              const std::vector<Timestamp>& stamps,
                              const std::vector<int>& graph_indices, const char& prefix,
                              const gtsam::Values& optimized_values, size_t k = 4,
                              double tol_t = 10.0) {
    pcl::PolygonMesh new_mesh;
    detail::BlobCloud in;
    detail::Trace trace("deformMesh");
    if (detail::blobView(original_mesh.cloud, &in)) {
      // The blob already has the layout toPCLPointCloud2 writes for a PointXYZRGBA cloud, so
      // blob -> cloud -> blob (:540, :557) would change nothing but x / y / z: copy the blob
      // once (threads) and deform the copy in place through the traits layer.
      detail::blobCloneShape(original_mesh.cloud, new_mesh.cloud);
      trace.mark("blob allocate");
      detail::BlobCloud out;
      detail::blobView(new_mesh.cloud, &out);
      detail::parallelFor(in.n, [&](size_t begin, size_t end) {
        std::memcpy(out.base + begin * out.step, in.base + begin * in.step, (end - begin) * in.step);
      });
      trace.mark("blob copy");
      in.stamps = &stamps;
      out.stamps = &stamps;
      deformPoints(out, in, prefix, optimized_values, k, tol_t, &graph_indices);
      trace.mark("deformPoints");
    } else {
      pcl::PointCloud<pcl::PointXYZRGBA> new_vertices;
      pcl::fromPCLPointCloud2(original_mesh.cloud, new_vertices);
      deformPoints(new_vertices, new_vertices, stamps, prefix, optimized_values, k, tol_t,
                   &graph_indices);
      pcl::toPCLPointCloud2(new_vertices, new_mesh.cloud);
    }
    // new_mesh.polygons = original_mesh.polygons (:555), one heap block per face: on a few
    // threads for large meshes
    detail::resizePrefaulted(new_mesh.polygons, original_mesh.polygons.size());
    detail::parallelFor(original_mesh.polygons.size(), [&](size_t begin, size_t end) {
      for (size_t f = begin; f < end; f++) new_mesh.polygons[f] = original_mesh.polygons[f];
    }, size_t(1) << 15);
    trace.mark("polygons");
    return new_mesh;
  }

  // Extension (no reference counterpart): deformMesh's full recompute (start index 0) on a
  // cloud that is already in HBM — 16-byte {x, y, z, rgba} records + u64 stamps, e.g. a
  // compressor's MeshCompression::deviceView() — written to d_out_xyzw (device, same layout,
  // may alias the input).  Same arithmetic and results as deformMesh; nothing is cached
  // (last_calculated_vertices_ is a host-side cache).  Returns false where the reference
  // returns early (unknown prefix / fewer control points than k).
  bool deformMeshResident(const void* d_vertices_xyzw, const uint64_t* d_stamps_ns,
                          size_t num_vertices, char prefix, const gtsam::Values& optimized_values,
                          void* d_out_xyzw, size_t k = 4, double tol_t = 10.0,
                          void* cuda_stream = nullptr) {
    if (vertex_positions_.find(prefix) == vertex_positions_.end()) return false;
    deformation::Backend& b = backend(prefix);
    b.setControls(prefix, vertex_positions_.at(prefix), &vertex_stamps_.at(prefix),
                  optimized_values);
    const int rc = pgmo_deform_mesh_device(b.get(), num_vertices, d_vertices_xyzw, d_stamps_ns,
                                           static_cast<uint32_t>(k), tol_t, PGMO_PRECISION_FP64,
                                           d_out_xyzw, nullptr, nullptr, cuda_stream);
    detail::throwOnError(rc);
    return rc == PGMO_OK;
  }

  // deformation_graph.cpp:561-581
  void deformPoints(pcl::PointCloud<pcl::PointXYZRGBA>& vertices,
                    const pcl::PointCloud<pcl::PointXYZRGBA>& old_vertices,
                    const std::vector<Timestamp>& stamps, char prefix,
                    const gtsam::Values& optimized_values, size_t k = 4, double tol_t = 10.0,
                    const std::vector<int>* graph_indices = nullptr, int start_index_hint = -1,
                    std::vector<std::set<size_t>>* vertex_graph_map = nullptr) {
    deformPoints(vertices, ConstStampedCloud<pcl::PointXYZRGBA>{old_vertices, stamps}, prefix,
                 optimized_values, k, tol_t, graph_indices, start_index_hint, vertex_graph_map);
  }

  // deformation_graph.h:808-870
  template <typename CloudIn, typename CloudOut>
  void deformPoints(CloudOut& vertices, const CloudIn& old_vertices, char prefix,
                    const gtsam::Values& optimized_values, size_t k = 4, double tol_t = 10.0,
                    const std::vector<int>* graph_indices = nullptr, int start_index_hint = -1,
                    std::vector<std::set<size_t>>* vertex_graph_map = nullptr) {
    if (vertex_positions_.find(prefix) == vertex_positions_.end()) return;  // :819-823

    const size_t start_idx = findStartIndex(prefix, start_index_hint, old_vertices, tol_t);
    fillPreviousPoints(vertices, prefix, start_idx);

    deformation::Backend& b = backend(prefix);
    std::vector<size_t> to_deform;
    bool controls_uploaded = false;
    if (start_idx != 0) {
      if (graph_indices) {
        b.setControls(prefix, vertex_positions_.at(prefix), &vertex_stamps_.at(prefix),
                      optimized_values);
        controls_uploaded = true;
        predeformPoints(b, vertices, old_vertices, *graph_indices, to_deform, start_idx);
      } else {
        to_deform.resize(traits::num_vertices(vertices) - start_idx);
        std::iota(to_deform.begin(), to_deform.end(), start_idx);
      }
    }

    const std::vector<size_t>* indices_ptr = start_idx == 0 ? nullptr : &to_deform;
    std::vector<std::set<size_t>> vertex_graph_map_deformed;
    // the table predeformPoints needed is the one deformation::deformPoints is about to upload
    // (for a stamped cloud; an unstamped one uploads it without the stamps)
    const deformation::Backend::KeepControls keep(
        b, controls_uploaded && traits::has_get_stamp<CloudIn>::value);
    if (vertex_graph_map) {
      deformation::deformPoints(vertices, vertex_graph_map_deformed, old_vertices, prefix,
                                vertex_positions_.at(prefix), vertex_stamps_.at(prefix),
                                optimized_values, k, tol_t, indices_ptr, &b);
    } else {
      // the reference fills a local map nobody reads (:848-855); deformMesh comes this way
      deformation::deformPointsOnly(vertices, old_vertices, prefix, vertex_positions_.at(prefix),
                                    vertex_stamps_.at(prefix), optimized_values, k, tol_t,
                                    indices_ptr, &b);
    }

    if (vertex_graph_map) {  // :857-866
      if (start_idx == 0) {
        *vertex_graph_map = vertex_graph_map_deformed;
      } else {
        vertex_graph_map->resize(traits::num_vertices(vertices));
        for (size_t i = 0; i < indices_ptr->size(); i++) {
          vertex_graph_map->at(indices_ptr->at(i)) = vertex_graph_map_deformed.at(i);
        }
      }
    }

    detail::Trace trace("deformPoints");
    cacheNewPoints(vertices, prefix, start_idx);
    trace.mark("cacheNewPoints");
    recalculate_vertices_ = false;
  }

  // deformation_graph.h:715-746
  template <typename Cloud>
  size_t findStartIndex(char prefix, int start_index_hint, const Cloud& cloud,
                        double tol_t) const {
    const auto cache = last_calculated_vertices_.find(prefix);
    if (cache == last_calculated_vertices_.end()) return 0;
    if (recalculate_vertices_) return 0;
    if (start_index_hint >= 0) return static_cast<size_t>(start_index_hint);
    // unsigned arithmetic like the reference: the subtraction wraps for small stamps
    const Timestamp min_stamp = std::max(static_cast<Timestamp>(0),
                                         vertex_stamps_.at(prefix).back() - stampFromSec(tol_t));
    size_t lo = 0, count = traits::num_vertices(cloud);
    while (count > 0) {  // std::upper_bound(gen, min_stamp, v < stamp(i))
      const size_t step = count / 2;
      if (!(min_stamp < traits::get_timestamp(cloud, lo + step))) {
        lo += step + 1;
        count -= step + 1;
      } else {
        count = step;
      }
    }
    return std::min(lo, cache->second.size());
  }

 private:
  deformation::Backend& backend(char prefix) {
    auto it = backends_.find(prefix);
    if (it == backends_.end()) {
      it = backends_.emplace(prefix, deformation::Backend(device_, n_devices_)).first;
    }
    return it->second;
  }

  // :748-760
  template <typename Cloud>
  void fillPreviousPoints(Cloud& vertices, char prefix, size_t start_idx) const {
    const auto iter = last_calculated_vertices_.find(prefix);
    if (iter == last_calculated_vertices_.end()) return;
    const auto& cache = iter->second;
    detail::parallelFor(start_idx, [&](size_t begin, size_t end) {
      for (size_t i = begin; i < end; i++) {
        traits::set_vertex(vertices, i, traits::get_vertex(cache, i));
      }
    });
  }

  // :762-777
  template <typename Cloud>
  void cacheNewPoints(const Cloud& vertices, char prefix, size_t start_idx) {
    auto& cache = last_calculated_vertices_[prefix];
    const size_t total_vertices = traits::num_vertices(vertices);
    cache.resize(total_vertices);
    if (start_idx >= total_vertices) return;
    detail::parallelFor(total_vertices - start_idx, [&](size_t begin, size_t end) {
      for (size_t i = start_idx + begin; i < start_idx + end; i++) {
        traits::set_vertex(cache, i, traits::get_vertex(vertices, i));
      }
    });
  }

  // :779-806 on the GPU (pgmo_deform_predeform)
  template <typename CloudIn, typename CloudOut>
  void predeformPoints(deformation::Backend& b, CloudOut& new_vertices, const CloudIn& vertices,
                       const std::vector<int>& graph_indices,
                       std::vector<size_t>& indices_to_deform, size_t start_idx) {
    const size_t n_v = traits::num_vertices(vertices);
    if (start_idx >= n_v) return;
    if (graph_indices.size() < n_v) throw std::out_of_range("graph_indices.at()");
    // only [start_idx, n_v) is read or written (:786): that tail is what gets staged and shipped
    static_assert(sizeof(int) == sizeof(int32_t), "graph_indices is passed as int32");
    const size_t n = n_v - start_idx;
    detail::PinnedBlock& stage = b.staging();
    stage.reserve(6 * detail::PinnedBlock::padded(n * sizeof(float)) +
                  detail::PinnedBlock::padded(n * sizeof(uint64_t)) + 256);
    stage.rewind();
    float* x = stage.take<float>(n);
    float* y = stage.take<float>(n);
    float* z = stage.take<float>(n);
    float* ox = stage.take<float>(n);
    float* oy = stage.take<float>(n);
    float* oz = stage.take<float>(n);
    uint64_t* to_deform = stage.take<uint64_t>(n);
    detail::parallelFor(n, [&](size_t begin, size_t end) {
      for (size_t i = begin; i < end; ++i) {
        const traits::Pos p = traits::get_vertex(vertices, start_idx + i);
        x[i] = p.x();
        y[i] = p.y();
        z[i] = p.z();
      }
    });
    uint64_t n_to = 0;
    detail::throwOnError(pgmo_deform_predeform(
        b.get(), n, x, y, z, reinterpret_cast<const int32_t*>(graph_indices.data()) + start_idx, 0,
        ox, oy, oz, to_deform, &n_to));
    indices_to_deform.reserve(indices_to_deform.size() + n_to);
    size_t next = 0;
    for (size_t i = 0; i < n; ++i) {
      if (next < n_to && to_deform[next] == i) {
        indices_to_deform.push_back(start_idx + i);
        ++next;
      } else {
        traits::set_vertex(new_vertices, start_idx + i, traits::Pos(ox[i], oy[i], oz[i]));
      }
    }
  }

  int device_;
  int n_devices_;
  std::map<char, std::vector<gtsam::Point3>> vertex_positions_;
  std::map<char, std::vector<Timestamp>> vertex_stamps_;
  bool recalculate_vertices_ = false;
  std::map<char, pcl::PointCloud<pcl::PointXYZ>> last_calculated_vertices_;
  std::map<char, deformation::Backend> backends_;
};

}  // namespace kimera_pgmo_b200
