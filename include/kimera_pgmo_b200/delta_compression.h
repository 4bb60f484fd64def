// Drop-in for kimera_pgmo::DeltaCompression and the part of kimera_pgmo::MeshDelta it fills:
//   include/kimera_pgmo/compression/delta_compression.h:60-149, src/compression/delta_compression.cpp
//   include/kimera_pgmo/mesh_delta.h:24-150
// Same class / member names and meaning; the work happens in libpgmo_b200 (pgmo_delta_*).
// Iteration order of the reference's unordered containers is implementation-defined; here it is
// insertion order (vertices of a delta are ordered "first observed first"), see DESIGN.md 4.4.
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include <numeric>
#include <optional>
#include <stdexcept>
#include <string>

#include "kimera_pgmo_b200/compression.h"

namespace kimera_pgmo_b200 {

struct Face {  // mesh_delta.h:24-37, mesh_delta.cpp:66-77
  Face() = default;
  Face(size_t a, size_t b, size_t c)
      : v1(static_cast<uint32_t>(a)), v2(static_cast<uint32_t>(b)), v3(static_cast<uint32_t>(c)) {}
  Face(const std::vector<size_t>& indices, size_t i)
      : Face(indices.at(i), indices.at(i + 1), indices.at(i + 2)) {}
  bool valid() const { return v1 != v2 && v1 != v3 && v2 != v3; }
  void fill(std::vector<uint32_t>& other) const { other = {v1, v2, v3}; }
  uint32_t v1 = 0, v2 = 0, v3 = 0;
};

class MeshDelta {  // mesh_delta.h:65-150: what DeltaCompression::update produces
 public:
  using Ptr = std::shared_ptr<MeshDelta>;
  MeshDelta() : MeshDelta(0, 0) {}  // mesh_delta.cpp:79
  MeshDelta(size_t vertex_start_, size_t face_start_)
      : vertex_start(vertex_start_), face_start(face_start_),
        vertex_updates(new pcl::PointCloud<pcl::PointXYZRGBA>()) {}
  // a whole mesh as one delta (mesh_delta.cpp:86-106)
  MeshDelta(const pcl::PointCloud<pcl::PointXYZRGBA>& vertices, const std::vector<Timestamp>& stamps,
            const std::vector<pcl::Vertices>& faces,
            std::optional<std::vector<uint32_t>> semantics = std::nullopt)
      : vertex_updates(new pcl::PointCloud<pcl::PointXYZRGBA>(vertices)), stamp_updates(stamps) {
    if (semantics) semantic_updates = *semantics;
    face_updates.reserve(faces.size());
    for (const auto& face : faces) {
      face_updates.emplace_back(face.vertices[0], face.vertices[1], face.vertices[2]);
    }
  }

  // mesh_delta.cpp:144-168: what DeltaCompression::update calls per vertex / face in the
  // reference (here the device fills the vectors); kept for callers that build deltas by hand
  size_t addVertex(Timestamp timestamp_ns, const pcl::PointXYZRGBA& point,
                   std::optional<uint32_t> semantics = std::nullopt, bool archive = false) {
    const size_t index = vertex_start + vertex_updates->size();
    vertex_updates->push_back(point);
    stamp_updates.push_back(timestamp_ns);
    if (semantics) semantic_updates.push_back(*semantics);
    if (archive) ++num_archived_vertices_;
    return index;
  }
  void addFace(const Face& face, bool archive = false) {
    (archive ? face_archive_updates : face_updates).push_back(face);
  }

  size_t getNumArchivedVertices() const { return num_archived_vertices_; }
  size_t getNumArchivedFaces() const { return face_archive_updates.size(); }
  size_t getTotalArchivedVertices() const { return vertex_start + num_archived_vertices_; }
  size_t getTotalArchivedFaces() const { return face_start + face_archive_updates.size(); }
  size_t getTotalVertices() const { return vertex_start + vertex_updates->size(); }
  size_t getTotalFaces() const { return face_start + face_archive_updates.size() + face_updates.size(); }
  size_t getLocalIndex(size_t index) const { return index - vertex_start; }
  size_t getGlobalIndex(size_t index) const { return index + vertex_start; }
  // mesh_delta.cpp:112-114
  bool hasSemantics() const { return semantic_updates.size() == vertex_updates->size(); }

  // updateVertices / updateFaces (mesh_delta.h:171-215, mesh_delta.cpp:116-142) for the PCL
  // containers; stamps / semantics optional as in the reference (semantics are written only
  // when the delta has a label for every vertex)
  void updateVertices(pcl::PointCloud<pcl::PointXYZRGBA>& vertices, std::vector<Timestamp>* stamps,
                      std::vector<uint32_t>* semantics = nullptr) const {
    const bool semantics_valid = semantics && hasSemantics();
    const size_t total_vertices = vertex_start + vertex_updates->size();
    vertices.resize(total_vertices);
    if (stamps) stamps->resize(total_vertices);
    if (semantics_valid) semantics->resize(total_vertices);
    for (size_t i = 0; i < vertex_updates->size(); ++i) {
      vertices.at(i + vertex_start) = vertex_updates->at(i);
      if (stamps) stamps->at(i + vertex_start) = stamp_updates.at(i);
      if (semantics_valid) semantics->at(i + vertex_start) = semantic_updates.at(i);
    }
  }
  void updateFaces(std::vector<pcl::Vertices>& faces) const {
    faces.resize(getTotalFaces());
    size_t face_idx = face_start;
    for (const auto* list : {&face_archive_updates, &face_updates}) {
      for (const auto& face : *list) {
        faces.at(face_idx).vertices = {face.v1, face.v2, face.v3};
        ++face_idx;
      }
    }
  }
  void updateMesh(pcl::PointCloud<pcl::PointXYZRGBA>& vertices, std::vector<Timestamp>& stamps,
                  std::vector<pcl::Vertices>& faces, std::vector<uint32_t>* semantics = nullptr) const {
    updateVertices(vertices, &stamps, semantics);
    updateFaces(faces);
  }
  // mesh_delta.cpp:194-200: the vertices of this delta that were not archived
  std::shared_ptr<std::vector<int>> getActiveIndices() const {
    auto indices = std::make_shared<std::vector<int>>(vertex_updates->size() - num_archived_vertices_);
    std::iota(indices->begin(), indices->end(), static_cast<int>(vertex_start + num_archived_vertices_));
    return indices;
  }
  // mesh_delta.cpp:202-262
  void checkFaces(const std::string& name) const {
    const size_t total_archived = getTotalArchivedVertices();
    for (const auto& face : face_archive_updates) {
      if (!(face.v1 < total_archived && face.v2 < total_archived && face.v3 < total_archived)) {
        throw std::runtime_error(name + " invalid archived face: (" + std::to_string(face.v1) + ", " +
                                 std::to_string(face.v2) + ", " + std::to_string(face.v3) +
                                 ") (total archived: " + std::to_string(total_archived) + ")");
      }
    }
    const size_t total_vertices = vertex_start + vertex_updates->size();
    for (const auto& face : face_updates) {
      if (!(face.v1 < total_vertices && face.v2 < total_vertices && face.v3 < total_vertices)) {
        throw std::runtime_error(name + " invalid face: (" + std::to_string(face.v1) + ", " +
                                 std::to_string(face.v2) + ", " + std::to_string(face.v3) +
                                 ") (total vertices: " + std::to_string(total_vertices) + ")");
      }
    }
  }
  void validate(pcl::PointCloud<pcl::PointXYZRGBA>& vertices, std::vector<Timestamp>& stamps,
                std::vector<pcl::Vertices>& faces, std::vector<uint32_t>* semantics = nullptr) const {
    auto fail = [](const char* what, size_t have, const char* start_name, size_t start) {
      throw std::runtime_error(std::string("invalid update: ") + what + " size (" + std::to_string(have) +
                               ") < " + start_name + " (" + std::to_string(start) + ")");
    };
    if (vertices.size() < vertex_start) fail("vertices", vertices.size(), "vertex start", vertex_start);
    if (stamps.size() < vertex_start) fail("stamps", stamps.size(), "vertex start", vertex_start);
    if (semantics && semantics->size() < vertex_start) fail("semantics", semantics->size(), "vertex start", vertex_start);
    if (faces.size() < face_start) fail("faces", faces.size(), "faces start", face_start);
    checkFaces("[validate]");
  }

  size_t vertex_start = 0;
  size_t face_start = 0;
  pcl::PointCloud<pcl::PointXYZRGBA>::Ptr vertex_updates;
  std::vector<Timestamp> stamp_updates;
  std::vector<uint32_t> semantic_updates;
  std::vector<Face> face_updates;
  std::vector<Face> face_archive_updates;
  std::map<size_t, size_t> prev_to_curr;
  std::set<size_t> deleted_indices;
  std::set<size_t> observed_indices;
  std::set<size_t> new_indices;

 protected:
  friend class DeltaCompression;
  size_t num_archived_vertices_ = 0;
};

// what a BlockFilter sees of a block (delta_compression.h:49-58; the voxel set and the index
// list stay on the device)
struct BlockInfo {
  uint64_t update_time_ns = 0;
  uint16_t sequence_number = 0;
};

class DeltaCompression {
 public:
  using Ptr = std::shared_ptr<DeltaCompression>;
  using BlockFilter = std::function<bool(const BlockIndex&, const BlockInfo&)>;

  explicit DeltaCompression(double resolution, int device = 0) {
    pgmo_delta_t* h = nullptr;
    detail::throwOnError(pgmo_delta_create(resolution, device, &h));
    handle_.reset(h, [](pgmo_delta_t* p) { pgmo_delta_destroy(p); });
  }
  virtual ~DeltaCompression() = default;

  // delta_compression.cpp:211-259
  MeshDelta::Ptr update(MeshInterface& mesh, uint64_t timestamp_ns,
                        HashedIndexMapping* remapping = nullptr) {
    // any MeshInterface (voxblox messages included) through getActiveVertex /
    // getActiveSemantics, like the reference (:278-288)
    const detail::FlatMesh flat(mesh, /* allow_packed */ false, /* with_semantics */ true);
    std::vector<uint64_t> remap(flat.numVertices());
    detail::throwOnError(pgmo_delta_update_labeled(
        handle_.get(), flat.numBlocks(), flat.block_idx.data(), flat.offsets.data(), flat.xyz.data(),
        flat.rgba.data(), flat.labels.empty() ? nullptr : flat.labels.data(), timestamp_ns, remap.data()));
    if (remapping) {  // :171-174, :186-192: insert-if-absent into the caller's maps
      size_t b = 0;
      for (const auto& block : mesh.blockIndices()) {
        auto& map = remapping->insert({block, IndexMapping()}).first->second;
        for (uint64_t i = flat.offsets[b]; i < flat.offsets[b + 1]; ++i) {
          map.insert({i - flat.offsets[b], static_cast<size_t>(remap[i])});
        }
        ++b;
      }
    }
    return fetch();
  }

  void archiveBlocksByTime(uint64_t earliest_time_ns) {  // :306-310
    detail::throwOnError(pgmo_delta_archive_by_time(handle_.get(), earliest_time_ns));
  }

  void clearArchivedBlocks(const BlockIndices& blocks) {  // :312-316
    std::vector<int32_t> idx;
    for (const auto& b : blocks) {
      idx.push_back(b.x());
      idx.push_back(b.y());
      idx.push_back(b.z());
    }
    detail::throwOnError(pgmo_delta_archive_blocks(handle_.get(), static_cast<uint32_t>(blocks.size()), idx.data()));
  }

  void archiveBlocks(const BlockFilter& should_archive) {  // :318-360
    uint64_t n = 0;
    detail::throwOnError(pgmo_delta_blocks(handle_.get(), 0, nullptr, nullptr, nullptr, &n));
    std::vector<int32_t> idx(3 * n);
    std::vector<uint64_t> times(n);
    std::vector<uint32_t> seqs(n);
    detail::throwOnError(pgmo_delta_blocks(handle_.get(), n, idx.data(), times.data(), seqs.data(), &n));
    std::vector<int32_t> chosen;
    for (uint64_t i = 0; i < n; ++i) {
      BlockInfo info;
      info.update_time_ns = times[i];
      info.sequence_number = static_cast<uint16_t>(seqs[i]);
      if (should_archive(BlockIndex(idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]), info)) {
        chosen.insert(chosen.end(), idx.begin() + 3 * i, idx.begin() + 3 * i + 3);
      }
    }
    detail::throwOnError(pgmo_delta_archive_blocks(handle_.get(), static_cast<uint32_t>(chosen.size() / 3),
                                                   chosen.data()));
  }

 protected:
  MeshDelta::Ptr fetch() {
    uint64_t s[12];
    detail::throwOnError(pgmo_delta_sizes(handle_.get(), s));
    auto delta = std::make_shared<MeshDelta>(s[0], s[1]);
    std::vector<float> xyz(3 * s[2]);
    std::vector<uint32_t> rgba(s[2]), faces(3 * s[4]), faces_arch(3 * s[5]);
    std::vector<uint64_t> stamps(s[2]), p2c(2 * s[6]), deleted(s[7]), observed(s[8]), fresh(s[9]);
    uint64_t n_p2c = 0, n_del = 0, n_obs = 0, n_new = 0;
    detail::throwOnError(pgmo_delta_fetch(handle_.get(), xyz.data(), rgba.data(), stamps.data(), faces.data(),
                                          faces_arch.data(), p2c.data(), &n_p2c, deleted.data(), &n_del,
                                          observed.data(), &n_obs, fresh.data(), &n_new));
    for (uint64_t i = 0; i < s[2]; ++i) {
      delta->vertex_updates->push_back(detail::makePoint(xyz.data() + 3 * i, rgba[i]));
      delta->stamp_updates.push_back(stamps[i]);
    }
    int any_label = 0;
    detail::throwOnError(pgmo_delta_fetch_labels(handle_.get(), nullptr, &any_label));
    if (any_label && s[2]) {  // addVertex pushes a label only where there is one (mesh_delta.cpp:152-154)
      std::vector<uint32_t> labels(s[2]);
      detail::throwOnError(pgmo_delta_fetch_labels(handle_.get(), labels.data(), nullptr));
      for (uint32_t label : labels) {
        if (label != 0xFFFFFFFFu) delta->semantic_updates.push_back(label);
      }
    }
    delta->num_archived_vertices_ = s[3];
    for (uint64_t f = 0; f < s[4]; ++f) delta->face_updates.emplace_back(faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]);
    for (uint64_t f = 0; f < s[5]; ++f) {
      delta->face_archive_updates.emplace_back(faces_arch[3 * f], faces_arch[3 * f + 1], faces_arch[3 * f + 2]);
    }
    for (uint64_t i = 0; i < n_p2c; ++i) delta->prev_to_curr.emplace_hint(delta->prev_to_curr.end(), p2c[2 * i], p2c[2 * i + 1]);
    delta->deleted_indices.insert(deleted.begin(), deleted.begin() + n_del);
    delta->observed_indices.insert(observed.begin(), observed.begin() + n_obs);
    delta->new_indices.insert(fresh.begin(), fresh.begin() + n_new);
    return delta;
  }

  std::shared_ptr<pgmo_delta_t> handle_;
};

}  // namespace kimera_pgmo_b200
