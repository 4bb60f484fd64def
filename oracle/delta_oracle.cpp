// ORACLE — TEST INFRASTRUCTURE ONLY.  See delta_oracle.h for scope and the container-order note.
#include "delta_oracle.h"

#include <cmath>
#include <stdexcept>

namespace pgmo_oracle {

namespace {

// delta_compression.cpp:19-33
size_t getRemappedIndex(const std::map<size_t, size_t>& remapping, size_t original) {
  const auto iter = remapping.find(original);
  return iter == remapping.end() ? original : iter->second;
}
size_t getRemappedIndex(const std::map<size_t, size_t>& remapping, size_t original,
                        size_t threshold) {
  if (original < threshold) return original;
  return getRemappedIndex(remapping, original);
}

bool canBeArchived(const DeltaFace& face, size_t archive_threshold) {  // :448-451
  return face.v1 < archive_threshold && face.v2 < archive_threshold && face.v3 < archive_threshold;
}

}  // namespace

// mesh_delta.cpp:144-169
size_t MeshDelta::addVertex(uint64_t timestamp_ns, const PointXYZRGBA& point, uint32_t label,
                            bool archive) {
  const size_t index = vertex_start + vertex_updates.size();
  vertex_updates.push_back(point);
  stamp_updates.push_back(timestamp_ns);
  if (label != MeshInterface::kNoLabel) semantic_updates.push_back(label);  // :152-154
  label_of_vertex.push_back(label);
  if (archive) num_archived_vertices++;
  return index;
}
void MeshDelta::addFace(const DeltaFace& face, bool archive) {
  if (archive) {
    face_archive_updates.push_back(face);
  } else {
    face_updates.push_back(face);
  }
}

// redundancy_checker.cpp:12-43 (the second clause really is checked twice; v3 -> v1 is not)
bool RedunancyChecker::check(const DeltaFace& face) const {
  const bool present =
      hasEdge(face.v1, face.v2) && hasEdge(face.v2, face.v3) && hasEdge(face.v2, face.v3);
  return !present;
}
void RedunancyChecker::add(const DeltaFace& face) {
  addEdge(face.v1, face.v2);
  addEdge(face.v2, face.v3);
  addEdge(face.v3, face.v1);
}
bool RedunancyChecker::hasEdge(size_t source, size_t target) const {
  const auto iter = A.find(source);
  if (iter == A.end()) return false;
  return iter->second.count(target) > 0;
}
void RedunancyChecker::addEdge(size_t source, size_t target) { A[source].insert(target); }

DeltaCompression::DeltaCompression(double resolution)  // :47-52
    : resolution_(resolution),
      index_scale_(1.0 / resolution),
      sequence_number_(1),
      num_archived_vertices_(0),
      num_archived_faces_(0) {}

// :54-92
void DeltaCompression::addPoint(const PointXYZRGBA& point, uint64_t timestamp_ns, uint32_t semantic_label,
                                std::vector<size_t>& face_map, std::vector<LongIndex>& curr_voxels,
                                std::set<std::array<int64_t, 3>>& curr_set) {
  // float * double product, std::round half away from zero, narrowed to int64 (:61-63)
  const LongIndex vertex_index{static_cast<int64_t>(std::round(point.x * index_scale_)),
                               static_cast<int64_t>(std::round(point.y * index_scale_)),
                               static_cast<int64_t>(std::round(point.z * index_scale_))};
  VertexInfo* info = vertices_map_.find(vertex_index);
  if (!info) info = &vertices_map_.insert(vertex_index, VertexInfo());
  if (info->sequence_number != sequence_number_) {
    info->timestamp_ns = timestamp_ns;
    info->point = point;
    info->label = semantic_label;  // :76
    const size_t prev_index = info->mesh_index;
    info->mesh_index = active_remapping_.size();
    active_remapping_.push_back(prev_index);  // cache previous index
    info->sequence_number = sequence_number_;
  }
  face_map.push_back(info->mesh_index);
  const std::array<int64_t, 3> key{vertex_index.x, vertex_index.y, vertex_index.z};
  if (!curr_set.count(key)) {
    ++info->active_refs;  // one observation per block
    curr_set.insert(key);
    curr_voxels.push_back(vertex_index);
  }
}

// :94-126
void DeltaCompression::removeBlockObservations(const std::vector<LongIndex>& to_remove) {
  for (const auto& prev : to_remove) {
    VertexInfo* found = vertices_map_.find(prev);
    if (!found) throw std::out_of_range("vertices_map_.at()");
    VertexInfo& info = *found;
    --info.active_refs;
    if (info.notObserved()) {
      delta_->deleted_indices.insert(info.mesh_index);
      vertices_map_.erase(prev);
      continue;
    }
    if (!info.shouldArchive()) continue;
    const size_t mesh_index = delta_->addVertex(info.timestamp_ns, info.point, info.label, true);
    delta_->prev_to_curr[info.mesh_index] = mesh_index;
    vertices_map_.erase(prev);
  }
}

// :128-154
void DeltaCompression::addActiveVertices() {
  vertices_map_.forEach([&](const LongIndex&, VertexInfo& info) {
    const size_t new_idx = delta_->addVertex(info.timestamp_ns, info.point, info.label);
    if (info.sequence_number != sequence_number_) {
      delta_->prev_to_curr[info.mesh_index] = new_idx;
    } else {
      if (!info.is_new) {
        const size_t prev_mesh_index = active_remapping_[info.mesh_index];
        delta_->prev_to_curr[prev_mesh_index] = new_idx;
      } else {
        delta_->new_indices.insert(new_idx);
        info.is_new = false;
      }
      delta_->observed_indices.insert(new_idx);
      active_remapping_[info.mesh_index] = new_idx;
    }
    info.mesh_index = new_idx;
  });
}

// :156-209
void DeltaCompression::addActiveFaces(const MeshInterface& mesh, std::vector<uint64_t>* remap_out) {
  const auto& prev_to_curr = delta_->prev_to_curr;
  RedunancyChecker checker;
  // where each block of the message starts in the flat remap output
  std::map<std::array<int, 3>, size_t> msg_offset;
  if (remap_out) {
    size_t off = 0;
    for (const auto& b : mesh.blockIndices()) {
      mesh.markBlockActive(b);
      msg_offset[{b.x, b.y, b.z}] = off;  // a repeated block index: the last one wins, like the map
      off += mesh.activeBlockSize();
    }
    remap_out->assign(off, ~0ull);
  }
  block_info_map_.forEach([&](const BlockIndex& id, DeltaBlockInfo& block_info) {
    const bool was_updated = block_info.sequence_number == sequence_number_;
    auto& indices = block_info.indices;
    for (size_t i = 0; i < indices.size(); i += 3) {
      for (size_t c = 0; c < 3; ++c) {
        indices[i + c] = was_updated ? active_remapping_[indices[i + c]] : prev_to_curr.at(indices[i + c]);
      }
      if (remap_out && was_updated) {
        const size_t off = msg_offset.at({id.x, id.y, id.z});
        // block_remap->insert({i, indices[i]}): the block's map starts empty each pass
        for (size_t c = 0; c < 3; ++c) (*remap_out)[off + i + c] = indices[i + c];
      }
      DeltaFace face;
      face.v1 = static_cast<uint32_t>(indices[i]);
      face.v2 = static_cast<uint32_t>(indices[i + 1]);
      face.v3 = static_cast<uint32_t>(indices[i + 2]);
      if (!face.valid()) continue;
      if (!checker.check(face)) continue;
      checker.add(face);
      delta_->addFace(face);
    }
  });
}

// :211-259
std::shared_ptr<MeshDelta> DeltaCompression::update(const MeshInterface& mesh, uint64_t timestamp_ns,
                                                    std::vector<uint64_t>* remap_out) {
  if (archive_delta_) {
    delta_ = archive_delta_;
    archive_delta_.reset();
  } else {
    delta_.reset(new MeshDelta(num_archived_vertices_, num_archived_faces_));
  }
  updateRemapping(mesh, timestamp_ns);
  addActiveVertices();
  addActiveFaces(mesh, remap_out);
  updateAndAddArchivedFaces();
  num_archived_vertices_ = delta_->getTotalArchivedVertices();
  num_archived_faces_ = delta_->getTotalArchivedFaces();
  ++sequence_number_;
  if (sequence_number_ == 0) sequence_number_ = 1;
  return delta_;
}

// :261-296
void DeltaCompression::updateRemapping(const MeshInterface& mesh, uint64_t stamp_ns) {
  active_remapping_.clear();
  for (const auto& block_index : mesh.blockIndices()) {
    bool is_block_new = false;
    DeltaBlockInfo* block_info = block_info_map_.find(block_index);
    if (!block_info) {
      is_block_new = true;
      block_info = &block_info_map_.insert(block_index, DeltaBlockInfo());
    }
    block_info->update_time_ns = stamp_ns;
    block_info->sequence_number = sequence_number_;
    block_info->indices.clear();
    std::vector<LongIndex> curr_voxels;
    std::set<std::array<int64_t, 3>> curr_set;
    mesh.markBlockActive(block_index);
    for (size_t i = 0; i < mesh.activeBlockSize(); ++i) {
      addPoint(mesh.getActiveVertex(i), stamp_ns,
               mesh.hasSemantics() ? mesh.getActiveSemantics(i) : MeshInterface::kNoLabel,  // :285
               block_info->indices, curr_voxels, curr_set);
    }
    if (!is_block_new) removeBlockObservations(block_info->vertices);
    block_info->vertices = curr_voxels;
  }
}

void DeltaCompression::archiveBlocksByTime(uint64_t earliest_time_ns) {
  archiveBlocks([earliest_time_ns](const BlockIndex&, const DeltaBlockInfo& info) {
    return info.update_time_ns <= earliest_time_ns;
  });
}

void DeltaCompression::clearArchivedBlocks(const BlockIndices& blocks) {
  archiveBlocks([&blocks](const BlockIndex& idx, const DeltaBlockInfo&) {
    for (const auto& b : blocks) {
      if (b == idx) return true;
    }
    return false;
  });
}

// :318-360
void DeltaCompression::archiveImpl(const std::vector<BlockIndex>& to_erase) {
  if (!archive_delta_) {
    archive_delta_.reset(new MeshDelta(num_archived_vertices_, num_archived_faces_));
  }
  for (const auto& idx : to_erase) {
    const DeltaBlockInfo& block_info = *block_info_map_.find(idx);
    for (const auto& voxel : block_info.vertices) {
      VertexInfo* info = vertices_map_.find(voxel);
      if (!info) info = &vertices_map_.insert(voxel, VertexInfo());  // operator[]
      info->needs_archive = true;
      --info->active_refs;
      if (!info->shouldArchive()) continue;
      const size_t new_index = archive_delta_->addVertex(info->timestamp_ns, info->point, info->label, true);
      archive_delta_->prev_to_curr[info->mesh_index] = new_index;
      vertices_map_.erase(voxel);
    }
  }
  RedunancyChecker checker;
  for (const auto& idx : to_erase) {
    archiveBlockFaces(*block_info_map_.find(idx), checker);
    block_info_map_.erase(idx);
  }
}

// :362-414
void DeltaCompression::archiveBlockFaces(const DeltaBlockInfo& block_info, RedunancyChecker& checker) {
  const auto archive_threshold = archive_delta_->getTotalArchivedVertices();
  const auto& prev_to_curr = archive_delta_->prev_to_curr;
  const auto& indices = block_info.indices;
  for (size_t i = 0; i + 2 < indices.size(); i += 3) {
    DeltaFace face;
    face.v1 = static_cast<uint32_t>(indices[i]);
    face.v2 = static_cast<uint32_t>(indices[i + 1]);
    face.v3 = static_cast<uint32_t>(indices[i + 2]);
    const bool v1_can_archive = prev_to_curr.count(face.v1) || face.v1 < num_archived_vertices_;
    const bool v2_can_archive = prev_to_curr.count(face.v2) || face.v2 < num_archived_vertices_;
    const bool v3_can_archive = prev_to_curr.count(face.v3) || face.v3 < num_archived_vertices_;
    if (!v1_can_archive || !v2_can_archive || !v3_can_archive) {
      archived_faces_.push_back(face);
      continue;
    }
    face.v1 = static_cast<uint32_t>(getRemappedIndex(prev_to_curr, face.v1));
    face.v2 = static_cast<uint32_t>(getRemappedIndex(prev_to_curr, face.v2));
    face.v3 = static_cast<uint32_t>(getRemappedIndex(prev_to_curr, face.v3));
    if (!face.valid()) continue;
    if (!canBeArchived(face, archive_threshold)) {
      archived_faces_.push_back(face);
      continue;
    }
    if (!checker.check(face)) continue;
    checker.add(face);
    archive_delta_->addFace(face, true);
  }
}

// :416-446
void DeltaCompression::updateAndAddArchivedFaces() {
  RedunancyChecker checker;
  const auto& prev_to_curr = delta_->prev_to_curr;
  const auto archive_threshold = delta_->getTotalArchivedVertices();
  std::vector<DeltaFace> kept;
  for (DeltaFace face : archived_faces_) {
    face.v1 = static_cast<uint32_t>(getRemappedIndex(prev_to_curr, face.v1, num_archived_vertices_));
    face.v2 = static_cast<uint32_t>(getRemappedIndex(prev_to_curr, face.v2, num_archived_vertices_));
    face.v3 = static_cast<uint32_t>(getRemappedIndex(prev_to_curr, face.v3, num_archived_vertices_));
    if (!face.valid()) continue;          // erased
    if (!checker.check(face)) continue;   // erased: a pending face already covers it
    const bool can_archive = canBeArchived(face, archive_threshold);
    checker.add(face);
    delta_->addFace(face, can_archive);
    if (!can_archive) kept.push_back(face);  // stays pending, with its remapped indices
  }
  archived_faces_.swap(kept);
}

}  // namespace pgmo_oracle
