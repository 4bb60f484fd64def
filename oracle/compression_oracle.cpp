// ORACLE — TEST INFRASTRUCTURE ONLY (see compression_oracle.h header for the reference map).
#include "compression_oracle.h"

#include <algorithm>
#include <cassert>
#include <cmath>

namespace pgmo_oracle {

// common_functions.cpp:166-185 — exact ordered-triple match among the faces adjacent to
// vertices[0] only.
bool SurfaceExists(const Face& new_surface,
                   const std::map<size_t, std::vector<size_t>>& adjacent_surfaces,
                   const std::vector<Face>& surfaces) {
  const size_t idx0 = new_surface[0];
  const auto it = adjacent_surfaces.find(idx0);
  if (it == adjacent_surfaces.end()) return false;
  for (const auto& s_idx : it->second) {
    if (surfaces.at(s_idx) == new_surface) return true;
  }
  return false;
}

// mesh_compression.cpp:216-417
void MeshCompression::compressAndIntegrate(const MeshInterface& mesh,
                                           std::vector<PointXYZRGBA>& new_vertices,
                                           std::vector<Face>& new_triangles,
                                           std::vector<size_t>& new_indices,
                                           HashedIndexMapping& remapping,
                                           double stamp_in_sec,
                                           RemapRecord* record) {
  new_vertices.clear();
  new_triangles.clear();
  new_indices.clear();
  if (record) {
    record->index.clear();
    record->kind.clear();
  }

  const size_t num_original_vertices = all_vertices_.size();

  std::unordered_map<size_t, size_t> reindex;
  std::vector<size_t> temp_reindex;
  std::vector<size_t> potential_new_vertices;
  std::vector<bool> potential_new_vertices_check;
  std::vector<Face> input_surfaces;  // holds *count* triples (input numbering)

  std::vector<PointXYZ> temp_new_vertices;
  initializeTempStructure(temp_new_vertices);

  size_t count = 0;
  std::unordered_map<size_t, std::pair<BlockIndex, size_t>> count_to_block;
  std::vector<PointXYZRGBA> all_parsed_points;
  std::unordered_map<size_t, std::vector<size_t>> converged_vertices;

  for (const auto& block_index : mesh.blockIndices()) {
    mesh.markBlockActive(block_index);
    remapping.insert({block_index, IndexMapping()});  // :256

    for (size_t i = 0; i < mesh.activeBlockSize(); ++i) {
      const PointXYZRGBA p = mesh.getActiveVertex(i);
      PointXYZ p_xyz;
      p_xyz.x = p.x;
      p_xyz.y = p.y;
      p_xyz.z = p.z;
      count_to_block[count] = {block_index, i};
      all_parsed_points.push_back(p);
      if (record) {
        record->index.push_back(-1);
        record->kind.push_back(0);
      }

      int result_idx = -1;
      if (checkIfVertexUnique(p_xyz, &result_idx)) {  // :268
        if (checkIfVertexTempUnique(p_xyz, &result_idx)) {
          temp_new_vertices.push_back(p_xyz);
          updateTempStructure(temp_new_vertices);
          potential_new_vertices.push_back(count);
          potential_new_vertices_check.push_back(false);
          temp_reindex.push_back(num_original_vertices + temp_new_vertices.size() - 1);
          converged_vertices.insert({count, std::vector<size_t>()});
        } else {
          temp_reindex.push_back(num_original_vertices + result_idx);
          converged_vertices[potential_new_vertices[result_idx]].push_back(count);
        }
      } else {
        // reobservation :284-301
        const size_t g = active_vertices_index_[result_idx];
        reindex[count] = g;
        remapping.at(block_index)[i] = g;  // overwrite
        if (record) {
          record->index[count] = static_cast<int64_t>(g);
          record->kind[count] = 1;
        }
        temp_reindex.push_back(g);
        if (!std::binary_search(new_indices.begin(), new_indices.end(), g)) {
          new_indices.insert(std::lower_bound(new_indices.begin(), new_indices.end(), g), g);
        }
        active_vertex_stamps_[result_idx] = stamp_in_sec;
      }

      // :304-343 every third vertex closes a face
      if (i % 3 == 2 && temp_reindex.at(count) != temp_reindex.at(count - 1)) {
        const size_t r_idx_0 = temp_reindex.at(count - 2);
        const size_t r_idx_1 = temp_reindex.at(count - 1);
        const size_t r_idx_2 = temp_reindex.at(count);
        const bool has_new_vertex =
            (r_idx_0 >= num_original_vertices || r_idx_1 >= num_original_vertices ||
             r_idx_2 >= num_original_vertices);
        if (!has_new_vertex) {
          count++;
          continue;
        }
        const bool degenerate = (r_idx_0 == r_idx_1 || r_idx_0 == r_idx_2 || r_idx_1 == r_idx_2);
        if (degenerate) {
          count++;
          continue;
        }
        input_surfaces.push_back({static_cast<uint32_t>(count - 2),
                                  static_cast<uint32_t>(count - 1),
                                  static_cast<uint32_t>(count)});
        if (r_idx_0 >= num_original_vertices)
          potential_new_vertices_check[r_idx_0 - num_original_vertices] = true;
        if (r_idx_1 >= num_original_vertices)
          potential_new_vertices_check[r_idx_1 - num_original_vertices] = true;
        if (r_idx_2 >= num_original_vertices)
          potential_new_vertices_check[r_idx_2 - num_original_vertices] = true;
      }
      count++;
    }
  }

  // :349-378 accept potential vertices that passed the face check, in creation order
  for (size_t i = 0; i < potential_new_vertices.size(); i++) {
    if (!potential_new_vertices_check[i]) continue;
    const size_t input_idx = potential_new_vertices[i];
    const PointXYZRGBA p = all_parsed_points[input_idx];
    PointXYZ p_xyz;
    p_xyz.x = p.x;
    p_xyz.y = p.y;
    p_xyz.z = p.z;
    active_vertices_xyz_.push_back(p_xyz);
    updateStructure(active_vertices_xyz_);
    all_vertices_.push_back(p);
    all_vertex_stamps_.push_back(stampFromSec(stamp_in_sec));
    active_vertices_index_.push_back(all_vertices_.size() - 1);
    active_vertex_stamps_.push_back(stamp_in_sec);
    const size_t g = all_vertices_.size() - 1;
    reindex[input_idx] = g;
    remapping.at(count_to_block[input_idx].first).insert({count_to_block[input_idx].second, g});
    if (record) {
      record->index[input_idx] = static_cast<int64_t>(g);
      record->kind[input_idx] = 2;
    }
    for (const auto& m : converged_vertices[input_idx]) {
      reindex[m] = g;
      remapping.at(count_to_block[m].first).insert({count_to_block[m].second, g});
      if (record) {
        record->index[m] = static_cast<int64_t>(g);
        record->kind[m] = 2;
      }
    }
    new_indices.push_back(g);
    new_vertices.push_back(p);
  }

  // :382-416 second pass over the candidate faces
  for (const auto& s : input_surfaces) {
    std::vector<uint32_t> reindex_v;
    bool new_surface = false;
    for (size_t idx : s) {
      if (reindex.find(idx) == reindex.end()) break;
      reindex_v.push_back(static_cast<uint32_t>(reindex[idx]));
      if (reindex[idx] >= num_original_vertices) new_surface = true;
    }
    if (reindex_v.size() < 3) continue;
    const Face reindex_s = {reindex_v[0], reindex_v[1], reindex_v[2]};
    if (!new_surface) new_surface = !SurfaceExists(reindex_s, adjacent_polygons_, polygons_);
    if (!new_surface) continue;
    if (reindex_s[0] == reindex_s[1] || reindex_s[1] == reindex_s[2] ||
        reindex_s[2] == reindex_s[0])
      continue;
    polygons_.push_back(reindex_s);
    new_triangles.push_back(reindex_s);
    for (size_t v : reindex_s) adjacent_polygons_[v].push_back(polygons_.size() - 1);
  }
}

// mesh_compression.cpp:419-464
void MeshCompression::pruneStoredMesh(double earliest_time_sec) {
  if (active_vertices_xyz_.size() == 0) return;
  std::vector<PointXYZ> temp_active_vertices;
  std::vector<double> temp_vertices_time;
  std::vector<size_t> temp_vertices_index;
  std::map<size_t, std::vector<size_t>> temp_adjacent_polygons;
  for (size_t i = 0; i < active_vertex_stamps_.size(); i++) {
    if (active_vertex_stamps_[i] > earliest_time_sec) {
      temp_active_vertices.push_back(active_vertices_xyz_[i]);
      temp_vertices_time.push_back(active_vertex_stamps_[i]);
      temp_vertices_index.push_back(active_vertices_index_[i]);
      temp_adjacent_polygons[active_vertices_index_[i]] =
          adjacent_polygons_[active_vertices_index_[i]];
    }
  }
  if (temp_active_vertices.size() < active_vertices_xyz_.size()) {
    active_vertices_xyz_.swap(temp_active_vertices);
    std::swap(active_vertex_stamps_, temp_vertices_time);
    std::swap(active_vertices_index_, temp_vertices_index);
    std::swap(adjacent_polygons_, temp_adjacent_polygons);
    reInitializeStructure(active_vertices_xyz_);
  }
}

// ---------------------------------------------------------------------------------------
// octree_compression.cpp:17-40 + octree_compression.h:20-27
static bool inOctree(const OctreeSearchRestated& octree, const PointXYZ& v, int* matched_ind) {
  if (!octree.getLeafCount()) return false;
  double min_x, min_y, min_z, max_x, max_y, max_z;
  octree.getBoundingBox(min_x, min_y, min_z, max_x, max_y, max_z);
  const bool in_box = (v.x >= min_x && v.x <= max_x) && (v.y >= min_y && v.y <= max_y) &&
                      (v.z >= min_z && v.z <= max_z);
  if (!in_box) return false;
  const double resolution = octree.getResolution();
  const double threshold = resolution * resolution;
  float sqr_distance = static_cast<float>(threshold + 1.0f);
  octree.approxNearestSearch(v, *matched_ind, sqr_distance);
  if (sqr_distance >= threshold) {
    *matched_ind = -1;
    return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------
// block_compression.cpp:19-81
LongIndex BlockCompression::toIndex(const PointXYZ& p) const {
  LongIndex i;
  i.x = static_cast<int64_t>(std::floor(p.x * voxel_size_inv_));
  i.y = static_cast<int64_t>(std::floor(p.y * voxel_size_inv_));
  i.z = static_cast<int64_t>(std::floor(p.z * voxel_size_inv_));
  return i;
}

void BlockCompression::reInitializeStructure(const std::vector<PointXYZ>& active_vertices) {
  cell_hash_.clear();
  size_t idx = 0;
  for (const auto& p : active_vertices) {
    cell_hash_.emplace(toIndex(p), idx);  // first point of a cell keeps it
    idx++;
  }
}

bool BlockCompression::checkIfVertexUnique(const PointXYZ& v, int* matched_ind) const {
  const auto it = cell_hash_.find(toIndex(v));
  if (it == cell_hash_.end()) return true;
  *matched_ind = static_cast<int>(it->second);
  return false;
}

void BlockCompression::updateStructure(const std::vector<PointXYZ>& vertices) {
  const size_t vertex_index = vertices.size() - 1;
  cell_hash_.emplace(toIndex(vertices[vertex_index]), vertex_index);
}

bool BlockCompression::checkIfVertexTempUnique(const PointXYZ& v, int* matched_ind) const {
  const auto it = temp_cell_hash_.find(toIndex(v));
  if (it == temp_cell_hash_.end()) return true;
  *matched_ind = static_cast<int>(it->second);
  return false;
}

void BlockCompression::initializeTempStructure(const std::vector<PointXYZ>& vertices) {
  temp_cell_hash_.clear();
  size_t idx = 0;
  for (const auto& p : vertices) {
    temp_cell_hash_.emplace(toIndex(p), idx);
    idx++;
  }
}

void BlockCompression::updateTempStructure(const std::vector<PointXYZ>& vertices) {
  const size_t vertex_index = vertices.size() - 1;
  temp_cell_hash_.emplace(toIndex(vertices[vertex_index]), vertex_index);
}

OctreeCompression::OctreeCompression(double resolution) : MeshCompression(resolution) {
  octree_.reset(new OctreeSearchRestated(resolution));
  octree_->setInputCloud(&active_vertices_xyz_);
  temp_octree_.reset(new OctreeSearchRestated(resolution));
}

void OctreeCompression::reInitializeStructure(const std::vector<PointXYZ>& active_vertices) {
  octree_.reset(new OctreeSearchRestated(resolution_));
  octree_->setInputCloud(&active_vertices);
  octree_->addPointsFromInputCloud();
}

bool OctreeCompression::checkIfVertexUnique(const PointXYZ& v, int* matched_ind) const {
  return !inOctree(*octree_, v, matched_ind);
}

void OctreeCompression::updateStructure(const std::vector<PointXYZ>& vertices) {
  octree_->addPointFromCloud(static_cast<int>(vertices.size() - 1));
}

bool OctreeCompression::checkIfVertexTempUnique(const PointXYZ& v, int* matched_ind) const {
  return !inOctree(*temp_octree_, v, matched_ind);
}

void OctreeCompression::initializeTempStructure(const std::vector<PointXYZ>& vertices) {
  temp_octree_.reset(new OctreeSearchRestated(resolution_));
  temp_octree_->setInputCloud(&vertices);
  temp_octree_->addPointsFromInputCloud();
}

void OctreeCompression::updateTempStructure(const std::vector<PointXYZ>& vertices) {
  temp_octree_->addPointFromCloud(static_cast<int>(vertices.size() - 1));
}

// ---------------------------------------------------------------------------------------
// voxel_clearing_compression.cpp:17-27
static void updatePolygons(std::vector<Face>& faces, const std::vector<size_t>& indices) {
  std::vector<uint32_t> face;
  for (const auto& point_index : indices) {
    face.push_back(static_cast<uint32_t>(point_index));
    if (face.size() == 3) {
      faces.push_back({face[0], face[1], face[2]});
      face.clear();
    }
  }
}

// :34-47
void VoxelClearingCompression::compressAndIntegrate(const MeshInterface& mesh,
                                                    std::vector<PointXYZRGBA>& new_vertices,
                                                    std::vector<Face>& new_faces,
                                                    std::vector<size_t>& new_indices,
                                                    HashedIndexMapping& remapping,
                                                    double stamp_in_sec,
                                                    RemapRecord* record) {
  new_vertices.clear();
  new_faces.clear();
  new_indices.clear();
  updateRemapping(mesh, stamp_in_sec, remapping, record);
  updateVertices();
  updateActiveIndices();
}

// :49-55
void VoxelClearingCompression::updateActiveIndices() {
  active_vertices_index_.clear();
  active_vertices_index_.reserve(indices_to_active_refs_.size());
  for (const auto& id_count_pair : indices_to_active_refs_) {
    active_vertices_index_.push_back(id_count_pair.first);
  }
}

// :57-66
void VoxelClearingCompression::pruneStoredMesh(double earliest_time_s) {
  BlockIndices to_clear;
  for (const auto& idx_time_pair : block_update_times_) {
    if (idx_time_pair.second <= earliest_time_s) to_clear.push_back(idx_time_pair.first);
  }
  pruneMeshBlocks(to_clear);
}

// :68-70
void VoxelClearingCompression::clearArchivedBlocks(const BlockIndices& blocks) {
  pruneMeshBlocks(blocks);
}

// :72-101
void VoxelClearingCompression::pruneMeshBlocks(const BlockIndices& to_clear) {
  for (const auto& idx : to_clear) {
    auto voxels = prev_meshes_[idx];
    for (const auto& voxel : voxels) {
      const size_t mesh_idx = vertices_map_[voxel];
      indices_to_active_refs_[mesh_idx]--;
      if (indices_to_active_refs_[mesh_idx] != 0) continue;
      indices_to_active_refs_.erase(mesh_idx);
      indices_to_inactive_refs_.erase(mesh_idx);
      vertices_map_.erase(voxel);
    }
    polygons_.resize(archived_polygon_size_);
    updatePolygons(polygons_, block_face_map_[idx]);
    archived_polygon_size_ = polygons_.size();
    block_face_map_.erase(idx);
    prev_meshes_.erase(idx);
    block_update_times_.erase(idx);
  }
  updateActiveIndices();
}

// :103-109
void VoxelClearingCompression::updateVertices() {
  polygons_.resize(archived_polygon_size_);
  for (const auto& idx_indices_pair : block_face_map_) {
    updatePolygons(polygons_, idx_indices_pair.second);
  }
}

// :111-210
void VoxelClearingCompression::updateRemapping(const MeshInterface& mesh, double stamp_in_sec,
                                               HashedIndexMapping& remapping,
                                               RemapRecord* record) {
  const double threshold_inv = 1.0 / resolution_;
  const Timestamp vertex_stamp = stampFromSec(stamp_in_sec);
  PointXYZRGBA fake_point;  // pcl::PointXYZRGBA(): xyz = 0, r = g = b = 0, a = 255 (PCL >= 1.10)
  fake_point.rgba = 0xFF000000u;
  last_keys_.clear();
  if (record) {
    record->index.clear();
    record->kind.clear();
  }

  for (const auto& block_index : mesh.blockIndices()) {
    block_update_times_[block_index] = stamp_in_sec;
    remapping.insert({block_index, IndexMapping()});
    mesh.markBlockActive(block_index);

    const size_t block_size = mesh.activeBlockSize();
    std::vector<size_t> face_map;
    face_map.reserve(block_size);

    LongIndexSet curr_voxels;
    for (size_t i = 0; i < block_size; ++i) {
      const PointXYZRGBA p = mesh.getActiveVertex(i);
      LongIndex vertex_index;  // :134-136  float * double -> double, round half away, -> int64
      vertex_index.x = static_cast<int64_t>(std::round(p.x * threshold_inv));
      vertex_index.y = static_cast<int64_t>(std::round(p.y * threshold_inv));
      vertex_index.z = static_cast<int64_t>(std::round(p.z * threshold_inv));
      last_keys_.push_back(vertex_index);

      size_t mesh_index;
      auto map_iter = vertices_map_.find(vertex_index);
      if (map_iter != vertices_map_.end()) {
        mesh_index = map_iter->second;
      } else {
        mesh_index = max_index_;
        ++max_index_;
        all_vertex_stamps_.push_back(vertex_stamp);
        vertices_map_[vertex_index] = mesh_index;
        indices_to_active_refs_[mesh_index] = 0;
        indices_to_inactive_refs_[mesh_index] = 0;
        all_vertices_.push_back(p);
      }

      remapping.at(block_index)[i] = mesh_index;
      if (record) {
        record->index.push_back(static_cast<int64_t>(mesh_index));
        record->kind.push_back(1);
      }
      all_vertices_[mesh_index] = p;
      face_map.push_back(mesh_index);
      if (!curr_voxels.count(vertex_index)) {
        ++indices_to_active_refs_[mesh_index];
        ++indices_to_inactive_refs_[mesh_index];
      }
      curr_voxels.insert(vertex_index);
    }

    block_face_map_[block_index] = face_map;

    auto prev_mesh = prev_meshes_.find(block_index);
    if (prev_mesh == prev_meshes_.end()) {
      prev_meshes_[block_index] = curr_voxels;
      continue;
    }

    for (const auto& prev : prev_mesh->second) {
      const size_t mesh_index = vertices_map_.at(prev);
      --indices_to_active_refs_[mesh_index];
      --indices_to_inactive_refs_[mesh_index];

      if (indices_to_inactive_refs_[mesh_index] == 0) {
        vertices_map_.erase(prev);
        indices_to_active_refs_.erase(mesh_index);
        indices_to_inactive_refs_.erase(mesh_index);
        all_vertices_[mesh_index] = fake_point;
        empty_slots_.push_back(mesh_index);
        continue;
      }
      if (indices_to_active_refs_[mesh_index] != 0) continue;
      indices_to_active_refs_.erase(mesh_index);
      indices_to_inactive_refs_.erase(mesh_index);
      vertices_map_.erase(prev);
    }
    prev_meshes_[block_index] = curr_voxels;
  }
}

}  // namespace pgmo_oracle
