// ORACLE — TEST INFRASTRUCTURE ONLY (see pcl_octree_restated.h header).
//
// CPU restatement of Kimera-PGMO's mesh-frontend compression path:
//   kimera_pgmo/include/kimera_pgmo/compression/mesh_compression.h:25-148
//   kimera_pgmo/src/compression/mesh_compression.cpp:216-417 (compressAndIntegrate), :419-464 (prune)
//   kimera_pgmo/src/compression/octree_compression.cpp:17-81, include/.../octree_compression.h:14-66
//   kimera_pgmo/src/compression/voxel_clearing_compression.cpp:17-210, include/.../voxel_clearing_compression.h
//   kimera_pgmo/src/compression/block_compression.cpp (cell-hash variant, SURVEY §8f "next" #1)
//   kimera_pgmo/src/utils/common_functions.cpp:166-185 (SurfaceExists)
//   kimera_pgmo/include/kimera_pgmo/utils/mesh_interface.h:17-37, pcl_mesh_interface.h:17-45
// spatial_hash (MIT-SPARK/Spatial-Hash, unpinned git dependency, absent here) supplies only
// containers on this path: Index/LongIndex (Eigen int / int64 3-vectors) and IndexHash
// (x + y*17191 + z*17191^2 truncated to unsigned int, the voxblox AnyIndexHash it derives
// from).  Iteration order of the unordered containers (which fixes the order of `polygons_`
// and of `empty_slots_`) follows from that hash + libstdc++; it is "parity unpinned"
// (SURVEY.md §8c iv) beyond what test_voxel_clearing_compression.cpp:59-104 checks.
#pragma once

#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "deform_oracle.h"
#include "pcl_octree_restated.h"

namespace pgmo_oracle {

struct PointXYZRGBA {
  float x = 0.f, y = 0.f, z = 0.f;
  uint32_t rgba = 0;
};

struct BlockIndex {
  int x = 0, y = 0, z = 0;
  bool operator==(const BlockIndex& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct LongIndex {
  int64_t x = 0, y = 0, z = 0;
  bool operator==(const LongIndex& o) const { return x == o.x && y == o.y && z == o.z; }
};

struct IndexHash {
  static constexpr size_t sl = 17191;
  static constexpr size_t sl2 = sl * sl;
  size_t operator()(const BlockIndex& i) const {
    return static_cast<unsigned int>(i.x + i.y * sl + i.z * sl2);
  }
  size_t operator()(const LongIndex& i) const {
    return static_cast<unsigned int>(i.x + i.y * sl + i.z * sl2);
  }
};

using BlockIndices = std::vector<BlockIndex>;
template <typename V>
using BlockIndexMap = std::unordered_map<BlockIndex, V, IndexHash>;
using LongIndexSet = std::unordered_set<LongIndex, IndexHash>;
template <typename V>
using LongIndexHashMap = std::unordered_map<LongIndex, V, IndexHash>;
using IndexMapping = std::unordered_map<size_t, size_t>;
using HashedIndexMapping = BlockIndexMap<IndexMapping>;
using Face = std::array<uint32_t, 3>;  // pcl::Vertices::vertices restricted to triangles

// utils/mesh_interface.h:17-37 — cursor-style block accessor (markBlockActive mutates a cursor)
struct MeshInterface {
  virtual ~MeshInterface() = default;
  virtual const BlockIndices& blockIndices() const = 0;
  virtual void markBlockActive(const BlockIndex& block) const = 0;
  virtual size_t activeBlockSize() const = 0;
  virtual PointXYZRGBA getActiveVertex(size_t i) const = 0;
  // mesh_interface.h:28-33 (kNoLabel stands for std::nullopt)
  static constexpr uint32_t kNoLabel = 0xFFFFFFFFu;
  virtual bool hasSemantics() const { return false; }
  virtual uint32_t getActiveSemantics(size_t /* i */) const { return kNoLabel; }
};

// Flat block mesh: block b owns vertices [offsets[b], offsets[b+1]).  Covers both the
// reference's PclMeshInterface (one block (0,0,0) of 3*faces unindexed vertices,
// pcl_mesh_interface.h:17-45) and block-structured inputs (test fixture
// OrderedBlockMeshInterface, pgmo_fixtures.h:28-48; VoxbloxMsgInterface).
struct FlatBlockMesh : MeshInterface {
  BlockIndices blocks;
  std::vector<size_t> offsets;  // blocks.size() + 1
  std::vector<PointXYZRGBA> vertices;
  std::vector<uint32_t> labels;  // empty: no semantics
  mutable size_t active = 0;
  const BlockIndices& blockIndices() const override { return blocks; }
  void markBlockActive(const BlockIndex& block) const override {
    for (size_t b = 0; b < blocks.size(); ++b) {
      if (blocks[b] == block) active = b;  // last duplicate wins, as with a block map
    }
  }
  size_t activeBlockSize() const override { return offsets[active + 1] - offsets[active]; }
  PointXYZRGBA getActiveVertex(size_t i) const override { return vertices[offsets[active] + i]; }
  bool hasSemantics() const override { return !labels.empty(); }
  uint32_t getActiveSemantics(size_t i) const override { return labels.at(offsets[active] + i); }
};

// common_functions.cpp:166-185
bool SurfaceExists(const Face& new_surface,
                   const std::map<size_t, std::vector<size_t>>& adjacent_surfaces,
                   const std::vector<Face>& surfaces);

// Dense per-input-vertex record of what compressAndIntegrate did to `remapping` this call.
// kind: 0 = no write, 1 = overwrite (operator[]) , 2 = insert-if-absent (.insert).
struct RemapRecord {
  std::vector<int64_t> index;
  std::vector<uint8_t> kind;
};

class MeshCompression {
 public:
  explicit MeshCompression(double resolution) : resolution_(resolution) {}
  virtual ~MeshCompression() = default;

  virtual void compressAndIntegrate(const MeshInterface& mesh,
                                    std::vector<PointXYZRGBA>& new_vertices,
                                    std::vector<Face>& new_triangles,
                                    std::vector<size_t>& new_indices,
                                    HashedIndexMapping& remapping,
                                    double stamp_in_sec,
                                    RemapRecord* record = nullptr);
  virtual void pruneStoredMesh(double earliest_time_sec);

  virtual void reInitializeStructure(const std::vector<PointXYZ>& active_vertices) = 0;
  virtual bool checkIfVertexUnique(const PointXYZ& v, int* matched_ind) const = 0;
  virtual void updateStructure(const std::vector<PointXYZ>& vertices) = 0;
  virtual bool checkIfVertexTempUnique(const PointXYZ& v, int* matched_ind) const = 0;
  virtual void initializeTempStructure(const std::vector<PointXYZ>& vertices) = 0;
  virtual void updateTempStructure(const std::vector<PointXYZ>& vertices) = 0;
  virtual std::vector<size_t> getInvalidIndices() const { return {}; }
  virtual void clearArchivedBlocks(const BlockIndices&) {}

  const std::vector<PointXYZRGBA>& getVertices() const { return all_vertices_; }
  const std::vector<Timestamp>& getTimestamps() const { return all_vertex_stamps_; }
  const std::vector<PointXYZ>& getActiveVertices() const { return active_vertices_xyz_; }
  const std::vector<Face>& getStoredPolygons() const { return polygons_; }
  const std::vector<double>& getActiveVerticesTimestamps() const { return active_vertex_stamps_; }
  size_t getNumVertices() const { return all_vertices_.size(); }
  const std::vector<size_t>& getActiveVerticesIndex() const { return active_vertices_index_; }

 protected:
  double resolution_;
  std::vector<PointXYZ> active_vertices_xyz_;
  std::vector<PointXYZRGBA> all_vertices_;
  std::vector<Timestamp> all_vertex_stamps_;
  std::vector<size_t> active_vertices_index_;
  std::vector<Face> polygons_;
  std::map<size_t, std::vector<size_t>> adjacent_polygons_;
  std::vector<double> active_vertex_stamps_;
};

class OctreeCompression : public MeshCompression {
 public:
  explicit OctreeCompression(double resolution);
  void reInitializeStructure(const std::vector<PointXYZ>& active_vertices) override;
  bool checkIfVertexUnique(const PointXYZ& v, int* matched_ind) const override;
  void updateStructure(const std::vector<PointXYZ>& vertices) override;
  bool checkIfVertexTempUnique(const PointXYZ& v, int* matched_ind) const override;
  void initializeTempStructure(const std::vector<PointXYZ>& vertices) override;
  void updateTempStructure(const std::vector<PointXYZ>& vertices) override;

 protected:
  std::unique_ptr<OctreeSearchRestated> octree_;
  std::unique_ptr<OctreeSearchRestated> temp_octree_;
};

// compression/block_compression.{h,cpp}: uniqueness by cell hash (SURVEY.md 8f "next" #1).
// spatial_hash::Grid<LongIndex>::toIndex (MIT-SPARK/Spatial-Hash, unpinned, absent here) is
// restated as floor(p * (1 / voxel_size)) per axis in float, the voxblox getGridIndexFromPoint
// family it derives from without the epsilon; PARITY UNPINNED beyond what
// test_block_compression.cpp checks (integer-aligned vertices).
class BlockCompression : public MeshCompression {
 public:
  explicit BlockCompression(double resolution)
      : MeshCompression(resolution), voxel_size_inv_(1.0f / static_cast<float>(resolution)) {}
  void reInitializeStructure(const std::vector<PointXYZ>& active_vertices) override;
  bool checkIfVertexUnique(const PointXYZ& v, int* matched_ind) const override;
  void updateStructure(const std::vector<PointXYZ>& vertices) override;
  bool checkIfVertexTempUnique(const PointXYZ& v, int* matched_ind) const override;
  void initializeTempStructure(const std::vector<PointXYZ>& vertices) override;
  void updateTempStructure(const std::vector<PointXYZ>& vertices) override;
  LongIndex toIndex(const PointXYZ& p) const;

 protected:
  float voxel_size_inv_;
  LongIndexHashMap<size_t> cell_hash_;
  LongIndexHashMap<size_t> temp_cell_hash_;
};

class VoxelClearingCompression : public MeshCompression {
 public:
  explicit VoxelClearingCompression(double resolution) : MeshCompression(resolution) {}

  void compressAndIntegrate(const MeshInterface& mesh,
                            std::vector<PointXYZRGBA>& new_vertices,
                            std::vector<Face>& new_triangles,
                            std::vector<size_t>& new_indices,
                            HashedIndexMapping& remapping,
                            double stamp_in_sec,
                            RemapRecord* record = nullptr) override;
  void pruneStoredMesh(double earliest_time_sec) override;
  void clearArchivedBlocks(const BlockIndices& blocks) override;
  std::vector<size_t> getInvalidIndices() const override { return empty_slots_; }

  void reInitializeStructure(const std::vector<PointXYZ>&) override { throw std::logic_error("not implemented"); }
  bool checkIfVertexUnique(const PointXYZ&, int*) const override { throw std::logic_error("not implemented"); }
  void updateStructure(const std::vector<PointXYZ>&) override { throw std::logic_error("not implemented"); }
  bool checkIfVertexTempUnique(const PointXYZ&, int*) const override { throw std::logic_error("not implemented"); }
  void initializeTempStructure(const std::vector<PointXYZ>&) override { throw std::logic_error("not implemented"); }
  void updateTempStructure(const std::vector<PointXYZ>&) override { throw std::logic_error("not implemented"); }

  // per-input-vertex voxel keys of the last call (test aid; keys are part of the parity bar)
  const std::vector<LongIndex>& lastKeys() const { return last_keys_; }

 protected:
  void pruneMeshBlocks(const BlockIndices& to_clear);
  void updateRemapping(const MeshInterface& mesh, double stamp_in_sec,
                       HashedIndexMapping& remapping, RemapRecord* record);
  void updateVertices();
  void updateActiveIndices();

  BlockIndexMap<LongIndexSet> prev_meshes_;
  BlockIndexMap<std::vector<size_t>> block_face_map_;
  BlockIndexMap<double> block_update_times_;
  LongIndexHashMap<size_t> vertices_map_;
  std::map<size_t, size_t> indices_to_active_refs_;
  std::map<size_t, size_t> indices_to_inactive_refs_;
  std::vector<size_t> empty_slots_;
  size_t max_index_ = 0;
  size_t archived_polygon_size_ = 0;
  std::vector<LongIndex> last_keys_;
};

}  // namespace pgmo_oracle
