// ORACLE — TEST INFRASTRUCTURE ONLY (see pcl_octree_restated.h header).
//
// CPU restatement of Kimera-PGMO's mesh deformation path:
//   kimera_pgmo/include/kimera_pgmo/mesh_deformation.h:22-213   (SearchTree, deformPoints x2)
//   kimera_pgmo/src/mesh_deformation.cpp:22-106                 (SearchTree::Impl, interpPoint)
//   kimera_pgmo/include/kimera_pgmo/deformation_graph.h:715-870 (start index, cache, predeform)
//   kimera_pgmo/src/deformation_graph.cpp:523-581               (deformMesh wrappers)
//   kimera_pgmo/src/utils/common_structs.cpp:21-24              (stampFromSec)
// GTSAM (Pose3/Rot3/Values/Symbol) is absent from /root/reference and this image
// (CMakeLists.txt:15, unpinned 4.x): `values` is restated as a dense pose table
// (row-major R then t) plus a validity byte per control index, which is exactly what
// values.exists(Symbol(prefix, j)) / values.at<Pose3>(Symbol(prefix, j)) expose to this path.
//
// Arithmetic conventions defined by this oracle where the reference's are not
// establishable offline: R*p is accumulated row-wise left to right,
// ((r0*p0 + r1*p1) + r2*p2), no FMA contraction anywhere (build with -ffp-contract=off).
#pragma once

#include <cstdint>
#include <map>
#include <set>
#include <vector>

#include "pcl_octree_restated.h"

namespace pgmo_oracle {

using Timestamp = uint64_t;

struct Point3d {
  double x = 0, y = 0, z = 0;
};

struct Pose3 {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
  double t[3] = {0, 0, 0};
};

// common_structs.cpp:21-24: duration_cast<nanoseconds>(duration<double>(sec)).count()
inline Timestamp stampFromSec(double sec) {
  return static_cast<Timestamp>(static_cast<int64_t>(sec * 1e9));
}

// gtsam "values" as seen by this path for one prefix.
struct ValuesView {
  const uint8_t* valid = nullptr;  // valid[j] != 0 <=> values.exists(Symbol(prefix, j))
  const Pose3* poses = nullptr;
  size_t size = 0;
  bool exists(size_t j) const { return j < size && valid[j]; }
};

// mesh_deformation.h:22-42 + mesh_deformation.cpp:22-70
class SearchTree {
 public:
  explicit SearchTree(double resolution = 1.0) : tree_(resolution) { tree_.setInputCloud(&cloud_); }
  size_t getLeafCount() const { return tree_.getLeafCount(); }
  void addPoint(const Point3d& p, bool valid) {
    PointXYZ f;
    f.x = static_cast<float>(p.x);
    f.y = static_cast<float>(p.y);
    f.z = static_cast<float>(p.z);
    cloud_.push_back(f);
    if (valid) tree_.addPointFromCloud(static_cast<int>(cloud_.size() - 1));
  }
  void removePoint(size_t index) { tree_.deleteVoxelAtPoint(static_cast<int>(index)); }
  void search(const PointXYZ& p, size_t k, std::vector<int>& nn_index,
              std::vector<float>& nn_sq_dist) const {
    tree_.nearestKSearch(p, static_cast<uint32_t>(k + 1), nn_index, nn_sq_dist);
  }
  const OctreeSearchRestated& tree() const { return tree_; }

 private:
  std::vector<PointXYZ> cloud_;
  OctreeSearchRestated tree_;
};

struct InterpDetail {
  std::vector<int> order;  // neighbour indices in blend order (n-1 entries)
  double pre_cast[3];      // blended point before the final cast<float>()
};

// mesh_deformation.cpp:73-106
PointXYZ interpPoint(std::set<size_t>& control_points_seen,
                     const std::vector<Point3d>& control_points,
                     const ValuesView& values,
                     const SearchTree& tree,
                     size_t k,
                     const PointXYZ& old_point,
                     InterpDetail* detail = nullptr);

// Flat SoA view of a (stamped) point cloud; out may alias in (deformation_graph.cpp:545).
struct DeformIO {
  size_t num_vertices = 0;
  const float* in_x = nullptr;
  const float* in_y = nullptr;
  const float* in_z = nullptr;
  const Timestamp* stamps = nullptr;  // nullptr selects the unstamped overload
  float* out_x = nullptr;
  float* out_y = nullptr;
  float* out_z = nullptr;
};

struct DeformTrace {
  // one entry per *processed* vertex, in processing order (== control_point_map)
  std::vector<size_t> vertex;            // which vertex index (ii)
  std::vector<size_t> position;          // processing position (point_index)
  std::vector<std::vector<int>> order;   // neighbour indices in blend order
  std::vector<double> pre_cast;          // 3 doubles per processed vertex
};

// Return value: 0 = ran; 1 = reference returned early without touching outputs.
// mesh_deformation.h:63-113 (unstamped) and :129-213 (stamped).
int deformPoints(const DeformIO& io,
                 std::vector<std::set<size_t>>& control_point_map,
                 const std::vector<Point3d>& control_points,
                 const std::vector<Timestamp>& control_point_stamps,
                 const ValuesView& values,
                 size_t k,
                 double tol_t,
                 const std::vector<size_t>* indices,
                 DeformTrace* trace = nullptr);

// deformation_graph.h:715-870 / deformation_graph.cpp:523-581 — the part of DeformationGraph
// that sits on the hot path (no solver): per-prefix control table, cache of the last deformed
// vertices, start-index search, rigid predeform.
class DeformationGraphMeshPath {
 public:
  void setControlPoints(char prefix, const std::vector<Point3d>& positions,
                        const std::vector<Timestamp>& stamps) {
    vertex_positions_[prefix] = positions;
    vertex_stamps_[prefix] = stamps;
  }
  void setRecalculate(bool v) { recalculate_vertices_ = v; }
  bool recalculate() const { return recalculate_vertices_; }

  // deformation_graph.h:808-870.  graph_indices may be null; returns the start index used.
  size_t deformPoints(const DeformIO& io, char prefix, const ValuesView& values, size_t k,
                      double tol_t, const std::vector<int>* graph_indices,
                      int start_index_hint,
                      std::vector<std::set<size_t>>* vertex_graph_map,
                      DeformTrace* trace = nullptr);

  size_t findStartIndex(char prefix, int start_index_hint, const DeformIO& io, double tol_t) const;

 private:
  std::map<char, std::vector<Point3d>> vertex_positions_;
  std::map<char, std::vector<Timestamp>> vertex_stamps_;
  bool recalculate_vertices_ = false;
  std::map<char, std::vector<PointXYZ>> last_calculated_vertices_;
};

}  // namespace pgmo_oracle
