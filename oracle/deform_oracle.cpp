// ORACLE — TEST INFRASTRUCTURE ONLY (see deform_oracle.h header for the reference map).
#include "deform_oracle.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace pgmo_oracle {

// mesh_deformation.cpp:73-106
PointXYZ interpPoint(std::set<size_t>& control_points_seen,
                     const std::vector<Point3d>& control_points,
                     const ValuesView& values,
                     const SearchTree& tree,
                     size_t k,
                     const PointXYZ& old_point,
                     InterpDetail* detail) {
  std::vector<int> nn_index;
  std::vector<float> nn_sq_dist;
  tree.search(old_point, k + 1, nn_index, nn_sq_dist);  // :83 (Impl::search adds one more, :46)

  // :85-86  std::sqrt(float) stays float, then widens
  const double d_max = std::sqrt(nn_sq_dist[nn_index.size() - 1]);
  const bool use_const_weight = std::sqrt(nn_sq_dist[0]) == d_max || d_max == 0;

  double weight_sum = 0;
  double acc[3] = {0, 0, 0};
  const double vi[3] = {static_cast<double>(old_point.x), static_cast<double>(old_point.y),
                        static_cast<double>(old_point.z)};
  for (size_t j = 0; j < nn_index.size() - 1; j++) {  // :91
    const Point3d& gj = control_points.at(nn_index[j]);
    const double w = use_const_weight ? 1 : (1 - std::sqrt(nn_sq_dist[j]) / d_max);  // :94
    weight_sum += w;
    const Pose3& T = values.poses[nn_index[j]];  // values.at<Pose3>(Symbol(prefix, idx)) :96
    const double p[3] = {vi[0] - gj.x, vi[1] - gj.y, vi[2] - gj.z};
    double delta[3];
    for (int r = 0; r < 3; ++r) {
      const double rot = (T.R[3 * r + 0] * p[0] + T.R[3 * r + 1] * p[1]) + T.R[3 * r + 2] * p[2];
      delta[r] = rot + T.t[r];  // :97-98
    }
    for (int r = 0; r < 3; ++r) acc[r] += w * delta[r];  // :100
    control_points_seen.insert(nn_index[j]);            // :101
    if (detail) detail->order.push_back(nn_index[j]);
  }
  for (int r = 0; r < 3; ++r) acc[r] /= weight_sum;  // :104
  if (detail) {
    detail->pre_cast[0] = acc[0];
    detail->pre_cast[1] = acc[1];
    detail->pre_cast[2] = acc[2];
  }
  PointXYZ out;
  out.x = static_cast<float>(acc[0]);  // :105
  out.y = static_cast<float>(acc[1]);
  out.z = static_cast<float>(acc[2]);
  return out;
}

namespace {

inline PointXYZ getVertex(const DeformIO& io, size_t i) {
  PointXYZ p;
  p.x = io.in_x[i];
  p.y = io.in_y[i];
  p.z = io.in_z[i];
  return p;
}

inline void setVertex(const DeformIO& io, size_t i, const PointXYZ& p) {
  io.out_x[i] = p.x;
  io.out_y[i] = p.y;
  io.out_z[i] = p.z;
}

void record(DeformTrace* trace, size_t ii, size_t pos, const InterpDetail& d) {
  if (!trace) return;
  trace->vertex.push_back(ii);
  trace->position.push_back(pos);
  trace->order.push_back(d.order);
  trace->pre_cast.push_back(d.pre_cast[0]);
  trace->pre_cast.push_back(d.pre_cast[1]);
  trace->pre_cast.push_back(d.pre_cast[2]);
}

// mesh_deformation.h:63-113
int deformPointsUnstamped(const DeformIO& io,
                          std::vector<std::set<size_t>>& control_point_map,
                          const std::vector<Point3d>& control_points,
                          const ValuesView& values,
                          size_t k,
                          const std::vector<size_t>* indices,
                          DeformTrace* trace) {
  const size_t num_points = indices ? indices->size() : io.num_vertices;
  if (!num_points) return 1;               // :78-80
  if (control_points.size() == 0) return 1;  // :83-86

  control_point_map.clear();  // :88
  SearchTree search_tree;     // :91
  for (size_t j = 0; j < control_points.size(); j++) {
    search_tree.addPoint(control_points[j], values.exists(j));  // :92-94
  }
  if (search_tree.getLeafCount() < k) return 1;  // :96-99

  for (size_t p_idx = 0; p_idx < num_points; ++p_idx) {
    const size_t ii = indices ? indices->at(p_idx) : p_idx;
    control_point_map.emplace_back();
    InterpDetail detail;
    const PointXYZ p_new = interpPoint(control_point_map.back(), control_points, values,
                                       search_tree, k, getVertex(io, ii), &detail);
    setVertex(io, ii, p_new);
    record(trace, ii, p_idx, detail);
  }
  return 0;
}

// mesh_deformation.h:129-213
int deformPointsStamped(const DeformIO& io,
                        std::vector<std::set<size_t>>& control_point_map,
                        const std::vector<Point3d>& control_points,
                        const std::vector<Timestamp>& control_point_stamps,
                        const ValuesView& values,
                        size_t k,
                        double tol_t,
                        const std::vector<size_t>* indices,
                        DeformTrace* trace) {
  const size_t num_points = indices ? indices->size() : io.num_vertices;
  if (!num_points) return 1;                // :143-146
  if (control_points.size() < k) return 1;  // :149-152

  control_point_map.clear();  // :154
  SearchTree search_tree;     // :155

  size_t ctrl_pt_idx = 0;        // :159
  size_t lower_ctrl_pt_idx = 0;  // :160
  const Timestamp tol = stampFromSec(tol_t);
  for (size_t point_index = 0; point_index < num_points; ++point_index) {
    const size_t ii = indices ? indices->at(point_index) : point_index;
    const Timestamp stamp = io.stamps[ii];
    size_t num_ctrl_pts = search_tree.getLeafCount();  // :164 (leaves, not points)
    while (ctrl_pt_idx < control_points.size() &&
           (control_point_stamps[ctrl_pt_idx] <= stamp + tol || num_ctrl_pts < k + 1)) {  // :167-169
      const bool ctrl_valid = values.exists(ctrl_pt_idx);
      search_tree.addPoint(control_points[ctrl_pt_idx], ctrl_valid);
      ctrl_pt_idx++;
      if (!ctrl_valid) continue;
      num_ctrl_pts++;  // :177 (points, not leaves)
    }

    if (search_tree.getLeafCount() < k + 1) {  // :180
      if (num_ctrl_pts > 1) {
        k = num_ctrl_pts - 1;  // :183 — by-value parameter, stays changed
      } else {
        continue;  // :185 — no control_point_map entry, no removal pass
      }
    }

    control_point_map.emplace_back();  // :189
    InterpDetail detail;
    const PointXYZ p_old = getVertex(io, ii);
    const PointXYZ p_new = interpPoint(control_point_map.back(), control_points, values,
                                       search_tree, k, p_old, &detail);
    setVertex(io, ii, p_new);
    record(trace, ii, point_index, detail);

    size_t num_leaves = search_tree.getLeafCount();  // :200
    while (lower_ctrl_pt_idx < control_points.size() && num_leaves > k + 1 &&
           control_point_stamps[lower_ctrl_pt_idx] < stamp - tol) {  // :201-202 (unsigned wrap)
      if (!values.exists(lower_ctrl_pt_idx)) {
        lower_ctrl_pt_idx++;
        continue;
      }
      search_tree.removePoint(lower_ctrl_pt_idx);  // whole 1 m voxel, cpp:40
      num_leaves--;                                // even when the voxel was already gone
      lower_ctrl_pt_idx++;
    }
  }
  return 0;
}

}  // namespace

int deformPoints(const DeformIO& io,
                 std::vector<std::set<size_t>>& control_point_map,
                 const std::vector<Point3d>& control_points,
                 const std::vector<Timestamp>& control_point_stamps,
                 const ValuesView& values,
                 size_t k,
                 double tol_t,
                 const std::vector<size_t>* indices,
                 DeformTrace* trace) {
  // traits::has_get_stamp<CloudIn> picks the overload (mesh_deformation.h:65,131)
  if (io.stamps) {
    return deformPointsStamped(io, control_point_map, control_points, control_point_stamps,
                               values, k, tol_t, indices, trace);
  }
  return deformPointsUnstamped(io, control_point_map, control_points, values, k, indices, trace);
}

// deformation_graph.h:715-746
size_t DeformationGraphMeshPath::findStartIndex(char prefix, int start_index_hint,
                                                const DeformIO& io, double tol_t) const {
  const auto cache = last_calculated_vertices_.find(prefix);
  if (cache == last_calculated_vertices_.end()) return 0;
  if (recalculate_vertices_) return 0;
  if (start_index_hint >= 0) return static_cast<size_t>(start_index_hint);

  // std::max on unsigned operands is a no-op: the subtraction wraps (:735-737)
  const Timestamp min_stamp = std::max(static_cast<Timestamp>(0),
                                       vertex_stamps_.at(prefix).back() - stampFromSec(tol_t));
  // upper_bound over vertex indices with comp(v, i) = v < stamp[i]
  size_t lo = 0, count = io.num_vertices;
  while (count > 0) {
    const size_t step = count / 2;
    const size_t mid = lo + step;
    if (!(min_stamp < io.stamps[mid])) {
      lo = mid + 1;
      count -= step + 1;
    } else {
      count = step;
    }
  }
  return std::min(lo, cache->second.size());
}

// deformation_graph.h:808-870
size_t DeformationGraphMeshPath::deformPoints(const DeformIO& io, char prefix,
                                              const ValuesView& values, size_t k, double tol_t,
                                              const std::vector<int>* graph_indices,
                                              int start_index_hint,
                                              std::vector<std::set<size_t>>* vertex_graph_map,
                                              DeformTrace* trace) {
  if (vertex_positions_.find(prefix) == vertex_positions_.end()) return 0;  // :819-823

  const size_t start_idx = findStartIndex(prefix, start_index_hint, io, tol_t);
  {  // fillPreviousPoints :748-760
    const auto iter = last_calculated_vertices_.find(prefix);
    if (iter != last_calculated_vertices_.end()) {
      for (size_t i = 0; i < start_idx; i++) setVertex(io, i, iter->second.at(i));
    }
  }

  std::vector<size_t> to_deform;
  if (start_idx != 0) {
    if (graph_indices) {
      // predeformPoints :779-806
      for (size_t i = start_idx; i < io.num_vertices; i++) {
        const int index = graph_indices->at(i);
        if (index < 0 || !values.exists(static_cast<size_t>(index))) {
          to_deform.push_back(i);
          continue;
        }
        const PointXYZ v = getVertex(io, i);
        const double vi[3] = {static_cast<double>(v.x), static_cast<double>(v.y),
                              static_cast<double>(v.z)};
        const Pose3& T = values.poses[index];
        const Point3d& g = vertex_positions_[prefix].at(index);
        const double p[3] = {vi[0] - g.x, vi[1] - g.y, vi[2] - g.z};
        PointXYZ out;
        float* o[3] = {&out.x, &out.y, &out.z};
        for (int r = 0; r < 3; ++r) {
          const double rot =
              (T.R[3 * r + 0] * p[0] + T.R[3 * r + 1] * p[1]) + T.R[3 * r + 2] * p[2];
          *o[r] = static_cast<float>(rot + T.t[r]);
        }
        setVertex(io, i, out);
      }
    } else {
      to_deform.resize(io.num_vertices - start_idx);
      std::iota(to_deform.begin(), to_deform.end(), start_idx);
    }
  }

  const std::vector<size_t>* indices_ptr = start_idx == 0 ? nullptr : &to_deform;
  std::vector<std::set<size_t>> vertex_graph_map_deformed;
  pgmo_oracle::deformPoints(io, vertex_graph_map_deformed, vertex_positions_.at(prefix),
                            vertex_stamps_.at(prefix), values, k, tol_t, indices_ptr, trace);

  if (vertex_graph_map) {  // :857-866
    if (start_idx == 0) {
      *vertex_graph_map = vertex_graph_map_deformed;
    } else {
      vertex_graph_map->resize(io.num_vertices);
      for (size_t i = 0; i < indices_ptr->size(); i++) {
        vertex_graph_map->at(indices_ptr->at(i)) = vertex_graph_map_deformed.at(i);
      }
    }
  }

  {  // cacheNewPoints :762-777
    auto& cache = last_calculated_vertices_[prefix];
    cache.resize(io.num_vertices);
    for (size_t i = start_idx; i < io.num_vertices; i++) {
      PointXYZ p;
      p.x = io.out_x[i];
      p.y = io.out_y[i];
      p.z = io.out_z[i];
      cache[i] = p;
    }
  }
  recalculate_vertices_ = false;  // :869
  return start_idx;
}

}  // namespace pgmo_oracle
