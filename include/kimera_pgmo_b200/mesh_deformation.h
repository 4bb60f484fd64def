// Drop-in for kimera_pgmo/include/kimera_pgmo/mesh_deformation.h: the same
// deformation::deformPoints overloads (unstamped :63-75, stamped :129-141), same argument
// meaning and error behaviour, executed by libpgmo_b200.so on a B200.
//
// What the shim does per call (host side, thin):
//   1. gathers the vertices through the ADL traits into SoA float buffers (+ u64 stamps),
//   2. flattens `values` into poses12[n_c][12] (row-major R then t) and valid[n_c] by probing
//      gtsam::Symbol(prefix, j) for j < control_points.size(),
//   3. one C-ABI call (pgmo_deform_points),
//   4. scatters the processed vertices back with traits::set_vertex (colour untouched) and
//      rebuilds std::vector<std::set<size_t>> from the dense association output.
#pragma once

#include <cstdint>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "kimera_pgmo_b200/deps.h"
#include "pgmo_b200.h"

namespace kimera_pgmo_b200 {

// common_structs.cpp:21-24
inline Timestamp stampFromSec(double sec) { return pgmo_stamp_from_sec(sec); }

class PgmoB200Error : public std::runtime_error {
 public:
  PgmoB200Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

namespace detail {
inline void throwOnError(int rc) {
  if (rc >= 0) return;
  const std::string msg = std::string("pgmo_b200: ") + pgmo_last_error();
  if (rc == PGMO_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);  // what .at() would have thrown
  throw PgmoB200Error(rc, msg);
}
}  // namespace detail

namespace deformation {

// One GPU-side control table + window-schedule cache.  Owned by whoever owns the deformation
// graph (one per prefix); not thread-safe, like the reference's DeformationGraph.
class Backend {
 public:
  explicit Backend(int device = 0) {
    pgmo_deform_t* h = nullptr;
    detail::throwOnError(pgmo_deform_create(device, &h));
    handle_.reset(h, [](pgmo_deform_t* p) { pgmo_deform_destroy(p); });
  }
  pgmo_deform_t* get() const { return handle_.get(); }

  // Uploads control positions / stamps and the flattened `values`.
  void setControls(char prefix, const std::vector<gtsam::Point3>& control_points,
                   const std::vector<Timestamp>* control_point_stamps,
                   const gtsam::Values& values) {
    const size_t n_c = control_points.size();
    pos_.resize(3 * n_c);
    valid_.resize(n_c);
    poses_.assign(12 * n_c, 0.0);
    for (size_t j = 0; j < n_c; ++j) {
      pos_[3 * j + 0] = control_points[j].x();
      pos_[3 * j + 1] = control_points[j].y();
      pos_[3 * j + 2] = control_points[j].z();
      const gtsam::Key key = gtsam::Symbol(prefix, j);
      valid_[j] = values.exists(key) ? 1 : 0;
      if (!valid_[j]) continue;
      const gtsam::Pose3& T = values.at<gtsam::Pose3>(key);
      const auto R = T.rotation().matrix();
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) poses_[12 * j + 3 * r + c] = R(r, c);
      }
      poses_[12 * j + 9] = T.translation().x();
      poses_[12 * j + 10] = T.translation().y();
      poses_[12 * j + 11] = T.translation().z();
    }
    const uint64_t* stamps = nullptr;
    if (control_point_stamps && control_point_stamps->size() == n_c) {
      stamps = control_point_stamps->data();
    }
    detail::throwOnError(pgmo_deform_set_controls(handle_.get(), n_c, pos_.data(), stamps,
                                                  valid_.data(), poses_.data()));
  }

 private:
  std::shared_ptr<pgmo_deform_t> handle_;
  std::vector<double> pos_, poses_;
  std::vector<uint8_t> valid_;
};

namespace detail {

// process-wide default backend for the free functions (the reference's are stateless)
inline Backend& defaultBackend() {
  static thread_local Backend backend(0);
  return backend;
}

template <typename CloudOut, typename CloudIn>
void run(Backend& backend, CloudOut& new_points,
         std::vector<std::set<size_t>>& control_point_map, const CloudIn& points, bool stamped,
         size_t k, double tol_t, const std::vector<size_t>* indices, int precision) {
  const size_t n_v = traits::num_vertices(points);
  std::vector<float> x(n_v), y(n_v), z(n_v);
  std::vector<uint64_t> stamps(stamped ? n_v : 0);
  for (size_t i = 0; i < n_v; ++i) {
    const traits::Pos p = traits::get_vertex(points, i);
    x[i] = p.x();
    y[i] = p.y();
    z[i] = p.z();
    if (stamped) stamps[i] = traits::get_timestamp(points, i);
  }
  std::vector<uint64_t> subset;
  if (indices) subset.assign(indices->begin(), indices->end());
  const size_t n_pos = indices ? indices->size() : n_v;
  const uint32_t stride = static_cast<uint32_t>(k + 2 > 8 ? k + 2 : 8);
  std::vector<float> ox(x), oy(y), oz(z);
  std::vector<uint32_t> assoc(n_pos * stride);
  std::vector<uint8_t> cnt(n_pos), processed(n_pos);
  int rc = pgmo_deform_points(backend.get(), n_v, x.data(), y.data(), z.data(),
                              stamped ? stamps.data() : nullptr,
                              indices ? subset.data() : nullptr, subset.size(),
                              static_cast<uint32_t>(k), tol_t, precision, ox.data(), oy.data(),
                              oz.data(), nullptr, assoc.data(), stride, cnt.data(),
                              processed.data());
  kimera_pgmo_b200::detail::throwOnError(rc);
  if (rc == PGMO_NOOP) return;  // the reference returned before control_point_map.clear()
  control_point_map.clear();
  for (size_t p = 0; p < n_pos; ++p) {
    if (!processed[p]) continue;  // skipped vertices own no entry (mesh_deformation.h:184-186)
    const size_t ii = indices ? (*indices)[p] : p;
    control_point_map.emplace_back();
    // neighbour counts above the stride only occur for degenerate k; the set is what callers use
    const uint32_t n = cnt[p] < stride ? cnt[p] : stride;
    for (uint32_t j = 0; j < n; ++j) control_point_map.back().insert(assoc[p * stride + j]);
    traits::set_vertex(new_points, ii, traits::Pos(ox[ii], oy[ii], oz[ii]));
  }
}

}  // namespace detail

// mesh_deformation.h:63-113 — clouds without stamps
template <typename CloudOut, typename CloudIn,
          std::enable_if_t<!traits::has_get_stamp<CloudIn>::value, bool> = true>
void deformPoints(CloudOut& new_points, std::vector<std::set<size_t>>& control_point_map,
                  const CloudIn& points, char prefix,
                  const std::vector<gtsam::Point3>& control_points,
                  const std::vector<Timestamp>& /* control_point_stamps */,
                  const gtsam::Values& values, size_t k = 4, double /* tol_t */ = 10.0,
                  const std::vector<size_t>* indices = nullptr, Backend* backend = nullptr,
                  int precision = PGMO_PRECISION_FP64) {
  const size_t num_points = indices ? indices->size() : traits::num_vertices(points);
  if (!num_points) return;
  if (control_points.size() == 0) return;  // "No control points. No deformation"
  Backend& b = backend ? *backend : detail::defaultBackend();
  b.setControls(prefix, control_points, nullptr, values);
  detail::run(b, new_points, control_point_map, points, false, k, 0.0, indices, precision);
}

// mesh_deformation.h:129-213 — stamped clouds
template <typename CloudOut, typename CloudIn,
          std::enable_if_t<traits::has_get_stamp<CloudIn>::value, bool> = true>
void deformPoints(CloudOut& new_points, std::vector<std::set<size_t>>& control_point_map,
                  const CloudIn& points, char prefix,
                  const std::vector<gtsam::Point3>& control_points,
                  const std::vector<Timestamp>& control_point_stamps,
                  const gtsam::Values& values, size_t k = 4, double tol_t = 10.0,
                  const std::vector<size_t>* indices = nullptr, Backend* backend = nullptr,
                  int precision = PGMO_PRECISION_FP64) {
  const size_t num_points = indices ? indices->size() : traits::num_vertices(points);
  if (!num_points) return;
  if (control_points.size() < k) return;  // "Not enough valid control points to deform points"
  if (control_point_stamps.size() < control_points.size()) {
    throw std::out_of_range("control_point_stamps shorter than control_points");
  }
  Backend& b = backend ? *backend : detail::defaultBackend();
  b.setControls(prefix, control_points, &control_point_stamps, values);
  detail::run(b, new_points, control_point_map, points, true, k, tol_t, indices, precision);
}

}  // namespace deformation
}  // namespace kimera_pgmo_b200
